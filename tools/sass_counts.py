"""SASS opcode counts per kernel of the built library (evidence that the tcgen05 / TMEM / bulk-copy / cluster paths are real).
usage: sass_counts.py [path/to/libtrmps_b200.so]  ->  table on stdout (kept as profiles/*_sass_opcode_counts.txt)"""
import collections, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mps-mnist_b200", "lib", "libtrmps_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "SYNCS", "UCGABAR_", "DFMA", "FFMA2", "FFMA", "SHFL", "MUFU", "BAR.SYNC", "LDS", "STS", "LDG", "STG"]
name, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").replace("trmps::", "")
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter(); total[name] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1); total[name] += 1
        for c in COLS:
            if op.startswith(c):
                if c == "FFMA" and op.startswith("FFMA2"):
                    continue
                counts[name][c] += 1
print("# cuobjdump -sass %s  (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -lineinfo)" % os.path.relpath(lib, ROOT))
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,")
print("# UCGABAR_* = barrier.cluster, DFMA = FP64 FMA, FFMA2 = packed FP32 FMA.  No UTMALDG: operands move with 1-D cp.async.bulk, not tensor-map TMA.")
print("%-58s %7s " % ("kernel", "instrs") + " ".join("%8s" % c for c in COLS))
for n in sorted(counts, key=lambda k: -total[k]):
    print("%-58s %7d " % (n[:58], total[n]) + " ".join("%8d" % counts[n][c] for c in COLS))
