"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: launch_list_summary.py launches.csv"""
import collections, csv, re, sys
tot, cnt = collections.Counter(), collections.Counter()
rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rows[1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu].strip().replace("second", "s").replace("usecond", "us").replace("nsecond", "ns"), 1e-6)
    name = re.sub(r"\(.*", "", r[kn])[:72]
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print("%-72s %9s %12s %8s" % ("kernel", "launches", "total ms", "share"))
for name, v in tot.most_common():
    print("%-72s %9d %12.3f %7.1f%%" % (name, cnt[name], v, 100 * v / total))
print("%-72s %9d %12.3f" % ("all listed launches", sum(cnt.values()), total))
