"""Event-timed datasource kernels (csrc/preprocess.cu): GB/s of algorithmic bytes (images read + batch written) per launch.
`--once` launches each case once (for ncu)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps"))
import torch
import trmps_native as nat, trmps_device as dev
once = "--once" in sys.argv
nat.require_gpu(0)
d = torch.device("cuda", 0)
CASES = [("pool f32 60k", 60000, torch.float32, True, nat.FM_PRECOMPUTED),
         ("pool u8  60k", 60000, torch.uint8, True, nat.FM_PRECOMPUTED),
         ("pool f32 60k -> phi", 60000, torch.float32, True, nat.FM_ONE_SQRT3),
         ("pool f32 1M", 1000000, torch.float32, True, nat.FM_PRECOMPUTED),
         ("flat f32 1M", 1000000, torch.float32, False, nat.FM_PRECOMPUTED),
         ("flat u8  1M", 1000000, torch.uint8, False, nat.FM_PRECOMPUTED)]
for name, B, dt, shrink, fm in CASES:
    imgs = torch.rand((B, 784), device=d) if dt == torch.float32 else torch.randint(0, 256, (B, 784), device=d, dtype=torch.uint8)
    n = 196 if shrink else 784
    out = torch.empty((n, B) if fm == nat.FM_PRECOMPUTED else (n, B, 2), device=d)
    nbytes = imgs.numel() * imgs.element_size() + out.numel() * 4
    if once:
        dev.preprocess_batch(imgs, 28, 28, shrink, 0, B, out_map=fm, out=out); torch.cuda.synchronize(); continue
    for _ in range(3): dev.preprocess_batch(imgs, 28, 28, shrink, 0, B, out_map=fm, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for r in range(reps): dev.preprocess_batch(imgs, 28, 28, shrink, (r * 977) % B, B, out_map=fm, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-22s %8.3f ms  %7.1f MB  %7.1f GB/s  %6.1f M images/s" % (name, ms, nbytes / 1e6, nbytes / ms / 1e6, B / ms / 1e3))
    del imgs, out
# phi dataset gather (total, N, 2) -> (N, B, 2)
for name, total, N in (("gather phi 60k x196", 60000, 196), ("gather phi 250k x784", 250000, 784)):
    data = torch.rand((total, N, 2), device=d); lab = torch.rand((total, 10), device=d)
    fo = torch.empty((N, total, 2), device=d); lo = torch.empty((total, 10), device=d)
    nbytes = 2 * data.numel() * 4
    if once:
        dev.batch_gather(data, lab, 5, total, fo, lo); torch.cuda.synchronize(); continue
    for _ in range(3): dev.batch_gather(data, lab, 5, total, fo, lo)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(20): dev.batch_gather(data, None, r * 31, total, fo, None)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("%-22s %8.3f ms  %7.1f MB  %7.1f GB/s" % (name, ms, nbytes / 1e6, nbytes / ms / 1e6))
    del data, fo
