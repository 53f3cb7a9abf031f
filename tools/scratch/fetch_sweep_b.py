import os, sys
sys.path.insert(0, "mps-mnist_b200/trmps")
import torch
import trmps_native as nat, trmps_device as dev
d = torch.device("cuda", 0)
for shrink in (False, True):
  for B in (60000, 125000, 250000, 500000, 1000000, 2000000):
    imgs = torch.rand((B, 784), device=d)
    n = 196 if shrink else 784
    out = torch.empty((n, B), device=d)
    nbytes = imgs.numel() * 4 + out.numel() * 4
    for _ in range(3): dev.preprocess_batch(imgs, 28, 28, shrink, 0, B, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(20): dev.preprocess_batch(imgs, 28, 28, shrink, 0, B, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    # plain device copy of the same bytes for comparison
    src = imgs.view(-1)[: (nbytes // 8)]
    dst = torch.empty_like(src)
    for _ in range(3): dst.copy_(src)
    a.record()
    for r in range(20): dst.copy_(src)
    b.record(); torch.cuda.synchronize()
    cms = a.elapsed_time(b) / 20
    print("shrink=%d B=%8d  %8.3f ms %7.1f GB/s   | torch copy of the same bytes %8.3f ms %7.1f GB/s" % (shrink, B, ms, nbytes / ms / 1e6, cms, nbytes / cms / 1e6))
    del imgs, out, src, dst
