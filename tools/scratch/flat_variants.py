import os, sys
sys.path.insert(0, "mps-mnist_b200/trmps")
import torch
import trmps_native as nat, trmps_device as dev
d = torch.device("cuda", 0)
B = 1000000
imgs = torch.rand((B, 784), device=d)
out = torch.empty((784, B), device=d)
nbytes = 2 * imgs.numel() * 4
for v in (0, 1, 2, 3):
    nat.set_option(3, v)
    out.zero_()
    dev.preprocess_batch(imgs, 28, 28, False, 0, B, out=out)
    ok = torch.equal(out, imgs.t())
    small = dev.preprocess_batch(imgs[:1003].contiguous(), 28, 28, False, 7, 517)
    ok2 = torch.equal(small, imgs[:1003][(7 + torch.arange(517, device=d)) % 1003].t())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(20): dev.preprocess_batch(imgs, 28, 28, False, (r * 977) % B, B, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("variant %d  correct %s %s  %8.3f ms  %7.1f GB/s" % (v, ok, ok2, ms, nbytes / ms / 1e6))
a.record()
for r in range(20): out.copy_(imgs.t())
b.record(); torch.cuda.synchronize()
print("torch transpose copy %8.3f ms %7.1f GB/s" % (a.elapsed_time(b) / 20, nbytes / (a.elapsed_time(b) / 20) / 1e6))
