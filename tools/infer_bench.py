"""Inference throughput (default cfg5: N=784, D=32, B images; args: B N D): MPS.predict on raw pixels with the fused cos/sin feature map."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, ROOT)
import numpy as np, torch
import trmps_native as nat, trmps_device as dev
from bench import full_bond_mps
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 784
D = int(sys.argv[3]) if len(sys.argv) > 3 else 32
L = 10
loc = N // 2
nodes = full_bond_mps(N, 2, L, D, loc)
w = dev.DeviceWeights(nodes, loc, L, dev.bond_stride(nodes, loc, L), torch.device("cuda", 0))
g = torch.Generator(device="cuda").manual_seed(0)
pix = torch.rand((N, B), device="cuda", generator=g)
out = torch.empty((B, L), device="cuda")
for _ in range(2):
    dev.predict(w, pix, nat.FM_COS_SIN, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.predict(w, pix, nat.FM_COS_SIN, out=out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = min(ts)
flops = B * ((N - 1) * 2.0 * 2 * D * D + 2.0 * L * 2 * D * D)
print("predict N=%d D=%d B=%d: %.2f ms  -> %.2f M images/s, %.1f TFLOP/s algorithmic, pixel stream %.0f GB/s; logits mean %.4f" %
      (N, D, B, ms, B / ms / 1e3, flops / ms / 1e9, 4.0 * N * B / ms / 1e6, float(out.mean())))
import ctypes as C
buf = (C.c_float * 64)()
nat.load().trmps_debug_chain(buf)
print("chain epilogue (CTA0 tile0) cycles: other/pixel %.0f  wait x_ready %.0f  tmem ld + combine %.0f  store_state+arrive %.0f  (last launch, %d steps)" % (buf[0], buf[1], buf[2], buf[3], N - 1 - loc))
print("chain issuer cycles: wait node %.0f  wait a_ready %.0f  issue+commit %.0f" % (buf[4], buf[5], buf[6]))
t0 = buf[16]
print("events of the middle step (cycles after the issuer entered the step):")
print("  issuer: issued (tile,half) " + "  ".join("T%dh%d %+.0f" % (i % 2, i // 2, buf[17 + i] - t0) for i in range(4)))
for k in range(2):
    print("  tile %d epilogue: x_ready h0 %+.0f  h0 converted %+.0f  x_ready h1 %+.0f  published %+.0f" % (k, buf[32 + 8 * k] - t0, buf[32 + 8 * k + 3] - t0, buf[32 + 8 * k + 1] - t0, buf[32 + 8 * k + 4] - t0))
