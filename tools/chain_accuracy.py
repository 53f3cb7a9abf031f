"""Relative error of MPS.predict against the fp64 oracle on a sample of images at full chain length
(cfg5: N=784, D=32; cfg3: N=196, D=120; plus N=784, D=64/120), fused tcgen05 chain vs the per-site FP32 SIMT path."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
L = 10
for N, D, B in ((784, 32, 4096), (196, 120, 2048), (784, 64, 2048), (784, 120, 1024)):
    loc = N // 2
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=5)
    w = dev.DeviceWeights(nodes, loc, L, dev.bond_stride(nodes, loc, L), torch.device("cuda", 0))
    pix = torch.rand((N, B), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    sub = pix[:, :128].contiguous()
    phi64 = o.feature_map(sub.cpu().numpy().astype(np.float64), o.FM_COS_SIN, dtype=np.float64)
    want = o.predict([n.astype(np.float64) for n in nodes], loc, phi64, L)
    line = "N=%d D=%d:" % (N, D)
    for name, impl in (("default", nat.CHAIN_TCGEN05), ("tcgen05 fp16-split chain (forced)", nat.CHAIN_TCGEN05_ALWAYS), ("per-site FP32 SIMT", nat.CHAIN_SIMT)):
        nat.set_option(nat.OPT_CHAIN_IMPL, impl)
        got = dev.predict(w, pix, nat.FM_COS_SIN)[:128].cpu().numpy().astype(np.float64)
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        scale = float(np.sum(got * want) / np.sum(want * want))      # common factor (1 = unbiased)
        line += "  %s: rel err %.2e (common factor 1%+.2e)" % (name, err, scale - 1)
    nat.set_option(nat.OPT_CHAIN_IMPL, nat.CHAIN_TCGEN05)
    print(line)
