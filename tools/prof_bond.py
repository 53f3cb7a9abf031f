"""Minimal program for ncu: a short chain (N sites) at cfg3 bond size, one L->R half-sweep of bond updates."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, ROOT)
import numpy as np, torch
import trmps_native as nat, trmps_device as dev
from bench import synthetic_pixels, full_bond_mps
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6
loc = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # start in the middle of a longer chain for full-size bonds
B, L, D = 60000, 10, 120
pix, lab = synthetic_pixels(N, B, L, seed=100)
nodes = full_bond_mps(N, 2, L, D, loc)
st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(pix, lab)
opt = dev.make_opt(lr=1e-3, max_size=D, min_singular_value=0.0)
for rep in range(2):
    st.set_nodes(nodes, loc)
    st.init_env(loc)
    st.sweep(loc, min(N - 1, loc + 3), +1, opt)
torch.cuda.synchronize()
iw = st.iwork.cpu().numpy()
print("updates", iw[nat.I_N_UPDATES], "reverts", iw[nat.I_N_REVERT], "m", iw[nat.I_M], "svd sweeps", iw[nat.I_SVD_SWEEPS])
