"""Minimal program for ncu: a few merge+split calls at cfg3 bond size."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
N, B, L, D, loc = 6, 4096, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 120, 2
pix, lab = o.synthetic_batch(N, B, L, seed=0)
nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=1)
st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(pix, lab)
opt = dev.make_opt(lr=1e-3, max_size=D, min_singular_value=0.0)
for _ in range(2):
    st.set_nodes(nodes, loc); st.bond_merge(loc, +1); st.bond_split(loc, +1, opt)
torch.cuda.synchronize()
print("sweeps", int(st.iwork.cpu()[nat.I_SVD_SWEEPS]))
