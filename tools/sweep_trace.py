"""Per-bond trace (cost0, cost1, Armijo trials, accepted step, counters) of the first bonds of the cfg3 benchmark sweep."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, ROOT)
import numpy as np, torch
import trmps_native as nat, trmps_device as dev
from bench import synthetic_pixels, full_bond_mps
N, B, L, D, loc = 24, 60000, 10, 120, 0
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5
pix, lab = synthetic_pixels(N, B, L, seed=100)
nodes = full_bond_mps(N, 2, L, D, loc)
st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(pix, lab)
opt = dev.make_opt(lr=lr, max_size=D, min_singular_value=0.0)
st.init_env(loc)
for c in range(N - 1):
    st.bond_step(c, +1, opt)
    iw, sc = st.iwork.cpu().numpy(), st.scalars().cpu().numpy()
    print("bond %2d cost0 %.6f cost1 %.6f gdc %.3e trials %2d accept %d lr* %.2e m %d svd_sweeps %d | exhaust %d revert %d" %
          (c, sc[0], sc[3], sc[1], iw[1], iw[0], sc[2], iw[2], iw[3], iw[5], iw[4]))
