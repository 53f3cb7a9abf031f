"""Per-bond Jacobi convergence history (rotating CTA-rounds and worst |cos| per sweep) and spectrum summary in the cfg3 sweep."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, ROOT)
import numpy as np, torch
import trmps_native as nat, trmps_device as dev
from bench import synthetic_pixels, full_bond_mps
N, B, L, D, loc = 12, 20000, 10, int(sys.argv[2]) if len(sys.argv) > 2 else 120, 0
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-4
pix, lab = synthetic_pixels(N, B, L, seed=100)
nodes = full_bond_mps(N, 2, L, D, loc)
st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(pix, lab)
opt = dev.make_opt(lr=lr, max_size=D, min_singular_value=0.0)
st.init_env(loc)
for c in range(N - 1):
    st.bond_step(c, +1, opt)
    iw = st.iwork.cpu().numpy()
    ns = int(iw[nat.I_SVD_SWEEPS])
    sv = np.sort(st.singular_values().cpu().numpy())[::-1]
    nrot = iw[8: 8 + ns]
    worst = iw[8 + 32: 8 + 32 + ns].view(np.float32)
    print("bond %2d accept %d sweeps %2d  sigma[0,60,119,120,180,239] = %s" % (c, iw[0], ns, ["%.2e" % sv[min(i, len(sv) - 1)] for i in (0, D // 2, D - 1, D, 3 * D // 2, 2 * D - 1)]))
    print("    nrot  ", list(nrot)); print("    worst ", ["%.1e" % w for w in worst])
    R = 2 * st.Ds; nbk = (R + 7) // 8
    off = st.layout.red + 32 + (nbk + 1) * 8
    wt = st.work[off: off + ns].cpu().numpy().view(np.float32)
    print("    w_top ", ["%.1e" % w for w in wt])
