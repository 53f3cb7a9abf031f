"""Bond-by-bond comparison of the CUDA sweep with the oracle trace (debug aid)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o

N, B, L, max_size, loc, lr = 10, 257, 10, 32, 0, 0.05
if len(sys.argv) > 1:
    N, B, max_size, loc = (int(x) for x in sys.argv[1:5])
pix, lab = o.synthetic_batch(N, B, L, seed=51)
phi = o.feature_map(pix, o.FM_ONE_SQRT3)
nodes, loc = o.initial_nodes(N, 2, L, loc=loc)
p = o.SweepParams(max_size=max_size)
tr = o.Trace()
want = o.train_step([n.astype(np.float64) for n in nodes], loc, phi.astype(np.float64), lab.astype(np.float64), lr, p, L, tr)
Ds = dev.bond_stride(nodes, loc, L, max_size)
st = dev.SweepState(N, B, L, Ds, nat.FM_PRECOMPUTED, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(phi, lab)
opt = dev.make_opt(lr=lr, max_size=max_size)
st.init_env(loc)
sched = [(c, +1) for c in range(loc, N - 1)] + [(c, -1) for c in range(N - 1, 0, -1)] + [(c, +1) for c in range(0, loc)]
for i, (c, d) in enumerate(sched):
    st.bond_step(c, d, opt)
    torch.cuda.synchronize()
    iw, sc = st.iwork.cpu().numpy(), st.scalars().cpu().numpy()
    t = tr[i]
    sv = st.singular_values().cpu().numpy()[:len(t['s'])]
    k = min(8, len(sv))
    print("bond %2d site %2d dir %+d | cost0 %.6f/%.6f cost1 %.6f/%.6f trials %d/%d m %d/%d sweeps %d | s0 %.5f/%.5f  max|ds|/s0 %.2e  gdc %.4e/%.4e" %
          (i, c, d, sc[0], t['cost0'], sc[3], t['cost1'], iw[1], t['trials'], iw[2], t['m'], iw[3], sv[0], t['s'][0],
           np.max(np.abs(sv - t['s'])) / t['s'][0], sc[1], t['gdc']))
