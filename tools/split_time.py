"""Event-timed bond_split at a given (L, D): how long does the Jacobi kernel take on an R x (L R) bond?"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
D = int(sys.argv[1]) if len(sys.argv) > 1 else 120
for L in [int(v) for v in (sys.argv[2:] or ["1", "10"])]:
    N, B, loc = 6, 1024, 2
    pix, lab = o.synthetic_batch(N, B, L, seed=0)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=1)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc); st.set_batch(pix, lab)
    opt = dev.make_opt(lr=1e-3, max_size=D, min_singular_value=0.0)
    g = torch.Generator(device="cuda").manual_seed(5)
    ts = []
    for it in range(4):
        st.set_nodes(nodes, loc); st.bond_merge(loc, +1)
        bm = st.bond_matrix()
        bm.copy_(torch.randn(bm.shape, device="cuda", generator=g) * (bm != 0))     # dense random bond, padding kept zero
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.bond_split(loc, +1, opt); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ns = int(st.iwork.cpu()[nat.I_SVD_SWEEPS]); R = 2 * D
    nbk = (R + 7) // 8; rounds = (nbk + (nbk & 1) - 1) * ns
    dbg = st.work[st.layout.red: st.layout.red + 8].cpu().numpy()
    print("   per round (CTA 0, cycles): load %.0f  gram %.0f  rotate %.0f  replay %.0f  store+publish %.0f  grid sync/sweep %.0f  flag wait %.0f   (%d rounds, total %.0f cycles)" %
          (dbg[0] / rounds, dbg[1] / rounds, dbg[2] / rounds, dbg[3] / rounds, dbg[4] / rounds, dbg[5] / ns, dbg[6] / rounds, rounds, dbg[:7].sum()))
    print("L=%d D=%d bond %s: split %.3f ms (min of %s), sweeps %d" % (L, D, tuple(bm.shape), min(ts), ["%.3f" % t for t in ts], int(st.iwork.cpu()[nat.I_SVD_SWEEPS])))
