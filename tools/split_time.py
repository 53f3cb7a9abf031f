"""Event-timed bond_split at a given (L, D) for both split implementations (trmps_opt.split_impl):
0 = Gram + Cholesky + Jacobi on the factor inside one cluster (split_factor.cu), 1 = Jacobi on the bond rows (svd.cu).
usage: split_time.py [D] [L ...]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
D = int(sys.argv[1]) if len(sys.argv) > 1 else 120
for L in [int(v) for v in (sys.argv[2:] or ["10"])]:
    N, B, loc = 6, 1024, 2
    pix, lab = o.synthetic_batch(N, B, L, seed=0)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=1)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc); st.set_batch(pix, lab)
    for kind in ("gauss", "decay1e3"):
        for impl in (0, 1):
            opt = dev.make_opt(lr=1e-3, max_size=D, min_singular_value=0.0, split_impl=impl)
            g = torch.Generator(device="cuda").manual_seed(5)
            ts, res = [], None
            for it in range(5):
                st.set_nodes(nodes, loc); st.bond_merge(loc, +1)
                bm = st.bond_matrix()
                rnd = torch.randn(bm.shape, device="cuda", generator=g) * (bm != 0)     # dense random bond, padding kept zero
                if kind != "gauss":
                    rows = (rnd.abs().sum(1) > 0).nonzero().flatten()
                    scale = torch.ones(bm.shape[0], device="cuda")
                    scale[rows] = torch.logspace(0, -3, len(rows), device="cuda")
                    q, _ = torch.linalg.qr(torch.randn(bm.shape[0], bm.shape[0], device="cuda", generator=g))
                    rnd = ((q * scale) @ q.T @ rnd) * (bm != 0)
                bm.copy_(rnd)
                ref = torch.linalg.svdvals(bm.double()).cpu().numpy()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); st.bond_split(loc, +1, opt); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
                sv = st.singular_values().cpu().numpy()
                m = int(st.iwork.cpu()[nat.I_M])
                res = float(np.max(np.abs(sv[:m] - ref[:m]) / ref[:m]))
            if impl == 0:
                dbg = st.work[st.layout.red: st.layout.red + 3].cpu().numpy(); ns = int(st.iwork.cpu()[nat.I_SVD_SWEEPS])
                cd = st.work[st.layout.red + 8: st.layout.red + 14].cpu().numpy()
                print("   cholesky cycles: G+D load %.0f  update %.0f  reduce %.0f  diagonal block %.0f  solve+write %.0f  norms+rank %.0f" % tuple(cd))
                print("   jacobi CTA 0 warp 0, cycles per sweep: rotations + hand-over %.0f  cluster barrier %.0f" % tuple(dbg[:2] / ns))
            print("impl %d %-9s L=%d D=%d bond %s: split %.3f ms (min of %s), sweeps %d, m %d, sigma rel err %.1e" %
                  (impl, kind, L, D, tuple(bm.shape), min(ts), ["%.3f" % t for t in ts], int(st.iwork.cpu()[nat.I_SVD_SWEEPS]), m, res))
