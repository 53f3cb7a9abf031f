"""Full-size parity run of BASELINE.json configs[0] (cfg1: N=196, d=2, D=10, 10 labels, 1k images, one sweep):
the device train_step against the NumPy oracle (the reference graph restated) on the same inputs and initial MPS.
Prints per-bond cost agreement, accept/reject agreement and the final logits' distance.  args: N B D lr"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
N = int(sys.argv[1]) if len(sys.argv) > 1 else 196
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
L = 10
pix, lab = o.synthetic_batch(N, B, L, seed=7)
phi = o.feature_map(pix, o.FM_ONE_SQRT3)
nodes, loc = o.initial_nodes(N, 2, L)
p = o.SweepParams(max_size=D)
tr = o.Trace()
t0 = time.perf_counter()
want_nodes = o.train_step(nodes, loc, phi, lab, lr, p, L, tr)
t_cpu = time.perf_counter() - t0
want_f = o.predict(want_nodes, loc, phi, L)

Ds = dev.bond_stride(nodes, loc, L, D)
st = dev.SweepState(N, B, L, Ds, nat.FM_PRECOMPUTED, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(phi, lab)
opt = dev.make_opt(lr=lr, max_size=D)
# bond by bond so that the per-bond scalars can be compared with the oracle's trace
st.init_env(loc)
got = []
def run(frm, to, d):
    for c in range(frm, to, d):
        st.bond_step(c, d, opt)
        iw, sc = st.iwork.cpu().numpy(), st.scalars().cpu().numpy()
        got.append((float(sc[0]), float(sc[3]), int(iw[nat.I_ACCEPT]), int(iw[nat.I_M]), int(iw[nat.I_TRIALS])))
torch.cuda.synchronize(); t0 = time.perf_counter()
run(loc, N - 1, +1); run(N - 1, 0, -1); run(0, loc, +1)
torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
got_f = dev.predict(st.weights_view(), st.feat, nat.FM_PRECOMPUTED).cpu().numpy()
rel = lambda a, b: np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30)
print("cfg: N=%d B=%d D=%d lr=%g: %d bond updates; oracle %.1f s on the host, device %.2f s (bond-by-bond with host reads)" % (N, B, D, lr, len(got), t_cpu, t_gpu))
c0 = np.array([[g[0], t["cost0"]] for g, t in zip(got, tr)])
acc = np.array([[g[2], int(t["accepted"])] for g, t in zip(got, tr)])
mm = np.array([[g[3], t["m"]] for g, t in zip(got, tr)])
trl = np.array([[g[4], t["trials"]] for g, t in zip(got, tr)])
d = np.abs(c0[:, 0] - c0[:, 1]) / np.abs(c0[:, 1])
print("cost before each bond update: max rel diff %.2e (bond %d), median %.2e; first/last oracle cost %.6f / %.6f" % (d.max(), int(d.argmax()), np.median(d), c0[0, 1], c0[-1, 1]))
print("accept decisions equal: %d / %d   bond dimensions equal: %d / %d   Armijo trial counts equal: %d / %d" %
      ((acc[:, 0] == acc[:, 1]).sum(), len(acc), (mm[:, 0] == mm[:, 1]).sum(), len(mm), (trl[:, 0] == trl[:, 1]).sum(), len(trl)))
print("final logits: rel diff %.2e, argmax agreement %.4f, final cost device %.6f oracle %.6f" %
      (rel(got_f, want_f), np.mean(got_f.argmax(1) == want_f.argmax(1)), o.cost(got_f, lab), o.cost(want_f, lab)))
first = int(np.argmax(trl[:, 0] != trl[:, 1])) if (trl[:, 0] != trl[:, 1]).any() else -1
print("first Armijo trial-count mismatch at update %d; cost rel diffs of the updates before it: max %.2e" % (first, d[:max(first, 1)].max()))
for i in range(max(0, first - 3), min(len(got), first + 3)):
    print("  update %3d site %s: cost0 dev %.7f oracle %.7f | cost1 dev %.7f oracle %.7f | trials %d / %d | gdc oracle %.3e" %
          (i, tr.sites[i], got[i][0], tr[i]["cost0"], got[i][1], tr[i]["cost1"], got[i][4], tr[i]["trials"], tr[i]["gdc"]))
print("cost rel diff (device vs fp32 oracle) at updates 0,1,2,5,10,20,40,60,80:", ["%.1e" % d[i] for i in (0, 1, 2, 5, 10, 20, 40, 60, 80) if i < len(d)])
# natural sensitivity of the trajectory: the same oracle in float64
tr64 = o.Trace()
o.train_step([n.astype(np.float64) for n in nodes], loc, phi.astype(np.float64), lab.astype(np.float64), lr, p, L, tr64)
c64 = np.array([t["cost0"] for t in tr64])
d64 = np.abs(c0[:, 1] - c64) / np.abs(c64)
dd = np.abs(c0[:, 0] - c64) / np.abs(c64)
print("cost rel diff (fp32 oracle vs fp64 oracle) at the same updates:           ", ["%.1e" % d64[i] for i in (0, 1, 2, 5, 10, 20, 40, 60, 80) if i < len(d64)])
print("cost rel diff (device vs fp64 oracle) at the same updates:                ", ["%.1e" % dd[i] for i in (0, 1, 2, 5, 10, 20, 40, 60, 80) if i < len(dd)])
print("trial counts fp32 oracle vs fp64 oracle equal: %d / %d; device vs fp64 oracle: %d / %d" %
      (sum(int(a["trials"] == b["trials"]) for a, b in zip(tr, tr64)), len(tr), sum(int(g[4] == b["trials"]) for g, b in zip(got, tr64)), len(tr)))
