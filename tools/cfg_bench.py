"""Full sweeps (reference schedule loc -> N-1 -> 0 -> loc, baseoptimizer.py:219-236) at the other BASELINE.json
configurations: cfg2 (N=784, D=20, 20k images) and cfg4 (N=784, D=64, 60k images, two sweeps with cached environments).
Prints images*bonds/s and the per-category device-time breakdown.  usage: cfg_bench.py N D B sweeps [span] [single]
`single` as sixth argument runs the single-site optimizer (singlesiteOptimizer.py) on the same schedule.
span limits every sweep to the first `span` bonds (out and back): the random full-bond MPS of the benchmark is not
norm-neutral (the special node gains a factor ~1.12 per bond), so a walk over all 783 bonds of a 28x28 chain overflows
FP32 -- as it would in the reference with this initial state; 195 bonds (cfg3's length) stay finite."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, ROOT)
import numpy as np, torch
import trmps_native as nat, trmps_device as dev
from bench import synthetic_pixels, full_bond_mps
N, D, B, sweeps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 1
span = min(int(sys.argv[5]), N - 1) if len(sys.argv) > 5 else N - 1
single = len(sys.argv) > 6 and sys.argv[6] == "single"
L, loc = 10, 0
pix, lab = synthetic_pixels(N, B, L, seed=100)
nodes = full_bond_mps(N, 2, L, D, loc)
st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(pix, lab)
opt = dev.make_opt(lr=1e-4, max_size=D, min_singular_value=0.0, reg=1e-3, single_site=single)
def run():
    st.set_nodes(nodes, loc)
    st.init_env(loc)
    for _ in range(sweeps):
        st.sweep(loc, span, +1, opt); st.sweep(span, 0, -1, opt); st.sweep(0, loc, +1, opt)
run(); torch.cuda.synchronize()
nat.profile_begin(400000)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
prof = nat.profile_end()
ms = a.elapsed_time(b)
bonds = sweeps * 2 * span
iw = st.iwork.cpu().numpy()
print(("single-site " if single else "") + "N=%d D=%d B=%d, %d sweep(s) out and back over the first %d bonds = %d bond updates (+ environment build over all %d sites): %.1f ms -> %.2f M images*bonds/s (%.3f ms per update); m=%d, Jacobi sweeps of the last split %d" %
      (N, D, B, sweeps, span, bonds, N, ms, B * bonds / ms / 1e3, ms / bonds, iw[nat.I_M], iw[nat.I_SVD_SWEEPS]))
print("   breakdown ms:", {k: round(v[0], 1) for k, v in prof.items()})
sc = st.scalars().cpu().numpy(); sv = st.singular_values().cpu().numpy()
print("   last update: cost0 %.6g cost1 %.6g; sigma max %.4g min %.4g finite=%s; updates %d reverts %d armijo-exhausted %d" %
      (sc[0], sc[3], np.nanmax(sv), np.nanmin(sv), bool(np.isfinite(sv).all()), iw[nat.I_N_UPDATES], iw[nat.I_N_REVERT], iw[nat.I_N_EXHAUST]))
