"""Event-timed gradient pass (loss + gradient kernels, no forward) at cfg3 bond size for each implementation."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
N, B, L, D, loc = 8, 60000, 10, 120, 3
pix, lab = o.synthetic_batch(N, B, L, seed=0)
nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=1)
st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
st.set_nodes(nodes, loc); st.set_batch(pix, lab)
opt = dev.make_opt(lr=1e-3, max_size=D, min_singular_value=0.0)
st.init_env(loc); st.bond_merge(loc, +1)
def t_of(fn, reps=7):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
fwd = t_of(lambda: st.bond_forward(loc, +1, 0))
for name, impl in (("3xTF32", nat.GRAD_TCGEN05), ("fp16 hi/lo", nat.GRAD_TCGEN05_F16), ("SIMT", nat.GRAD_SIMT)):
    nat.set_option(nat.OPT_GRAD_IMPL, impl)
    both = t_of(lambda: st.bond_gradient(loc, +1, opt))
    print("%-10s forward+loss+gradient %.3f ms -> loss+gradient %.3f ms" % (name, both, both - fwd))
nat.set_option(nat.OPT_GRAD_IMPL, nat.GRAD_DEFAULT)
