"""Event-timed forward contraction (merged two-site bond, trmps_bond_forward) at cfg3 bond size: one CTA per image tile vs CTA
pairs (tcgen05 cta_group::2, trmps_set_option(6, v)), for a few batch sizes (wave quantisation / column-split tail)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import trmps_native as nat, trmps_device as dev, trmps_oracle as o
N, L, D, loc = 8, 10, 120, 3
def t_of(fn, reps=9):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for B in [int(v) for v in (sys.argv[1:] or ["60000", "56832", "7500"])]:
    pix, lab = o.synthetic_batch(N, B, L, seed=0)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=1)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc); st.set_batch(pix, lab)
    st.init_env(loc); st.bond_merge(loc, +1)
    res = []
    for pair in (0, 1):
        nat.set_option(6, pair)
        ms = t_of(lambda: st.bond_forward(loc, +1, 0))
        f = st.work[st.layout.f0: st.layout.f0 + B * L].clone()
        res.append((ms, f))
    nat.set_option(6, 1)
    d = float((res[0][1] - res[1][1]).abs().max() / res[0][1].abs().max())
    print("B=%6d: single-CTA %.3f ms, CTA pairs %.3f ms (phi + absmax + prep + forward + logits copy), max rel diff %.1e" % (B, res[0][0], res[1][0], d))
    del st; torch.cuda.empty_cache()
