"""Per-piece CUDA-event timings of one bond update at a given size (development aid, not the bench)."""
import argparse
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "mps-mnist_b200", "trmps"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch

import trmps_native as nat
import trmps_device as dev
import trmps_oracle as o


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=12)
    ap.add_argument("--B", type=int, default=60000)
    ap.add_argument("--D", type=int, default=120)
    ap.add_argument("--L", type=int, default=10)
    a = ap.parse_args()
    N, B, D, L = a.N, a.B, a.D, a.L
    pix, lab = o.synthetic_batch(N, B, L, seed=0)
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=0, seed=1)
    loc = 3
    nodes = o.random_full_bond_mps(N, 2, L, D, loc=loc, seed=1)
    st = dev.SweepState(N, B, L, dev.bond_stride(nodes, loc, L, D), nat.FM_COS_SIN, nat.LOSS_SOFTMAX_XENT, torch.device("cuda", 0))
    st.set_nodes(nodes, loc)
    st.set_batch(pix, lab)
    opt = dev.make_opt(lr=1e-3, max_size=D, min_singular_value=0.0)
    print("layout gsplit=%d nsplit=%d svd_block=%d Ds=%d" % (st.layout.gsplit, st.layout.nsplit, st.layout.svd_block, st.Ds))
    t = timed(lambda: st.init_env(loc)); print("init_env (N-2=%d env steps)   min %.3f ms  med %.3f ms" % (N - 2, *t))
    flops = 2.0 * B * L * 4 * D * D
    t = timed(lambda: st.bond_merge(loc, +1)); print("merge                         min %.3f ms  med %.3f ms" % t)
    t = timed(lambda: st.bond_forward(loc, +1, 0)); print("forward                       min %.3f ms  med %.3f ms   %.1f TFLOP/s" % (*t, flops / t[0] / 1e9))
    t = timed(lambda: st.bond_gradient(loc, +1, opt)); print("forward+loss+gradient         min %.3f ms  med %.3f ms" % t)
    if os.environ.get("TRMPS_GRAD_DEBUG"):
        ph = st.work[st.layout.hres:st.layout.hres + 6].cpu().numpy()
        print("  grad_tc CTA0 warp1 cycles (loads+prev mma issue, wait empty, convert+store, fence, sync, -):", [int(x) for x in ph], "sum", int(ph.sum()))
    st.bond_merge(loc, +1)
    t = timed(lambda: (st.bond_merge(loc, +1), st.bond_split(loc, +1, opt), st.set_nodes(nodes, loc))[0], reps=3, warm=1)
    iw = st.iwork.cpu().numpy()
    print("merge+split(+reset)           min %.3f ms  med %.3f ms  sweeps=%d m=%d" % (*t, iw[nat.I_SVD_SWEEPS], iw[nat.I_M]))
    ns = int(iw[nat.I_SVD_SWEEPS])
    print("  CTAs that rotated per sweep:", iw[8:8 + ns].tolist())
    ph = st.work[st.layout.red:st.layout.red + 8].cpu().numpy()
    print("  CTA0 phase cycles (load, gram, rotate, replay, store+publish, gridsync, flagwait, -):", [int(x) for x in ph], " total %.2f ms @1.9GHz" % (ph.sum() / 1.9e6))
    print("  max rel off-diagonal per sweep:", ["%.1e" % x for x in iw[40:40 + ns].view(np.float32).tolist()])

    def full():
        st.set_nodes(nodes, loc)
        st.bond_step(loc, +1, opt)
    t = timed(full, reps=3, warm=1); print("bond_step (+node reset)       min %.3f ms  med %.3f ms" % t)
    iw = st.iwork.cpu().numpy(); sc = st.scalars().cpu().numpy()
    print("accept=%d trials=%d m=%d svd_sweeps=%d cost0=%.5f cost1=%.5f lr*=%.3g" %
          (iw[nat.I_ACCEPT], iw[nat.I_TRIALS], iw[nat.I_M], iw[nat.I_SVD_SWEEPS], sc[0], sc[3], sc[2]))


if __name__ == "__main__":
    main()
