"""Generates tests/golden/*.npz from the oracle (and, when available, from the reference's own Python executed
under oracle/tf1_shim.py -- see make_golden_from_reference.py).  Run from the repo root:

    python oracle/make_golden.py

The vectors are small seeded cases of the hot path: logits of the initial MPS, then per-bond-update records of
one full training step (cost before/after, Armijo trials, kept bond dimension, leading singular values) and the
logits of the updated MPS.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import trmps_oracle as o

OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = {
    # name: (N, B, L, max_size, loc, lr, loss, updates_per_step, seed)
    "xent_n8": (8, 96, 10, 24, None, 0.05, o.SOFTMAX_XENT, 1, 7),
    "xent_n10_loc0": (10, 64, 10, 32, 0, 0.05, o.SOFTMAX_XENT, 1, 8),
    "squared_n6": (6, 80, 10, 24, 2, 0.02, o.SQUARED, 2, 9),
}


def run_case(N, B, L, max_size, loc, lr, loss, updates, seed, dtype=np.float32):
    pix, lab = o.synthetic_batch(N, B, L, seed=seed)
    phi = o.feature_map(pix, o.FM_ONE_SQRT3)
    nodes, loc = o.initial_nodes(N, 2, L, loc=loc)
    nodes = [n.astype(dtype) for n in nodes]
    phi_t, lab_t = phi.astype(dtype), lab.astype(dtype)
    f_init = o.predict(nodes, loc, phi_t, L)
    p = o.SweepParams(max_size=max_size, updates_per_step=updates, loss_kind=loss)
    tr = o.Trace()
    new_nodes = o.train_step(nodes, loc, phi_t, lab_t, lr, p, L, tr)
    f_final = o.predict(new_nodes, loc, phi_t, L)
    rec = dict(pixels=pix, labels=lab, loc=np.int32(loc), f_init=f_init, f_final=f_final,
               cost0=np.array([t['cost0'] for t in tr]), cost1=np.array([t['cost1'] for t in tr]),
               trials=np.array([t['trials'] for t in tr]), accepted=np.array([t['accepted'] for t in tr]),
               gdc=np.array([t['gdc'] for t in tr]), grad_norm=np.array([t['grad_norm'] for t in tr]),
               m=np.array([t['m'] for t in tr if 'm' in t]),
               s_top=np.array([t['s'][:6] for t in tr if 's' in t]),
               dims=np.array([w.shape[-1] for w in new_nodes]))
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, cfg in CASES.items():
        rec = run_case(*cfg)
        rec64 = run_case(*cfg, dtype=np.float64)
        rec['f_final_f64'] = rec64['f_final']
        rec['cfg'] = np.array([x if x is not None else -1 for x in cfg], dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "updates", len(rec['cost0']), "final cost", rec['cost1'][-1], "dims", rec['dims'])


if __name__ == "__main__":
    main()
