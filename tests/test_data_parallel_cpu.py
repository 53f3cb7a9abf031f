"""world_size-2 gloo test (CPU) of the data-parallel exchange the bond step performs: every rank
holds a replicated bond and a shard of the images; the un-normalised gradient partial plus the cost sum ride in one
all-reduce, the 19 Armijo trial sums in a second; everything after is replicated.  Checked against the oracle on the
concatenated batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import trmps_oracle as o


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L, D, B = 10, 6, 64
    rng = np.random.default_rng(0)                       # same stream on every rank: replicated model, full data
    C1 = rng.normal(size=(B, D)).astype(np.float32)
    C2 = rng.normal(size=(B, D)).astype(np.float32)
    pix = rng.random((2, B), dtype=np.float32)
    phi = o.feature_map(pix, o.FM_COS_SIN)
    lab = np.eye(L, dtype=np.float32)[rng.integers(0, L, B)]
    bond = (0.1 * rng.normal(size=(L, 2, 2, D, D))).astype(np.float32)
    lr, reg, coeff = 0.05, 1e-3, 1e-4
    sl = slice(rank * B // world, (rank + 1) * B // world)
    # ---- exchange 1: gradient partial (no regulariser) + cost sum
    f0 = o.forward_factored(bond, C1[sl], C2[sl], phi[0][sl], phi[1][sl])
    h = o.softmax(f0)
    z = f0 - f0.max(1, keepdims=True)
    cost_sum = float(-(lab[sl] * (z - np.log(np.exp(z).sum(1, keepdims=True)))).sum())
    G = o.gradient_factored(lab[sl] - h, C1[sl], C2[sl], phi[0][sl], phi[1][sl])
    buf = torch.from_numpy(np.concatenate([G.ravel(), [cost_sum]]).astype(np.float32))
    dist.all_reduce(buf)
    G = buf[:-1].numpy().reshape(bond.shape) - np.float32(2 * reg) * bond      # replicated from here
    cost0 = float(buf[-1]) / B
    gdc = float(np.sum(G.astype(np.float64) ** 2)) / B
    # ---- exchange 2: Armijo trial sums from f0 and f(delta)
    fd = o.forward_factored(G, C1[sl], C2[sl], phi[0][sl], phi[1][sl])
    sums = []
    step = lr
    for _ in range(19):
        zz = f0 + np.float32(step) * fd
        zz = zz - zz.max(1, keepdims=True)
        sums.append(float(-(lab[sl] * (zz - np.log(np.exp(zz).sum(1, keepdims=True)))).sum()))
        step *= 0.5
    tb = torch.tensor(sums, dtype=torch.float32)
    dist.all_reduce(tb)
    trial = tb.numpy() / B
    step, chosen = lr, None
    for c in range(19):
        if not trial[c] > cost0 - coeff * step * gdc:
            chosen = (c + 1, step, float(trial[c]))
            break
        step *= 0.5
    q.put((rank, cost0, gdc, chosen, G))
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_rank_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # ranks agree bit for bit (replicated state must not drift)
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2] and res[0][3] == res[1][3]
    assert np.array_equal(res[0][4], res[1][4])
    # and equal the single-process oracle on the full batch
    L, D, B = 10, 6, 64
    rng = np.random.default_rng(0)
    C1 = rng.normal(size=(B, D)).astype(np.float32)
    C2 = rng.normal(size=(B, D)).astype(np.float32)
    pix = rng.random((2, B), dtype=np.float32)
    phi = o.feature_map(pix, o.FM_COS_SIN)
    lab = np.eye(L, dtype=np.float32)[rng.integers(0, L, B)]
    bond = (0.1 * rng.normal(size=(L, 2, 2, D, D))).astype(np.float32)
    p = o.SweepParams(max_size=D, reg=1e-3)
    C = o.calculate_C(C1, C2, phi[0], phi[1])
    tr = o.Trace()
    o.update_bond(bond, C, lab, 0.05, p, tr)
    t = tr[0]
    assert abs(res[0][1] - t['cost0']) < 1e-5 * abs(t['cost0'])
    assert abs(res[0][2] - t['gdc']) < 1e-4 * abs(t['gdc'])
    assert res[0][3][0] == t['trials'] and abs(res[0][3][2] - t['cost1']) < 1e-5 * abs(t['cost1'])
    assert np.linalg.norm(res[0][4] - t['gradient']) < 1e-5 * np.linalg.norm(t['gradient'])
