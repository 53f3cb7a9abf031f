"""fp32 simulation: does re-sorting the rows by norm between sweeps (de Rijk kept up to date) cut the sweep count?
Block structure as in the kernel: blocks of 8 rows, round-robin over blocks, all cross pairs of a block pair per round."""
import numpy as np, sys
rng = np.random.default_rng(1)
R, C = 240, 1200
x = np.arange(R) / (R - 1)
sig = np.exp(np.interp(x, [0, 0.02, 0.25, 0.5, 0.52, 0.75, 1.0], np.log([21, 6, 2.3, 1.3, 0.9, 0.05, 0.01])))
U = np.linalg.qr(rng.standard_normal((R, R)))[0]; V = np.linalg.qr(rng.standard_normal((C, R)))[0]
M0 = ((U * sig) @ V.T).astype(np.float32)

def rot_pairs(M, p, q):
    x, y = M[p], M[q]
    a = (x * x).sum(1); b = (y * y).sum(1); g = (x * y).sum(1)
    den = np.sqrt(a * b); c = np.where(den > 0, np.abs(g) / np.maximum(den, 1e-37), 0)
    zeta = (b - a) / (2 * np.where(g == 0, 1, g)); t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta)); t = np.where(g == 0, 0, t)
    cs = 1 / np.sqrt(1 + t * t); sn = cs * t
    M[p] = (cs[:, None] * x - sn[:, None] * y).astype(np.float32); M[q] = (sn[:, None] * x + cs[:, None] * y).astype(np.float32)
    return float(c.max())

def sweeps(M, resort, maxs=40, tol=1e-3):
    M = M.copy(); n = M.shape[0]; BLK = 8; nb = n // BLK
    order = np.argsort(-(M * M).sum(1)); hist = []
    for sw in range(maxs):
        if resort and sw > 0: order = order[np.argsort(-(M[order] * M[order]).sum(1), kind="stable")]
        worst = 0.0
        blocks = list(range(nb))
        for r in range(nb - 1):
            pairs = [(blocks[i], blocks[nb - 1 - i]) for i in range(nb // 2)]
            # within-block pairs once per sweep (first round), then all cross pairs: 8 steps of 8 disjoint pairs
            for step in range(15 if r == 0 else 8):
                P, Q = [], []
                for (bi, bj) in pairs:
                    rows = list(order[bi * BLK:(bi + 1) * BLK]) + list(order[bj * BLK:(bj + 1) * BLK])
                    if r == 0:
                        ring = list(range(16)); ring = [ring[0]] + [ring[1 + (i + step) % 15] for i in range(15)]
                        for i in range(8): P.append(rows[ring[i]]); Q.append(rows[ring[15 - i]])
                    else:
                        for i in range(8): P.append(rows[i]); Q.append(rows[8 + (i + step) % 8])
                worst = max(worst, rot_pairs(M, np.array(P), np.array(Q)))
            blocks = [blocks[0]] + [blocks[-1]] + blocks[1:-1]
        hist.append(worst)
        if worst < tol: break
    return len(hist), hist
for resort in (False, True):
    n, h = sweeps(M0, resort)
    print("re-sort each sweep" if resort else "initial sort only ", n, ["%.0e" % v for v in h])
