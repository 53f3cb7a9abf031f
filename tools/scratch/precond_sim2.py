"""fp32 simulation: sweeps of one-sided row Jacobi on M, on L (=chol(M M^T)) and on L^T, for a bond-like spectrum."""
import numpy as np, sys
sys.path.insert(0, "tools/scratch")
from precond_sim import sweeps_needed
rng = np.random.default_rng(1)
R, C = 128, 1280
x = np.arange(R) / (R - 1)
sig = np.interp(x, [0, 0.02, 0.25, 0.5, 0.52, 0.75, 1.0], np.log([21, 6, 2.3, 1.3, 0.9, 0.05, 0.01])); sig = np.exp(sig)
U = np.linalg.qr(rng.standard_normal((R, R)))[0]; V = np.linalg.qr(rng.standard_normal((C, R)))[0]
M = ((U * sig) @ V.T).astype(np.float32)
def sort_rows(A):
    return A[np.argsort(-(A * A).sum(1))]
def pivoted_chol(G):
    G = G.astype(np.float64).copy(); n = G.shape[0]; L = np.zeros_like(G); perm = np.arange(n)
    for k in range(n):
        j = k + int(np.argmax(np.diag(G)[k:]))
        G[[k, j]] = G[[j, k]]; G[:, [k, j]] = G[:, [j, k]]; L[[k, j]] = L[[j, k]]; perm[[k, j]] = perm[[j, k]]
        d = np.sqrt(max(G[k, k], 1e-30)); L[k, k] = d
        L[k + 1:, k] = G[k + 1:, k] / d
        G[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], L[k + 1:, k])
    return L.astype(np.float32), perm
G = (M @ M.T).astype(np.float32)
n, r = sweeps_needed(sort_rows(M)); print("rows of M (sorted)      ", n, ["%.0e" % v for v in r])
Ms = sort_rows(M); Gs = (Ms @ Ms.T).astype(np.float32)
L0 = np.linalg.cholesky(Gs.astype(np.float64) + 1e-7 * np.eye(R) * Gs.max()).astype(np.float32)
n, r = sweeps_needed(L0); print("rows of L (norm-sorted) ", n, ["%.0e" % v for v in r])
n, r = sweeps_needed(L0.T.copy()); print("rows of L^T (norm-sorted)", n, ["%.0e" % v for v in r])
Lp, perm = pivoted_chol(G)
n, r = sweeps_needed(Lp.T.copy()); print("rows of L^T (pivoted)   ", n, ["%.0e" % v for v in r])
n, r = sweeps_needed(Lp); print("rows of L (pivoted)     ", n, ["%.0e" % v for v in r])
# two-stage: L^T Jacobi -> W (normalised rows) -> M' = W^T M -> sweeps on M'
def jacobi_rows(A, tol=1e-3, maxs=40):
    A = A.astype(np.float32).copy(); n = A.shape[0]
    for sw in range(maxs):
        worst = 0.0; order = list(range(n))
        for r in range(n - 1):
            p = np.array(order[: n // 2]); q = np.array(order[n // 2:][::-1])
            x, y = A[p], A[q]; a = (x * x).sum(1); b = (y * y).sum(1); g = (x * y).sum(1)
            den = np.sqrt(a * b); c = np.where(den > 0, np.abs(g) / np.maximum(den, 1e-37), 0); worst = max(worst, float(c.max()))
            zeta = (b - a) / (2 * np.where(g == 0, 1, g)); t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta)); t = np.where(g == 0, 0, t)
            cs = 1 / np.sqrt(1 + t * t); sn = cs * t
            A[p] = (cs[:, None] * x - sn[:, None] * y).astype(np.float32); A[q] = (sn[:, None] * x + cs[:, None] * y).astype(np.float32)
            order = [order[0]] + [order[-1]] + order[1:-1]
        if worst < tol: return A, sw + 1
    return A, maxs
Xf, ns = jacobi_rows(Lp.T.copy())
W = Xf / np.linalg.norm(Xf, axis=1, keepdims=True)      # rows = w_p^T  (in pivoted row order)
Mp = (W @ M[perm]).astype(np.float32)
n, r = sweeps_needed(Mp); print("stage 1 sweeps", ns, "| stage 2 on W^T M:", n, ["%.0e" % v for v in r], "| orth err of W", np.abs(W @ W.T - np.eye(R)).max())
