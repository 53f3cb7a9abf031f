"""Offline study: does Gram-eigenvector preconditioning cut the number of one-sided Jacobi sweeps?  (fp32 simulation)"""
import numpy as np
rng = np.random.default_rng(0)

def sweeps_needed(M, tol=1e-3, maxs=40):
    M = M.astype(np.float32).copy(); n = M.shape[0]
    idx = list(range(n)); res = []
    for sw in range(maxs):
        worst = 0.0
        order = list(range(n))
        for r in range(n - 1):
            p = np.array(order[: n // 2]); q = np.array(order[n // 2:][::-1])
            x, y = M[p], M[q]
            a = (x * x).sum(1); b = (y * y).sum(1); g = (x * y).sum(1)
            den = np.sqrt(a * b); c = np.where(den > 0, np.abs(g) / np.maximum(den, 1e-37), 0)
            worst = max(worst, float(c.max()))
            zeta = (b - a) / (2 * np.where(g == 0, 1, g))
            t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta)); t = np.where(g == 0, 0, t)
            cs = 1 / np.sqrt(1 + t * t); sn = cs * t
            M[p] = (cs[:, None] * x - sn[:, None] * y).astype(np.float32)
            M[q] = (sn[:, None] * x + cs[:, None] * y).astype(np.float32)
            order = [order[0]] + [order[-1]] + order[1:-1]
        res.append(worst)
        if worst < tol: break
    return len(res), res

def gram_precond(M):
    G = (M.astype(np.float32) @ M.astype(np.float32).T).astype(np.float32)
    w, U = np.linalg.eigh(G.astype(np.float32))
    return (U.T.astype(np.float32) @ M.astype(np.float32)).astype(np.float32)

R, C = 128, 1280
for name, M in [("gauss", rng.standard_normal((R, C))),
                ("decay", (np.linalg.qr(rng.standard_normal((R, R)))[0] * np.logspace(0, -5, R)) @ np.linalg.qr(rng.standard_normal((C, R)))[0].T),
                ("ident+noise", np.hstack([np.eye(R), np.zeros((R, C - R))]) + 1e-2 * rng.standard_normal((R, C)))]:
    n0, r0 = sweeps_needed(M)
    n1, r1 = sweeps_needed(gram_precond(M))
    print(name, "plain sweeps", n0, ["%.1e" % v for v in r0[-3:]], "| preconditioned", n1, ["%.1e" % v for v in r1])

def gram2_precond(M):
    M = M.astype(np.float32)
    G = (M @ M.T).astype(np.float32)
    G2 = (G @ G.T).astype(np.float32)
    w, U = np.linalg.eigh(G2)
    return (U.T.astype(np.float32) @ M).astype(np.float32)
print("-- eigenvectors taken from G^2 (what one-sided Jacobi on the rows of G resolves)")
for name, M in [("gauss", rng.standard_normal((R, C))),
                ("decay1e2", (np.linalg.qr(rng.standard_normal((R, R)))[0] * np.logspace(0, -2, R)) @ np.linalg.qr(rng.standard_normal((C, R)))[0].T),
                ("decay1e3", (np.linalg.qr(rng.standard_normal((R, R)))[0] * np.logspace(0, -3, R)) @ np.linalg.qr(rng.standard_normal((C, R)))[0].T)]:
    n0, r0 = sweeps_needed(M)
    n1, r1 = sweeps_needed(gram_precond(M))
    n2, r2 = sweeps_needed(gram2_precond(M))
    print(name, "plain", n0, "| G-precond", n1, ["%.1e" % v for v in r1], "| G2-precond", n2, ["%.1e" % v for v in r2])
