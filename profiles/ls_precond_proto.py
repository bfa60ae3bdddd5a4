#!/usr/bin/env python
"""CPU prototype (scipy) behind DESIGN.md §8 item 1: the least-squares system of decision 6 on a heightfield, PCG
iterations with Jacobi vs two-level aggregation preconditioners (piecewise-constant and piecewise-linear coarse
spaces, exact coarse solve, damped-Jacobi smoothing).  Measured here (tolerance 1e-6):
    n = 200 (40 k unknowns): Jacobi 142; 4x4 aggregates: constant 45, linear 23; 8x8: constant 56, linear 39
    n = 400 (161 k unknowns): Jacobi 300; 8x8 aggregates: constant 97, linear 57
Jacobi grows linearly with n (config 5, n = 3163: 2496 on the GPU); unsmoothed aggregation with a fixed aggregate
size still grows with n (the piecewise coarse functions carry large gradient-jump energy), so a production
preconditioner needs smoothed prolongation and several levels — not built.
usage: ls_precond_proto.py <n> [aggregate size]"""
import sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
sys.path.insert(0, '/root/repo')
from optix_prime_baking_b200 import scenes

def build_system(n=200, seed=3, frac_sampled=0.5, w=0.1, warp=0.6):
    m = scenes._heightfield(n, seed, 40.0*n/3163, 3.0*n/3163*8, 24.0*n/3163, warp) if False else scenes._heightfield(n, seed, 40.0, 3.0, 24.0, warp)
    V = m.vertices.astype(np.float64); T = m.tris.astype(np.int64)
    nV, nT = len(V), len(T)
    rng = np.random.default_rng(0)
    # samples: one per triangle for the first half of the triangles (like the leftover rule), random bary
    ns = int(frac_sampled * nT)
    tri = np.arange(ns)
    r1, r2 = rng.random(ns), rng.random(ns)
    s = np.sqrt(r1); bary = np.stack([1 - s, r2 * s, 1 - (1 - s) - r2 * s], axis=1)
    e0 = V[T[:, 1]] - V[T[:, 0]]; e1 = V[T[:, 2]] - V[T[:, 0]]
    area = 0.5 * np.linalg.norm(np.cross(e0, e1), axis=1)
    dA = area[tri]
    ao = rng.uniform(0.2, 1.0, ns)
    rows = np.repeat(T[tri], 3, axis=1).ravel(); cols = np.tile(T[tri], (1, 3)).ravel()
    vals = (dA[:, None, None] * bary[:, :, None] * bary[:, None, :]).ravel()
    M = sp.coo_matrix((vals, (rows, cols)), shape=(nV, nV)).tocsr()
    b = np.zeros(nV); np.add.at(b, T[tri].ravel(), (dA[:, None] * ao[:, None] * bary).ravel())
    d = M.diagonal(); fixed = ~(d > 0)
    M = M + sp.diags(fixed.astype(float)); b[fixed] = 0
    # interior edges
    he = np.concatenate([T[:, [0, 1, 2]], T[:, [1, 2, 0]], T[:, [2, 0, 1]]])  # (a,b,opp)
    key = np.minimum(he[:, 0], he[:, 1]) * nV + np.maximum(he[:, 0], he[:, 1])
    order = np.argsort(key, kind='stable'); key = key[order]; he = he[order]
    same = key[1:] == key[:-1]
    k0 = np.nonzero(same)[0]
    i = np.minimum(he[k0, 0], he[k0, 1]); j = np.maximum(he[k0, 0], he[k0, 1]); p = he[k0, 2]; q = he[k0 + 1, 2]
    e = V[j] - V[i]; L2 = (e * e).sum(1)
    def foot(o):
        d = V[o] - V[i]; s = (d * e).sum(1) / L2; r = d - s[:, None] * e; h = np.linalg.norm(r, axis=1); return s, h, r
    s1, h1, ra = foot(p); s2, h2, rb = foot(q)
    A1 = 0.5 * np.sqrt(L2) * h1; A2 = 0.5 * np.sqrt(L2) * h2
    c = (ra * rb).sum(1) / (h1 * h2); W = A1 + A2
    al = np.stack([-(1 - s1) / h1, -s1 / h1, 1 / h1, 0 * h1], axis=1)
    be = np.stack([-(1 - s2) / h2, -s2 / h2, 0 * h2, 1 / h2], axis=1)
    idx = np.stack([i, j, p, q], axis=1)
    blk = W[:, None, None] * (al[:, :, None] * al[:, None, :] + be[:, :, None] * be[:, None, :] - c[:, None, None] * (al[:, :, None] * be[:, None, :] + be[:, :, None] * al[:, None, :]))
    R = sp.coo_matrix((blk.ravel(), (np.repeat(idx, 4, axis=1).ravel(), np.tile(idx, (1, 4)).ravel())), shape=(nV, nV)).tocsr()
    A = (M + w * R).tocsr()
    return A, b, V, n

def pcg(A, b, Minv, tol=1e-6, maxit=20000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; bn = np.linalg.norm(b)
    for it in range(maxit):
        if np.linalg.norm(r) <= tol * bn: return x, it
        Ap = A @ p; alpha = rz / (p @ Ap); x += alpha * p; r -= alpha * Ap; z = Minv(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    blk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    A, b, V, n = build_system(n)
    nV = len(b); D = A.diagonal()
    x0, it0 = pcg(A, b, lambda r: r / D)
    print(f"n={n} unknowns {nV}: Jacobi {it0} iterations")
    ii, jj = np.divmod(np.arange(nV), n + 1)
    _, agg = np.unique((ii // blk) * (n + blk) + (jj // blk), return_inverse=True)
    na = agg.max() + 1
    cnt = np.bincount(agg, minlength=na)
    cen = np.stack([np.bincount(agg, weights=V[:, k], minlength=na) / cnt for k in range(3)], axis=1)
    for name, basis in (("constant", np.ones((nV, 1))), ("linear", np.concatenate([np.ones((nV, 1)), V - cen[agg]], axis=1))):
        k = basis.shape[1]
        P = sp.coo_matrix((basis.ravel(), (np.repeat(np.arange(nV), k), (k * agg[:, None] + np.arange(k)[None, :]).ravel())), shape=(nV, k * na)).tocsr()
        Ac = P.T @ A @ P
        lu = spla.splu((Ac + 1e-12 * Ac.diagonal().max() * sp.identity(Ac.shape[0])).tocsc())

        def Minv(r, omega=0.7):
            z = omega * r / D
            z += P @ lu.solve(P.T @ (r - A @ z))
            z += omega * (r - A @ z) / D
            return z
        x, it = pcg(A, b, Minv)
        print(f"  two-level, {blk}x{blk} aggregates, {name} coarse space ({k * na} dofs): {it} iterations, max |x - x_jacobi| {np.abs(x - x0).max():.1e}")
