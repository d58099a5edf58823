"""Smoothed-aggregation AMG preconditioner for the CPU oracle (TEST INFRASTRUCTURE ONLY, see ``oracle/__init__.py``).

The reference preconditions its iterative solves with hypre BoomerAMG (``cashocs/_utils/linalg.py:44-52``:
``cg`` + ``hypre``/``boomeramg``); hypre is not installable here, so the CPU arm of ``bench.py`` uses this
scipy restatement of a multigrid preconditioner of the same class instead of one-level Jacobi -- the timed CPU
baseline is then algorithmically comparable (mesh-independent iteration counts) with the device path, which uses
smoothed aggregation too.  Nothing here is shared with the product (own aggregation, own cycle):

* aggregation: MIS(2) roots with hashed priorities, neighbours join the best adjacent aggregate;
* scalar problems: piecewise-constant tentative prolongator, one damped-Jacobi smoothing step;
* vector problems (elasticity): all six rigid-body modes, carried by *two* coarse nodes per aggregate (translations,
  rotations about the aggregate centroid scaled by its radius), so every level keeps d x d blocks;
* V(1,1) cycle with one Chebyshev-Jacobi step on [lmax/10, lmax] (block Jacobi), dense (pseudo-)inverse at the bottom.

"parity unpinned": PETSc / hypre iteration counts are not reproduced (nor are they by the device path).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _neighbour_max(indptr, indices, v):
    """max of v over the closed neighbourhood of every node (the graph must hold the diagonal)."""
    return np.maximum.reduceat(v[indices], indptr[:-1])


def mis2_aggregate(indptr, indices, n):
    """MIS(2) aggregation on a symmetric graph with self loops -> (agg [n], number of aggregates)."""
    idx = np.arange(n, dtype=np.int64)
    h = (idx * 2654435761 + 12345) % 2147483629
    h = (h * 40503 + 7) % 2147483629
    prio = h * n + idx
    state = np.zeros(n, dtype=np.int8)  # 0 undecided, 1 root, 2 covered
    for _ in range(400):
        und = state == 0
        if not und.any():
            break
        p = np.where(und, prio, -1)
        m = _neighbour_max(indptr, indices, _neighbour_max(indptr, indices, p))
        newroot = und & (p == m)
        state[newroot] = 1
        cov = _neighbour_max(indptr, indices, _neighbour_max(indptr, indices, newroot.astype(np.int64)))
        state[(state == 0) & (cov > 0)] = 2
    roots = state == 1
    nc = int(roots.sum())
    agg = np.full(n, -1, dtype=np.int64)
    agg[roots] = np.arange(nc)
    for _ in range(2):
        assigned = agg >= 0
        best = _neighbour_max(indptr, indices, np.where(assigned, prio, -1))
        join = (~assigned) & (best >= 0)
        agg[join] = agg[(best % n)[join]]
    if (agg < 0).any():  # isolated leftovers become their own aggregates
        left = np.flatnonzero(agg < 0)
        agg[left] = nc + np.arange(left.size)
        nc += left.size
    return agg, nc


def _cross_blocks(r):
    """[m, 3, 3] matrices C with C @ w = w x r."""
    C = np.zeros((r.shape[0], 3, 3))
    C[:, 0, 1], C[:, 0, 2] = r[:, 2], -r[:, 1]
    C[:, 1, 0], C[:, 1, 2] = -r[:, 2], r[:, 0]
    C[:, 2, 0], C[:, 2, 1] = r[:, 1], -r[:, 0]
    return C


def _rot2_blocks(r):
    """2-D: one rotation; [m, 2, 2] blocks whose first column is the displacement of a unit rotation at offset r."""
    C = np.zeros((r.shape[0], 2, 2))
    C[:, 0, 0], C[:, 1, 0] = -r[:, 1], r[:, 0]
    return C


class SmoothedAggregation:
    """``z = M r`` with ``M`` one V(1,1) cycle; ``bs`` = dofs per node, ``coords`` [n, d] switches the rigid-body
    modes on (vector problems)."""

    def __init__(self, A, bs: int = 1, coords=None, coarsest: int = 500, max_levels: int = 12,
                 point_jacobi: bool = False) -> None:
        self.bs = bs
        self.point_jacobi = point_jacobi  # diagonal instead of block-diagonal D (what the device smoothers use)
        self.levels = []
        A = sp.csr_matrix(A)
        rbm = coords is not None and bs in (2, 3)
        n = A.shape[0] // bs
        centre = None if not rbm else np.asarray(coords, dtype=float).copy()
        scale = np.zeros(n)   # 0: translation node; > 0: rotation node (coefficient scale = aggregate radius)
        group = np.arange(n)  # nodes of one aggregate pair stay together when aggregating again
        while True:
            n = A.shape[0] // bs
            G = A.tobsr(blocksize=(bs, bs)) if bs > 1 else A
            rows = np.repeat(np.arange(n), np.diff(G.indptr))
            diag = G.indices == rows
            if bs > 1:
                D = G.data[diag].copy()
                null = np.abs(D).sum(axis=(1, 2)) == 0.0  # rotation node of a one-node aggregate: no energy
                D[null] = np.eye(bs)
                # 2-D rotation nodes use one of their two dofs: regularise the unused one
                for k in range(bs):
                    z = D[:, k, k] == 0.0
                    D[z, k, k] = 1.0
                if point_jacobi:
                    Dp = np.zeros_like(D)
                    for k in range(bs):
                        Dp[:, k, k] = D[:, k, k]
                    D = Dp
                dinv = np.linalg.inv(D)
            else:
                dinv = 1.0 / G.data[diag]
            L = dict(A=A, dinv=dinv)
            L["lmax"] = self._lambda_max(L)
            self.levels.append(L)
            if n <= coarsest or len(self.levels) >= max_levels:
                L["inv"] = np.linalg.pinv(A.toarray())
                return
            # graph of groups
            ng = int(group.max()) + 1
            key = np.unique(group[rows].astype(np.int64) * ng + group[G.indices])
            g_rows, g_cols = key // ng, key % ng
            g_indptr = np.searchsorted(g_rows, np.arange(ng + 1))
            agg_g, nc = mis2_aggregate(g_indptr, g_cols, ng)
            if nc >= 0.8 * ng:
                L["inv"] = np.linalg.pinv(A.toarray()) if A.shape[0] <= 4000 else None
                return
            agg = agg_g[group]
            if not rbm:
                T = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
                T = sp.kron(T, sp.identity(bs)).tocsr() if bs > 1 else T
                ncoarse = nc
            else:
                d = bs
                is_t = scale == 0.0
                cnt = np.maximum(np.bincount(agg[is_t], minlength=nc), 1)
                C = np.stack([np.bincount(agg[is_t], weights=centre[is_t, k], minlength=nc) for k in range(d)], axis=1)
                C /= cnt[:, None]
                off = centre - C[agg]
                radius = np.linalg.norm(off, axis=1) + scale
                S = np.zeros(nc)
                np.maximum.at(S, agg, radius)
                S = np.where(S > 0.0, S, 1.0)
                tn = np.flatnonzero(is_t)
                rn = np.flatnonzero(~is_t)
                rot = _cross_blocks(off[tn]) if d == 3 else _rot2_blocks(off[tn])
                eye = np.broadcast_to(np.eye(d), (tn.size, d, d))
                conv = np.eye(d)[None, :, :] * (scale[rn] / S[agg[rn]])[:, None, None]
                ri = np.concatenate([tn, tn, rn])
                ci = np.concatenate([2 * agg[tn], 2 * agg[tn] + 1, 2 * agg[rn] + 1])
                data = np.concatenate([eye, rot / S[agg[tn]][:, None, None], conv])
                order = np.lexsort((ci, ri))
                ri, ci, data = ri[order], ci[order], data[order]
                T = sp.bsr_matrix((data, ci, np.searchsorted(ri, np.arange(n + 1))), shape=(n * d, 2 * nc * d)).tocsr()
                ncoarse = 2 * nc
            omega = 4.0 / 3.0 / L["lmax"]
            P = (T - omega * self._dinv_apply_matrix(L, A @ T)).tocsr()
            L["P"] = P
            A = (P.T @ A @ P).tocsr()
            A.eliminate_zeros()
            if rbm:
                centre = np.repeat(C, 2, axis=0)
                scale = np.zeros(ncoarse)
                scale[1::2] = S
                group = np.repeat(np.arange(nc), 2)
            else:
                group = np.arange(ncoarse)

    # -- helpers
    def _dinv(self, L, x):
        if self.bs == 1:
            return L["dinv"] * x
        return np.einsum("nij,nj->ni", L["dinv"], x.reshape(-1, self.bs)).reshape(-1)

    def _dinv_apply_matrix(self, L, M):
        n = L["A"].shape[0] // self.bs
        if self.bs == 1:
            return sp.diags(L["dinv"]) @ M
        Dm = sp.bsr_matrix((L["dinv"], np.arange(n), np.arange(n + 1)), shape=L["A"].shape)
        return Dm @ M

    def _lambda_max(self, L, its: int = 15) -> float:
        x = np.random.RandomState(0).rand(L["A"].shape[0]) + 0.5
        lam = 1.0
        for _ in range(its):
            y = self._dinv(L, L["A"] @ x)
            nrm = np.linalg.norm(y)
            if nrm == 0.0:
                return 1.0
            lam = nrm / np.linalg.norm(x)
            x = y / nrm
        return 1.15 * lam

    def _smooth(self, L, b, x):
        theta = 0.5 * (L["lmax"] + L["lmax"] / 10.0)
        if x is None:
            return self._dinv(L, b) / theta
        return x + self._dinv(L, b - L["A"] @ x) / theta

    def _cycle(self, k, b):
        L = self.levels[k]
        if "inv" in L and L["inv"] is not None:
            return L["inv"] @ b
        if "P" not in L:  # stalled coarsening on a large level: a few smoothing steps
            x = None
            for _ in range(8):
                x = self._smooth(L, b, x)
            return x
        x = self._smooth(L, b, None)
        x = x + L["P"] @ self._cycle(k + 1, L["P"].T @ (b - L["A"] @ x))
        return self._smooth(L, b, x)

    def __call__(self, r):
        return self._cycle(0, r)

    @property
    def sizes(self):
        return [L["A"].shape[0] for L in self.levels]

    @property
    def operator_complexity(self) -> float:
        return sum(L["A"].nnz for L in self.levels) / self.levels[0]["A"].nnz
