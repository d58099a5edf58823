"""CPU prototype (numpy / scipy, no device code): how many PCG iterations does the elasticity Riesz solve of the
bench workload need when the smoothed-aggregation near-null-space holds the six rigid-body modes instead of the
three translations the round-1 device hierarchy uses (scalar prolongator shared by the components)?

    python tools/proto_amg_rbm.py --n 16 24

The elasticity matrix is the oracle's (all boundaries deformable, damping 0.2 -> no Dirichlet rows, rotations are
low-energy modes).  Same ingredients as the device preconditioner: MIS(2) aggregation on the node graph
(``cashocs_b200/_utils/amg.py``), damped-Jacobi prolongator smoothing, Galerkin products, V(1,1) with
degree-1 Chebyshev-Jacobi (ratio 10), dense coarsest solve, PCG to rtol 1e-10 on the preconditioned norm.
Result of the experiment is recorded in DESIGN.md section 8 ("next").
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cashocs_b200._utils import amg as host_amg  # noqa: E402  (host-side symbolic code, runs on CPU tensors)
from oracle import meshes, pipeline  # noqa: E402


def lam_max(A, dinv_blocks, bs, its=15):
    """Power iteration for lambda_max(D^-1 A), D = block diagonal."""
    rng = np.random.RandomState(0)
    x = rng.rand(A.shape[0])
    lam = 1.0
    for _ in range(its):
        y = apply_dinv(dinv_blocks, A @ x, bs)
        lam = np.linalg.norm(y) / np.linalg.norm(x)
        x = y / np.linalg.norm(y)
    return 1.15 * lam


def block_diag_inv(A, bs):
    n = A.shape[0] // bs
    A = A.tobsr(blocksize=(bs, bs))
    out = np.zeros((n, bs, bs))
    for i in range(n):
        for k in range(A.indptr[i], A.indptr[i + 1]):
            if A.indices[k] == i:
                out[i] = np.linalg.inv(A.data[k])
    return out


def apply_dinv(dinv, x, bs):
    return np.einsum("nij,nj->ni", dinv, x.reshape(-1, bs)).reshape(-1)


def tentative(B, agg, nc, bs):
    """Per-aggregate QR of the near-null-space B [n*bs, k] -> P_tent [n*bs, nc*k], coarse B [nc*k, k]."""
    k = B.shape[1]
    n = agg.size
    rows, cols, vals = [], [], []
    Bc = np.zeros((nc * k, k))
    order = np.argsort(agg, kind="stable")
    starts = np.searchsorted(agg[order], np.arange(nc + 1))
    for a in range(nc):
        nodes = order[starts[a]: starts[a + 1]]
        dofs = (nodes[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)
        Q, R = np.linalg.qr(B[dofs])
        if Q.shape[1] < k:  # tiny aggregate: pad
            Qp = np.zeros((dofs.size, k))
            Qp[:, : Q.shape[1]] = Q
            Rp = np.zeros((k, k))
            Rp[: R.shape[0]] = R
            Q, R = Qp, Rp
        rows.append(np.repeat(dofs, k))
        cols.append(np.tile(a * k + np.arange(k), dofs.size))
        vals.append(Q.reshape(-1))
        Bc[a * k: (a + 1) * k] = R
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n * bs, nc * k))
    return P, Bc


def build(A, B, bs, coarsest=600, max_levels=10):
    levels = []
    while True:
        n = A.shape[0] // bs
        dinv = block_diag_inv(A, bs)
        lmax = lam_max(A, dinv, bs)
        levels.append(dict(A=A, dinv=dinv, lmax=lmax, bs=bs))
        if n <= coarsest or len(levels) >= max_levels:
            levels[-1]["inv"] = np.linalg.inv(A.toarray())
            return levels
        G = A.tobsr(blocksize=(bs, bs))
        rows = np.repeat(np.arange(n), np.diff(G.indptr))
        agg, nc = host_amg.mis_aggregate(torch.from_numpy(rows.astype(np.int64)), torch.from_numpy(G.indices.astype(np.int64)), n, 2)
        agg = agg.numpy().astype(np.int64)
        Pt, Bc = tentative(B, agg, nc, bs)
        omega = 4.0 / 3.0 / lmax
        DA = sp.csr_matrix(sp.bsr_matrix((dinv, np.arange(n), np.arange(n + 1)), shape=A.shape)) @ A
        P = (Pt - omega * (DA @ Pt)).tocsr()
        levels[-1]["P"] = P
        A = (P.T @ A @ P).tocsr()
        B, bs = Bc, B.shape[1]


def cross_block(r):
    """3 x 3 matrix C with C @ w = w x r (the displacement at offset r of a rotation w)."""
    return np.array([[0.0, r[2], -r[1]], [-r[2], 0.0, r[0]], [r[1], -r[0], 0.0]])


def scalar_coarsen(A, types, surrogate=True):
    """The round-1 device coarsening (scalar prolongator (x) I_3, smoothed with S = trace(A_blk) / 3), with the
    aggregation restricted to nodes of the same type (translation / rotation nodes never share an aggregate)."""
    bs = 3
    n = A.shape[0] // bs
    G = A.tobsr(blocksize=(bs, bs))
    rows = np.repeat(np.arange(n), np.diff(G.indptr))
    same = types[rows] == types[G.indices]
    agg, nc = host_amg.mis_aggregate(torch.from_numpy(rows[same].astype(np.int64)),
                                     torch.from_numpy(G.indices[same].astype(np.int64)), n, 2)
    agg = agg.numpy().astype(np.int64)
    Sg = sp.csr_matrix((np.trace(G.data, axis1=1, axis2=2) / 3.0, G.indices, G.indptr), shape=(n, n))
    ds = Sg.diagonal()
    x0 = np.random.RandomState(0).rand(n)
    for _ in range(15):
        y0 = (Sg @ x0) / ds
        ls = np.linalg.norm(y0) / np.linalg.norm(x0)
        x0 = y0 / np.linalg.norm(y0)
    T = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
    Ps = (T - (4.0 / 3.0 / (1.15 * ls)) * sp.diags(1.0 / ds) @ Sg @ T).tocsr()
    ctype = np.zeros(nc, dtype=types.dtype)
    ctype[agg] = types
    return sp.kron(Ps, sp.identity(3)).tocsr(), ctype


def build_pairs(A, X, coarsest=600, max_levels=10, surrogate0=False, rbm_levels=99):
    """Rigid-body modes with 3 x 3 blocks only: every aggregate becomes TWO coarse nodes, one carrying the three
    translations, one the three rotations about the aggregate's centroid (scaled by its radius).  No QR, all
    operators stay 3 x 3 block matrices, so the device kernels for A keep their block size."""
    bs = 3
    levels = []
    centre = X.copy()                     # position of every node
    scale = np.zeros(X.shape[0])          # 0: translation node; s > 0: rotation node with coefficient scale s
    pair = np.arange(X.shape[0])          # level 0: every node is its own group; level >= 1: (2a, 2a+1) share a group
    while True:
        n = A.shape[0] // bs
        dinv = block_diag_inv(A, bs)
        lmax = lam_max(A, dinv, bs)
        levels.append(dict(A=A, dinv=dinv, lmax=lmax, bs=bs))
        if n <= coarsest or len(levels) >= max_levels:
            levels[-1]["inv"] = np.linalg.pinv(A.toarray())
            return levels
        if len(levels) > rbm_levels:   # below the RBM levels: the existing scalar coarsening, types kept apart
            P, types = scalar_coarsen(A, (scale > 0).astype(np.int64))
            levels[-1]["P"] = P
            A = (P.T @ A @ P).tocsr()
            scale = types.astype(float)
            continue
        # aggregate the graph of groups
        G = A.tobsr(blocksize=(bs, bs))
        rows = np.repeat(np.arange(n), np.diff(G.indptr))
        ng = int(pair.max()) + 1
        gr, gc = pair[rows], pair[G.indices]
        key = np.unique(gr.astype(np.int64) * ng + gc)
        agg_g, nc = host_amg.mis_aggregate(torch.from_numpy(key // ng), torch.from_numpy(key % ng), ng, 2)
        agg_g = agg_g.numpy().astype(np.int64)
        agg = agg_g[pair]                                        # aggregate of every node
        is_t = scale == 0.0
        cnt = np.bincount(agg[is_t], minlength=nc)
        C = np.stack([np.bincount(agg[is_t], weights=centre[is_t, k], minlength=nc) for k in range(3)], axis=1) / cnt[:, None]
        off = centre - C[agg]
        radius = np.linalg.norm(off, axis=1) + scale
        S = np.maximum.reduceat(radius[np.argsort(agg, kind="stable")],
                                np.searchsorted(np.sort(agg), np.arange(nc)))
        S = np.maximum(S, 1e-300)
        data, ri, ci = [], [], []
        for i in range(n):
            a = agg[i]
            if scale[i] == 0.0:   # translation node: I to the translations, (w x r) / S to the rotations
                ri += [i, i]
                ci += [2 * a, 2 * a + 1]
                data += [np.eye(3), cross_block(off[i]) / S[a]]
            else:                  # rotation node: coefficient conversion
                ri.append(i)
                ci.append(2 * a + 1)
                data.append(np.eye(3) * (scale[i] / S[a]))
        order = np.lexsort((ci, ri))
        ri, ci, data = np.asarray(ri)[order], np.asarray(ci)[order], np.asarray(data)[order]
        indptr = np.searchsorted(ri, np.arange(n + 1))
        Pt = sp.bsr_matrix((data, ci, indptr), shape=(n * bs, 2 * nc * bs)).tocsr()
        omega = 4.0 / 3.0 / lmax
        if surrogate0 and len(levels) == 1:
            # level 0 as on the device: smooth with the scalar surrogate S = trace(A_blk) / 3 (x) I_3, so that a
            # prolongator entry is one scalar p_iJ and one 3-vector q_iJ (P[i,(J,r)] = cross(q_iJ) / s_J)
            Sg = sp.csr_matrix((np.trace(G.data, axis1=1, axis2=2) / 3.0, G.indices, G.indptr), shape=(n, n))
            ds = Sg.diagonal()
            x0 = np.random.RandomState(0).rand(n)
            for _ in range(15):
                y0 = (Sg @ x0) / ds
                ls = np.linalg.norm(y0) / np.linalg.norm(x0)
                x0 = y0 / np.linalg.norm(y0)
            M = sp.identity(n) - (4.0 / 3.0 / (1.15 * ls)) * sp.diags(1.0 / ds) @ Sg
            P = (sp.kron(M, sp.identity(3)) @ Pt).tocsr()
        else:
            DA = sp.csr_matrix(sp.bsr_matrix((dinv, np.arange(n), np.arange(n + 1)), shape=A.shape)) @ A
            P = (Pt - omega * (DA @ Pt)).tocsr()
        levels[-1]["P"] = P
        A = (P.T @ A @ P).tocsr()
        centre = np.repeat(C, 2, axis=0)
        scale = np.zeros(2 * nc)
        scale[1::2] = S
        pair = np.repeat(np.arange(nc), 2)


def cheb_smooth(L, b, x):
    """One degree-1 Chebyshev-Jacobi sweep on [lmax/10, lmax]: x += 1/theta * D^-1 (b - A x)."""
    theta = 0.5 * (L["lmax"] + L["lmax"] / 10.0)
    r = b - L["A"] @ x if x is not None else b
    dx = apply_dinv(L["dinv"], r, L["bs"]) / theta
    return dx if x is None else x + dx


def vcycle(levels, k, b):
    L = levels[k]
    if "inv" in L:
        return L["inv"] @ b
    x = cheb_smooth(L, b, None)
    r = b - L["A"] @ x
    x = x + L["P"] @ vcycle(levels, k + 1, L["P"].T @ r)
    return cheb_smooth(L, b, x)


def pcg(A, b, M, rtol=1e-10, maxit=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    z0 = np.sqrt(z @ z)
    for it in range(1, maxit + 1):
        w = A @ p
        alpha = rz / (p @ w)
        x += alpha * p
        r -= alpha * w
        z = M(r)
        if np.sqrt(z @ z) <= rtol * z0:
            return x, it
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[12, 16])
    args = ap.parse_args()
    for n in args.n:
        om = meshes.regular_mesh(n, 1.0, 1.0, 1.0)
        cfg = pipeline.ShapeConfig(shape_bdry_def=(1, 2, 3, 4, 5, 6))
        prob = pipeline.PoissonShapeProblem(om, bc_markers=(1, 2, 3, 4, 5, 6), config=cfg)
        prob.update_scalar_product()
        A = prob.A_elas.tocsr()
        prob.solve_state()
        prob.solve_adjoint()
        b = prob.shape_derivative_vector()
        X = om.coords
        nv = X.shape[0]
        T = np.zeros((3 * nv, 3))
        for i in range(3):
            T[i::3, i] = 1.0
        c = X - X.mean(axis=0)
        R = np.zeros((3 * nv, 3))
        R[0::3, 0], R[1::3, 0] = -c[:, 1], c[:, 0]          # rotation about z
        R[1::3, 1], R[2::3, 1] = -c[:, 2], c[:, 1]          # about x
        R[0::3, 2], R[2::3, 2] = c[:, 2], -c[:, 0]          # about y
        for name, B in (("translations (round 1)", T), ("rigid-body modes", np.hstack([T, R]))):
            t = time.time()
            levels = build(A, B, 3)
            sizes = [L["A"].shape[0] for L in levels]
            nnz = [L["A"].nnz for L in levels]
            x, its = pcg(A, b, lambda r: vcycle(levels, 0, r))
            res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
            print(f"n={n:3d} dofs={A.shape[0]:7d} {name:24s} its={its:4d} levels={sizes} opc={sum(nnz) / nnz[0]:.2f} "
                  f"res={res:.1e} ({time.time() - t:.0f}s)", flush=True)
        t = time.time()
        levels = build_pairs(A, X)
        sizes = [L["A"].shape[0] for L in levels]
        nnz = [L["A"].nnz for L in levels]
        x, its = pcg(A, b, lambda r: vcycle(levels, 0, r))
        res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        print(f"n={n:3d} dofs={A.shape[0]:7d} {'RBM as node pairs, 3x3':24s} its={its:4d} levels={sizes} "
              f"opc={sum(nnz) / nnz[0]:.2f} res={res:.1e} ({time.time() - t:.0f}s)", flush=True)
        t = time.time()
        levels = build_pairs(A, X, surrogate0=True)
        sizes = [L["A"].shape[0] for L in levels]
        nnz = [L["A"].nnz for L in levels]
        x, its = pcg(A, b, lambda r: vcycle(levels, 0, r))
        res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        print(f"n={n:3d} dofs={A.shape[0]:7d} {'  + scalar smoothing, l0':24s} its={its:4d} levels={sizes} "
              f"opc={sum(nnz) / nnz[0]:.2f} res={res:.1e} ({time.time() - t:.0f}s)", flush=True)
        t = time.time()
        levels = build_pairs(A, X, surrogate0=True, rbm_levels=1)
        sizes = [L["A"].shape[0] for L in levels]
        nnz = [L["A"].nnz for L in levels]
        x, its = pcg(A, b, lambda r: vcycle(levels, 0, r))
        res = np.linalg.norm(b - A @ x) / np.linalg.norm(b)
        print(f"n={n:3d} dofs={A.shape[0]:7d} {'  + RBM on level 0 only':24s} its={its:4d} levels={sizes} "
              f"opc={sum(nnz) / nnz[0]:.2f} res={res:.1e} ({time.time() - t:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
