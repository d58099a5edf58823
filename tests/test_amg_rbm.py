"""Symbolic phase of the rigid-body-mode hierarchy (``cashocs_b200/_utils/amg_rbm.py``, round-2 work in progress) and
the arithmetic of its numeric kernels, restated in numpy *over the term lists the symbolic phase built* and checked
against scipy's sparse products.  CPU only: the CUDA kernels of ``csrc/amg_rbm.cu`` implement exactly the functions
``emulate_*`` below (same operands, same summation order)."""

import numpy as np
import scipy.sparse as sp
import torch

from cashocs_b200._utils import amg_rbm
from oracle import amg as oracle_amg
from oracle import krylov, meshes, pipeline

ALL = (1, 2, 3, 4, 5, 6)


def elasticity(n):
    om = meshes.regular_mesh(n, 1.0, 1.0, 1.0)
    cfg = pipeline.ShapeConfig(shape_bdry_def=ALL, test_for_intersections=False)
    prob = pipeline.PoissonShapeProblem(om, bc_markers=ALL, config=cfg)
    prob.solve_state()
    prob.solve_adjoint()
    A = sp.bsr_matrix(prob.A_elas, blocksize=(3, 3))
    A.sort_indices()
    return om, A, prob.shape_derivative_vector()


def cross_blocks(r):
    C = np.zeros((r.shape[0], 3, 3))
    C[:, 0, 1], C[:, 0, 2] = r[:, 2], -r[:, 1]
    C[:, 1, 0], C[:, 1, 2] = -r[:, 2], r[:, 0]
    C[:, 2, 0], C[:, 2, 1] = r[:, 1], -r[:, 0]
    return C


# ---- numpy restatement of the numeric kernels (one function per kernel of csrc/amg_rbm.cu) -----------------------
def emulate_aggregate_geometry(L, centre, scale):
    """k_rbm_aggregate_geometry: centroid of the translation nodes of every aggregate, then its radius."""
    ptr, idx = L.member_ptr.numpy(), L.member_idx.numpy()
    kind = L.kind.numpy()
    C, S = np.zeros((L.nc, 3)), np.zeros(L.nc)
    for a in range(L.nc):
        nodes = idx[ptr[a]: ptr[a + 1]]
        t = nodes[kind[nodes] == 0]
        C[a] = centre[t].sum(axis=0) / max(len(t), 1)
        rad = np.linalg.norm(centre[nodes] - C[a], axis=1) + scale[nodes]
        S[a] = rad.max() if rad.max() > 0.0 else 1.0
    return C, S


def emulate_tentative(L, centre, scale, C, S):
    """k_rbm_tentative: one 3 x 3 block per entry of T."""
    rows = np.repeat(np.arange(L.n), np.diff(L.t_rowptr.numpy()))
    col, agg, kind = L.t_col.numpy(), L.agg.numpy(), L.kind.numpy()
    T = np.zeros((col.size, 3, 3))
    a = agg[rows]
    trans = (kind[rows] == 0) & (col == 2 * a)
    rot = (kind[rows] == 0) & (col == 2 * a + 1)
    conv = kind[rows] == 1
    T[trans] = np.eye(3)
    T[rot] = cross_blocks(centre[rows[rot]] - C[a[rot]]) / S[a[rot]][:, None, None]
    T[conv] = np.eye(3)[None] * (scale[rows[conv]] / S[a[conv]])[:, None, None]
    return T


def emulate_spgemm_bb(seg, sl, sr, lhs, rhs, trans_l=False):
    """k_spgemm_terms_bb: out[q] = sum_k op(lhs[sl[k]]) @ rhs[sr[k]] over the segment of q, ascending k."""
    seg, sl, sr = seg.numpy(), sl.numpy(), sr.numpy()
    out = np.zeros((seg.size - 1, 3, 3))
    for q in range(seg.size - 1):
        for k in range(seg[q], seg[q + 1]):
            lb = lhs[sl[k]].T if trans_l else lhs[sl[k]]
            out[q] += lb @ rhs[sr[k]]
    return out


def emulate_prolongator_finish(L, T, AT, dinv, omega):
    """k_rbm_prolongator_finish: P = -omega * dinv_row .* (A T), then + T on T's positions."""
    rows = np.repeat(np.arange(L.n), np.diff(L.p_rowptr.numpy()))
    P = -omega * dinv.reshape(-1, 3)[rows][:, :, None] * AT
    P[L.t2p.numpy()] += T
    return P


def bsr(rowptr, col, val, shape):
    return sp.bsr_matrix((val, col.numpy(), rowptr.numpy()), shape=shape)


def lambda_max(A, dinv, its=15):
    x = np.random.RandomState(0).rand(A.shape[0]) + 0.5
    lam = 1.0
    for _ in range(its):
        y = dinv * (A @ x)
        lam = np.linalg.norm(y) / np.linalg.norm(x)
        x = y / np.linalg.norm(y)
    return 1.15 * lam


def build_hierarchy(om, A):
    sym = amg_rbm.RBMSymbolic(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int32)),
                              A.shape[0] // 3)
    centre, scale = om.coords.copy(), np.zeros(om.num_vertices)
    val = A.data.copy()
    levels = []
    for L in sym.levels:
        Al = bsr(L.rowptr, L.col, val, (3 * L.n, 3 * L.n))
        dinv = 1.0 / np.where(Al.diagonal() != 0.0, Al.diagonal(), 1.0)
        lv = dict(A=sp.csr_matrix(Al), dinv=dinv, lmax=lambda_max(Al, dinv))
        levels.append(lv)
        if L.last:
            lv["inv"] = np.linalg.pinv(Al.toarray())
            break
        C, S = emulate_aggregate_geometry(L, centre, scale)
        T = emulate_tentative(L, centre, scale, C, S)
        AT = emulate_spgemm_bb(L.at_seg, L.at_sl, L.at_sr, val, T)
        omega = 4.0 / 3.0 / lv["lmax"]
        P = emulate_prolongator_finish(L, T, AT, dinv, omega)
        AP = emulate_spgemm_bb(L.ap_seg, L.ap_sl, L.ap_sr, val, P)
        Ac = emulate_spgemm_bb(L.ac_seg, L.ac_sl, L.ac_sr, P, AP, trans_l=True)
        lv.update(L=L, T=T, P=P, AP=AP, Ac=Ac, omega=omega, C=C, S=S)
        nxt = sym.levels[len(levels)]
        val = Ac
        centre = np.repeat(C, 2, axis=0)
        scale = np.zeros(2 * L.nc)
        scale[1::2] = S
        assert nxt.n == 2 * L.nc
    return sym, levels


def test_symbolic_phase_and_kernel_arithmetic_against_scipy():
    om, A, _ = elasticity(6)
    sym, levels = build_hierarchy(om, A)
    assert len(levels) >= 2 and sym.sizes[1] == 2 * sym.levels[0].nc
    for lv in levels[:-1]:
        L = lv["L"]
        n, ncn = L.n, 2 * L.nc
        Ts = sp.csr_matrix(bsr(L.t_rowptr, L.t_col, lv["T"], (3 * n, 3 * ncn)))
        Ps = sp.csr_matrix(bsr(L.p_rowptr, L.p_col, lv["P"], (3 * n, 3 * ncn)))
        P_ref = Ts - lv["omega"] * sp.diags(lv["dinv"]) @ (lv["A"] @ Ts)
        assert abs(Ps - P_ref).max() <= 1e-13 * abs(P_ref).max()
        APs = sp.csr_matrix(bsr(L.ap_rowptr, L.ap_col, lv["AP"], (3 * n, 3 * ncn)))
        assert abs(APs - lv["A"] @ P_ref).max() <= 1e-12 * abs(APs).max()
        nxt = levels[levels.index(lv) + 1]
        Ac_ref = P_ref.T @ lv["A"] @ P_ref
        assert abs(nxt["A"] - Ac_ref).max() <= 1e-12 * abs(Ac_ref).max()
        assert abs(nxt["A"] - nxt["A"].T).max() <= 1e-12 * abs(nxt["A"]).max()
        # P^T pattern: permutation of P's entries sorted by column
        perm = L.pt_perm.numpy()
        assert np.array_equal(L.p_col.numpy()[perm], np.repeat(np.arange(ncn), np.diff(L.pt_rowptr.numpy())))
    # the six global rigid-body modes are interpolated exactly by the tentative prolongators of every level
    c = om.coords - om.coords.mean(axis=0)
    nv = om.num_vertices
    B = np.zeros((3 * nv, 6))
    for i in range(3):
        B[i::3, i] = 1.0
    B[0::3, 3], B[1::3, 3] = -c[:, 1], c[:, 0]
    B[1::3, 4], B[2::3, 4] = -c[:, 2], c[:, 1]
    B[0::3, 5], B[2::3, 5] = c[:, 2], -c[:, 0]
    for lv in levels[:-1]:
        L = lv["L"]
        Ts = sp.csr_matrix(bsr(L.t_rowptr, L.t_col, lv["T"], (3 * L.n, 6 * L.nc)))
        coef = np.linalg.lstsq(Ts.toarray(), B, rcond=None)[0]
        assert np.max(np.abs(Ts @ coef - B)) <= 1e-10 * np.max(np.abs(B))
        B = coef


def test_rbm_hierarchy_cuts_the_iterations():
    """V(1,1) point-Jacobi cycle over the emulated hierarchy: PCG needs about a third of the iterations of the
    translations-only hierarchy (the round-1 device algorithm) on the bench's elasticity operator."""
    om, A, b = elasticity(10)
    _, levels = build_hierarchy(om, A)
    for lv in levels[:-1]:
        L = lv["L"]
        lv["Ps"] = sp.csr_matrix(bsr(L.p_rowptr, L.p_col, lv["P"], (3 * L.n, 6 * L.nc)))

    def smooth(lv, rhs, x):
        theta = 0.5 * (lv["lmax"] + lv["lmax"] / 10.0)
        return lv["dinv"] * rhs / theta if x is None else x + lv["dinv"] * (rhs - lv["A"] @ x) / theta

    def cycle(k, rhs):
        lv = levels[k]
        if "inv" in lv:
            return lv["inv"] @ rhs
        x = smooth(lv, rhs, None)
        x = x + lv["Ps"] @ cycle(k + 1, lv["Ps"].T @ (rhs - lv["A"] @ x))
        return smooth(lv, rhs, x)

    As = sp.csr_matrix(A)
    _, reason, its_rbm, _ = krylov.cg(As, b, lambda r: cycle(0, r), rtol=1e-10, atol=1e-30)
    M0 = oracle_amg.SmoothedAggregation(As, bs=3, coords=None, point_jacobi=True, coarsest=100)
    _, _, its_trans, _ = krylov.cg(As, b, M0, rtol=1e-10, atol=1e-30)
    assert reason > 0 and its_rbm <= 40 and its_rbm <= 0.55 * its_trans, (its_rbm, its_trans)
