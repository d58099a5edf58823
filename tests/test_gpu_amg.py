"""GPU: smoothed-aggregation AMG (device counterpart of pc_type hypre / gamg): Galerkin
operators against scipy, PCG solutions against a direct solve, determinism."""

import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import cashocs_b200 as cb
from cashocs_b200._forms import forms
from cashocs_b200._utils import amg, linalg

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def demo_poly(dim):
    x = cb.spatial_coordinate(dim)
    return 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1


def systems(kind):
    if kind == "cube":
        dm = cb.regular_mesh(14, 1.0, 1.0, 1.0)
        markers = [1, 2, 3, 4, 5, 6]
    elif kind == "square":
        dm = cb.regular_mesh(48)
        markers = [1, 2, 3, 4]
    else:
        dm = cb.import_mesh(os.path.join(GOLDEN, "mesh3d.npz"))
        markers = [1, 2]
    flag, val = forms.bc_flag_arrays(dm, forms.create_dirichlet_bcs(0.0, markers), 1)
    A, b = linalg.assemble_scalar_system(dm, 1.0, 0.0, f_poly=demo_poly(dm.dim), bc_flag=flag, bc_val=val)
    mu = torch.full((dm.num_vertices,), 0.35714285714285715, dtype=torch.float64, device="cuda")
    E = linalg.assemble_elasticity_matrix(dm, mu, 1.428571428571429, 0.2)
    rhs = torch.from_numpy(np.random.RandomState(1).rand(dm.dim * dm.num_vertices) - 0.5).cuda()
    return dm, A, b, E, rhs


def csr(rowptr, col, val, shape):
    return sp.csr_matrix((val.cpu().numpy(), col.cpu().numpy(), rowptr.cpu().numpy()), shape=shape)


@pytest.mark.parametrize("kind", ["cube", "square", "mesh3d"])
def test_hierarchy_is_galerkin(kind):
    dm, A, b, E, rhs = systems(kind)
    for M in (A, E):
        h = amg.AMGHierarchy(M)
        h.numeric(M)
        assert len(h.levels) >= 2 and h.operator_complexity < 2.5
        bs = M.bs
        cur = M.to_scipy()
        for li, L in enumerate(h.levels[:-1]):
            nxt, D = h.levels[li + 1], h.data[li]
            nxt.val = h.data[li + 1].val
            P = csr(L.p_rowptr, L.p_col, D.p_val, (L.n, L.nc))
            Pt = csr(L.pt_rowptr, L.pt_col, D.pt_val, (L.nc, L.n))
            assert abs(Pt - P.T).max() == 0.0
            # partition of unity of the smoothed prolongator on interior rows: row sums of P == 1 - w*rowsum(S)/S_ii
            assert np.all(np.diff(L.p_rowptr.cpu().numpy()) >= 1)
            Pb = sp.kron(P, sp.identity(bs)).tocsr() if bs > 1 else P
            ref = (Pb.T @ cur @ Pb).tocsr()
            if bs == 1:
                got = csr(nxt.rowptr, nxt.col, nxt.val, (nxt.n, nxt.n))
            else:
                got = sp.bsr_matrix((nxt.val.cpu().numpy().reshape(-1, bs, bs), nxt.col.cpu().numpy(), nxt.rowptr.cpu().numpy()),
                                    shape=(nxt.n * bs, nxt.n * bs)).tocsr()
            assert abs(got - ref).max() <= 1e-12 * abs(ref).max()
            cur = got
        # dense inverse of the coarsest operator
        if h.coarse_inv is not None:
            nn = h.levels[-1].n * bs
            inv = h.coarse_inv.cpu().numpy().reshape(nn, nn)
            assert np.abs(inv @ cur.toarray() - np.eye(nn)).max() < 1e-9
        # aggregates cover every node exactly once
        agg = h.levels[0].agg.cpu().numpy()
        assert agg.min() == 0 and agg.max() == h.levels[0].nc - 1 and np.unique(agg).size == h.levels[0].nc


@pytest.mark.parametrize("kind", ["cube", "square", "mesh3d"])
@pytest.mark.parametrize("pc", ["hypre", "gamg"])
def test_amg_pcg_matches_direct_solve(kind, pc):
    dm, A, b, E, rhs = systems(kind)
    for M, f in ((A, b), (E, rhs)):
        n = M.nrows * M.bs
        x = torch.zeros(n, dtype=torch.float64, device="cuda")
        opts = {"ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-11, "ksp_atol": 1e-30, "ksp_max_it": 500, "ksp_monitor": None}
        if pc == "hypre":
            opts["pc_hypre_type"] = "boomeramg"
        res = cb.LinearSolver().solve(x, A=M, b=f, ksp_options=opts)
        xj = torch.zeros_like(x)
        resj = cb.LinearSolver().solve(xj, A=M, b=f, ksp_options={"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-11,
                                                               "ksp_atol": 1e-30, "ksp_max_it": 20000})
        assert res.reason == 2 and res.its < resj.its
        assert np.all(np.diff(res.history) < 0) or res.history[-1] <= 1e-11 * res.history[0]
        xd = spla.spsolve(M.to_scipy().tocsc(), f.cpu().numpy())
        assert np.max(np.abs(x.cpu().numpy() - xd)) < 1e-8 * np.max(np.abs(xd))
        # bitwise reproducible
        x2 = torch.zeros_like(x)
        res2 = cb.LinearSolver().solve(x2, A=M, b=f, ksp_options=opts)
        assert res2.its == res.its and torch.equal(x, x2)


def test_reference_iterative_defaults_use_amg():
    """cashocs/_utils/linalg.py:44-52 (cg + boomeramg, rtol 1e-20 clamped) runs on the device AMG."""
    dm, A, b, E, rhs = systems("cube")
    o = linalg.parse_ksp_options(linalg.iterative_ksp_options)
    assert o.pc_type == 3 and o.rtol == 1e-14
    x = torch.zeros(A.nrows, dtype=torch.float64, device="cuda")
    opts = dict(linalg.iterative_ksp_options)
    opts["ksp_rtol"] = 1e-9  # gradient_tol (shape_gradient_problem.py:87-91)
    res = cb.LinearSolver().solve(x, A=A, b=b, ksp_options=opts)
    assert res.reason == 2 and res.its <= 40


@pytest.mark.parametrize("kind", ["cube", "mesh3d"])
def test_cuda_graph_replay_is_bitwise_identical(kind):
    """The PCG iteration captured into a CUDA graph (ksp_use_graph = 1; automatic on partitioned meshes)
    replays exactly the kernels of the eager loop: same iteration count, bitwise the same solution."""
    dm, A, b, E, rhs = systems(kind)
    for M, f in ((A, b), (E, rhs)):
        out = {}
        for mode in (-1, 1):
            for pc in ("hypre", "jacobi"):
                x = torch.zeros(M.nrows * M.bs, dtype=torch.float64, device="cuda")
                res = cb.LinearSolver().solve(x, A=M, b=f, ksp_options={
                    "ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-11, "ksp_atol": 1e-30, "ksp_max_it": 5000,
                    "ksp_use_graph": mode, "ksp_check_every": 3, "ksp_monitor": None})
                out[(mode, pc)] = (res.its, x.clone(), res.history.copy())
        for pc in ("hypre", "jacobi"):
            assert out[(1, pc)][0] == out[(-1, pc)][0] > 3
            assert torch.equal(out[(1, pc)][1], out[(-1, pc)][1])
            assert np.array_equal(out[(1, pc)][2], out[(-1, pc)][2])


@pytest.mark.parametrize("kind", ["cube", "mesh3d"])
def test_rowwise_galerkin_products_match_term_lists(kind):
    """Large operators build their Galerkin products by searching the output rows instead of term lists
    (AMGSymbolic.TERM_LIST_MAX_NNZ): same summation order, bitwise the same coarse operators."""
    dm, A, b, E, rhs = systems(kind)
    for M in (A, E):
        vals = {}
        try:
            for limit in (60_000_000, 0):
                amg.AMGSymbolic.TERM_LIST_MAX_NNZ = limit
                h = amg.AMGHierarchy(M)
                h.numeric(M)
                assert all(L.rowwise == (limit == 0) for L in h.levels[:-1])
                vals[limit] = [D.val.clone() for D in h.data[1:]] + [D.ap_val.clone() for D in h.data[:-1]]
        finally:
            amg.AMGSymbolic.TERM_LIST_MAX_NNZ = 60_000_000
        assert len(vals[0]) == len(vals[60_000_000]) >= 2
        for x, y in zip(vals[0], vals[60_000_000]):
            assert torch.equal(x, y)


@pytest.mark.parametrize("kind", ["cube", "square", "mesh3d"])
def test_fused_coarse_cycle_equals_the_launch_per_operation_cycle(kind):
    """The levels >= 1 of the V-cycle as one persistent kernel (``csrc/coarse_cycle.cuh``, grid barriers between the
    phases) against the same cycle launched operation by operation (``CASHOCS_B200_FUSED_COARSE=0``): same operators,
    same Chebyshev coefficients, only the summation order inside a row differs -- same iteration counts, solutions
    equal to round-off, and the fused path is bitwise reproducible (eager loop == graph replay, run == rerun)."""
    import os

    dm, A, b, E, rhs = systems(kind)
    for M, f in ((A, b), (E, rhs)):
        for lsd in (1, 2, 3):
            opts = {"ksp_type": "cg", "pc_type": "hypre", "ksp_rtol": 1e-11, "ksp_atol": 1e-30, "ksp_max_it": 500,
                    "pc_amg_level_smooth_degree": lsd}
            sol = {}
            for fused in ("1", "0"):
                os.environ["CASHOCS_B200_FUSED_COARSE"] = fused
                try:
                    x = torch.zeros(M.nrows * M.bs, dtype=torch.float64, device="cuda")
                    res = cb.LinearSolver().solve(x, A=M, b=f, ksp_options=opts)
                finally:
                    os.environ.pop("CASHOCS_B200_FUSED_COARSE", None)
                sol[fused] = (res.its, x)
            assert abs(sol["1"][0] - sol["0"][0]) <= 1, (sol["1"][0], sol["0"][0])
            scale = float(sol["0"][1].abs().max())
            assert float((sol["1"][1] - sol["0"][1]).abs().max()) <= 1e-9 * scale
            for graph in (-1, 0):
                x2 = torch.zeros_like(sol["1"][1])
                res2 = cb.LinearSolver().solve(x2, A=M, b=f, ksp_options=dict(opts, ksp_use_graph=graph))
                assert res2.its == sol["1"][0] and torch.equal(x2, sol["1"][1])


@pytest.mark.parametrize("kind", ["cube", "square"])
def test_rigid_body_coarse_correction(kind):
    """``pc_amg_rigid_body_correction``: M + Z (Z^T A Z)^-1 Z^T with the global rigid-body modes -- same solution as
    without (and as the direct solve), markedly fewer iterations for the damped free-boundary elasticity operator,
    the Galerkin matrix is the one scipy computes, eager and graph-replayed loops agree bitwise."""
    dm, A, b, E, rhs = systems(kind)
    d = dm.dim
    Es = E.to_scipy()
    want = spla.spsolve(Es.tocsc(), rhs.cpu().numpy())
    base = {"ksp_type": "cg", "pc_type": "hypre", "ksp_rtol": 1e-11, "ksp_atol": 1e-30, "ksp_max_it": 2000}
    out = {}
    for rbm in (False, True):
        for graph in (-1, 0):
            x = torch.zeros_like(rhs)
            opts = dict(base, pc_amg_rigid_body_correction=rbm, ksp_use_graph=graph)
            res = linalg.LinearSolver().solve(x, A=E, b=rhs, ksp_options=opts)
            assert np.max(np.abs(x.cpu().numpy() - want)) <= 1e-8 * np.max(np.abs(want))
            out[(rbm, graph)] = (res.its, x.clone())
    for rbm in (False, True):
        assert out[(rbm, -1)][0] == out[(rbm, 0)][0] and torch.equal(out[(rbm, -1)][1], out[(rbm, 0)][1])
    # 3-D: the three rotations are outliers of the spectrum (130 -> 89 iterations on the 10 M-tet bench mesh);
    # 2-D: one rotation, no gain expected, no harm allowed
    limit = 0.85 if d == 3 else 1.1
    assert out[(True, 0)][0] <= limit * out[(False, 0)][0], (out[(True, 0)][0], out[(False, 0)][0])
    # the small Galerkin matrix against scipy
    h = linalg.LinearSolver.amg_hierarchy(E, {})
    xc, einv, _ = linalg.LinearSolver.rigid_body_correction(E, h)
    c = xc.cpu().numpy().reshape(-1, d)[: dm.num_vertices]
    k = 6 if d == 3 else 3
    Z = np.zeros((k, dm.num_vertices, d))
    for a in range(d):
        Z[a, :, a] = 1.0
    if d == 3:
        Z[3, :, 0], Z[3, :, 1] = -c[:, 1], c[:, 0]
        Z[4, :, 1], Z[4, :, 2] = -c[:, 2], c[:, 1]
        Z[5, :, 0], Z[5, :, 2] = c[:, 2], -c[:, 0]
    else:
        Z[2, :, 0], Z[2, :, 1] = -c[:, 1], c[:, 0]
    Z = Z.reshape(k, -1)
    G = Z @ (Es @ Z.T)
    assert np.max(np.abs(einv.cpu().numpy().reshape(k, k) @ G - np.eye(k))) < 1e-9
