"""GPU parity: every CUDA kernel through the C-ABI against the CPU oracle (bit-exact for
indices and flags, <= 1e-12 relative for assembled values)."""

import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import cashocs_b200 as cb
from cashocs_b200._forms import forms
from cashocs_b200._utils import linalg
from oracle import fem, krylov, meshes, pipeline, quality

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def demo_poly(dim):
    x = cb.spatial_coordinate(dim)
    return 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1


def oracle_mesh(kind):
    if kind == "sq8":
        return meshes.regular_mesh(8)
    if kind == "cube4":
        return meshes.regular_mesh(4, 1.0, 1.0, 1.0)
    if kind == "box":
        return meshes.regular_mesh(3, 1.0, 2.0, 1.0)
    if kind == "circle":
        return meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    if kind == "mesh3d":
        return meshes.load_npz(os.path.join(GOLDEN, "mesh3d.npz"))
    raise KeyError(kind)


def device_mesh(kind):
    if kind == "sq8":
        return cb.regular_mesh(8)
    if kind == "cube4":
        return cb.regular_mesh(4, 1.0, 1.0, 1.0)
    if kind == "box":
        return cb.regular_mesh(3, 1.0, 2.0, 1.0)
    return cb.import_mesh(os.path.join(GOLDEN, {"circle": "unit_circle.npz", "mesh3d": "mesh3d.npz"}[kind]))


KINDS = ["sq8", "cube4", "box", "circle", "mesh3d"]


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("kind", ["sq8", "cube4", "box"])
def test_synthetic_mesh_matches_oracle(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    assert np.array_equal(dm.cells.cpu().numpy(), om.cells)
    assert np.array_equal(dm.coords.cpu().numpy(), om.coords)
    for marker in np.unique(om.facet_tags):
        assert np.array_equal(dm.boundary_vertices([int(marker)]).cpu().numpy(), om.boundary_vertices([marker]))


@pytest.mark.parametrize("kind", KINDS)
def test_pattern_bit_exact(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    dm.build_pattern()
    rowptr, col = fem.csr_pattern(om.cells, om.num_vertices)
    assert np.array_equal(dm.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(dm.col.cpu().numpy(), col)
    # scatter map: every (cell, a, b) points at the entry (row = dof a, col = dof b)
    k = om.dim + 1
    c2n = torch.empty_like(dm.cc2n)
    c2n[dm.colour_perm] = dm.cc2n
    c2n = c2n.cpu().numpy().reshape(-1, k, k)
    rows_of_nnz = np.repeat(np.arange(om.num_vertices), np.diff(rowptr))
    assert np.array_equal(rows_of_nnz[c2n], np.repeat(om.cells[:, :, None], k, axis=2))
    assert np.array_equal(col[c2n], np.repeat(om.cells[:, None, :], k, axis=1))
    # colouring: no two cells of one colour share a vertex
    cc = dm.ccells.cpu().numpy()
    for c in range(dm.ncolours):
        v = cc[dm.colour_off[c]:dm.colour_off[c + 1]].reshape(-1)
        assert np.unique(v).size == v.size
    assert np.array_equal(dm.valence.cpu().numpy(), quality.vertex_valence(om.cells, om.num_vertices))


@pytest.mark.parametrize("kind", KINDS)
def test_scalar_assembly_matches_oracle(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    d = om.dim
    bc = om.boundary_vertices([1, 3])
    g = 0.7
    _, vol, G = fem.cell_geometry(om.coords, om.cells)
    Ke = 1.3 * fem.stiffness_elements(vol, G) + 0.4 * fem.mass_elements(vol, d)
    fn = np.random.RandomState(1).rand(om.num_vertices)
    be = 0.9 * fem.load_elements_callable(om.coords, om.cells, vol, pipeline.demo_f, 5) - 2.0 * fem.load_elements_p1(vol, om.cells, fn, d)
    A0, b0 = fem.assemble_system(Ke, be, om.cells, om.num_vertices, bc, g)
    flag, val = forms.bc_flag_arrays(dm, [cb.DirichletBC(g, (1, 3))], 1)
    A, b = linalg.assemble_scalar_system(dm, 1.3, 0.4, f_poly=demo_poly(d), f_poly_scale=0.9,
                                         f_nodal=torch.from_numpy(fn).cuda(), f_nodal_scale=-2.0,
                                         bc_flag=flag, bc_val=val)
    As = A.to_scipy()
    assert np.array_equal(As.indptr, A0.indptr) and np.array_equal(As.indices, A0.indices)
    assert rel(As.data, A0.data) < 1e-12
    assert rel(b.cpu().numpy(), b0) < 1e-12
    # determinism: bitwise identical on a second assembly
    A2, b2 = linalg.assemble_scalar_system(dm, 1.3, 0.4, f_poly=demo_poly(d), f_poly_scale=0.9,
                                           f_nodal=torch.from_numpy(fn).cuda(), f_nodal_scale=-2.0,
                                           bc_flag=flag, bc_val=val)
    assert torch.equal(A.val, A2.val) and torch.equal(b, b2)


@pytest.mark.parametrize("kind", KINDS)
def test_elasticity_assembly_matches_oracle(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    d = om.dim
    mu = 0.5 + np.random.RandomState(2).rand(om.num_vertices)
    _, vol, G = fem.cell_geometry(om.coords, om.cells)
    Ae = fem.elasticity_elements(vol, G, mu, om.cells, 1.4, 0.2)
    vdofs = fem.vector_cell_dofs(om.cells, d)
    fix = om.boundary_vertices([1])
    bc = np.concatenate([d * fix + k for k in range(d)] + [d * om.boundary_vertices([2]) + 1])
    A0, _ = fem.assemble_system(Ae, None, vdofs, d * om.num_vertices, np.unique(bc), 0.0)
    flag, _ = forms.bc_flag_arrays(dm, [cb.DirichletBC(0.0, (1,)), cb.DirichletBC(0.0, (2,), 1)], d)
    A = linalg.assemble_elasticity_matrix(dm, torch.from_numpy(mu).cuda(), 1.4, 0.2, bc_flag=flag)
    As = A.to_scipy()
    As.sort_indices()
    diff = (As - A0)
    assert abs(diff).max() / abs(A0).max() < 1e-12
    assert abs(As - As.T).max() / abs(As).max() < 1e-15  # symmetry check (shape_form_handler.py:372-387)


@pytest.mark.parametrize("kind", KINDS)
def test_shape_derivative_matches_oracle(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    d = om.dim
    prob = pipeline.PoissonShapeProblem(om, bc_markers=(1,), config=pipeline.ShapeConfig(shape_bdry_fix=(2,) if kind != "circle" else ()))
    r = np.random.RandomState(3)
    prob.u, prob.p = r.rand(om.num_vertices), r.rand(om.num_vertices)
    b0 = prob.shape_derivative_vector()
    flag, _ = forms.bc_flag_arrays(dm, forms.create_dirichlet_bcs(0.0, [2] if kind != "circle" else []), d)
    b = linalg.assemble_shape_derivative(dm, torch.from_numpy(prob.u).cuda(), torch.from_numpy(prob.p).cuda(),
                                         demo_poly(d), 1.0, bc_flag=flag)
    assert rel(b.cpu().numpy(), b0) < 1e-12


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("pull_back", [True, False])
def test_general_shape_derivative_matches_oracle(kind, pull_back):
    """cb_assemble_shape_derivative for the whole Lagrangian family: kappa != 1, a reaction term and the tracking
    functional c_track/2 (u - u_d)^2 with a P1 u_d (the `J_func` case of tests/test_shape_optimization.py:247-299,
    including the pull-back term of shape_form_handler.py:507-531) against the oracle, <= 1e-12."""
    om, dm = oracle_mesh(kind), device_mesh(kind)
    d = om.dim
    r = np.random.RandomState(7)
    ud = om.coords[:, 0] ** 2 + om.coords[:, 1] ** 4 - np.pi
    prob = pipeline.PoissonShapeProblem(om, bc_markers=(1,), config=pipeline.ShapeConfig(shape_bdry_fix=(2,) if kind != "circle" else ()),
                                        kappa=1.7, reaction=0.8, c_state=0.3, c_track=2.0, u_d=ud, use_pull_back=pull_back)
    prob.u, prob.p = r.rand(om.num_vertices), r.rand(om.num_vertices)
    b0 = prob.shape_derivative_vector()
    flag, _ = forms.bc_flag_arrays(dm, forms.create_dirichlet_bcs(0.0, [2] if kind != "circle" else []), d)
    b = linalg.assemble_shape_derivative(dm, torch.from_numpy(prob.u).cuda(), torch.from_numpy(prob.p).cuda(),
                                         demo_poly(d), 0.3, bc_flag=flag, kappa=1.7, reaction=0.8, c_track=2.0,
                                         u_d=torch.from_numpy(ud).cuda(), pull_back=pull_back)
    assert rel(b.cpu().numpy(), b0) < 1e-12


@pytest.mark.parametrize("kind", KINDS)
def test_spmv_and_scalar_product(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    d = om.dim
    r = np.random.RandomState(4)
    A, _ = linalg.assemble_scalar_system(dm, 1.0, 1.0, rhs=False)
    x, y = r.rand(om.num_vertices), r.rand(om.num_vertices)
    As = A.to_scipy()
    y0 = As @ x
    y1 = A.mult(torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel(y1, y0) < 1e-14
    sp0 = float(y @ y0)
    assert abs(A.scalar_product(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()) - sp0) < 1e-13 * abs(sp0)
    E = linalg.assemble_elasticity_matrix(dm, torch.ones(om.num_vertices, dtype=torch.float64, device="cuda"), 1.4, 0.2)
    xv, yv = r.rand(d * om.num_vertices), r.rand(d * om.num_vertices)
    Es = E.to_scipy()
    assert rel(E.mult(torch.from_numpy(xv).cuda()).cpu().numpy(), Es @ xv) < 1e-14
    sp0 = float(yv @ (Es @ xv))
    assert abs(E.scalar_product(torch.from_numpy(xv).cuda(), torch.from_numpy(yv).cuda()) - sp0) < 1e-13 * abs(sp0)


@pytest.mark.parametrize("kind", ["sq8", "cube4", "circle", "mesh3d"])
def test_fp32_value_pass_equals_fp64_pass_on_the_rounded_matrix(kind):
    """cb_spmv_f32 (the pass the mixed-precision V-cycle runs over cb_mat.val32): values are stored in fp32, products
    and sums are fp64 in the same order as cb_spmv, so it equals cb_spmv on the rounded matrix bit for bit and the
    exact product to fp32 round-off of the entries."""
    dm = device_mesh(kind)
    r = np.random.RandomState(5)
    A, _ = linalg.assemble_scalar_system(dm, 1.0, 1.0, rhs=False)
    E = linalg.assemble_elasticity_matrix(dm, torch.ones(dm.num_vertices, dtype=torch.float64, device="cuda"), 1.4, 0.2)
    for M in (A, E):
        x = torch.from_numpy(r.rand(M.ncols * M.bs)).cuda()
        y32 = M.mult_f32(x)
        rounded = linalg.Matrix(dm, M.bs, val=M.val.to(torch.float32).to(torch.float64))
        assert torch.equal(y32, rounded.mult(x))
        y64 = M.mult(x)
        assert float((y32 - y64).abs().max()) <= 1e-6 * float(y64.abs().max())
        assert float((y32 - y64).abs().max()) > 0.0


def _poisson_system(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    markers = [1] if kind == "circle" else [1, 2, 3, 4]
    flag, val = forms.bc_flag_arrays(dm, forms.create_dirichlet_bcs(0.0, markers), 1)
    A, b = linalg.assemble_scalar_system(dm, 1.0, 0.0, f_poly=demo_poly(om.dim), bc_flag=flag, bc_val=val)
    return om, dm, A, b


@pytest.mark.parametrize("kind", ["sq8", "cube4", "circle", "mesh3d"])
@pytest.mark.parametrize("pc", ["jacobi", "none", "chebyshev"])
def test_cg_matches_oracle(kind, pc):
    om, dm, A, b = _poisson_system(kind)
    opts = {"ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-10, "ksp_atol": 1e-30, "ksp_max_it": 5000,
            "ksp_monitor": None}
    x = torch.zeros(om.num_vertices, dtype=torch.float64, device="cuda")
    solver = cb.LinearSolver()
    res = solver.solve(x, A=A, b=b, ksp_options=opts)
    As, bs = A.to_scipy(), b.cpu().numpy()
    hist = []
    o = {k: v for k, v in opts.items() if k != "ksp_monitor"}
    x0, reason0, its0, rn0 = krylov.solve(As, bs, o, history=hist)
    assert res.reason == reason0 == 2
    assert abs(res.its - its0) <= 1
    n = min(len(hist), len(res.history))
    assert np.allclose(res.history[: n // 2], hist[: n // 2], rtol=1e-6)
    xd = sp.linalg.spsolve(As.tocsc(), bs)
    assert rel(x.cpu().numpy(), xd) < 1e-8
    assert rel(x.cpu().numpy(), x0) < 1e-8


@pytest.mark.parametrize("kind", ["sq8", "cube4"])
def test_minres_matches_oracle(kind):
    om, dm, A, b = _poisson_system(kind)
    opts = {"ksp_type": "minres", "pc_type": "jacobi", "ksp_rtol": 1e-10, "ksp_atol": 1e-30, "ksp_max_it": 5000,
            "ksp_monitor": None}
    x = torch.zeros(om.num_vertices, dtype=torch.float64, device="cuda")
    res = cb.LinearSolver().solve(x, A=A, b=b, ksp_options=opts)
    As, bs = A.to_scipy(), b.cpu().numpy()
    hist = []
    x0, reason0, its0, _ = krylov.solve(As, bs, {k: v for k, v in opts.items() if k != "ksp_monitor"}, history=hist)
    assert res.reason == reason0 == 2
    assert abs(res.its - its0) <= 1
    assert np.allclose(res.history[: its0 // 2], hist[: its0 // 2], rtol=1e-6)
    assert rel(x.cpu().numpy(), sp.linalg.spsolve(As.tocsc(), bs)) < 1e-8


def test_ksp_error_reasons_and_options():
    om, dm, A, b = _poisson_system("sq8")
    x = torch.zeros(om.num_vertices, dtype=torch.float64, device="cuda")
    with pytest.raises(cb.PETScKSPError) as e:
        cb.LinearSolver().solve(x, A=A, b=b, ksp_options={"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12, "ksp_max_it": 3})
    assert e.value.error_code == -3
    # indefinite operator -> -10 like PETSc
    Aneg = linalg.Matrix(dm, 1, val=-A.val.clone())
    with pytest.raises(cb.PETScKSPError) as e:
        cb.LinearSolver().solve(x, A=Aneg, b=b, ksp_options={"ksp_type": "cg", "pc_type": "none"})
    assert e.value.error_code == -10
    # b None -> zero solution (linalg.py:522-524)
    x.fill_(3.0)
    cb.LinearSolver().solve(x, A=A, b=None)
    assert float(x.abs().max()) == 0.0
    with pytest.raises(cb.InputError):
        cb.LinearSolver().solve(x, A=A, b=b, ksp_options={"ksp_type": "gmres", "pc_type": "ilu"})
    # preonly + jacobi = one diagonal scaling (mesh_testing.py:74-81)
    cb.LinearSolver().solve(x, A=A, b=b, ksp_options={"ksp_type": "preonly", "pc_type": "jacobi"})
    assert rel(x.cpu().numpy(), b.cpu().numpy() / A.to_scipy().diagonal()) < 1e-15
    # None options = the reference's direct solver -> solution to 1e-8 of spsolve
    cb.LinearSolver().solve(x, A=A, b=b, ksp_options=None)
    assert rel(x.cpu().numpy(), sp.linalg.spsolve(A.to_scipy().tocsc(), b.cpu().numpy())) < 1e-8


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("measure", ["skewness", "maximum_angle", "radius_ratios", "condition_number"])
def test_quality_matches_oracle(kind, measure):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    r = np.random.RandomState(5)
    h = (np.prod(om.coords.max(0) - om.coords.min(0)) / om.num_cells) ** (1 / om.dim)
    pert = 0.15 * h * (r.rand(*om.coords.shape) - 0.5)
    om.coords += pert
    dm.coords += torch.from_numpy(pert).cuda()
    q0 = quality.MEASURES[measure](om.coords, om.cells)
    from cashocs_b200.geometry import quality as dq

    q1 = dq._CALCULATORS[measure]().compute(dm).cpu().numpy()
    assert np.max(np.abs(q1 - q0)) < 1e-13
    assert abs(cb.compute_mesh_quality(dm, "min", measure) - q0.min()) < 1e-13
    assert abs(cb.compute_mesh_quality(dm, "avg", measure) - np.average(q0)) < 1e-13
    assert abs(cb.compute_mesh_quality(dm, "quantile", measure, 0.351) - np.quantile(q0, 0.351)) < 1e-13


def test_quality_known_answers():
    """tests/test_geometry.py:138-269 of the reference."""
    m2 = cb.regular_mesh(4)
    assert abs(cb.compute_mesh_quality(m2, "min", "condition_number") - 0.4714045207910318) < 1e-14
    assert abs(cb.compute_mesh_quality(m2, "avg", "condition_number") - 0.4714045207910318) < 1e-14
    opt, a1, a2 = np.pi / 3, np.pi / 2, np.pi / 4
    q = min(1 - max((a1 - opt) / (np.pi - opt), (opt - a1) / opt), 1 - max((a2 - opt) / (np.pi - opt), (opt - a2) / opt))
    for meas in ("skewness", "maximum_angle"):
        for typ in ("min", "avg"):
            assert abs(cb.compute_mesh_quality(m2, typ, meas) - q) < 1e-14
    assert abs(cb.compute_mesh_quality(m2, "quantile", "skewness", 0.167) - q) < 1e-14
    m3 = cb.regular_mesh(4, 1.0, 1.0, 1.0)
    assert abs(cb.compute_mesh_quality(m3, "min", "condition_number") - 0.3162277660168379) < 1e-14
    assert abs(cb.compute_mesh_quality(m3, "avg", "condition_number") - 0.3162277660168379) < 1e-14
    assert abs(cb.compute_mesh_quality(m3, "min", "radius_ratios") - cb.compute_mesh_quality(m3, "avg", "radius_ratios")) < 1e-14


@pytest.mark.parametrize("kind", KINDS)
def test_det_check_move_revert_intersections(kind):
    om, dm = oracle_mesh(kind), device_mesh(kind)
    d = om.dim
    r = np.random.RandomState(6)
    h = (np.prod(om.coords.max(0) - om.coords.min(0)) / om.num_cells) ** (1 / d)
    tester = cb.APrioriMeshTester(dm)
    itest = cb.IntersectionTester(dm)
    handler = cb.DeformationHandler(dm, tester, itest)
    assert itest.test()
    assert np.array_equal(itest.compute_collisions().cpu().numpy(), quality.compute_collisions(om.coords, om.cells))
    for scale, expect in ((0.2 * h, True), (1e3, False)):
        V = scale * (2 * r.rand(om.num_vertices, d) - 1)
        ok0, mn0, mx0 = quality.a_priori_test(om.coords, om.cells, V)
        Vd = torch.from_numpy(V.reshape(-1)).cuda()
        mn1, mx1 = tester.determinants(Vd)
        assert abs(mn1 - mn0) <= 1e-12 * max(1, abs(mn0)) and abs(mx1 - mx0) <= 1e-12 * max(1, abs(mx0))
        assert tester.test(Vd, float("inf")) == ok0 == expect
        before = dm.coords.clone()
        moved = handler.move_mesh(Vd)
        assert moved == expect
        if moved:
            assert np.allclose(dm.coords.cpu().numpy(), om.coords + V, rtol=0, atol=1e-15)
            assert itest.test() == quality.intersection_test(om.coords + V, om.cells)
            handler.revert_transformation()
        assert torch.equal(dm.coords, before)  # rejected or reverted: coordinates untouched
    # a fold-over that keeps all determinants positive is impossible; instead overlap two far parts:
    V = np.zeros((om.num_vertices, d))
    far = np.argmax(om.coords[:, 0])
    V[far] = om.coords[np.argmin(om.coords[:, 0])] - om.coords[far] + 0.3 * h
    om2 = om.coords + V
    dm.coords.copy_(torch.from_numpy(om2).cuda())
    assert itest.test() == quality.intersection_test(om2, om.cells)
    assert np.array_equal(itest.counts.cpu().numpy() == dm.valence.cpu().numpy(),
                          quality.compute_collisions(om2, om.cells) == quality.vertex_valence(om.cells, om.num_vertices)) or not itest.test()


def test_translation_moves_and_reverts():
    """tests/test_shape_optimization.py:157-179 of the reference."""
    dm = device_mesh("circle")
    handler = cb.DeformationHandler(dm, cb.APrioriMeshTester(dm), cb.IntersectionTester(dm))
    before = dm.coords.clone()
    V = torch.zeros_like(dm.coords)
    V[:, 0] = 0.5
    assert handler.move_mesh(V.reshape(-1))
    assert float((dm.coords - before - V).abs().max()) < 1e-10
    handler.revert_transformation()
    assert torch.equal(dm.coords, before)
    W = torch.from_numpy(np.random.RandomState(300696).uniform(-1e3, 1e3, size=tuple(dm.coords.shape))).cuda()
    assert not handler.move_mesh(W.reshape(-1))
    assert torch.equal(dm.coords, before)


@pytest.mark.parametrize("kind", KINDS)
def test_gather_and_scatter_assembly_agree(kind):
    """The row-centric kernel, the entry-centric gather kernels and the coloured cell-scatter kernels assemble the
    same matrices (different summation order -> round-off only), including the per-cell scaling."""
    dm = device_mesh(kind)
    d = dm.dim
    dm.build_pattern()
    assert dm.n2c_ptr is not None and int(dm.n2c_ptr[-1]) == dm.n2c_idx.numel()
    mu = torch.from_numpy(0.5 + np.random.RandomState(3).rand(dm.num_vertices)).cuda()
    scale = torch.from_numpy(0.5 + np.random.RandomState(4).rand(dm.num_cells)).cuda()
    flag_v, _ = forms.bc_flag_arrays(dm, [cb.DirichletBC(0.0, (1,)), cb.DirichletBC(0.0, (2,), 1)], d)
    flag_s, val_s = forms.bc_flag_arrays(dm, [cb.DirichletBC(0.3, (1, 3))], 1)
    out = {}
    try:
        for gather in (True, "entries", False):  # row-centric (default), entry-centric gather, coloured scatter
            linalg.USE_GATHER_ASSEMBLY = gather
            E = linalg.assemble_elasticity_matrix(dm, mu, 1.4, 0.2, bc_flag=flag_v, cell_scale=scale)
            A, b = linalg.assemble_scalar_system(dm, 1.3, 0.4, f_poly=demo_poly(d), bc_flag=flag_s, bc_val=val_s)
            out[gather] = (E.val.clone(), A.val.clone(), b.clone())
            E2 = linalg.assemble_elasticity_matrix(dm, mu, 1.4, 0.2, bc_flag=flag_v, cell_scale=scale)
            assert torch.equal(E2.val, E.val)  # bitwise repeatable
    finally:
        linalg.USE_GATHER_ASSEMBLY = True
    for other in ("entries", False):
        for x, y in zip(out[True], out[other]):
            assert float((x - y).abs().max() / y.abs().max()) < 1e-13


def test_degenerate_inputs():
    """Smallest meshes and degenerate data: one cell, all-Dirichlet systems, zero right-hand sides, NaNs."""
    tri = cb.mesh_from_arrays(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]]),
                              np.array([[0, 1], [1, 2], [0, 2]]), np.array([1, 2, 3]))
    tet = cb.mesh_from_arrays(np.array([[0.0, 0, 0], [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]), np.array([[0, 1, 2, 3]]),
                              np.array([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]]), np.array([1, 2, 3, 4]))
    for m, vol in ((tri, 0.5), (tet, 1.0 / 6.0)):
        d = m.dim
        M, b = linalg.assemble_scalar_system(m, 0.0, 1.0, f_poly=cb.spatial_coordinate(d)[0] * 0.0 + 1.0)
        assert M.nnzb == (d + 1) ** 2 and abs(float(M.val.sum()) - vol) < 1e-15 and abs(float(b.sum()) - vol) < 1e-15
        assert cb.compute_mesh_quality(m, "min", "skewness") > 0.0
        assert cb.IntersectionTester(m).test()
        # every dof Dirichlet: identity system, solution = boundary value, zero iterations needed
        flag, val = forms.bc_flag_arrays(m, [cb.DirichletBC(2.5, tuple(m.boundary_markers))], 1)
        A, rhs = linalg.assemble_scalar_system(m, 1.0, 0.0, bc_flag=flag, bc_val=val)
        x = torch.zeros(m.num_vertices, dtype=torch.float64, device="cuda")
        cb.LinearSolver().solve(x, A=A, b=rhs, ksp_options={"ksp_type": "cg", "pc_type": "jacobi"})
        assert float((x - 2.5).abs().max()) < 1e-14
        # zero right-hand side: PETSc returns x = 0 with CONVERGED_ATOL
        res = cb.LinearSolver().solve(x, A=A, b=torch.zeros_like(rhs), ksp_options={"ksp_type": "cg", "pc_type": "jacobi"})
        assert res.reason > 0 and float(x.abs().max()) == 0.0
        # P2 space of a single cell
        sp2 = m.p2_space()
        assert sp2.ndofs == (d + 1) + (3 if d == 2 else 6) and sp2.nnz == sp2.ndofs**2
    # NaN coordinates: mesh measures propagate NaN (never silently pass a threshold), Krylov reports -9
    bad = cb.regular_mesh(3)
    bad.coords[5, 0] = float("nan")
    assert np.isnan(cb.compute_mesh_quality(bad, "min", "skewness"))
    V = torch.zeros(bad.num_vertices * 2, dtype=torch.float64, device="cuda")
    assert not cb.APrioriMeshTester(bad).test(V, float("inf"))
    A, b = linalg.assemble_scalar_system(bad, 1.0, 1.0, f_poly=cb.spatial_coordinate(2)[0])
    x = torch.zeros(bad.num_vertices, dtype=torch.float64, device="cuda")
    with pytest.raises(cb.PETScKSPError) as e:
        cb.LinearSolver().solve(x, A=A, b=b, ksp_options={"ksp_type": "cg", "pc_type": "jacobi"})
    assert e.value.error_code == -9


def test_trace_level_profiles_the_four_reference_timers():
    """cashocs/log.py:361-389 wraps assemble / solve / a-priori test / intersection test
    (linalg.py:107,483; mesh_testing.py:83,160); the same four actions are timed here with CUDA events at TRACE
    level.  Also: a-priori test with a finite volume_change (mesh_testing.py:114-117) decides like the oracle."""
    om, dm = oracle_mesh("cube4"), device_mesh("cube4")
    cb.log.set_log_level(cb.log.TRACE)
    cb.log.timings(reset=True)
    try:
        A, b = linalg.assemble_scalar_system(dm, 1.0, 1.0, f_poly=demo_poly(3))
        x = torch.zeros(dm.num_vertices, dtype=torch.float64, device="cuda")
        cb.LinearSolver().solve(x, A=A, b=b, ksp_options={"ksp_type": "cg", "pc_type": "jacobi"})
        tester = cb.APrioriMeshTester(dm)
        r = np.random.RandomState(11)
        h = 0.25
        for scale in (0.05, 0.3):
            for vc in (1.05, 1.5, 4.0):
                V = scale * h * (2 * r.rand(om.num_vertices, 3) - 1)
                ok0, _, _ = quality.a_priori_test(om.coords, om.cells, V, vc)
                assert tester.test(torch.from_numpy(V.reshape(-1)).cuda(), vc) == ok0
        cb.IntersectionTester(dm).test()
    finally:
        cb.log.set_log_level(cb.log.WARNING)
    t = cb.log.timings(reset=True)
    assert set(t) == {"assembling the system", "solving the linear system", "a priori mesh testing", "intersection testing"}
    assert all(n >= 1 and ms > 0.0 for n, ms in t.values()) and t["a priori mesh testing"][0] == 6
