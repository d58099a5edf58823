"""CPU: pin the oracle against every known answer / invariant the reference's tests hold for the
hot path and against the committed mesh fixtures."""

import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import fem, krylov, meshes, pipeline, quality

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_quality_known_answers_2d():
    """tests/test_geometry.py:138-195 (reference)."""
    m = meshes.regular_mesh(4)
    opt, a1, a2 = np.pi / 3, np.pi / 2, np.pi / 4
    q = min(1 - max((a1 - opt) / (np.pi - opt), (opt - a1) / opt), 1 - max((a2 - opt) / (np.pi - opt), (opt - a2) / opt))
    for meas in ("skewness", "maximum_angle"):
        for typ in ("min", "avg"):
            assert abs(quality.compute_mesh_quality(m.coords, m.cells, typ, meas) - q) < 1e-14
    assert abs(quality.compute_mesh_quality(m.coords, m.cells, "quantile", "skewness", 0.167) - q) < 1e-14
    for typ, qt in (("min", 0.0), ("avg", 0.0), ("quantile", 0.75189)):
        assert abs(quality.compute_mesh_quality(m.coords, m.cells, typ, "condition_number", qt) - 0.4714045207910318) < 1e-14
    rr = quality.radius_ratios(m.coords, m.cells)
    assert abs(rr.min() - rr.mean()) < 1e-14


def test_quality_known_answers_3d():
    """tests/test_geometry.py:198-269 (reference)."""
    m = meshes.regular_mesh(4, 1.0, 1.0, 1.0)
    ang = quality.cell_angles(m.coords, m.cells)
    opt = np.arccos(1 / 3)
    amin, amax = ang.min(), ang.max()
    q = min(1 - max((amax - opt) / (np.pi - opt), (opt - amax) / opt), 1 - max((amin - opt) / (np.pi - opt), (opt - amin) / opt))
    r = min(1 - max((amax - opt) / (np.pi - opt), 0.0), 1 - max((amin - opt) / (np.pi - opt), 0.0))
    assert abs(quality.compute_mesh_quality(m.coords, m.cells, "min", "skewness") - q) < 1e-14
    assert abs(quality.compute_mesh_quality(m.coords, m.cells, "avg", "skewness") - q) < 1e-14
    assert abs(quality.compute_mesh_quality(m.coords, m.cells, "min", "maximum_angle") - r) < 1e-14
    assert abs(quality.compute_mesh_quality(m.coords, m.cells, "quantile", "maximum_angle", 0.715) - r) < 1e-14
    for typ, qt in (("min", 0.0), ("avg", 0.0), ("quantile", 0.91576)):
        assert abs(quality.compute_mesh_quality(m.coords, m.cells, typ, "condition_number", qt) - 0.3162277660168379) < 1e-14
    rr = quality.radius_ratios(m.coords, m.cells)
    assert abs(rr.min() - rr.mean()) < 1e-14


def test_regular_mesh_invariants():
    """tests/test_geometry.py:51-99 style invariants: volume 1, each side 1, Euler counts."""
    m = meshes.regular_mesh(5)
    _, vol, _ = fem.cell_geometry(m.coords, m.cells)
    assert abs(vol.sum() - 1.0) < 1e-14 and m.num_cells == 50 and m.num_vertices == 36
    seg = np.linalg.norm(m.coords[m.facets[:, 1]] - m.coords[m.facets[:, 0]], axis=1)
    for tag in (1, 2, 3, 4):
        assert abs(seg[m.facet_tags == tag].sum() - 1.0) < 1e-14
    m3 = meshes.regular_mesh(3, 1.0, 2.0, 1.0)
    _, vol, _ = fem.cell_geometry(m3.coords, m3.cells)
    assert abs(vol.sum() - 2.0) < 1e-13 and m3.num_cells == 6 * 3 * 6 * 3
    assert sorted(np.unique(m3.facet_tags)) == [1, 2, 3, 4, 5, 6]
    # interior vertex of the Kuhn mesh: 15 nnz per row, 24 incident tets
    m4 = meshes.regular_mesh(4, 1.0, 1.0, 1.0)
    rowptr, _ = fem.csr_pattern(m4.cells, m4.num_vertices)
    centre = 2 + 5 * 2 + 25 * 2
    assert rowptr[centre + 1] - rowptr[centre] == 15
    assert quality.vertex_valence(m4.cells, m4.num_vertices)[centre] == 24


def test_golden_mesh_fixtures():
    """The reference's gmsh fixtures (tests/mesh/*.msh) converted by tests/golden/make_fixtures.py."""
    mc = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    assert (mc.num_vertices, mc.num_cells) == (713, 1340) and set(mc.facet_tags) == {1}
    _, vol, _ = fem.cell_geometry(mc.coords, mc.cells)
    assert abs(vol.sum() - np.pi) < 5e-3
    assert np.allclose(np.linalg.norm(mc.coords[mc.boundary_vertices([1])], axis=1), 1.0, atol=1e-12)
    m2 = meshes.load_npz(os.path.join(GOLDEN, "mesh2d.npz"))
    _, vol, _ = fem.cell_geometry(m2.coords, m2.cells)
    assert abs(vol.sum() - 1.0) < 1e-14 and m2.num_vertices == 12  # int 1 dx = 1 (tests/test_geometry.py:51-99)
    seg = np.linalg.norm(m2.coords[m2.facets[:, 1]] - m2.coords[m2.facets[:, 0]], axis=1)
    assert abs(seg.sum() - 4.0) < 1e-14
    for tag in (1, 2, 3, 4):
        assert abs(seg[m2.facet_tags == tag].sum() - 1.0) < 1e-14
    m3 = meshes.load_npz(os.path.join(GOLDEN, "mesh3d.npz"))
    _, vol, _ = fem.cell_geometry(m3.coords, m3.cells)
    assert abs(vol.sum() - 1.0) < 1e-13 and sorted(set(m3.facet_tags)) == [1, 2, 3, 4, 5, 6]


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/mesh/mesh.msh"), reason="reference tree not present")
def test_fixtures_match_reference_files():
    for name, rel in (("unit_circle", "unit_circle/mesh.msh"), ("mesh2d", "mesh.msh"), ("mesh3d", "mesh3.msh")):
        a = meshes.read_gmsh41(os.path.join("/root/reference/tests/mesh", rel))
        b = meshes.load_npz(os.path.join(GOLDEN, name + ".npz"))
        assert np.array_equal(a.coords, b.coords) and np.array_equal(a.cells, b.cells)
        assert np.array_equal(a.facets, b.facets) and np.array_equal(a.facet_tags, b.facet_tags)


def test_quadrature_exactness():
    from math import factorial

    for d in (2, 3):
        lam, w = fem.quadrature(d, 6)
        rng = np.random.RandomState(0)
        for _ in range(10):
            e = rng.randint(0, 3, size=d + 1)
            if e.sum() > 6:
                continue
            exact = factorial(d) * np.prod([factorial(int(k)) for k in e]) / factorial(int(e.sum()) + d)
            assert abs(np.sum(w * np.prod(lam ** e, axis=1)) - exact) < 1e-14


def test_state_adjoint_gradient_definitions(rng):
    """tests/test_pde_problems.py:87-128: oracle pipeline == textbook systems solved directly."""
    m = meshes.regular_mesh(10)
    y_d, u = rng.rand(m.num_vertices), rng.rand(m.num_vertices)
    ocp = pipeline.PoissonControlProblem(m, y_d, u, alpha=1e-6, bc_markers=(1, 2))
    g = ocp.compute_gradient()
    _, vol, G = fem.cell_geometry(m.coords, m.cells)
    K = fem.assemble_matrix(fem.stiffness_elements(vol, G), m.cells, m.num_vertices).tolil()
    M = fem.assemble_matrix(fem.mass_elements(vol, 2), m.cells, m.num_vertices)
    bc = m.boundary_vertices([1, 2])
    free = np.setdiff1d(np.arange(m.num_vertices), bc)
    Kff = K.tocsr()[free][:, free].tocsc()
    y = np.zeros(m.num_vertices)
    y[free] = spla.spsolve(Kff, (M @ u)[free])
    p = np.zeros(m.num_vertices)
    p[free] = spla.spsolve(Kff, -(M @ (y - y_d))[free])
    gref = spla.spsolve(M.tocsc(), 1e-6 * (M @ u) - M @ p)
    assert np.allclose(ocp.y, y) and np.allclose(ocp.p, p) and np.allclose(g, gref)


def test_bc_elimination_is_symmetric_and_counts_cells():
    """BC diagonal = number of cells sharing the dof, matrix symmetric, solution = g on BC."""
    m = meshes.regular_mesh(6)
    _, vol, G = fem.cell_geometry(m.coords, m.cells)
    bc = m.boundary_vertices([1, 2, 3, 4])
    be = fem.load_elements_callable(m.coords, m.cells, vol, pipeline.demo_f, 5)
    A, b = fem.assemble_system(fem.stiffness_elements(vol, G), be, m.cells, m.num_vertices, bc, 0.25)
    assert abs(A - A.T).max() < 1e-15
    assert np.array_equal(A.diagonal()[bc], quality.vertex_valence(m.cells, m.num_vertices)[bc].astype(float))
    x = spla.spsolve(A.tocsc(), b)
    assert np.allclose(x[bc], 0.25)


def test_krylov_petsc_semantics():
    m = meshes.regular_mesh(12)
    _, vol, G = fem.cell_geometry(m.coords, m.cells)
    bc = m.boundary_vertices([1, 2, 3, 4])
    be = fem.load_elements_callable(m.coords, m.cells, vol, pipeline.demo_f, 5)
    A, b = fem.assemble_system(fem.stiffness_elements(vol, G), be, m.cells, m.num_vertices, bc, 0.0)
    xd = spla.spsolve(A.tocsc(), b)
    for ksp in ("cg", "minres"):
        for pc in ("jacobi", "none"):
            hist = []
            x, reason, its, rn = krylov.solve(A, b, {"ksp_type": ksp, "pc_type": pc, "ksp_rtol": 1e-10}, history=hist)
            assert reason == krylov.CONVERGED_RTOL and len(hist) == its + 1
            assert hist[-1] <= 1e-10 * hist[0] < hist[-2]
            assert np.max(np.abs(x - xd)) < 1e-8 * np.max(np.abs(xd))
    x, reason, its, _ = krylov.solve(A, b, {"ksp_type": "cg", "pc_type": "chebyshev", "pc_chebyshev_degree": 3, "ksp_rtol": 1e-10})
    assert reason == 2 and np.max(np.abs(x - xd)) < 1e-8 * np.max(np.abs(xd))
    _, reason, its, _ = krylov.solve(A, b, {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-12, "ksp_max_it": 3})
    assert reason == krylov.DIVERGED_ITS and its == 3
    _, reason, _, _ = krylov.solve(-A, b, {"ksp_type": "cg", "pc_type": "none"})
    assert reason == krylov.DIVERGED_INDEFINITE_MAT
    x, reason, its, _ = krylov.solve(A, None)
    assert np.all(x == 0.0)


def test_taylor_test_and_iteration_budgets_unit_circle():
    """tests/test_shape_optimization.py:302-353 (reference): rate > 1.9; L-BFGS <= 7, GD <= 32 iterations."""
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02)
    rate, _, _ = pipeline.shape_gradient_test(pipeline.PoissonShapeProblem(load(), config=cfg))
    assert rate > 1.9
    ok, hist = pipeline.lbfgs(pipeline.PoissonShapeProblem(load(), config=cfg), rtol=1e-2, max_iter=7)
    assert ok and hist[-1][2] <= 1e-2
    ok, hist = pipeline.gradient_descent(pipeline.PoissonShapeProblem(load(), config=cfg), rtol=1e-2, max_iter=32)
    assert ok and hist[-1][2] <= 1e-2


def test_reextension_surface_reference_budget():
    """tests/test_shape_optimization.py:1072-1081 (``reextend_from_boundary``, mode ``surface``): L-BFGS reaches
    rtol 1e-2 within the reference's 7 iterations; the re-extended deformation keeps the boundary values of the
    Riesz gradient and solves the homogeneous elasticity equation in the interior."""
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, reextend_from_boundary=True)
    prob = pipeline.PoissonShapeProblem(load(), config=cfg)
    g = prob.compute_shape_gradient()
    g0 = pipeline.PoissonShapeProblem(load(), config=pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02)).compute_shape_gradient()
    bdry = prob.mesh.boundary_vertices((1,))
    bd = (2 * bdry[:, None] + np.arange(2)[None, :]).reshape(-1)
    assert np.array_equal(g[bd], g0[bd]) and np.max(np.abs(g - g0)) > 1e-6
    interior = np.setdiff1d(np.arange(g.size), bd)
    free = pipeline.PoissonShapeProblem(load(), config=pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02))
    r = free.A_elas @ g  # shape_bdry_fix is empty here: A_elas is the unconstrained operator
    assert np.max(np.abs(r[interior])) <= 1e-9 * np.max(np.abs(r[bd]))
    ok, hist = pipeline.lbfgs(pipeline.PoissonShapeProblem(load(), config=cfg), rtol=1e-2, max_iter=7)
    assert ok and len(hist) - 1 == 6 and hist[-1][2] <= 1e-2


def test_taylor_test_inhomogeneous_mu():
    """tests/test_shape_optimization.py:645-670 style: mu_fix=1, mu_def=10 on regular_mesh(20)."""
    m = meshes.regular_mesh(20)
    cfg = pipeline.ShapeConfig(shape_bdry_def=(1, 2), shape_bdry_fix=(3, 4), mu_def=10.0, mu_fix=1.0, inhomogeneous=True)
    prob = pipeline.PoissonShapeProblem(m, bc_markers=(1, 2, 3, 4), config=cfg)
    assert prob.mu.min() >= 1.0 - 1e-12 and prob.mu.max() <= 10.0 + 1e-12
    rate, _, _ = pipeline.shape_gradient_test(prob)
    assert rate > 1.9
    # non-uniform cells (the scaling 1 / (|K| / max|K|)^gamma is not constant), sqrt(mu), gamma = 0.7
    m2 = meshes.regular_mesh(12)
    m2.coords += 0.2 / 12 * (np.random.RandomState(9).rand(*m2.coords.shape) - 0.5) * (m2.coords > 0) * (m2.coords < 1)
    cfg = pipeline.ShapeConfig(shape_bdry_def=(1, 2), shape_bdry_fix=(3, 4), mu_def=10.0, mu_fix=1.0, inhomogeneous=True,
                               inhomogeneous_exponent=0.7, use_sqrt_mu=True)
    assert pipeline.shape_gradient_test(pipeline.PoissonShapeProblem(m2, bc_markers=(1, 2, 3, 4), config=cfg))[0] > 1.9


def test_move_revert_and_rejection(rng):
    """tests/test_shape_optimization.py:157-179 (reference)."""
    m = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    prob = pipeline.PoissonShapeProblem(m)
    before = m.coords.copy()
    V = np.zeros_like(m.coords)
    V[:, 0] = 0.5
    assert prob.move_mesh(V.reshape(-1))
    assert np.max(np.abs(m.coords - before - V)) < 1e-10
    prob.revert_transformation()
    assert np.array_equal(m.coords, before)
    assert not prob.move_mesh(rng.uniform(-1e3, 1e3, size=m.coords.size))
    assert np.array_equal(m.coords, before)


def test_collisions_equal_valence_on_valid_meshes():
    for m in (meshes.regular_mesh(5), meshes.regular_mesh(3, 1.0, 1.0, 1.0), meshes.load_npz(os.path.join(GOLDEN, "mesh2d.npz"))):
        assert np.array_equal(quality.compute_collisions(m.coords, m.cells), quality.vertex_valence(m.cells, m.num_vertices))
        assert quality.intersection_test(m.coords, m.cells)


# ----------------------------------------------------------------------------- P2 states (BASELINE configs[4])
@pytest.mark.parametrize("dim", [2, 3])
def test_p2_patch_test_and_reference_tensors(dim):
    """Quadratics are reproduced exactly by the P2 stiffness / load tensors; the reference mass matrix
    integrates 1; edge counts follow the closed formula of SURVEY.md 9.1."""
    from oracle import p2

    om = meshes.regular_mesh(4) if dim == 2 else meshes.regular_mesh(3, 1.0, 1.0, 1.0)
    nv = om.num_vertices
    dofs, edges = p2.cell_dofs(om.cells, nv)
    if dim == 3:
        n = 3
        assert len(edges) == 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n**3
    nd = nv + len(edges)
    _, vol, G = fem.cell_geometry(om.coords, om.cells)
    K = fem.assemble_matrix(p2.stiffness_elements(vol, G), dofs, nd)
    M = fem.assemble_matrix(p2.mass_elements(vol, dim), dofs, nd)
    q = lambda x: x[..., 0] ** 2 + 2 * x[..., 0] * x[..., 1] - 0.5 * x[..., 1] ** 2 + x[..., 0] + (0.25 * x[..., 2] ** 2 if dim == 3 else 0.0)  # noqa: E731
    lap = 2.0 - 1.0 + (0.5 if dim == 3 else 0.0)
    u = p2.interpolate(q, om.coords, edges)
    b = fem.assemble_vector(p2.load_elements_callable(om.coords, om.cells, vol, lambda x: -lap * np.ones(x.shape[:-1]), 0), dofs, nd)
    bd = p2.boundary_dofs(om, np.unique(om.facet_tags), edges)
    interior = np.setdiff1d(np.arange(nd), bd)
    assert np.abs((K @ u - b)[interior]).max() < 1e-13
    assert abs(M.sum() - 1.0) < 1e-13
    assert abs(p2.integral(u, dofs, vol, dim) - u @ (M @ np.ones(nd))) < 1e-13


@pytest.mark.parametrize("kind", ["square", "cube", "circle"])
def test_p2_taylor_test(kind):
    """Shape gradient with P2 state / adjoint passes the reference's Taylor test (rate > 1.9,
    cashocs/verification.py:182-302; tests/test_shape_optimization.py:302-318)."""
    if kind == "square":
        om, mk = meshes.regular_mesh(10), (1, 2, 3, 4)
    elif kind == "cube":
        om, mk = meshes.regular_mesh(4, 1.0, 1.0, 1.0), (1, 2, 3, 4, 5, 6)
    else:
        om, mk = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz")), (1,)
    cfg = pipeline.ShapeConfig(shape_bdry_def=mk, test_for_intersections=False)
    prob = pipeline.PoissonShapeProblem(om, bc_markers=mk, config=cfg, state_degree=2)
    rate, rates, _ = pipeline.shape_gradient_test(prob)
    assert rate > 1.9 and min(rates) > 1.9


@pytest.mark.parametrize("method,budget", [("FR", 21), ("PR", 16), ("HS", 18), ("DY", 18), ("HZ", 18)])
def test_ncg_iteration_budgets_unit_circle(method, budget):
    """tests/test_shape_optimization.py:321-353 (reference): NCG FR / PR / HS / DY / HZ reach rtol 1e-2 within
    21 / 16 / 18 / 18 / 18 iterations on the unit-disc fixture."""
    om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    prob = pipeline.PoissonShapeProblem(om, config=pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02))
    ok, hist = pipeline.ncg(prob, method=method, rtol=1e-2, max_iter=budget)
    assert ok and hist[-1][2] <= 1e-2


CONTROL_BUDGETS = [("gd", {}, 46), ("bfgs", {}, 7), ("ncg", {"method": "FR"}, 20), ("ncg", {"method": "PR"}, 25),
                   ("ncg", {"method": "HS"}, 27), ("ncg", {"method": "DY"}, 9), ("ncg", {"method": "HZ"}, 27),
                   ("ncg", {"method": "DY", "periodic_restart": True, "periodic_its": 5}, 10),
                   ("ncg", {"method": "DY", "relative_restart": True, "restart_tol": 1.0}, 24)]


def control_problem():
    om = meshes.regular_mesh(10)
    yd = np.sin(2 * np.pi * om.coords[:, 0]) * np.sin(2 * np.pi * om.coords[:, 1])
    return pipeline.PoissonControlProblem(om, yd, np.zeros(om.num_vertices), alpha=1e-6, bc_markers=(1, 2, 3, 4))


@pytest.mark.parametrize("algorithm,kwargs,budget", CONTROL_BUDGETS)
def test_control_iteration_budgets(algorithm, kwargs, budget):
    """tests/test_optimal_control.py:147-195 (reference, regular_mesh(10), y_d = sin(2 pi x) sin(2 pi y),
    alpha = 1e-6, tests/config_ocp.ini): gd <= 46, L-BFGS(3) <= 7, NCG FR / PR / HS / DY / HZ <= 20 / 25 / 27 / 9 / 27,
    DY with periodic restart (5) <= 10, DY with relative restart (1.0) <= 24 iterations to rtol 1e-2.
    The oracle needs exactly one iteration less than every one of these nine budgets."""
    prob = control_problem()
    if algorithm == "gd":
        ok, hist = pipeline.gradient_descent(prob, rtol=1e-2, max_iter=budget)
    elif algorithm == "bfgs":
        ok, hist = pipeline.lbfgs(prob, rtol=1e-2, max_iter=budget, memory=3, scaling=True)
    else:
        ok, hist = pipeline.ncg(prob, rtol=1e-2, max_iter=budget, **kwargs)
    assert ok and hist[-1][2] <= 1e-2 and len(hist) - 1 == budget - 1


def test_angle_change_and_fixed_dimensions_unit_circle():
    """tests/test_shape_optimization.py:934-968 (reference): with angle_change = 0.1 L-BFGS converges within 15
    iterations; with fixed_dimensions = [k] component k of the shape gradient is exactly zero and the Taylor
    test still has rate > 1.9."""
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, angle_change=0.1)
    ok, hist = pipeline.lbfgs(pipeline.PoissonShapeProblem(load(), config=cfg), rtol=1e-2, max_iter=15)
    assert ok and hist[-1][2] < 1e-2
    for k in (0, 1):
        cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, fixed_dimensions=(k,))
        prob = pipeline.PoissonShapeProblem(load(), config=cfg)
        g = prob.compute_shape_gradient().reshape(-1, 2)
        assert np.all(g[:, k] == 0.0) and np.abs(g[:, 1 - k]).max() > 0.0
        assert pipeline.shape_gradient_test(prob)[0] > 1.9


def test_small_initial_stepsize_breaks_the_armijo_rule():
    """tests/test_shape_optimization.py:1003-1012 (reference): initial_stepsize = 1e-3, NCG (DY), rtol 1e-3 ->
    "Armijo rule failed." -- the step collapses below the limits of armijo_line_search.py:75-94."""
    om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, initial_stepsize=1e-3)
    ok, hist = pipeline.ncg(pipeline.PoissonShapeProblem(om, config=cfg), method="DY", rtol=1e-3, max_iter=1000)
    assert not ok and len(hist) - 1 < 1000 and hist[-1][2] > 1e-3


def test_oracle_multigrid_preconditioner():
    """oracle/amg.py (the CPU arm's stand-in for hypre): same solutions as the direct solver, iteration counts that
    do not grow with the mesh, rigid-body modes for the elasticity system."""
    from oracle import amg, krylov

    its = {}
    for n in (8, 14):
        om = meshes.regular_mesh(n, 1.0, 1.0, 1.0)
        cfg = pipeline.ShapeConfig(shape_bdry_def=(1, 2, 3, 4, 5, 6), test_for_intersections=False)
        prob = pipeline.PoissonShapeProblem(om, bc_markers=(1, 2, 3, 4, 5, 6), config=cfg)
        g_direct = prob.compute_shape_gradient().copy()
        u_direct = prob.u.copy()
        ksp = {"ksp_type": "cg", "pc_type": "hypre", "ksp_rtol": 1e-11, "ksp_atol": 1e-30}
        prob2 = pipeline.PoissonShapeProblem(om, bc_markers=(1, 2, 3, 4, 5, 6), config=cfg, state_ksp=ksp,
                                             adjoint_ksp=ksp, gradient_ksp=ksp)
        g_amg = prob2.compute_shape_gradient()
        assert np.max(np.abs(prob2.u - u_direct)) <= 1e-8 * np.max(np.abs(u_direct))
        assert np.max(np.abs(g_amg - g_direct)) <= 1e-8 * np.max(np.abs(g_direct))
        its[n] = dict(prob2.its)
        M = amg.SmoothedAggregation(prob2.A_elas, bs=3, coords=om.coords, coarsest=100)
        # the six rigid-body modes are reproduced by the coarse space: P maps some coarse vector onto each of them
        P = M.levels[0]["P"]
        c = om.coords - om.coords.mean(axis=0)
        rot = np.zeros(3 * om.num_vertices)
        rot[0::3], rot[1::3] = -c[:, 1], c[:, 0]
        # smoothing changes the modes only where A B != 0, i.e. by the damping term; check the tentative part:
        # the residual of the least-squares fit of the mode in range(P) is small compared with the mode itself
        import scipy.sparse.linalg as spla

        coef = spla.lsqr(P, rot, atol=1e-12, btol=1e-12, iter_lim=2000)[0]
        assert np.linalg.norm(P @ coef - rot) <= 5e-2 * np.linalg.norm(rot)
    assert its[14]["gradient"] <= its[8]["gradient"] + 8 and its[14]["gradient"] <= 40
    assert its[14]["state"] <= 35
    # 2-D: two translations and one rotation per aggregate
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02)
    g_direct = pipeline.PoissonShapeProblem(load(), config=cfg).compute_shape_gradient()
    ksp = {"ksp_type": "cg", "pc_type": "hypre", "ksp_rtol": 1e-12, "ksp_atol": 1e-30}
    prob = pipeline.PoissonShapeProblem(load(), config=cfg, state_ksp=ksp, adjoint_ksp=ksp, gradient_ksp=ksp)
    g_amg = prob.compute_shape_gradient()
    assert np.max(np.abs(g_amg - g_direct)) <= 1e-9 * np.max(np.abs(g_direct)) and prob.its["gradient"] <= 40


def test_volume_and_barycenter_regularization_reference_scenarios():
    """tests/test_shape_optimization.py:356-378,406-434 (reference): Taylor rate > 1.9 with the regularisation terms,
    L-BFGS drives the unit disc to the target radius / barycenter within the reference's tolerances."""
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    rng = np.random.RandomState(300696)
    radius = rng.uniform(0.33, 0.66)
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, factor_volume=1.0, target_volume=np.pi * radius**2,
                               volume_change=10.0)
    prob = pipeline.PoissonShapeProblem(load(), config=cfg, c_state=0.0)
    assert pipeline.shape_gradient_test(prob, rng=rng)[0] > 1.9
    ok, hist = pipeline.lbfgs(prob, rtol=1e-6, max_iter=50, memory=3)
    assert ok
    assert abs(prob.coords.max() - radius) < 5e-3 and abs(-prob.coords.min() - radius) < 5e-3
    assert 0.5 * (prob.volume() - np.pi * radius**2) ** 2 < 1e-10

    rng = np.random.RandomState(300696)
    pos_x, pos_y = rng.uniform(0.2, 0.4), rng.uniform(-0.4, -0.2)
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, factor_volume=1e2, use_initial_volume=True,
                               factor_barycenter=1.0, target_barycenter=(pos_x, pos_y), volume_change=10.0)
    prob = pipeline.PoissonShapeProblem(load(), config=cfg, c_state=0.0)
    v0 = prob.volume()
    assert pipeline.shape_gradient_test(prob, rng=rng)[0] > 1.9
    ok, hist = pipeline.lbfgs(prob, rtol=1e-5, max_iter=50, memory=3)
    bc = prob.barycenter()
    assert ok and abs(prob.volume() - v0) < 1e-2 and abs(bc[0] - pos_x) < 1e-4 and abs(bc[1] - pos_y) < 1e-4


def test_surface_regularization_reference_scenario():
    """tests/test_shape_optimization.py:381-403 (reference): surface regularisation with target 2 pi r."""
    rng = np.random.RandomState(300696)
    radius = rng.uniform(0.33, 0.66)
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, factor_surface=1.0, target_surface=2 * np.pi * radius,
                               volume_change=10.0)
    prob = pipeline.PoissonShapeProblem(meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz")), config=cfg, c_state=0.0)
    assert pipeline.shape_gradient_test(prob, rng=rng)[0] > 1.9
    ok, hist = pipeline.lbfgs(prob, rtol=1e-6, max_iter=50, memory=3)
    assert ok
    assert abs(prob.coords.max() - radius) < 5e-3 and abs(-prob.coords.min() - radius) < 5e-3
    assert 0.5 * (prob.surface() - 2 * np.pi * radius) ** 2 < 1e-10
    # 3-D: closed-form gradient of the boundary measure against a difference quotient
    om = meshes.regular_mesh(3, 1.0, 1.0, 1.0)
    p3 = pipeline.PoissonShapeProblem(om, bc_markers=(1, 2, 3, 4, 5, 6), c_state=0.0,
                                      config=pipeline.ShapeConfig(shape_bdry_def=(1, 2, 3, 4, 5, 6), factor_surface=1.0))
    assert abs(p3.surface() - 6.0) < 1e-13
    g = p3.surface_derivative_vector()
    h = 1e-6 * np.random.RandomState(1).rand(g.size)
    s0 = p3.surface()
    p3.coords += h.reshape(-1, 3)
    assert abs((p3.surface() - s0) - g @ h) < 1e-9 * abs(g @ h) + 1e-11


def test_control_bcs_reference_budget():
    """tests/test_optimal_control.py:569-580 (reference): Dirichlet conditions on the control, L-BFGS <= 9 iterations,
    the control keeps its boundary value exactly."""
    value = np.random.RandomState(300696).rand()
    om = meshes.regular_mesh(10)
    yd = np.sin(2 * np.pi * om.coords[:, 0]) * np.sin(2 * np.pi * om.coords[:, 1])
    prob = pipeline.PoissonControlProblem(om, yd, np.zeros(om.num_vertices), alpha=1e-6, bc_markers=(1, 2, 3, 4),
                                          control_bcs=((1, 2, 3, 4), value))
    ok, hist = pipeline.lbfgs(prob, rtol=1e-2, max_iter=9, memory=3)
    b = om.boundary_vertices((1, 2, 3, 4))
    assert ok and len(hist) - 1 <= 9 and np.all(prob.u[b] == value)
    g = prob.compute_gradient()
    assert np.all(g[b] == 0.0)


def test_newton_solver_and_semilinear_control_gradient():
    """tests/test_nonlinear_solvers.py:28-61 (reference): -Laplace u + 1e2 u^3 = 1 on regular_mesh(5) by Newton from
    zero with rtol 1e-9 / atol 1e-10 equals the solution of an independent nonlinear solve;
    tests/test_optimal_control.py:364-376: control problem with the y^3 term, Taylor rate of the gradient > 1.9."""
    import scipy.optimize

    from oracle import fem, nonlinear

    om = meshes.regular_mesh(5)
    bc = om.boundary_vertices((1, 2, 3, 4))
    _, vol, G = fem.cell_geometry(om.coords, om.cells)
    load = fem.load_elements_p1(vol, om.cells, np.ones(om.num_vertices), 2)
    u, its = nonlinear.newton_solve(om.coords, om.cells, load, 1e2, bc, rtol=1e-9, atol=1e-10, max_iter=50)
    assert 2 <= its <= 10 and np.all(u[bc] == 0.0)
    # independent check: scipy's hybrid Powell solver on the free unknowns of the same discrete residual
    K = fem.assemble_matrix(fem.stiffness_elements(vol, G), om.cells, om.num_vertices)
    free = np.setdiff1d(np.arange(om.num_vertices), bc)

    def F(x):
        w = np.zeros(om.num_vertices)
        w[free] = x
        r = K @ w + fem.assemble_vector(nonlinear.cubic_load_elements(vol, om.cells, w, 2, 1e2) - load, om.cells, om.num_vertices)
        return r[free]

    sol = scipy.optimize.root(F, np.zeros(free.size), method="hybr", tol=1e-13)
    assert sol.success and np.allclose(u[free], sol.x, rtol=1e-5, atol=1e-8)
    with pytest.raises(RuntimeError, match="Maximum number of iterations"):
        nonlinear.newton_solve(om.coords, om.cells, load, 1e2, bc, rtol=1e-9, atol=1e-10, max_iter=1)

    om = meshes.regular_mesh(10)
    yd = np.sin(2 * np.pi * om.coords[:, 0]) * np.sin(2 * np.pi * om.coords[:, 1])
    prob = pipeline.PoissonControlProblem(om, yd, np.zeros(om.num_vertices), alpha=1e-6, bc_markers=(1, 2, 3, 4), cubic=1.0)
    assert pipeline.control_gradient_test(prob, rng=np.random.RandomState(300696))[0] > 1.9
    lin = pipeline.PoissonControlProblem(om, yd, np.zeros(om.num_vertices), alpha=1e-6, bc_markers=(1, 2, 3, 4))
    assert pipeline.control_gradient_test(lin, rng=np.random.RandomState(300696))[0] > 1.9


def test_curvature_regularization_known_answer():
    """tests/test_shape_optimization.py:528-546: on the reference's unit-disc fixture the mean-curvature vector of
    ``CurvatureRegularization._compute_curvature`` (shape_regularization.py:596-613) has mean norm 1 to 1e-3; the
    objective is then 1/2 int 1 ds = pi, and the oracle's derivative vector (``:554-575``) passes a Taylor test of the
    discrete functional with rate 2."""
    om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    prob = pipeline.PoissonShapeProblem(om, config=pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, factor_curvature=1.0))
    k = prob.compute_curvature()
    F, m = prob.boundary_facets(), prob._facet_measures()
    gp, gw = np.array([0.5 - np.sqrt(0.15), 0.5, 0.5 + np.sqrt(0.15)]), np.array([5.0, 8.0, 5.0]) / 18.0
    mean = sum(w * np.sum(m * np.linalg.norm((1 - t) * k[F[:, 0]] + t * k[F[:, 1]], axis=1)) for t, w in zip(gp, gw))
    assert abs(mean / m.sum() - 1.0) < 1e-3
    assert abs(prob.curvature_objective() - np.pi) < 5e-3
    interior = np.setdiff1d(np.arange(om.num_vertices), np.unique(F))
    assert np.max(np.abs(k[interior])) < 1e-14  # round-off of the normal projection of the opposite hat function
    rng = np.random.RandomState(300696)
    dJ = prob.curvature_derivative_vector()
    h = rng.rand(2 * om.num_vertices)
    J0, x0 = prob.curvature_objective(), prob.coords.copy()
    res = []
    eps = [1e-2 / 2**i for i in range(4)]
    for e in eps:
        prob.coords[:] = x0 + e * h.reshape(-1, 2)
        res.append(abs(prob.curvature_objective() - J0 - e * (dJ @ h)))
    prob.coords[:] = x0
    assert np.log(res[-1] / res[-2]) / np.log(0.5) > 1.9


def _p1_eval(coords, cells, values, point):
    """Value of a P1 function at a point (any cell containing it)."""
    x = np.asarray(point, dtype=float)
    T = coords[cells]
    M = np.transpose(T[:, 1:] - T[:, :1], (0, 2, 1))
    lam = np.linalg.solve(M, np.broadcast_to(x - T[:, 0], (len(cells), len(x)))[..., None])[..., 0]
    bary = np.concatenate([1.0 - lam.sum(axis=1, keepdims=True), lam], axis=1)
    c = int(np.argmax(bary.min(axis=1)))
    assert bary[c].min() > -1e-12
    return float(bary[c] @ values[cells[c]])


def test_distance_mu_known_answers():
    """tests/test_shape_optimization.py:707-771 (``use_distance_mu``, eikonal distance): on ``regular_mesh(64)`` with
    dist_min 0.1, dist_max 0.25, mu_min 1, mu_max 10 the Lame parameter is exactly 10 at the centre and at distance
    0.26 from the sides, exactly 1 at distance 0.05 and 0.09 -- the reference's ten point evaluations to 1e-10."""
    om = meshes.regular_mesh(64)
    cfg = pipeline.ShapeConfig(use_distance_mu=True, dist_min=0.1, dist_max=0.25, mu_min=1.0, mu_max=10.0)
    prob = pipeline.PoissonShapeProblem(om, config=cfg)
    want = {(0.5, 0.5): 10.0, (0.05, 0.5): 1.0, (0.09, 0.5): 1.0, (0.91, 0.5): 1.0, (0.5, 0.09): 1.0, (0.5, 0.91): 1.0,
            (0.5, 0.26): 10.0, (0.5, 0.74): 10.0, (0.26, 0.5): 10.0, (0.74, 0.5): 10.0}
    for pt, mu in want.items():
        assert abs(_p1_eval(om.coords, om.cells, prob.mu, pt) - mu) / mu < 1e-10, pt
    # a distance: zero on the boundary, close to the true distance near the walls, 1-Lipschitz-ish in the bulk
    d = prob.distance
    assert np.all(d[om.boundary_vertices((1, 2, 3, 4))] == 0.0) and d.min() >= 0.0
    near = np.minimum.reduce([om.coords[:, 0], 1 - om.coords[:, 0], om.coords[:, 1], 1 - om.coords[:, 1]])
    sel = near < 0.1
    assert np.max(np.abs(d[sel] - near[sel])) < 0.01
    # the blend itself (both variants) at hand-computed points
    dist = np.array([0.0, 0.1, 0.175, 0.25, 0.3])
    assert np.allclose(pipeline.distance_mu(dist, 0.1, 0.25, 1.0, 10.0), [1.0, 1.0, 5.5, 10.0, 10.0])
    sm = pipeline.distance_mu(dist, 0.1, 0.25, 1.0, 10.0, smooth=True)
    assert np.allclose(sm[[0, 1, 3, 4]], [1.0, 1.0, 10.0, 10.0]) and abs(sm[2] - 5.5) < 1e-12  # odd about the midpoint


def test_p_laplacian_reference_scenarios():
    """tests/test_p_laplacian.py:82-160: with ``use_p_laplacian`` (mu = 1, damping 1) gradient descent with p = 2
    converges to rtol 1e-2 within 22 iterations (oracle: 21), with p = 10 to rtol 1e-1 within 6 (oracle: 5), and the
    Taylor test -- the p-Laplace form as "scalar product" (shape_form_handler.py:761-770) -- has rate > 1.9."""
    base = dict(tol_lower=0.01, tol_upper=0.02, mu_def=1.0, mu_fix=1.0, damping_factor=1.0, use_p_laplacian=True)
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    prob = pipeline.PoissonShapeProblem(load(), config=pipeline.ShapeConfig(p_laplacian_power=2, **base))
    g = prob.compute_shape_gradient()
    assert prob.its["p_laplace_newton"] == [1]  # p = 2 is linear: one Newton step
    Fe, Je = prob.p_laplace_forms(g, 2)
    A, _ = fem.assemble_system(Je, None, prob.vcell_dofs, g.size)
    assert np.linalg.norm(A @ g - prob.dJ) <= 1e-12 * np.linalg.norm(prob.dJ)
    ok, hist = pipeline.gradient_descent(prob, rtol=1e-2, max_iter=22)
    assert ok and len(hist) - 1 == 21
    prob = pipeline.PoissonShapeProblem(load(), config=pipeline.ShapeConfig(p_laplacian_power=10, **base))
    ok, hist = pipeline.gradient_descent(prob, rtol=1e-1, max_iter=6)
    assert ok and len(hist) - 1 == 5 and hist[-1][2] <= 1e-1
    prob = pipeline.PoissonShapeProblem(load(), config=pipeline.ShapeConfig(p_laplacian_power=10, **base))
    assert pipeline.shape_gradient_test(prob)[0] > 1.9


def test_scaling_shape_regularization_known_answer():
    """tests/test_shape_optimization.py:800-831: with ``use_relative_scaling`` the initial cost of the four regularisation
    terms equals the sum of their weights (the reference asserts 1e-15) for five random weight sets."""
    weights = np.random.RandomState(300696).rand(5, 4)
    for w in weights:
        om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
        cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, use_relative_scaling=True, target_barycenter=(1.0, 1.0, 0.0),
                                   factor_volume=w[0], factor_surface=w[1], factor_curvature=w[2], factor_barycenter=w[3])
        prob = pipeline.PoissonShapeProblem(om, config=cfg, c_state=0.0)
        prob.solve_state()
        assert abs(prob.cost() - w.sum()) < 1e-15


def test_reextension_normal_reference_budget():
    """tests/test_shape_optimization.py:1084-1110 (``reextension_mode = normal``): L-BFGS reaches rtol 1e-2 within the
    reference's 11 iterations (oracle: 10); the normal projection reproduces a radial field on the disc's boundary and
    annihilates a tangential one."""
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, reextend_from_boundary=True, reextension_mode="normal")
    om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    prob = pipeline.PoissonShapeProblem(om, config=cfg)
    b = om.boundary_vertices((1,))
    w = prob.compute_normal_deformation(om.coords.reshape(-1).copy()).reshape(-1, 2)
    assert np.max(np.abs(w[b] - om.coords[b])) < 1e-3 and np.count_nonzero(np.delete(w, b, axis=0)) == 0
    tang = np.stack([-om.coords[:, 1], om.coords[:, 0]], axis=1).reshape(-1)
    assert np.max(np.abs(prob.compute_normal_deformation(tang))) < 1e-12
    ok, hist = pipeline.lbfgs(prob, rtol=1e-2, max_iter=11, memory=3)
    assert ok and len(hist) - 1 == 10


def test_geometric_cost_terms_and_regularization_demo_settings():
    """J = int u dx + alpha int 1 dx + alpha int 1 ds with volume + surface + curvature regularisation
    (demos/documented/shape_optimization/regularization: alpha = 0.1, factor_volume = factor_surface = 1, targets 1.5 /
    4.5, factor_curvature = 1e-4, damping 0.5): the shape gradient of the full functional passes the Taylor test."""
    om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    cfg = pipeline.ShapeConfig(lambda_lame=0.0, mu_def=1.0, mu_fix=1.0, damping_factor=0.5, factor_volume=1.0,
                               target_volume=1.5, factor_surface=1.0, target_surface=4.5, factor_curvature=1e-4)
    prob = pipeline.PoissonShapeProblem(om, config=cfg, c_volume=0.1, c_surface=0.1)
    prob.solve_state()
    plain = pipeline.PoissonShapeProblem(meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz")), config=cfg)
    plain.solve_state()
    assert abs(prob.cost() - plain.cost() - 0.1 * (prob.volume() + prob.surface())) < 1e-13
    assert pipeline.shape_gradient_test(prob)[0] > 1.9


def test_scalar_tracking_reference_scenarios():
    """tests/test_shape_optimization.py:549-592: J = 1/2 (int 1 dx - pi r^2)^2 -- Taylor rate > 1.9, L-BFGS to rtol 1e-6
    within 50 iterations (oracle: 8), radius r to 5e-3, 1/2 (vol - goal)^2 < 1e-10; J = 1/2 (int u^2 dx - goal)^2 --
    L-BFGS to rtol 1e-5 within 50 iterations (oracle: 10) leaves 1/2 (int u^2 - goal)^2 < 1e-14."""
    rng = np.random.RandomState(300696)
    radius = rng.uniform(0.33, 0.66)
    goal = np.pi * radius**2
    load = lambda: meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))  # noqa: E731
    cfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, volume_change=10.0)
    prob = pipeline.PoissonShapeProblem(load(), config=cfg, c_state=0.0, scalar_tracking=[dict(c_volume=1.0, goal=goal)])
    assert pipeline.shape_gradient_test(prob, rng=rng)[0] > 1.9
    ok, hist = pipeline.lbfgs(prob, rtol=1e-6, max_iter=50, memory=3)
    assert ok and len(hist) - 1 <= 50
    assert abs(prob.coords.max() - radius) < 5e-3 and abs(-prob.coords.min() - radius) < 5e-3
    assert 0.5 * (prob.volume() - goal) ** 2 < 1e-10
    goal2 = np.random.RandomState(300696).uniform(0.25, 0.75)
    prob = pipeline.PoissonShapeProblem(load(), config=pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02), c_state=0.0,
                                        scalar_tracking=[dict(c_track=2.0, goal=goal2)])
    ok, hist = pipeline.lbfgs(prob, rtol=1e-5, max_iter=50, memory=3)
    assert ok and 0.5 * (prob._tracking_values()[0] - goal2) ** 2 < 1e-14
