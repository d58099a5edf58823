"""GPU parity of the P2 (quadratic Lagrange) state / adjoint path (BASELINE configs[4]) against the
CPU oracle ``oracle/p2.py``: dof tables and CSR pattern bit-exact, assembled values <= 1e-12
relative, shape derivative / functional <= 1e-12, pipeline gradients <= 1e-8."""

import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla
import torch

import cashocs_b200 as cb
from cashocs_b200._forms import forms
from cashocs_b200._utils import linalg
from oracle import fem, meshes, p2, pipeline

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KINDS = ["sq8", "cube4", "circle", "mesh3d"]


def demo_poly(dim):
    x = cb.spatial_coordinate(dim)
    return 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1


def both(kind):
    if kind == "sq8":
        return meshes.regular_mesh(8), cb.regular_mesh(8)
    if kind == "cube4":
        return meshes.regular_mesh(4, 1.0, 1.0, 1.0), cb.regular_mesh(4, 1.0, 1.0, 1.0)
    name = {"circle": "unit_circle.npz", "mesh3d": "mesh3d.npz"}[kind]
    return meshes.load_npz(os.path.join(GOLDEN, name)), cb.import_mesh(os.path.join(GOLDEN, name))


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("kind", KINDS)
def test_p2_dofs_and_pattern_bit_exact(kind):
    om, dm = both(kind)
    sp = dm.p2_space()
    dofs, edges = p2.cell_dofs(om.cells, om.num_vertices)
    assert np.array_equal(sp.edges.cpu().numpy(), edges)
    assert np.array_equal(sp.cell_dofs.cpu().numpy(), dofs)
    rowptr, col = fem.csr_pattern(dofs, sp.ndofs)
    assert np.array_equal(sp.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(sp.col.cpu().numpy(), col)
    markers = [int(t) for t in np.unique(om.facet_tags)]
    for mk in (markers, markers[:1]):
        assert np.array_equal(np.sort(sp.boundary_dofs(mk).cpu().numpy()), np.sort(p2.boundary_dofs(om, mk, edges)))


@pytest.mark.parametrize("kind", KINDS)
def test_p2_system_matches_oracle(kind):
    om, dm = both(kind)
    d = om.dim
    sp = dm.p2_space()
    dofs, edges = p2.cell_dofs(om.cells, om.num_vertices)
    nd = sp.ndofs
    markers = [int(t) for t in np.unique(om.facet_tags)]
    bc = p2.boundary_dofs(om, markers[:1], edges)
    g = 0.7
    _, vol, G = fem.cell_geometry(om.coords, om.cells)
    Ke = 1.3 * p2.stiffness_elements(vol, G) + 0.4 * p2.mass_elements(vol, d)
    fn = np.random.RandomState(1).rand(nd)
    be = (0.9 * p2.load_elements_callable(om.coords, om.cells, vol, pipeline.demo_f, 4)
          - 2.0 * np.einsum("cab,cb->ca", p2.mass_elements(vol, d), fn[dofs])
          + 0.25 * vol[:, None] * p2.reference_tensors(d)[2][None])
    A0, b0 = fem.assemble_system(Ke, be, dofs, nd, bc, g)
    flag, val = forms.bc_flag_arrays(dm, [cb.DirichletBC(g, tuple(markers[:1]))], 1, space=sp)
    A, b = linalg.assemble_p2_system(sp, 1.3, 0.4, f_poly=demo_poly(d), f_poly_scale=0.9, f_const=0.25,
                                     f_nodal=torch.from_numpy(fn).cuda(), f_nodal_scale=-2.0, bc_flag=flag, bc_val=val)
    As = A.to_scipy()
    assert np.array_equal(As.indptr, A0.indptr) and np.array_equal(As.indices, A0.indices)
    assert rel(As.data, A0.data) < 1e-12
    assert rel(b.cpu().numpy(), b0) < 1e-12
    assert abs(As - As.T).max() == 0.0
    A2, b2 = linalg.assemble_p2_system(sp, 1.3, 0.4, f_poly=demo_poly(d), f_poly_scale=0.9, f_const=0.25,
                                       f_nodal=torch.from_numpy(fn).cuda(), f_nodal_scale=-2.0, bc_flag=flag, bc_val=val)
    assert torch.equal(A.val, A2.val) and torch.equal(b, b2)  # bit-reproducible
    # the Krylov / AMG solvers work on the P2 pattern
    x = torch.zeros(nd, dtype=torch.float64, device="cuda")
    xd = spla.spsolve(A0.tocsc(), b0)
    for pc in ("jacobi", "hypre"):
        res = cb.LinearSolver().solve(x, A=A, b=b, ksp_options={"ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-12,
                                                                 "ksp_atol": 1e-30, "ksp_max_it": 5000})
        assert res.reason > 0 and rel(x.cpu().numpy(), xd) < 1e-8


@pytest.mark.parametrize("kind", KINDS)
def test_p2_shape_derivative_and_functional_match_oracle(kind):
    om, dm = both(kind)
    d = om.dim
    sp = dm.p2_space()
    dofs, edges = p2.cell_dofs(om.cells, om.num_vertices)
    rng = np.random.RandomState(3)
    u, p = rng.rand(sp.ndofs), rng.rand(sp.ndofs)
    ud = rng.rand(sp.ndofs)
    be = p2.shape_derivative_poisson(om.coords, om.cells, dofs, u, p, pipeline.demo_f, pipeline.demo_grad_f, 4, 0.8)
    b0 = fem.assemble_vector(be, fem.vector_cell_dofs(om.cells, d), d * om.num_vertices)
    b = linalg.assemble_p2_shape_derivative(sp, torch.from_numpy(u).cuda(), torch.from_numpy(p).cuda(), demo_poly(d), 0.8)
    assert rel(b.cpu().numpy(), b0) < 1e-12
    _, vol, _ = fem.cell_geometry(om.coords, om.cells)
    J0 = 0.8 * p2.integral(u, dofs, vol, d) + 0.5 * 0.3 * float(
        np.sum(np.einsum("ca,cab,cb->c", (u - ud)[dofs], p2.mass_elements(vol, d), (u - ud)[dofs])))
    yd = cb.Function(dm, degree=2)
    yd.vector.copy_(torch.from_numpy(ud))
    y = cb.Function(dm, degree=2)
    y.vector.copy_(torch.from_numpy(u))
    from cashocs_b200._optimization import cost_functional

    class _Solved:
        def solve(self):
            return None

    rcf = cost_functional.ReducedCostFunctional(dm, cb.IntegralFunctional(c_state=0.8, c_track=0.3, desired_state=yd),
                                                [y], _Solved())
    assert abs(rcf.evaluate() - J0) < 1e-12 * max(abs(J0), 1.0)


@pytest.mark.parametrize("kind", ["sq8", "cube4", "circle"])
def test_p2_shape_gradient_pipeline(kind):
    om, dm = both(kind)
    d = om.dim
    markers = tuple(int(t) for t in np.unique(om.facet_tags))
    cfg = cb.Config()
    cfg.set("StateSystem", "is_linear", "True")
    cfg.set("ShapeGradient", "shape_bdry_def", str(list(markers)))
    cfg.set("ShapeGradient", "lambda_lame", "1.428571428571429")
    cfg.set("ShapeGradient", "damping_factor", "0.2")
    cfg.set("ShapeGradient", "mu_fix", "0.35714285714285715")
    cfg.set("ShapeGradient", "mu_def", "0.35714285714285715")
    ksp = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-13, "ksp_atol": 1e-30, "ksp_max_it": 20000}
    u, p = cb.Function(dm, degree=2), cb.Function(dm, degree=2)
    sop = cb.ShapeOptimizationProblem(cb.PoissonForm(source=demo_poly(d)), cb.create_dirichlet_bcs(0.0, list(markers)),
                                      cb.IntegralFunctional(c_state=1.0), u, p, config=cfg, ksp_options=ksp,
                                      adjoint_ksp_options=ksp, gradient_ksp_options=ksp)
    g = sop.compute_shape_gradient()[0].vector.cpu().numpy()
    ocfg = pipeline.ShapeConfig(shape_bdry_def=markers, test_for_intersections=False)
    prob = pipeline.PoissonShapeProblem(om, bc_markers=markers, config=ocfg, state_degree=2)
    g0 = prob.compute_shape_gradient()
    assert rel(u.vector.cpu().numpy(), prob.u) < 1e-8
    assert rel(p.vector.cpu().numpy(), prob.p) < 1e-8
    assert rel(g, g0) < 1e-8
    assert abs(sop.reduced_cost_functional.evaluate() - prob.cost()) < 1e-10
    assert abs(sop.gradient_problem.gradient_norm_squared - prob.gradient_norm_squared) < 1e-8 * prob.gradient_norm_squared
    rate = sop.gradient_test()
    assert rate > 1.9
