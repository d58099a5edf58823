"""GPU: the hot path at BASELINE configs[1] FULL size (unit cube, n = 119: 10.1 M tets, 1.73 M vertices,
25.6 M 3x3 blocks) checked through size-independent properties -- the oracle cannot run there in seconds:
constants in the kernel of K, partition of unity of M, symmetry and linearity of the SpMV, true residual of
the Krylov solutions (recomputed with cb_spmv), bitwise determinism, move -> revert round trip, analytic
values of the mesh measures on the undeformed Kuhn mesh."""

import numpy as np
import pytest
import torch

import cashocs_b200 as cb
from cashocs_b200._forms import forms
from cashocs_b200._utils import linalg

pytestmark = pytest.mark.gpu
N = 119


@pytest.fixture(scope="module")
def mesh():
    m = cb.regular_mesh(N, 1.0, 1.0, 1.0)
    m.build_pattern()
    return m


def test_full_size_counts_and_measures(mesh):
    assert mesh.num_cells == 6 * N**3 == 10110954 and mesh.num_vertices == (N + 1) ** 3
    assert mesh.nnz == 25575838  # P1 nnz of the n = 119 cube
    # every Kuhn tetrahedron has the condition number of tests/test_geometry.py:269 of the reference
    assert abs(cb.compute_mesh_quality(mesh, "min", "condition_number") - 0.3162277660168379) < 1e-12
    assert abs(cb.compute_mesh_quality(mesh, "avg", "condition_number") - 0.3162277660168379) < 1e-12
    tester = cb.APrioriMeshTester(mesh)
    V = torch.zeros(mesh.num_vertices * 3, dtype=torch.float64, device="cuda")
    assert tester.determinants(V) == (1.0, 1.0) and tester.test(V, float("inf"))
    # a uniform dilation x -> 1.5 x has det(I + grad V) = 1.5^3 in every cell
    mn, mx = tester.determinants((0.5 * mesh.coords).reshape(-1).contiguous())
    assert abs(mn - 3.375) < 1e-9 and abs(mx - 3.375) < 1e-9
    assert cb.IntersectionTester(mesh).test()


def test_full_size_matrices(mesh):
    nv = mesh.num_vertices
    one = torch.ones(nv, dtype=torch.float64, device="cuda")
    K, _ = linalg.assemble_scalar_system(mesh, 1.0, 0.0, rhs=False)
    M, _ = linalg.assemble_scalar_system(mesh, 0.0, 1.0, rhs=False)
    assert float(K.mult(one).abs().max()) < 1e-9 * float(K.val.abs().max())  # constants are in the kernel of K
    assert abs(float(M.mult(one).sum()) - 1.0) < 1e-12                       # sum of M = |Omega| = 1
    x = torch.from_numpy(np.random.RandomState(1).rand(nv)).cuda()
    y = torch.from_numpy(np.random.RandomState(2).rand(nv)).cuda()
    A, _ = linalg.assemble_scalar_system(mesh, 1.0, 0.3, rhs=False)
    lin = A.mult(2.0 * x - 3.0 * y) - (2.0 * A.mult(x) - 3.0 * A.mult(y))
    assert float(lin.abs().max()) < 1e-10 * float(A.mult(x).abs().max())
    assert abs(A.scalar_product(x, y) - A.scalar_product(y, x)) < 1e-12 * abs(A.scalar_product(x, y))
    A2, _ = linalg.assemble_scalar_system(mesh, 1.0, 0.3, rhs=False)
    assert torch.equal(A.val, A2.val) and torch.equal(A.mult(x), A2.mult(x))  # bitwise reproducible
    # elasticity: rigid translations see only the damping term: E t = delta * M t
    mu = torch.full((nv,), 0.35714285714285715, dtype=torch.float64, device="cuda")
    E = linalg.assemble_elasticity_matrix(mesh, mu, 1.428571428571429, 0.2)
    t = torch.zeros(nv * 3, dtype=torch.float64, device="cuda")
    t[1::3] = 1.0
    Et = E.mult(t).view(-1, 3)
    ref = 0.2 * M.mult(one)
    assert float((Et[:, 1] - ref).abs().max()) < 1e-9 * float(E.val.abs().max())
    assert float(Et[:, 0].abs().max()) < 1e-9 * float(E.val.abs().max())
    u = torch.from_numpy(np.random.RandomState(3).rand(nv * 3)).cuda()
    v = torch.from_numpy(np.random.RandomState(4).rand(nv * 3)).cuda()
    assert abs(E.scalar_product(u, v) - E.scalar_product(v, u)) < 1e-12 * abs(E.scalar_product(u, v))


def test_full_size_solves_have_small_true_residuals(mesh):
    nv = mesh.num_vertices
    x = cb.spatial_coordinate(3)
    f = 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1
    flag, val = forms.bc_flag_arrays(mesh, forms.create_dirichlet_bcs(0.0, [1, 2, 3, 4, 5, 6]), 1)
    A, b = linalg.assemble_scalar_system(mesh, 1.0, 0.0, f_poly=f, bc_flag=flag, bc_val=val)
    ksp = {"ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg", "ksp_rtol": 1e-10, "ksp_atol": 1e-30,
           "ksp_max_it": 500}
    u = torch.zeros(nv, dtype=torch.float64, device="cuda")
    res = cb.LinearSolver().solve(u, A=A, b=b, ksp_options=ksp)
    assert res.reason == 2 and res.its < 60
    # true residual, recomputed with cb_spmv (the KSP test is on the preconditioned norm, like PETSc's default)
    assert float((A.mult(u) - b).norm() / b.norm()) < 1e-7
    assert float(u[mesh.boundary_vertices([1, 2, 3, 4, 5, 6])].abs().max()) < 1e-12 * float(u.abs().max())
    u2 = torch.zeros_like(u)
    res2 = cb.LinearSolver().solve(u2, A=A, b=b, ksp_options=ksp)
    assert res2.its == res.its and torch.equal(u, u2)                # bitwise reproducible
    mu = torch.full((nv,), 0.35714285714285715, dtype=torch.float64, device="cuda")
    E = linalg.assemble_elasticity_matrix(mesh, mu, 1.428571428571429, 0.2)
    rhs = torch.from_numpy(np.random.RandomState(5).rand(nv * 3) - 0.5).cuda()
    g = torch.zeros(nv * 3, dtype=torch.float64, device="cuda")
    res = cb.LinearSolver().solve(g, A=E, b=rhs, ksp_options=ksp)
    assert res.reason == 2 and res.its < 250
    # the KSP test is on the preconditioned norm (PETSc's default): with the rigid-body coarse correction in the
    # preconditioner |M b| is about twice as large, the true residual at convergence about three times (CPU oracle)
    assert float((E.mult(g) - rhs).norm() / rhs.norm()) < 5e-7


def test_full_size_move_revert_round_trip(mesh):
    handler = cb.DeformationHandler(mesh, cb.APrioriMeshTester(mesh), cb.IntersectionTester(mesh))
    before = mesh.coords.clone()
    h = 1.0 / N
    V = torch.from_numpy(0.2 * h * (2 * np.random.RandomState(300696).rand(mesh.num_vertices * 3) - 1)).cuda()
    assert handler.move_mesh(V)
    assert float((mesh.coords - before - V.view(-1, 3)).abs().max()) < 1e-15
    q = cb.compute_mesh_quality(mesh, "min", "skewness")
    assert 0.0 < q < 1.0
    handler.revert_transformation()
    assert torch.equal(mesh.coords, before)
    W = torch.from_numpy(np.random.RandomState(1).uniform(-1e3, 1e3, size=mesh.num_vertices * 3)).cuda()
    assert not handler.move_mesh(W) and torch.equal(mesh.coords, before)
