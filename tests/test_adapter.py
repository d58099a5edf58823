"""``cashocs_b200/fenics_adapter.py``: the numerical cores of the reference-side adapter without FEniCS -- PETSc
AIJ -> block CSR, numerical recognition of the supported form family from element tensors (the oracle's element
tensors stand in for ``fenics.assemble_local``), polynomial fitting of a coordinate expression -- and, on a GPU, the
``LinearSolver`` plug-in driven through duck-typed PETSc / dolfin stand-ins."""

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from cashocs_b200 import fenics_adapter as fa
from oracle import fem, meshes


def test_aij_to_baij_round_trip():
    rng = np.random.RandomState(0)
    om = meshes.regular_mesh(3, 1.0, 1.0, 1.0)
    rowptr, col = fem.csr_pattern(om.cells, om.num_vertices)
    nnzb = col.size
    blocks = rng.rand(nnzb, 3, 3)
    B = sp.bsr_matrix((blocks, col, rowptr), shape=(3 * om.num_vertices, 3 * om.num_vertices))
    A = B.tocsr()
    A.sort_indices()
    rp, cl, vl = fa.aij_to_baij(A.indptr, A.indices, A.data, 3)
    assert np.array_equal(rp, rowptr) and np.array_equal(cl, col) and np.array_equal(vl, blocks)
    rp1, cl1, vl1 = fa.aij_to_baij(A.indptr, A.indices, A.data, 1)
    assert np.array_equal(rp1, A.indptr) and np.array_equal(vl1, A.data)
    broken = A.tolil()
    broken[0, 4] = 0.0
    broken = sp.csr_matrix(broken)
    broken.eliminate_zeros()
    with pytest.raises(Exception):
        fa.aij_to_baij(broken.indptr, broken.indices, broken.data, 3)


@pytest.mark.parametrize("dim", [2, 3])
def test_recognise_bilinear_and_linear_forms(dim):
    om = meshes.regular_mesh(3) if dim == 2 else meshes.regular_mesh(2, 1.0, 1.0, 1.0)
    rng = np.random.RandomState(1)
    coords = om.coords + 0.02 * rng.rand(*om.coords.shape)
    cells = om.cells[:6]
    _, vol, G = fem.cell_geometry(coords, om.cells)
    K, M = fem.stiffness_elements(vol, G)[:6], fem.mass_elements(vol, dim)[:6]
    X = coords[cells]
    assert fa.recognise_bilinear(1.7 * K + 0.8 * M, X) == pytest.approx((1.7, 0.8))
    assert fa.recognise_bilinear(K, X) == pytest.approx((1.0, 0.0))
    # outside the family: an anisotropic diffusion tensor
    D = np.diag([1.0, 3.0, 2.0][:dim])
    aniso = vol[:6, None, None] * np.einsum("cai,ij,cbj->cab", G[:6], D, G[:6])
    assert fa.recognise_bilinear(aniso, X) is None
    # linear forms: polynomial source and a nodal (P1) source
    f = lambda p: 2.5 * (p[:, 0] + 0.4 - p[:, 1] ** 2) ** 2 + p[:, 0] ** 2 + p[:, 1] ** 2 - 1  # noqa: E731
    be = fem.load_elements_callable(coords, om.cells, vol, lambda x: f(x.reshape(-1, dim)).reshape(x.shape[:-1]), 5)[:6]
    s_p, s_k = fa.recognise_linear(-3.0 * be, X, polynomial=f)
    assert s_p == pytest.approx(-3.0) and s_k == []
    w = rng.rand(om.num_vertices)
    bn = fem.load_elements_p1(vol, om.cells, w, dim)[:6]
    s_p, s_k = fa.recognise_linear(2.0 * bn, X, nodal_fields=[w[cells]])
    assert s_p == 0.0 and s_k == pytest.approx([2.0])
    assert fa.recognise_linear(bn ** 2, X, nodal_fields=[w[cells]]) is None


@pytest.mark.parametrize("dim", [2, 3])
def test_recognise_semilinear_forms(dim):
    """The residual of ``grad y . grad p + 1e2 y^3 p - u p`` (demos/documented/optimal_control/nonlinear_pdes) tabulated
    at a random P1 state is recognised with its coefficients; a quadratic nonlinearity is refused."""
    from oracle import nonlinear

    om = meshes.regular_mesh(3) if dim == 2 else meshes.regular_mesh(2, 1.0, 1.0, 1.0)
    rng = np.random.RandomState(2)
    coords = om.coords + 0.02 * rng.rand(*om.coords.shape)
    _, vol, G = fem.cell_geometry(coords, om.cells)
    y, u = rng.rand(om.num_vertices) + 0.5, rng.rand(om.num_vertices)
    K, M = fem.stiffness_elements(vol, G), fem.mass_elements(vol, dim)
    cubic = nonlinear.cubic_load_elements(vol, om.cells, y, dim, 1e2)
    Fe = np.einsum("cab,cb->ca", K, y[om.cells]) + cubic - np.einsum("cab,cb->ca", M, u[om.cells])
    sel = slice(0, 8)
    X = coords[om.cells[sel]]
    got = fa.recognise_semilinear(Fe[sel], X, y[om.cells[sel]], u[om.cells[sel]])
    assert got == pytest.approx((1.0, 0.0, 1e2, 1.0), rel=1e-8, abs=1e-8)
    lin = 2.0 * np.einsum("cab,cb->ca", K, y[om.cells]) + 0.3 * np.einsum("cab,cb->ca", M, y[om.cells])
    assert fa.recognise_semilinear(lin[sel], X, y[om.cells[sel]]) == pytest.approx((2.0, 0.3, 0.0, 0.0), abs=1e-8)
    lam, w = fem.quadrature(dim, 4)
    quad = vol[:, None] * np.einsum("cq,q,qa->ca", (y[om.cells] @ lam.T) ** 2, w, lam)
    assert fa.recognise_semilinear((Fe + quad)[sel], X, y[om.cells[sel]], u[om.cells[sel]]) is None
    assert fa.recognise_semilinear(Fe[sel] * 0.0, X, 0.0 * y[om.cells[sel]]) is None  # a zero state tells nothing


def test_fit_polynomial_recovers_the_demo_source():
    f = lambda p: 2.5 * (p[:, 0] + 0.4 - p[:, 1] ** 2) ** 2 + p[:, 0] ** 2 + p[:, 1] ** 2 - 1  # noqa: E731
    poly = fa.fit_polynomial(f, 2)
    want = {(2, 0, 0): 3.5, (1, 0, 0): 2.0, (1, 2, 0): -5.0, (0, 4, 0): 2.5, (0, 2, 0): -1.0, (0, 0, 0): -0.6}
    assert set(poly.terms) == set(want)
    for k, v in want.items():
        assert poly.terms[k] == pytest.approx(v, rel=1e-9)
    assert fa.fit_polynomial(lambda p: np.sin(p[:, 0]), 2) is None
    g = fa.fit_polynomial(lambda p: p[:, 0] * p[:, 1] * p[:, 2] - 2.0, 3)
    assert {k: round(v, 9) for k, v in g.terms.items()} == {(1, 1, 1): 1.0, (0, 0, 0): -2.0}


class _Vec:
    def __init__(self, a):
        self.a = np.array(a, dtype=float)

    def getArray(self):
        return self.a

    def setArray(self, v):
        self.a[:] = v

    def set(self, v):
        self.a[:] = v


class _Mat:
    def __init__(self, A, bs):
        self.A, self.bs = sp.csr_matrix(A), bs
        self.A.sort_indices()

    def getBlockSize(self):
        return self.bs

    def getValuesCSR(self):
        return self.A.indptr, self.A.indices, self.A.data


class _Function:
    def __init__(self, n):
        self._v = _Vec(np.full(n, 7.0))
        self.applied = 0

    def vector(self):
        return self

    def vec(self):
        return self._v

    def apply(self, mode):
        self.applied += 1


@pytest.mark.gpu
def test_linear_solver_plug_in_with_petsc_like_objects():
    """``B200LinearSolver.solve(function, A, b, ksp_options, rtol, atol)`` -- the reference's documented plug-in point --
    fed with objects that expose petsc4py's / dolfin's methods: scalar Poisson (cg + hypre) and the vertex-blocked
    elasticity operator with node coordinates (rigid-body correction on), both against a direct solve; ``b is None``
    zeroes the function; a negative reason raises ``PETScKSPError``."""
    import cashocs_b200 as cb
    from oracle import pipeline

    om = meshes.regular_mesh(6, 1.0, 1.0, 1.0)
    prob = pipeline.PoissonShapeProblem(om, bc_markers=(1, 2, 3, 4, 5, 6), config=pipeline.ShapeConfig(shape_bdry_def=(1, 2, 3, 4, 5, 6)))
    prob.solve_state()
    ksp = {"ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg", "ksp_rtol": 1e-12, "ksp_atol": 1e-30}
    fun = _Function(om.num_vertices)
    fa.B200LinearSolver().solve(fun, _Mat(prob.A_state, 1), _Vec(prob.b_state), ksp)
    want = spla.spsolve(sp.csc_matrix(prob.A_state), prob.b_state)
    assert np.max(np.abs(fun.vec().a - want)) <= 1e-9 * np.max(np.abs(want)) and fun.applied == 1
    rhs = np.random.RandomState(2).rand(3 * om.num_vertices) - 0.5
    fun3 = _Function(3 * om.num_vertices)
    solver = fa.B200LinearSolver(node_coordinates=om.coords)
    solver.solve(fun3, _Mat(prob.A_elas, 3), _Vec(rhs), ksp)
    want3 = spla.spsolve(sp.csc_matrix(prob.A_elas), rhs)
    assert np.max(np.abs(fun3.vec().a - want3)) <= 1e-8 * np.max(np.abs(want3))
    no_rbm = fa.B200LinearSolver()
    no_rbm.solve(_Function(3 * om.num_vertices), _Mat(prob.A_elas, 3), _Vec(rhs), ksp)
    assert solver.last_result.its <= no_rbm.last_result.its
    fa.B200LinearSolver().solve(fun, _Mat(prob.A_state, 1), None, ksp)
    assert np.all(fun.vec().a == 0.0)
    with pytest.raises(cb.PETScKSPError):
        fa.B200LinearSolver().solve(fun, _Mat(prob.A_state, 1), _Vec(prob.b_state), dict(ksp, ksp_max_it=1))
