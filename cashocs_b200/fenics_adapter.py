"""Adapter between a FEniCS / PETSc based cashocs and the device library (``INTEGRATION.md`` sections 1-3 as code).

The reference derives every system symbolically with UFL and hands dolfin forms, ``PETSc.Mat`` and ``PETSc.Vec``
around (``cashocs/_utils/linalg.py:107-188,480-546``, ``cashocs/_forms/general_form_handler.py:101-140,170-244``).
This module is the glue a cashocs maintainer installs on the *reference* side so that user scripts -- UFL forms,
``.ini`` files, ``ksp_options`` dicts -- stay unchanged:

* :class:`B200LinearSolver` -- the documented plug-in point (``linear_solver=`` kwarg; example
  ``demos/undocumented/shape_optimization/python_pc/python_pc.py:48-108``): a PETSc AIJ matrix and right-hand side in,
  the solution written into ``function.vector()``, ``PETScKSPError`` on a negative reason.
* :func:`recognise_bilinear` / :func:`recognise_linear` / :func:`recognise_semilinear` / :func:`fit_polynomial` -- which member of the supported
  form family a UFL form is, decided *numerically*: dolfin tabulates the element tensor of the user's form on a few
  cells (``fenics.assemble_local``), the adapter fits ``kappa K_e + c M_e`` (resp. a polynomial / nodal source) to
  it and accepts the form only if the fit is exact to round-off.  No UFL tree matching, so any way of writing the
  same form is recognised and anything outside the family falls through to the stock dolfin path.
* :func:`device_mesh_from_dolfin`, :func:`install` -- mesh hand-over and the monkey-patches of section 2 / 3.

FEniCS, petsc4py and cashocs are imported lazily: the numerical cores (everything that does arithmetic) work on
numpy arrays and duck-typed PETSc-like objects and are tested without them (``tests/test_adapter.py``); only the
few lines that call dolfin need it.  Nothing here is on the hot path of ``bench.py`` -- the native classes of
``cashocs_b200`` are.
"""

from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from cashocs_b200 import _exceptions

__all__ = ["aij_to_baij", "fit_polynomial", "recognise_bilinear", "recognise_linear", "recognise_semilinear", "B200LinearSolver",
           "device_mesh_from_dolfin", "install"]


# ----------------------------------------------------------------------------- PETSc AIJ -> block CSR
def aij_to_baij(rowptr, col, val, bs: int):
    """Scalar CSR of a vertex-blocked vector operator (dof = bs * node + component, PETSc AIJ with sorted columns)
    -> block CSR ``(rowptr_b [nb+1], col_b [nnzb], val_b [nnzb, bs, bs])`` as ``cb_mat`` wants it.

    dolfin assembles the full cell pattern, so every block is dense; a row whose pattern is not a union of complete
    blocks raises (the operator is then not vertex-blocked and has to stay scalar)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.float64)
    n = rowptr.size - 1
    if bs == 1:
        return rowptr, col.astype(np.int32), val
    if n % bs:
        raise _exceptions.InputError("fenics_adapter.aij_to_baij", "bs", "matrix size is not a multiple of the block size")
    nb = n // bs
    first = rowptr[:-1:bs]                       # first scalar row of every block row
    lens = np.diff(rowptr)
    if np.any(lens.reshape(nb, bs) != lens[::bs, None]) or np.any(lens % bs):
        raise _exceptions.InputError("fenics_adapter.aij_to_baij", "A", "rows of one block have different patterns")
    nnzb_row = lens[::bs] // bs
    rowptr_b = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum(nnzb_row, out=rowptr_b[1:])
    nnzb = int(rowptr_b[-1])
    # block columns from the first scalar row of every block row: every bs-th entry
    take = np.repeat(first, nnzb_row) + (np.arange(nnzb) - np.repeat(rowptr_b[:-1], nnzb_row)) * bs
    if np.any(col[take] % bs):
        raise _exceptions.InputError("fenics_adapter.aij_to_baij", "A", "pattern is not made of complete blocks")
    col_b = (col[take] // bs).astype(np.int32)
    val_b = np.empty((nnzb, bs, bs))
    blk_row = np.repeat(np.arange(nb), nnzb_row)
    for i in range(bs):
        start = rowptr[blk_row * bs + i] + (np.arange(nnzb) - rowptr_b[blk_row]) * bs
        for j in range(bs):
            if np.any(col[start + j] != col_b.astype(np.int64) * bs + j):
                raise _exceptions.InputError("fenics_adapter.aij_to_baij", "A", "pattern is not made of complete blocks")
            val_b[:, i, j] = val[start + j]
    return rowptr_b, col_b, val_b


# ----------------------------------------------------------------------------- numerical form recognition
def _p1_reference_tensors(x):
    """Stiffness and mass element tensors of the P1 simplex with vertices ``x [d+1, d]``."""
    d = x.shape[1]
    J = (x[1:] - x[0]).T
    vol = abs(np.linalg.det(J)) / float(np.prod(np.arange(1, d + 1)))
    G = np.zeros((d + 1, d))
    G[1:] = np.linalg.inv(J)
    G[0] = -G[1:].sum(axis=0)
    K = vol * G @ G.T
    M = vol * (np.ones((d + 1, d + 1)) + np.eye(d + 1)) / ((d + 1) * (d + 2))
    return K, M, vol


def recognise_bilinear(element_tensors, cell_vertices, tol: float = 1e-10):
    """Is ``a(u, v)`` of the family ``kappa grad u . grad v + c u v`` on scalar P1?

    ``element_tensors [ncell, d+1, d+1]``: the form's element matrices on a few cells (``fenics.assemble_local``, rows
    / columns in vertex order), ``cell_vertices [ncell, d+1, d]`` their vertex coordinates.  Returns ``(kappa, c)`` when
    ``A_e = kappa K_e + c M_e`` holds on every sampled cell to ``tol`` (relative), else ``None``."""
    A = np.asarray(element_tensors, dtype=float)
    X = np.asarray(cell_vertices, dtype=float)
    rows, rhs = [], []
    for Ae, x in zip(A, X):
        K, M, _ = _p1_reference_tensors(x)
        rows.append(np.stack([K.reshape(-1), M.reshape(-1)], axis=1))
        rhs.append(Ae.reshape(-1))
    B, y = np.concatenate(rows), np.concatenate(rhs)
    coef, *_ = np.linalg.lstsq(B, y, rcond=None)
    if np.linalg.norm(B @ coef - y) > tol * max(np.linalg.norm(y), 1e-300):
        return None
    kappa, c = float(coef[0]), float(coef[1])
    scale = max(abs(kappa), abs(c), 1e-300)
    return (0.0 if abs(kappa) < tol * scale else kappa), (0.0 if abs(c) < tol * scale else c)


def fit_polynomial(f, dim: int, max_degree: int = 8, tol: float = 1e-9, rng=None):
    """The polynomial in the spatial coordinates that a callable ``f(points [n, dim]) -> [n]`` is (a UFL expression in
    ``SpatialCoordinate`` is evaluated point-wise by calling it), or ``None`` if it is not a polynomial of degree <=
    ``max_degree``: least-squares fit of increasing degree on random points, accepted once a second, independent point
    set is reproduced to ``tol``.  Coefficients below round-off of the largest one are dropped."""
    from cashocs_b200._forms import expressions

    rng = rng or np.random.RandomState(20240229)
    for degree in range(0, max_degree + 1):
        exps = [e for e in itertools.product(range(degree + 1), repeat=dim) if sum(e) <= degree]
        n = 4 * len(exps) + 8

        def vandermonde(p):
            return np.stack([np.prod(p ** np.asarray(e)[None, :], axis=1) for e in exps], axis=1)

        p_fit, p_chk = rng.rand(n, dim) * 2.0 - 0.5, rng.rand(n, dim) * 2.0 - 0.5
        y_fit, y_chk = np.asarray(f(p_fit), dtype=float).reshape(-1), np.asarray(f(p_chk), dtype=float).reshape(-1)
        coef, *_ = np.linalg.lstsq(vandermonde(p_fit), y_fit, rcond=None)
        scale = max(np.max(np.abs(y_chk)), 1e-300)
        if np.max(np.abs(vandermonde(p_chk) @ coef - y_chk)) <= tol * scale:
            cmax = max(np.max(np.abs(coef)), 1e-300)
            terms = {}
            for e, c in zip(exps, coef):
                if abs(c) > 1e-11 * cmax:
                    terms[tuple(list(e) + [0] * (3 - dim))] = float(c)
            return expressions.Polynomial(terms)
    return None


def recognise_linear(element_vectors, cell_vertices, nodal_fields=(), polynomial=None, tol: float = 1e-9):
    """Is ``L(v)`` of the family ``s_p int f v + sum_k s_k int w_k v`` (f a known polynomial, w_k P1 fields)?

    ``element_vectors [ncell, d+1]``: the form's element vectors on a few cells; ``nodal_fields``: for every candidate
    P1 coefficient its vertex values on those cells ``[ncell, d+1]``; ``polynomial``: a callable ``f(points)`` or
    ``None``.  Returns the scale factors ``(s_p, [s_k])`` if the fit is exact to ``tol``, else ``None``."""
    from cashocs_b200._forms import expressions

    b = np.asarray(element_vectors, dtype=float)
    X = np.asarray(cell_vertices, dtype=float)
    d = X.shape[2]
    cols = []
    if polynomial is not None:
        lam, w = expressions.simplex_quadrature(d, 10)
        lam = lam[:, : d + 1]
        col = np.zeros_like(b)
        for c, x in enumerate(X):
            _, _, vol = _p1_reference_tensors(x)
            fq = np.asarray(polynomial(lam @ x), dtype=float).reshape(-1)
            col[c] = vol * (w * fq) @ lam
        cols.append(col.reshape(-1))
    for field in nodal_fields:
        col = np.zeros_like(b)
        for c, x in enumerate(X):
            _, M, _ = _p1_reference_tensors(x)
            col[c] = M @ np.asarray(field[c], dtype=float)
        cols.append(col.reshape(-1))
    if not cols:
        return None if np.linalg.norm(b) > tol else (0.0, [])
    B = np.stack(cols, axis=1)
    y = b.reshape(-1)
    coef, *_ = np.linalg.lstsq(B, y, rcond=None)
    if np.linalg.norm(B @ coef - y) > tol * max(np.linalg.norm(y), 1e-300):
        return None
    coef = [float(c) for c in coef]
    return (coef[0], coef[1:]) if polynomial is not None else (0.0, coef)


def recognise_semilinear(residual_vectors, cell_vertices, state_values, source_values=None, tol: float = 1e-9):
    """Is the residual ``F(u; v)`` of the semilinear family ``kappa grad u . grad v + c u v + gamma u^3 v - s w v``
    (``PoissonForm`` with ``cubic``; the equation of ``tests/test_nonlinear_solvers.py:28-61`` and of the
    ``nonlinear_pdes`` demo)?

    ``residual_vectors [ncell, d+1]``: the form's element vectors on a few cells, tabulated by dolfin at the P1 state
    with the vertex values ``state_values [ncell, d+1]`` (``fenics.assemble_local`` of the nonlinear form);
    ``source_values``: vertex values of the P1 source ``w`` (the control) on those cells, or ``None``.  The candidate
    columns are ``K_e u_e``, ``M_e u_e``, ``int u^3 phi`` (degree-4 rule: exact for P1 ``u``) and ``M_e w_e``; returns
    ``(kappa, c, gamma, s)`` when the fit is exact to ``tol``, else ``None`` -- the state must not be (close to) zero,
    otherwise the columns cannot be told apart."""
    from cashocs_b200._forms import expressions

    F = np.asarray(residual_vectors, dtype=float)
    X = np.asarray(cell_vertices, dtype=float)
    U = np.asarray(state_values, dtype=float)
    d = X.shape[2]
    lam, w = expressions.simplex_quadrature(d, 4)
    lam = lam[:, : d + 1]
    cols = [np.zeros_like(F) for _ in range(4 if source_values is not None else 3)]
    for c, (x, u) in enumerate(zip(X, U)):
        K, M, vol = _p1_reference_tensors(x)
        cols[0][c], cols[1][c] = K @ u, M @ u
        cols[2][c] = vol * (w * (lam @ u) ** 3) @ lam
        if source_values is not None:
            cols[3][c] = -(M @ np.asarray(source_values[c], dtype=float))
    B = np.stack([col.reshape(-1) for col in cols], axis=1)
    y = F.reshape(-1)
    if np.linalg.matrix_rank(B, tol=1e-12 * max(np.linalg.norm(B), 1e-300)) < B.shape[1]:
        return None
    coef, *_ = np.linalg.lstsq(B, y, rcond=None)
    if np.linalg.norm(B @ coef - y) > tol * max(np.linalg.norm(y), 1e-300):
        return None
    scale = max(np.max(np.abs(coef)), 1e-300)
    out = [0.0 if abs(v) < tol * scale else float(v) for v in coef]
    return tuple(out) if source_values is not None else (out[0], out[1], out[2], 0.0)


# ----------------------------------------------------------------------------- the LinearSolver plug-in
def _linear_solver_base():
    try:  # the reference's class when cashocs (and with it FEniCS) is importable
        import cashocs

        return cashocs._utils.linalg.LinearSolver  # noqa: SLF001
    except Exception:  # noqa: BLE001
        return object


def _ksp_error(reason: int):
    try:
        import cashocs

        return cashocs._exceptions.PETScKSPError(reason)  # noqa: SLF001  (cashocs/_utils/linalg.py:543-544)
    except Exception:  # noqa: BLE001
        return _exceptions.PETScKSPError(reason)


class B200LinearSolver(_linear_solver_base()):
    """Drop-in for ``cashocs.LinearSolver`` (``cashocs/_utils/linalg.py:480-546``): ``solve(function, A, b, ksp_options,
    rtol, atol, P)`` with a PETSc AIJ matrix (``getValuesCSR``, ``getBlockSize``) and right-hand side (``getArray``); the
    solution goes into ``function.vector().vec()``; ``b is None`` => x = 0; a negative reason raises ``PETScKSPError``.
    The matrix is copied to the device per solve (the zero-copy path replaces assembly as well, ``INTEGRATION.md`` 2).

    ``node_coordinates`` ([n_nodes, d], optional): what PETSc takes through ``PCSetCoordinates`` -- enables the rigid-body
    coarse correction of the AMG preconditioner for vertex-blocked vector operators."""

    def __init__(self, node_coordinates=None, device: str = "cuda") -> None:
        base = type(self).__mro__[1]
        if base is not object:
            super().__init__()
        self.node_coordinates = node_coordinates
        self.device = device
        self.last_result = None

    def solve(self, function, A=None, b=None, ksp_options=None, rtol=None, atol=None, P=None) -> None:
        import torch

        from cashocs_b200 import _lib
        from cashocs_b200._utils import linalg

        vec = function.vector().vec()
        if b is None:  # cashocs/_utils/linalg.py:507-509
            vec.set(0.0)
            function.vector().apply("")
            return
        if A is None:
            raise _exceptions.InputError("cashocs_b200.fenics_adapter.B200LinearSolver.solve", "A",
                                         "The KSP object has to be initialized with some Matrix in case A is None.")
        bs = int(A.getBlockSize())
        rowptr, col, val = A.getValuesCSR()
        rowptr, col, val = aij_to_baij(rowptr, col, val, bs)
        dev = torch.device(self.device)
        t_rowptr = torch.as_tensor(np.ascontiguousarray(rowptr, dtype=np.int64)).to(dev)
        t_col = torch.as_tensor(np.ascontiguousarray(col, dtype=np.int32)).to(dev)
        t_val = torch.as_tensor(np.ascontiguousarray(val, dtype=np.float64).reshape(-1)).to(dev)
        rhs = torch.as_tensor(np.ascontiguousarray(b.getArray(), dtype=np.float64)).to(dev)
        x = torch.zeros_like(rhs)
        holder = _PatternHolder(t_rowptr, t_col, rowptr.size - 1, dev, self.node_coordinates)
        M = linalg.Matrix(holder, bs, val=t_val)
        res = linalg.LinearSolver().solve(x, A=M, b=rhs, ksp_options=_translate_options(ksp_options), rtol=rtol, atol=atol)
        self.last_result = res
        vec.setArray(x.cpu().numpy())
        function.vector().apply("")


def _translate_options(ksp_options):
    """cashocs' ``None`` means "direct solve" (``linalg.py:306-325``); the device layer maps that itself."""
    return None if ksp_options is None else dict(ksp_options)


class _PatternHolder:
    """The part of ``cashocs_b200.Mesh`` the solver touches for a matrix that did not come from a device mesh: the
    sparsity pattern, the coordinates of the nodes (for the rigid-body modes) and a reduction scratch."""

    dist = None

    def __init__(self, rowptr, col, n, device, coords) -> None:
        import torch

        self.rowptr, self.col = rowptr, col
        self.nnz = int(col.numel())
        self.n_owned_vertices = self.num_vertices = int(n)
        self.device = device
        self.scratch = None
        self.coords = None
        self.dim = 0
        if coords is not None:
            self.coords = torch.as_tensor(np.ascontiguousarray(coords, dtype=np.float64)).to(device)
            self.dim = int(self.coords.shape[1])

    def build_pattern(self) -> None:
        pass

    def reduce_scratch(self):
        import torch

        from cashocs_b200 import _lib

        if self.scratch is None:
            self.scratch = torch.zeros(int(_lib.load().cb_reduce_scratch_len()), dtype=torch.float64, device=self.device)
        return self.scratch


# ----------------------------------------------------------------------------- dolfin glue (needs FEniCS)
def device_mesh_from_dolfin(mesh, boundaries, device: str = "cuda"):
    """dolfin ``Mesh`` + facet ``MeshFunction`` -> ``cashocs_b200.Mesh`` (coords, cells, tagged boundary facets)."""
    import fenics

    from cashocs_b200.geometry import mesh as mesh_module

    tdim = mesh.topology().dim()
    mesh.init(tdim - 1, tdim)
    facets, tags = [], []
    for f in fenics.facets(mesh):
        if f.exterior():
            facets.append(f.entities(0))
            tags.append(int(boundaries[f]))
    return mesh_module.mesh_from_arrays(mesh.coordinates(), mesh.cells(), np.asarray(facets), np.asarray(tags), device=device)


def _sample_cells(mesh, count: int = 6):
    import fenics

    cells = [c for _, c in zip(range(count), fenics.cells(mesh))]
    coords = mesh.coordinates()
    return cells, np.stack([coords[c.entities(0)] for c in cells])


def recognise_forms(lhs_form, rhs_form, V):
    """``PoissonForm`` descriptor of a dolfin (lhs, rhs) pair on the scalar P1 space ``V``, or ``None`` when the pair is
    outside the supported family -- by tabulating element tensors on sample cells (``fenics.assemble_local``) and
    fitting them (:func:`recognise_bilinear`, :func:`recognise_linear`, :func:`fit_polynomial`)."""
    import fenics
    import ufl

    import cashocs_b200 as cb

    mesh = V.mesh()
    if V.ufl_element().degree() != 1 or V.num_sub_spaces() != 0:
        return None
    cells, X = _sample_cells(mesh)
    d2v = fenics.dof_to_vertex_map(V)
    perm = []
    for c in cells:  # element tensors come in cell-dof order: permute to vertex order
        dofs = V.dofmap().cell_dofs(c.index())
        verts = list(c.entities(0))
        perm.append([list(dofs).index(dd) for dd in sorted(dofs, key=lambda q: verts.index(d2v[q]))])
    Ae = np.stack([np.asarray(fenics.assemble_local(lhs_form, c))[np.ix_(p, p)] for c, p in zip(cells, perm)])
    fit = recognise_bilinear(Ae, X)
    if fit is None:
        return None
    kappa, reaction = fit
    if rhs_form is None or rhs_form.empty():
        return cb.PoissonForm(source=None, kappa=kappa, reaction=reaction)
    be = np.stack([np.asarray(fenics.assemble_local(rhs_form, c))[p] for c, p in zip(cells, perm)])
    coeffs = [w for w in rhs_form.coefficients() if isinstance(w, fenics.Function) and w.function_space() == V]
    x = ufl.SpatialCoordinate(mesh)
    integrand_terms = [itg.integrand() for itg in rhs_form.integrals() if itg.integral_type() == "cell"]
    poly = None
    if any(xi in ufl.algorithms.extract_type(t, ufl.classes.SpatialCoordinate) for t in integrand_terms for xi in [x]):
        # the source is an expression of x: divide the test function out by evaluating the integrand with v = 1
        v = rhs_form.arguments()[0]
        expr = sum(ufl.replace(t, {v: ufl.constantvalue.FloatValue(1.0)}) for t in integrand_terms)
        poly = fit_polynomial(lambda pts: np.array([expr(tuple(p)) for p in pts]), mesh.geometry().dim())
        if poly is None:
            return None
    fields = [np.stack([w.compute_vertex_values(mesh)[c.entities(0)] for c in cells]) for w in coeffs]
    fit = recognise_linear(be, X, fields, (lambda pts: poly_eval(poly, pts)) if poly is not None else None)
    if fit is None:
        return None
    s_p, s_k = fit
    if poly is not None and not s_k:
        return cb.PoissonForm(source=poly, kappa=kappa, reaction=reaction, source_scale=s_p)
    if poly is None and len(s_k) == 1:
        return ("nodal", coeffs[0], cb.PoissonForm(source=None, kappa=kappa, reaction=reaction, source_scale=s_k[0]))
    return None


def poly_eval(poly, pts):
    pts = np.asarray(pts, dtype=float)
    out = np.zeros(pts.shape[0])
    for e, c in poly.terms.items():
        out += c * np.prod(pts ** np.asarray(e[: pts.shape[1]])[None, :], axis=1)
    return out


def install(problem, node_coordinates=None):
    """Put the device linear solver behind the seams of a cashocs problem (``INTEGRATION.md`` 1): the state, adjoint and
    gradient solves then run through ``cb_ksp_solve``.  ``problem`` is a ``cashocs.ShapeOptimizationProblem`` or
    ``cashocs.OptimalControlProblem`` (``optimization_problem.py:285-299``, ``pde_problem.py:53-56``)."""
    problem.state_problem.linear_solver = B200LinearSolver()
    problem.adjoint_problem.linear_solver = B200LinearSolver()
    problem.gradient_problem.linear_solver = B200LinearSolver(node_coordinates=node_coordinates)
    return problem
