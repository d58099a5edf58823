"""Device linear algebra behind the reference's ``cashocs/_utils/linalg.py`` surface.

* :class:`Matrix` -- block-CSR in HBM (PETSc AIJ / BAIJ layout),
* :func:`assemble_scalar_system` / :func:`assemble_elasticity_matrix` /
  :func:`assemble_shape_derivative` -- the ``assemble_petsc_system`` boundary
  (``cashocs/_utils/linalg.py:107-188``) for the supported form family,
* :class:`LinearSolver` -- same ``solve(function, A, b, ksp_options, rtol, atol, P)`` contract
  as ``cashocs/_utils/linalg.py:480-546`` (zero initial guess, ``b is None`` -> x = 0,
  negative reason -> :class:`PETScKSPError`),
* :func:`parse_ksp_options` -- the PETSc option names the reference passes as dicts
  (``cashocs/_typing.py:58``, ``setup_petsc_options`` ``linalg.py:191-217``).
"""

from __future__ import annotations

import copy
import ctypes as C
from typing import Any

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200 import log
from cashocs_b200._forms import expressions

KspOption = dict[str, Any]

# the reference's defaults (cashocs/_utils/linalg.py:44-59) -- kept verbatim as *names*; the
# device library maps them explicitly in parse_ksp_options (no silent substitution).
iterative_ksp_options: KspOption = {
    "ksp_type": "cg",
    "pc_type": "hypre",
    "pc_hypre_type": "boomeramg",
    "pc_hypre_boomeramg_strong_threshold": 0.7,
    "ksp_rtol": 1e-20,
    "ksp_atol": 1e-50,
    "ksp_max_it": 1000,
}

direct_ksp_options: KspOption = {
    "ksp_type": "preonly",
    "pc_type": "lu",
    "pc_factor_mat_solver_type": "mumps",
    "mat_mumps_icntl_24": 1,
}

# What "solve to machine precision" means on the device: there is no sparse LU on the GPU, so
# the direct / AMG option sets of the reference are mapped to Chebyshev-Jacobi PCG driven to
# a tight relative tolerance.  The mapping is explicit and documented ("KSP mapping").
DIRECT_EQUIVALENT: KspOption = {
    "ksp_type": "cg",
    "pc_type": "jacobi",
    "ksp_rtol": 1e-13,
    "ksp_atol": 1e-50,
    "ksp_max_it": 200000,
}

_KSP_TYPES = {"cg": 0, "minres": 1, "preonly": 2}
_PC_TYPES = {"none": 0, "jacobi": 1, "chebyshev": 2, "amg": 3}
_IGNORED_KEYS = {
    "ksp_monitor", "ksp_monitor_true_residual", "ksp_converged_reason", "ksp_view",
    "pc_jacobi_type", "mat_mumps_icntl_24",
    "pc_factor_mat_solver_type", "pc_hypre_type", "pc_hypre_boomeramg_strong_threshold",
    "ksp_cg_single_reduction",
}


def parse_ksp_options(ksp_options: KspOption | None, rtol=None, atol=None, distributed: bool = False) -> _lib.cb_ksp_opts:
    """Translate a cashocs ``ksp_options`` dict into the C-ABI options struct.

    ``None`` means "direct solve" in the reference (``define_ksp_options``,
    ``cashocs/_utils/linalg.py:306-325``).  ``preonly``+``lu`` and ``pc_type hypre`` have no
    device counterpart and are mapped to :data:`DIRECT_EQUIVALENT` resp. Jacobi with the
    caller's tolerances; anything else that is unknown raises -- never a silent fallback.
    """
    opts = copy.deepcopy(direct_ksp_options) if ksp_options is None else dict(ksp_options)
    # options whose semantics are fixed on the device: accepted only with the value that is implemented
    # (zero initial guess, convergence test on the preconditioned residual norm) -- never silently dropped
    if "ksp_initial_guess_nonzero" in opts:
        v = opts.pop("ksp_initial_guess_nonzero")
        if v is None or _truthy(v):
            raise _exceptions.InputError("cashocs_b200.LinearSolver", "ksp_options",
                                         "ksp_initial_guess_nonzero is not supported: cb_ksp_solve always starts from x = 0 "
                                         "(as cashocs/_utils/linalg.py:480-546 does by default)")
    if "ksp_norm_type" in opts:
        v = str(opts.pop("ksp_norm_type")).casefold()
        if v not in ("preconditioned", "default"):
            raise _exceptions.InputError("cashocs_b200.LinearSolver", "ksp_options",
                                         f"ksp_norm_type {v!r} is not supported: the device tests the preconditioned norm")
    ksp = str(opts.get("ksp_type", "gmres")).lower()
    pc = str(opts.get("pc_type", "none")).lower()
    flexible = opts.get("ksp_cg_flexible", None)
    if ksp == "fcg":  # PETSc KSPFCG (mmax = 1): CG with the Polak-Ribiere beta
        ksp, flexible = "cg", True
    if ksp == "preonly" and pc in ("lu", "cholesky"):
        base = dict(DIRECT_EQUIVALENT)
        ksp, pc = base["ksp_type"], base["pc_type"]
        opts = {**base}
    elif pc in ("hypre", "gamg", "ml", "amg"):
        # algebraic multigrid -> the device smoothed-aggregation AMG (also on a partitioned mesh:
        # rank-local aggregates, tentative interface rows, per-level halo plans)
        pc = "amg"
    if pc == "ksp":  # PETSc spelling of a polynomial preconditioner
        if str(opts.get("ksp_ksp_type", "")).lower() != "chebyshev":
            raise _exceptions.InputError("cashocs_b200.LinearSolver", "ksp_options", "pc_type ksp needs ksp_ksp_type chebyshev")
        pc = "chebyshev"
        opts.setdefault("pc_chebyshev_degree", int(opts.get("ksp_ksp_max_it", 3)))
    if ksp not in _KSP_TYPES:
        raise _exceptions.InputError(
            "cashocs_b200.LinearSolver", "ksp_options",
            f"ksp_type {ksp!r} is not supported on the device (supported: {sorted(_KSP_TYPES)})",
        )
    if pc not in _PC_TYPES:
        raise _exceptions.InputError(
            "cashocs_b200.LinearSolver", "ksp_options",
            f"pc_type {pc!r} is not supported on the device (supported: {sorted(_PC_TYPES)}, lu/hypre are mapped)",
        )
    known = {"ksp_type", "pc_type", "ksp_rtol", "ksp_atol", "ksp_divtol", "ksp_max_it",
             "pc_chebyshev_degree", "pc_chebyshev_ratio", "ksp_check_every", "ksp_ksp_type",
             "ksp_ksp_max_it", "ksp_pc_type", "ksp_profile_every", "ksp_use_graph", "pc_amg_smooth_degree",
             "pc_amg_coarse_degree", "pc_amg_smooth_ratio", "pc_amg_coarse_ratio", "pc_amg_min_coarse",
             "pc_amg_rigid_body_correction", "pc_amg_nullspace", "pc_amg_precision", "ksp_cg_flexible", "pc_amg_level_smooth_degree", "pc_amg_lambda_safety",
             "pc_amg_prolongator_damping",
             "pc_gamg_type", "pc_gamg_agg_nsmooths"} | _IGNORED_KEYS
    unknown = set(opts) - known
    if unknown:
        raise _exceptions.InputError(
            "cashocs_b200.LinearSolver", "ksp_options", f"unsupported PETSc options: {sorted(unknown)}"
        )
    o = _lib.cb_ksp_opts()
    o.ksp_type = _KSP_TYPES[ksp]
    o.pc_type = _PC_TYPES[pc]
    o.rtol = float(opts.get("ksp_rtol", 1e-5)) if rtol is None else float(rtol)
    if pc == "amg" and o.rtol < 1e-14:
        o.rtol = 1e-14  # rtol 1e-16 / 1e-20 is below what fp64 round-off lets PCG reach (also for an explicit rtol argument)
    o.atol = float(opts.get("ksp_atol", 1e-50)) if atol is None else float(atol)
    o.dtol = float(opts.get("ksp_divtol", 1e5))
    o.max_it = int(opts.get("ksp_max_it", 10000))
    o.cheb_degree = int(opts.get("pc_chebyshev_degree", 3))
    o.cheb_ratio = float(opts.get("pc_chebyshev_ratio", 30.0))
    o.check_every = int(opts.get("ksp_check_every", 0))
    o.monitor = 1 if ("ksp_monitor" in opts) else 0
    o.profile_every = int(opts.get("ksp_profile_every", 0))
    o.use_graph = int(opts.get("ksp_use_graph", 0))
    # -1: unset (the solver turns it on for a mixed-precision AMG hierarchy), 0 / 1 explicit
    o.flexible = -1 if flexible is None else int(_truthy(flexible))
    return o


def row_split_hint(nnzb: int, nrows: int) -> int:
    """Lane groups per row for the SpMV kernel: 1 for mesh stencils (~15 blocks / row), more for the
    denser coarse AMG operators so that every lane still owns ~16 blocks."""
    avg = nnzb / max(nrows, 1)
    return 1 if avg <= 24 else (2 if avg <= 48 else (4 if avg <= 96 else 8))


def cta_warps_hint(nnzb: int, nrows: int, bs: int, row_split: int = 1) -> int:
    """Warps per CTA for the bulk-copy SpMV-type pass over an operator with short rows (``cb_mat.cta_warps``): a CTA
    stages the values of all its rows with one bulk copy into a tile of 16 blocks per row of the default shape
    (csrc/spmv.cuh: 2 warps x 10 rows for 3 x 3 blocks, 4 warps x 32 / 16 rows otherwise); with fewer blocks per row
    more rows fit, which keeps the copy size -- the bytes in flight per SM -- where the sweeps found the DRAM pipe
    saturated.  0 = default."""
    default = 2 if bs == 3 else 4
    g = max(int(row_split), 1)
    rpw = 32 // (bs * g)
    tile = 16 * rpw * default if g == 1 else 16 * (32 // bs) * default
    avg = nnzb / max(nrows, 1)
    want = int(0.9 * tile / max(avg * rpw, 1.0))
    w = min(max(want, default), 2 * default)
    return 0 if w == default else w


# matrices are assembled row by row (True; csrc/assemble.cu k_assemble_rows); "entries" selects the entry-centric
# gather kernels over the nnz -> cell lists, False the coloured cell-scatter kernels (both kept as cross-checks in the
# tests; the scatter kernels also serve meshes whose scatter map exceeds int32)
USE_GATHER_ASSEMBLY: bool | str = True


class Matrix:
    """Block-CSR matrix in HBM sharing the mesh's P1 pattern (``bs`` = 1 scalar, ``dim`` vector)."""

    def __init__(self, mesh, bs: int, val: torch.Tensor | None = None) -> None:
        mesh.build_pattern()
        self.mesh = mesh
        self.bs = int(bs)
        self.nrows = mesh.n_owned_vertices
        self.ncols = mesh.num_vertices
        self.rowptr = mesh.rowptr
        self.col = mesh.col
        self.nnzb = mesh.nnz
        self.val = (
            val if val is not None
            else torch.zeros(self.nnzb * self.bs * self.bs, dtype=torch.float64, device=mesh.device)
        )
        self.version = 0  # bumped by every (re-)assembly; AMG hierarchies key their numeric phase on it

    def touch(self) -> None:
        """Tell cached preconditioners that the values changed."""
        self.version += 1

    @property
    def shape(self):
        return (self.nrows * self.bs, self.ncols * self.bs)

    def struct(self) -> _lib.cb_mat:
        m = _lib.cb_mat()
        m.nrows, m.ncols, m.bs = self.nrows, self.ncols, self.bs
        m.row_split = row_split_hint(self.nnzb, self.nrows)
        m.rowptr, m.col, m.val = self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr()
        return m

    def mult(self, x: torch.Tensor, y: torch.Tensor | None = None) -> torch.Tensor:
        """y = A x (``PETSc.Mat.mult``)."""
        _lib.require_cuda(x, y)
        if y is None:
            y = torch.empty(self.nrows * self.bs, dtype=torch.float64, device=x.device)
        m = self.struct()
        _lib.check(_lib.load().cb_spmv(_lib.stream_ptr(), C.byref(m), _lib.ptr(x), _lib.ptr(y)), "cb_spmv")
        return y

    def mult_f32(self, x: torch.Tensor, y: torch.Tensor | None = None) -> torch.Tensor:
        """y = fl32(A) x: the preconditioner's pass over the fp32 copy of the values (made on first use)."""
        _lib.require_cuda(x, y)
        if y is None:
            y = torch.empty(self.nrows * self.bs, dtype=torch.float64, device=x.device)
        v32 = self.__dict__.get("_val32")
        if v32 is None or self.__dict__.get("_val32_version") != self.version:
            v32 = self.__dict__.setdefault("_val32", torch.empty(self.val.numel(), dtype=torch.float32, device=self.val.device))
            _lib.check(_lib.load().cb_to_f32(_lib.stream_ptr(), self.val.numel(), _lib.ptr(self.val), _lib.ptr(v32)), "cb_to_f32")
            self._val32_version = self.version
        m = self.struct()
        m.val32 = v32.data_ptr()
        _lib.check(_lib.load().cb_spmv_f32(_lib.stream_ptr(), C.byref(m), _lib.ptr(x), _lib.ptr(y)), "cb_spmv_f32")
        return y

    def scalar_product(self, x: torch.Tensor, y: torch.Tensor) -> float:
        """y^T A x fused on the device (``ShapeFormHandler.scalar_product``, shape_form_handler.py:748-780)."""
        _lib.require_cuda(x, y)
        out = torch.empty(1, dtype=torch.float64, device=x.device)
        scratch = self.mesh.reduce_scratch()
        m = self.struct()
        _lib.check(
            _lib.load().cb_scalar_product(_lib.stream_ptr(), C.byref(m), _lib.ptr(x), _lib.ptr(y),
                                          _lib.ptr(out), _lib.ptr(scratch), scratch.numel()),
            "cb_scalar_product",
        )
        if self.mesh.dist is not None:
            self.mesh.dist.allreduce(out, "sum")
        return float(out.item())

    def ident_zeros(self) -> int:
        """``PETScMatrix.ident_zeros`` (``cashocs/_utils/linalg.py:167``)."""
        nfix = torch.zeros(1, dtype=torch.int64, device=self.val.device)
        _lib.check(
            _lib.load().cb_ident_zeros(_lib.stream_ptr(), self.nrows, self.bs, _lib.ptr(self.rowptr),
                                       _lib.ptr(self.col), _lib.ptr(self.val), _lib.ptr(nfix)),
            "cb_ident_zeros",
        )
        return int(nfix.item())

    def to_scipy(self):
        """Host copy as scipy CSR/BSR (tests and debugging only)."""
        import scipy.sparse as sp

        rp = self.rowptr.cpu().numpy()
        col = self.col.cpu().numpy()
        if self.bs == 1:
            return sp.csr_matrix((self.val.cpu().numpy(), col, rp), shape=(self.nrows, self.ncols))
        data = self.val.cpu().numpy().reshape(-1, self.bs, self.bs)
        return sp.bsr_matrix((data, col, rp), shape=(self.nrows * self.bs, self.ncols * self.bs)).tocsr()


def _bc_arrays(n: int, bc_dofs, bc_vals, device):
    """uint8 flag + fp64 value arrays for a Dirichlet dof set (later entries win, like a bcs list)."""
    if bc_dofs is None:
        return None, None
    flag = torch.zeros(n, dtype=torch.uint8, device=device)
    val = torch.zeros(n, dtype=torch.float64, device=device)
    if not isinstance(bc_dofs, (list, tuple)):
        bc_dofs, bc_vals = [bc_dofs], [bc_vals]
    for dofs, v in zip(bc_dofs, bc_vals):
        if dofs.numel() == 0:
            continue
        flag[dofs] = 1
        val[dofs] = v if isinstance(v, torch.Tensor) else float(v)
    return flag, val


@log.profile_execution_time("assembling the system")
def assemble_scalar_system(
    mesh,
    c_stiff: float = 1.0,
    c_mass: float = 0.0,
    f_poly: expressions.Polynomial | None = None,
    f_poly_scale: float = 1.0,
    f_nodal: torch.Tensor | None = None,
    f_nodal_scale: float = 1.0,
    bc_flag: torch.Tensor | None = None,
    bc_val: torch.Tensor | None = None,
    A: Matrix | None = None,
    b: torch.Tensor | None = None,
    matrix: bool = True,
    rhs: bool = True,
):
    """Scalar P1 system ``(c_stiff K + c_mass M) x = s_p int f phi + s_n M f_h`` with symmetric BCs.

    The device restatement of ``assemble_petsc_system`` (``cashocs/_utils/linalg.py:107-188``):
    full cell pattern, ``keep_diagonal``, per-element symmetric Dirichlet elimination,
    ``ident_zeros``.
    """
    lib = _lib.load()
    view = mesh.view(gather=USE_GATHER_ASSEMBLY)
    if matrix and A is None:
        A = Matrix(mesh, 1)
    if rhs and b is None:
        b = torch.empty(mesh.n_owned_vertices, dtype=torch.float64, device=mesh.device)
    _lib.require_cuda(mesh.coords, f_nodal, bc_flag, bc_val, b)
    poly = quad = None
    s_poly = 0.0
    if f_poly is not None and rhs:
        poly = f_poly.struct()
        quad = expressions.quadrature_struct(mesh.dim, f_poly.degree + 1)
        s_poly = float(f_poly_scale)
    s_nodal = float(f_nodal_scale) if (f_nodal is not None and rhs) else 0.0
    _lib.check(
        lib.cb_assemble_scalar(
            _lib.stream_ptr(), C.byref(view), float(c_stiff), float(c_mass), s_poly,
            C.byref(poly) if poly is not None else None, C.byref(quad) if quad is not None else None,
            s_nodal, _lib.ptr(f_nodal), _lib.ptr(bc_flag), _lib.ptr(bc_val),
            _lib.ptr(A.val) if matrix else None, _lib.ptr(b) if rhs else None,
        ),
        "cb_assemble_scalar",
    )
    if matrix:
        _lib.check(lib.cb_ident_zeros(_lib.stream_ptr(), A.nrows, 1, _lib.ptr(A.rowptr), _lib.ptr(A.col),
                                      _lib.ptr(A.val), None), "cb_ident_zeros")
        A.touch()
    return (A if matrix else None), (b if rhs else None)


def assemble_cubic(mesh, u: torch.Tensor, c: float, bc_flag: torch.Tensor | None = None, A: "Matrix | None" = None,
                   vec: torch.Tensor | None = None) -> None:
    """The cubic term ``c u^3 p`` of a semilinear state equation: adds its Jacobian ``3 c int u^2 phi_a phi_b`` to the
    values of ``A`` (Dirichlet rows / columns untouched) and / or writes its load ``c int u^3 phi_a`` into ``vec`` --
    what ``cashocs.newton_solve`` (``newton_solver.py:285-362``) re-assembles in every Newton step."""
    lib = _lib.load()
    view = mesh.view(gather=True)
    _lib.require_cuda(u, bc_flag, vec)
    quad = expressions.quadrature_struct(mesh.dim, 4)
    _lib.check(lib.cb_assemble_cubic(_lib.stream_ptr(), C.byref(view), _lib.ptr(u), float(c), C.byref(quad), _lib.ptr(bc_flag),
                                     _lib.ptr(A.val) if A is not None else None, _lib.ptr(vec)), "cb_assemble_cubic")
    if A is not None:
        A.touch()


@log.profile_execution_time("assembling the system")
def assemble_elasticity_matrix(mesh, mu: torch.Tensor, lambda_lame: float, damping: float,
                               bc_flag: torch.Tensor | None = None, cell_scale: torch.Tensor | None = None,
                               A: Matrix | None = None) -> Matrix:
    """Elasticity Riesz scalar product (``cashocs/_forms/shape_form_handler.py:655-685,737-738``)."""
    lib = _lib.load()
    view = mesh.view(gather=USE_GATHER_ASSEMBLY)
    if A is None:
        A = Matrix(mesh, mesh.dim)
    _lib.require_cuda(mu, bc_flag, cell_scale)
    if cell_scale is not None and not view.n2c_ptr:
        cell_scale = cell_scale[mesh.colour_perm].contiguous()  # the scatter kernels walk colour-sorted cells
    _lib.check(
        lib.cb_assemble_elasticity(_lib.stream_ptr(), C.byref(view), _lib.ptr(mu), float(lambda_lame),
                                   float(damping), _lib.ptr(cell_scale), _lib.ptr(bc_flag), _lib.ptr(A.val)),
        "cb_assemble_elasticity",
    )
    _lib.check(lib.cb_ident_zeros(_lib.stream_ptr(), A.nrows, A.bs, _lib.ptr(A.rowptr), _lib.ptr(A.col),
                                  _lib.ptr(A.val), None), "cb_ident_zeros")
    A.touch()
    return A


@log.profile_execution_time("assembling the system")
def assemble_shape_derivative(mesh, u: torch.Tensor, p: torch.Tensor, f_poly: expressions.Polynomial,
                              c_int_u: float = 1.0, bc_flag: torch.Tensor | None = None,
                              b: torch.Tensor | None = None, kappa: float = 1.0, reaction: float = 0.0,
                              c_track: float = 0.0, u_d: torch.Tensor | None = None,
                              pull_back: bool = True) -> torch.Tensor:
    """Shape-derivative vector of the Lagrangian ``int c_int_u u + c_track/2 (u - u_d)^2 + kappa grad u . grad p +
    reaction u p - f p`` (``shape_gradient_problem.py:142-144``, form ``shape_form_handler.py:533-549``); ``u_d`` is a P1
    field whose gradient enters through the pull-back of ``shape_form_handler.py:507-531`` when ``pull_back``."""
    lib = _lib.load()
    view = mesh.view(gather=USE_GATHER_ASSEMBLY)
    d = mesh.dim
    if b is None:
        b = torch.empty(mesh.n_owned_vertices * d, dtype=torch.float64, device=mesh.device)
    _lib.require_cuda(u, p, bc_flag, b, u_d)
    f = f_poly.struct()
    grads = (_lib.cb_poly * 3)()
    for i in range(d):
        grads[i] = f_poly.derivative(i).struct()
    quad = expressions.quadrature_struct(d, f_poly.degree + 1)
    _lib.check(
        lib.cb_assemble_shape_derivative(
            _lib.stream_ptr(), C.byref(view), _lib.ptr(u), _lib.ptr(p), C.byref(f), grads, C.byref(quad),
            float(c_int_u), float(kappa), float(reaction), float(c_track), _lib.ptr(u_d), int(bool(pull_back)),
            _lib.ptr(bc_flag), _lib.ptr(b)),
        "cb_assemble_shape_derivative",
    )
    return b


@log.profile_execution_time("assembling the system")
def assemble_p2_system(space, c_stiff: float = 1.0, c_mass: float = 0.0, f_poly: expressions.Polynomial | None = None,
                       f_poly_scale: float = 1.0, f_const: float = 0.0, f_nodal: torch.Tensor | None = None,
                       f_nodal_scale: float = 1.0, bc_flag: torch.Tensor | None = None,
                       bc_val: torch.Tensor | None = None, A: "Matrix | None" = None, b: torch.Tensor | None = None,
                       matrix: bool = True, rhs: bool = True):
    """Scalar P2 system ``(c_stiff K + c_mass M) x = s_p int f phi + f_const int phi + s_n M f_h`` with
    symmetric BCs -- ``assemble_petsc_system`` (``cashocs/_utils/linalg.py:107-188``) for
    ``FunctionSpace(mesh, "CG", 2)`` (BASELINE configs[4])."""
    lib = _lib.load()
    view = space.view()
    if matrix and A is None:
        A = Matrix(space, 1)
    if rhs and b is None:
        b = torch.empty(space.n_owned_dofs, dtype=torch.float64, device=space.device)
    _lib.require_cuda(f_nodal, bc_flag, bc_val, b)
    if matrix:
        _lib.check(lib.cb_p2_assemble_matrix(_lib.stream_ptr(), C.byref(view), float(c_stiff), float(c_mass),
                                             _lib.ptr(bc_flag), _lib.ptr(A.val)), "cb_p2_assemble_matrix")
        _lib.check(lib.cb_ident_zeros(_lib.stream_ptr(), A.nrows, 1, _lib.ptr(A.rowptr), _lib.ptr(A.col),
                                      _lib.ptr(A.val), None), "cb_ident_zeros")
        A.touch()
    if rhs:
        poly, nq, qtab, s_poly = None, 0, None, 0.0
        if f_poly is not None:
            poly = f_poly.struct()
            nq, qtab = space.qtab(f_poly.degree + 2)
            s_poly = float(f_poly_scale)
        s_nodal = float(f_nodal_scale) if f_nodal is not None else 0.0
        _lib.check(
            lib.cb_p2_assemble_rhs(_lib.stream_ptr(), C.byref(view), float(c_stiff), float(c_mass), s_poly,
                                   C.byref(poly) if poly is not None else None, nq, _lib.ptr(qtab), float(f_const),
                                   s_nodal, _lib.ptr(f_nodal), _lib.ptr(bc_flag), _lib.ptr(bc_val), _lib.ptr(b)),
            "cb_p2_assemble_rhs")
    return (A if matrix else None), (b if rhs else None)


@log.profile_execution_time("assembling the system")
def assemble_p2_shape_derivative(space, u: torch.Tensor, p: torch.Tensor, f_poly: expressions.Polynomial,
                                 c_int_u: float = 1.0, bc_flag: torch.Tensor | None = None,
                                 b: torch.Tensor | None = None) -> torch.Tensor:
    """Shape-derivative vector (vector-P1 test functions) for P2 state / adjoint."""
    lib = _lib.load()
    mesh = space.mesh
    mesh.build_pattern()
    d = mesh.dim
    if b is None:
        b = torch.empty(mesh.n_owned_vertices * d, dtype=torch.float64, device=mesh.device)
    _lib.require_cuda(u, p, bc_flag, b)
    view = space.view()
    f = f_poly.struct()
    grads = (_lib.cb_poly * 3)()
    for i in range(d):
        grads[i] = f_poly.derivative(i).struct()
    nq, qtab = space.qtab(f_poly.degree + 2)
    _lib.check(
        lib.cb_p2_shape_derivative_poisson(_lib.stream_ptr(), C.byref(view), _lib.ptr(mesh.v2c_ptr),
                                           _lib.ptr(mesh.v2c_idx), _lib.ptr(u), _lib.ptr(p), C.byref(f), grads, nq,
                                           _lib.ptr(qtab), float(c_int_u), _lib.ptr(bc_flag), _lib.ptr(b)),
        "cb_p2_shape_derivative_poisson")
    return b


class KSPResult:
    def __init__(self, reason: int, its: int, rnorm: float, rnorm0: float, history=None, spmv_ms: float = 0.0,
                 spmv_profiled: int = 0, spmv_launches: int = 0) -> None:
        self.reason, self.its, self.rnorm, self.rnorm0, self.history = reason, its, rnorm, rnorm0, history
        self.spmv_ms, self.spmv_profiled, self.spmv_launches = spmv_ms, spmv_profiled, spmv_launches

    def __repr__(self) -> str:
        return f"KSPResult(reason={self.reason}, its={self.its}, rnorm={self.rnorm:.3e}, rnorm0={self.rnorm0:.3e})"


def _truthy(v) -> bool:
    return v if isinstance(v, bool) else str(v).casefold() in ("1", "true", "yes", "on")


class LinearSolver:
    """A solver for linear problems (``cashocs/_utils/linalg.py:480-546``), device resident."""

    def __init__(self) -> None:
        self._work: torch.Tensor | None = None
        self.last_result: KSPResult | None = None
        self.total_iterations = 0

    # AMG data lives on the objects it describes, so it is released with them: the *symbolic* phase (pattern-only) is an
    # attribute of the mesh / function space that owns the sparsity pattern and is shared by every operator on it;
    # the *values* of a hierarchy belong to one Matrix and one option set and are refreshed when the matrix version
    # changes (state, mu-Laplace and mass matrices on one mesh no longer evict each other's numeric phase)
    @staticmethod
    def _amg_option_key(options: dict, bs: int) -> tuple:
        prec = str(options.get("pc_amg_precision", "mixed")).casefold()
        if not (prec in ("mixed", "double") or prec.startswith("mixed:")):
            raise _exceptions.InputError("cashocs_b200.LinearSolver", "pc_amg_precision",
                                         "mixed, double or mixed:<a0|a1|ap0|coarse joined by +>")
        # Defaults from the sweeps on the 10 M-tet elasticity / Poisson systems (tools/sweep_precision.py, gpurun_out/r02_sweep*):
        # smoother interval [lmax/20, lmax]; lmax = 1.1 x the (warm-started, hence converged) power-method estimate;
        # prolongator damping 1.6 / lmax(D_S^-1 S) instead of the textbook 4/3 (89 -> 65 PCG iterations together).
        # levels >= 1 hold a tenth of the fine level's bytes: vector operators smooth them with degree 2 (measured on
        # the 10 M-tet elasticity solve: 89 -> 70 PCG iterations for 11 % more time per iteration); scalar ones gain
        # nothing (35 -> 33)
        lsd = int(options.get("pc_amg_level_smooth_degree", 2 if bs > 1 else 0))
        return (int(options.get("pc_amg_smooth_degree", 1)), int(options.get("pc_amg_coarse_degree", 16)),
                float(options.get("pc_amg_smooth_ratio", 20.0)), float(options.get("pc_amg_coarse_ratio", 40.0)), prec, lsd,
                float(options.get("pc_amg_lambda_safety", 1.1)), float(options.get("pc_amg_prolongator_damping", 1.6)))

    @classmethod
    def amg_hierarchy(cls, A: "Matrix", options: dict):
        from cashocs_b200._utils import amg

        key = cls._amg_option_key(options, A.bs)
        holder = A.mesh
        sym = getattr(holder, "_amg_symbolic", None)
        if sym is None or sym.levels[0].rowptr is not A.rowptr or sym.distance != amg.AMGSymbolic.DISTANCE:
            sym = amg.AMGSymbolic(A.rowptr, A.col, A.nrows, A.ncols, dist=holder.dist)
            sym.distance = amg.AMGSymbolic.DISTANCE
            holder._amg_symbolic = sym
        cache = A.__dict__.setdefault("_amg", {})
        h = cache.get(key)
        if h is None or h.symbolic is not sym:
            if len(cache) >= 2:  # option sweeps: keep the memory of at most two value sets per matrix
                cache.clear()
            prec = key[4]
            h = amg.AMGHierarchy(A, sym, smooth_degree=key[0], coarse_degree=key[1], smooth_ratio=key[2],
                                 coarse_ratio=key[3], mixed=prec != "double", level_smooth_degree=key[5],
                                 mixed_mask=set(prec[6:].split("+")) if prec.startswith("mixed:") else None,
                                 lambda_safety=key[6], prolongator_damping=key[7])
            cache[key] = h
        if h.version != A.version:
            h.numeric(A)
            h.version = A.version
            h.rbm = None
        return h

    @classmethod
    def rbm_hierarchy(cls, A: "Matrix"):
        from cashocs_b200._utils import amg_rbm

        h = A.__dict__.get("_amg_rbm")
        if h is None or h.levels[0].rowptr is not A.rowptr:
            sym = getattr(A.mesh, "_amg_rbm_symbolic", None)
            if sym is None or sym.levels[0].rowptr is not A.rowptr:
                sym = amg_rbm.RBMSymbolic(A.rowptr, A.col, A.nrows)
                A.mesh._amg_rbm_symbolic = sym
            h = amg_rbm.RBMHierarchy(A, sym)
            h.version = None
            A._amg_rbm = h
        if h.version != A.version:
            h.A = A
            h.numeric(A.mesh.coords)
            h.version = A.version
        return h

    @staticmethod
    def rigid_body_correction(A: "Matrix", h):
        """Data of the coarse correction ``Z (Z^T A Z)^-1 Z^T`` with the global rigid-body modes (see
        ``cb_ksp_opts.rbm_*``): centred node coordinates, the inverse of the k x k Galerkin matrix (k = 3 in 2-D,
        6 in 3-D) and the work vector.  Per matrix version: k SpMVs + k projections + one k x k inverse, all native
        (``csrc/rbm.cu``); the only torch op is the one-off centre of the initial mesh."""
        if getattr(h, "rbm", None) is not None:
            return h.rbm
        mesh, bs = A.mesh, A.bs
        if bs != mesh.dim or A.nrows != mesh.n_owned_vertices:
            raise _exceptions.InputError("cashocs_b200.LinearSolver", "pc_amg_rigid_body_correction",
                                         "needs a vertex-blocked vector operator (bs = dim) on the mesh vertices")
        dist = mesh.dist
        n_own, k = mesh.n_owned_vertices, (6 if bs == 3 else 3)
        centre = mesh.__dict__.get("_rbm_centre")
        if centre is None:
            # any fixed point spans the same modes; the centroid of the initial mesh keeps the rotations well scaled
            acc = torch.cat([mesh.coords[:n_own].sum(dim=0), torch.tensor([float(n_own)], dtype=torch.float64,
                                                                          device=mesh.device)])
            if dist is not None:
                dist.allreduce(acc, "sum")
            acc = acc.cpu()
            centre = mesh._rbm_centre = (C.c_double * 3)(*[float(acc[a] / acc[bs]) for a in range(bs)])
        buf = h.__dict__.get("_rbm_buf")
        if buf is None:
            f64 = dict(dtype=torch.float64, device=mesh.device)
            buf = h._rbm_buf = dict(c=torch.zeros(A.ncols * bs, **f64), v=torch.zeros(A.ncols * bs, **f64),
                                    w=torch.zeros(A.nrows * bs, **f64), G=torch.zeros(k * k, **f64),
                                    einv=torch.zeros(k * k, **f64), work=torch.zeros(8, **f64),
                                    # own reduction scratch (the set-up may be issued from another stream than the
                                    # mesh-level reductions)
                                    scratch=torch.zeros(int(_lib.load().cb_reduce_scratch_len()), **f64))
        lib, s = _lib.load(), _lib.stream_ptr()
        scratch = buf["scratch"]
        m = A.struct()
        _lib.check(lib.cb_rbm_galerkin(s, C.byref(m), _lib.ptr(mesh.coords), C.cast(centre, C.c_void_p), _lib.ptr(buf["c"]),
                                       _lib.ptr(buf["v"]), _lib.ptr(buf["w"]), _lib.ptr(buf["G"]), _lib.ptr(scratch),
                                       scratch.numel()), "cb_rbm_galerkin")
        if dist is not None:
            dist.allreduce(buf["G"], "sum")
        _lib.check(lib.cb_rbm_invert(s, k, _lib.ptr(buf["G"]), _lib.ptr(buf["einv"])), "cb_rbm_invert")
        h.rbm = (buf["c"], buf["einv"], buf["work"])
        return h.rbm

    def _setup_preconditioner(self, A: Matrix, ksp_options, opts):
        """AMG hierarchy (numeric phase if the matrix changed) and rigid-body coarse correction for ``opts``; fills the
        pointers of ``opts`` and returns the objects that own them."""
        hierarchy = hstruct = None
        nullspace = str((ksp_options or {}).get("pc_amg_nullspace", "translations")).casefold()
        if opts.pc_type == _PC_TYPES["amg"] and nullspace == "rigid_body":
            # smoothed aggregation with the six rigid-body modes inside the hierarchy (block prolongator, p_block = 1)
            if A.bs != 3 or A.mesh.dist is not None:
                raise _exceptions.InputError("cashocs_b200.LinearSolver", "pc_amg_nullspace",
                                             "rigid_body needs a 3 x 3 block operator on an unpartitioned mesh")
            hierarchy = self.rbm_hierarchy(A)
            hstruct = hierarchy.struct()
            opts.amg = C.cast(C.pointer(hstruct), C.c_void_p)
        elif opts.pc_type == _PC_TYPES["amg"]:
            hierarchy = self.amg_hierarchy(A, ksp_options or {})
            hstruct = hierarchy.struct()
            opts.amg = C.cast(C.pointer(hstruct), C.c_void_p)
            if opts.flexible < 0:
                opts.flexible = 1 if hierarchy.mixed else 0
            # rigid-body coarse correction: default for vertex-blocked vector operators (the elasticity Riesz
            # problem), on one GPU and on partitioned meshes (one extra 6-value all-reduce per iteration)
            applicable = (A.bs >= 2 and A.bs == A.mesh.dim and A.nrows == A.mesh.n_owned_vertices
                          and nullspace != "rigid_body")
            wanted = (ksp_options or {}).get("pc_amg_rigid_body_correction", None)
            use_rbm = applicable and (True if wanted is None else _truthy(wanted))
            if use_rbm:
                rbm = self.rigid_body_correction(A, hierarchy)
                opts.rbm_coords, opts.rbm_einv, opts.rbm_work = (t.data_ptr() for t in rbm)
        return hierarchy, hstruct

    def prepare_preconditioner(self, A: Matrix, ksp_options: KspOption | None = None) -> None:
        """Set up the preconditioner of a later ``solve(A=A, ksp_options=ksp_options)`` now -- everything is cached on
        the matrix version, so the solve finds it ready (used to time the set-up apart from the iteration)."""
        opts = parse_ksp_options(ksp_options, None, None, distributed=A.mesh.dist is not None)
        self._setup_preconditioner(A, ksp_options, opts)

    def _workspace(self, n_doubles: int, device) -> torch.Tensor:
        if self._work is None or self._work.numel() < n_doubles or self._work.device != device:
            self._work = torch.empty(n_doubles, dtype=torch.float64, device=device)
        return self._work

    @log.profile_execution_time("solving the linear system")
    def solve(
        self,
        function: torch.Tensor,
        A: Matrix | None = None,
        b: torch.Tensor | None = None,
        ksp_options: KspOption | None = None,
        rtol: float | None = None,
        atol: float | None = None,
        P: Matrix | None = None,
    ) -> KSPResult:
        """Solve ``A x = b`` into ``function`` (a device vector); raises ``PETScKSPError`` on failure."""
        if A is None:
            raise _exceptions.InputError(
                "cashocs_b200.LinearSolver.solve", "A",
                "The KSP object has to be initialized with some Matrix in case A is None.",
            )
        if P is not None:
            raise _exceptions.InputError("cashocs_b200.LinearSolver.solve", "P", "separate preconditioner matrices are not supported")
        _lib.require_cuda(function, b, A.val)
        if b is None:
            function.zero_()
            self.last_result = KSPResult(3, 0, 0.0, 0.0)
            return self.last_result
        opts = parse_ksp_options(ksp_options, rtol, atol, distributed=A.mesh.dist is not None)
        lib = _lib.load()
        keep = self._setup_preconditioner(A, ksp_options, opts)  # noqa: F841  (keeps the C structs alive)
        if opts.flexible < 0:
            opts.flexible = 0
        n, nc = A.nrows * A.bs, A.ncols * A.bs
        if function.numel() < n or b.numel() < n:
            raise _exceptions.InputError("cashocs_b200.LinearSolver.solve", "function", "vector too short for the matrix")
        work = self._workspace(int(lib.cb_ksp_workspace_len(n, nc, C.byref(opts))), function.device)
        res = _lib.cb_ksp_result()
        hist = None
        hist_len = 0
        if opts.monitor:
            hist_len = opts.max_it + 1
            hist = np.zeros(hist_len)
        m = A.struct()
        dist = A.mesh.dist.handle if A.mesh.dist is not None else None
        _lib.check(
            lib.cb_ksp_solve(_lib.stream_ptr(), C.byref(m), _lib.ptr(b), _lib.ptr(function), C.byref(opts),
                             dist, _lib.ptr(work), work.numel(), C.byref(res), _lib.ptr(hist), hist_len),
            "cb_ksp_solve",
        )
        if A.mesh.dist is not None:
            A.mesh.dist.check_p2p()
        history = hist[: res.its + 1].copy() if hist is not None else None
        self.last_result = KSPResult(int(res.reason), int(res.its), float(res.rnorm), float(res.rnorm0), history,
                                     float(res.spmv_ms), int(res.spmv_profiled), int(res.spmv_launches))
        self.total_iterations += int(res.its)
        if res.reason < 0:
            raise _exceptions.PETScKSPError(int(res.reason))
        return self.last_result
