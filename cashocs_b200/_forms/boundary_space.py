"""The P1 space on (a part of) the boundary: compressed vertex numbering, CSR pattern of the facet graph, the facet
mass matrix ``int_Gamma phi_a phi_b ds`` and L2 projections with it.

The reference assembles such forms on ALL vertices of the mesh and lets ``ident_zeros`` turn the interior rows into
identities (``cashocs/_forms/shape_regularization.py:596-613`` for the curvature,
``cashocs/_pde_problems/shape_gradient_problem.py:225-299`` for the normal deformation); the solutions vanish there,
so the device solves on the boundary vertices only.  Used by the curvature regularisation and by the re-extension mode
``normal``.
"""

from __future__ import annotations

import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200._utils import linalg


class _PatternHolder:
    """What ``linalg.Matrix`` / ``LinearSolver`` read from a mesh."""

    def __init__(self, mesh, nb: int, rowptr: torch.Tensor, col: torch.Tensor, nnz: int) -> None:
        self._mesh = mesh
        self.dim = mesh.dim
        self.device = mesh.device
        self.dist = None
        self.n_owned_vertices = self.num_vertices = nb
        self.rowptr, self.col, self.nnz = rowptr, col, nnz

    def build_pattern(self) -> None:
        return None

    def reduce_scratch(self) -> torch.Tensor:
        return self._mesh.reduce_scratch()


class BoundarySpace:
    """``facets`` [nf, d] int32 (mesh vertex ids) -> the P1 space on their vertices."""

    def __init__(self, mesh, facets: torch.Tensor, what: str = "boundary forms") -> None:
        if mesh.dist is not None:
            raise _exceptions.InputError("cashocs_b200.BoundarySpace", what, "runs on unpartitioned meshes")
        self.mesh = mesh
        d, dev = mesh.dim, mesh.device
        lib, s = _lib.load(), _lib.stream_ptr()
        self.facets = facets.to(torch.int32).contiguous()
        long_facets = self.facets.long()
        self.vertices = torch.unique(long_facets.reshape(-1))  # sorted mesh vertex ids
        self.vertices32 = self.vertices.to(torch.int32).contiguous()
        nb, nf = int(self.vertices.numel()), int(self.facets.shape[0])
        self.n, self.nf = nb, nf
        self.facets_c = torch.searchsorted(self.vertices, long_facets.reshape(-1)).to(torch.int32).reshape(nf, d).contiguous()
        # vertex -> facet adjacency in both numberings: compressed rows for the matrix, mesh rows for the vector gathers
        self.cv2f_ptr = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
        self.cv2f_idx = torch.zeros(max(nf * d, 1), dtype=torch.int32, device=dev)
        self.v2f_ptr = torch.zeros(mesh.num_vertices + 1, dtype=torch.int64, device=dev)
        self.v2f_idx = torch.zeros(max(nf * d, 1), dtype=torch.int32, device=dev)
        rowptr = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
        nnz = 0
        col = torch.zeros(0, dtype=torch.int32, device=dev)
        if nf:
            _lib.check(lib.cb_v2c_build(s, nb, nf, d, _lib.ptr(self.facets_c), _lib.ptr(self.cv2f_ptr),
                                        _lib.ptr(self.cv2f_idx)), "cb_v2c_build")
            _lib.check(lib.cb_v2c_build(s, mesh.num_vertices, nf, d, _lib.ptr(self.facets), _lib.ptr(self.v2f_ptr),
                                        _lib.ptr(self.v2f_idx)), "cb_v2c_build")
            _lib.check(lib.cb_pattern_rowptr(s, nb, d, _lib.ptr(self.facets_c), _lib.ptr(self.cv2f_ptr),
                                             _lib.ptr(self.cv2f_idx), _lib.ptr(rowptr)), "cb_pattern_rowptr")
            nnz = int(rowptr[-1].item())
            col = torch.empty(nnz, dtype=torch.int32, device=dev)
            f2n = torch.empty(nf * d * d, dtype=torch.int32, device=dev)
            _lib.check(lib.cb_pattern_fill(s, nb, nf, d, _lib.ptr(self.facets_c), _lib.ptr(self.cv2f_ptr),
                                           _lib.ptr(self.cv2f_idx), _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(f2n)),
                       "cb_pattern_fill")
        self.holder = _PatternHolder(mesh, nb, rowptr, col, nnz)
        self.mass = linalg.Matrix(self.holder, 1)
        f64 = dict(dtype=torch.float64, device=dev)
        self.rhs = torch.zeros(nb, **f64)
        self.sol = torch.zeros(nb, **f64)
        self.facet_vec = torch.zeros(max(nf * d * d, 1), **f64)
        self.solver = linalg.LinearSolver()
        self.iterations = 0

    def assemble_mass(self) -> linalg.Matrix:
        """``int_Gamma phi_a phi_b ds`` for the current coordinates (``cb_boundary_mass``)."""
        A, mesh = self.mass, self.mesh
        _lib.check(
            _lib.load().cb_boundary_mass(_lib.stream_ptr(), mesh.dim, A.nrows, _lib.ptr(mesh.coords), _lib.ptr(self.vertices32),
                                         _lib.ptr(self.facets_c), _lib.ptr(self.cv2f_ptr), _lib.ptr(self.cv2f_idx),
                                         _lib.ptr(A.rowptr), _lib.ptr(A.col), _lib.ptr(A.val)),
            "cb_boundary_mass",
        )
        A.touch()
        return A

    def project(self, rhs_on_mesh: torch.Tensor, out_on_mesh: torch.Tensor, ksp_options=None) -> torch.Tensor:
        """Solve ``M x = rhs`` for one scalar component given / returned as (strided views of) mesh-vertex vectors; the
        mass matrix must be current (:meth:`assemble_mass`).  ``ksp_options = None``: the reference's default direct
        solver, i.e. the device's tight CG (``_utils/linalg.py``).  Returns the compressed solution (a scratch buffer)."""
        torch.index_select(rhs_on_mesh, 0, self.vertices, out=self.rhs)
        res = self.solver.solve(self.sol, A=self.mass, b=self.rhs, ksp_options=ksp_options)
        self.iterations += res.its
        out_on_mesh.index_copy_(0, self.vertices, self.sol)
        return self.sol
