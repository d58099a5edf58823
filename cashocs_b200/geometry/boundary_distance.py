"""Distance to the boundary (``cashocs/geometry/boundary_distance.py:40-282``) on the device.

``compute_boundary_distance(mesh, boundary_idcs, tol, max_iter, method="eikonal")``: the iterated linearisation of the
eikonal equation ``|grad u| = 1`` (``:191-282``): start from ``-Laplace u = 1``, then repeat
``int grad u . grad v = int (grad u_prev / |grad u_prev|) . grad v`` until ``int (|grad u| - 1)^2`` has dropped by
``tol`` or ``max_iter`` sweeps are done; ``u = 0`` on the chosen boundaries (all boundary vertices when none are
given).  One stiffness matrix (``cb_assemble_scalar``) and one AMG hierarchy serve every sweep; the right-hand side
and the residual are the kernels of ``csrc/distance.cu``.  The Poisson variant (``:104-188``) integrates
``sqrt(|grad u|^2 + 2 u)`` with a quadrature rule chosen by the form compiler and is not built.
"""

from __future__ import annotations

import copy

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200._utils import linalg


class EikonalDistance:
    """Reusable state of ``compute_boundary_distance_eikonal`` for one mesh topology and boundary selection."""

    def __init__(self, mesh, boundary_idcs=None, ksp_options=None, linear_solver=None) -> None:
        self.mesh = mesh
        mesh.build_pattern()
        dev = mesh.device
        nv = mesh.num_vertices
        if boundary_idcs is not None and len(boundary_idcs) > 0:
            verts = mesh.boundary_vertices([int(m) for m in boundary_idcs])
        else:  # fenics.CompiledSubDomain("on_boundary"): every vertex of a boundary facet
            facets = mesh.boundary_facet_table()[0]
            verts = torch.unique(facets.reshape(-1).long())
            if mesh.dist is not None:  # facets of the partition: add the tagged boundary vertices this rank sees
                verts = torch.unique(torch.cat([verts, mesh.boundary_vertices(mesh.boundary_markers)]))
        self.bc_flag = torch.zeros(nv, dtype=torch.uint8, device=dev)
        self.bc_flag[verts] = 1
        self.bc_val = torch.zeros(nv, dtype=torch.float64, device=dev)
        # boundary_distance.py:222: the reference's iterative default (cg + boomeramg)
        self.ksp_options = ksp_options if ksp_options is not None else copy.deepcopy(linalg.iterative_ksp_options)
        self.linear_solver = linear_solver or linalg.LinearSolver()
        self.A = linalg.Matrix(mesh, 1)
        f64 = dict(dtype=torch.float64, device=dev)
        self.b = torch.zeros(mesh.n_owned_vertices, **f64)
        self.u_curr, self.u_prev = torch.zeros(nv, **f64), torch.zeros(nv, **f64)
        self._ones = torch.ones(nv, **f64)
        self._out = torch.zeros(1, **f64)
        self.sweeps = 0
        self.iterations = 0
        self.residuals: list[float] = []

    def _solve(self) -> None:
        self.u_curr.zero_()
        res = self.linear_solver.solve(self.u_curr, A=self.A, b=self.b, ksp_options=self.ksp_options)
        self.iterations += res.its
        if self.mesh.dist is not None:
            self.mesh.dist.halo_exchange(self.u_curr, 1)

    def _residual(self) -> float:
        mesh = self.mesh
        scratch = mesh.reduce_scratch()
        _lib.check(
            _lib.load().cb_eikonal_residual(_lib.stream_ptr(), mesh.dim, mesh.n_owned_cells, _lib.ptr(mesh.coords),
                                            _lib.ptr(mesh.cells), _lib.ptr(self.u_curr), _lib.ptr(self._out),
                                            _lib.ptr(scratch), scratch.numel()),
            "cb_eikonal_residual",
        )
        if mesh.dist is not None:
            mesh.dist.allreduce(self._out, "sum")
        return float(np.sqrt(self._out.item()))

    def compute(self, tol: float = 1e-1, max_iter: int = 10) -> torch.Tensor:
        mesh = self.mesh
        self.iterations = 0
        # -Laplace u = 1 (boundary_distance.py:255-258)
        linalg.assemble_scalar_system(mesh, 1.0, 0.0, f_nodal=self._ones, f_nodal_scale=1.0, bc_flag=self.bc_flag,
                                      bc_val=self.bc_val, A=self.A, b=self.b)
        self._solve()
        res_0 = self._residual()
        self.residuals = [res_0]
        self.sweeps = 0
        for _ in range(max_iter):
            self.u_prev.copy_(self.u_curr)
            _lib.check(
                _lib.load().cb_eikonal_rhs(_lib.stream_ptr(), mesh.dim, mesh.n_owned_vertices, _lib.ptr(mesh.v2c_ptr),
                                           _lib.ptr(mesh.v2c_idx), _lib.ptr(mesh.cells), _lib.ptr(mesh.coords),
                                           _lib.ptr(self.u_prev), _lib.ptr(self.bc_flag), _lib.ptr(self.b)),
                "cb_eikonal_rhs",
            )
            self._solve()
            self.sweeps += 1
            res = self._residual()
            self.residuals.append(res)
            if res <= res_0 * tol:
                break
        return self.u_curr


def compute_boundary_distance(mesh, boundaries=None, boundary_idcs=None, tol: float = 1e-1, max_iter: int = 10,
                              minimum_distance: float = 0.0, method: str = "eikonal") -> torch.Tensor:
    """``cashocs.geometry.compute_boundary_distance`` (``boundary_distance.py:40-101``); ``boundaries`` is accepted
    for signature compatibility (the device mesh carries its own markers).  Returns the P1 vertex values."""
    if method == "eikonal":
        return EikonalDistance(mesh, boundary_idcs).compute(tol=tol, max_iter=max_iter).clone()
    if method == "poisson":
        raise _exceptions.InputError("compute_boundary_distance", "method",
                                     "the Poisson approach is not built on the device, use 'eikonal'")
    raise _exceptions.InputError("compute_boundary_distance", "method", "The method can only be 'poisson' or 'eikonal'.")
