"""P2 (quadratic Lagrange) function space on a device mesh -- BASELINE ``configs[4]``.

Host mirror of ``fenics.FunctionSpace(mesh, "CG", 2)`` as far as the hot path needs it: dof
numbering (vertex ``v`` -> ``v``, edge ``e`` -> ``Nv + e``; dolfin's own numbering is a graph
reordering that cannot be reproduced, SURVEY.md 9.1 -- the contract is the cell->dof table), the
CSR pattern of the dofmap graph (``bilinear_boundary_form_modification``,
``cashocs/_utils/forms.py:274-288``), the dof->cell adjacency the gather kernels walk, Dirichlet dofs
by a topological search over tagged boundary facets, and the reference tensors of the affine P2
element.  The object quacks like a mesh for :class:`cashocs_b200.Matrix` (``rowptr``, ``col``,
``nnz``, ``n_owned_vertices`` = ``num_vertices`` = number of dofs).

The deformation space stays vector P1 (the reference forces it,
``cashocs/_optimization/shape_optimization/shape_optimization_problem.py:251-256``); P2 is the space
of the states and adjoints.  One GPU for now (the partitioned P2 dofmap is the next row).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200._forms import expressions

LOCAL_EDGES = {2: ((0, 1), (0, 2), (1, 2)), 3: ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))}


def p2_basis(lam: np.ndarray):
    """phi [nq, nd], dphi/dlam [nq, nd, d+1] at barycentric points lam [nq, d+1]."""
    nq, k = lam.shape
    loc = LOCAL_EDGES[k - 1]
    nd = k + len(loc)
    phi = np.zeros((nq, nd))
    dphi = np.zeros((nq, nd, k))
    for i in range(k):
        phi[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
        dphi[:, i, i] = 4.0 * lam[:, i] - 1.0
    for e, (i, j) in enumerate(loc):
        phi[:, k + e] = 4.0 * lam[:, i] * lam[:, j]
        dphi[:, k + e, i] = 4.0 * lam[:, j]
        dphi[:, k + e, j] = 4.0 * lam[:, i]
    return phi, dphi


class P2Space:
    degree = 2

    def __init__(self, mesh) -> None:
        if mesh.dist is not None:
            raise _exceptions.InputError("cashocs_b200.P2Space", "mesh",
                                         "P2 spaces on a partitioned mesh are not implemented yet (one GPU)")
        _lib.require_cuda(mesh.coords, mesh.cells)
        lib = _lib.load()
        s = _lib.stream_ptr()
        self.mesh = mesh
        self.device = dev = mesh.device
        self.dim = d = mesh.dim
        self.dist = None
        k = d + 1
        loc = LOCAL_EDGES[d]
        self.ndof_e = nde = k + len(loc)
        nv, nc = mesh.num_vertices, mesh.num_cells
        cells = mesh.cells.long()
        # edges = sorted unique vertex pairs (cells are ascending per cell, so pairs are ordered)
        key = torch.stack([cells[:, i] * nv + cells[:, j] for i, j in loc], dim=1)
        uniq, inv = torch.unique(key.reshape(-1), return_inverse=True)
        self.edges = torch.stack([uniq // nv, uniq % nv], dim=1)
        self.num_edges = int(uniq.numel())
        self.ndofs = nv + self.num_edges
        if self.ndofs >= 2**31:
            raise _exceptions.CashocsException("P2 dof count exceeds int32")
        self.cell_dofs = torch.cat([cells, nv + inv.view(nc, len(loc))], dim=1).to(torch.int32).contiguous()
        del key, inv, cells
        # mesh-like attributes for Matrix / LinearSolver / AMG
        self.num_vertices = self.n_owned_vertices = self.ndofs
        self.d2c_ptr = torch.empty(self.ndofs + 1, dtype=torch.int64, device=dev)
        self.d2c_idx = torch.empty(nc * nde, dtype=torch.int32, device=dev)
        _lib.check(lib.cb_v2c_build(s, self.ndofs, nc, nde, _lib.ptr(self.cell_dofs), _lib.ptr(self.d2c_ptr),
                                    _lib.ptr(self.d2c_idx)), "cb_v2c_build")
        self.rowptr = torch.empty(self.ndofs + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.cb_pattern_rowptr(s, self.ndofs, nde, _lib.ptr(self.cell_dofs), _lib.ptr(self.d2c_ptr),
                                         _lib.ptr(self.d2c_idx), _lib.ptr(self.rowptr)), "cb_pattern_rowptr")
        self.nnz = int(self.rowptr[-1].item())
        if self.nnz >= 2**31:
            raise _exceptions.CashocsException("P2 nnz exceeds the int32 column-index range of one GPU")
        self.col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        _lib.check(lib.cb_pattern_fill(s, self.ndofs, nc, nde, _lib.ptr(self.cell_dofs), _lib.ptr(self.d2c_ptr),
                                       _lib.ptr(self.d2c_idx), _lib.ptr(self.rowptr), _lib.ptr(self.col), None),
                   "cb_pattern_fill")
        # reference tensors of the affine element (degree-4 exact rule)
        lam, w = expressions.simplex_quadrature(d, 4)
        phi, dphi = p2_basis(lam[:, :k])
        f64 = dict(dtype=torch.float64, device=dev)
        self.ref_R = torch.tensor(np.einsum("q,qak,qbl->abkl", w, dphi, dphi), **f64).contiguous()
        self.ref_M = torch.tensor(np.einsum("q,qa,qb->ab", w, phi, phi), **f64).contiguous()
        self.ref_I = torch.tensor(np.einsum("q,qa->a", w, phi), **f64).contiguous()
        self.cell_scratch = torch.empty(nc * max(nde, k * d), **f64)
        self._qtabs: dict = {}
        self.scratch = None
        torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ mesh-like plumbing
    def build_pattern(self) -> None:
        pass

    def reduce_scratch(self) -> torch.Tensor:
        return self.mesh.reduce_scratch()

    def view(self) -> _lib.cb_p2_view:
        v = _lib.cb_p2_view()
        m = self.mesh
        v.dim, v.ndof_e = self.dim, self.ndof_e
        v.nc, v.nv, v.ndofs, v.nnz = m.num_cells, m.num_vertices, self.ndofs, self.nnz
        v.cells, v.cell_dofs, v.coords = m.cells.data_ptr(), self.cell_dofs.data_ptr(), m.coords.data_ptr()
        v.rowptr, v.col = self.rowptr.data_ptr(), self.col.data_ptr()
        v.d2c_ptr, v.d2c_idx = self.d2c_ptr.data_ptr(), self.d2c_idx.data_ptr()
        v.ref_R, v.ref_M, v.ref_I = self.ref_R.data_ptr(), self.ref_M.data_ptr(), self.ref_I.data_ptr()
        v.cell_scratch = self.cell_scratch.data_ptr()
        return v

    def qtab(self, degree: int):
        """Device table [nq, (d+1) + 1 + ndof_e] of a rule exact for ``degree``: point, weight, basis."""
        if degree not in self._qtabs:
            k = self.dim + 1
            lam, w = expressions.simplex_quadrature(self.dim, degree)
            phi, _ = p2_basis(lam[:, :k])
            tab = np.concatenate([lam[:, :k], w[:, None], phi], axis=1)
            self._qtabs[degree] = (len(w), torch.tensor(tab, dtype=torch.float64, device=self.device).contiguous())
        return self._qtabs[degree]

    # ------------------------------------------------------------------ boundary dofs
    def boundary_dofs(self, markers) -> torch.Tensor:
        """Vertex + edge dofs on the boundary facets carrying ``markers`` (topological search, like
        dolfin's DirichletBC).  Meshes without stored facets (the synthetic boxes, whose faces are
        planar and convex) use: an edge lies on a marked face iff both end points do."""
        m = self.mesh
        nv = m.num_vertices
        verts = m.boundary_vertices(markers)
        facets = getattr(m, "boundary_facets", None)
        ekey = self.edges[:, 0] * nv + self.edges[:, 1]
        if facets:
            parts = [facets[int(k)] for k in (markers if not isinstance(markers, int) else [markers]) if int(k) in facets]
            if not parts:
                return torch.zeros(0, dtype=torch.int64, device=self.device)
            f = torch.sort(torch.cat(parts).long(), dim=1).values
            pairs = [(0, 1)] if self.dim == 2 else [(0, 1), (0, 2), (1, 2)]
            fkey = torch.unique(torch.cat([f[:, i] * nv + f[:, j] for i, j in pairs]))
            eid = torch.searchsorted(ekey, fkey)
        else:
            if isinstance(markers, int):
                markers = [markers]
            eids = []
            for mk in markers:  # per marker: both end points on the same face
                on = torch.zeros(nv, dtype=torch.bool, device=self.device)
                on[m.boundary_vertices([mk])] = True
                eids.append(torch.nonzero(on[self.edges[:, 0]] & on[self.edges[:, 1]]).reshape(-1))
            eid = torch.unique(torch.cat(eids)) if eids else torch.zeros(0, dtype=torch.int64, device=self.device)
        return torch.cat([verts, nv + eid])

    def interpolate(self, expr) -> torch.Tensor:
        """Nodal interpolant (vertex values, then edge-midpoint values) of a polynomial / callable."""
        x = self.mesh.coords
        mid = 0.5 * (x[self.edges[:, 0]] + x[self.edges[:, 1]])
        pts = torch.cat([x, mid])
        if isinstance(expr, expressions.Polynomial):
            out = torch.zeros(pts.shape[0], dtype=torch.float64, device=pts.device)
            for (a, b, c), v in expr.terms.items():
                t = v * pts[:, 0] ** a * pts[:, 1] ** b
                if pts.shape[1] == 3:
                    t = t * pts[:, 2] ** c
                out += t
            return out
        return expr(pts).reshape(-1)
