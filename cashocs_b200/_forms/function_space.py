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
of the states and adjoints.  On a partitioned mesh the space carries its own halo plan over vertex +
edge dofs (``p2_numbering``).
"""

from __future__ import annotations

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


def p2_numbering(mesh):
    """Dof numbering of the scalar P2 space on a (partition of a) mesh -- pure index plumbing, runs on
    CPU tensors too (``tests/test_distributed_cpu.py``).

    Local edges are the unique vertex pairs of the local cells.  One GPU: dof = vertex, then
    ``nv + edge``.  Partitioned mesh: an edge is owned by the owner of its end point with the smaller
    *global* id, so every cell around an owned edge is a local cell (local cells = all cells touching an
    owned vertex) and owned rows assemble without communication.  Numbering: owned vertices, owned
    edges (by global key), then the ghosts grouped by owner -- per owner its ghost vertices (vertex-plan
    order) followed by its ghost edges (by global key) -- which is the layout the halo plan needs.

    Returns a dict: ``edges`` [ne, 2] (local vertex ids, lower first), ``cell_dofs`` [nc, nde] int32,
    ``vertex_dof`` [nv], ``edge_dof`` [ne], ``ndofs``, ``n_owned``, ``dist`` (dof halo plan or None) and,
    on a partitioned mesh, ``edge_gkey`` [ne, 2] (global end points, lower first).
    """
    import torch.distributed as td

    d = mesh.dim
    k = d + 1
    loc = LOCAL_EDGES[d]
    nv, nc = mesh.num_vertices, mesh.num_cells
    dev = mesh.cells.device
    cells = mesh.cells.long()
    a = torch.stack([cells[:, i] for i, _ in loc], dim=1)
    b = torch.stack([cells[:, j] for _, j in loc], dim=1)
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    uniq, inv = torch.unique((lo * nv + hi).reshape(-1), return_inverse=True)
    edges = torch.stack([uniq // nv, uniq % nv], dim=1)
    ne = int(uniq.numel())
    cell_edges = inv.view(nc, len(loc))
    dist = mesh.dist
    if dist is None:
        vertex_dof = torch.arange(nv, device=dev)
        edge_dof = nv + torch.arange(ne, device=dev)
        out = dict(edges=edges, vertex_dof=vertex_dof, edge_dof=edge_dof, ndofs=nv + ne, n_owned=nv + ne, dist=None)
    else:
        from cashocs_b200 import distributed

        n_ov = mesh.n_owned_vertices
        gid = mesh.global_vertex_ids.to(dev)
        ga, gb = gid[edges[:, 0]], gid[edges[:, 1]]
        glo, ghi = torch.minimum(ga, gb), torch.maximum(ga, gb)
        off_t = torch.as_tensor(mesh.global_offsets, device=dev)
        owner = torch.searchsorted(off_t, glo, right=True) - 1
        nglob = int(off_t[-1].item())
        gkey = glo * nglob + ghi
        mine = owner == dist.rank
        own_idx = torch.nonzero(mine).reshape(-1)
        own_idx = own_idx[torch.argsort(gkey[own_idx])]
        n_oe = int(own_idx.numel())
        # ghost edges per neighbour (sorted by global key); what I need from each owner
        need = {}
        ghost_idx = {}
        for r in torch.unique(owner[~mine]).tolist():
            idx = torch.nonzero(owner == r).reshape(-1)
            idx = idx[torch.argsort(gkey[idx])]
            ghost_idx[int(r)] = idx
            need[int(r)] = gkey[idx].cpu()
        all_need = [None] * dist.world
        td.all_gather_object(all_need, need)
        neigh = sorted(set(dist.neigh) | set(need) | {r for r in range(dist.world) if r != dist.rank and dist.rank in all_need[r]})
        own_keys = gkey[own_idx]
        vertex_dof = torch.empty(nv, dtype=torch.int64, device=dev)
        vertex_dof[:n_ov] = torch.arange(n_ov, device=dev)
        edge_dof = torch.empty(ne, dtype=torch.int64, device=dev)
        edge_dof[own_idx] = n_ov + torch.arange(n_oe, device=dev)
        n_owned = n_ov + n_oe
        send_lists, recv_counts = [], []
        cursor = n_owned
        vneigh = {int(r): j for j, r in enumerate(dist.neigh)}
        empty = torch.zeros(0, dtype=torch.int64, device=dev)
        for r in neigh:
            # what r needs from me: my owned vertices in its ghost list, then my owned edges it asked for
            if r in vneigh:
                j = vneigh[r]
                sv = dist.send_idx[int(dist.send_off[j]): int(dist.send_off[j + 1])].long()
                gv = n_ov + torch.arange(int(dist.recv_off[j]), int(dist.recv_off[j + 1]), device=dev)
            else:
                sv, gv = empty, empty
            want = all_need[r].get(dist.rank)
            if want is not None and want.numel():
                pos = torch.searchsorted(own_keys, want.to(dev))
                if bool((own_keys[pos.clamp(max=max(n_oe - 1, 0))] != want.to(dev)).any()):
                    raise _exceptions.CashocsException("P2 halo plan: a requested edge is not owned here")
                se = n_ov + pos
            else:
                se = empty
            send_lists.append(torch.cat([sv, se]))
            ge = ghost_idx.get(r, empty)
            vertex_dof[gv] = cursor + torch.arange(gv.numel(), device=dev)
            cursor += int(gv.numel())
            edge_dof[ge] = cursor + torch.arange(ge.numel(), device=dev)
            cursor += int(ge.numel())
            recv_counts.append(int(gv.numel() + ge.numel()))
        ndofs = nv + ne
        assert cursor == ndofs
        plan = distributed.DistContext(dist.rank, dist.world, dist.transport)
        plan.parent = dist if dist.parent is None else dist.parent
        plan.set_plan(neigh, send_lists, recv_counts, n_owned, ndofs - n_owned, dev)
        out = dict(edges=edges, vertex_dof=vertex_dof, edge_dof=edge_dof, ndofs=ndofs, n_owned=n_owned, dist=plan,
                   edge_gkey=torch.stack([glo, ghi], dim=1))
    out["cell_dofs"] = torch.cat([out["vertex_dof"][cells], out["edge_dof"][cell_edges]], dim=1).to(torch.int32).contiguous()
    return out


class P2Space:
    degree = 2

    def __init__(self, mesh) -> None:
        _lib.require_cuda(mesh.coords, mesh.cells)
        lib = _lib.load()
        s = _lib.stream_ptr()
        self.mesh = mesh
        self.device = dev = mesh.device
        self.dim = d = mesh.dim
        k = d + 1
        self.ndof_e = nde = k + len(LOCAL_EDGES[d])
        nc = mesh.num_cells
        num = p2_numbering(mesh)
        self.edges = num["edges"]              # [ne, 2] local vertex ids
        self.vertex_dof = num["vertex_dof"]    # local vertex -> dof
        self.edge_dof = num["edge_dof"]        # local edge -> dof
        self.cell_dofs = num["cell_dofs"]
        self.num_edges = int(self.edges.shape[0])
        self.ndofs = int(num["ndofs"])          # owned + ghost dofs of this rank
        self.n_owned_dofs = int(num["n_owned"])
        self.dist = num["dist"]                 # halo plan over dofs (None on one GPU)
        if self.ndofs >= 2**31:
            raise _exceptions.CashocsException("P2 dof count exceeds int32")
        # mesh-like attributes for Matrix / LinearSolver / AMG
        self.num_vertices, self.n_owned_vertices = self.ndofs, self.n_owned_dofs
        nrows = self.n_owned_dofs
        self.d2c_ptr = torch.empty(self.ndofs + 1, dtype=torch.int64, device=dev)
        self.d2c_idx = torch.empty(nc * nde, dtype=torch.int32, device=dev)
        _lib.check(lib.cb_v2c_build(s, self.ndofs, nc, nde, _lib.ptr(self.cell_dofs), _lib.ptr(self.d2c_ptr),
                                    _lib.ptr(self.d2c_idx)), "cb_v2c_build")
        self.rowptr = torch.empty(nrows + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.cb_pattern_rowptr(s, nrows, nde, _lib.ptr(self.cell_dofs), _lib.ptr(self.d2c_ptr),
                                         _lib.ptr(self.d2c_idx), _lib.ptr(self.rowptr)), "cb_pattern_rowptr")
        self.nnz = int(self.rowptr[-1].item())
        if self.nnz >= 2**31:
            raise _exceptions.CashocsException("P2 nnz exceeds the int32 column-index range of one GPU")
        self.col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        _lib.check(lib.cb_pattern_fill(s, nrows, nc, nde, _lib.ptr(self.cell_dofs), _lib.ptr(self.d2c_ptr),
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
        v.nrows_owned, v.nv_owned, v.nc_owned = self.n_owned_dofs, m.n_owned_vertices, m.n_owned_cells
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
        verts = self.vertex_dof[m.boundary_vertices(markers)]
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
        return torch.cat([verts, self.edge_dof[eid]])

    def interpolate(self, expr) -> torch.Tensor:
        """Nodal interpolant (vertex values, then edge-midpoint values) of a polynomial / callable."""
        x = self.mesh.coords
        pts = torch.empty(self.ndofs, x.shape[1], dtype=torch.float64, device=x.device)
        pts[self.vertex_dof] = x
        pts[self.edge_dof] = 0.5 * (x[self.edges[:, 0]] + x[self.edges[:, 1]])
        if isinstance(expr, expressions.Polynomial):
            out = torch.zeros(pts.shape[0], dtype=torch.float64, device=pts.device)
            for (a, b, c), v in expr.terms.items():
                t = v * pts[:, 0] ** a * pts[:, 1] ** b
                if pts.shape[1] == 3:
                    t = t * pts[:, 2] ** c
                out += t
            return out
        return expr(pts).reshape(-1)
