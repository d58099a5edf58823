"""Device-resident simplicial meshes: synthetic boxes, imported meshes, P1 sparsity data.

Host mirror of ``cashocs/geometry/mesh.py`` (``regular_mesh`` :155-217, ``regular_box_mesh``
:221-391) and of the subset of ``cashocs.import_mesh`` (``cashocs/io/mesh.py:186-293``) the hot
path needs.  All arrays live in HBM as torch tensors (plumbing); the sparsity pattern, the
cell->nnz scatter map and the cell colouring are built by the C-ABI (``csrc/pattern.cu``).

Layout in HBM
-------------
``coords``  [Nv, d] fp64 (AoS: one vertex = 16/24 contiguous bytes, one or two sectors per gather)
``cells``   [Nc, d+1] int32, ascending per cell (dolfin ``mesh.order()``)
``rowptr``  [Nrows+1] int64, ``col`` [nnz] int32   -- P1 dofmap graph (scalar dof == vertex;
            the vector space uses the same graph with d x d blocks, dof = d*vertex + comp)
``ccells`` / ``cc2n``  colour-sorted copies of ``cells`` and of the [Nc, (d+1)^2] cell->nnz map,
            ``colour_off`` (host) the start of every colour
"""

from __future__ import annotations


import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib


class Mesh:
    """A P1 simplicial mesh on one GPU (or one partition of a distributed mesh)."""

    def __init__(
        self,
        coords: torch.Tensor,
        cells: torch.Tensor,
        boundary_vertices: dict[int, torch.Tensor],
        n_owned_vertices: int | None = None,
        n_owned_cells: int | None = None,
        dist=None,
    ) -> None:
        self.coords = coords.contiguous()
        self.cells = cells.to(torch.int32).contiguous()
        self.device = self.coords.device
        self.dim = int(self.coords.shape[1])
        self.num_vertices = int(self.coords.shape[0])
        self.num_cells = int(self.cells.shape[0])
        # distributed: owned vertices / cells come first, ghosts after
        self.n_owned_vertices = self.num_vertices if n_owned_vertices is None else int(n_owned_vertices)
        self.n_owned_cells = self.num_cells if n_owned_cells is None else int(n_owned_cells)
        self.dist = dist
        self._boundary_vertices = {int(k): v.to(torch.int64) for k, v in boundary_vertices.items()}
        self.coords_old: torch.Tensor | None = None
        self._pattern_ready = False
        self.scratch = None
        self.cell_scratch = None
        self.cell_geo = None
        self.boundary_facets: dict[int, torch.Tensor] | None = None  # tagged facets (imported meshes)
        self._p2_space = None

    # ------------------------------------------------------------------ boundary data
    def boundary_vertices(self, markers) -> torch.Tensor:
        """Sorted unique vertices on boundary facets carrying one of ``markers``.

        Equivalent of dolfin's topological DirichletBC search used by
        ``create_dirichlet_bcs`` (``cashocs/_utils/forms.py:214-271``).
        """
        if isinstance(markers, (int, np.integer)):
            markers = [int(markers)]
        parts = [self._boundary_vertices[int(m)] for m in markers if int(m) in self._boundary_vertices]
        if not parts:
            return torch.zeros(0, dtype=torch.int64, device=self.device)
        return torch.unique(torch.cat(parts))

    def p2_space(self):
        """The scalar P2 space on this mesh (built once per mesh topology)."""
        if self._p2_space is None:
            from cashocs_b200._forms import function_space

            self._p2_space = function_space.P2Space(self)
        return self._p2_space

    def boundary_facets_local(self):
        """Boundary facets seen by this rank, from the topology alone: ``(facets [nf, d] int64, counted [nf] bool)``.

        A facet lies on the boundary iff it belongs to exactly one cell.  On a partition (all cells touching an owned
        vertex are local) that is decidable for every facet with at least one owned vertex -- both its cells touch
        that vertex -- and those are exactly the facets whose measure gradients reach owned vertices.  ``counted``
        marks the facets this rank adds to the global boundary measure: the owner of the facet's vertex with the
        lowest global id counts it, so every boundary facet is counted exactly once."""
        d = self.dim
        cells = self.cells.long()
        keep = [[j for j in range(d + 1) if j != omit] for omit in range(d + 1)]
        sub = torch.sort(torch.cat([cells[:, k] for k in keep], dim=0), dim=1).values  # a no-op for ascending cells
        uniq, counts = torch.unique(sub, dim=0, return_counts=True)
        facets = uniq[counts == 1]
        if self.dist is None:
            return facets, torch.ones(facets.shape[0], dtype=torch.bool, device=facets.device)
        n_own = self.n_owned_vertices
        facets = facets[(facets < n_own).any(dim=1)]
        gid = torch.as_tensor(self.global_vertex_ids, device=facets.device).long()
        first = torch.gather(facets, 1, torch.argmin(gid[facets], dim=1, keepdim=True)).squeeze(1)
        return facets, first < n_own

    def boundary_facet_table(self):
        """The ``ds`` measure on the device: ``(facets [nf, d] int32, v2f_ptr, v2f_idx, counted)`` -- the boundary facets
        of :meth:`boundary_facets_local`, their vertex -> facet adjacency (rows of the owned vertices are used) and the
        facets this rank counts in the global measure (all of them on one GPU).  Built once per topology."""
        if getattr(self, "_facet_table", None) is not None:
            return self._facet_table
        d = self.dim
        facets64, counted = self.boundary_facets_local()
        facets = facets64.to(torch.int32).contiguous()
        nf = int(facets.shape[0])
        n = self.num_vertices
        ptr = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        idx = torch.zeros(max(nf * d, 1), dtype=torch.int32, device=self.device)
        if nf:
            _lib.check(_lib.load().cb_v2c_build(_lib.stream_ptr(), n, nf, d, _lib.ptr(facets), _lib.ptr(ptr), _lib.ptr(idx)),
                       "cb_v2c_build")
        self._facet_table = (facets, ptr, idx, counted)
        return self._facet_table

    @property
    def boundary_markers(self) -> list[int]:
        return sorted(self._boundary_vertices)

    # ------------------------------------------------------------------ sparsity pattern
    def reduce_scratch(self) -> torch.Tensor:
        if self.scratch is None:
            n = int(_lib.load().cb_reduce_scratch_len())
            self.scratch = torch.zeros(n, dtype=torch.float64, device=self.device)
        return self.scratch

    def build_pattern(self) -> None:
        """Build v2c adjacency, CSR pattern, cell->nnz map and the colour-sorted cell arrays."""
        if self._pattern_ready:
            return
        lib = _lib.load()
        _lib.require_cuda(self.coords, self.cells)
        k = self.dim + 1
        nv, nc, nrows = self.num_vertices, self.num_cells, self.n_owned_vertices
        dev = self.device
        s = _lib.stream_ptr()
        self.v2c_ptr = torch.empty(nv + 1, dtype=torch.int64, device=dev)
        self.v2c_idx = torch.empty(nc * k, dtype=torch.int32, device=dev)
        _lib.check(lib.cb_v2c_build(s, nv, nc, k, _lib.ptr(self.cells), _lib.ptr(self.v2c_ptr),
                                    _lib.ptr(self.v2c_idx)), "cb_v2c_build")
        self.rowptr = torch.empty(nrows + 1, dtype=torch.int64, device=dev)
        _lib.check(lib.cb_pattern_rowptr(s, nrows, k, _lib.ptr(self.cells), _lib.ptr(self.v2c_ptr),
                                         _lib.ptr(self.v2c_idx), _lib.ptr(self.rowptr)), "cb_pattern_rowptr")
        self.nnz = int(self.rowptr[-1].item())
        if self.nnz >= 2**31:
            raise _exceptions.CashocsException("block nnz exceeds int32 scatter-map range")
        self.col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        cell2nnz = torch.empty(nc * k * k, dtype=torch.int32, device=dev)
        _lib.check(lib.cb_pattern_fill(s, nrows, nc, k, _lib.ptr(self.cells), _lib.ptr(self.v2c_ptr),
                                       _lib.ptr(self.v2c_idx), _lib.ptr(self.rowptr), _lib.ptr(self.col),
                                       _lib.ptr(cell2nnz)), "cb_pattern_fill")
        # IntersectionTester occurrences (cashocs/geometry/mesh_testing.py:148-158)
        self.valence = (self.v2c_ptr[1:] - self.v2c_ptr[:-1]).to(torch.int32).contiguous()
        # deterministic greedy colouring on the host, then colour-sorted device copies
        cells_h = self.cells.cpu().numpy()
        colour_h = np.empty(nc, dtype=np.int32)
        ncol = lib.cb_colour_cells_host(nv, nc, k, cells_h.ctypes.data, colour_h.ctypes.data)
        _lib.check(ncol, "cb_colour_cells_host")
        self.ncolours = int(ncol)
        counts = np.bincount(colour_h, minlength=ncol).astype(np.int64)
        self.colour_off = np.zeros(ncol + 1, dtype=np.int64)
        np.cumsum(counts, out=self.colour_off[1:])
        colour = torch.from_numpy(colour_h).to(dev)
        self.colour_perm = torch.sort(colour, stable=True).indices  # cell ids grouped by colour
        self.ccells = self.cells[self.colour_perm].contiguous()
        self.cc2n = cell2nnz.view(nc, k * k)[self.colour_perm].contiguous()
        # gather data (nnz -> contributing (cell, a, b), original cell order) for the matrix kernels
        self.n2c_ptr = self.n2c_idx = None
        if nc * k * k < 2**31:
            valid = torch.nonzero(cell2nnz >= 0).reshape(-1)
            pos = cell2nnz[valid].long()
            order = torch.sort(pos, stable=True).indices
            self.n2c_idx = valid[order].to(torch.int32).contiguous()
            self.n2c_ptr = torch.zeros(self.nnz + 1, dtype=torch.int64, device=dev)
            self.n2c_ptr[1:] = torch.cumsum(torch.bincount(pos, minlength=self.nnz), 0)
            del valid, pos, order
        del cell2nnz
        torch.cuda.current_stream().synchronize()
        self._pattern_ready = True

    def view(self, gather: bool | str = True) -> _lib.cb_mesh_view:
        """The struct the assembly entry points take (keeps no ownership).  ``gather=True``: matrices row by row
        (the default), ``"entries"``: the entry-centric gather kernels over the nnz -> cell lists, ``False`` leaves
        out all gather data, which selects the coloured cell-scatter kernels."""
        self.build_pattern()
        v = _lib.cb_mesh_view()
        v.dim = self.dim
        v.ncolours = self.ncolours
        v.colour_off = self.colour_off.ctypes.data
        v.cells = self.ccells.data_ptr()
        v.cell2nnz = self.cc2n.data_ptr()
        v.coords = self.coords.data_ptr()
        v.nv = self.num_vertices
        v.nrows_owned = self.n_owned_vertices
        v.nnz = self.nnz
        if self.n2c_ptr is not None and gather:
            v.cells_orig = self.cells.data_ptr()
            v.n2c_ptr = self.n2c_ptr.data_ptr()
            v.n2c_idx = self.n2c_idx.data_ptr()
            if self.cell_scratch is None:
                k = self.dim + 1
                self.cell_scratch = torch.empty(self.num_cells * k * self.dim, dtype=torch.float64, device=self.device)
            v.v2c_ptr = self.v2c_ptr.data_ptr()
            v.v2c_idx = self.v2c_idx.data_ptr()
            v.cell_scratch = self.cell_scratch.data_ptr()
            if self.cell_geo is None:
                self.cell_geo = torch.empty(self.num_cells * 16, dtype=torch.float64, device=self.device)
            v.cell_geo = self.cell_geo.data_ptr()
            if gather is True:
                v.rowptr = self.rowptr.data_ptr()
                v.col = self.col.data_ptr()
        return v


def _box_counts(n: int, lengths: list[float]) -> list[int]:
    lmin = min(lengths)
    return [int(np.round(length / lmin * n)) for length in lengths]


def regular_box_mesh(
    n: int = 10,
    start_x: float = 0.0,
    start_y: float = 0.0,
    start_z: float | None = None,
    end_x: float = 1.0,
    end_y: float = 1.0,
    end_z: float | None = None,
    device: str | torch.device = "cuda",
) -> Mesh:
    """``cashocs.regular_box_mesh`` (``cashocs/geometry/mesh.py:221-391``), diagonal="right".

    Vertices are numbered lexicographically (x fastest); every square is split along its
    (v0, v3) diagonal, every cube into the six Kuhn tetrahedra around its (v0, v7) diagonal
    [dolfin RectangleMesh / BoxMesh, recalled]; boundary markers 1..6 = x0, x1, y0, y1, z0, z1.
    """
    n = int(n)
    if n <= 0:
        raise _exceptions.InputError("cashocs_b200.regular_box_mesh", "n", "This needs to be positive.")
    if (start_z is None) != (end_z is None):
        raise _exceptions.InputError(
            "cashocs_b200.regular_box_mesh", "start_z",
            "Incorrect input for the z-coordinate. Both start_z and end_z need to be specified.",
        )
    lo = [start_x, start_y] + ([] if start_z is None else [start_z])
    hi = [end_x, end_y] + ([] if end_z is None else [end_z])
    lengths = [h - l for l, h in zip(lo, hi)]
    if any(length <= 0.0 for length in lengths):
        raise _exceptions.InputError("cashocs_b200.regular_box_mesh", "end", "end must be larger than start")
    npts = _box_counts(n, lengths)
    dev = torch.device(device)
    i64 = dict(dtype=torch.int64, device=dev)
    axes = [
        lo[a] + torch.arange(npts[a] + 1, dtype=torch.float64, device=dev) * (lengths[a] / npts[a])
        for a in range(len(npts))
    ]
    for a in range(len(npts)):
        axes[a][-1] = hi[a]
    bverts: dict[int, torch.Tensor] = {}
    if len(npts) == 2:
        nx, ny = npts
        sx, sy = 1, nx + 1
        yy, xx = torch.meshgrid(axes[1], axes[0], indexing="ij")
        coords = torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=1)
        iy, ix = torch.meshgrid(torch.arange(ny, **i64), torch.arange(nx, **i64), indexing="ij")
        v0 = (iy * sy + ix).reshape(-1)
        tri = torch.stack(
            [torch.stack([v0, v0 + 1, v0 + sy + 1], dim=1), torch.stack([v0, v0 + sy, v0 + sy + 1], dim=1)],
            dim=1,
        ).reshape(-1, 3)
        cells = tri
        vid = torch.arange((nx + 1) * (ny + 1), **i64).view(ny + 1, nx + 1)
        bverts = {1: vid[:, 0], 2: vid[:, nx], 3: vid[0, :], 4: vid[ny, :]}
    else:
        nx, ny, nz = npts
        sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
        zz, yy, xx = torch.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=1)
        iz, iy, ix = torch.meshgrid(
            torch.arange(nz, **i64), torch.arange(ny, **i64), torch.arange(nx, **i64), indexing="ij"
        )
        base = (iz * sz + iy * sy + ix).reshape(-1)
        off = torch.tensor(
            [(b & 1) * sx + ((b >> 1) & 1) * sy + ((b >> 2) & 1) * sz for b in range(8)], **i64
        )
        # Kuhn tetrahedra with ascending corner ids (corner id increases with its bit pattern)
        kuhn = torch.tensor(
            [[0, 1, 3, 7], [0, 1, 5, 7], [0, 4, 5, 7], [0, 2, 3, 7], [0, 4, 6, 7], [0, 2, 6, 7]], **i64
        )
        cells = (base[:, None, None] + off[kuhn][None, :, :]).reshape(-1, 4)
        vid = torch.arange((nx + 1) * (ny + 1) * (nz + 1), **i64).view(nz + 1, ny + 1, nx + 1)
        bverts = {
            1: vid[:, :, 0], 2: vid[:, :, nx], 3: vid[:, 0, :], 4: vid[:, ny, :], 5: vid[0, :, :], 6: vid[nz, :, :],
        }
    bverts = {k: v.reshape(-1).contiguous() for k, v in bverts.items()}
    return Mesh(coords.contiguous(), cells.to(torch.int32).contiguous(), bverts)


def regular_mesh(
    n: int = 10,
    length_x: float = 1.0,
    length_y: float = 1.0,
    length_z: float | None = None,
    device: str | torch.device = "cuda",
) -> Mesh:
    """``cashocs.regular_mesh`` (``cashocs/geometry/mesh.py:155-217``)."""
    if length_x <= 0.0 or length_y <= 0.0 or (length_z is not None and length_z <= 0.0):
        raise _exceptions.InputError("cashocs_b200.regular_mesh", "length", "lengths need to be positive")
    if length_z is None:
        return regular_box_mesh(n, 0.0, 0.0, None, length_x, length_y, None, device=device)
    return regular_box_mesh(n, 0.0, 0.0, 0.0, length_x, length_y, length_z, device=device)


def mesh_from_arrays(coords, cells, facets, facet_tags, device="cuda") -> Mesh:
    """Mesh from host arrays (coords [Nv,d], cells [Nc,d+1], tagged boundary facets [Nf,d])."""
    dev = torch.device(device)
    cells = np.sort(np.asarray(cells), axis=1)
    facets = np.asarray(facets)
    facet_tags = np.asarray(facet_tags)
    bverts = {}
    for tag in np.unique(facet_tags):
        bverts[int(tag)] = torch.from_numpy(
            np.unique(facets[facet_tags == tag].reshape(-1)).astype(np.int64)
        ).to(dev)
    m = Mesh(
        torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).to(dev),
        torch.from_numpy(np.ascontiguousarray(cells, dtype=np.int32)).to(dev),
        bverts,
    )
    m.boundary_facets = {int(tag): torch.from_numpy(np.ascontiguousarray(facets[facet_tags == tag], dtype=np.int64)).to(dev)
                         for tag in np.unique(facet_tags)}
    return m


def import_mesh(path: str, device="cuda") -> Mesh:
    """Load a mesh from a gmsh 4.1 ASCII ``.msh`` file (``cashocs_b200/io/mesh.py``) or from ``.npz``
    (coords, cells, facets, facet_tags).

    The reference's ``cashocs.import_mesh`` reads XDMF/HDF5 (``cashocs/io/mesh.py:186-293``); that
    format is outside the hot path -- the committed fixtures are the
    reference's gmsh meshes converted by ``tests/golden/make_fixtures.py``.
    """
    if str(path).endswith(".msh"):
        from cashocs_b200.io import mesh as mesh_io

        coords, cells, facets, facet_tags, _, node_tags = mesh_io.read_gmsh(path, return_node_tags=True)
        m = mesh_from_arrays(coords, cells, facets, facet_tags, device=device)
        m.gmsh_node_tags = node_tags  # vertex -> gmsh node, for io.write_out_mesh
        m.gmsh_file = str(path)
        return m
    z = np.load(path)
    return mesh_from_arrays(z["coords"], z["cells"], z["facets"], z["facet_tags"], device=device)
