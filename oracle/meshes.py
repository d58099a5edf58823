"""Oracle meshes: synthetic unit-square / unit-cube meshes and a gmsh 4.1 ASCII reader.

Test infrastructure only (see ``oracle/__init__.py``).

Follows ``cashocs/geometry/mesh.py:221-391`` (``regular_box_mesh``: ``n`` cells along the
shortest side, ``round(L/Lmin*n)`` along the others, boundary markers 1..6 = x0, x1, y0, y1,
z0, z1) and the recalled dolfin ``RectangleMesh(diagonal="right")`` / ``BoxMesh``
connectivity [third party, recalled -- SURVEY.md 9.1]: vertices lexicographic with x
fastest; per square ``(v0,v1,v3),(v0,v2,v3)``; per cube the six Kuhn tetrahedra around the
diagonal v0-v7; cell vertex lists sorted ascending (dolfin ``mesh.order()``).
"""

from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class Mesh:
    """A simplicial mesh with tagged boundary facets."""

    coords: np.ndarray  # [Nv, d] float64
    cells: np.ndarray  # [Nc, d+1] int32, ascending per cell
    facets: np.ndarray  # [Nf, d] int32 boundary facets
    facet_tags: np.ndarray  # [Nf] int32 boundary markers

    @property
    def dim(self) -> int:
        return self.coords.shape[1]

    @property
    def num_vertices(self) -> int:
        return self.coords.shape[0]

    @property
    def num_cells(self) -> int:
        return self.cells.shape[0]

    def boundary_vertices(self, markers) -> np.ndarray:
        """Sorted unique vertices on facets carrying one of ``markers``.

        dolfin's topological DirichletBC search: all dofs on the marked facets
        (``cashocs/_utils/forms.py:214-271``, ``create_dirichlet_bcs``).
        """
        markers = np.atleast_1d(np.asarray(markers, dtype=np.int64))
        if markers.size == 0 or self.facets.size == 0:
            return np.zeros(0, dtype=np.int64)
        mask = np.isin(self.facet_tags, markers)
        return np.unique(self.facets[mask].reshape(-1)).astype(np.int64)


_KUHN = np.array(
    [
        [0, 1, 3, 7],
        [0, 1, 7, 5],
        [0, 5, 7, 4],
        [0, 3, 2, 7],
        [0, 6, 4, 7],
        [0, 2, 6, 7],
    ],
    dtype=np.int64,
)


def regular_mesh(n: int, lx: float = 1.0, ly: float = 1.0, lz: float | None = None) -> Mesh:
    """``cashocs.regular_mesh(n, lx, ly, lz)`` (``cashocs/geometry/mesh.py:155-217``)."""
    sizes = [lx, ly] if lz is None else [lx, ly, lz]
    smin = min(sizes)
    npts = [int(np.round(s / smin * n)) for s in sizes]
    if lz is None:
        nx, ny = npts
        xs = np.linspace(0.0, lx, nx + 1)
        ys = np.linspace(0.0, ly, ny + 1)
        X, Y = np.meshgrid(xs, ys, indexing="xy")  # y slow, x fast
        coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
        ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
        v0 = (iy * (nx + 1) + ix).reshape(-1)
        v1 = v0 + 1
        v2 = v0 + nx + 1
        v3 = v1 + nx + 1
        cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
        cells[0::2] = np.stack([v0, v1, v3], axis=1)
        cells[1::2] = np.stack([v0, v2, v3], axis=1)
    else:
        nx, ny, nz = npts
        xs = np.linspace(0.0, lx, nx + 1)
        ys = np.linspace(0.0, ly, ny + 1)
        zs = np.linspace(0.0, lz, nz + 1)
        Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
        coords = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
        iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        base = (iz * (ny + 1) * (nx + 1) + iy * (nx + 1) + ix).reshape(-1)
        sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
        corner = np.stack(
            [base + (b & 1) * sx + ((b >> 1) & 1) * sy + ((b >> 2) & 1) * sz for b in range(8)],
            axis=1,
        )  # [ncube, 8], bit order x,y,z
        cells = corner[:, _KUHN].reshape(-1, 4)
    cells = np.sort(cells, axis=1).astype(np.int32)
    facets, tags = _box_boundary_facets(coords, cells, sizes)
    return Mesh(np.ascontiguousarray(coords), np.ascontiguousarray(cells), facets, tags)


def boundary_facets(cells: np.ndarray) -> np.ndarray:
    """Facets that belong to exactly one cell (each returned sorted ascending)."""
    d1 = cells.shape[1]
    allf = []
    for k in range(d1):
        keep = [j for j in range(d1) if j != k]
        allf.append(cells[:, keep])
    allf = np.sort(np.concatenate(allf, axis=0), axis=1)
    uniq, counts = np.unique(allf, axis=0, return_counts=True)
    return uniq[counts == 1].astype(np.int32)


def _box_boundary_facets(coords, cells, sizes):
    facets = boundary_facets(cells)
    tags = np.zeros(facets.shape[0], dtype=np.int32)
    eps = 3.0e-16  # DOLFIN_EPS (cashocs/geometry/mesh.py:339-365)
    for axis, length in enumerate(sizes):
        c = coords[facets, axis]  # [Nf, d]
        lo = np.all(np.abs(c - 0.0) < eps, axis=1)
        hi = np.all(np.abs(c - length) < eps, axis=1)
        tags[lo] = 2 * axis + 1
        tags[hi] = 2 * axis + 2
    return facets, tags


def read_gmsh41(path: str) -> Mesh:
    """Read a gmsh 4.1 ASCII ``.msh`` file (triangles / tetrahedra + tagged boundary facets).

    Stands in for ``cashocs.import_mesh`` (``cashocs/io/mesh.py:186-293``) on the
    reference's own fixtures; the ``.h5`` twins are unreadable here (no h5py).  Boundary
    tags are the gmsh *physical* tags of the codim-1 entities, which is what
    ``cashocs-convert`` stores in the boundaries MeshFunction.
    """
    with open(path, "r", encoding="utf-8") as fh:
        lines = fh.read().split("\n")
    idx = {name: i for i, name in enumerate(lines) if name.startswith("$")}

    # entities -> physical tag per (dim, entity tag)
    phys: dict[tuple[int, int], int] = {}
    i = idx["$Entities"] + 1
    npnt, ncur, nsur, nvol = (int(t) for t in lines[i].split())
    i += 1
    for _ in range(npnt):
        i += 1
    for dim, cnt in ((1, ncur), (2, nsur), (3, nvol)):
        for _ in range(cnt):
            tok = lines[i].split()
            tag = int(tok[0])
            nphys = int(tok[7])
            if nphys > 0:
                phys[(dim, tag)] = int(tok[8])
            i += 1

    # nodes
    i = idx["$Nodes"] + 1
    nblocks, nnodes, _, _ = (int(t) for t in lines[i].split())
    i += 1
    node_tags = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3))
    pos = 0
    for _ in range(nblocks):
        _, _, _, nb = (int(t) for t in lines[i].split())
        i += 1
        for k in range(nb):
            node_tags[pos + k] = int(lines[i + k])
        for k in range(nb):
            xyz[pos + k] = [float(t) for t in lines[i + nb + k].split()]
        i += 2 * nb
        pos += nb
    remap = -np.ones(node_tags.max() + 1, dtype=np.int64)
    remap[node_tags] = np.arange(nnodes)

    # elements
    i = idx["$Elements"] + 1
    nblocks, _, _, _ = (int(t) for t in lines[i].split())
    i += 1
    blocks = []
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(t) for t in lines[i].split())
        i += 1
        conn = np.array([[int(t) for t in lines[i + k].split()[1:]] for k in range(nb)], dtype=np.int64)
        i += nb
        blocks.append((edim, etag, etype, conn))
    tdim = max(b[0] for b in blocks)
    cells = np.concatenate([remap[b[3]] for b in blocks if b[0] == tdim], axis=0)
    fac, ftag = [], []
    for edim, etag, _, conn in blocks:
        if edim == tdim - 1 and (edim, etag) in phys:
            fac.append(remap[conn])
            ftag.append(np.full(conn.shape[0], phys[(edim, etag)], dtype=np.int32))
    coords = xyz[:, :tdim].copy()
    used = np.zeros(nnodes, dtype=bool)
    used[cells.reshape(-1)] = True
    if not used.all():  # drop orphan nodes (e.g. the circle centre point)
        new = -np.ones(nnodes, dtype=np.int64)
        new[used] = np.arange(int(used.sum()))
        coords = coords[used]
        cells = new[cells]
        fac = [new[f] for f in fac]
    cells = np.sort(cells, axis=1).astype(np.int32)
    if fac:
        facets = np.sort(np.concatenate(fac, axis=0), axis=1).astype(np.int32)
        tags = np.concatenate(ftag)
    else:
        facets = np.zeros((0, tdim), dtype=np.int32)
        tags = np.zeros(0, dtype=np.int32)
    return Mesh(np.ascontiguousarray(coords), np.ascontiguousarray(cells), facets, tags)


def save_npz(mesh: Mesh, path: str) -> None:
    np.savez_compressed(
        path, coords=mesh.coords, cells=mesh.cells, facets=mesh.facets, facet_tags=mesh.facet_tags
    )


def load_npz(path: str) -> Mesh:
    z = np.load(path)
    return Mesh(z["coords"], z["cells"], z["facets"], z["facet_tags"])
