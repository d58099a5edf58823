"""Mesh import for the device layer (host side).

The reference reads XDMF/HDF5 produced by ``cashocs-convert`` from gmsh files
(``cashocs/io/mesh.py:186-293``, ``cashocs/_cli``); h5py is not available here, so the device layer reads
the *source* of that conversion directly: gmsh ``.msh`` files in format 4.1 ASCII (what ``cashocs-convert``
consumes), plus the small ``.npz`` container used for the committed fixtures.  Physical groups of dimension
``d-1`` become the boundary markers (``boundaries`` MeshFunction of the reference), physical groups of
dimension ``d`` the subdomain markers.  Only simplices (lines, triangles, tetrahedra) are accepted, like the
reference.
"""

from __future__ import annotations

import numpy as np

from cashocs_b200 import _exceptions

# gmsh element type -> (topological dimension, nodes per element) for first-order simplices
_SIMPLEX = {15: (0, 1), 1: (1, 2), 2: (2, 3), 4: (3, 4)}


def _blocks(path: str) -> dict[str, list[str]]:
    out: dict[str, list[str]] = {}
    name = None
    with open(path) as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line.startswith("$End"):
                name = None
            elif line.startswith("$"):
                name = line[1:]
                out[name] = []
            elif name is not None:
                out[name].append(line)
    return out


def read_gmsh(path: str):
    """Parse a gmsh 4.1 ASCII mesh.

    Returns ``(coords [Nv, d], cells [Nc, d+1], facets [Nf, d], facet_tags [Nf], cell_tags [Nc])`` with
    0-based, densely renumbered vertices (only vertices used by a cell are kept); ``d`` is the highest
    simplex dimension present; facets are the ``d-1`` simplices carrying a physical tag.
    """
    blk = _blocks(path)
    fmt = blk.get("MeshFormat", ["0"])[0].split()
    if not fmt or not fmt[0].startswith("4.1") or (len(fmt) > 1 and fmt[1] != "0"):
        raise _exceptions.InputError("cashocs_b200.io.mesh.read_gmsh", "path", "only gmsh format 4.1 ASCII is supported")
    # entity -> first physical tag, per dimension
    phys: dict[tuple[int, int], int] = {}
    ent = blk.get("Entities", [])
    if ent:
        counts = [int(t) for t in ent[0].split()]
        pos = 1
        for dim, n in enumerate(counts):
            for _ in range(n):
                t = ent[pos].split()
                pos += 1
                tag = int(t[0])
                k = 4 if dim == 0 else 7  # points: tag x y z ; others: tag + bounding box
                nphys = int(t[k])
                if nphys > 0:
                    phys[(dim, tag)] = abs(int(t[k + 1]))
    # nodes
    nodes = blk["Nodes"]
    nblocks, nnodes = (int(t) for t in nodes[0].split()[:2])
    tags = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3))
    pos, filled = 1, 0
    for _ in range(nblocks):
        _, _, parametric, nb = (int(t) for t in nodes[pos].split())
        pos += 1
        tags[filled: filled + nb] = [int(nodes[pos + i]) for i in range(nb)]
        pos += nb
        for i in range(nb):
            xyz[filled + i] = [float(t) for t in nodes[pos + i].split()[:3]]
        pos += nb
        filled += nb
    # elements
    elems = blk["Elements"]
    nblocks = int(elems[0].split()[0])
    by_dim: dict[int, list] = {0: [], 1: [], 2: [], 3: []}
    tag_by_dim: dict[int, list] = {0: [], 1: [], 2: [], 3: []}
    pos = 1
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(t) for t in elems[pos].split())
        pos += 1
        if etype not in _SIMPLEX:
            raise _exceptions.InputError("cashocs_b200.io.mesh.read_gmsh", "path",
                                         f"element type {etype} is not a first-order simplex")
        nn = _SIMPLEX[etype][1]
        for i in range(nb):
            by_dim[edim].append([int(t) for t in elems[pos + i].split()[1: 1 + nn]])
        tag_by_dim[edim] += [phys.get((edim, etag), 0)] * nb
        pos += nb
    d = max(k for k, v in by_dim.items() if v)
    if d < 2:
        raise _exceptions.InputError("cashocs_b200.io.mesh.read_gmsh", "path", "no triangles or tetrahedra in the mesh")
    cells = np.asarray(by_dim[d], dtype=np.int64)
    cell_tags = np.asarray(tag_by_dim[d], dtype=np.int64)
    facets = np.asarray(by_dim[d - 1], dtype=np.int64).reshape(-1, d)
    facet_tags = np.asarray(tag_by_dim[d - 1], dtype=np.int64)
    keep = facet_tags > 0
    facets, facet_tags = facets[keep], facet_tags[keep]
    # dense 0-based renumbering of the vertices that cells use, in ascending gmsh tag order
    used = np.unique(cells)
    lookup = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    lookup[used] = np.arange(used.size)
    coord_of_tag = np.full((int(tags.max()) + 1, 3), np.nan)
    coord_of_tag[tags] = xyz
    coords = coord_of_tag[used][:, :d]
    if np.isnan(coords).any() or (facets.size and (lookup[facets] < 0).any()):
        raise _exceptions.InputError("cashocs_b200.io.mesh.read_gmsh", "path", "inconsistent node references")
    return coords, np.sort(lookup[cells], axis=1), np.sort(lookup[facets], axis=1), facet_tags, cell_tags
