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


def read_gmsh(path: str, return_node_tags: bool = False):
    """Parse a gmsh 4.1 ASCII mesh.

    Returns ``(coords [Nv, d], cells [Nc, d+1], facets [Nf, d], facet_tags [Nf], cell_tags [Nc])`` with
    0-based, densely renumbered vertices (only vertices used by a cell are kept); ``d`` is the highest
    simplex dimension present; facets are the ``d-1`` simplices carrying a physical tag.  With
    ``return_node_tags`` the gmsh node tag of every kept vertex is appended (what :func:`write_out_mesh`
    needs to put moved coordinates back into the file).
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
    out = (coords, np.sort(lookup[cells], axis=1), np.sort(lookup[facets], axis=1), facet_tags, cell_tags)
    return out + (used,) if return_node_tags else out


def gather_coordinates(mesh) -> np.ndarray | None:
    """Global vertex coordinates on rank 0 as a host array (``cashocs/io/mesh.py:403-431``); ``None`` elsewhere.

    One device->host copy of the owned vertices per rank; on a partitioned mesh the owned blocks are gathered
    over the process group and placed by their global vertex ids.
    """
    owned = mesh.coords[: mesh.n_owned_vertices].detach().cpu().numpy()
    dist = getattr(mesh, "dist", None)
    if dist is None or dist.world == 1:
        return owned.copy()
    import torch.distributed as td

    gids = np.asarray(mesh.global_vertex_ids[: mesh.n_owned_vertices])
    gathered = [None] * dist.world if dist.rank == 0 else None
    td.gather_object((gids, owned), gathered, dst=0)
    if dist.rank != 0:
        return None
    n_global = int(mesh.global_offsets[-1])
    points = np.zeros((n_global, owned.shape[1]))
    for ids, xyz in gathered:
        points[ids] = xyz
    return points


def _nodes_header(path: str) -> int:
    with open(path, encoding="utf-8") as fh:
        for line in fh:
            if line.strip() == "$Nodes":
                return int(fh.readline().split()[1])
    raise _exceptions.InputError("cashocs_b200.io.mesh", "path", f"{path} has no $Nodes section")


def check_mesh_compatibility(mesh, original_msh_file: str) -> None:
    """``cashocs/io/mesh.py:489-523``: the file must hold as many nodes as the mesh has (global) vertices."""
    dist = getattr(mesh, "dist", None)
    n_global = int(mesh.global_offsets[-1]) if dist is not None and dist.world > 1 else mesh.num_vertices
    n_file = _nodes_header(original_msh_file)
    tags = getattr(mesh, "gmsh_node_tags", None)
    if n_global != (n_file if tags is None else len(tags)) or (tags is not None and int(np.max(tags)) > _max_tag(original_msh_file)):
        raise _exceptions.CashocsException(
            "The mesh supplied in the configuration file is not compatible with the mesh used."
        )


def _max_tag(path: str) -> int:
    with open(path, encoding="utf-8") as fh:
        for line in fh:
            if line.strip() == "$Nodes":
                return int(fh.readline().split()[3])
    return 0


def write_out_mesh(mesh, original_msh_file: str, out_msh_file: str) -> None:
    """Write the moved mesh as a gmsh ``.msh`` file (``cashocs/io/mesh.py:526-557``).

    Only the coordinates in the ``$Nodes`` section of ``original_msh_file`` (format 4.1 ASCII) are replaced;
    topology, entities and physical groups are copied verbatim, so the result can be handed to gmsh for
    remeshing or re-imported.  Coordinates are printed with 17 significant digits like the reference
    (``create_point_representation``, ``:366-401``: ``x y 0`` in 2-D), i.e. the file round-trips fp64 exactly.
    Vertex ``v`` of the mesh belongs to gmsh node ``mesh.gmsh_node_tags[v]`` (meshes imported by
    :func:`read_gmsh`) or to node ``v + 1`` (the reference's convention, meshes built from converted arrays).
    """
    import os

    check_mesh_compatibility(mesh, original_msh_file)
    points = gather_coordinates(mesh)
    dist = getattr(mesh, "dist", None)
    if points is not None:
        os.makedirs(os.path.dirname(os.path.abspath(out_msh_file)), exist_ok=True)
        tags = getattr(mesh, "gmsh_node_tags", None)
        if tags is None:
            vertex_of_tag = None
        else:
            vertex_of_tag = np.full(int(np.max(tags)) + 1, -1, dtype=np.int64)
            vertex_of_tag[np.asarray(tags)] = np.arange(len(tags))
        _rewrite_nodes(original_msh_file, out_msh_file, points, mesh.dim, vertex_of_tag)
    if dist is not None and dist.world > 1:
        import torch.distributed as td

        td.barrier()


def _rewrite_nodes(src: str, dst: str, points: np.ndarray, dim: int, vertex_of_tag) -> None:
    with open(src, encoding="utf-8") as fh:
        lines = fh.readlines()
    try:
        start = next(i for i, ln in enumerate(lines) if ln.strip() == "$Nodes")
        stop = next(i for i, ln in enumerate(lines) if ln.strip() == "$EndNodes")
    except StopIteration:
        raise _exceptions.InputError("cashocs_b200.io.mesh.write_out_mesh", "original_msh_file",
                                     "no $Nodes section") from None
    out = lines[: start + 2]  # everything up to and including the section header line
    nblocks = int(lines[start + 1].split()[0])
    pos = start + 2
    for _ in range(nblocks):
        header = lines[pos]
        _, _, parametric, nb = (int(t) for t in header.split())
        if parametric:
            raise _exceptions.InputError("cashocs_b200.io.mesh.write_out_mesh", "original_msh_file",
                                         "parametric node blocks are not supported")
        out.append(header)
        out += lines[pos + 1: pos + 1 + nb]
        for i in range(nb):
            tag = int(lines[pos + 1 + i])
            v = tag - 1 if vertex_of_tag is None else (int(vertex_of_tag[tag]) if tag < len(vertex_of_tag) else -1)
            if v < 0:  # a node no cell uses (e.g. a circle centre): keep it where it was
                out.append(lines[pos + 1 + nb + i])
            elif dim == 2:
                out.append(f"{points[v][0]:.16e} {points[v][1]:.16e} 0\n")
            else:
                out.append(f"{points[v][0]:.16e} {points[v][1]:.16e} {points[v][2]:.16e}\n")
        pos += 1 + 2 * nb
    if pos != stop:
        raise _exceptions.InputError("cashocs_b200.io.mesh.write_out_mesh", "original_msh_file",
                                     "malformed $Nodes section")
    out += lines[stop:]
    with open(dst, "w", encoding="utf-8") as fh:
        fh.writelines(out)
