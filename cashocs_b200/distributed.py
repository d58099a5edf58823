"""Mesh partitioning and halo plumbing for one-process-per-GPU runs.

The reference distributes the mesh with MPI through dolfin/PETSc: owned +
ghost entities per rank, ``Vec.ghostUpdate`` per SpMV, ``MPI_Allreduce`` per Krylov dot.
Here every rank owns a contiguous range of (globally numbered) vertices and all cells touching
them; interface cells are duplicated as ghost cells so that the rows of owned dofs assemble
without communication (dolfin ``ghost_mode="shared_vertex"``).  Local numbering: owned
vertices first (global order), then ghost vertices ordered by owner; owned cells (owner = owner
of the cell's lowest vertex) first, then ghost cells.

Two transports share one halo plan:
* ``native`` -- NCCL inside the C library (``csrc/dist.cu``); the Krylov loop exchanges halos
  and all-reduces without returning to Python.  Used on GPUs.
* ``torch``  -- ``torch.distributed`` point-to-point / all-reduce (gloo on CPU); used by the
  CPU tests that check the partitioning and the halo plan.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as td

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200.geometry import mesh as mesh_module


class DistContext:
    """Communicator + halo plan of one rank."""

    def __init__(self, rank: int, world: int, transport: str | None = None) -> None:
        self.rank, self.world = int(rank), int(world)
        self.transport = transport or ("native" if torch.cuda.is_available() else "torch")
        self.handle = None
        self.neigh: list[int] = []
        self.send_off = np.zeros(1, dtype=np.int64)
        self.recv_off = np.zeros(1, dtype=np.int64)
        self.send_idx: torch.Tensor | None = None
        self.n_owned = self.n_ghost = 0
        self._sendbuf: torch.Tensor | None = None
        self.cell_counts: list[int] | None = None
        self.parent: DistContext | None = None  # plans derived for AMG coarse levels share the communicator
        self.p2p = False

    # ------------------------------------------------------------------ setup
    def set_plan(self, neigh, send_lists, recv_counts, n_owned: int, n_ghost: int, device) -> None:
        """``neigh[j]`` exchanges ``send_lists[j]`` (owned local vertex ids) for a contiguous ghost range."""
        self.neigh = [int(r) for r in neigh]
        self.send_off = np.zeros(len(neigh) + 1, dtype=np.int64)
        self.recv_off = np.zeros(len(neigh) + 1, dtype=np.int64)
        for j in range(len(neigh)):
            self.send_off[j + 1] = self.send_off[j] + int(send_lists[j].numel())
            self.recv_off[j + 1] = self.recv_off[j] + int(recv_counts[j])
        self.send_idx = (torch.cat([s.to(torch.int32) for s in send_lists]) if len(neigh)
                         else torch.zeros(0, dtype=torch.int32)).to(device).contiguous()
        self.n_owned, self.n_ghost = int(n_owned), int(n_ghost)
        if self.recv_off[-1] != n_ghost:
            raise _exceptions.CashocsException("halo plan does not cover all ghost vertices")
        if self.transport == "native":
            self._create_native(device)

    def _create_native(self, device) -> None:
        lib = _lib.load()
        neigh = np.asarray(self.neigh, dtype=np.int32)
        if self.parent is not None:
            handle = C.c_void_p()
            _lib.check(
                lib.cb_dist_clone_plan(self.parent.handle, C.byref(handle), len(self.neigh), neigh.ctypes.data,
                                       self.send_off.ctypes.data, _lib.ptr(self.send_idx), self.recv_off.ctypes.data,
                                       self.n_owned, self.n_ghost),
                "cb_dist_clone_plan",
            )
            self.handle = handle
            self._enable_p2p()
            return
        ident = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            _lib.check(lib.cb_dist_unique_id(buf), "cb_dist_unique_id")
            ident = [bytes(buf.raw)]
        td.broadcast_object_list(ident, src=0)
        neigh = np.asarray(self.neigh, dtype=np.int32)
        handle = C.c_void_p()
        _lib.check(
            lib.cb_dist_create(C.byref(handle), ident[0], self.rank, self.world, len(self.neigh),
                               neigh.ctypes.data, self.send_off.ctypes.data, _lib.ptr(self.send_idx),
                               self.recv_off.ctypes.data, self.n_owned, self.n_ghost),
            "cb_dist_create",
        )
        self.handle = handle
        self._enable_p2p()

    def _enable_p2p(self) -> None:
        """Switch this plan to the one-sided NVLink transport (``csrc/dist.cu``): export this rank's
        window, all-gather the CUDA IPC handles and the ghost offsets, map the peers' windows.
        Collective; any rank failing to map a peer sends every rank back to NCCL.
        ``CASHOCS_B200_P2P=0`` keeps NCCL send/recv + all-reduce."""
        self.p2p = False
        if os.environ.get("CASHOCS_B200_P2P", "1") == "0" or self.world > 32:
            return
        lib = _lib.load()
        buf = C.create_string_buffer(64)
        ok = lib.cb_dist_p2p_export(self.handle, buf) == 0
        info = (bytes(buf.raw), {int(r): int(self.recv_off[j]) for j, r in enumerate(self.neigh)}, ok,
                int(lib.cb_dist_p2p_ghost_cap(self.handle)))
        allinfo = [None] * self.world
        td.all_gather_object(allinfo, info)
        if all(i[2] for i in allinfo):
            handles = b"".join(i[0] for i in allinfo)
            remote_off = np.array([allinfo[r][1][self.rank] for r in self.neigh] + [0], dtype=np.int64)
            remote_cap = np.array([allinfo[r][3] for r in self.neigh] + [0], dtype=np.int64)
            ok = lib.cb_dist_p2p_open(self.handle, handles, remote_off.ctypes.data, remote_cap.ctypes.data) == 0
        else:
            ok = False
        oks = [None] * self.world
        td.all_gather_object(oks, ok)
        if all(oks):
            self.p2p = True
        else:
            lib.cb_dist_p2p_disable(self.handle)

    def check_p2p(self) -> None:
        """Raise if a peer timed out inside a one-sided wait on any plan of this communicator (the vectors are then
        invalid).  The error word is pinned host memory shared by the root plan and its clones: this is a plain load,
        so it runs after every exchange, all-reduce and solve; the wait limit is ``CASHOCS_B200_P2P_TIMEOUT_S`` (60 s)."""
        if self.handle is not None and self.transport == "native" and _lib.load().cb_dist_p2p_error(self.handle):
            raise _exceptions.CashocsException(
                "one-sided NVLink exchange: a peer did not answer within the wait limit (CASHOCS_B200_P2P_TIMEOUT_S); "
                "the communicator is unusable")

    def owner_of_ghost(self) -> torch.Tensor:
        """Owner rank of every ghost vertex (ghosts are ordered by owner)."""
        counts = torch.from_numpy(np.diff(self.recv_off))
        return torch.repeat_interleave(torch.tensor(self.neigh, dtype=torch.int64), counts)

    def derive_plan(self, owners: torch.Tensor, ids: torch.Tensor, n_owned: int, device) -> "DistContext":
        """Halo plan of a derived (coarser) numbering: this rank needs the entities ``ids`` (numbered
        locally by their owner) of ranks ``owners``; both sorted by (owner, id).  Collective."""
        need = {int(r): ids[owners == r].cpu() for r in torch.unique(owners).tolist()}
        all_need = [None] * self.world
        td.all_gather_object(all_need, need)
        neigh = sorted(set(need) | {r for r in range(self.world) if r != self.rank and self.rank in all_need[r]})
        empty = torch.zeros(0, dtype=torch.int64)
        send_lists = [all_need[r].get(self.rank, empty).to(torch.int64) for r in neigh]
        recv_counts = [int(need[r].numel()) if r in need else 0 for r in neigh]
        new = DistContext(self.rank, self.world, self.transport)
        new.parent = self if self.parent is None else self.parent
        new.set_plan(neigh, send_lists, recv_counts, n_owned, int(ids.numel()), device)
        return new

    def destroy(self) -> None:
        if self.handle is not None:
            _lib.load().cb_dist_destroy(self.handle)
            self.handle = None

    # ------------------------------------------------------------------ collectives
    def allreduce(self, t: torch.Tensor, op: str = "sum") -> torch.Tensor:
        """In-place all-reduce of a small fp64 tensor."""
        if self.world == 1:
            return t
        if self.transport == "native":
            code = {"sum": 0, "min": 1, "max": 2}[op]
            _lib.check(_lib.load().cb_dist_allreduce(_lib.stream_ptr(), self.handle, _lib.ptr(t), t.numel(), code),
                       "cb_dist_allreduce")
            self.check_p2p()
        else:
            td.all_reduce(t, {"sum": td.ReduceOp.SUM, "min": td.ReduceOp.MIN, "max": td.ReduceOp.MAX}[op])
        return t

    def allreduce_scalar(self, v: float, op: str = "sum") -> float:
        dev = self.send_idx.device if self.send_idx is not None else "cpu"
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        out = float(self.allreduce(t, op).item())
        self.check_p2p()  # .item() synchronised: a time-out inside this very exchange is visible now
        return out

    def halo_exchange(self, x: torch.Tensor, bs: int = 1) -> torch.Tensor:
        """Ghost update of a vertex-blocked vector ``[(n_owned + n_ghost) * bs]`` (``Vec.ghostUpdate``)."""
        if self.world == 1 or not self.neigh:
            return x
        if self.transport == "native":
            need = int(self.send_off[-1]) * bs
            if self._sendbuf is None or self._sendbuf.numel() < need:
                self._sendbuf = torch.empty(max(need, 1), dtype=torch.float64, device=x.device)
            _lib.check(_lib.load().cb_dist_halo_exchange(_lib.stream_ptr(), self.handle, _lib.ptr(x), bs,
                                                         _lib.ptr(self._sendbuf)), "cb_dist_halo_exchange")
            self.check_p2p()
            return x
        xv = x.view(-1, bs)
        reqs, bufs = [], []
        for j, r in enumerate(self.neigh):
            idx = self.send_idx[self.send_off[j]:self.send_off[j + 1]].long()
            if idx.numel():
                reqs.append(td.isend(xv[idx].contiguous(), r))
            cnt = int(self.recv_off[j + 1] - self.recv_off[j])
            if cnt:
                buf = torch.empty(cnt, bs, dtype=x.dtype, device=x.device)
                bufs.append((j, buf))
                reqs.append(td.irecv(buf, r))
        for q in reqs:
            q.wait()
        for j, buf in bufs:
            xv[self.n_owned + int(self.recv_off[j]): self.n_owned + int(self.recv_off[j + 1])] = buf
        return x

    def gather_cells(self, q: torch.Tensor) -> torch.Tensor:
        """Concatenate per-rank owned-cell arrays on every rank (quantile reductions only)."""
        if self.world == 1:
            return q
        objs = [None] * self.world
        td.all_gather_object(objs, q.cpu())
        return torch.cat([o.to(q.device) for o in objs])


def vertex_offsets(num_vertices: int, world: int, align: int = 1) -> np.ndarray:
    """Balanced contiguous vertex ranges, optionally aligned to multiples of ``align`` (mesh layers)."""
    units = num_vertices // align
    if units * align != num_vertices:
        raise _exceptions.InputError("vertex_offsets", "align", "num_vertices must be a multiple of align")
    off = np.array([(units * r) // world for r in range(world + 1)], dtype=np.int64) * align
    return off


def partition_mesh(coords: torch.Tensor, cells: torch.Tensor, boundary_vertices: dict, offsets: np.ndarray,
                   dist: DistContext) -> mesh_module.Mesh:
    """Extract rank ``dist.rank``'s partition of a globally numbered mesh and build its halo plan.

    ``coords`` / ``cells`` are the *global* arrays (every rank generates or loads them once; only
    the local part is kept).  Cells must be ascending per cell in the global numbering, so that
    column 0 is the cell's lowest vertex (cell ownership).
    """
    rank, world = dist.rank, dist.world
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    dev = coords.device
    cells = cells.long()
    inside = (cells >= lo) & (cells < hi)
    local_mask = inside.any(dim=1)
    owned_cell = inside[:, 0][local_mask]
    lcells = cells[local_mask]
    order = torch.sort((~owned_cell).to(torch.int8), stable=True).indices  # owned cells first
    lcells = lcells[order]
    n_owned_cells = int(owned_cell.sum().item())
    flat = lcells.reshape(-1)
    ghost_glob = torch.unique(flat[(flat < lo) | (flat >= hi)])  # ascending == ordered by owner
    n_owned, n_ghost = hi - lo, int(ghost_glob.numel())
    # global -> local
    is_owned = (flat >= lo) & (flat < hi)
    loc = torch.where(is_owned, flat - lo, n_owned + torch.searchsorted(ghost_glob, flat))
    lcells = loc.view(-1, cells.shape[1]).to(torch.int32).contiguous()
    vert_glob = torch.cat([torch.arange(lo, hi, device=dev), ghost_glob])
    lcoords = coords[vert_glob].contiguous()
    # boundary vertices in local numbering (ghosts included: BC flags are needed on ghost dofs too)
    g2l = torch.full((int(coords.shape[0]),), -1, dtype=torch.int64, device=dev)
    g2l[vert_glob] = torch.arange(vert_glob.numel(), device=dev)
    bverts = {}
    for k, v in boundary_vertices.items():
        l = g2l[v.to(dev).long()]
        bverts[int(k)] = torch.sort(l[l >= 0]).values
    # halo plan: who owns my ghosts, and who needs my vertices
    off_t = torch.from_numpy(offsets).to(dev)
    owner = torch.searchsorted(off_t, ghost_glob, right=True) - 1
    need = {int(r): ghost_glob[owner == r].cpu() for r in torch.unique(owner).tolist()}
    all_need = [None] * world
    td.all_gather_object(all_need, need)
    neigh = sorted(set(need) | {s for s in range(world) if s != rank and rank in all_need[s]})
    send_lists, recv_counts = [], []
    for r in neigh:
        wanted = all_need[r].get(rank, torch.zeros(0, dtype=torch.int64))
        send_lists.append((wanted - lo).to(torch.int64))
        recv_counts.append(int(need[r].numel()) if r in need else 0)
    dist.set_plan(neigh, send_lists, recv_counts, n_owned, n_ghost, dev)
    m = mesh_module.Mesh(lcoords, lcells, bverts, n_owned_vertices=n_owned, n_owned_cells=n_owned_cells, dist=dist)
    m.global_vertex_ids = vert_glob
    m.global_offsets = offsets
    return m


def rcb_order(coords: np.ndarray, parts: int) -> tuple[np.ndarray, np.ndarray]:
    """Recursive coordinate bisection of a point set into ``parts`` pieces of (nearly) equal size: returns the vertex
    permutation ``order`` (old ids, grouped by part) and the part offsets.  The geometric stand-in for the graph
    partitioner dolfin calls (SCOTCH / ParMETIS [third party]): compact parts with small interfaces for any mesh, no
    structure assumed; deterministic (stable sorts), identical on every rank."""
    coords = np.asarray(coords, dtype=np.float64)
    n = coords.shape[0]
    order = np.arange(n, dtype=np.int64)
    offsets = [0]

    def split(idx: np.ndarray, k: int) -> None:
        if k == 1:
            order[offsets[-1]: offsets[-1] + idx.size] = idx
            offsets.append(offsets[-1] + idx.size)
            return
        pts = coords[idx]
        axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))
        k_lo = k // 2
        cut = (idx.size * k_lo) // k
        srt = idx[np.argsort(pts[:, axis], kind="stable")]
        split(srt[:cut], k_lo)
        split(srt[cut:], k - k_lo)

    split(np.arange(n, dtype=np.int64), int(parts))
    return order, np.asarray(offsets, dtype=np.int64)


def distribute_mesh(coords, cells, facets, facet_tags, dist: DistContext, device=None) -> mesh_module.Mesh:
    """Partition ANY simplicial mesh given by global host arrays (e.g. from ``io.mesh.read_gmsh``) over the ranks of
    ``dist``: vertices are renumbered part by part along a recursive coordinate bisection, then every rank keeps its
    contiguous vertex range plus the ghost layer (``partition_mesh``) -- what dolfin does at mesh import with a graph
    partitioner (owned-then-ghost ordering, ``cashocs/geometry/quality.py:319-320``).  Collective."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    order, offsets = rcb_order(coords, dist.world)
    new_id = np.empty_like(order)
    new_id[order] = np.arange(order.size)
    ncells = np.sort(new_id[cells], axis=1)                       # ascending per cell: column 0 owns the cell
    ncells = ncells[np.lexsort(ncells.T[::-1])]                    # deterministic cell order
    nfacets = new_id[np.asarray(facets, dtype=np.int64)] if len(facets) else np.zeros((0, coords.shape[1]), dtype=np.int64)
    tags = np.asarray(facet_tags)
    bverts = {int(t): torch.from_numpy(np.unique(nfacets[tags == t].reshape(-1))) for t in np.unique(tags)}
    m = partition_mesh(torch.from_numpy(coords[order]).to(device), torch.from_numpy(ncells).to(device), bverts, offsets, dist)
    m.global_boundary_facets = {int(t): nfacets[tags == t] for t in np.unique(tags)}
    m.rcb_order = order                                             # new global id -> id in the input arrays
    return m


def distributed_regular_mesh(n: int, dist: DistContext, lengths=(1.0, 1.0, 1.0), device=None) -> mesh_module.Mesh:
    """Unit-cube mesh partitioned into z-slabs of whole vertex layers (one halo layer per side)."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    g = mesh_module.regular_mesh(n, *lengths, device=device)
    npts = mesh_module._box_counts(n, list(lengths))
    layer = (npts[0] + 1) * (npts[1] + 1)
    offsets = vertex_offsets(g.num_vertices, dist.world, align=layer)
    bverts = {k: g.boundary_vertices([k]) for k in g.boundary_markers}
    m = partition_mesh(g.coords, g.cells, bverts, offsets, dist)
    del g
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return m
