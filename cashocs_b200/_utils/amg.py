"""Smoothed-aggregation AMG hierarchy: symbolic phase (host layer, torch index plumbing) and the
driver of the numeric phase (C-ABI kernels in ``csrc/amg.cu``).

This is what the reference's iterative default ``pc_type hypre`` / ``boomeramg``
(``cashocs/_utils/linalg.py:44-52``; the Lame-mu solve ``cashocs/_forms/shape_form_handler.py:113-120``)
and PETSc's ``pc_type gamg`` map to on the device.  Unknown-based smoothed aggregation with one
scalar prolongator shared by all components, so every level is a block-CSR matrix of the fine
block size and runs through the same SpMV kernel.

* Symbolic (once per sparsity pattern, i.e. once per mesh topology): MIS(2) aggregation on the
  matrix graph with hashed priorities (integer arithmetic only -> deterministic), patterns of
  ``P``, ``P^T``, ``A P`` and ``P^T A P`` with their term lists.
* Numeric (every time the matrix values change): strength matrix ``S = tr(A_ij)/bs``, smoothed
  prolongator ``P = (I - 4/(3 lmax) D_S^-1 S) T``, Galerkin operators, inverse diagonals and
  Gershgorin bounds for the Chebyshev smoothers.
"""

from __future__ import annotations

import ctypes as C
import os

import torch

from cashocs_b200 import _lib


def _rows_of(rowptr: torch.Tensor) -> torch.Tensor:
    n = rowptr.numel() - 1
    return torch.repeat_interleave(torch.arange(n, device=rowptr.device), rowptr[1:] - rowptr[:-1])


def _rowptr_from_rows(rows: torch.Tensor, n: int) -> torch.Tensor:
    rp = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return rp


def mis_aggregate(rows: torch.Tensor, cols: torch.Tensor, n: int, distance: int = 2):
    """MIS(k) aggregation on the matrix graph (k = ``distance``).

    Roots form a maximal independent set of the k-th power of the graph (hashed integer priorities
    -> deterministic).  k = 2 is the classical smoothed-aggregation choice: roots are >= 3 edges
    apart, so the root neighbourhoods are disjoint aggregates of diameter 2; nodes at distance 2
    from every root join the aggregate of their highest-priority already-aggregated neighbour.
    k = 1 gives small aggregates (root + part of its neighbourhood).
    Returns (agg [n] int32, number of aggregates)."""
    dev = rows.device
    off = rows != cols
    r, c = rows[off], cols[off]
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    h = (idx * 2654435761 + 12345) % 2147483629
    h = (h * 40503 + 7) % 2147483629
    prio = h * n + idx  # unique, < 2^31 * n
    state = torch.zeros(n, dtype=torch.int8, device=dev)  # 0 undecided, 1 root, 2 covered
    neg = torch.full((n,), -1, dtype=torch.int64, device=dev)

    def nbmax(v):  # max over the closed neighbourhood
        return v.clone().scatter_reduce_(0, r, v[c], "amax", include_self=True)

    for _ in range(400):
        und = state == 0
        if not bool(und.any()):
            break
        p = torch.where(und, prio, neg)
        m = p
        for _k in range(distance):
            m = nbmax(m)
        newroot = und & (p == m)
        state[newroot] = 1
        cov = newroot.to(torch.int64)
        for _k in range(distance):
            cov = nbmax(cov)
        state[(state == 0) & (cov > 0)] = 2
    is_root = state == 1
    nc = int(is_root.sum().item())
    agg = torch.full((n,), -1, dtype=torch.int64, device=dev)
    agg[is_root] = torch.arange(nc, device=dev)
    for _k in range(distance):
        # unassigned nodes join the aggregate of their highest-priority assigned neighbour
        assigned = agg >= 0
        best = neg.clone().scatter_reduce_(0, r, torch.where(assigned[c], prio[c], neg[c]), "amax", include_self=True)
        join = (~assigned) & (best >= 0)
        agg[join] = agg[(best % n)[join]]
    if bool((agg < 0).any()):
        raise RuntimeError("aggregation left unassigned nodes")
    return agg.to(torch.int32), nc


def _expand_terms(lhs_rows, lhs_cols, lhs_index, rhs_rowptr, rhs_col, nrows_out, ncols_out, chunk_rows=300_000):
    """Symbolic sparse product L * R.

    Returns the CSR pattern of the product (rowptr, col) and, for every output entry, the list of
    contributing (lhs entry, rhs entry) pairs in a fixed order (seg_ptr, src_l, src_r) -- the
    numeric phase is then a pure gather (``cb_amg_spgemm_terms``).  ``lhs_rows`` must be ascending.
    """
    dev = lhs_rows.device
    lens_all = rhs_rowptr[1:] - rhs_rowptr[:-1]
    keys, counts, sl, sr = [], [], [], []
    bounds = torch.searchsorted(lhs_rows, torch.arange(0, nrows_out + chunk_rows, chunk_rows, device=dev)).tolist()
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b <= a:
            continue
        i, k = lhs_rows[a:b], lhs_cols[a:b]
        lens = lens_all[k]
        total = int(lens.sum().item())
        if total == 0:
            continue
        rep = torch.repeat_interleave(torch.arange(b - a, device=dev), lens)
        start = torch.cumsum(lens, 0) - lens
        rpos = rhs_rowptr[k][rep] + (torch.arange(total, device=dev) - start[rep])
        key = i[rep] * ncols_out + rhs_col[rpos].long()
        uniq, inv = torch.unique(key, return_inverse=True)
        order = torch.argsort(inv, stable=True)  # expansion order inside a segment -> fixed summation order
        keys.append(uniq)
        counts.append(torch.bincount(inv, minlength=uniq.numel()))
        sl.append(lhs_index[a:b][rep][order].to(torch.int32))
        sr.append(rpos[order].to(torch.int32))
        del rep, start, rpos, key, inv, order
    key = torch.cat(keys)
    seg_ptr = torch.zeros(key.numel() + 1, dtype=torch.int64, device=dev)
    seg_ptr[1:] = torch.cumsum(torch.cat(counts), 0)
    out_rows = key // ncols_out
    return (_rowptr_from_rows(out_rows, nrows_out), (key % ncols_out).to(torch.int32).contiguous(), seg_ptr,
            torch.cat(sl).contiguous(), torch.cat(sr).contiguous())


def _expand_pattern(lhs_rows, lhs_cols, rhs_rowptr, rhs_col, nrows_out, ncols_out, chunk_rows=300_000):
    """CSR pattern (rowptr, col) of L * R without the term lists (see ``_expand_terms``); ``lhs_rows``
    ascending.  Peak memory is one chunk of keys."""
    dev = lhs_rows.device
    lens_all = rhs_rowptr[1:] - rhs_rowptr[:-1]
    keys = []
    bounds = torch.searchsorted(lhs_rows, torch.arange(0, nrows_out + chunk_rows, chunk_rows, device=dev)).tolist()
    for a, b in zip(bounds[:-1], bounds[1:]):
        if b <= a:
            continue
        i, k = lhs_rows[a:b], lhs_cols[a:b]
        lens = lens_all[k]
        total = int(lens.sum().item())
        if total == 0:
            continue
        rep = torch.repeat_interleave(torch.arange(b - a, device=dev), lens)
        start = torch.cumsum(lens, 0) - lens
        rpos = rhs_rowptr[k][rep] + (torch.arange(total, device=dev) - start[rep])
        keys.append(torch.unique(i[rep] * ncols_out + rhs_col[rpos].long()))
        del rep, start, rpos
    key = torch.cat(keys)
    return _rowptr_from_rows(key // ncols_out, nrows_out), (key % ncols_out).to(torch.int32).contiguous()


class _Level:
    pass


class AMGSymbolic:
    """Everything that depends only on the sparsity pattern (shared by all block sizes).

    On a partitioned mesh (``dist`` given) aggregates never cross rank boundaries and prolongator
    rows are smoothed over local couplings only (ghost couplings lumped into the diagonal), so ``P``
    maps owned fine nodes to owned coarse nodes only; the rows of the ghost nodes are imported from
    their owners, the Galerkin product of the owned coarse rows then needs no further communication
    and the coarse operator is the exact, globally symmetric ``P^T A P``.  Coarse ghost nodes are the
    owners' aggregates appearing in imported rows; every level gets its own halo plan
    (see ``_distributed_prolongator_pattern``).
    """

    MAX_DENSE = 256  # scalar rows of a coarsest level that is inverted densely (bs <= 3 -> <= 85 nodes)
    AGGLOMERATE = 20000  # global node count below which a distributed level is replicated on every rank
    DISTANCE = 2  # MIS(k) aggregation (see mis_aggregate)
    # levels with more block entries than this use the search-based Galerkin products instead of term lists
    # (8 bytes per scalar product term: 2-4x the operator itself)
    TERM_LIST_MAX_NNZ = int(os.environ.get("CASHOCS_B200_TERM_LIST_MAX_NNZ", 60_000_000))

    def __init__(self, rowptr, col, n, ncols, dist=None, max_levels: int = 12) -> None:
        self.levels: list[_Level] = []
        dev = col.device
        while True:
            L = _Level()
            L.rep_dist, L.rep_off, L.rep_pos, L.rep_nnzb_local, L.rep_gather = None, 0, None, 0, None
            n_glob = n if dist is None else int(round(dist.allreduce_scalar(float(n), "sum")))
            if dist is not None and self.levels and n_glob <= self.AGGLOMERATE:
                # agglomeration: below this size halo latency dominates, so the level (and everything
                # coarser) is replicated on every rank; the restriction all-reduces the rhs instead
                L.rep_dist, L.rep_nnzb_local = dist, int(col.numel())
                L.rep_rowptr_local, L.rep_col_local = rowptr, col  # this rank's rows in its own numbering
                rowptr, col, L.rep_off, L.rep_pos, gid, counts = self._replicate(rowptr, col, n, ncols, n_glob, dist)
                # all-gather of the restricted residual as a halo plan: the "ghosts" are all other ranks' rows
                owners = torch.cat([torch.full((c,), r, dtype=torch.int64) for r, c in enumerate(counts) if r != dist.rank]
                                   + [torch.zeros(0, dtype=torch.int64)])
                ids = torch.cat([torch.arange(c, dtype=torch.int64) for r, c in enumerate(counts) if r != dist.rank]
                                + [torch.zeros(0, dtype=torch.int64)])
                L.rep_gather = dist.derive_plan(owners, ids, n, dev)
                self.levels[-1].coarse_gid = gid.to(torch.int32).contiguous()  # local coarse node -> replicated row
                n = ncols = n_glob
                dist = None
            L.n, L.ncols, L.rowptr, L.col, L.dist = n, ncols, rowptr, col, dist
            L.nnzb = int(rowptr[-1].item())
            L.last = True
            L.iface = None
            L.coarse_gid = None
            self.levels.append(L)
            small = n_glob * 3 <= self.MAX_DENSE
            if small or len(self.levels) >= max_levels:
                break
            rows = _rows_of(rowptr)
            local = col < n
            agg, nc = mis_aggregate(rows[local], col[local].long(), n, self.DISTANCE)
            nc_glob = nc if dist is None else int(round(dist.allreduce_scalar(float(nc), "sum")))
            if nc_glob >= 0.8 * n_glob or nc_glob < 1:
                break
            agg = agg.long()
            new_dist = None
            if dist is None:
                agg_ext, gc = agg, 0
                pr, pc = rows, agg[col.long()]
                nct = nc
                key = torch.unique(pr * nct + pc)
                p_rows = key // nct
                L.p_rowptr = _rowptr_from_rows(p_rows, ncols)
                L.p_col = (key % nct).to(torch.int32).contiguous()
                L.nnzp_owned = int(L.p_col.numel())
                L.pplan = None
            else:
                gc, agg_ext, new_dist = self._distributed_prolongator_pattern(L, rows, col, local, agg, nc, dist)
                nct = nc + gc
            po_rows = _rows_of(L.p_rowptr[: n + 1])
            po_col = L.p_col[: L.nnzp_owned].long()
            order = torch.argsort(po_col * n + po_rows)
            L.pt_perm = order
            L.pt_col = po_rows[order].to(torch.int32).contiguous()
            L.pt_rowptr = _rowptr_from_rows(po_col[order], nc)
            L.agg, L.nc, L.gc = agg_ext.to(torch.int32).contiguous(), nc, gc
            nnza = col.numel()
            L.rowwise = nnza > self.TERM_LIST_MAX_NNZ
            if L.rowwise:  # patterns only; the numeric phase searches the output rows (cb_amg_spgemm_rowwise)
                L.ap_rowptr, L.ap_col = _expand_pattern(rows, col.long(), L.p_rowptr, L.p_col, n, nct)
                c_rowptr, c_col = _expand_pattern(_rows_of(L.pt_rowptr), L.pt_col.long(), L.ap_rowptr, L.ap_col, nc, nct)
                L.ap_seg = L.ap_sl = L.ap_sr = L.ac_seg = L.ac_sl = L.ac_sr = None
            else:
                L.ap_rowptr, L.ap_col, L.ap_seg, L.ap_sl, L.ap_sr = _expand_terms(
                    rows, col.long(), torch.arange(nnza, device=dev), L.p_rowptr, L.p_col, n, nct)
                c_rowptr, c_col, L.ac_seg, L.ac_sl, L.ac_sr = _expand_terms(
                    _rows_of(L.pt_rowptr), L.pt_col.long(), torch.arange(L.nnzp_owned, device=dev), L.ap_rowptr,
                    L.ap_col, nc, nct)
            L.last = False
            rowptr, col, n, ncols, dist = c_rowptr, c_col, nc, nct, new_dist

    @staticmethod
    def _distributed_prolongator_pattern(L, rows, col, local, agg, nc, dist):
        """Pattern of P on a partitioned level (collective).

        Aggregates are rank-local.  Owned rows are smoothed over their *local* couplings only (the
        numeric kernel lumps the dropped ghost couplings into the diagonal, so rows still sum to one):
        row i holds the aggregates of its owned neighbours, all owned coarse nodes.  The rows of the
        ghost nodes are imported from their owners (pattern here, values by ``L.pplan`` in every
        numeric phase), so ``A P`` of the owned rows -- and with it the owned rows of ``P^T A P`` -- are
        exact products with the global P and need no further communication.  Coarse ghost nodes are
        the owners' aggregates that appear in imported rows; they get their own halo plan.
        Returns (number of coarse ghosts, agg extended by a dummy ghost tail, coarse halo plan).
        """
        import torch.distributed as td

        from cashocs_b200 import distributed

        dev = col.device
        n, ncols = L.n, L.ncols
        key = torch.unique(rows[local] * nc + agg[col[local].long()])
        own_rows, own_cols = key // nc, key % nc
        own_rowptr = _rowptr_from_rows(own_rows, n)
        nnz_own = int(own_cols.numel())
        # export the rows of my send nodes, neighbour by neighbour, in send (= the receiver's ghost) order
        export, send_pos = {}, []
        for j, r in enumerate(dist.neigh):
            nodes = dist.send_idx[int(dist.send_off[j]): int(dist.send_off[j + 1])].long()
            lens = own_rowptr[nodes + 1] - own_rowptr[nodes]
            tot = int(lens.sum().item())
            rep = torch.repeat_interleave(torch.arange(nodes.numel(), device=dev), lens)
            start = torch.cumsum(lens, 0) - lens
            pos = own_rowptr[nodes][rep] + (torch.arange(tot, device=dev) - start[rep])
            send_pos.append(pos)
            export[int(r)] = (lens.cpu(), own_cols[pos].cpu())
        allexp = [None] * dist.world
        td.all_gather_object(allexp, export)
        g_lens, g_owner, g_id, recv_counts = [], [], [], []
        for j, r in enumerate(dist.neigh):
            lens, ids = allexp[int(r)][dist.rank]
            assert lens.numel() == int(dist.recv_off[j + 1] - dist.recv_off[j])
            g_lens.append(lens)
            g_id.append(ids)
            g_owner.append(torch.full((ids.numel(),), int(r), dtype=torch.int64))
            recv_counts.append(int(ids.numel()))
        g_lens = torch.cat(g_lens).to(dev) if g_lens else torch.zeros(0, dtype=torch.int64, device=dev)
        g_id = torch.cat(g_id).to(dev) if g_id else torch.zeros(0, dtype=torch.int64, device=dev)
        g_owner = torch.cat(g_owner).to(dev) if g_owner else torch.zeros(0, dtype=torch.int64, device=dev)
        big = int(g_id.max().item()) + 1 if g_id.numel() else 1
        ukey, inv = torch.unique(g_owner * big + g_id, return_inverse=True)  # sorted by (owner, id)
        gc = int(ukey.numel())
        new_dist = dist.derive_plan(ukey // big, ukey % big, nc, dev)
        L.p_rowptr = torch.cat([own_rowptr, own_rowptr[-1] + torch.cumsum(g_lens, 0)])
        L.p_col = torch.cat([own_cols, nc + inv]).to(torch.int32).contiguous()
        L.nnzp_owned = nnz_own
        # halo plan over prolongator ENTRIES: ghost-row values follow the owners' values
        pplan = distributed.DistContext(dist.rank, dist.world, dist.transport)
        pplan.parent = dist if dist.parent is None else dist.parent
        pplan.set_plan(dist.neigh, send_pos, recv_counts, nnz_own, int(g_id.numel()), dev)
        L.pplan = pplan
        agg_ext = torch.cat([agg, torch.zeros(ncols - n, dtype=agg.dtype, device=dev)])
        return gc, agg_ext, new_dist

    @staticmethod
    def _replicate(rowptr, col, n, ncols, n_glob, dist):
        """Global pattern of a distributed level, identical on every rank (collective).

        Global node ids are rank offset + local id.  Returns (rowptr, col) of the global CSR, this
        rank's node offset and, for every local entry, its position in the global value array.
        """
        import torch.distributed as td

        dev = col.device
        counts = [None] * dist.world
        td.all_gather_object(counts, int(n))
        off = int(sum(counts[: dist.rank]))
        v = torch.zeros(ncols, dtype=torch.float64, device=dev)
        v[:n] = torch.arange(off, off + n, dtype=torch.float64, device=dev)
        dist.halo_exchange(v, 1)
        gid = v.round().long()
        key = gid[_rows_of(rowptr)] * n_glob + gid[col.long()]
        allk = [None] * dist.world
        td.all_gather_object(allk, key.cpu())
        start = int(sum(k.numel() for k in allk[: dist.rank]))
        allkeys = torch.cat(allk).to(dev)
        skeys, order = torch.sort(allkeys)
        if skeys.numel() > 1 and bool((skeys[1:] == skeys[:-1]).any()):
            raise RuntimeError("replicated AMG level: a matrix row is owned by two ranks")
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=dev)
        pos = inv[start: start + key.numel()].contiguous()
        g_rowptr = _rowptr_from_rows(skeys // n_glob, n_glob)
        g_col = (skeys % n_glob).to(torch.int32).contiguous()
        return g_rowptr, g_col, off, pos, gid, [int(c) for c in counts]

    @property
    def sizes(self):
        return [L.n for L in self.levels]

    @property
    def operator_complexity(self) -> float:
        return sum(L.nnzb for L in self.levels) / self.levels[0].nnzb


class AMGHierarchy:
    """Values of the hierarchy for one block size; ``numeric(A)`` refreshes them for new matrix values."""

    def __init__(self, A, symbolic: AMGSymbolic | None = None, smooth_degree: int = 1, coarse_degree: int = 16,
                 smooth_ratio: float = 20.0, coarse_ratio: float = 40.0, mixed: bool = True,
                 level_smooth_degree: int = 0, mixed_mask=None, lambda_safety: float = 1.1,
                 prolongator_damping: float = 1.6) -> None:
        self.bs = bs = A.bs
        # Mixed precision: the fine-level operator gets an fp32 copy of its values, which is what the residual pass
        # of the V-cycle streams (cb_mat.val32; vectors, products and sums stay fp64 -- see csrc/spmv.cuh).  Measured
        # on the 10 M-tet elasticity operator (tools/sweep_precision.py): fp32 for the fine-level residual leaves the
        # PCG iteration count alone (89 -> 86), fp32 copies of A*P ("ap0") or of the coarse operators ("coarse") do
        # not (-> 110: the rounding defect of an entry is O(1e-7 |A|) while A e is O(|A| / kappa) for the smooth
        # vectors those operators act on), so they stay fp64 unless asked for.
        self.mixed = bool(mixed)
        self.mixed_mask = set(mixed_mask) if mixed_mask is not None else {"a0"}
        self.device = dev = A.val.device
        self.smooth_degree, self.coarse_degree = int(smooth_degree), int(coarse_degree)
        self.level_smooth_degree = int(level_smooth_degree)
        self.lambda_safety, self.prolongator_damping = float(lambda_safety), float(prolongator_damping)
        self.smooth_ratio, self.coarse_ratio = float(smooth_ratio), float(coarse_ratio)
        self.symbolic = symbolic if symbolic is not None else AMGSymbolic(A.rowptr, A.col, A.nrows, A.ncols,
                                                                         dist=A.mesh.dist)
        self.levels = self.symbolic.levels
        self.distributed = self.levels[0].dist is not None
        self.version = None
        self.power_its = 12      # power-method steps of the first numeric phase (cold start) ...
        self.power_its_warm = 4  # ... and of the later ones (restart from the previous dominant vector)
        self.use_ap = True  # first post-smoothing step through A*P (csrc/spmv.cu k_spmv_post)
        self._power_x: dict = {}
        f64 = dict(dtype=torch.float64, device=dev)
        self.data = []
        for li, L in enumerate(self.levels):
            D = _Level()
            nb = L.n * bs
            D.val = None if li == 0 else torch.zeros(L.nnzb * bs * bs, **f64)
            # rows of a replicated level computed by this rank (scattered + all-reduced into D.val)
            D.val_local = torch.zeros(L.rep_nnzb_local * bs * bs, **f64) if L.rep_dist is not None else None
            D.rep_tmp = torch.zeros(nb, **f64) if L.rep_gather is not None else None
            want32 = self.mixed and (("a0" in self.mixed_mask) if li == 0 else ("coarse" in self.mixed_mask or "a1" in self.mixed_mask))
            D.val32 = torch.zeros(L.nnzb * bs * bs, dtype=torch.float32, device=dev) if want32 else None
            D.ap_val32 = None
            D.dinv = torch.zeros(nb, **f64)
            D.t = torch.zeros(nb, **f64)
            D.d = torch.zeros(nb, **f64)
            D.x = None if li == 0 else torch.zeros(L.ncols * bs, **f64)
            D.b = None if li == 0 else torch.zeros(nb, **f64)
            D.sendbuf = None
            if L.dist is not None:
                D.sendbuf = torch.zeros(max(int(L.dist.send_off[-1]) * bs, 1), **f64)
            D.lmax = 2.0
            if not L.last:
                D.p_val = torch.ones(L.p_col.numel(), **f64)  # ghost rows stay 1.0 (tentative)
                D.pt_val = torch.zeros(L.nnzp_owned, **f64)
                D.ap_val = torch.zeros(L.ap_col.numel() * bs * bs, **f64)
                if self.mixed and (("ap0" in self.mixed_mask) if li == 0 else ("coarse" in self.mixed_mask)):
                    D.ap_val32 = torch.zeros(L.ap_col.numel() * bs * bs, dtype=torch.float32, device=dev)
                D.xc_local = torch.zeros(L.coarse_gid.numel() * bs, **f64) if L.coarse_gid is not None else None
                D.s_val = torch.zeros(L.nnzb, **f64) if bs > 1 else None
                D.s_dinv = torch.zeros(L.n, **f64)
            self.data.append(D)
        last = self.levels[-1]
        self.coarse_inv = self.coarse_work = None
        if last.dist is None and last.n * bs <= AMGSymbolic.MAX_DENSE and len(self.levels) > 1:
            nn = last.n * bs
            self.coarse_inv = torch.zeros(nn * nn, **f64)
            self.coarse_work = torch.zeros(2 * nn * nn, **f64)
        self.scratch = torch.zeros(int(_lib.load().cb_reduce_scratch_len()), **f64)
        self.lmax_dev = torch.zeros(1, **f64)
        self._struct = None
        self._level_array = None

    @property
    def sizes(self):
        return self.symbolic.sizes

    @property
    def operator_complexity(self) -> float:
        return self.symbolic.operator_complexity

    # ------------------------------------------------------------------ numeric
    def _mat(self, L, val, bs, val32=None) -> _lib.cb_mat:
        from cashocs_b200._utils import linalg

        m = _lib.cb_mat()
        m.nrows, m.ncols, m.bs = L.n, L.ncols, bs
        m.row_split = linalg.row_split_hint(L.nnzb, L.n)
        m.rowptr, m.col, m.val = L.rowptr.data_ptr(), L.col.data_ptr(), val.data_ptr()
        m.val32 = val32.data_ptr() if val32 is not None else None
        return m

    def _to_f32(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        _lib.check(_lib.load().cb_to_f32(_lib.stream_ptr(), src.numel(), _lib.ptr(src), _lib.ptr(dst)), "cb_to_f32")

    def _diag_inv(self, m, dinv, L, nscalar: int, safe: bool = False) -> float:
        """Inverse diagonal and an upper estimate of lambda_max(D^-1 A).

        Gershgorin is a safe bound but 2-6x too large for elasticity rows and for smoothed-aggregation
        coarse operators, which would push the Chebyshev interval off the spectrum; so the bound is
        tightened with `power_its` steps of the power method on D^-1 A (fixed start vector ->
        deterministic) times a 1.15 safety factor, never exceeding Gershgorin.  On a partitioned
        mesh both numbers are global (halo exchange per step, all-reduced norms), so every rank
        builds the same Chebyshev polynomials.
        """
        lib = _lib.load()
        s = _lib.stream_ptr()
        _lib.check(lib.cb_extract_diag_inv(s, C.byref(m), _lib.ptr(dinv), _lib.ptr(self.lmax_dev),
                                           _lib.ptr(self.scratch), self.scratch.numel()), "cb_extract_diag_inv")
        if safe:
            # coarsest level without a dense inverse: a degree-16 Chebyshev "solve" turns indefinite
            # if lambda_max is underestimated at all, so take the rigorous Gershgorin bound
            if L.dist is not None:
                L.dist.allreduce(self.lmax_dev, "max")
            return float(self.lmax_dev.item())
        comp = nscalar // L.n
        ncol = L.ncols * comp
        st = self._power_x.get((id(L), nscalar))
        if st is None:
            gen = torch.Generator(device="cpu").manual_seed(1234 + nscalar)
            x0 = (torch.rand(nscalar, generator=gen, dtype=torch.float64) + 0.5).to(self.device)
            st = _Level()
            st.bufs = [torch.zeros(ncol, dtype=torch.float64, device=self.device) for _ in range(2)]
            st.bufs[0][:nscalar] = x0
            st.cur, st.alpha, st.warm = 0, 1.0, False
            st.norms = torch.zeros(max(self.power_its, self.power_its_warm) + 1, dtype=torch.float64, device=self.device)
            self._power_x[(id(L), nscalar)] = st
        # No normalisation between the steps (lambda^steps stays far from overflow): every step is ONE fused pass
        # y = alpha D^-1 A x with |y|^2, and the estimate is the last growth factor.  After the first numeric phase
        # the iteration restarts from the previous matrix's dominant vector (the mesh moved a little: it is still an
        # excellent start), so a few steps reproduce what twelve cold ones give.
        steps = self.power_its_warm if st.warm else self.power_its
        alpha = st.alpha
        for k in range(steps):
            xin, y = st.bufs[st.cur], st.bufs[1 - st.cur]
            if L.dist is not None:
                L.dist.halo_exchange(xin, comp)
            _lib.check(lib.cb_power_step(s, C.byref(m), _lib.ptr(dinv), alpha, _lib.ptr(xin), _lib.ptr(y),
                                         _lib.ptr(st.norms[k:]), _lib.ptr(self.scratch), self.scratch.numel(), 1),
                       "cb_power_step")
            alpha = 1.0
            st.cur = 1 - st.cur
        if L.dist is not None:
            L.dist.allreduce(self.lmax_dev, "max")
            L.dist.allreduce(st.norms, "sum")
        both = torch.cat([self.lmax_dev, st.norms[:steps]]).cpu()  # one host sync per estimate
        n_last, n_prev = float(both[steps]), float(both[steps - 1])
        if st.warm or steps >= 2:
            # warm start: the input of the first step was normalised (alpha), so for steps == 1 the growth is sqrt(n_last)
            lam = (n_last / n_prev) ** 0.5 if steps >= 2 else n_last ** 0.5
        else:
            lam = float("inf")
        st.alpha = 1.0 / n_last ** 0.5 if n_last > 0.0 else 1.0
        st.warm = True
        return min(float(both[0]), self.lambda_safety * lam)

    def numeric(self, A) -> None:
        """Recompute prolongator values, Galerkin operators and smoother data for the values of ``A``."""
        lib = _lib.load()
        s = _lib.stream_ptr()
        bs = self.bs
        val = A.val
        for li, (L, D) in enumerate(zip(self.levels, self.data)):
            D.cur_val = val
            if D.val32 is not None:
                self._to_f32(val, D.val32)
            m = self._mat(L, val, bs, D.val32)
            if L.last and self.coarse_inv is not None:
                _lib.check(lib.cb_amg_dense_inverse(s, C.byref(m), _lib.ptr(self.coarse_work), _lib.ptr(self.coarse_inv)),
                           "cb_amg_dense_inverse")
                break
            D.lmax = self._diag_inv(m, D.dinv, L, L.n * bs, safe=L.last)
            if L.last:
                break
            if bs > 1:
                _lib.check(lib.cb_amg_block_trace(s, C.byref(m), _lib.ptr(D.s_val)), "cb_amg_block_trace")
                sval = D.s_val
            else:
                sval = val
            ms = self._mat(L, sval, 1)
            if bs == 1:  # the strength matrix is the level operator itself
                D.s_dinv, lmax_s = D.dinv, D.lmax
            else:
                lmax_s = self._diag_inv(ms, D.s_dinv, L, L.n)
            omega = self.prolongator_damping / lmax_s
            _lib.check(lib.cb_amg_smooth_prolongator(s, C.byref(ms), omega, _lib.ptr(L.agg), _lib.ptr(L.iface),
                                                     _lib.ptr(L.p_rowptr), _lib.ptr(L.p_col), _lib.ptr(D.p_val)),
                       "cb_amg_smooth_prolongator")
            if L.pplan is not None:  # prolongator rows of the ghost nodes follow their owners' values
                L.pplan.halo_exchange(D.p_val, 1)
            torch.index_select(D.p_val[: L.nnzp_owned], 0, L.pt_perm, out=D.pt_val)
            nxt, nxtL = self.data[li + 1], self.levels[li + 1]
            rep = nxtL.rep_dist is not None
            out = nxt.val_local if rep else nxt.val
            if L.rowwise:
                _lib.check(lib.cb_amg_spgemm_rowwise(s, bs, 1, L.n, _lib.ptr(L.rowptr), _lib.ptr(L.col), _lib.ptr(val),
                                                     _lib.ptr(L.p_rowptr), _lib.ptr(L.p_col), _lib.ptr(D.p_val),
                                                     _lib.ptr(L.ap_rowptr), _lib.ptr(L.ap_col), _lib.ptr(D.ap_val)),
                           "cb_amg_spgemm_rowwise(AP)")
                _lib.check(lib.cb_amg_spgemm_rowwise(s, bs, 0, L.nc, _lib.ptr(L.pt_rowptr), _lib.ptr(L.pt_col),
                                                     _lib.ptr(D.pt_val), _lib.ptr(L.ap_rowptr), _lib.ptr(L.ap_col),
                                                     _lib.ptr(D.ap_val), _lib.ptr(nxtL.rep_rowptr_local if rep else nxtL.rowptr),
                                                     _lib.ptr(nxtL.rep_col_local if rep else nxtL.col), _lib.ptr(out)),
                           "cb_amg_spgemm_rowwise(PtAP)")
            else:
                _lib.check(lib.cb_amg_spgemm_terms(s, bs, 1, L.ap_col.numel(), _lib.ptr(L.ap_seg), _lib.ptr(L.ap_sl),
                                                   _lib.ptr(L.ap_sr), _lib.ptr(val), _lib.ptr(D.p_val), _lib.ptr(D.ap_val)),
                           "cb_amg_spgemm_terms(AP)")
                _lib.check(lib.cb_amg_spgemm_terms(s, bs, 0, out.numel() // (bs * bs), _lib.ptr(L.ac_seg), _lib.ptr(L.ac_sl),
                                                   _lib.ptr(L.ac_sr), _lib.ptr(D.pt_val), _lib.ptr(D.ap_val), _lib.ptr(out)),
                           "cb_amg_spgemm_terms(PtAP)")
            if D.ap_val32 is not None:
                self._to_f32(D.ap_val, D.ap_val32)
            if rep:  # every rank contributes its rows of the replicated operator
                nxt.val.zero_()
                nxt.val.view(-1, bs * bs)[nxtL.rep_pos] = out.view(-1, bs * bs)
                nxtL.rep_dist.allreduce(nxt.val, "sum")
            val = nxt.val
        self._struct = None

    def struct(self) -> _lib.cb_amg:
        """The C struct (kept alive by this object) for ``cb_ksp_opts.amg``."""
        if self._struct is not None:
            return self._struct
        arr = (_lib.cb_amg_level * len(self.levels))()
        for li, (L, D) in enumerate(zip(self.levels, self.data)):
            e = arr[li]
            e.A = self._mat(L, D.cur_val, self.bs, D.val32)
            e.dinv = D.dinv.data_ptr()
            e.lmax = D.lmax
            e.nc = 0 if L.last else L.nc
            if not L.last:
                e.p_rowptr, e.p_col, e.p_val = L.p_rowptr.data_ptr(), L.p_col.data_ptr(), D.p_val.data_ptr()
                e.pt_rowptr, e.pt_col, e.pt_val = L.pt_rowptr.data_ptr(), L.pt_col.data_ptr(), D.pt_val.data_ptr()
            e.x = D.x.data_ptr() if D.x is not None else None
            e.b = D.b.data_ptr() if D.b is not None else None
            e.t, e.d = D.t.data_ptr(), D.d.data_ptr()
            if L.dist is not None:
                e.dist = L.dist.handle
                e.sendbuf = D.sendbuf.data_ptr()
            if L.rep_dist is not None:
                e.rep_dist = L.rep_dist.handle
                e.rep_off = L.rep_off
                if L.rep_gather is not None and L.rep_gather.handle is not None:
                    e.rep_gather = L.rep_gather.handle
                    e.rep_tmp = D.rep_tmp.data_ptr()
            if not L.last and self.use_ap:
                from cashocs_b200._utils import linalg

                ap = _lib.cb_mat()
                ap.nrows, ap.ncols, ap.bs = L.n, L.nc + L.gc, self.bs
                ap.row_split = linalg.row_split_hint(L.ap_col.numel(), L.n)
                ap.cta_warps = linalg.cta_warps_hint(L.ap_col.numel(), L.n, self.bs, ap.row_split)
                ap.rowptr, ap.col, ap.val = L.ap_rowptr.data_ptr(), L.ap_col.data_ptr(), D.ap_val.data_ptr()
                ap.val32 = D.ap_val32.data_ptr() if D.ap_val32 is not None else None
                e.AP = ap
                if L.coarse_gid is not None:
                    e.coarse_gid = L.coarse_gid.data_ptr()
                    e.xc_local = D.xc_local.data_ptr()
        h = _lib.cb_amg()
        h.nlevels = len(self.levels)
        h.smooth_degree, h.coarse_degree = self.smooth_degree, self.coarse_degree
        h.smooth_ratio, h.coarse_ratio = self.smooth_ratio, self.coarse_ratio
        h.level_smooth_degree = self.level_smooth_degree
        h.levels = C.cast(arr, C.c_void_p)
        h.coarse_inv = self.coarse_inv.data_ptr() if self.coarse_inv is not None else None
        self._level_array = arr
        self._struct = h
        return h
