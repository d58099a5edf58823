"""Smoothed-aggregation AMG hierarchy: symbolic phase (host layer, torch index plumbing) and the
driver of the numeric phase (C-ABI kernels in ``csrc/amg.cu``).

This is what the reference's iterative default ``pc_type hypre`` / ``boomeramg``
(``cashocs/_utils/linalg.py:44-52``; the Lame-mu solve ``cashocs/_forms/shape_form_handler.py:113-120``)
and PETSc's ``pc_type gamg`` map to on the device.  Unknown-based smoothed aggregation with one
scalar prolongator shared by all components, so every level is a block-CSR matrix of the fine
block size and runs through the same SpMV kernel.

* Symbolic (once per sparsity pattern, i.e. once per mesh topology): MIS(1) aggregation on the
  matrix graph with hashed priorities (integer arithmetic only -> deterministic), patterns of
  ``P``, ``P^T``, ``A P`` and ``P^T A P``.
* Numeric (every time the matrix values change): strength matrix ``S = tr(A_ij)/bs``, smoothed
  prolongator ``P = (I - 4/(3 lmax) D_S^-1 S) T``, Galerkin operators, inverse diagonals and
  Gershgorin bounds for the Chebyshev smoothers.
"""

from __future__ import annotations

import ctypes as C

import torch

from cashocs_b200 import _lib


def _rows_of(rowptr: torch.Tensor) -> torch.Tensor:
    n = rowptr.numel() - 1
    return torch.repeat_interleave(torch.arange(n, device=rowptr.device), rowptr[1:] - rowptr[:-1])


def _rowptr_from_rows(rows: torch.Tensor, n: int) -> torch.Tensor:
    rp = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return rp


def mis_aggregate(rows: torch.Tensor, cols: torch.Tensor, n: int):
    """MIS(1) aggregation: roots form a maximal independent set, every other node joins its
    highest-priority adjacent root.  Returns (agg [n] int32, number of aggregates)."""
    dev = rows.device
    off = rows != cols
    r, c = rows[off], cols[off]
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    h = (idx * 2654435761 + 12345) % 2147483629
    h = (h * 40503 + 7) % 2147483629
    prio = h * n + idx  # unique, < 2^31 * n
    state = torch.zeros(n, dtype=torch.int8, device=dev)  # 0 undecided, 1 root, 2 covered
    neg = torch.full((n,), -1, dtype=torch.int64, device=dev)
    for _ in range(200):
        und = state == 0
        if not bool(und.any()):
            break
        p = torch.where(und, prio, neg)
        nbmax = neg.clone().scatter_reduce_(0, r, p[c], "amax", include_self=True)
        newroot = und & (p > nbmax)
        state[newroot] = 1
        cov = torch.zeros(n, dtype=torch.int8, device=dev).scatter_reduce_(0, r, newroot[c].to(torch.int8), "amax")
        state[(state == 0) & (cov > 0)] = 2
    is_root = state == 1
    nc = int(is_root.sum().item())
    agg = torch.full((n,), -1, dtype=torch.int64, device=dev)
    agg[is_root] = torch.arange(nc, device=dev)
    best = neg.clone().scatter_reduce_(0, r, torch.where(is_root[c], prio[c], neg[c]), "amax", include_self=True)
    nonroot = ~is_root
    agg[nonroot] = agg[(best % n)[nonroot]]
    return agg.to(torch.int32), nc


def _expand_unique(lhs_rows, lhs_cols, rhs_rowptr, rhs_col, nrows_out, ncols_out, chunk_rows=400_000):
    """Pattern of (L * R): unique (row(L), col(R)) pairs, returned as CSR (rowptr, col int32)."""
    dev = lhs_rows.device
    lens_all = rhs_rowptr[1:] - rhs_rowptr[:-1]
    keys = []
    # rows of L are sorted ascending: process row ranges so the expansion stays bounded
    bounds = torch.searchsorted(lhs_rows, torch.arange(0, nrows_out + chunk_rows, chunk_rows, device=dev))
    bounds = bounds.tolist()
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
        offs = torch.arange(total, device=dev) - start[rep]
        j = rhs_col[rhs_rowptr[k][rep] + offs].long()
        keys.append(torch.unique(i[rep] * ncols_out + j))
    key = torch.cat(keys) if keys else torch.zeros(0, dtype=torch.int64, device=dev)
    out_rows = key // ncols_out
    return _rowptr_from_rows(out_rows, nrows_out), (key % ncols_out).to(torch.int32).contiguous()


class _Level:
    pass


class AMGHierarchy:
    """Hierarchy for one block-CSR sparsity pattern; ``numeric(A)`` refreshes the values."""

    def __init__(self, A, max_levels: int = 10, min_coarse: int = 300, smooth_degree: int = 2,
                 coarse_degree: int = 16, smooth_ratio: float = 4.0, coarse_ratio: float = 40.0) -> None:
        self.bs = A.bs
        self.device = A.val.device
        self.smooth_degree, self.coarse_degree = int(smooth_degree), int(coarse_degree)
        self.smooth_ratio, self.coarse_ratio = float(smooth_ratio), float(coarse_ratio)
        self.levels: list[_Level] = []
        self.version = None
        self._symbolic(A, max_levels, min_coarse)
        self._struct = None
        self._level_array = None

    # ------------------------------------------------------------------ symbolic
    def _symbolic(self, A, max_levels, min_coarse) -> None:
        dev, bs = self.device, self.bs
        rowptr, col, n = A.rowptr, A.col, A.nrows
        ncols = A.ncols
        while True:
            L = _Level()
            L.n, L.ncols, L.rowptr, L.col = n, ncols, rowptr, col
            L.nnzb = int(rowptr[-1].item())
            L.last = True
            self.levels.append(L)
            if n <= min_coarse or len(self.levels) >= max_levels:
                break
            rows = _rows_of(rowptr)
            keep = col < n  # local columns only (ghost couplings are not part of the preconditioner)
            r, c = rows[keep], col[keep].long()
            agg, nc = mis_aggregate(r, c, n)
            if nc >= 0.8 * n or nc < 1:
                break
            key = torch.unique(r * nc + agg.long()[c])
            p_rows = key // nc
            L.p_rowptr = _rowptr_from_rows(p_rows, n)
            L.p_col = (key % nc).to(torch.int32).contiguous()
            order = torch.argsort(L.p_col.long() * n + p_rows)
            L.pt_perm = order
            L.pt_col = p_rows[order].to(torch.int32).contiguous()
            L.pt_rowptr = _rowptr_from_rows(L.p_col.long()[order], nc)
            L.agg, L.nc = agg.contiguous(), nc
            L.ap_rowptr, L.ap_col = _expand_unique(r, c, L.p_rowptr, L.p_col, n, nc)
            c_rowptr, c_col = _expand_unique(_rows_of(L.pt_rowptr), L.pt_col.long(), L.ap_rowptr, L.ap_col, nc, nc)
            L.last = False
            rowptr, col, n, ncols = c_rowptr, c_col, nc, nc
        f64 = dict(dtype=torch.float64, device=dev)
        for li, L in enumerate(self.levels):
            nb = L.n * bs
            L.val = None if li == 0 else torch.zeros(L.nnzb * bs * bs, **f64)
            L.dinv = torch.zeros(nb, **f64)
            L.t = torch.zeros(nb, **f64)
            L.d = torch.zeros(nb, **f64)
            L.x = None if li == 0 else torch.zeros(L.ncols * bs, **f64)
            L.b = None if li == 0 else torch.zeros(nb, **f64)
            L.lmax = 2.0
            if not L.last:
                nnzp = L.p_col.numel()
                L.p_val = torch.zeros(nnzp, **f64)
                L.pt_val = torch.zeros(nnzp, **f64)
                L.ap_val = torch.zeros(L.ap_col.numel() * bs * bs, **f64)
                L.s_val = torch.zeros(L.nnzb, **f64) if bs > 1 else None
                L.s_dinv = torch.zeros(L.n, **f64)
        self.scratch = torch.zeros(int(_lib.load().cb_reduce_scratch_len()), **f64)
        self.lmax_dev = torch.zeros(1, **f64)

    @property
    def sizes(self):
        return [L.n for L in self.levels]

    @property
    def operator_complexity(self) -> float:
        return sum(L.nnzb for L in self.levels) / self.levels[0].nnzb

    # ------------------------------------------------------------------ numeric
    def _mat(self, L, val, bs) -> _lib.cb_mat:
        m = _lib.cb_mat()
        m.nrows, m.ncols, m.bs = L.n, L.ncols, bs
        m.rowptr, m.col, m.val = L.rowptr.data_ptr(), L.col.data_ptr(), val.data_ptr()
        return m

    def _diag_inv(self, m, dinv) -> float:
        lib = _lib.load()
        _lib.check(lib.cb_extract_diag_inv(_lib.stream_ptr(), C.byref(m), _lib.ptr(dinv), _lib.ptr(self.lmax_dev),
                                           _lib.ptr(self.scratch), self.scratch.numel()), "cb_extract_diag_inv")
        return float(self.lmax_dev.item())

    def numeric(self, A) -> None:
        """Recompute prolongator values, Galerkin operators and smoother data for the values of ``A``."""
        lib = _lib.load()
        s = _lib.stream_ptr()
        bs = self.bs
        val = A.val
        for li, L in enumerate(self.levels):
            L.cur_val = val
            m = self._mat(L, val, bs)
            L.lmax = self._diag_inv(m, L.dinv)
            if L.last:
                break
            if bs > 1:
                _lib.check(lib.cb_amg_block_trace(s, C.byref(m), _lib.ptr(L.s_val)), "cb_amg_block_trace")
                sval = L.s_val
            else:
                sval = val
            ms = self._mat(L, sval, 1)
            lmax_s = self._diag_inv(ms, L.s_dinv)
            omega = (4.0 / 3.0) / lmax_s
            _lib.check(lib.cb_amg_smooth_prolongator(s, C.byref(ms), omega, _lib.ptr(L.agg), _lib.ptr(L.p_rowptr),
                                                     _lib.ptr(L.p_col), _lib.ptr(L.p_val)), "cb_amg_smooth_prolongator")
            torch.index_select(L.p_val, 0, L.pt_perm, out=L.pt_val)
            nxt = self.levels[li + 1]
            mc = self._mat(nxt, nxt.val, bs)
            _lib.check(lib.cb_amg_galerkin(s, C.byref(m), _lib.ptr(L.p_rowptr), _lib.ptr(L.p_col), _lib.ptr(L.p_val),
                                           _lib.ptr(L.pt_rowptr), _lib.ptr(L.pt_col), _lib.ptr(L.pt_val),
                                           _lib.ptr(L.ap_rowptr), _lib.ptr(L.ap_col), _lib.ptr(L.ap_val), C.byref(mc)),
                       "cb_amg_galerkin")
            val = nxt.val
        self._struct = None

    def struct(self) -> _lib.cb_amg:
        """The C struct (kept alive by this object) for ``cb_ksp_opts.amg``."""
        if self._struct is not None:
            return self._struct
        arr = (_lib.cb_amg_level * len(self.levels))()
        for li, L in enumerate(self.levels):
            e = arr[li]
            e.A = self._mat(L, L.cur_val, self.bs)
            e.dinv = L.dinv.data_ptr()
            e.lmax = L.lmax
            e.nc = 0 if L.last else L.nc
            if not L.last:
                e.p_rowptr, e.p_col, e.p_val = L.p_rowptr.data_ptr(), L.p_col.data_ptr(), L.p_val.data_ptr()
                e.pt_rowptr, e.pt_col, e.pt_val = L.pt_rowptr.data_ptr(), L.pt_col.data_ptr(), L.pt_val.data_ptr()
            e.x = L.x.data_ptr() if L.x is not None else None
            e.b = L.b.data_ptr() if L.b is not None else None
            e.t, e.d = L.t.data_ptr(), L.d.data_ptr()
        h = _lib.cb_amg()
        h.nlevels = len(self.levels)
        h.smooth_degree, h.coarse_degree = self.smooth_degree, self.coarse_degree
        h.smooth_ratio, h.coarse_ratio = self.smooth_ratio, self.coarse_ratio
        h.levels = C.cast(arr, C.c_void_p)
        self._level_array = arr
        self._struct = h
        return h
