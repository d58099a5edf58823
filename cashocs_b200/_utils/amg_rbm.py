"""Smoothed aggregation with the rigid-body modes *inside* the hierarchy -- symbolic phase.

WORK IN PROGRESS (round 2): this module and ``csrc/amg_rbm.cu`` are not wired into ``cb_ksp_solve`` yet.  The
symbolic phase below and the arithmetic of the kernels (restated in numpy by ``tests/test_amg_rbm.py``, which
walks the very term lists built here) are checked against scipy on the CPU; the kernels themselves have been
compiled but not yet run on a GPU.  The measured motivation is in ``DESIGN.md`` section 8 and
``tools/proto_amg_rbm.py``: 89 -> ~36 PCG iterations for the elasticity Riesz solve.

Design (everything stays a d x d block-CSR matrix, so all level operators keep running through the tuned SpMV):
every aggregate J becomes TWO coarse nodes, ``2J`` carrying the translations and ``2J+1`` the rotations about the
aggregate's centroid ``C_J`` scaled by its radius ``S_J``.  The tentative prolongator has the blocks

    translation node i in J:  T[i, 2J] = I,   T[i, 2J+1] = cross(x_i - C_J) / S_J      (w -> w x r)
    rotation node i in J:                      T[i, 2J+1] = (s_i / S_J) I

(level 0 has translation nodes only; on coarser levels the two nodes of an aggregate stay in one group and are
aggregated together).  ``P = T - omega D^-1 (A T)`` with the point-Jacobi ``D`` the smoothers use, on the pattern
of ``A T``; ``A_c = P^T (A P)``.  All three products are gathers over term lists (one thread per output block
element, fixed summation order), like the scalar-prolongator hierarchy of ``amg.py``.

Only the pattern-dependent part lives here; centroids, radii and all values belong to the numeric phase because the
mesh moves between two matrices.
"""

from __future__ import annotations

import torch

from cashocs_b200._utils.amg import _expand_terms, _rowptr_from_rows, _rows_of, mis_aggregate


class _Level:
    pass


class RBMSymbolic:
    """Patterns, term lists and aggregate membership of every level (torch tensors on the device of ``col``)."""

    MAX_DENSE_NODES = 85  # coarsest level inverted densely (<= 256 scalar rows)
    DISTANCE = 2

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, n: int, max_levels: int = 12) -> None:
        dev = col.device
        self.levels: list[_Level] = []
        kind = torch.zeros(n, dtype=torch.uint8, device=dev)      # 0 translation / ordinary node, 1 rotation node
        group = torch.arange(n, dtype=torch.int64, device=dev)    # nodes of one aggregate pair share a group
        while True:
            L = _Level()
            L.n, L.rowptr, L.col = n, rowptr, col
            L.nnzb = int(rowptr[-1].item())
            L.kind, L.group = kind, group
            L.last = True
            self.levels.append(L)
            if n <= self.MAX_DENSE_NODES or len(self.levels) >= max_levels:
                break
            rows = _rows_of(rowptr)
            # aggregate the graph of groups
            ng = int(group.max().item()) + 1
            key = torch.unique(group[rows] * ng + group[col.long()])
            agg_g, nc = mis_aggregate(key // ng, key % ng, ng, self.DISTANCE)
            if nc >= 0.8 * ng or nc < 1:
                break
            agg = agg_g.long()[group]
            # tentative prolongator: translation nodes -> (2J, 2J+1), rotation nodes -> (2J+1)
            is_t = kind == 0
            t_len = torch.where(is_t, 2, 1)
            L.t_rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            L.t_rowptr[1:] = torch.cumsum(t_len, 0)
            nt = int(L.t_rowptr[-1].item())
            t_rows = _rows_of(L.t_rowptr)
            first = L.t_rowptr[:-1][t_rows] == torch.arange(nt, device=dev)  # first entry of its row
            t_col = 2 * agg[t_rows] + 1
            t_col[first & is_t[t_rows]] -= 1
            L.t_col = t_col.to(torch.int32).contiguous()
            L.agg = agg.to(torch.int32).contiguous()
            L.nc = nc
            ncn = 2 * nc
            # aggregate -> member nodes (for centroids and radii in the numeric phase)
            order = torch.argsort(agg, stable=True)
            L.member_idx = order.to(torch.int32).contiguous()
            L.member_ptr = _rowptr_from_rows(agg[order], nc)
            # P has the pattern of A*T; terms (A entry, T entry)
            nnza = int(col.numel())
            L.p_rowptr, L.p_col, L.at_seg, L.at_sl, L.at_sr = _expand_terms(
                rows, col.long(), torch.arange(nnza, device=dev), L.t_rowptr, L.t_col, n, ncn)
            p_rows = _rows_of(L.p_rowptr)
            # position of every T entry inside P (T's pattern is a subset: A holds its diagonal)
            pkey = p_rows * ncn + L.p_col.long()
            L.t2p = torch.searchsorted(pkey, t_rows * ncn + L.t_col.long()).contiguous()
            # P^T: entries of P sorted by (column, row)
            perm = torch.argsort(L.p_col.long() * n + p_rows)
            L.pt_perm = perm.contiguous()
            L.pt_col = p_rows[perm].to(torch.int32).contiguous()
            L.pt_rowptr = _rowptr_from_rows(L.p_col.long()[perm], ncn)
            # A*P and P^T*(A*P)
            L.ap_rowptr, L.ap_col, L.ap_seg, L.ap_sl, L.ap_sr = _expand_terms(
                rows, col.long(), torch.arange(nnza, device=dev), L.p_rowptr, L.p_col, n, ncn)
            c_rowptr, c_col, L.ac_seg, L.ac_sl, L.ac_sr = _expand_terms(
                _rows_of(L.pt_rowptr), L.pt_col.long(), perm, L.ap_rowptr, L.ap_col, ncn, ncn)
            L.last = False
            rowptr, col, n = c_rowptr, c_col, ncn
            kind = torch.zeros(ncn, dtype=torch.uint8, device=dev)
            kind[1::2] = 1
            group = torch.arange(nc, dtype=torch.int64, device=dev).repeat_interleave(2)

    @property
    def sizes(self):
        return [L.n for L in self.levels]

    @property
    def operator_complexity(self) -> float:
        return sum(L.nnzb for L in self.levels) / self.levels[0].nnzb


class RBMHierarchy:
    """Numeric phase + a host-driven V(1,1) cycle over the kernels of ``csrc/amg_rbm.cu``.

    WORK IN PROGRESS: a *validation harness* for round 2 (one kernel launch per operation from Python, no fusion, no
    graph) -- it exists so that the new kernels can be proven end to end on a GPU (``tests/test_gpu_amg_rbm.py``)
    before the cycle is moved into ``cb_ksp_solve``.  ``A`` is a 3 x 3 block :class:`~cashocs_b200._utils.linalg.Matrix`
    on the mesh vertices.
    """

    def __init__(self, A, symbolic: RBMSymbolic | None = None, smooth_ratio: float = 10.0, power_its: int = 12) -> None:
        import ctypes as C

        from cashocs_b200 import _lib

        if A.bs != 3:
            raise ValueError("RBMHierarchy: 3 x 3 blocks (3-D vector problems) only")
        self._C, self._lib = C, _lib
        self.A = A
        self.device = A.val.device
        self.symbolic = symbolic if symbolic is not None else RBMSymbolic(A.rowptr, A.col, A.nrows)
        self.levels = self.symbolic.levels
        self.smooth_ratio, self.power_its = float(smooth_ratio), int(power_its)
        f64 = dict(dtype=torch.float64, device=self.device)
        self.data = []
        for li, L in enumerate(self.levels):
            D = _Level()
            D.val = None if li == 0 else torch.zeros(L.nnzb * 9, **f64)
            D.dinv = torch.zeros(L.n * 3, **f64)
            D.lmax = 2.0
            if not L.last:
                D.C = torch.zeros(L.nc * 3, **f64)
                D.S = torch.zeros(L.nc, **f64)
                D.t_val = torch.zeros(L.t_col.numel() * 9, **f64)
                D.p_val = torch.zeros(L.p_col.numel() * 9, **f64)
                D.pt_val = torch.zeros(L.p_col.numel() * 9, **f64)
                D.ap_val = torch.zeros(L.ap_col.numel() * 9, **f64)
            self.data.append(D)
        self.coarse_inv = None

    # ------------------------------------------------------------------ helpers
    def _mat(self, L, val):
        m = self._lib.cb_mat()
        m.nrows, m.ncols, m.bs = L.n, L.n, 3
        m.row_split = 0
        m.rowptr, m.col, m.val = L.rowptr.data_ptr(), L.col.data_ptr(), val.data_ptr()
        return m

    def _spmv(self, L, val, x):
        y = torch.empty(L.n * 3, dtype=torch.float64, device=self.device)
        m = self._mat(L, val)
        self._lib.check(self._lib.load().cb_spmv(self._lib.stream_ptr(), self._C.byref(m), self._lib.ptr(x), self._lib.ptr(y)),
                        "cb_spmv")
        return y

    def _diag_inv(self, L, val, dinv) -> float:
        """Point-Jacobi diagonal inverse (zero diagonals -> 1) and 1.15 x the power-method estimate of lmax(D^-1 A)."""
        rows = _rows_of(L.rowptr)
        diag_blocks = val.view(-1, 9)[rows == L.col.long()]          # ascending rows, every row holds its diagonal
        d = diag_blocks[:, [0, 4, 8]].reshape(-1)
        dinv.copy_(torch.where(d != 0.0, 1.0 / d, torch.ones_like(d)))
        gen = torch.Generator(device="cpu").manual_seed(1234)
        x = (torch.rand(L.n * 3, generator=gen, dtype=torch.float64) + 0.5).to(self.device)
        lam = 1.0
        for _ in range(self.power_its):
            y = self._spmv(L, val, x) * dinv
            nrm = torch.linalg.norm(y)
            lam = float(nrm / torch.linalg.norm(x))
            x = y / nrm
        return 1.15 * lam

    # ------------------------------------------------------------------ numeric phase
    def numeric(self, coords: torch.Tensor) -> None:
        """Values of every level for the current values of ``A`` and the current vertex positions ``coords`` [n, 3]."""
        lib, s, ptr, check = self._lib.load(), self._lib.stream_ptr(), self._lib.ptr, self._lib.check
        val = self.A.val
        centre = coords[: self.levels[0].n].contiguous().view(-1)
        scale = torch.zeros(self.levels[0].n, dtype=torch.float64, device=self.device)
        for li, (L, D) in enumerate(zip(self.levels, self.data)):
            D.cur_val = val
            D.lmax = self._diag_inv(L, val, D.dinv)
            if L.last:
                dense = torch.zeros(L.n * 3, L.n * 3, dtype=torch.float64, device=self.device)
                rows = _rows_of(L.rowptr)
                blk = val.view(-1, 3, 3)
                for r in range(3):
                    for c in range(3):
                        dense[rows * 3 + r, L.col.long() * 3 + c] = blk[:, r, c]
                self.coarse_inv = torch.linalg.pinv(dense)  # rotation nodes of one-node aggregates carry no energy
                break
            check(lib.cb_amg_rbm_aggregate_geometry(s, L.nc, ptr(L.member_ptr), ptr(L.member_idx), ptr(L.kind), ptr(centre),
                                                    ptr(scale), ptr(D.C), ptr(D.S)), "cb_amg_rbm_aggregate_geometry")
            check(lib.cb_amg_rbm_tentative(s, L.n, ptr(L.t_rowptr), ptr(L.agg), ptr(L.kind), ptr(centre), ptr(scale), ptr(D.C),
                                           ptr(D.S), ptr(D.t_val)), "cb_amg_rbm_tentative")
            check(lib.cb_amg_rbm_spgemm_terms(s, 0, L.p_col.numel(), ptr(L.at_seg), ptr(L.at_sl), ptr(L.at_sr), ptr(val),
                                              ptr(D.t_val), ptr(D.p_val)), "cb_amg_rbm_spgemm_terms(A*T)")
            omega = (4.0 / 3.0) / D.lmax
            check(lib.cb_amg_rbm_prolongator_finish(s, L.n, ptr(L.p_rowptr), L.t_col.numel(), ptr(L.t2p), ptr(D.t_val),
                                                    ptr(D.dinv), omega, ptr(D.p_val)), "cb_amg_rbm_prolongator_finish")
            check(lib.cb_amg_rbm_transpose_blocks(s, L.p_col.numel(), ptr(L.pt_perm), ptr(D.p_val), ptr(D.pt_val)),
                  "cb_amg_rbm_transpose_blocks")
            check(lib.cb_amg_rbm_spgemm_terms(s, 0, L.ap_col.numel(), ptr(L.ap_seg), ptr(L.ap_sl), ptr(L.ap_sr), ptr(val),
                                              ptr(D.p_val), ptr(D.ap_val)), "cb_amg_rbm_spgemm_terms(A*P)")
            nxt = self.data[li + 1]
            check(lib.cb_amg_rbm_spgemm_terms(s, 1, nxt.val.numel() // 9, ptr(L.ac_seg), ptr(L.ac_sl), ptr(L.ac_sr),
                                              ptr(D.p_val), ptr(D.ap_val), ptr(nxt.val)), "cb_amg_rbm_spgemm_terms(Pt*AP)")
            val = nxt.val
            # geometry of the coarse nodes: both nodes of aggregate J sit at C_J; the rotation node carries S_J
            centre = D.C.view(-1, 3).repeat_interleave(2, dim=0).contiguous().view(-1)
            scale = torch.zeros(2 * L.nc, dtype=torch.float64, device=self.device)
            scale[1::2] = D.S

    # ------------------------------------------------------------------ cycle (host driven)
    def _smooth(self, L, D, b, x):
        theta = 0.5 * (D.lmax + D.lmax / self.smooth_ratio)
        if x is None:
            return D.dinv * b / theta
        return x + D.dinv * (b - self._spmv(L, D.cur_val, x)) / theta

    def _cycle(self, li: int, b: torch.Tensor) -> torch.Tensor:
        lib, s, ptr, check = self._lib.load(), self._lib.stream_ptr(), self._lib.ptr, self._lib.check
        L, D = self.levels[li], self.data[li]
        if L.last:
            return self.coarse_inv @ b
        x = self._smooth(L, D, b, None)
        t = self._spmv(L, D.cur_val, x)
        bc = torch.empty(2 * L.nc * 3, dtype=torch.float64, device=self.device)
        check(lib.cb_amg_rbm_restrict(s, 2 * L.nc, ptr(L.pt_rowptr), ptr(L.pt_col), ptr(D.pt_val), ptr(b), ptr(t), ptr(bc)),
              "cb_amg_rbm_restrict")
        xc = self._cycle(li + 1, bc)
        check(lib.cb_amg_rbm_prolong(s, L.n, ptr(L.p_rowptr), ptr(L.p_col), ptr(D.p_val), ptr(xc.contiguous()), ptr(x)),
              "cb_amg_rbm_prolong")
        return self._smooth(L, D, b, x)

    def apply(self, r: torch.Tensor) -> torch.Tensor:
        """``z = M r`` with one V(1,1) cycle."""
        return self._cycle(0, r)

    def pcg(self, b: torch.Tensor, rtol: float = 1e-10, max_it: int = 500):
        """Host-driven PCG on the preconditioned norm (validation only; the product loop is ``cb_ksp_solve``)."""
        L0 = self.levels[0]
        x = torch.zeros_like(b)
        r = b.clone()
        z = self.apply(r)
        p = z.clone()
        rz = float(r @ z)
        z0 = float(torch.linalg.norm(z))
        for it in range(1, max_it + 1):
            w = self._spmv(L0, self.A.val, p)
            alpha = rz / float(p @ w)
            x.add_(p, alpha=alpha)
            r.add_(w, alpha=-alpha)
            z = self.apply(r)
            if float(torch.linalg.norm(z)) <= rtol * z0:
                return x, it
            rz_new = float(r @ z)
            p.mul_(rz_new / rz).add_(z)
            rz = rz_new
        return x, max_it

    # ------------------------------------------------------------------ the C struct for the fused V-cycle
    def struct(self):
        """``cb_amg`` with ``p_block = 1`` for ``cb_ksp_opts.amg``: the fused V-cycle of ``cb_ksp_solve`` then restricts
        and prolongates with the block kernels and runs everything else (smoothers, A*P post-smoothing, dense
        coarsest solve) unchanged."""
        from cashocs_b200._utils import linalg

        C, _lib = self._C, self._lib
        f64 = dict(dtype=torch.float64, device=self.device)
        arr = (_lib.cb_amg_level * len(self.levels))()
        self._keep = []
        for li, (L, D) in enumerate(zip(self.levels, self.data)):
            e = arr[li]
            m = self._mat(L, D.cur_val)
            m.row_split = linalg.row_split_hint(L.nnzb, L.n)
            e.A = m
            e.dinv = D.dinv.data_ptr()
            e.lmax = D.lmax
            e.nc = 0 if L.last else 2 * L.nc
            bufs = [torch.zeros(L.n * 3, **f64) for _ in range(4)]  # x, b, t, d
            self._keep.append(bufs)
            e.x = bufs[0].data_ptr() if li > 0 else None
            e.b = bufs[1].data_ptr() if li > 0 else None
            e.t, e.d = bufs[2].data_ptr(), bufs[3].data_ptr()
            if not L.last:
                e.p_rowptr, e.p_col, e.p_val = L.p_rowptr.data_ptr(), L.p_col.data_ptr(), D.p_val.data_ptr()
                e.pt_rowptr, e.pt_col, e.pt_val = L.pt_rowptr.data_ptr(), L.pt_col.data_ptr(), D.pt_val.data_ptr()
                ap = _lib.cb_mat()
                ap.nrows, ap.ncols, ap.bs = L.n, 2 * L.nc, 3
                ap.row_split = linalg.row_split_hint(L.ap_col.numel(), L.n)
                ap.rowptr, ap.col, ap.val = L.ap_rowptr.data_ptr(), L.ap_col.data_ptr(), D.ap_val.data_ptr()
                e.AP = ap
        h = _lib.cb_amg()
        h.nlevels = len(self.levels)
        h.smooth_degree, h.coarse_degree = 1, 16
        h.smooth_ratio, h.coarse_ratio = self.smooth_ratio, 40.0
        h.p_block = 1
        h.levels = C.cast(arr, C.c_void_p)
        self._coarse_inv_flat = self.coarse_inv.contiguous().view(-1) if self.coarse_inv is not None else None
        h.coarse_inv = self._coarse_inv_flat.data_ptr() if self._coarse_inv_flat is not None else None
        self._level_array, self._struct = arr, h
        return h
