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
