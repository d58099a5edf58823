"""Cell measures needed by the ``inhomogeneous`` elasticity scaling (``shape_form_handler.py:629-644``)."""

from __future__ import annotations

import torch

from cashocs_b200 import _lib


def cell_volumes(mesh, with_max: bool = False):
    """|K| per cell (``cb_cell_volumes``: one pass over the cells, the maximum fused into the same kernel).  This is
    the DG0 ``l2_projection`` of ``CellVolume`` (``cashocs/_utils/linalg.py:768-809``), which is the identity for a
    cellwise constant.  ``with_max``: also return the (rank-local) maximum as a one-element device tensor."""
    _lib.require_cuda(mesh.coords)
    vol = torch.empty(mesh.num_cells, dtype=torch.float64, device=mesh.device)
    vmax = torch.empty(1, dtype=torch.float64, device=mesh.device)
    scratch = mesh.reduce_scratch()
    _lib.check(_lib.load().cb_cell_volumes(_lib.stream_ptr(), mesh.dim, mesh.num_cells, _lib.ptr(mesh.coords),
                                           _lib.ptr(mesh.cells), _lib.ptr(vol), _lib.ptr(vmax), _lib.ptr(scratch),
                                           scratch.numel()), "cb_cell_volumes")
    return (vol, vmax) if with_max else vol
