"""Cell measures needed by the ``inhomogeneous`` elasticity scaling (``shape_form_handler.py:629-644``)."""

from __future__ import annotations

import math

import torch


def cell_volumes(mesh) -> torch.Tensor:
    """|K| per cell.  Off the hot path by default (``inhomogeneous = False``), torch plumbing."""
    x = mesh.coords[mesh.cells.long()]
    J = (x[:, 1:, :] - x[:, :1, :]).transpose(1, 2)
    return torch.linalg.det(J).abs() / math.factorial(mesh.dim)
