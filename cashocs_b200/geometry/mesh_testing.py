"""Mesh testers (``cashocs/geometry/mesh_testing.py``), fused device kernels.

The reference assembles and "solves" a DG0 mass system with preonly+jacobi to obtain
``det(I + grad V)`` per cell (``mesh_testing.py:61-81,103-110``); for a P1 field the
determinant is constant per cell, so one kernel evaluates it and reduces min / max exactly.
"""

from __future__ import annotations

import torch

from cashocs_b200 import _lib
from cashocs_b200 import log


class APrioriMeshTester:
    """Tests the mesh before it is modified (``mesh_testing.py:40-132``)."""

    def __init__(self, mesh) -> None:
        self.mesh = mesh
        self.last_min_det = float("nan")
        self.last_max_det = float("nan")

    def determinants(self, transformation: torch.Tensor, step: float = 1.0):
        """Return (min, max) of det(I + step*grad(transformation)) over all cells."""
        mesh = self.mesh
        _lib.require_cuda(transformation)
        red = torch.empty(2, dtype=torch.float64, device=mesh.device)
        scratch = mesh.reduce_scratch()
        _lib.check(
            _lib.load().cb_det_check(_lib.stream_ptr(), mesh.dim, mesh.num_cells, _lib.ptr(mesh.coords),
                                     _lib.ptr(mesh.cells), _lib.ptr(transformation), float(step), None,
                                     _lib.ptr(red), _lib.ptr(scratch), scratch.numel()),
            "cb_det_check",
        )
        if mesh.dist is not None:
            mesh.dist.allreduce(red[0:1], "min")
            mesh.dist.allreduce(red[1:2], "max")
        r = red.cpu()
        return float(r[0]), float(r[1])

    @log.profile_execution_time("a priori mesh testing")
    def test(self, transformation: torch.Tensor, volume_change: float, step: float = 1.0) -> bool:
        """``min_det >= 1/volume_change and max_det <= volume_change`` (``mesh_testing.py:84-132``)."""
        min_det, max_det = self.determinants(transformation, step)
        self.last_min_det, self.last_max_det = min_det, max_det
        return bool((min_det >= 1 / volume_change) and (max_det <= volume_change))


class IntersectionTester:
    """Tests the mesh after it has been modified (``mesh_testing.py:135-195``).

    Per vertex the number of cells whose closed simplex contains it must equal the vertex'
    topological valence (``:154-158,177-179``).  The reference queries dolfin's bounding-box
    tree per vertex (``compute_collisions`` ``:201-241``); the device version bins vertices in a
    uniform grid and lets every cell test the vertices inside its bounding box.
    """

    def __init__(self, mesh) -> None:
        self.mesh = mesh
        mesh.build_pattern()
        self.occurrences = mesh.valence
        self.counts = torch.empty(mesh.num_vertices, dtype=torch.int32, device=mesh.device)

    def compute_collisions(self) -> torch.Tensor:
        self._run()
        return self.counts

    def _run(self) -> int:
        mesh = self.mesh
        mism = torch.zeros(1, dtype=torch.int64, device=mesh.device)
        _lib.check(
            _lib.load().cb_intersection_test(_lib.stream_ptr(), mesh.dim, mesh.num_vertices, mesh.num_cells,
                                             _lib.ptr(mesh.coords), _lib.ptr(mesh.cells),
                                             _lib.ptr(self.occurrences), _lib.ptr(self.counts), _lib.ptr(mism)),
            "cb_intersection_test",
        )
        return int(mism.item())

    @log.profile_execution_time("intersection testing")
    def test(self) -> bool:
        bad = self._run() > 0
        if self.mesh.dist is not None:
            flag = torch.tensor([1.0 if bad else 0.0], dtype=torch.float64, device=self.mesh.device)
            self.mesh.dist.allreduce(flag, "max")
            bad = bool(flag.item() > 0)
        return not bad
