"""Mesh quality on the device (``cashocs/geometry/quality.py``).

``compute_mesh_quality`` keeps the reference signature (``quality.py:41-95``); the four
calculators keep their names and ``compute(mesh) -> per-cell array`` contract
(``quality.py:273-461``).  One fused kernel computes the per-cell measure together with the
min / sum reductions; ``min`` / ``avg`` never materialise the per-cell array.  The reference
gathers every cell value on rank 0 (``quality.py:321``) -- here ranks all-reduce two doubles.
"""

from __future__ import annotations

import abc

import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib

_MEASURE_ID = {"skewness": 0, "maximum_angle": 1, "radius_ratios": 2, "condition_number": 3}


def _launch(mesh, measure: str, want_cells: bool):
    lib = _lib.load()
    _lib.require_cuda(mesh.coords, mesh.cells)
    nc = mesh.n_owned_cells
    q = torch.empty(nc, dtype=torch.float64, device=mesh.device) if want_cells else None
    red = torch.empty(2, dtype=torch.float64, device=mesh.device)
    scratch = mesh.reduce_scratch()
    _lib.check(
        lib.cb_mesh_quality(_lib.stream_ptr(), mesh.dim, nc, _lib.ptr(mesh.coords), _lib.ptr(mesh.cells),
                            _MEASURE_ID[measure], _lib.ptr(q), _lib.ptr(red), _lib.ptr(scratch), scratch.numel()),
        "cb_mesh_quality",
    )
    return q, red


class MeshQualityCalculator(abc.ABC):
    """Interface for calculating mesh quality (``quality.py:98-284``)."""

    measure: str = ""

    def compute(self, mesh) -> torch.Tensor:
        """Per-cell quality of the owned cells (device tensor)."""
        return _launch(mesh, self.measure, True)[0]

    def reductions(self, mesh) -> tuple[float, float, int]:
        """(min, sum, number of cells) over the whole (distributed) mesh."""
        _, red = _launch(mesh, self.measure, False)
        n = mesh.n_owned_cells
        if mesh.dist is not None:
            mn = red[0:1].clone()
            sm = torch.stack([red[1], torch.tensor(float(n), dtype=torch.float64, device=red.device)])
            mesh.dist.allreduce(mn, "min")
            mesh.dist.allreduce(sm, "sum")
            return float(mn.item()), float(sm[0].item()), int(round(float(sm[1].item())))
        r = red.cpu()
        return float(r[0]), float(r[1]), n


class SkewnessCalculator(MeshQualityCalculator):
    """Skewness (``quality.py:287-327``, C++ ``:181-221``)."""

    measure = "skewness"


class MaximumAngleCalculator(MeshQualityCalculator):
    """Maximum angle (``quality.py:330-368``, C++ ``:225-262``)."""

    measure = "maximum_angle"


class RadiusRatiosCalculator(MeshQualityCalculator):
    """Radius ratios d*r/R (``quality.py:371-400``)."""

    measure = "radius_ratios"


class ConditionNumberCalculator(MeshQualityCalculator):
    """sqrt(d) / (|J|_F |J^-1|_F) (``quality.py:403-461``)."""

    measure = "condition_number"


class MeshQuality:
    """Reductions of the per-cell quality (``quality.py:464-546``)."""

    @classmethod
    def min(cls, calculator: MeshQualityCalculator, mesh) -> float:
        return calculator.reductions(mesh)[0]

    @classmethod
    def avg(cls, calculator: MeshQualityCalculator, mesh) -> float:
        _, s, n = calculator.reductions(mesh)
        return s / n

    @classmethod
    def quantile(cls, calculator: MeshQualityCalculator, mesh, quantile: float) -> float:
        q = calculator.compute(mesh)
        if mesh.dist is not None:
            q = mesh.dist.gather_cells(q)
        # np.quantile (linear interpolation) semantics; rarely used (default type is "min")
        return float(torch.quantile(q, float(quantile), interpolation="linear").item())


_CALCULATORS = {
    "skewness": SkewnessCalculator,
    "maximum_angle": MaximumAngleCalculator,
    "radius_ratios": RadiusRatiosCalculator,
    "condition_number": ConditionNumberCalculator,
}


def compute_mesh_quality(mesh, quality_type: str = "min", quality_measure: str = "skewness",
                         quantile: float = 0.0) -> float:
    """This computes the mesh quality of a given mesh (``cashocs/geometry/quality.py:41-95``)."""
    if quality_measure not in _CALCULATORS:
        raise _exceptions.InputError("cashocs_b200.compute_mesh_quality", "quality_measure",
                                     f"unknown measure {quality_measure!r}")
    calculator = _CALCULATORS[quality_measure]()
    if quality_type in ("min", "minimum"):
        return MeshQuality.min(calculator, mesh)
    if quality_type in ("avg", "average"):
        return MeshQuality.avg(calculator, mesh)
    if quality_type == "quantile":
        return MeshQuality.quantile(calculator, mesh, quantile)
    raise _exceptions.InputError("cashocs_b200.compute_mesh_quality", "quality_type",
                                 f"unknown type {quality_type!r}")
