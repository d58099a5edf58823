"""Mesh handling for shape optimisation (``cashocs/geometry/mesh_handler.py``)."""

from __future__ import annotations

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200.geometry import deformations
from cashocs_b200.geometry import quality


def check_mesh_quality_tolerance(mesh_quality: float, tolerance: float) -> None:
    """Raise when the mesh quality is below ``tol_upper`` (``mesh_handler.py:54-74``)."""
    if mesh_quality < tolerance:
        fail_msg = (
            "The quality of the mesh file you have specified is not "
            "sufficient for computing the shape gradient.\n "
            + f"It currently is {mesh_quality:.3e} but has to "
            f"be at least {tolerance:.3e}."
        )
        raise _exceptions.InputError("cashocs.geometry.import_mesh", "input_arg", fail_msg)


class _MeshHandler:
    """Quality thresholds + move / revert orchestration (``mesh_handler.py:77-377``)."""

    def __init__(self, mesh, config, a_priori_tester, a_posteriori_tester, is_remeshed: bool = False) -> None:
        self.mesh = mesh
        self.config = config
        self.a_priori_tester = a_priori_tester
        self.a_posteriori_tester = a_posteriori_tester
        self.deformation_handler = deformations.DeformationHandler(mesh, a_priori_tester, a_posteriori_tester)
        self.volume_change = float(config.get("MeshQuality", "volume_change"))
        self.angle_change = float(config.get("MeshQuality", "angle_change"))
        self.test_for_intersections = config.getboolean("ShapeGradient", "test_for_intersections")
        self.mesh_quality_tol_lower = config.getfloat("MeshQuality", "tol_lower")
        self.mesh_quality_tol_upper = config.getfloat("MeshQuality", "tol_upper")
        self.mesh_quality_measure = config.get("MeshQuality", "measure")
        self.mesh_quality_type = config.get("MeshQuality", "type")
        self.quality_quantile = config.getfloat("MeshQuality", "quantile")
        self.current_mesh_quality = quality.compute_mesh_quality(
            mesh, self.mesh_quality_type, self.mesh_quality_measure, self.quality_quantile
        )
        if not is_remeshed:
            check_mesh_quality_tolerance(self.current_mesh_quality, self.mesh_quality_tol_upper)

    def move_mesh(self, transformation: torch.Tensor, step: float = 1.0) -> bool:
        """a-priori test -> move -> (intersection test) -> quality (``mesh_handler.py:264-303``)."""
        if transformation.numel() != self.mesh.coords.numel():
            raise _exceptions.CashocsException("Not a valid mesh transformation")
        if not self.a_priori_tester.test(transformation, self.volume_change, step):
            return False
        success_flag = self.deformation_handler.move_mesh(
            transformation, validated_a_priori=True, test_for_intersections=self.test_for_intersections, step=step
        )
        self.current_mesh_quality = quality.compute_mesh_quality(
            self.mesh, self.mesh_quality_type, self.mesh_quality_measure, self.quality_quantile
        )
        return success_flag

    def revert_transformation(self) -> None:
        self.deformation_handler.revert_transformation()

    def compute_decreases(self, search_direction: torch.Tensor, stepsize: float) -> int:
        """Number of a-priori Armijo halvings for ``angle_change`` (``mesh_handler.py:329-377``)."""
        if self.angle_change == float("inf"):
            return 0
        mesh = self.mesh
        red = torch.empty(2, dtype=torch.float64, device=mesh.device)
        scratch = mesh.reduce_scratch()
        _lib.check(
            _lib.load().cb_gradient_frobenius_max(_lib.stream_ptr(), mesh.dim, mesh.num_cells,
                                                  _lib.ptr(mesh.coords), _lib.ptr(mesh.cells),
                                                  _lib.ptr(search_direction), _lib.ptr(red),
                                                  _lib.ptr(scratch), scratch.numel()),
            "cb_gradient_frobenius_max",
        )
        if mesh.dist is not None:
            mesh.dist.allreduce(red[1:2], "max")
        frobenius_norm = float(red[1].item())
        beta_armijo = self.config.getfloat("LineSearch", "beta_armijo")
        return int(
            np.maximum(
                np.ceil(np.log(self.angle_change / stepsize / frobenius_norm) / np.log(1 / beta_armijo)),
                0.0,
            )
        )
