"""Mesh deformation (``cashocs/geometry/deformations.py``): ALE move / revert on the device."""

from __future__ import annotations

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib


class DeformationHandler:
    """Moves the mesh by ``id + V`` and reverts (``deformations.py:38-205``).

    A deformation is a vertex-blocked device vector ``[Nv*d]`` (dof ``d*v + c``), which is both
    the "dof" and the "coordinate" layout, so ``dof_to_coordinate`` / ``coordinate_to_dof``
    (``:146-185``) are reshapes.
    """

    def __init__(self, mesh, a_priori_tester, intersection_tester) -> None:
        self.mesh = mesh
        self.a_priori_tester = a_priori_tester
        self.intersection_tester = intersection_tester
        self.dim = mesh.dim
        self.mesh.coords_old = torch.empty_like(mesh.coords)
        self._has_old = False

    def coordinate_to_dof(self, coordinate_deformation) -> torch.Tensor:
        t = coordinate_deformation
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float64)).to(self.mesh.device)
        return t.reshape(-1).contiguous()

    def dof_to_coordinate(self, dof_deformation: torch.Tensor) -> torch.Tensor:
        return dof_deformation.view(self.mesh.num_vertices, self.dim)

    def revert_transformation(self) -> None:
        """``mesh.coordinates()[:, :] = old_coordinates`` (``deformations.py:75-84``)."""
        if not self._has_old:
            raise _exceptions.CashocsException("no transformation to revert")
        self.mesh.coords.copy_(self.mesh.coords_old)
        self._has_old = False

    def move_mesh(self, transformation, validated_a_priori: bool = False,
                  test_for_intersections: bool = True, step: float = 1.0) -> bool:
        """Perturbation of identity ``x -> x + step * V(x)`` (``deformations.py:86-144``)."""
        mesh = self.mesh
        if isinstance(transformation, np.ndarray):
            if tuple(transformation.shape) != tuple(mesh.coords.shape):
                raise _exceptions.CashocsException("Not a valid dimension for the transformation")
        V = self.coordinate_to_dof(transformation)
        if V.numel() != mesh.coords.numel():
            raise _exceptions.CashocsException("Not a valid dimension for the transformation")
        if not validated_a_priori:
            if not self.a_priori_tester.test(V, float("inf"), step):
                return False
        _lib.check(
            _lib.load().cb_move_mesh(_lib.stream_ptr(), mesh.coords.numel(), _lib.ptr(mesh.coords),
                                     _lib.ptr(mesh.coords_old), _lib.ptr(V), float(step)),
            "cb_move_mesh",
        )
        self._has_old = True
        if test_for_intersections:
            check = self.intersection_tester.test()
            if not check:
                self.revert_transformation()
            return check
        return True

    def assign_coordinates(self, coordinates) -> bool:
        """Assign new coordinates, reverting when the mesh self-intersects (``:187-205``)."""
        new = self.coordinate_to_dof(coordinates).view_as(self.mesh.coords)
        self.mesh.coords_old.copy_(self.mesh.coords)
        self.mesh.coords.copy_(new)
        self._has_old = True
        check = self.intersection_tester.test()
        if not check:
            self.revert_transformation()
        return check
