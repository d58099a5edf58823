"""Geometric regularisation of shape problems (``cashocs/_forms/shape_regularization.py``): volume (``:131-222``),
surface (``:224-316``) and barycenter (``:318-510``).  Curvature regularisation (a surface Laplace-Beltrami solve)
stays outside the hot path (``Config._check_supported`` refuses it).

The cell-integral terms need no new kernels: everything is expressed through two existing device entry points.

* ``int_Omega w dx`` for a P1 field ``w`` is ``cb_functional`` -- volume with ``w = 1``, first moments with the
  coordinate fields.
* For a polynomial ``f`` the shape derivative of ``int_Omega f dx`` in direction ``V = phi_a e_i`` is
  ``int (f div V + grad f . V) dx``, which is what ``cb_assemble_shape_derivative_poisson`` computes from its source
  terms ``-G_a[i] int_K f p - int_K (d_i f) phi_a p`` when the "adjoint" is the constant ``-1`` and the "state" is
  zero.  Both regularisations are linear combinations of ``d(vol)`` (``f = 1``) and ``d(int x_k)`` (``f = x_k``), so
  the whole regularisation derivative is ONE launch with the affine polynomial ``f = a_0 + sum_k a_k x_k``.

The surface term integrates over the boundary facets (``csrc/boundary.cu``): ``cb_boundary_measure`` and
``cb_boundary_measure_gradient`` -- for P1 deformations ``int_F t_div(V, n) ds = d|F|[V]`` exactly, so the derivative
vector is the gradient of the facet measures, gathered per vertex in a fixed order.
"""

from __future__ import annotations

import torch

from cashocs_b200 import _lib
from cashocs_b200._forms import expressions
from cashocs_b200._utils import linalg


class ShapeRegularization:
    def __init__(self, mesh, config, bc_flag) -> None:
        self.mesh, self.config, self.bc_flag = mesh, config, bc_flag
        d = mesh.dim
        self.mu_volume = config.getfloat("Regularization", "factor_volume")
        self.mu_barycenter = config.getfloat("Regularization", "factor_barycenter")
        self.mu_surface = config.getfloat("Regularization", "factor_surface")
        self.is_active = self.mu_volume > 0.0 or self.mu_barycenter > 0.0 or self.mu_surface > 0.0
        if not self.is_active:
            return
        f64 = dict(dtype=torch.float64, device=mesh.device)
        self._ones = torch.ones(mesh.num_vertices, **f64)
        self._minus_ones = -self._ones
        self._zeros = torch.zeros(mesh.num_vertices, **f64)
        self._field = torch.zeros(mesh.num_vertices, **f64)
        self._out = torch.zeros(1, **f64)
        self._vector = torch.zeros(mesh.n_owned_vertices * d, **f64)
        self.target_volume = config.getfloat("Regularization", "target_volume")
        if config.getboolean("Regularization", "use_initial_volume"):
            self.target_volume = self.compute_volume()
        target = [float(t) for t in config.getlist("Regularization", "target_barycenter")]
        self.target_barycenter = (target + [0.0] * 3)[:3]
        if config.getboolean("Regularization", "use_initial_barycenter"):
            self.target_barycenter = self.compute_barycenter()
        self.current_volume = 1.0
        self.current_barycenter = [0.0, 0.0, 0.0]
        self.current_surface = 1.0
        self.target_surface = config.getfloat("Regularization", "target_surface")
        if self.mu_surface > 0.0:
            self._facets, self._v2f_ptr, self._v2f_idx, counted = mesh.boundary_facet_table()
            # facets this rank adds to the global measure (all of them on one GPU; on a partition every boundary
            # facet is counted by exactly one rank)
            self._measure_facets = self._facets if bool(counted.all()) else self._facets[counted].contiguous()
            nf = int(self._facets.shape[0])
            self._facet_vec = torch.zeros(max(nf * d * d, 1), **f64)
            self._surface_vector = torch.zeros(mesh.n_owned_vertices * d, **f64)
            if config.getboolean("Regularization", "use_initial_surface"):
                self.target_surface = self.compute_surface()

    # -- geometric quantities
    def _integral(self, w: torch.Tensor) -> float:
        mesh = self.mesh
        scratch = mesh.reduce_scratch()
        _lib.check(
            _lib.load().cb_functional(_lib.stream_ptr(), mesh.dim, mesh.n_owned_cells, _lib.ptr(mesh.coords),
                                      _lib.ptr(mesh.cells), 1.0, _lib.ptr(w), 0.0, None, _lib.ptr(self._out),
                                      _lib.ptr(scratch), scratch.numel()),
            "cb_functional",
        )
        if mesh.dist is not None:
            mesh.dist.allreduce(self._out, "sum")
        return float(self._out.item())

    def compute_volume(self) -> float:
        """``_compute_volume`` (``:213-222``)."""
        return self._integral(self._ones)

    def compute_barycenter(self) -> list[float]:
        """``_compute_barycenter_list`` (``:486-510``); z = 0 in two dimensions."""
        volume = self.compute_volume()
        out = [0.0, 0.0, 0.0]
        for k in range(self.mesh.dim):
            self._field.copy_(self.mesh.coords[:, k])
            out[k] = self._integral(self._field) / volume
        return out

    def compute_surface(self) -> float:
        """``_compute_surface`` (``:307-316``)."""
        mesh = self.mesh
        scratch = mesh.reduce_scratch()
        _lib.check(
            _lib.load().cb_boundary_measure(_lib.stream_ptr(), mesh.dim, int(self._measure_facets.shape[0]),
                                            _lib.ptr(mesh.coords), _lib.ptr(self._measure_facets), _lib.ptr(self._out),
                                            _lib.ptr(scratch), scratch.numel()),
            "cb_boundary_measure",
        )
        if mesh.dist is not None:
            mesh.dist.allreduce(self._out, "sum")
        return float(self._out.item())

    def update_geometric_quantities(self) -> None:
        """``ShapeRegularization.update_geometric_quantities`` (``:670-673``)."""
        if self.is_active:
            self.current_volume = self.compute_volume()
            if self.mu_barycenter > 0.0:
                self.current_barycenter = self.compute_barycenter()
            if self.mu_surface > 0.0:
                self.current_surface = self.compute_surface()

    def compute_objective(self) -> float:
        """``:180-192`` + ``:437-459``."""
        value = 0.0
        if not self.is_active:
            return value
        if self.mu_volume > 0.0:
            value += 0.5 * self.mu_volume * (self.compute_volume() - self.target_volume) ** 2
        if self.mu_surface > 0.0:
            value += 0.5 * self.mu_surface * (self.compute_surface() - self.target_surface) ** 2
        if self.mu_barycenter > 0.0:
            bc = self.compute_barycenter()
            value += 0.5 * self.mu_barycenter * sum((bc[k] - self.target_barycenter[k]) ** 2 for k in range(3))
        return value

    def add_shape_derivative(self, b: torch.Tensor) -> torch.Tensor:
        """``b += dJ_reg[V]`` with the forms of ``:154-174`` and ``:353-425`` (the barycenter form verbatim, i.e. with
        the reference's ``+ bc_k / vol * div V``); call after :meth:`update_geometric_quantities`."""
        if not self.is_active:
            return b
        d = self.mesh.dim
        if self.mu_surface > 0.0:  # mu (surface - target) int_Gamma t_div(V, n) ds  (:248-269)
            mesh = self.mesh
            _lib.check(
                _lib.load().cb_boundary_measure_gradient(
                    _lib.stream_ptr(), d, int(self._facets.shape[0]), mesh.n_owned_vertices, _lib.ptr(mesh.coords),
                    _lib.ptr(self._facets), _lib.ptr(self._v2f_ptr), _lib.ptr(self._v2f_idx), _lib.ptr(self._facet_vec),
                    _lib.ptr(self._surface_vector)),
                "cb_boundary_measure_gradient",
            )
            if self.bc_flag is not None:
                self._surface_vector.masked_fill_(self.bc_flag[: self._surface_vector.numel()].bool(), 0.0)
            b.add_(self._surface_vector, alpha=self.mu_surface * (self.current_surface - self.target_surface))
        if self.mu_volume <= 0.0 and self.mu_barycenter <= 0.0:
            return b
        vol, bc, target = self.current_volume, self.current_barycenter, self.target_barycenter
        a0 = self.mu_volume * (vol - self.target_volume) if self.mu_volume > 0.0 else 0.0
        f = expressions.Polynomial.constant(0.0)
        if self.mu_barycenter > 0.0:
            for k in range(d):
                w = self.mu_barycenter * (bc[k] - target[k])
                a0 += w * bc[k] / vol
                f = f + (w / vol) * expressions.Polynomial.coordinate(k)
        f = f + a0
        linalg.assemble_shape_derivative(self.mesh, self._zeros, self._minus_ones, f, 0.0, bc_flag=self.bc_flag,
                                         b=self._vector)
        b.add_(self._vector)
        return b
