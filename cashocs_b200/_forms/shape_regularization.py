"""Geometric regularisation of shape problems (``cashocs/_forms/shape_regularization.py``): volume (``:131-222``),
surface (``:224-316``), barycenter (``:318-510``) and curvature (``:512-646``).

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

The curvature term ``mu/2 int_Gamma |kappa|^2 ds`` needs the mean-curvature vector ``kappa``:
``int kappa . W ds = int t_grad(x, n) : t_grad(W, n) ds`` for all vector-P1 ``W`` (``_compute_curvature``, ``:596-613``).
The right-hand side equals ``int t_div(W, n) ds`` (``t_grad(x, n)`` is the tangential projector), i.e. the facet-measure
gradient again; the matrix is the P1 mass matrix of the boundary, the same for every component, which lives on the
boundary vertices only (the reference assembles it on all vertices and ``ident_zeros`` the interior rows, where
``kappa`` is therefore zero): ``cb_boundary_mass`` on the compressed boundary numbering, d solves with the device CG
(the reference's default direct solver is replaced like everywhere else, ``_utils/linalg.py``), then one facet
kernel for the derivative (``cb_boundary_curvature_derivative``).
"""

from __future__ import annotations

import torch

from cashocs_b200 import _lib
from cashocs_b200._forms import expressions
from cashocs_b200._utils import linalg


class ShapeRegularization:
    def __init__(self, mesh, config, bc_flag, cost=None) -> None:
        self.mesh, self.config, self.bc_flag = mesh, config, bc_flag
        d = mesh.dim
        # geometric terms of the cost functional itself (IntegralFunctional.c_volume / .c_surface): linear in the
        # volume / the boundary measure, they ride on the same kernels as the quadratic regularisation terms
        # (read at use: the effective coefficients of a CostFunctionalList follow the state)
        self._cost = cost
        self._j_volume_possible = bool(getattr(cost, "has_volume_terms", getattr(cost, "c_volume", 0.0) != 0.0))
        self._j_surface_possible = bool(getattr(cost, "has_surface_terms", getattr(cost, "c_surface", 0.0) != 0.0))
        self.mu_volume = config.getfloat("Regularization", "factor_volume")
        self.mu_barycenter = config.getfloat("Regularization", "factor_barycenter")
        self.mu_surface = config.getfloat("Regularization", "factor_surface")
        self.mu_curvature = config.getfloat("Regularization", "factor_curvature")
        self.is_active = (self.mu_volume > 0.0 or self.mu_barycenter > 0.0 or self.mu_surface > 0.0
                          or self.mu_curvature > 0.0 or self._j_volume_possible or self._j_surface_possible)
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
        if self.mu_surface > 0.0 or self.mu_curvature > 0.0 or self._j_surface_possible:
            self._facets, self._v2f_ptr, self._v2f_idx, counted = mesh.boundary_facet_table()
            # facets this rank adds to the global measure (all of them on one GPU; on a partition every boundary
            # facet is counted by exactly one rank)
            self._measure_facets = self._facets if bool(counted.all()) else self._facets[counted].contiguous()
            nf = int(self._facets.shape[0])
            self._facet_vec = torch.zeros(max(nf * d * d, 1), **f64)
            self._surface_vector = torch.zeros(mesh.n_owned_vertices * d, **f64)
            if self.mu_surface > 0.0 and config.getboolean("Regularization", "use_initial_surface"):
                self.target_surface = self.compute_surface()
        if self.mu_curvature > 0.0:
            self._setup_curvature()
        if config.getboolean("Regularization", "use_relative_scaling"):
            self.scale()

    def scale(self) -> None:
        """``scale`` of the four terms (``:194-211,289-305,461-484,621-643``): with ``use_relative_scaling`` every active
        term's factor is divided by the term's value on the initial geometry, unless that vanishes."""
        from cashocs_b200 import log

        def scaled(name: str, mu: float, value: float) -> float:
            if mu <= 0.0:
                return mu
            if abs(value) < 1e-15:
                log.info(f"The {name} regularization vanishes for the initial iteration. Multiplying this term with "
                         "the factor you supplied as weight.")
                return mu
            return mu / abs(value)

        if self.mu_volume > 0.0:
            self.mu_volume = scaled("volume", self.mu_volume, 0.5 * (self.compute_volume() - self.target_volume) ** 2)
        if self.mu_surface > 0.0:
            self.mu_surface = scaled("surface", self.mu_surface, 0.5 * (self.compute_surface() - self.target_surface) ** 2)
        if self.mu_barycenter > 0.0:
            bc = self.compute_barycenter()
            self.mu_barycenter = scaled("barycenter", self.mu_barycenter,
                                        0.5 * sum((bc[k] - self.target_barycenter[k]) ** 2 for k in range(3)))
        if self.mu_curvature > 0.0:
            mu, self.mu_curvature = self.mu_curvature, 1.0
            self.mu_curvature = scaled("curvature", mu, self.compute_curvature_objective())

    # -- geometric quantities
    @property
    def j_volume(self) -> float:
        return float(getattr(self._cost, "c_volume", 0.0) or 0.0) if self._j_volume_possible else 0.0

    @property
    def j_surface(self) -> float:
        return float(getattr(self._cost, "c_surface", 0.0) or 0.0) if self._j_surface_possible else 0.0

    def geometric_cost_terms(self) -> float:
        """``c_volume int 1 dx + c_surface int 1 ds`` of a plain IntegralFunctional (lists evaluate their own terms)."""
        value = 0.0
        if self.j_volume != 0.0:
            value += self.j_volume * self.compute_volume()
        if self.j_surface != 0.0:
            value += self.j_surface * self.compute_surface()
        return value

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

    # -- curvature (shape_regularization.py:512-646)
    def _setup_curvature(self) -> None:
        """The P1 space on the boundary vertices (``boundary_space.BoundarySpace``: compressed numbering, pattern of the
        facet graph, mass matrix) and the buffers of the mean-curvature vector."""
        from cashocs_b200._forms import boundary_space

        mesh = self.mesh
        d = mesh.dim
        f64 = dict(dtype=torch.float64, device=mesh.device)
        self._bspace = boundary_space.BoundarySpace(mesh, self._facets, "factor_curvature")
        self.boundary_mass = self._bspace.mass
        self.kappa_curvature = torch.zeros(mesh.num_vertices * d, **f64)
        self._kappa_b = [torch.zeros(self._bspace.n, **f64) for _ in range(d)]
        self._curv_vector = torch.zeros(mesh.n_owned_vertices * d, **f64)
        self.curvature_iterations = 0

    def _facet_measure_gradient(self, out: torch.Tensor) -> torch.Tensor:
        mesh = self.mesh
        _lib.check(
            _lib.load().cb_boundary_measure_gradient(
                _lib.stream_ptr(), mesh.dim, int(self._facets.shape[0]), mesh.n_owned_vertices, _lib.ptr(mesh.coords),
                _lib.ptr(self._facets), _lib.ptr(self._v2f_ptr), _lib.ptr(self._v2f_idx), _lib.ptr(self._facet_vec),
                _lib.ptr(out)),
            "cb_boundary_measure_gradient",
        )
        return out

    def compute_curvature(self) -> torch.Tensor:
        """``_compute_curvature`` (``:596-613``): the mean-curvature vector, zero at the interior vertices."""
        d, sp = self.mesh.dim, self._bspace
        sp.assemble_mass()
        b = self._facet_measure_gradient(self._curv_vector).view(-1, d)
        kappa = self.kappa_curvature.view(-1, d)
        sp.iterations = 0
        for c in range(d):
            self._kappa_b[c].copy_(sp.project(b[:, c], kappa[:, c]))
        self.curvature_iterations = sp.iterations
        return self.kappa_curvature

    def compute_curvature_objective(self) -> float:
        """``mu/2 int_Gamma kappa . kappa ds`` (``:577-594``)."""
        self.compute_curvature()
        A = self.boundary_mass
        return 0.5 * self.mu_curvature * sum(A.scalar_product(k, k) for k in self._kappa_b)

    def update_geometric_quantities(self) -> None:
        """``ShapeRegularization.update_geometric_quantities`` (``:670-673``)."""
        if self.is_active:
            self.current_volume = self.compute_volume()
            if self.mu_barycenter > 0.0:
                self.current_barycenter = self.compute_barycenter()
            if self.mu_surface > 0.0 or self._j_surface_possible:
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
        if self.mu_curvature > 0.0:
            value += self.compute_curvature_objective()
        return value

    def add_shape_derivative(self, b: torch.Tensor) -> torch.Tensor:
        """``b += dJ_reg[V]`` with the forms of ``:154-174`` and ``:353-425`` (the barycenter form verbatim, i.e. with
        the reference's ``+ bc_k / vol * div V``); call after :meth:`update_geometric_quantities`."""
        if not self.is_active:
            return b
        d = self.mesh.dim
        if self.mu_surface > 0.0 or self.j_surface != 0.0:  # (mu (surface - target) + c_surface) int_Gamma t_div(V, n) ds  (:248-269)
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
            coef = self.j_surface
            if self.mu_surface > 0.0:
                coef += self.mu_surface * (self.current_surface - self.target_surface)
            b.add_(self._surface_vector, alpha=coef)
        if self.mu_curvature > 0.0:  # :554-575 with the mean-curvature vector of the current mesh
            mesh = self.mesh
            self.compute_curvature()
            _lib.check(
                _lib.load().cb_boundary_curvature_derivative(
                    _lib.stream_ptr(), d, int(self._facets.shape[0]), mesh.n_owned_vertices, _lib.ptr(mesh.coords),
                    _lib.ptr(self._facets), _lib.ptr(self.kappa_curvature), _lib.ptr(self._v2f_ptr),
                    _lib.ptr(self._v2f_idx), _lib.ptr(self._facet_vec), _lib.ptr(self._curv_vector)),
                "cb_boundary_curvature_derivative",
            )
            if self.bc_flag is not None:
                self._curv_vector.masked_fill_(self.bc_flag[: self._curv_vector.numel()].bool(), 0.0)
            b.add_(self._curv_vector, alpha=self.mu_curvature)
        if self.mu_volume <= 0.0 and self.mu_barycenter <= 0.0 and self.j_volume == 0.0:
            return b
        vol, bc, target = self.current_volume, self.current_barycenter, self.target_barycenter
        a0 = self.j_volume + (self.mu_volume * (vol - self.target_volume) if self.mu_volume > 0.0 else 0.0)
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
