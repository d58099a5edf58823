"""Shape form handling (``cashocs/_forms/shape_form_handler.py``): Lame parameters, the
linear-elasticity Riesz scalar product, its boundary conditions and the shape derivative."""

from __future__ import annotations

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200._forms import forms
from cashocs_b200._utils import linalg


class Stiffness:
    """Lame mu from an auxiliary Laplace problem (``shape_form_handler.py:56-235``).

    ``-Laplace(mu) = 0`` with ``mu = mu_fix`` on ``shape_bdry_fix`` and ``mu_def`` on
    ``shape_bdry_def`` whenever ``|mu_def - mu_fix| / mu_fix > 1e-2`` (``:110``), otherwise the
    constant ``mu_fix`` (``:210-214``).  The reference solves with CG + BoomerAMG at rtol 1e-16
    / max_it 100 (``:113-120``); the device maps that option set explicitly
    (``linalg.parse_ksp_options``).
    """

    def __init__(self, mu_lame: forms.Function, config, mesh, shape_bdry_def, shape_bdry_fix,
                 linear_solver: linalg.LinearSolver | None = None, ksp_options=None) -> None:
        self.mu_lame = mu_lame
        self.config = config
        self.mesh = mesh
        self.shape_bdry_def = list(shape_bdry_def)
        self.shape_bdry_fix = list(shape_bdry_fix)
        self.linear_solver = linear_solver or linalg.LinearSolver()
        self.inhomogeneous_mu = False
        self.use_distance_mu = config.getboolean("ShapeGradient", "use_distance_mu")
        self.distance = None
        if self.use_distance_mu:
            self._setup_distance_mu()
            return
        mu_def = config.getfloat("ShapeGradient", "mu_def")
        mu_fix = config.getfloat("ShapeGradient", "mu_fix")
        self.options_mu = ksp_options or {
            "ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg",
            "ksp_rtol": 1e-16, "ksp_atol": 1e-50, "ksp_max_it": 100,
        }
        if np.abs(mu_def - mu_fix) / mu_fix > 1e-2:
            self.inhomogeneous_mu = True
            self.bcs_mu = forms.create_dirichlet_bcs(mu_fix, self.shape_bdry_fix)
            self.bcs_mu += forms.create_dirichlet_bcs(mu_def, self.shape_bdry_def)
            self.bc_flag, self.bc_val = forms.bc_flag_arrays(mesh, self.bcs_mu, 1)
            self.A_mu_matrix = linalg.Matrix(mesh, 1)
            self.b_mu = torch.zeros(mesh.n_owned_vertices, dtype=torch.float64, device=mesh.device)
        self.last_its = 0

    def _setup_distance_mu(self) -> None:
        """``_setup_mu_computation``, distance branch (``shape_form_handler.py:141-186``): mu_min within dist_min of
        the boundaries ``boundaries_dist``, mu_max beyond dist_max, a linear (or, ``smooth_mu``, cubic) blend between."""
        from cashocs_b200.geometry import boundary_distance

        cfg = self.config
        self.mu_min, self.mu_max = cfg.getfloat("ShapeGradient", "mu_min"), cfg.getfloat("ShapeGradient", "mu_max")
        self.dist_min, self.dist_max = cfg.getfloat("ShapeGradient", "dist_min"), cfg.getfloat("ShapeGradient", "dist_max")
        self.smooth_mu = cfg.getboolean("ShapeGradient", "smooth_mu")
        self.last_its = 0
        self.distance_active = np.abs(self.mu_min - self.mu_max) / self.mu_min > 1e-2
        method = cfg.get("ShapeGradient", "distance_method")
        if method != "eikonal":
            raise _exceptions.InputError("cashocs_b200.Stiffness", "distance_method",
                                         "the device path computes the distance with the eikonal iteration")
        if self.distance_active:
            idcs = [int(m) for m in cfg.getlist("ShapeGradient", "boundaries_dist")]
            self.eikonal = boundary_distance.EikonalDistance(self.mesh, idcs, linear_solver=self.linear_solver)

    def _compute_distance_mu(self) -> None:
        """``Stiffness.compute``, distance branch (``shape_form_handler.py:216-235``)."""
        from cashocs_b200 import _lib

        if not self.distance_active:
            # the reference leaves mu_expression undefined in this case (:148) and fails at the first compute
            raise _exceptions.InputError("cashocs_b200.Stiffness", "mu_max",
                                         "use_distance_mu needs |mu_min - mu_max| / mu_min > 1e-2")
        self.distance = self.eikonal.compute()
        self.last_its = self.eikonal.iterations
        mu = self.mu_lame.vector
        _lib.check(_lib.load().cb_distance_mu(_lib.stream_ptr(), mu.numel(), _lib.ptr(self.distance), self.dist_min,
                                              self.dist_max, self.mu_min, self.mu_max, int(self.smooth_mu), _lib.ptr(mu)),
                   "cb_distance_mu")

    def compute(self) -> None:
        """Computes the elasticity parameter mu (``shape_form_handler.py:188-235``)."""
        if self.use_distance_mu:
            self._compute_distance_mu()
            return
        if self.inhomogeneous_mu:
            linalg.assemble_scalar_system(self.mesh, 1.0, 0.0, bc_flag=self.bc_flag, bc_val=self.bc_val,
                                          A=self.A_mu_matrix, b=self.b_mu)
            res = self.linear_solver.solve(self.mu_lame.vector, A=self.A_mu_matrix, b=self.b_mu,
                                           ksp_options=self.options_mu)
            self.last_its = res.its
            if self.mesh.dist is not None:
                self.mesh.dist.halo_exchange(self.mu_lame.vector, 1)
            if self.config.getboolean("ShapeGradient", "use_sqrt_mu"):
                self.mu_lame.vector.abs_().sqrt_()
        else:
            self.mu_lame.vector.fill_(self.config.getfloat("ShapeGradient", "mu_fix"))


class ShapeFormHandler:
    """Riesz scalar product + shape derivative for a shape problem (``shape_form_handler.py:238-860``)."""

    def __init__(self, mesh, config, state_form: forms.PoissonForm, cost: forms.IntegralFunctional,
                 state: forms.Function, adjoint: forms.Function, mu_linear_solver=None) -> None:
        self.mesh = mesh
        self.config = config
        self.state_form, self.cost = state_form, cost
        self.state, self.adjoint = state, adjoint
        d = mesh.dim
        g = lambda k: [int(m) for m in config.getlist("ShapeGradient", k)]  # noqa: E731
        self.shape_bdry_def = g("shape_bdry_def")
        self.shape_bdry_fix = g("shape_bdry_fix")
        self.shape_bdry_fix_x = g("shape_bdry_fix_x")
        self.shape_bdry_fix_y = g("shape_bdry_fix_y")
        self.shape_bdry_fix_z = g("shape_bdry_fix_z")
        self.fixed_dimensions = [int(i) for i in config.getlist("ShapeGradient", "fixed_dimensions")]
        self.use_fixed_dimensions = len(self.fixed_dimensions) > 0
        # p-Laplace projection (shape_gradient_problem.py:93-115): incompatible with fixed dimensions, where the
        # reference falls back to the linear-elasticity projection with a warning
        self.use_p_laplacian = config.getboolean("ShapeGradient", "use_p_laplacian")
        if self.use_p_laplacian and self.use_fixed_dimensions:
            from cashocs_b200 import log

            log.warning("Incompatible config settings: use_p_laplacian and fixed_dimensions are incompatible. "
                        "Falling back to use_p_laplacian=False.")
            self.use_p_laplacian = False
        if self.use_p_laplacian and mesh.dist is not None:
            raise _exceptions.InputError("cashocs_b200.ShapeFormHandler", "use_p_laplacian",
                                         "the p-Laplace projection runs on unpartitioned meshes")
        self.p_laplacian_power = config.getint("ShapeGradient", "p_laplacian_power")
        self.p_laplacian_stabilization = config.getfloat("ShapeGradient", "p_laplacian_stabilization")
        self.reextend_from_boundary = config.getboolean("ShapeGradient", "reextend_from_boundary")
        self.reextension_mode = config.get("ShapeGradient", "reextension_mode")
        self._normal_space = None
        self.inhomogeneous = config.getboolean("ShapeGradient", "inhomogeneous")
        self.update_inhomogeneous = config.getboolean("ShapeGradient", "update_inhomogeneous")
        self.lambda_lame = config.getfloat("ShapeGradient", "lambda_lame")
        self.damping_factor = config.getfloat("ShapeGradient", "damping_factor")
        self.mu_lame = forms.Function(mesh, 1, "mu_lame")
        self.stiffness = Stiffness(self.mu_lame, config, mesh, self.shape_bdry_def, self.shape_bdry_fix,
                                   linear_solver=mu_linear_solver)
        self.bcs_shape = self._setup_bcs_shape()
        self.bc_flag, _ = forms.bc_flag_arrays(mesh, self.bcs_shape, d)
        if self.use_fixed_dimensions:
            if self.bc_flag is None:
                self.bc_flag = torch.zeros(mesh.num_vertices * d, dtype=torch.uint8, device=mesh.device)
            for c in self.fixed_dimensions:
                self.bc_flag.view(-1, d)[:, c] = 1  # _project_scalar_product (shape_form_handler.py:693-712)
        self.volumes = None
        if self.inhomogeneous:
            self._update_volumes()
        self.fe_scalar_product_matrix = linalg.Matrix(mesh, d)
        self.scalar_product_matrix = self.fe_scalar_product_matrix
        self.fe_shape_derivative_vector = torch.zeros(mesh.n_owned_vertices * d, dtype=torch.float64,
                                                      device=mesh.device)
        if self.reextend_from_boundary:  # shape_form_handler.py:343-367
            self.bcs_extension = self._setup_bcs_extension()
            self.bc_flag_ext, _ = forms.bc_flag_arrays(mesh, self.bcs_extension, d)
            if self.bc_flag_ext is None:
                self.bc_flag_ext = torch.zeros(mesh.num_vertices * d, dtype=torch.uint8, device=mesh.device)
            for c in self.fixed_dimensions:
                self.bc_flag_ext.view(-1, d)[:, c] = 1
            self.fe_reextension_matrix = linalg.Matrix(mesh, d)
            self.reextension_matrix = self.fe_reextension_matrix
            self.fe_reextension_vector = torch.zeros(mesh.n_owned_vertices * d, dtype=torch.float64, device=mesh.device)
            # mode `normal`: the boundary values are the normal deformation, which does not vanish on the fixed
            # boundaries, so the lifting needs the operator without the shape boundary conditions
            self.fe_lifting_matrix = linalg.Matrix(mesh, d) if self.reextension_mode == "normal" else None
        from cashocs_b200._forms import shape_regularization

        self.shape_regularization = shape_regularization.ShapeRegularization(mesh, config, self.bc_flag, cost)
        self.update_scalar_product()

    def _setup_bcs_shape(self) -> list[forms.DirichletBC]:
        """``shape_form_handler.py:551-588``."""
        bcs = forms.create_dirichlet_bcs(0.0, self.shape_bdry_fix)
        bcs += forms.create_dirichlet_bcs(0.0, self.shape_bdry_fix_x, 0)
        bcs += forms.create_dirichlet_bcs(0.0, self.shape_bdry_fix_y, 1)
        if self.mesh.dim == 3:
            bcs += forms.create_dirichlet_bcs(0.0, self.shape_bdry_fix_z, 2)
        return bcs

    def _setup_bcs_extension(self) -> list[forms.DirichletBC]:
        """``shape_form_handler.py:590-611``: every component on every boundary named in ``[ShapeGradient]`` (the values
        are those of the current gradient deformation, supplied when the right-hand side is assembled)."""
        all_boundaries = (self.shape_bdry_def + self.shape_bdry_fix + self.shape_bdry_fix_x + self.shape_bdry_fix_y
                          + self.shape_bdry_fix_z)
        return forms.create_dirichlet_bcs(0.0, all_boundaries)

    def _update_volumes(self) -> None:
        """Cell volumes normalised by their maximum (l2_projection of CellVolume, ``:629-644``)."""
        from cashocs_b200.geometry import measures

        vol, vmax = measures.cell_volumes(self.mesh, with_max=True)
        if self.mesh.dist is not None:
            self.mesh.dist.allreduce(vmax, "max")
        expo = self.config.getfloat("ShapeGradient", "inhomogeneous_exponent")
        self.volumes = vol / vmax
        self.cell_scale = torch.pow(self.volumes, -expo).contiguous()

    def update_scalar_product(self) -> None:
        """mu solve + elasticity re-assembly + ident_zeros (``shape_form_handler.py:714-746``)."""
        self.stiffness.compute()
        if self.inhomogeneous and self.update_inhomogeneous:
            self._update_volumes()
        linalg.assemble_elasticity_matrix(
            self.mesh, self.mu_lame.vector, self.lambda_lame, self.damping_factor, bc_flag=self.bc_flag,
            cell_scale=self.cell_scale if self.inhomogeneous else None, A=self.fe_scalar_product_matrix,
        )
        if self.reextend_from_boundary:  # shape_form_handler.py:741-744
            linalg.assemble_elasticity_matrix(
                self.mesh, self.mu_lame.vector, self.lambda_lame, self.damping_factor, bc_flag=self.bc_flag_ext,
                cell_scale=self.cell_scale if self.inhomogeneous else None, A=self.fe_reextension_matrix,
            )
            if self.fe_lifting_matrix is not None:
                linalg.assemble_elasticity_matrix(
                    self.mesh, self.mu_lame.vector, self.lambda_lame, self.damping_factor, bc_flag=None,
                    cell_scale=self.cell_scale if self.inhomogeneous else None, A=self.fe_lifting_matrix,
                )

    def compute_normal_deformation(self, g: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """``_compute_normal_deformation`` (``shape_gradient_problem.py:225-299``): the part of the deformation ``g`` normal
        to the boundaries the [ShapeGradient] section names, by two L2 projections with the boundary mass matrix --
        ``phi = P(g . n)``, then ``out = P(phi n)`` -- zero in the interior.  A boundary facet carries a marker when all
        its vertices do.  The reference solves both with cg + boomeramg at rtol 1e-8 (``:281-287``); the boundary systems
        are tiny, the device solves them tight."""
        from cashocs_b200 import _lib
        from cashocs_b200._forms import boundary_space

        mesh = self.mesh
        d = mesh.dim
        if self._normal_space is None:
            markers = (self.shape_bdry_def + self.shape_bdry_fix + self.shape_bdry_fix_x + self.shape_bdry_fix_y
                       + self.shape_bdry_fix_z)
            tagged = torch.zeros(mesh.num_vertices, dtype=torch.bool, device=mesh.device)
            tagged[mesh.boundary_vertices(markers)] = True
            facets = mesh.boundary_facet_table()[0]
            facets = facets[tagged[facets.long()].all(dim=1)].contiguous()
            sp = self._normal_space = boundary_space.BoundarySpace(mesh, facets, "reextension_mode = normal")
            mesh.build_pattern()
            sp.opposite = torch.zeros(max(sp.nf, 1), dtype=torch.int32, device=mesh.device)
            _lib.check(_lib.load().cb_boundary_facet_opposite(
                _lib.stream_ptr(), d, sp.nf, _lib.ptr(sp.facets), _lib.ptr(mesh.v2c_ptr), _lib.ptr(mesh.v2c_idx),
                _lib.ptr(mesh.cells), _lib.ptr(sp.opposite)), "cb_boundary_facet_opposite")
            f64 = dict(dtype=torch.float64, device=mesh.device)
            sp.phi = torch.zeros(mesh.num_vertices, **f64)
            sp.rhs_scalar = torch.zeros(mesh.n_owned_vertices, **f64)
            sp.rhs_vector = torch.zeros(mesh.n_owned_vertices * d, **f64)
        sp = self._normal_space
        lib, s = _lib.load(), _lib.stream_ptr()

        def rhs(vector_in: int, src: torch.Tensor, dst: torch.Tensor) -> None:
            _lib.check(lib.cb_boundary_normal_rhs(s, d, vector_in, sp.nf, mesh.n_owned_vertices, _lib.ptr(mesh.coords),
                                                  _lib.ptr(sp.facets), _lib.ptr(sp.opposite), _lib.ptr(src), _lib.ptr(sp.v2f_ptr),
                                                  _lib.ptr(sp.v2f_idx), _lib.ptr(sp.facet_vec), _lib.ptr(dst)),
                       "cb_boundary_normal_rhs")

        sp.assemble_mass()
        sp.iterations = 0
        rhs(1, g, sp.rhs_scalar)
        sp.phi.zero_()
        sp.project(sp.rhs_scalar, sp.phi)
        rhs(0, sp.phi, sp.rhs_vector)
        out.zero_()
        b, w = sp.rhs_vector.view(-1, d), out.view(-1, d)
        for c in range(d):
            sp.project(b[:, c], w[:, c])
        return out

    def scalar_product(self, a: torch.Tensor, b: torch.Tensor) -> float:
        """``b^T A a`` (``shape_form_handler.py:748-780``); with the p-Laplace projection the p-Laplace form with
        the gradient replaced by ``a`` and the test function by ``b`` (``:761-770,785-815``) -- nonlinear in ``a``."""
        if self.use_p_laplacian:
            from cashocs_b200 import _lib

            mesh = self.mesh
            out = torch.empty(1, dtype=torch.float64, device=mesh.device)
            scratch = mesh.reduce_scratch()
            _lib.require_cuda(a, b)
            _lib.check(
                _lib.load().cb_plaplace_product(
                    _lib.stream_ptr(), mesh.dim, mesh.n_owned_cells, _lib.ptr(mesh.coords), _lib.ptr(mesh.cells),
                    _lib.ptr(a), _lib.ptr(b), _lib.ptr(self.mu_lame.vector), self.p_laplacian_power,
                    self.p_laplacian_stabilization, self.damping_factor, _lib.ptr(out), _lib.ptr(scratch), scratch.numel()),
                "cb_plaplace_product",
            )
            return float(out.item())
        return self.scalar_product_matrix.scalar_product(a, b)

    def assemble_shape_derivative(self) -> torch.Tensor:
        """``assembler.assemble(fe_shape_derivative_vector)`` (``shape_gradient_problem.py:142-144``)."""
        src = self.state_form.source
        if not isinstance(src, (type(None),)) and not hasattr(src, "terms"):
            raise _exceptions.InputError("cashocs_b200.ShapeFormHandler", "source",
                                         "shape problems need a polynomial (coordinate) source term")
        if getattr(self.state_form, "cubic", 0.0) != 0.0:
            raise _exceptions.InputError("cashocs_b200.ShapeFormHandler", "forms",
                                         "the shape derivative of the cubic term is not implemented (semilinear states: "
                                         "optimal control problems)")
        general = self.cost.c_track != 0.0 or self.state_form.reaction != 0.0 or self.state_form.kappa != 1.0
        if general and self.state.degree == 2:
            raise _exceptions.InputError("cashocs_b200.ShapeFormHandler", "forms",
                                         "P2 states: shape derivative implemented for kappa=1, no reaction, J = c*int(u)")
        if self.cost.c_track != 0.0 and self.cost.desired_state.degree != 1:
            raise _exceptions.InputError("cashocs_b200.ShapeFormHandler", "desired_state", "u_d must be a P1 Function")
        from cashocs_b200._forms import expressions

        f = src if src is not None else expressions.Polynomial()
        f = f * self.state_form.source_scale
        if self.state.degree != self.adjoint.degree:
            raise _exceptions.InputError("cashocs_b200.ShapeFormHandler", "adjoint",
                                         "state and adjoint must use the same CG degree")
        if self.state.degree == 2:
            b = linalg.assemble_p2_shape_derivative(
                self.state.space, self.state.vector, self.adjoint.vector, f, self.cost.c_state,
                bc_flag=self.bc_flag, b=self.fe_shape_derivative_vector)
        else:
            b = linalg.assemble_shape_derivative(
                self.mesh, self.state.vector, self.adjoint.vector, f, self.cost.c_state, bc_flag=self.bc_flag,
                b=self.fe_shape_derivative_vector, kappa=self.state_form.kappa, reaction=self.state_form.reaction,
                c_track=self.cost.c_track,
                u_d=self.cost.desired_state.vector if self.cost.c_track != 0.0 else None,
                pull_back=self.config.getboolean("ShapeGradient", "use_pull_back"),
            )
        if self.shape_regularization.is_active:  # shape_gradient_problem.py:131 + shape_regularization.py:688-698
            self.shape_regularization.update_geometric_quantities()
            self.shape_regularization.add_shape_derivative(b)
        return b

    def assemble_reextension_vector(self, gradient: torch.Tensor) -> torch.Tensor:
        """``assembler_extension.assemble(fe_reextension_vector)`` (``shape_gradient_problem.py:193-195``): the zero
        source with the boundary values of ``gradient`` lifted symmetrically -- row i of a free dof gets
        ``-sum_j A_ij g_j`` over the Dirichlet dofs j, a Dirichlet row its value times the diagonal the elimination
        leaves there (the number of cells at the vertex).  The lifting is one SpMV with the scalar-product matrix:
        its own Dirichlet rows / columns (``bcs_shape``) belong to dofs where the gradient deformation is zero -- in
        mode ``normal`` they do not, and the operator without boundary conditions is used."""
        n, d = self.mesh.n_owned_vertices * self.mesh.dim, self.mesh.dim
        flag = self.bc_flag_ext.bool()
        gb = torch.where(flag, gradient, torch.zeros_like(gradient))
        b = self.fe_reextension_vector
        (self.fe_lifting_matrix if self.fe_lifting_matrix is not None else self.scalar_product_matrix).mult(gb, b)
        b.neg_()
        valence = self.mesh.valence[: self.mesh.n_owned_vertices].to(torch.float64).repeat_interleave(d)
        b.copy_(torch.where(flag[:n], valence * gradient[:n], b))
        return b

    def apply_reextension_bcs(self, function: torch.Tensor, gradient: torch.Tensor) -> None:
        """``shape_form_handler.py:843-862``: the boundary values of the gradient deformation (zero in the fixed
        dimensions, where the gradient deformation vanishes already)."""
        n = function.numel()
        flag = self.bc_flag_ext[:n].bool()
        function.copy_(torch.where(flag, gradient[:n], function))

    def apply_shape_bcs(self, function: torch.Tensor) -> None:
        """Impose the geometric constraints on a deformation (``shape_form_handler.py:825-841``)."""
        if self.bc_flag is not None:
            n = function.numel()
            function.masked_fill_(self.bc_flag[:n].bool(), 0.0)
