"""Riesz projection of the shape derivative (``cashocs/_pde_problems/shape_gradient_problem.py:117-173``) and the
p-Laplace projector (``:329-427``)."""

from __future__ import annotations

import copy

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200 import _lib
from cashocs_b200._forms import forms
from cashocs_b200._pde_problems import pde_problem
from cashocs_b200._utils import linalg


class ShapeGradientProblem(pde_problem.PDEProblem):
    def __init__(self, mesh, config, form_handler, state_problem, adjoint_problem, gradient: forms.Function,
                 gradient_ksp_options=None, linear_solver=None) -> None:
        super().__init__(linear_solver)
        self.mesh, self.config = mesh, config
        self.form_handler = form_handler
        self.state_problem, self.adjoint_problem = state_problem, adjoint_problem
        self.gradient = gradient
        self.gradient_norm_squared = 1.0
        self.iterations = 0
        gradient_tol = config.getfloat("OptimizationRoutine", "gradient_tol")
        gradient_method = config.get("OptimizationRoutine", "gradient_method")
        # option choice of shape_gradient_problem.py:81-91
        if gradient_ksp_options is not None:
            self.ksp_options = gradient_ksp_options
        elif gradient_method.casefold() == "direct":
            self.ksp_options = copy.deepcopy(linalg.direct_ksp_options)
        else:
            self.ksp_options = copy.deepcopy(linalg.iterative_ksp_options)
            self.ksp_options["ksp_rtol"] = gradient_tol
        self.ksp_options_given = self.ksp_options
        self.p_laplace_projector = None

    def solve(self):
        self.state_problem.solve()
        self.adjoint_problem.solve()
        if not self.has_solution:
            fh = self.form_handler
            b = fh.assemble_shape_derivative()
            g = self.gradient.vector
            if getattr(fh, "use_p_laplacian", False):  # shape_gradient_problem.py:133-139
                if self.p_laplace_projector is None:
                    self.p_laplace_projector = _PLaplaceProjector(self, self.ksp_options)
                self.p_laplace_projector.solve(b)
                self.iterations = self.p_laplace_projector.linear_iterations
            else:
                g.zero_()
                res = self.linear_solver.solve(g, A=fh.scalar_product_matrix, b=b, ksp_options=self.ksp_options)
                self.iterations = res.its
                fh.apply_shape_bcs(g[: b.numel()])
            if self.mesh.dist is not None:
                self.mesh.dist.halo_exchange(g, self.mesh.dim)
            self.reextend_gradient_from_boundaries()
            self.has_solution = True
            self.gradient_norm_squared = fh.scalar_product(g, g)
            if self.post_callback is not None:
                self.post_callback()
        return [self.gradient]

    def reextend_gradient_from_boundaries(self) -> None:
        """Re-extends the gradient deformation to the volume from its boundary values
        (``shape_gradient_problem.py:175-223``; mode ``normal`` first replaces it by its normal part): one more solve with the scalar product's operator,
        now with Dirichlet conditions on every boundary; the result overwrites the gradient deformation."""
        fh = self.form_handler
        if not getattr(fh, "reextend_from_boundary", False):
            return
        g = self.gradient.vector
        if fh.reextension_mode == "normal":  # shape_gradient_problem.py:186-191
            w = getattr(self, "_normal_deformation", None)
            if w is None or w.numel() != g.numel():
                w = self._normal_deformation = g.new_zeros(g.numel())
            g.copy_(fh.compute_normal_deformation(g, w))
        b = fh.assemble_reextension_vector(g)
        x = getattr(self, "_reextended", None)
        if x is None or x.numel() != g.numel():
            x = self._reextended = g.new_zeros(g.numel())
        x.zero_()
        res = self.linear_solver.solve(x, A=fh.reextension_matrix, b=b, ksp_options=self.ksp_options)
        self.reextension_iterations = res.its
        fh.apply_reextension_bcs(x[: b.numel()], g)
        g[: b.numel()].copy_(x[: b.numel()])
        if self.mesh.dist is not None:
            self.mesh.dist.halo_exchange(g, self.mesh.dim)


class _PLaplaceProjector:
    """Gradient deformation from a p-Laplace projection (``shape_gradient_problem.py:329-427``).

    For ``p = 2, 3, ..., p_laplacian_power`` solve
    ``int mu (eps + |grad U|^(p-2)) grad U : grad W + delta U . W dx = dJ[W]`` with Newton's method, each ``p`` starting
    from the previous solution (``U = 0`` for ``p = 2``, which is linear).  The reference drives PETSc's SNES
    (``newtonls``, ``basic`` line search = full steps, default tolerances rtol 1e-8 / atol 1e-50 / stol 1e-8 / 50
    iterations, Eisenstat-Walker forcing capped at 0.1 for iterative inner solvers, ``:404-411``); the loop below is
    that iteration with residual and Jacobian from the kernels of ``csrc/plaplace.cu`` and the inner solves through
    ``cb_ksp_solve``.
    """

    SNES_RTOL, SNES_ATOL, SNES_STOL, SNES_MAX_IT = 1e-8, 1e-50, 1e-8, 50

    def __init__(self, gradient_problem: ShapeGradientProblem, ksp_options) -> None:
        gp = self.gp = gradient_problem
        fh = gp.form_handler
        mesh = self.mesh = gp.mesh
        d = mesh.dim
        self.p_target = fh.p_laplacian_power
        self.eps = fh.p_laplacian_stabilization
        self.delta = fh.damping_factor
        self.p_list = list(range(2, self.p_target + 1))
        self.ksp_options = ksp_options
        self.A = linalg.Matrix(mesh, d)
        f64 = dict(dtype=torch.float64, device=mesh.device)
        self.residual = torch.zeros(mesh.n_owned_vertices * d, **f64)
        self.update = torch.zeros(mesh.num_vertices * d, **f64)
        self.newton_iterations: list[int] = []
        self.linear_iterations = 0
        # Eisenstat-Walker applies to inner Krylov solvers with a relative tolerance (not to preonly + lu)
        opts = ksp_options or {}
        self.inexact = str(opts.get("ksp_type", "preonly")).casefold() != "preonly"

    def _residual(self, p: int, shift: torch.Tensor) -> float:
        mesh, fh = self.mesh, self.gp.form_handler
        _lib.check(
            _lib.load().cb_plaplace_residual(
                _lib.stream_ptr(), mesh.dim, mesh.n_owned_vertices, _lib.ptr(mesh.v2c_ptr), _lib.ptr(mesh.v2c_idx),
                _lib.ptr(mesh.cells), _lib.ptr(mesh.coords), _lib.ptr(self.gp.gradient.vector), _lib.ptr(fh.mu_lame.vector),
                p, self.eps, self.delta, _lib.ptr(shift), _lib.ptr(fh.bc_flag), _lib.ptr(self.residual)),
            "cb_plaplace_residual",
        )
        return float(torch.linalg.vector_norm(self.residual).item())

    def _jacobian(self, p: int) -> None:
        mesh, fh, A = self.mesh, self.gp.form_handler, self.A
        _lib.check(
            _lib.load().cb_plaplace_jacobian(
                _lib.stream_ptr(), mesh.dim, mesh.n_owned_vertices, _lib.ptr(mesh.v2c_ptr), _lib.ptr(mesh.v2c_idx),
                _lib.ptr(mesh.cells), _lib.ptr(mesh.coords), _lib.ptr(self.gp.gradient.vector), _lib.ptr(fh.mu_lame.vector),
                p, self.eps, self.delta, _lib.ptr(fh.bc_flag), _lib.ptr(A.rowptr), _lib.ptr(A.col), _lib.ptr(A.val)),
            "cb_plaplace_jacobian",
        )
        A.touch()

    def solve(self, shift: torch.Tensor) -> None:
        U = self.gp.gradient.vector
        U.zero_()
        self.newton_iterations = []
        self.linear_iterations = 0
        for p in self.p_list:
            fnorm0 = fnorm = self._residual(p, shift)
            its = 0
            rtol_ew, fnorm_prev = 0.1, fnorm  # snes_ksp_ew: rtol_0 = 0.3 capped by rtolmax = 0.1
            converged = fnorm < self.SNES_ATOL
            while not converged and its < self.SNES_MAX_IT:
                self._jacobian(p)
                if self.inexact and its > 0:
                    # Eisenstat-Walker, version 2 (PETSc defaults gamma = 1, alpha = (1 + sqrt 5) / 2, threshold 0.1)
                    alpha = 0.5 * (1.0 + np.sqrt(5.0))
                    new = (fnorm / fnorm_prev) ** alpha
                    safeguard = rtol_ew ** alpha
                    if safeguard > 0.1:
                        new = max(new, safeguard)
                    rtol_ew = min(new, 0.1)
                self.update.zero_()
                res = self.gp.linear_solver.solve(self.update, A=self.A, b=self.residual, ksp_options=self.ksp_options,
                                                  rtol=rtol_ew if self.inexact else None)
                self.linear_iterations += res.its
                n = self.residual.numel()
                U[:n].sub_(self.update[:n])
                its += 1
                fnorm_prev = fnorm
                fnorm = self._residual(p, shift)
                ynorm = float(torch.linalg.vector_norm(self.update[:n]).item())
                xnorm = float(torch.linalg.vector_norm(U[:n]).item())
                converged = fnorm < self.SNES_ATOL or fnorm <= self.SNES_RTOL * fnorm0 or ynorm < self.SNES_STOL * xnorm
            if not converged:
                raise _exceptions.NotConvergedError("SNES solver", "The p-Laplace Newton iteration did not converge.")
            self.newton_iterations.append(its)
