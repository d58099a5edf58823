"""Riesz projection of the shape derivative (``cashocs/_pde_problems/shape_gradient_problem.py:117-173``)."""

from __future__ import annotations

import copy

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

    def solve(self):
        self.state_problem.solve()
        self.adjoint_problem.solve()
        if not self.has_solution:
            fh = self.form_handler
            b = fh.assemble_shape_derivative()
            g = self.gradient.vector
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
        (``shape_gradient_problem.py:175-223``, mode ``surface``): one more solve with the scalar product's operator,
        now with Dirichlet conditions on every boundary; the result overwrites the gradient deformation."""
        fh = self.form_handler
        if not getattr(fh, "reextend_from_boundary", False):
            return
        g = self.gradient.vector
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
