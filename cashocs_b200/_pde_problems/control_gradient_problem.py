"""Riesz projection for optimal control (``cashocs/_pde_problems/control_gradient_problem.py:96-125``).

``M g = dL/du`` with the L2 mass matrix assembled once (``control_form_handler.py:82-137``);
for ``F = grad y.grad p - u p`` and ``J = 1/2|y-y_d|^2 + alpha/2|u|^2`` the right-hand side is
``alpha M u - M p`` (``control_form_handler.py:209-220``, pinned by tests/test_pde_problems.py:111-128).
"""

from __future__ import annotations

import copy

import torch

from cashocs_b200._forms import forms
from cashocs_b200._pde_problems import pde_problem
from cashocs_b200._utils import linalg


class ControlGradientProblem(pde_problem.PDEProblem):
    def __init__(self, mesh, cost, state_form, control: forms.Function, adjoint: forms.Function,
                 gradient: forms.Function, state_problem, adjoint_problem, config=None,
                 riesz_ksp_options=None, linear_solver=None, control_bcs=None) -> None:
        super().__init__(linear_solver)
        self.mesh, self.cost, self.state_form = mesh, cost, state_form
        self.control, self.adjoint, self.gradient = control, adjoint, gradient
        self.state_problem, self.adjoint_problem = state_problem, adjoint_problem
        # option choice of control_gradient_problem.py:76-90
        if riesz_ksp_options is not None:
            self.riesz_ksp_options = riesz_ksp_options
        else:
            method = config.get("OptimizationRoutine", "gradient_method") if config is not None else "direct"
            if method.casefold() == "direct":
                self.riesz_ksp_options = copy.deepcopy(linalg.direct_ksp_options)
            else:
                self.riesz_ksp_options = copy.deepcopy(linalg.iterative_ksp_options)
                self.riesz_ksp_options["ksp_rtol"] = config.getfloat("OptimizationRoutine", "gradient_tol")
        # homogenised Dirichlet conditions of the control (optimal_control_problem.py:212-230) are eliminated
        # symmetrically from the Riesz matrix and its right-hand side (control_form_handler.py:82-137)
        self.bc_flag, self.bc_val = (None, None)
        if control_bcs:
            hom = [forms.DirichletBC(0.0, bc.markers, bc.component) for bc in control_bcs]
            self.bc_flag, self.bc_val = forms.bc_flag_arrays(mesh, hom, 1)
        self.riesz_projection_matrix, _ = linalg.assemble_scalar_system(mesh, 0.0, 1.0, bc_flag=self.bc_flag,
                                                                        bc_val=self.bc_val, rhs=False)
        # the cost functional's alpha/2 int u^2 needs the mass matrix without the eliminated rows
        self.mass_matrix = self.riesz_projection_matrix
        if self.bc_flag is not None:
            self.mass_matrix, _ = linalg.assemble_scalar_system(mesh, 0.0, 1.0, rhs=False)
        self.b_tensor = torch.zeros(mesh.n_owned_vertices, dtype=torch.float64, device=mesh.device)
        self._work = torch.zeros(mesh.num_vertices, dtype=torch.float64, device=mesh.device)
        self.gradient_norm_squared = 1.0
        self.iterations = 0

    def scalar_product(self, a: torch.Tensor, b: torch.Tensor) -> float:
        """``ControlFormHandler.scalar_product`` (``control_form_handler.py:184-207``)."""
        return self.riesz_projection_matrix.scalar_product(a, b)

    def mass_product(self, a: torch.Tensor, b: torch.Tensor) -> float:
        """``int a b dx`` (the control term of the cost functional)."""
        return self.mass_matrix.scalar_product(a, b)

    def solve(self):
        self.state_problem.solve()
        self.adjoint_problem.solve()
        if not self.has_solution:
            # dL/du[v] = int (c_control u - source_scale p) v dx   for F = ... - source_scale*u*p
            w = self._work
            torch.mul(self.control.vector, self.cost.c_control, out=w)
            w.add_(self.adjoint.vector, alpha=-self.state_form.source_scale)
            _, b = linalg.assemble_scalar_system(self.mesh, 0.0, 0.0, f_nodal=w, f_nodal_scale=1.0,
                                                 bc_flag=self.bc_flag, bc_val=self.bc_val, b=self.b_tensor, matrix=False)
            res = self.linear_solver.solve(self.gradient.vector, A=self.riesz_projection_matrix, b=b,
                                           ksp_options=self.riesz_ksp_options)
            self.iterations = res.its
            if self.mesh.dist is not None:
                self.mesh.dist.halo_exchange(self.gradient.vector, 1)
            self.has_solution = True
            self.gradient_norm_squared = self.scalar_product(self.gradient.vector, self.gradient.vector)
            if self.post_callback is not None:
                self.post_callback()
        return [self.gradient]
