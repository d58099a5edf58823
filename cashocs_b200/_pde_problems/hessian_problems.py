"""Reduced Hessian of optimal control problems for the truncated Newton method
(``cashocs/_pde_problems/hessian_problems.py:38-431``).

The reference differentiates the Lagrangian symbolically (``HessianFormHandler.compute_newton_forms``,
``cashocs/_forms/control_form_handler.py:281-360,425-542``).  For the form family of the device layer,
``F = kappa grad y.grad p + c y p - s u p`` and ``J = int c_s y + c_t/2 (y - y_d)^2 + alpha/2 u^2``, those forms are

* sensitivity equation          ``a(y', v) = s int h v``                      (``sensitivity_eqs_lhs / _rhs``),
* adjoint sensitivity equation  ``a(v, p') = w_1 = c_t int y' v``              (``L_yy[y']``; ``L_uy`` vanishes),
* ``hessian_rhs = w_2 + w_3 = alpha int h w + s int p' w``                      (``L_uu[h]``; ``L_yu`` vanishes),

all with homogenised Dirichlet conditions, followed by the Riesz solve with the control's scalar product.  One
application = two solves with the state operator (already assembled and, with AMG, already set up) + one mass solve:
the same device kernels as a gradient evaluation, nothing leaves HBM.  The inner CG / CR iterations and the
active-set splitting of ``reduced_hessian_application`` are host control flow over device vectors, as in the reference.
"""

from __future__ import annotations

import numpy as np
import torch

from cashocs_b200 import _exceptions
from cashocs_b200._utils import linalg


class HessianProblem:
    """PDE problem used to solve the (reduced) Hessian problem (``hessian_problems.py:38-147``)."""

    def __init__(self, problem) -> None:
        self.problem = problem
        self.config = problem.config
        self.form_handler = problem.form_handler
        self.gradient_problem = problem.gradient_problem
        self.ova = problem.optimization_variable_abstractions
        self.inner_newton = self.config.get("AlgoTNM", "inner_newton")
        self.max_it_inner_newton = self.config.getint("AlgoTNM", "max_it_inner_newton")
        self.inner_newton_rtol = self.config.getfloat("AlgoTNM", "inner_newton_rtol")
        self.inner_newton_atol = self.config.getfloat("AlgoTNM", "inner_newton_atol")
        g = problem.gradient.vector
        new = lambda: torch.zeros_like(g)  # noqa: E731
        self.residual, self.delta_control, self.temp1, self.temp2 = new(), new(), new(), new()
        self.p, self.q, self.s = new(), new(), new()
        self.hessian_actions, self.inactive_part, self.active_part = new(), new(), new()
        self.states_prime, self.adjoints_prime = new(), new()
        self.no_sensitivity_solves = 0
        self.inner_iterations = 0
        # hessian_problems.py:131-143: cg + boomeramg, rtol 1e-16 (clamped to the fp64 floor by the device solver)
        self.riesz_ksp_options = {
            "ksp_type": "cg", "pc_type": "hypre", "pc_hypre_type": "boomeramg",
            "pc_hypre_boomeramg_strong_threshold": 0.7, "ksp_rtol": 1e-16, "ksp_atol": 1e-50, "ksp_max_it": 100,
        }
        self.linear_solver = linalg.LinearSolver()
        self._setup_device()

    # ------------------------------------------------------------------ device part
    def _setup_device(self) -> None:
        pr = self.problem
        form = pr.state_forms[0]
        if getattr(form, "cubic", 0.0) != 0.0:
            raise _exceptions.InputError("cashocs_b200.HessianProblem", "state_forms",
                                         "second-order forms of the semilinear state equation are not built")
        if getattr(pr.cost, "tracking", None):
            # control_form_handler.py:506-520: "Newton's method is not available with scalar tracking or min_max terms."
            raise _exceptions.InputError("cashocs_b200.HessianProblem", "cost_functional_form",
                                         "Newton's method is not available with scalar tracking terms.")
        if pr.states[0].space is not None:
            raise _exceptions.InputError("cashocs_b200.HessianProblem", "states", "Newton's method drives P1 states")
        self.mesh = pr.mesh
        nv, no = self.mesh.num_vertices, self.mesh.n_owned_vertices
        f64 = dict(dtype=torch.float64, device=self.mesh.device)
        self._b = torch.zeros(no, **f64)
        self._w = torch.zeros(nv, **f64)
        self._x = torch.zeros(nv, **f64)

    def _state_solve(self, nodal: torch.Tensor, scale: float, out: torch.Tensor, ksp_options) -> None:
        """``a(out, v) = scale * int nodal v`` with the homogenised state conditions (the state operator of the last
        state solve is reused: symmetric elimination changes the right-hand side only)."""
        pr = self.problem
        sp, ap, form = pr.state_problem, pr.adjoint_problem, pr.state_forms[0]
        flag, val = ap.bc_arrays[0]
        A = sp.A_tensors[0]
        _, b = linalg.assemble_scalar_system(self.mesh, form.kappa, form.reaction, f_nodal=nodal, f_nodal_scale=scale,
                                             bc_flag=flag, bc_val=val, b=self._b, matrix=False)
        x = self._x
        x.zero_()
        self.linear_solver.solve(x, A=A, b=b, ksp_options=ksp_options)
        if self.mesh.dist is not None:
            self.mesh.dist.halo_exchange(x, 1)
        out.copy_(x)

    def hessian_application(self, h: torch.Tensor, out: torch.Tensor) -> None:
        """``hessian_problems.py:149-235``."""
        pr = self.problem
        pr.state_problem.solve()  # the operator is the one of the current state solve
        form, cost, gp = pr.state_forms[0], pr.cost, self.gradient_problem
        self._state_solve(h, form.source_scale, self.states_prime, pr.ksp_options[0])
        self._state_solve(self.states_prime, cost.c_track, self.adjoints_prime, pr.adjoint_ksp_options[0])
        w = self._w
        torch.mul(h, cost.c_control, out=w)
        w.add_(self.adjoints_prime, alpha=form.source_scale)
        _, b = linalg.assemble_scalar_system(self.mesh, 0.0, 0.0, f_nodal=w, f_nodal_scale=1.0, bc_flag=gp.bc_flag,
                                             bc_val=gp.bc_val, b=self._b, matrix=False)
        x = self._x
        x.zero_()
        self.linear_solver.solve(x, A=gp.riesz_projection_matrix, b=b, ksp_options=self.riesz_ksp_options)
        if self.mesh.dist is not None:
            self.mesh.dist.halo_exchange(x, 1)
        out.copy_(x)
        self.no_sensitivity_solves += 2

    # ------------------------------------------------------------------ host control flow (device vectors)
    def reduced_hessian_application(self, h: torch.Tensor, out: torch.Tensor) -> None:
        """``hessian_problems.py:237-265``: the identity on the active set, the Hessian on the inactive one."""
        self.ova.restrict_to_inactive_set(h, self.inactive_part)
        self.hessian_application(self.inactive_part, self.hessian_actions)
        self.ova.restrict_to_inactive_set(self.hessian_actions, self.inactive_part)
        self.ova.restrict_to_active_set(h, self.active_part)
        torch.add(self.active_part, self.inactive_part, out=out)

    def newton_solve(self) -> torch.Tensor:
        """``hessian_problems.py:267-286``."""
        self.gradient_problem.solve()
        self.ova.compute_active_sets()
        self.delta_control.zero_()
        if self.inner_newton.casefold() == "cg":
            self.cg()
        elif self.inner_newton.casefold() == "cr":
            self.cr()
        return self.delta_control

    def _split_product(self, a: torch.Tensor, b: torch.Tensor) -> float:
        """``<a_A, b_A> + <a_I, b_I>`` -- the reference evaluates these products set by set."""
        sp = self.form_handler.scalar_product
        self.ova.restrict_to_active_set(a, self.temp1)
        self.ova.restrict_to_active_set(b, self.temp2)
        v = sp(self.temp1, self.temp2)
        self.ova.restrict_to_inactive_set(a, self.temp1)
        self.ova.restrict_to_inactive_set(b, self.temp2)
        return v + sp(self.temp1, self.temp2)

    def cg(self) -> None:
        """Truncated conjugate gradients on the Newton system (``hessian_problems.py:288-343``)."""
        sp = self.form_handler.scalar_product
        torch.neg(self.problem.gradient.vector, out=self.residual)
        self.p.copy_(self.residual)
        rsold = sp(self.residual, self.residual)
        eps_0 = np.sqrt(rsold)
        self.inner_iterations = 0
        for _ in range(self.max_it_inner_newton):
            self.reduced_hessian_application(self.p, self.q)
            self.inner_iterations += 1
            self.ova.restrict_to_active_set(self.p, self.temp1)
            sp_val1 = sp(self.temp1, self.temp1)
            self.ova.restrict_to_inactive_set(self.p, self.temp1)
            self.ova.restrict_to_inactive_set(self.q, self.temp2)
            alpha = rsold / (sp_val1 + sp(self.temp1, self.temp2))
            self.delta_control.add_(self.p, alpha=alpha)
            self.residual.add_(self.q, alpha=-alpha)
            rsnew = sp(self.residual, self.residual)
            if np.sqrt(rsnew) < self.inner_newton_atol + self.inner_newton_rtol * eps_0:
                break
            self.p.mul_(rsnew / rsold).add_(self.residual)
            rsold = rsnew

    def cr(self) -> None:
        """Truncated conjugate residuals (``hessian_problems.py:345-431``)."""
        sp = self.form_handler.scalar_product
        torch.neg(self.problem.gradient.vector, out=self.residual)
        self.p.copy_(self.residual)
        eps_0 = np.sqrt(sp(self.residual, self.residual))
        self.reduced_hessian_application(self.residual, self.s)
        self.q.copy_(self.s)
        rar = self._split_product(self.residual, self.s)
        self.inner_iterations = 0
        for i in range(self.max_it_inner_newton):
            self.inner_iterations += 1
            alpha = rar / self._split_product(self.q, self.q)
            self.delta_control.add_(self.p, alpha=alpha)
            self.residual.add_(self.q, alpha=-alpha)
            eps = np.sqrt(sp(self.residual, self.residual))
            if eps < self.inner_newton_atol + self.inner_newton_rtol * eps_0 or i == self.max_it_inner_newton - 1:
                break
            self.reduced_hessian_application(self.residual, self.s)
            rar_new = self._split_product(self.residual, self.s)
            beta = rar_new / rar
            self.p.mul_(beta).add_(self.residual)
            self.q.mul_(beta).add_(self.s)
            rar = rar_new
