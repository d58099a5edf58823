"""State system (``cashocs/_pde_problems/state_problem.py``), linear branch ``:142-172``."""

from __future__ import annotations

import torch

from cashocs_b200 import _exceptions
from cashocs_b200._forms import expressions
from cashocs_b200._forms import forms
from cashocs_b200._pde_problems import pde_problem
from cashocs_b200._utils import linalg


class StateProblem(pde_problem.PDEProblem):
    """Assemble + solve the (linear) state equation for every state variable."""

    def __init__(self, mesh, state_forms, bcs_list, states, ksp_options, config, linear_solver=None) -> None:
        super().__init__(linear_solver)
        self.mesh = mesh
        self.state_forms = state_forms
        self.bcs_list = bcs_list
        self.states = states
        self.ksp_options = ksp_options
        self.config = config
        # is_linear = False is the reference's default (io/config.py:599): Newton's method.  The supported form
        # family is linear, so the Newton iteration of nonlinear_solvers/newton_solver.py:285-362 degenerates
        # to (at most two) linear solves with the residual tests kept -- see _newton_solve.
        self.is_linear = config.getboolean("StateSystem", "is_linear")
        self.newton_rtol = config.getfloat("StateSystem", "newton_rtol")
        self.newton_atol = config.getfloat("StateSystem", "newton_atol")
        self.newton_iter = config.getint("StateSystem", "newton_iter")
        self.newton_iterations = [0] * len(states)
        if config.getboolean("StateSystem", "picard_iteration"):
            raise _exceptions.InputError(
                "cashocs_b200.StateProblem", "config",
                "Picard iterations couple several state equations; one state variable is on the hot path.",
            )
        self.number_of_solves = 0
        self.iterations = [0] * len(states)
        # tensors preallocated once, like A_tensors / b_tensors (state_problem.py:101-108)
        # P2 states (BASELINE configs[4]) live on mesh.p2_space(); P1 states on the mesh's vertex space
        self.spaces = [y.space for y in states]
        self.A_tensors = [linalg.Matrix(sp if sp is not None else mesh, 1) for sp in self.spaces]
        self.b_tensors = [torch.zeros(sp.n_owned_dofs if sp is not None else mesh.n_owned_vertices, dtype=torch.float64,
                                      device=mesh.device) for sp in self.spaces]
        self.bc_arrays = [forms.bc_flag_arrays(mesh, bcs, 1, space=sp) for bcs, sp in zip(bcs_list, self.spaces)]
        self._checkpoint = None
        self.after_solve = None  # CostFunctionalList.update: effective cost coefficients follow the state

    def _generate_checkpoint(self) -> None:
        """``state_problem.py:252-258``."""
        self._checkpoint = [y.vector.clone() for y in self.states]

    def revert_to_checkpoint(self) -> None:
        """``state_problem.py:260-274``."""
        if self._checkpoint is not None:
            for y, c in zip(self.states, self._checkpoint):
                y.vector.copy_(c)

    def assemble(self, i: int):
        form = self.state_forms[i]
        flag, val = self.bc_arrays[i]
        src = form.source
        kw = {}
        if isinstance(src, expressions.Polynomial):
            kw = dict(f_poly=src, f_poly_scale=form.source_scale)
        elif isinstance(src, forms.Function):
            if src.degree != self.states[i].degree:
                raise _exceptions.InputError("cashocs_b200.StateProblem", "source",
                                             "a Function source must live in the state's space")
            kw = dict(f_nodal=src.vector, f_nodal_scale=form.source_scale)
        if self.spaces[i] is not None:
            return linalg.assemble_p2_system(self.spaces[i], form.kappa, form.reaction, bc_flag=flag, bc_val=val,
                                             A=self.A_tensors[i], b=self.b_tensors[i], **kw)
        return linalg.assemble_scalar_system(
            self.mesh, form.kappa, form.reaction, bc_flag=flag, bc_val=val,
            A=self.A_tensors[i], b=self.b_tensors[i], **kw,
        )

    def _newton_solve(self, i: int, A, b) -> None:
        """``newton_solve`` (``cashocs/nonlinear_solvers/newton_solver.py:285-362``) for F(u) = A u - b: the
        current state is the initial guess, Dirichlet values are imposed on it, then
        ``while ||F(u)|| > atol + rtol ||F(u_0)||: solve A du = -F(u); u += du`` (for a linear F the damped
        line search always accepts the full step)."""
        sp = self.spaces[i]
        plan = sp.dist if sp is not None else self.mesh.dist
        n = A.nrows
        u = self.states[i].vector
        flag, val = self.bc_arrays[i]
        if flag is not None:
            u.copy_(torch.where(flag.bool(), val, u))

        def residual():
            if plan is not None:
                plan.halo_exchange(u, 1)
            r = b - A.mult(u)
            nrm2 = torch.sum(r * r).reshape(1)
            if plan is not None:
                plan.allreduce(nrm2, "sum")
            return r, float(nrm2.sqrt().item())

        r, res0 = residual()
        self.newton_iterations[i] = 0
        self.iterations[i] = 0
        if res0 == 0.0:  # the input already is a solution (newton_solver.py:300-309)
            return
        tol = self.newton_atol + self.newton_rtol * res0
        res = res0
        du = torch.zeros_like(u)
        while res > tol and self.newton_iterations[i] < self.newton_iter:
            self.newton_iterations[i] += 1
            out = self.linear_solver.solve(du, A=A, b=r, ksp_options=self.ksp_options[i])
            self.iterations[i] += out.its
            u[:n].add_(du[:n])
            r, res = residual()
        if not (res <= tol):
            raise _exceptions.NotConvergedError("Newton solver", "Maximum number of iterations were exceeded.")

    def _newton_solve_semilinear(self, i: int, A, b) -> None:
        """``newton_solve`` (``newton_solver.py:285-362``) for F(u) = A u + c int u^3 phi - b: residual and Jacobian
        re-assembled on the device in every step (``cb_assemble_cubic``), undamped full steps, convergence on
        ``||F(u)|| <= atol + rtol ||F(u_0)||`` ("combined"), ``NotConvergedError`` after ``newton_iter`` steps.  ``A``,
        ``b`` hold the linear part with its Dirichlet elimination and lifting, so on the free rows A u - b is the
        linear residual for an iterate that carries the boundary values."""
        form = self.state_forms[i]
        if self.spaces[i] is not None:
            raise _exceptions.InputError("cashocs_b200.StateProblem", "state_forms", "semilinear states are P1")
        mesh, plan = self.mesh, self.mesh.dist
        n = A.nrows
        u = self.states[i].vector
        flag, val = self.bc_arrays[i]
        if flag is not None:
            u.copy_(torch.where(flag.bool(), val, u))
        lin_val = A.val.clone()  # linear part of the Jacobian; the cubic part is added to a fresh copy per step
        cub = torch.zeros(n, dtype=torch.float64, device=u.device)
        free = None if flag is None else (~flag[:n].bool())

        def residual():
            if plan is not None:
                plan.halo_exchange(u, 1)
            linalg.assemble_cubic(mesh, u, form.cubic, vec=cub)
            r = A_lin.mult(u) - b + cub
            if free is not None:
                r = r * free
            nrm2 = torch.sum(r * r).reshape(1)
            if plan is not None:
                plan.allreduce(nrm2, "sum")
            return r, float(nrm2.sqrt().item())

        A_lin = linalg.Matrix(A.mesh, 1, val=lin_val)
        r, res0 = residual()
        self.newton_iterations[i] = 0
        self.iterations[i] = 0
        if res0 == 0.0:
            return
        tol = self.newton_atol + self.newton_rtol * res0
        res = res0
        du = torch.zeros_like(u)
        while res > tol and self.newton_iterations[i] < self.newton_iter:
            self.newton_iterations[i] += 1
            A.val.copy_(lin_val)
            linalg.assemble_cubic(mesh, u, form.cubic, bc_flag=flag, A=A)  # J(u) = A_lin + 3 c M[u^2]
            out = self.linear_solver.solve(du, A=A, b=-r, ksp_options=self.ksp_options[i])
            self.iterations[i] += out.its
            u[:n].add_(du[:n])
            r, res = residual()
        if not (res <= tol):
            raise _exceptions.NotConvergedError("Newton solver", "Maximum number of iterations were exceeded.")
        # leave the Jacobian at the solution in A: the adjoint operator (self-adjoint linearisation)
        A.val.copy_(lin_val)
        linalg.assemble_cubic(mesh, u, form.cubic, bc_flag=flag, A=A)

    def solve(self):
        if not self.has_solution:
            if self.pre_callback is not None:
                self.pre_callback()
            self._generate_checkpoint()
            for i in range(len(self.states)):
                A, b = self.assemble(i)
                if self.state_forms[i].is_nonlinear:
                    if self.is_linear:
                        raise _exceptions.InputError("cashocs_b200.StateProblem", "config",
                                                     "the state equation is nonlinear: set [StateSystem] is_linear = False")
                    self._newton_solve_semilinear(i, A, b)
                    continue
                if not self.is_linear:
                    self._newton_solve(i, A, b)
                    continue
                res = self.linear_solver.solve(self.states[i].vector, A=A, b=b, ksp_options=self.ksp_options[i])
                self.iterations[i] = res.its
                plan = self.spaces[i].dist if self.spaces[i] is not None else self.mesh.dist
                if plan is not None:
                    plan.halo_exchange(self.states[i].vector, 1)
            self.has_solution = True
            self.number_of_solves += 1
            if self.after_solve is not None:
                self.after_solve()
        return self.states
