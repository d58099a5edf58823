"""``OptimalControlProblem`` (``cashocs/_optimization/optimal_control/optimal_control_problem.py:64-90``).

Distributed-control Poisson problems: ``F = kappa grad y.grad p + c y p - s u p``,
``J = 1/2|y - y_d|^2 + alpha/2 |u|^2``; ``compute_gradient`` (``:427-439``) returns the L2 Riesz
representative ``M g = alpha M u - s M p``.
"""

from __future__ import annotations

import copy

from cashocs_b200 import _exceptions
from cashocs_b200._forms import forms
from cashocs_b200._optimization import cost_functional
from cashocs_b200._optimization import line_search as ls
from cashocs_b200._optimization import optimization_algorithms as algs
from cashocs_b200._pde_problems import adjoint_problem
from cashocs_b200._pde_problems import control_gradient_problem
from cashocs_b200._pde_problems import state_problem
from cashocs_b200._utils import linalg
from cashocs_b200.io import config as config_module


def _enlist(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


class OptimalControlProblem:
    def __init__(self, state_forms, bcs_list, cost_functional_form, states, controls, adjoints, config=None,
                 ksp_options=None, adjoint_ksp_options=None, gradient_ksp_options=None, linear_solver=None,
                 adjoint_linear_solver=None, control_constraints=None, control_bcs_list=None, desired_weights=None) -> None:
        self.state_forms = _enlist(state_forms)
        self.states, self.controls, self.adjoints = _enlist(states), _enlist(controls), _enlist(adjoints)
        if not (len(self.states) == len(self.adjoints) == len(self.state_forms) == 1 and len(self.controls) == 1):
            raise _exceptions.InputError("cashocs_b200.OptimalControlProblem", "states",
                                         "one state / control / adjoint is on the hot path")
        bcs_list = _enlist(bcs_list) if bcs_list is not None else []
        self.bcs_list = [bcs_list] if (not bcs_list or isinstance(bcs_list[0], forms.DirichletBC)) else bcs_list
        self.cost = forms.as_cost_functional(cost_functional_form, force_list=desired_weights is not None)
        self.mesh = self.states[0].mesh
        self.config = config if config is not None else config_module.Config()
        self.config.validate_config()
        if self.state_forms[0].source is not self.controls[0]:
            raise _exceptions.InputError("cashocs_b200.OptimalControlProblem", "state_forms",
                                         "the control must be the source term of the state equation")
        norm = lambda o: [copy.deepcopy(linalg.direct_ksp_options)] if o is None else _enlist(o)  # noqa: E731
        self.ksp_options = norm(ksp_options)
        self.adjoint_ksp_options = norm(adjoint_ksp_options)
        self.mesh.build_pattern()
        self.gradient = forms.Function(self.mesh, 1, "gradient")
        self.state_problem = state_problem.StateProblem(
            self.mesh, self.state_forms, self.bcs_list, self.states, self.ksp_options, self.config, linear_solver)
        self.adjoint_problem = adjoint_problem.AdjointProblem(
            self.mesh, self.state_forms, self.bcs_list, self.cost, self.states, self.adjoints, self.state_problem,
            self.adjoint_ksp_options, adjoint_linear_solver)
        # Dirichlet conditions on the control (optimal_control_problem.py:212-230): the Riesz problem uses the
        # homogenised ones, `solve` imposes the values on the control (`_setup_control_bcs`, :306-311)
        self.control_bcs_list = _enlist(control_bcs_list) if control_bcs_list is not None else []
        self.gradient_problem = control_gradient_problem.ControlGradientProblem(
            self.mesh, self.cost, self.state_forms[0], self.controls[0], self.adjoints[0], self.gradient,
            self.state_problem, self.adjoint_problem, self.config,
            None if gradient_ksp_options is None else _enlist(gradient_ksp_options)[0],
            control_bcs=self.control_bcs_list)
        self.reduced_cost_functional = cost_functional.ReducedCostFunctional(
            self.mesh, self.cost, self.states, self.state_problem, self.controls[0],
            self.gradient_problem.mass_product)
        if desired_weights is not None:  # optimization_problem.py:300-303
            self.reduced_cost_functional.scale(list(desired_weights))

        # the optimisation loop sees the same surface as for shape problems
        self.problem_type = "control"
        self.form_handler = _ControlFormHandler(self.gradient_problem)
        self.mesh_handler = None
        # u_a <= u <= u_b, floats or Functions (optimal_control_problem.py:185-187, box_constraints.py)
        self.box_constraints = ls.BoxConstraints(self.controls[0], control_constraints)
        self.optimization_variable_abstractions = ls.ControlVariableAbstractions(self)
        self.solver = None

    def solve(self, algorithm: str | None = None, rtol: float | None = None, atol: float | None = None,
              max_iter: int | None = None) -> None:
        """``OptimalControlProblem.solve`` (``optimal_control_problem.py:362-425``): gradient descent, L-BFGS and
        the nonlinear CG methods and the truncated Newton method (``:324-325,346-349``)."""
        from cashocs_b200._optimization.shape_optimization_problem import _algorithm_name

        algorithm = _algorithm_name(algorithm or self.config.get("OptimizationRoutine", "algorithm"))
        if rtol is not None:
            self.config.set("OptimizationRoutine", "rtol", str(rtol))
        if atol is not None:
            self.config.set("OptimizationRoutine", "atol", str(atol))
        if max_iter is not None:
            self.config.set("OptimizationRoutine", "max_iter", str(max_iter))
        from cashocs_b200.io import managers

        self.output_manager = managers.OutputManager(self)
        self.config.set("OptimizationRoutine", "algorithm", algorithm)
        self._setup_control_bcs()
        line_search = ls.create_line_search(self)
        if algorithm == "gradient_descent":
            self.solver = algs.GradientDescentMethod(self, line_search)
        elif algorithm == "bfgs":
            self.solver = algs.LBFGSMethod(self, line_search)
        elif algorithm == "conjugate_gradient":
            self.solver = algs.NonlinearCGMethod(self, line_search)
        elif algorithm == "newton":
            from cashocs_b200._pde_problems import hessian_problems

            self.hessian_problem = hessian_problems.HessianProblem(self)
            self.solver = algs.NewtonMethod(self, line_search)
        else:
            raise _exceptions.InputError(
                "cashocs_b200.OptimalControlProblem.solve", "algorithm",
                f"{algorithm!r}: the device path drives gradient_descent, bfgs, conjugate_gradient and newton")
        self.solver.run()
        self.solver.post_processing()

    def _setup_control_bcs(self) -> None:
        """``_setup_control_bcs`` (``optimal_control_problem.py:306-311``): the control takes its boundary values."""
        if self.control_bcs_list:
            import torch

            flag, val = forms.bc_flag_arrays(self.mesh, self.control_bcs_list, 1)
            if flag is not None:
                u = self.controls[0].vector
                u.copy_(torch.where(flag.bool(), val, u))
                self._erase_pde_memory()

    # ------------------------------------------------------------------ callbacks (optimization_problem.py:600-640)
    def inject_pre_callback(self, function) -> None:
        """``function()`` is called before every solve of the state system."""
        self.state_problem.pre_callback = function
        self._forget_solutions()

    def inject_post_callback(self, function) -> None:
        """``function()`` is called after every computation of the gradient."""
        self.gradient_problem.post_callback = function
        self._forget_solutions()

    def inject_pre_post_callback(self, pre_function, post_function) -> None:
        self.inject_pre_callback(pre_function)
        self.inject_post_callback(post_function)

    def _forget_solutions(self) -> None:
        self.state_problem.has_solution = False
        self.adjoint_problem.has_solution = False
        self.gradient_problem.has_solution = False

    def compute_state_variables(self) -> None:
        self.state_problem.solve()

    def compute_adjoint_variables(self) -> None:
        self.state_problem.solve()
        self.adjoint_problem.solve()

    def compute_gradient(self):
        return self.gradient_problem.solve()

    def _erase_pde_memory(self) -> None:
        self.state_problem.has_solution = False
        self.adjoint_problem.has_solution = False
        self.gradient_problem.has_solution = False

    def gradient_test(self, u=None, h=None, rng=None) -> float:
        """``cashocs.verification.control_gradient_test`` on this problem."""
        from cashocs_b200 import verification

        return verification.control_gradient_test(self, u, h, rng)


class _ControlFormHandler:
    """The part of ``ControlFormHandler`` the optimisation loop touches: the Riesz scalar product
    (``control_form_handler.py:184-207``); the mass matrix never changes, so there is nothing to update."""

    def __init__(self, gradient_problem) -> None:
        self.gradient_problem = gradient_problem

    def scalar_product(self, a, b) -> float:
        return self.gradient_problem.scalar_product(a, b)

    def update_scalar_product(self) -> None:
        return None
