"""``ShapeOptimizationProblem`` (``cashocs/_optimization/shape_optimization/shape_optimization_problem.py``).

Same constructor vocabulary as the reference (``:125-153``): state forms, boundary conditions,
cost functional, states, adjoints, config, ``ksp_options`` / ``adjoint_ksp_options`` /
``gradient_ksp_options``, ``linear_solver`` / ``adjoint_linear_solver``; same methods
(``solve`` ``:468-530``, ``compute_shape_gradient`` ``:544-557``, ``compute_state_variables`` /
``compute_adjoint_variables`` ``optimization_problem.py:498-515``, ``gradient_test`` ``:624-641``).
Forms are descriptors of the supported family (``_forms/forms.py``) instead of UFL.
"""

from __future__ import annotations

import copy

from cashocs_b200 import _exceptions
from cashocs_b200._forms import forms
from cashocs_b200._forms import shape_form_handler
from cashocs_b200._optimization import cost_functional
from cashocs_b200._optimization import line_search as ls
from cashocs_b200._optimization import optimization_algorithms as algs
from cashocs_b200._pde_problems import adjoint_problem
from cashocs_b200._pde_problems import shape_gradient_problem
from cashocs_b200._pde_problems import state_problem
from cashocs_b200._utils import linalg
from cashocs_b200.geometry import mesh_handler
from cashocs_b200.geometry import mesh_testing
from cashocs_b200.io import config as config_module


def _enlist(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _algorithm_name(name: str) -> str:
    name = name.casefold()
    if name in ("gd", "gradient_descent"):
        return "gradient_descent"
    if name in ("bfgs", "lbfgs"):
        return "bfgs"
    if name in ("cg", "ncg", "nonlinear_cg", "conjugate_gradient"):
        return "conjugate_gradient"
    return name


class ShapeOptimizationProblem:
    def __init__(self, state_forms, bcs_list, cost_functional_form, states, adjoints, mesh=None, config=None,
                 ksp_options=None, adjoint_ksp_options=None, gradient_ksp_options=None, linear_solver=None,
                 adjoint_linear_solver=None, gradient_linear_solver=None, desired_weights=None) -> None:
        self.state_forms = _enlist(state_forms)
        self.states = _enlist(states)
        self.adjoints = _enlist(adjoints)
        if len(self.state_forms) != len(self.states) or len(self.states) != len(self.adjoints):
            raise _exceptions.InputError("cashocs_b200.ShapeOptimizationProblem", "states",
                                         "state_forms, states and adjoints must have the same length")
        if len(self.states) != 1:
            raise _exceptions.InputError("cashocs_b200.ShapeOptimizationProblem", "states",
                                         "coupled systems are outside the hot path (one state variable)")
        bcs_list = _enlist(bcs_list)
        self.bcs_list = [bcs_list] if (bcs_list and isinstance(bcs_list[0], forms.DirichletBC)) else bcs_list
        self.cost = forms.as_cost_functional(cost_functional_form, force_list=desired_weights is not None)
        self.mesh = mesh if mesh is not None else self.states[0].mesh
        self.config = config if config is not None else config_module.Config()
        self.config.validate_config()
        m = len(self.states)
        # None -> direct solver in the reference (optimization_problem.py:352-459, linalg.py:54-59)
        norm = lambda o: [copy.deepcopy(linalg.direct_ksp_options) for _ in range(m)] if o is None else _enlist(o)  # noqa: E731
        self.ksp_options = norm(ksp_options)
        self.adjoint_ksp_options = norm(adjoint_ksp_options)
        self.gradient_ksp_options = None if gradient_ksp_options is None else _enlist(gradient_ksp_options)[0]

        self.mesh.build_pattern()
        self.gradient = forms.Function(self.mesh, self.mesh.dim, "gradient")
        # sum of all accepted deformations ([ShapeGradient] global_deformation, shape_optimization_problem.py)
        self.global_deformation_function = forms.Function(self.mesh, self.mesh.dim, "global_deformation")
        self.state_problem = state_problem.StateProblem(
            self.mesh, self.state_forms, self.bcs_list, self.states, self.ksp_options, self.config, linear_solver)
        self.adjoint_problem = adjoint_problem.AdjointProblem(
            self.mesh, self.state_forms, self.bcs_list, self.cost, self.states, self.adjoints, self.state_problem,
            self.adjoint_ksp_options, adjoint_linear_solver)
        self.form_handler = shape_form_handler.ShapeFormHandler(
            self.mesh, self.config, self.state_forms[0], self.cost, self.states[0], self.adjoints[0])
        self.a_priori_tester = mesh_testing.APrioriMeshTester(self.mesh)
        self.intersection_tester = mesh_testing.IntersectionTester(self.mesh)
        self.mesh_handler = mesh_handler._MeshHandler(self.mesh, self.config, self.a_priori_tester, self.intersection_tester)
        self.gradient_problem = shape_gradient_problem.ShapeGradientProblem(
            self.mesh, self.config, self.form_handler, self.state_problem, self.adjoint_problem, self.gradient,
            self.gradient_ksp_options, gradient_linear_solver)
        self.reduced_cost_functional = cost_functional.ReducedCostFunctional(
            self.mesh, self.cost, self.states, self.state_problem)
        self.reduced_cost_functional.shape_regularization = self.form_handler.shape_regularization
        if desired_weights is not None:  # optimization_problem.py:300-303
            self.reduced_cost_functional.scale(list(desired_weights))
        self.optimization_variable_abstractions = ls.ShapeVariableAbstractions(self)
        self.solver = None

    # ------------------------------------------------------------------ reference API
    def compute_state_variables(self) -> None:
        self.state_problem.solve()

    def compute_adjoint_variables(self) -> None:
        self.state_problem.solve()
        self.adjoint_problem.solve()

    def compute_shape_gradient(self):
        return self.gradient_problem.solve()

    def _erase_pde_memory(self) -> None:
        """``shape_optimization_problem.py:416-425``."""
        self.state_problem.has_solution = False
        self.adjoint_problem.has_solution = False
        self.gradient_problem.has_solution = False
        self.form_handler.update_scalar_product()

    def solve(self, algorithm: str | None = None, rtol: float | None = None, atol: float | None = None,
              max_iter: int | None = None) -> None:
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
        line_search = ls.create_line_search(self)
        if algorithm == "gradient_descent":
            self.solver = algs.GradientDescentMethod(self, line_search)
        elif algorithm == "bfgs":
            self.solver = algs.LBFGSMethod(self, line_search)
        elif algorithm == "conjugate_gradient":
            self.solver = algs.NonlinearCGMethod(self, line_search)
        else:
            raise _exceptions.InputError("cashocs_b200.ShapeOptimizationProblem.solve", "algorithm",
                                         f"{algorithm!r}: the device path drives gradient_descent, bfgs and conjugate_gradient")
        self.solver.run()
        self.solver.post_processing()

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

    def gradient_test(self, h=None, rng=None) -> float:
        from cashocs_b200 import verification

        return verification.shape_gradient_test(self, h, rng)
