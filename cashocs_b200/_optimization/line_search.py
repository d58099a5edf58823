"""Line searches (``cashocs/_optimization/line_search/``): Armijo backtracking (``armijo_line_search.py:97-245``),
polynomial interpolation (``polynomial_line_search.py``), constant stepsize (``basic_line_search.py``), their
common part (``line_search.py:152-214,251-279``) and the variable update loops
(``shape_optimization/shape_variable_abstractions.py:91-135``, ``optimal_control/control_variable_abstractions.py``).
Host Python: only scalars cross the PCIe bus, every vector operation is a device op."""

from __future__ import annotations

import collections

import numpy as np
import torch

from cashocs_b200 import _exceptions


class ShapeVariableAbstractions:
    """Abstractions of the optimisation variables in the case of shape optimisation."""

    def __init__(self, problem) -> None:
        self.problem = problem
        self.mesh_handler = problem.mesh_handler
        self.form_handler = problem.form_handler
        self.deformation = torch.zeros_like(problem.gradient.vector)
        self.decisions: list[tuple[str, float]] = []

    def compute_decrease_measure(self, search_direction: torch.Tensor) -> float:
        return self.form_handler.scalar_product(self.problem.gradient.vector, search_direction)

    def compute_gradient_norm(self) -> float:
        g = self.problem.gradient.vector
        return float(np.sqrt(self.form_handler.scalar_product(g, g)))

    def revert_variable_update(self) -> None:
        self.mesh_handler.revert_transformation()

    def update_optimization_variables(self, search_direction: torch.Tensor, stepsize: float, beta: float) -> float:
        """Halve the step until the mesh move succeeds and the quality is >= tol_lower (``:91-135``)."""
        while True:
            torch.mul(search_direction, stepsize, out=self.deformation)
            if self.mesh_handler.move_mesh(self.deformation):
                if self.mesh_handler.current_mesh_quality < self.mesh_handler.mesh_quality_tol_lower:
                    self.decisions.append(("quality_reject", stepsize))
                    stepsize /= beta
                    self.mesh_handler.revert_transformation()
                    continue
                self.decisions.append(("accept", stepsize))
                break
            self.decisions.append(("move_reject", stepsize))
            stepsize /= beta
            if stepsize < 1e-300:
                raise _exceptions.NotConvergedError("mesh update", "no admissible step size")
        return stepsize

    def compute_a_priori_decreases(self, search_direction: torch.Tensor, stepsize: float) -> int:
        return self.mesh_handler.compute_decreases(search_direction, stepsize)

    def requires_remeshing(self) -> bool:
        return bool(self.mesh_handler.current_mesh_quality < self.mesh_handler.mesh_quality_tol_upper)

    has_box_constraints = False  # shape_variable_abstractions.py:176-232: the active-set hooks are no-ops


class BoxConstraints:
    """Box constraints ``u_a <= u <= u_b`` of an optimal control problem
    (``cashocs/_optimization/optimal_control/box_constraints.py:29-346``): bounds are floats or control-type
    functions, stored as vectors next to the control; projection, active / inactive sets as element-wise ops."""

    def __init__(self, control, control_constraints) -> None:
        u = control.vector
        if control_constraints is None:
            control_constraints = [float("-inf"), float("inf")]
        if len(control_constraints) != 2:
            raise _exceptions.InputError("cashocs_b200.OptimalControlProblem", "control_constraints",
                                         "control_constraints is [lower bound, upper bound] for the one control")
        bounds = []
        for b in control_constraints:
            bounds.append(b.vector.clone() if hasattr(b, "vector") else torch.full_like(u, float(b)))
        self.lower, self.upper = bounds
        if not bool(torch.all(self.lower < self.upper)):
            raise _exceptions.InputError(
                "cashocs._optimization.optimal_control.optimal_control_problem.OptimalControlProblem",
                "control_constraints",
                "The lower bound must always be smaller than the upper bound for the control_constraints.")
        self.require_control_constraints = not (float(self.lower.max()) == float("-inf")
                                                and float(self.upper.min()) == float("inf"))
        self.inactive = None

    def project_to_admissible_set(self, a: torch.Tensor) -> torch.Tensor:
        if self.require_control_constraints:
            torch.minimum(self.upper, a, out=a)
            torch.maximum(a, self.lower, out=a)
        return a


class ControlVariableAbstractions:
    """Abstractions of the optimisation variables for optimal control, with box constraints
    (``cashocs/_optimization/optimal_control/control_variable_abstractions.py:36-283``)."""

    def __init__(self, problem) -> None:
        self.problem = problem
        self.form_handler = problem.form_handler
        self.control = problem.controls[0]
        self.box_constraints = getattr(problem, "box_constraints", None) or BoxConstraints(self.control, None)
        self.has_box_constraints = self.box_constraints.require_control_constraints
        self.control_temp = self.control.vector.clone()
        self.projected_difference = torch.zeros_like(self.control.vector)
        self.deformation = None
        self.decisions: list[tuple[str, float]] = []

    def compute_decrease_measure(self, search_direction=None) -> float:
        """``g . (u - u_temp)`` with the actual update (``:74-96``)."""
        torch.sub(self.control.vector, self.control_temp, out=self.projected_difference)
        return self.form_handler.scalar_product(self.problem.gradient.vector, self.projected_difference)

    def compute_gradient_norm(self) -> float:
        """The stationary measure ``|u - P(u - g)|`` (``:152-191``); the gradient norm without constraints."""
        g = self.problem.gradient.vector
        if not self.has_box_constraints:
            return float(np.sqrt(self.form_handler.scalar_product(g, g)))
        w = self.projected_difference
        torch.sub(self.control.vector, g, out=w)
        self.box_constraints.project_to_admissible_set(w)
        torch.sub(self.control.vector, w, out=w)
        return float(np.sqrt(self.form_handler.scalar_product(w, w)))

    def store_optimization_variables(self) -> None:
        self.control_temp.copy_(self.control.vector)

    def revert_variable_update(self) -> None:
        self.control.vector.copy_(self.control_temp)

    def update_optimization_variables(self, search_direction: torch.Tensor, stepsize: float, beta: float) -> float:
        """``:109-150``: u = P(u_temp + stepsize * d)."""
        self.store_optimization_variables()
        self.control.vector.add_(search_direction, alpha=stepsize)
        self.box_constraints.project_to_admissible_set(self.control.vector)
        self.decisions.append(("accept", stepsize))
        return stepsize

    def compute_a_priori_decreases(self, search_direction: torch.Tensor, stepsize: float) -> int:
        return 0

    def requires_remeshing(self) -> bool:
        return False

    # -- active / inactive sets (box_constraints.py:215-324)
    def compute_active_sets(self) -> None:
        if self.has_box_constraints:
            u, bc = self.control.vector, self.box_constraints
            bc.inactive = (u > bc.lower) & (u < bc.upper)

    def restrict_to_inactive_set(self, a: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if self.has_box_constraints:
            out.copy_(torch.where(self.box_constraints.inactive, a, torch.zeros_like(a)))
        elif out is not a:
            out.copy_(a)
        return out

    def restrict_to_active_set(self, a: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        if self.has_box_constraints:
            out.copy_(torch.where(self.box_constraints.inactive, torch.zeros_like(a), a))
        else:
            out.zero_()
        return out

    def project_ncg_search_direction(self, search_direction: torch.Tensor) -> None:
        """``:212-243``: no component of the direction may point out of the admissible set."""
        if self.has_box_constraints:
            u, bc, d = self.control.vector, self.box_constraints, search_direction
            blocked = ((u <= bc.lower) & (d < 0.0)) | ((u >= bc.upper) & (d > 0.0))
            d.masked_fill_(blocked, 0.0)


class LineSearch:
    """What all line searches share (``line_search.py:40-279``)."""

    def __init__(self, problem) -> None:
        self.problem = problem
        self.problem_type = getattr(problem, "problem_type", "shape")
        cfg = problem.config
        self.config = cfg
        self.form_handler = problem.form_handler
        self.ova = problem.optimization_variable_abstractions
        self.cost_functional = problem.reduced_cost_functional
        self.state_problem = problem.state_problem
        self.epsilon_armijo = cfg.getfloat("LineSearch", "epsilon_armijo")
        self.beta_armijo = cfg.getfloat("LineSearch", "beta_armijo")
        self.stepsize = cfg.getfloat("LineSearch", "initial_stepsize")
        self.safeguard_stepsize = cfg.getboolean("LineSearch", "safeguard_stepsize")
        self.armijo_stepsize_initial = self.stepsize
        self.search_direction_inf = 1.0
        self.decrease_measure_w_o_step = 1.0
        # line_search.py:83-86: L-BFGS steps are "Newton-like", the relative stepsize criterion does not apply
        self.is_newton_like = cfg.get("OptimizationRoutine", "algorithm").casefold() in ("bfgs", "lbfgs")
        self.is_newton = cfg.get("OptimizationRoutine", "algorithm").casefold() == "newton"
        if self.is_newton:  # line_search.py:92-93
            self.stepsize = 1.0
        self.global_deformation = (self.problem_type == "shape" and cfg.getboolean("ShapeGradient", "global_deformation")
                                   and hasattr(problem, "global_deformation_function"))
        self.log: list[tuple[str, float]] = []
        # Armijo margins J_new - (J + eps * decrease), one per Armijo decision of `log`: a decision whose margin is at
        # round-off level (|margin| <~ 1e-10 |J|) is not reproducible across summation orders -- parity tests compare
        # decision sequences up to the first such trial
        self.margins: list[float] = []

    def initialize_stepsize(self, solver, search_direction: torch.Tensor, has_curvature_info: bool) -> None:
        """``line_search.py:152-214``."""
        self.search_direction_inf = float(search_direction.abs().max().item())
        if self.problem.mesh.dist is not None:
            t = torch.tensor([self.search_direction_inf], dtype=torch.float64, device=search_direction.device)
            self.problem.mesh.dist.allreduce(t, "max")
            self.search_direction_inf = float(t.item())
        if has_curvature_info:
            self.stepsize = 1.0
        if getattr(solver, "is_restarted", False):
            self.stepsize = self.config.getfloat("LineSearch", "initial_stepsize")
        num_decreases = self.ova.compute_a_priori_decreases(search_direction, self.stepsize)
        self.stepsize /= pow(self.beta_armijo, num_decreases)
        if self.safeguard_stepsize and solver.iteration == 0:
            nrm = np.sqrt(self.form_handler.scalar_product(search_direction, search_direction))
            self.stepsize = float(np.minimum(self.stepsize, 100.0 / (1.0 + nrm)))

    def _check_for_nonconvergence(self, solver) -> bool:
        """``armijo_line_search.py:59-94`` / ``polynomial_line_search.py:70-105``."""
        if solver.iteration >= solver.max_iter:
            return True
        if self.stepsize * self.search_direction_inf <= 1e-9:
            solver.line_search_broken = True
            return True
        if not self.is_newton_like and not self.is_newton and self.stepsize / self.armijo_stepsize_initial <= 1e-9:
            solver.line_search_broken = True
            return True
        return False

    def _compute_decrease_measure(self, search_direction) -> float:
        """``armijo_line_search.py:204-223``."""
        if self.problem_type == "shape":
            return self.decrease_measure_w_o_step * self.stepsize
        return self.ova.compute_decrease_measure(search_direction)

    def search(self, solver, search_direction: torch.Tensor, has_curvature_info: bool):
        raise NotImplementedError

    def perform(self, solver, search_direction, has_curvature_info):
        """``line_search.py:95-150``: search, then ``post_line_search`` (scalar product update)."""
        deformation = self.search(solver, search_direction, has_curvature_info)
        if deformation is not None and self.problem_type == "shape" and self.global_deformation:
            # line_search.py:133-148: accumulate the accepted deformations (the transfer matrix of the reference maps
            # dofs to vertices; the vertex-blocked deformation vector of the device layer needs none)
            self.problem.global_deformation_function.vector.add_(deformation)
        if self.problem_type == "shape":
            self.form_handler.update_scalar_product()  # line_search.py:247-249
        return deformation


class ArmijoLineSearch(LineSearch):
    """Armijo backtracking (``armijo_line_search.py:36-245``)."""

    def _next_trial(self, current_function_value: float, decrease_measure: float) -> float:
        return self.stepsize / self.beta_armijo

    def _enlarge(self) -> float:
        return self.stepsize * self.beta_armijo

    def _begin(self) -> None:
        pass

    def _trial_done(self, objective_step: float) -> None:
        pass

    def search(self, solver, search_direction: torch.Tensor, has_curvature_info: bool):
        """``armijo_line_search.py:97-202``; returns the accepted deformation or None."""
        self.initialize_stepsize(solver, search_direction, has_curvature_info)
        self._begin()
        while True:
            if self._check_for_nonconvergence(solver):
                self.log.append(("fail_stepsize", self.stepsize))
                return None
            if self.problem_type == "shape":
                self.decrease_measure_w_o_step = self.ova.compute_decrease_measure(search_direction)
            n0 = len(self.ova.decisions)
            self.stepsize = self.ova.update_optimization_variables(search_direction, self.stepsize, self.beta_armijo)
            self.log.extend(self.ova.decisions[n0:])
            current_function_value = solver.objective_value
            objective_step = self._compute_objective_at_new_iterate(current_function_value)
            self._trial_done(objective_step)
            decrease_measure = self._compute_decrease_measure(search_direction)
            self.margins.append(objective_step - (current_function_value + self.epsilon_armijo * decrease_measure))
            if objective_step < current_function_value + self.epsilon_armijo * decrease_measure:
                self.log.append(("armijo_accept", self.stepsize))
                if self.ova.requires_remeshing():
                    # armijo_line_search.py:169-181: remeshing needs the gmsh executable and is outside the hot
                    # path, so `mesh_handler.remesh` is a no-op as with `[Mesh] remesh = False`
                    break
                if solver.iteration == 0:
                    self.armijo_stepsize_initial = self.stepsize
                break
            self.log.append(("armijo_reject", self.stepsize))
            self.stepsize = self._next_trial(current_function_value, decrease_measure)
            self.ova.revert_variable_update()
        solver.stepsize = self.stepsize
        if not has_curvature_info:
            self.stepsize = self._enlarge()
        return self.ova.deformation

    def _compute_objective_at_new_iterate(self, current_function_value: float) -> float:
        """``armijo_line_search.py:225-245``: failed state solve -> J = 2|J| and state revert."""
        self.state_problem.has_solution = False
        try:
            return self.cost_functional.evaluate()
        except (_exceptions.PETScError, _exceptions.NotConvergedError):
            if self.config.getboolean("LineSearch", "fail_if_not_converged"):
                raise
            self.state_problem.revert_to_checkpoint()
            return 2.0 * abs(current_function_value)


class PolynomialLineSearch(ArmijoLineSearch):
    """Armijo acceptance with the rejected steps replaced by the minimiser of a quadratic / cubic model through
    the trial values (``polynomial_line_search.py:36-335``)."""

    def __init__(self, problem) -> None:
        super().__init__(problem)
        self.factor_low = self.config.getfloat("LineSearch", "factor_low")
        self.factor_high = self.config.getfloat("LineSearch", "factor_high")
        self.polynomial_model = self.config.get("LineSearch", "polynomial_model").casefold()
        self.f_vals: collections.deque[float] = collections.deque()
        self.alpha_vals: collections.deque[float] = collections.deque()

    def _begin(self) -> None:
        self.f_vals.clear()
        self.alpha_vals.clear()

    def _trial_done(self, objective_step: float) -> None:
        self.alpha_vals.append(self.stepsize)
        self.f_vals.append(objective_step)

    def _compute_objective_at_new_iterate(self, current_function_value: float) -> float:
        # polynomial_line_search.py:170-172: no fallback for a failed state solve
        self.state_problem.has_solution = False
        return self.cost_functional.evaluate()

    def _enlarge(self) -> float:
        return self.stepsize / self.factor_high  # polynomial_line_search.py:222-223

    def _next_trial(self, f_current: float, decrease_measure: float) -> float:
        """``_compute_polynomial_stepsize`` (``:228-266``) with the two models (``:268-335``)."""
        f, a = self.f_vals, self.alpha_vals
        if len(f) == 1:
            alpha = -(decrease_measure * self.stepsize**2) / (2.0 * (f[0] - f_current - decrease_measure * a[0]))
        elif len(f) == 2:
            r1 = f[1] - f_current - decrease_measure * a[1]
            r0 = f[0] - f_current - decrease_measure * a[0]
            scale = 1.0 / (a[0] ** 2 * a[1] ** 2 * (a[1] - a[0]))
            coeffs = scale * np.array([[a[0] ** 2, -(a[1] ** 2)], [-(a[0] ** 3), a[1] ** 3]]) @ np.array([r1, r0])
            ca, cb = coeffs
            alpha = float((-cb + np.sqrt(cb**2 - 3 * ca * decrease_measure)) / (3 * ca))
        else:
            raise _exceptions.CashocsException("This code should not be reached.")
        if alpha < self.factor_low * a[-1]:
            stepsize = self.factor_low * a[-1]
        elif alpha > self.factor_high * a[-1]:
            stepsize = self.factor_high * a[-1]
        else:
            stepsize = alpha
        if (self.polynomial_model == "quadratic" and len(f) == 1) or (self.polynomial_model == "cubic" and len(f) == 2):
            f.popleft()
            a.popleft()
        return float(stepsize)


class BasicLineSearch(LineSearch):
    """Constant stepsize (``basic_line_search.py:39-172``): the configured stepsize is only reduced by the mesh
    tests of the variable update or when the state equation cannot be solved."""

    def __init__(self, problem) -> None:
        super().__init__(problem)
        self.constant_stepsize = self.stepsize

    def search(self, solver, search_direction: torch.Tensor, has_curvature_info: bool):
        self.stepsize = self.constant_stepsize
        while True:
            n0 = len(self.ova.decisions)
            self.stepsize = self.ova.update_optimization_variables(search_direction, self.stepsize, self.beta_armijo)
            self.log.extend(self.ova.decisions[n0:])
            self.state_problem.has_solution = False
            try:
                self.cost_functional.evaluate()
                break
            except (_exceptions.PETScError, _exceptions.NotConvergedError):
                if self.config.getboolean("LineSearch", "fail_if_not_converged"):
                    raise
                self.state_problem.revert_to_checkpoint()
                self.stepsize /= self.beta_armijo
                self.ova.revert_variable_update()
        solver.stepsize = self.stepsize
        return self.ova.deformation


def create_line_search(problem) -> LineSearch:
    """``[LineSearch] method`` (``cashocs/_optimization/optimization_problem.py``: armijo | polynomial | basic)."""
    method = problem.config.get("LineSearch", "method").casefold()
    if method == "armijo":
        return ArmijoLineSearch(problem)
    if method == "polynomial":
        return PolynomialLineSearch(problem)
    if method == "basic":
        return BasicLineSearch(problem)
    raise _exceptions.ConfigError([f"Key method in section LineSearch has a wrong value: {method}.\n"])
