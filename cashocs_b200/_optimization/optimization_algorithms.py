"""Optimisation algorithms driving the hot path (``cashocs/_optimization/optimization_algorithms/``).

Gradient descent (``gradient_descent.py:28-60``), L-BFGS with the two-loop recursion
(``l_bfgs.py:91-300``), the nonlinear CG methods (``ncg.py``) and the truncated Newton method (``newton.py``); every inner product is ``form_handler.scalar_product`` (= fused SpMV + dot
on the device), vectors never leave HBM.  Convergence test and ascent check follow
``optimization_algorithm.py:281-331``.
"""

from __future__ import annotations

import collections

import torch


class OptimizationAlgorithm:
    def __init__(self, problem, line_search) -> None:
        self.problem = problem
        self.line_search = line_search
        cfg = problem.config
        self.config = cfg
        self.form_handler = problem.form_handler
        self.gradient_problem = problem.gradient_problem
        self.adjoint_problem = problem.adjoint_problem
        self.state_problem = problem.state_problem
        self.cost_functional = problem.reduced_cost_functional
        self.ova = problem.optimization_variable_abstractions
        self.gradient = problem.gradient.vector
        self.search_direction = torch.zeros_like(self.gradient)
        self.rtol = cfg.getfloat("OptimizationRoutine", "rtol")
        self.atol = cfg.getfloat("OptimizationRoutine", "atol")
        self.max_iter = cfg.getint("OptimizationRoutine", "max_iter")
        self.iteration = 0
        self.objective_value = 1.0
        self.gradient_norm = 1.0
        self.gradient_norm_initial = 1.0
        self.relative_norm = 1.0
        self.stepsize = 1.0
        self.converged = False
        self.line_search_broken = False
        self.converged_reason = 0
        self.has_curvature_info = False
        self.is_restarted = False
        self.history: list[dict] = []
        self.output_manager = getattr(problem, "output_manager", None)

    def compute_gradient(self) -> None:
        self.adjoint_problem.has_solution = False
        self.gradient_problem.has_solution = False
        self.gradient_problem.solve()

    def evaluate_cost_functional(self) -> None:
        self.objective_value = self.cost_functional.evaluate()

    def convergence_test(self) -> bool:
        if self.iteration == 0:
            self.gradient_norm_initial = self.gradient_norm
        self.relative_norm = self.gradient_norm / self.gradient_norm_initial if self.gradient_norm > 0.0 else 0.0
        if self.gradient_norm <= self.atol + self.rtol * self.gradient_norm_initial:
            self.objective_value = self.cost_functional.evaluate()
            self.converged = True
            return True
        return False

    def check_for_ascent(self) -> None:
        if self.form_handler.scalar_product(self.gradient, self.search_direction) >= 0:
            torch.neg(self.gradient, out=self.search_direction)
            self.has_curvature_info = False

    def nonconvergence(self) -> bool:
        """``optimization_algorithm.py:221-236``: -1 iteration budget, -2 Armijo failure."""
        if self.iteration >= self.max_iter:
            self.converged_reason = -1
        if self.line_search_broken:
            self.converged_reason = -2
        return self.converged_reason < 0

    def post_processing(self) -> None:
        """``optimization_algorithm.py:250-290``: a run that did not converge raises ``NotConvergedError``
        ("Maximum number of iterations exceeded." / "Armijo rule failed.") unless ``soft_exit`` is set."""
        om = self.output_manager
        if self.converged:
            if om is not None:  # the converged iterate was already written by the run loop
                om.post_process(self.optimization_state())
                om.output_summary(self.optimization_state())
            return
        from cashocs_b200 import _exceptions
        from cashocs_b200 import log

        self.objective_value = self.cost_functional.evaluate()
        self.gradient_norm = float("nan")
        self.relative_norm = float("nan")
        if self.converged_reason == -2:
            self.iteration -= 1
            message = "Armijo rule failed."
        else:
            message = "Maximum number of iterations exceeded."
            self._record()
        if om is not None:
            om.post_process(self.optimization_state())
        if self.config.getboolean("OptimizationRoutine", "soft_exit"):
            log.warning(message)
        else:
            raise _exceptions.NotConvergedError("Optimization Algorithm", message)

    def optimization_state(self) -> dict:
        """The ``optimization_state`` dictionary of the reference's parameter database
        (``optimization_algorithm.py:126-193``, ``state_problem.py:135``, ``adjoint_problem.py:121``,
        ``mesh_handler.py:252``)."""
        state = dict(stepsize=self.stepsize, no_state_solves=self.state_problem.number_of_solves,
                     no_adjoint_solves=self.adjoint_problem.number_of_solves)
        if self.problem.mesh_handler is not None:
            state["mesh_quality"] = float(self.problem.mesh_handler.current_mesh_quality)
        state.update(iteration=self.iteration, objective_value=float(self.objective_value),
                     gradient_norm=float(self.gradient_norm), gradient_norm_initial=float(self.gradient_norm_initial),
                     relative_norm=float(self.relative_norm))
        return state

    def _record(self) -> None:
        """``OptimizationAlgorithm.output`` (``optimization_algorithm.py:210-212``)."""
        self.history.append(
            dict(iteration=self.iteration, objective_value=self.objective_value, relative_norm=self.relative_norm,
                 gradient_norm=self.gradient_norm, mesh_quality=getattr(self.problem.mesh_handler, "current_mesh_quality", None),
                 stepsize=self.stepsize)
        )
        if self.output_manager is not None:
            self.output_manager.output(self.optimization_state())


class GradientDescentMethod(OptimizationAlgorithm):
    def run(self) -> None:
        while True:
            self.compute_gradient()
            self.gradient_norm = self.ova.compute_gradient_norm()
            if self.convergence_test():
                self._record()
                break
            self.evaluate_cost_functional()
            self._record()
            torch.neg(self.gradient, out=self.search_direction)
            self.line_search.perform(self, self.search_direction, self.has_curvature_info)
            self.iteration += 1
            if self.nonconvergence():
                break


class NewtonMethod(OptimizationAlgorithm):
    """A truncated Newton method (``optimization_algorithms/newton.py:30-101``): the search direction solves
    ``J''(u) d = -J'(u)`` inexactly with the CG / CR iterations of the Hessian problem."""

    def __init__(self, problem, line_search) -> None:
        super().__init__(problem, line_search)
        self.hessian_problem = problem.hessian_problem
        self.stepsize = 1.0

    def compute_search_direction(self) -> None:
        self.search_direction.copy_(self.hessian_problem.newton_solve())
        self.has_curvature_info = True

    def run(self) -> None:
        while True:
            self.compute_gradient()
            self.gradient_norm = self.ova.compute_gradient_norm()
            if self.convergence_test():
                self._record()
                break
            self.compute_search_direction()
            self.check_for_ascent()
            self.evaluate_cost_functional()
            self._record()
            self.line_search.perform(self, self.search_direction, self.has_curvature_info)
            if self.line_search_broken and self.has_curvature_info:
                # newton.py:80-92: the Newton direction failed, try the negative gradient once
                torch.neg(self.gradient, out=self.search_direction)
                self.has_curvature_info = False
                self.line_search_broken = False
                self.line_search.perform(self, self.search_direction, self.has_curvature_info)
            self.iteration += 1
            if self.nonconvergence():
                break


class LBFGSMethod(OptimizationAlgorithm):
    def __init__(self, problem, line_search) -> None:
        super().__init__(problem, line_search)
        self.bfgs_memory_size = self.config.getint("AlgoLBFGS", "bfgs_memory_size")
        self.use_bfgs_scaling = self.config.getboolean("AlgoLBFGS", "use_bfgs_scaling")
        self.bfgs_periodic_restart = self.config.getint("AlgoLBFGS", "bfgs_periodic_restart")
        self.periodic_its = 0
        self.damped = self.config.getboolean("AlgoLBFGS", "damped")
        self.history_s: collections.deque = collections.deque()
        self.history_y: collections.deque = collections.deque()
        self.history_rho: collections.deque = collections.deque()
        self.gradient_prev = torch.zeros_like(self.gradient)

    def compute_search_direction(self, grad: torch.Tensor | None = None) -> None:
        """``d = -H grad`` by the two-loop recursion (``l_bfgs.py:182-232``); ``grad`` defaults to the gradient."""
        sp = self.form_handler.scalar_product
        d = self.search_direction
        grad = self.gradient if grad is None else grad
        box = self.ova.has_box_constraints  # l_bfgs.py:182-222: recursion on the inactive set + active gradient
        if self.bfgs_memory_size > 0 and len(self.history_s) > 0:
            d.copy_(grad)
            if box:
                self.ova.restrict_to_inactive_set(d, d)
            alphas = []
            for s, y, rho in zip(self.history_s, self.history_y, self.history_rho):
                a = rho * sp(s, d)
                alphas.append(a)
                d.add_(y, alpha=-a)
            if self.use_bfgs_scaling and self.iteration > 0:
                d.mul_(sp(self.history_y[0], self.history_s[0]) / sp(self.history_y[0], self.history_y[0]))
            if box:
                self.ova.restrict_to_inactive_set(d, d)
            m = len(self.history_s)
            for i in range(m):
                b = self.history_rho[-1 - i] * sp(self.history_y[-1 - i], d)
                d.add_(self.history_s[-1 - i], alpha=alphas[-1 - i] - b)
            if box:
                self.ova.restrict_to_inactive_set(d, d)
                d.add_(self.ova.restrict_to_active_set(self.gradient, torch.empty_like(d)))
            d.neg_()
        else:
            torch.neg(grad, out=d)

    def _apply_inverse_hessian(self, y: torch.Tensor) -> torch.Tensor:
        """``H y`` with the current memory (``_compute_approximate_inverse_hessian_application``, ``l_bfgs.py:331-354``)."""
        keep = self.search_direction.clone()
        self.compute_search_direction(y)
        out = -self.search_direction
        self.search_direction.copy_(keep)
        return out

    def update_hessian_approximation(self) -> None:
        if self.bfgs_memory_size <= 0:
            return
        y_k = self.gradient - self.gradient_prev
        s_k = self.stepsize * self.search_direction
        if self.ova.has_box_constraints:
            self.ova.restrict_to_inactive_set(y_k, y_k)
            self.ova.restrict_to_inactive_set(s_k, s_k)
        curvature = self.form_handler.scalar_product(y_k, s_k)
        if self.damped:
            # Powell's damping (l_bfgs.py:284-304): s is blended with H y so that the pair keeps positive curvature
            hy = self._apply_inverse_hessian(y_k)
            gamma = self.form_handler.scalar_product(y_k, hy)
            phi = 1.0 if curvature >= 0.2 * gamma else (0.8 * gamma) / (gamma - curvature)
            r_k = phi * s_k + (1.0 - phi) * hy
            self.history_s.appendleft(r_k)
            self.history_y.appendleft(y_k)
            self.has_curvature_info = True
            self.history_rho.appendleft(1.0 / self.form_handler.scalar_product(y_k, r_k))
        else:
            self.history_y.appendleft(y_k)
            self.history_s.appendleft(s_k)
            if curvature <= 0.0:
                self.has_curvature_info = False
                self.history_s.clear(), self.history_y.clear(), self.history_rho.clear()
            else:
                self.has_curvature_info = True
                self.history_rho.appendleft(1.0 / curvature)
        if len(self.history_s) > self.bfgs_memory_size:
            self.history_s.pop(), self.history_y.pop(), self.history_rho.pop()

    def _compute_active_sets(self) -> None:
        if self.ova.has_box_constraints:
            self.ova.compute_active_sets()

    def check_restart(self) -> None:
        """``l_bfgs.py:312-334``: every ``bfgs_periodic_restart`` iterations the memory is dropped and the next
        step is a gradient step from the configured stepsize."""
        if self.bfgs_periodic_restart > 0:
            if self.periodic_its < self.bfgs_periodic_restart:
                self.periodic_its += 1
                self.is_restarted = False
            else:
                torch.neg(self.gradient, out=self.search_direction)
                self.periodic_its = 0
                self.has_curvature_info = False
                self.is_restarted = True
                self.history_s.clear(), self.history_y.clear(), self.history_rho.clear()

    def run(self) -> None:
        self.periodic_its = 0
        self.compute_gradient()
        self._compute_active_sets()
        self.gradient_norm = self.ova.compute_gradient_norm()
        self.converged = self.convergence_test()
        while not self.converged:
            self.compute_search_direction()
            self.check_restart()
            self.check_for_ascent()
            self.evaluate_cost_functional()
            self._record()
            self.line_search.perform(self, self.search_direction, self.has_curvature_info)
            self.iteration += 1
            if self.nonconvergence():
                break
            self.gradient_prev.copy_(self.gradient)
            self.compute_gradient()
            self._compute_active_sets()
            self.gradient_norm = self.ova.compute_gradient_norm()
            self.relative_norm = self.gradient_norm / self.gradient_norm_initial
            if self.convergence_test():
                break
            self.update_hessian_approximation()
        self._record()


class NonlinearCGMethod(OptimizationAlgorithm):
    """Nonlinear CG methods (``cashocs/_optimization/optimization_algorithms/ncg.py``): Fletcher-Reeves,
    Polak-Ribiere, Hestenes-Stiefel, Dai-Yuan, Hager-Zhang (``[AlgoCG] cg_method``), periodic and relative
    restarts (``:188-212``)."""

    def __init__(self, problem, line_search) -> None:
        super().__init__(problem, line_search)
        cfg = self.config
        self.gradient_prev = torch.zeros_like(self.gradient)
        self.difference = torch.zeros_like(self.gradient)
        self.cg_method = cfg.get("AlgoCG", "cg_method")
        self.cg_periodic_restart = cfg.getboolean("AlgoCG", "cg_periodic_restart")
        self.cg_periodic_its = cfg.getint("AlgoCG", "cg_periodic_its")
        self.cg_relative_restart = cfg.getboolean("AlgoCG", "cg_relative_restart")
        self.cg_restart_tol = cfg.getfloat("AlgoCG", "cg_restart_tol")
        self.memory = 0
        self.beta = 0.0

    def compute_beta(self) -> None:
        """``ncg.py:105-182``."""
        if self.iteration == 0:
            self.beta = 0.0
            return
        sp = self.form_handler.scalar_product
        g, gp, d, y = self.gradient, self.gradient_prev, self.search_direction, self.difference
        m = self.cg_method.casefold()
        if m != "fr":
            torch.sub(g, gp, out=y)
        if m == "fr":
            self.beta = sp(g, g) / sp(gp, gp)
        elif m == "pr":
            self.beta = sp(g, y) / sp(gp, gp)
        elif m == "hs":
            self.beta = sp(g, y) / sp(y, d)
        elif m == "dy":
            self.beta = sp(g, g) / sp(d, y)
        elif m == "hz":
            dy = sp(d, y)
            y2 = sp(y, y)
            y.add_(d, alpha=-2.0 * y2 / dy)
            self.beta = sp(y, g) / dy
        else:
            from cashocs_b200 import _exceptions

            raise _exceptions.ConfigError([f"[AlgoCG] cg_method = {self.cg_method!r} is not one of FR, PR, HS, DY, HZ"])

    def restart(self) -> None:
        """``ncg.py:188-212``."""
        if self.cg_periodic_restart:
            if self.memory < self.cg_periodic_its:
                self.memory += 1
            else:
                torch.neg(self.gradient, out=self.search_direction)
                self.memory = 0
        if self.cg_relative_restart:
            gg = abs(self.form_handler.scalar_product(self.gradient, self.gradient_prev))
            if gg / self.gradient_norm**2 >= self.cg_restart_tol:
                torch.neg(self.gradient, out=self.search_direction)

    def run(self) -> None:
        self.memory = 0
        while True:
            self.gradient_prev.copy_(self.gradient)
            self.compute_gradient()
            self.gradient_norm = self.ova.compute_gradient_norm()
            self.compute_beta()
            if self.convergence_test():
                self._record()
                break
            # d = -g + beta d  (aypx)
            self.search_direction.mul_(self.beta).sub_(self.gradient)
            self.restart()
            if self.ova.has_box_constraints:
                self.ova.project_ncg_search_direction(self.search_direction)
            self.check_for_ascent()
            self.evaluate_cost_functional()
            self._record()
            self.line_search.perform(self, self.search_direction, self.has_curvature_info)
            self.iteration += 1
            if self.nonconvergence():
                break
