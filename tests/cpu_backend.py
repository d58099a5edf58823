"""A CPU stand-in for the device layer *below* the optimisation loop, for ``-m "not gpu"`` tests only.

The host optimisation layer (``cashocs_b200/_optimization``: algorithms, line searches, variable abstractions)
only talks to a handful of objects -- state / adjoint / gradient problems, the form handler's scalar product, the
mesh handler, the reduced cost functional.  Here those objects are backed by the numpy oracle, with torch CPU
tensors where the device layer has CUDA tensors, so that the *shipped* host logic can be executed without a GPU
and compared with the oracle's own optimisation loops decision by decision.  Nothing of this is importable from
the product (it lives under ``tests/``); the GPU tests run the same host code on the CUDA layer.
"""

from __future__ import annotations

import types

import numpy as np
import torch

import cashocs_b200 as cb
from cashocs_b200._optimization import line_search as ls
from cashocs_b200._optimization import optimization_algorithms as algs
from cashocs_b200._pde_problems import hessian_problems
from cashocs_b200.io import managers


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().numpy()


class _Function:
    def __init__(self, n: int) -> None:
        self.vector = torch.zeros(n, dtype=torch.float64)


class _Problems:
    """State / adjoint / gradient problems of one oracle problem, with the ``has_solution`` protocol."""

    def __init__(self, owner) -> None:
        self.owner = owner
        self.state = types.SimpleNamespace(has_solution=False, number_of_solves=0, solve=self.solve_state,
                                           revert_to_checkpoint=lambda: None)
        self.adjoint = types.SimpleNamespace(has_solution=False, number_of_solves=0)
        self.gradient = types.SimpleNamespace(has_solution=False, solve=self.solve_gradient)

    def solve_state(self) -> None:
        if not self.state.has_solution:
            self.owner.sync_to_oracle()
            self.owner.oracle.solve_state()
            self.state.has_solution = True
            self.state.number_of_solves += 1

    def solve_gradient(self):
        if not self.gradient.has_solution:
            own = self.owner
            own.sync_to_oracle()
            g = own.oracle.compute_shape_gradient()  # state + adjoint + Riesz projection
            self.state.has_solution = self.adjoint.has_solution = self.gradient.has_solution = True
            self.state.number_of_solves += 1
            self.adjoint.number_of_solves += 1
            own.gradient.vector.copy_(torch.from_numpy(np.asarray(g)))
        return [self.owner.gradient]


class _MeshHandler:
    """``_MeshHandler`` over the oracle's mesh tests (``oracle/pipeline.py`` move_mesh / compute_decreases)."""

    def __init__(self, oracle) -> None:
        self.oracle = oracle
        self.mesh_quality_tol_lower = oracle.cfg.tol_lower
        self.mesh_quality_tol_upper = oracle.cfg.tol_upper

    @property
    def current_mesh_quality(self) -> float:
        return self.oracle.current_mesh_quality

    def move_mesh(self, deformation: torch.Tensor) -> bool:
        return bool(self.oracle.move_mesh(_np(deformation).copy()))

    def revert_transformation(self) -> None:
        self.oracle.revert_transformation()

    def compute_decreases(self, search_direction: torch.Tensor, stepsize: float) -> int:
        return self.oracle.compute_decreases(_np(search_direction), stepsize)


class _OracleHessianProblem(hessian_problems.HessianProblem):
    """The shipped inner CG / CR iterations and active-set splitting, with the Hessian application itself (the
    three device solves) taken from the oracle."""

    def _setup_device(self) -> None:
        pass

    def hessian_application(self, h: torch.Tensor, out: torch.Tensor) -> None:
        self.problem.sync_to_oracle()
        out.copy_(torch.from_numpy(np.asarray(self.problem.oracle.hessian_application(_np(h).copy()))))
        self.no_sensitivity_solves += 2


class HostProblem:
    """Looks like ``ShapeOptimizationProblem`` / ``OptimalControlProblem`` to the host optimisation layer."""

    def __init__(self, oracle, config: cb.Config) -> None:
        self.oracle = oracle
        self.config = config
        self.is_control = bool(getattr(oracle, "is_control", False))
        self.problem_type = "control" if self.is_control else "shape"
        self.mesh = types.SimpleNamespace(dist=None)
        n = oracle.nv if self.is_control else oracle.d * oracle.nv
        self.gradient = _Function(n)
        if not self.is_control:
            self.global_deformation_function = _Function(n)
        p = _Problems(self)
        self.state_problem, self.adjoint_problem, self.gradient_problem = p.state, p.adjoint, p.gradient
        self.form_handler = types.SimpleNamespace(
            scalar_product=lambda a, b: float(oracle.scalar_product(_np(a), _np(b))),
            update_scalar_product=oracle.update_scalar_product)
        self.reduced_cost_functional = types.SimpleNamespace(evaluate=self._evaluate)
        if self.is_control:
            self.controls = [_Function(n)]
            self.controls[0].vector.copy_(torch.from_numpy(oracle.u))
            self.mesh_handler = None
            bounds = [types.SimpleNamespace(vector=torch.from_numpy(b.copy())) for b in (oracle.u_a, oracle.u_b)]
            self.box_constraints = ls.BoxConstraints(self.controls[0], bounds)
            self.optimization_variable_abstractions = ls.ControlVariableAbstractions(self)
        else:
            self.mesh_handler = _MeshHandler(oracle)
            self.optimization_variable_abstractions = ls.ShapeVariableAbstractions(self)
        self.solver = None

    def sync_to_oracle(self) -> None:
        if self.is_control:
            self.oracle.u = _np(self.controls[0].vector).copy()

    def _evaluate(self) -> float:
        self.state_problem.solve()
        return float(self.oracle.cost())

    def solve(self, algorithm: str, rtol: float, atol: float, max_iter: int) -> None:
        """The body of ``ShapeOptimizationProblem.solve`` / ``OptimalControlProblem.solve``."""
        from cashocs_b200._optimization.shape_optimization_problem import _algorithm_name

        algorithm = _algorithm_name(algorithm)
        for key, val in (("rtol", rtol), ("atol", atol), ("max_iter", max_iter), ("algorithm", algorithm)):
            self.config.set("OptimizationRoutine", key, str(val))
        self.output_manager = managers.OutputManager(self)
        search = ls.create_line_search(self)
        cls = {"gradient_descent": algs.GradientDescentMethod, "bfgs": algs.LBFGSMethod,
               "conjugate_gradient": algs.NonlinearCGMethod, "newton": algs.NewtonMethod}[algorithm]
        if algorithm == "newton":
            self.hessian_problem = _OracleHessianProblem(self)
        self.solver = cls(self, search)
        self.solver.run()
        self.solver.post_processing()
