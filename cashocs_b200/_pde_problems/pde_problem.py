"""Base class of the PDE problems (``cashocs/_pde_problems/pde_problem.py:33-66``)."""

from __future__ import annotations

import abc

from cashocs_b200._utils import linalg


class PDEProblem(abc.ABC):
    """A PDE problem with a cached solution (``has_solution``) and its own ``LinearSolver``."""

    # user callbacks (``optimization_algorithms/callback.py``): ``pre_callback`` runs before every solve of the state
    # system (``state_problem.py:152``), ``post_callback`` after every gradient computation
    # (``shape_gradient_problem.py:166``, ``control_gradient_problem.py:123``)
    pre_callback = None
    post_callback = None

    def __init__(self, linear_solver: linalg.LinearSolver | None = None) -> None:
        self.has_solution = False
        self.linear_solver = linear_solver if linear_solver is not None else linalg.LinearSolver()

    @abc.abstractmethod
    def solve(self):
        """Solve the problem (no-op while ``has_solution``)."""
