"""Base class of the PDE problems (``cashocs/_pde_problems/pde_problem.py:33-66``)."""

from __future__ import annotations

import abc

from cashocs_b200._utils import linalg


class PDEProblem(abc.ABC):
    """A PDE problem with a cached solution (``has_solution``) and its own ``LinearSolver``."""

    def __init__(self, linear_solver: linalg.LinearSolver | None = None) -> None:
        self.has_solution = False
        self.linear_solver = linear_solver if linear_solver is not None else linalg.LinearSolver()

    @abc.abstractmethod
    def solve(self):
        """Solve the problem (no-op while ``has_solution``)."""
