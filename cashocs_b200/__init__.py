"""cashocs_b200 -- B200-native shape-gradient pipeline behind the cashocs API surface.

Hot path only: P1 assembly into CSR / block-CSR, preconditioned Krylov
solves (state, adjoint, Lame-mu, elasticity Riesz projection), mesh quality, a-priori /
a-posteriori mesh tests and ALE mesh moves -- hand-written CUDA for sm_100a behind a C-ABI
(``include/cashocs_b200.h``), host logic in Python mirroring the reference's classes.
"""

from cashocs_b200 import _exceptions
from cashocs_b200 import log
from cashocs_b200._exceptions import CashocsException, ConfigError, InputError, NotConvergedError, PETScKSPError
from cashocs_b200._forms.expressions import Polynomial, spatial_coordinate
from cashocs_b200._forms.forms import (DirichletBC, Function, IntegralFunctional, PoissonForm, ScalarTrackingFunctional,
                                       create_dirichlet_bcs)
from cashocs_b200._optimization.optimal_control_problem import OptimalControlProblem
from cashocs_b200._optimization.shape_optimization_problem import ShapeOptimizationProblem
from cashocs_b200._utils.linalg import LinearSolver, Matrix
from cashocs_b200.geometry.mesh import Mesh, import_mesh, mesh_from_arrays, regular_box_mesh, regular_mesh
from cashocs_b200.geometry.mesh_testing import APrioriMeshTester, IntersectionTester
from cashocs_b200.geometry.deformations import DeformationHandler
from cashocs_b200.geometry.quality import compute_mesh_quality
from cashocs_b200.io.config import Config, load_config
from cashocs_b200 import nonlinear_solvers
from cashocs_b200.nonlinear_solvers import linear_solve, newton_solve

__version__ = "0.1.0"

__all__ = [
    "APrioriMeshTester", "CashocsException", "Config", "ConfigError", "DeformationHandler", "DirichletBC",
    "Function", "InputError", "IntegralFunctional", "IntersectionTester", "LinearSolver", "Matrix", "Mesh",
    "NotConvergedError", "OptimalControlProblem", "PETScKSPError", "PoissonForm", "Polynomial", "ScalarTrackingFunctional",
    "ShapeOptimizationProblem", "compute_mesh_quality", "create_dirichlet_bcs", "import_mesh", "linear_solve", "load_config", "log", "newton_solve", "nonlinear_solvers",
    "mesh_from_arrays", "regular_box_mesh", "regular_mesh", "spatial_coordinate",
]
