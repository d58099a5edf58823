"""The reference's documented demo ``demos/documented/optimal_control/iterative_solvers/demo_iterative_solvers.py`` on
the B200 path: the ``ksp_options`` / ``adjoint_ksp_options`` dictionaries are passed through unchanged -- cg + hypre
(boomeramg; here the device's smoothed-aggregation AMG) for the state, minres + jacobi for the adjoint system.

Reference (FEniCS / UFL)                                         | here
-----------------------------------------------------------------+-------------------------------------------------
mesh, ... = cashocs.regular_mesh(50)                             | mesh = cb.regular_mesh(50)
ksp_options = {"ksp_type": "cg", "pc_type": "hypre", ...}        | the same dict
adjoint_ksp_options = {"ksp_type": "minres", "pc_type": "jacobi",| the same dict
                       "ksp_rtol": 1e-6, "ksp_atol": 1e-15}      |
ocp = cashocs.OptimalControlProblem(e, bcs, J, y, u, p, config,  | ocp = cb.OptimalControlProblem(e, bcs, J, y, u, p, config=config,
      ksp_options=ksp_options,                                   |       ksp_options=ksp_options,
      adjoint_ksp_options=adjoint_ksp_options)                   |       adjoint_ksp_options=adjoint_ksp_options)
ocp.solve()                                                      | ocp.solve()

Run on a GPU box:  python examples/optimal_control_iterative_solvers/demo_iterative_solvers.py
"""

import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import torch  # noqa: E402

import cashocs_b200 as cb  # noqa: E402

KSP_OPTIONS = {
    "ksp_type": "cg",
    "pc_type": "hypre",
    "pc_hypre_type": "boomeramg",
    "ksp_rtol": 1e-10,
    "ksp_atol": 1e-13,
    "ksp_max_it": 100,
}
ADJOINT_KSP_OPTIONS = {
    "ksp_type": "minres",
    "pc_type": "jacobi",
    "ksp_rtol": 1e-6,
    "ksp_atol": 1e-15,
}


def main(verbose: bool = True):
    config = cb.load_config(os.path.join(HERE, "config.ini"))
    mesh = cb.regular_mesh(50)
    y, p, u, y_d = (cb.Function(mesh) for _ in range(4))
    e = cb.PoissonForm(source=u)
    bcs = cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4])
    xy = mesh.coords
    y_d.vector.copy_(torch.sin(2 * math.pi * xy[:, 0]) * torch.sin(2 * math.pi * xy[:, 1]))
    alpha = 1e-6
    J = cb.IntegralFunctional(c_track=1.0, desired_state=y_d, c_control=alpha)
    ocp = cb.OptimalControlProblem(e, bcs, J, y, u, p, config=config, ksp_options=dict(KSP_OPTIONS),
                                   adjoint_ksp_options=dict(ADJOINT_KSP_OPTIONS))
    ocp.solve()
    if verbose:
        for h in ocp.solver.history:
            print(f"{h['iteration']:4d}   {h['objective_value']: .6e}   {h['relative_norm']:.3e}   {h['stepsize']:.3e}")
        print("Krylov iterations of the last state / adjoint solve:", ocp.state_problem.iterations, ocp.adjoint_problem.iterations)
    return ocp


if __name__ == "__main__":
    main()
