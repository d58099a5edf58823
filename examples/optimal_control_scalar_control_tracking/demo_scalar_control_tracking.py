"""The reference's documented demo
``demos/documented/optimal_control/scalar_control_tracking/demo_scalar_control_tracking.py`` on the B200 path: find a
control such that the squared L2 norm of the state hits a prescribed value, ``J = 1/2 (int y^2 dx - 1)^2``.

Reference (FEniCS / UFL)                                            | here
--------------------------------------------------------------------+-----------------------------------------------
u.vector()[:] = 1.0                                                 | u.vector.fill_(1.0)
integrand = y*y*dx; tracking_goal = 1.0                             | integrand = cb.IntegralFunctional(c_track=2.0, desired_state=zero)
J_tracking = cashocs.ScalarTrackingFunctional(integrand, goal)      | J_tracking = cb.ScalarTrackingFunctional(integrand, tracking_goal)
ocp = cashocs.OptimalControlProblem(e, bcs, J_tracking, y, u, p,    | ocp = cb.OptimalControlProblem(e, bcs, J_tracking, y, u, p, config=config)
                                    config=config)                  |
ocp.solve()                                                         | ocp.solve()
print("L2-Norm of y squared: ...")                                  | the same line

(``y*y*dx = 2 * 1/2 int (y - 0)^2``: the quadratic state term of the form family with a zero desired state.)

Run on a GPU box:  python examples/optimal_control_scalar_control_tracking/demo_scalar_control_tracking.py
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cashocs_b200 as cb  # noqa: E402


def main(verbose: bool = True):
    config = cb.load_config(os.path.join(HERE, "config.ini"))
    mesh = cb.regular_mesh(25)
    y, p, u, zero = (cb.Function(mesh) for _ in range(4))
    u.vector.fill_(1.0)
    e = cb.PoissonForm(source=u)
    bcs = cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4])
    integrand = cb.IntegralFunctional(c_track=2.0, desired_state=zero)
    tracking_goal = 1.0
    J_tracking = cb.ScalarTrackingFunctional(integrand, tracking_goal)
    ocp = cb.OptimalControlProblem(e, bcs, J_tracking, y, u, p, config=config)
    ocp.solve()
    ocp.compute_state_variables()
    norm = ocp.cost.tracking_values()[0]
    if verbose:
        for h in ocp.solver.history:
            print(f"{h['iteration']:4d}   {h['objective_value']: .6e}   {h['relative_norm']:.3e}   {h['stepsize']:.3e}")
        print("L2-Norm of y squared: " + format(norm, ".3e"))
    return ocp, norm


if __name__ == "__main__":
    main()
