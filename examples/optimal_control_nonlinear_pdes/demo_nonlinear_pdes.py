"""The reference's documented demo ``demos/documented/optimal_control/nonlinear_pdes/demo_nonlinear_pdes.py`` on the
B200 path: distributed control of the semilinear equation ``-Laplace y + c y^3 = u`` with the demo's own
``config.ini`` (``is_linear`` keeps its default ``False``: the state is solved by Newton's method on the device).

Reference (FEniCS / UFL)                                            | here
--------------------------------------------------------------------+------------------------------------------------
mesh, ... = cashocs.regular_mesh(25)                                | mesh = cb.regular_mesh(25)
c = Constant(1e2)                                                   | c = 1e2
e = inner(grad(y), grad(p))*dx + c*pow(y, 3)*p*dx - u*p*dx          | e = cb.PoissonForm(source=u, cubic=c)
bcs = cashocs.create_dirichlet_bcs(V, Constant(0), boundaries, 1-4) | bcs = cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4])
y_d = Expression("sin(2*pi*x[0])*sin(2*pi*x[1])", degree=1)         | y_d: the same function interpolated to P1
J = IntegralFunctional(0.5*(y-y_d)^2*dx + 0.5*alpha*u*u*dx)         | J = cb.IntegralFunctional(c_track=1.0, desired_state=y_d, c_control=alpha)
ocp = cashocs.OptimalControlProblem(e, bcs, J, y, u, p, config)     | ocp = cb.OptimalControlProblem(e, bcs, J, y, u, p, config=config)
ocp.solve()                                                         | ocp.solve()

Run on a GPU box:  python examples/optimal_control_nonlinear_pdes/demo_nonlinear_pdes.py
"""

import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import torch  # noqa: E402

import cashocs_b200 as cb  # noqa: E402


def main(verbose: bool = True):
    config = cb.load_config(os.path.join(HERE, "config.ini"))
    mesh = cb.regular_mesh(25)
    y, p, u, y_d = (cb.Function(mesh) for _ in range(4))
    c = 1e2
    e = cb.PoissonForm(source=u, cubic=c)
    bcs = cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4])
    xy = mesh.coords
    y_d.vector.copy_(torch.sin(2 * math.pi * xy[:, 0]) * torch.sin(2 * math.pi * xy[:, 1]))
    alpha = 1e-6
    J = cb.IntegralFunctional(c_track=1.0, desired_state=y_d, c_control=alpha)
    ocp = cb.OptimalControlProblem(e, bcs, J, y, u, p, config=config)
    ocp.solve()
    if verbose:
        for h in ocp.solver.history:
            print(f"{h['iteration']:4d}   {h['objective_value']: .6e}   {h['relative_norm']:.3e}   {h['stepsize']:.3e}")
    return ocp


if __name__ == "__main__":
    main()
