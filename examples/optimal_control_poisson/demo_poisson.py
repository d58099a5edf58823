"""The reference's documented demo ``demos/documented/optimal_control/poisson/demo_poisson.py`` (the 2-D sibling of
BASELINE configs[3]) on the B200 path, line by line -- optionally with the box constraints of the ``box_constraints``
demo (``--box 0.0 1e2``).

Reference (FEniCS / UFL)                                   | here (form descriptors, device tensors)
-----------------------------------------------------------+---------------------------------------------------
config = cashocs.load_config("config.ini")                 | config = cb.load_config("config.ini")
mesh, ..., boundaries, dx, ... = cashocs.regular_mesh(25)  | mesh = cb.regular_mesh(25)
V = FunctionSpace(mesh, "CG", 1); y, p, u = Function(V)    | y, p, u = cb.Function(mesh) x 3
e = inner(grad(y), grad(p))*dx - u*p*dx                    | e = cb.PoissonForm(source=u)
bcs = create_dirichlet_bcs(V, Constant(0), bdry, [1,2,3,4])| bcs = cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4])
y_d = Expression("sin(2 pi x0) sin(2 pi x1)", degree=1)    | y_d = P1 interpolant of the same expression
J = IntegralFunctional(0.5 (y-y_d)^2 dx + 0.5 alpha u^2 dx)| J = cb.IntegralFunctional(c_track=1, desired_state=y_d,
                                                           |                           c_control=alpha)
ocp = cashocs.OptimalControlProblem(e, bcs, J, y, u, p,    | ocp = cb.OptimalControlProblem(e, bcs, J, y, u, p,
                                    config=config)         |                                config=config)
ocp.solve()                                                | ocp.solve()

Run on a GPU box:  python examples/optimal_control_poisson/demo_poisson.py [--box LOWER UPPER]
"""

import argparse
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import torch  # noqa: E402

import cashocs_b200 as cb  # noqa: E402


def main(box=None, n: int = 25):
    config = cb.load_config(os.path.join(HERE, "config.ini"))
    mesh = cb.regular_mesh(n)
    y, p, u, y_d = (cb.Function(mesh) for _ in range(4))
    e = cb.PoissonForm(source=u)
    bcs = cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4])
    x = mesh.coords
    y_d.vector.copy_(torch.sin(2 * math.pi * x[:, 0]) * torch.sin(2 * math.pi * x[:, 1]))
    alpha = 1e-6
    J = cb.IntegralFunctional(c_track=1.0, desired_state=y_d, c_control=alpha)
    ocp = cb.OptimalControlProblem(e, bcs, J, y, u, p, config=config, control_constraints=box)
    ocp.solve()
    return ocp


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--box", type=float, nargs=2, default=None, metavar=("LOWER", "UPPER"))
    args = ap.parse_args()
    ocp = main(args.box)
    print(f"iterations {ocp.solver.iteration}, J = {ocp.solver.objective_value:.6e}, "
          f"relative gradient norm {ocp.solver.relative_norm:.3e}, "
          f"u in [{float(ocp.controls[0].vector.min()):.3f}, {float(ocp.controls[0].vector.max()):.3f}]")
