"""The reference's documented demo ``demos/documented/shape_optimization/p_laplacian/demo_p_laplacian.py`` on the
B200 path: the same problem as ``shape_poisson`` with the gradient deformation from a p-Laplace projection
(``[ShapeGradient] use_p_laplacian = True``, ``p_laplacian_power = 10``) and the demo's own ``config.ini`` keys.

Reference (FEniCS / UFL)                                   | here
-----------------------------------------------------------+---------------------------------------------------
config = cashocs.load_config("./config.ini")               | config = cb.load_config("./config.ini")
mesh, ... = cashocs.import_mesh("./mesh/mesh.xdmf")        | mesh = cb.import_mesh(unit_circle.npz)   (same gmsh mesh)
e = inner(grad(u), grad(p))*dx - f*p*dx                    | e = cb.PoissonForm(source=f)
J = cashocs.IntegralFunctional(u*dx)                       | J = cb.IntegralFunctional(c_state=1.0)
sop = cashocs.ShapeOptimizationProblem(e, bcs, J, u, p,    | sop = cb.ShapeOptimizationProblem(e, bcs, J, u, p, config=config)
                                       boundaries, config) |
sop.solve()                                                | sop.solve()

Run on a GPU box:  python examples/shape_p_laplacian/demo_p_laplacian.py
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cashocs_b200 as cb  # noqa: E402


def main(verbose: bool = True):
    config = cb.load_config(os.path.join(HERE, "config.ini"))
    mesh = cb.import_mesh(os.path.join(HERE, "..", "..", "tests", "golden", "unit_circle.npz"))
    u, p = cb.Function(mesh), cb.Function(mesh)
    x = cb.spatial_coordinate(2)
    f = 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1
    e = cb.PoissonForm(source=f)
    bcs = cb.create_dirichlet_bcs(0.0, [1])
    J = cb.IntegralFunctional(c_state=1.0)
    sop = cb.ShapeOptimizationProblem(e, bcs, J, u, p, config=config)
    sop.solve()
    if verbose:
        for h in sop.solver.history:
            print(f"{h['iteration']:4d}   {h['objective_value']: .6e}   {h['relative_norm']:.3e}   {h['mesh_quality']:.3f}   {h['stepsize']:.3e}")
        print("Newton steps of the last projection, p = 2 .. 10:", sop.gradient_problem.p_laplace_projector.newton_iterations)
    return sop


if __name__ == "__main__":
    main()
