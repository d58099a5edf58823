"""BASELINE configs[0] on the B200 path: the reference's documented demo
``demos/documented/shape_optimization/shape_poisson/demo_shape_poisson.py`` line by line.

Reference (FEniCS / UFL)                                   | here (form descriptors, device tensors)
-----------------------------------------------------------+---------------------------------------------------
config = cashocs.load_config("./config.ini")               | config = cb.load_config("./config.ini")
mesh, ..., boundaries, dx, ... = cashocs.import_mesh(xdmf) | mesh = cb.import_mesh(unit_circle.npz)   (same gmsh mesh)
V = FunctionSpace(mesh, "CG", 1); u, p = Function(V) x2    | u, p = cb.Function(mesh), cb.Function(mesh)
x = SpatialCoordinate(mesh); f = 2.5*(x0+0.4-x1^2)^2+...   | x = cb.spatial_coordinate(2); f = the same expression
e = inner(grad(u), grad(p))*dx - f*p*dx                    | e = cb.PoissonForm(source=f)
bcs = DirichletBC(V, Constant(0), boundaries, 1)           | bcs = cb.create_dirichlet_bcs(0.0, [1])
J = cashocs.IntegralFunctional(u*dx)                       | J = cb.IntegralFunctional(c_state=1.0)
sop = cashocs.ShapeOptimizationProblem(e, bcs, J, u, p,    | sop = cb.ShapeOptimizationProblem(e, bcs, J, u, p,
                                       boundaries, config) |                                    config=config)
sop.solve()                                                | sop.solve()

Run on a GPU box:  python examples/shape_poisson/demo_shape_poisson.py
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cashocs_b200 as cb  # noqa: E402


def main(verbose: bool = True):
    config = cb.load_config(os.path.join(HERE, "config.ini"))
    # the reference's mesh/mesh.xdmf (713 vertices / 1340 triangles), converted by tests/golden/make_fixtures.py
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
        print("iter   cost function   rel. grad. norm   mesh quality   step size")
        for h in sop.solver.history:
            print(f"{h['iteration']:4d}   {h['objective_value']: .6e}   {h['relative_norm']:.3e}        "
                  f"{h['mesh_quality']:.3f}          {h['stepsize']:.3e}")
    return sop


if __name__ == "__main__":
    main()
