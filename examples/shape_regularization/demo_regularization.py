"""The reference's documented demo ``demos/documented/shape_optimization/regularization/demo_regularization.py`` on
the B200 path: volume, surface and curvature regularisation from the demo's own ``config.ini`` keys, plus the geometric
terms the demo puts into the cost functional itself.

Reference (FEniCS / UFL)                                          | here
------------------------------------------------------------------+------------------------------------------------
alpha_vol = 1e-1; alpha_surf = 1e-1                               | the same
J = cashocs.IntegralFunctional(u*dx + Constant(alpha_vol)*dx      | J = cb.IntegralFunctional(c_state=1.0, c_volume=alpha_vol,
                               + Constant(alpha_surf)*ds)         |                           c_surface=alpha_surf)
sop = cashocs.ShapeOptimizationProblem(e, bcs, J, u, p,           | sop = cb.ShapeOptimizationProblem(e, bcs, J, u, p, config=config)
                                       boundaries, config=config) |
sop.solve()                                                       | sop.solve()

Run on a GPU box:  python examples/shape_regularization/demo_regularization.py
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
    alpha_vol = 1e-1
    alpha_surf = 1e-1
    J = cb.IntegralFunctional(c_state=1.0, c_volume=alpha_vol, c_surface=alpha_surf)
    sop = cb.ShapeOptimizationProblem(e, bcs, J, u, p, config=config)
    sop.solve()
    if verbose:
        for h in sop.solver.history:
            print(f"{h['iteration']:4d}   {h['objective_value']: .6e}   {h['relative_norm']:.3e}   {h['mesh_quality']:.3f}   {h['stepsize']:.3e}")
        reg = sop.form_handler.shape_regularization
        print(f"volume {reg.compute_volume():.4f} (target 1.5), surface {reg.compute_surface():.4f} (target 4.5)")
    return sop


if __name__ == "__main__":
    main()
