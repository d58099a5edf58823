"""CPU oracle for the cashocs shape-gradient hot path (TEST INFRASTRUCTURE ONLY).

This package is a numpy/scipy *restatement* of what the reference computes on the
path SURVEY.md section 8 names (P1 assembly -> CSR, Krylov solves, elasticity Riesz
projection, mesh quality, a-priori det check, move_mesh, Armijo decisions).

It is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing
under ``cashocs_b200/`` imports ``oracle``.

Parity pinning status
---------------------
The reference (sblauth/cashocs) cannot be imported in this container: all arithmetic
lives in un-vendored third-party code (legacy FEniCS 2019.1 -- dolfin / FFC / FIAT / UFL --
and PETSc through petsc4py; pins: ``.github/conda/openmpi.yml:6``,
``.github/docker/Dockerfile:6``, ``README.rst:88-89``).  The oracle therefore restates the
*published* algorithms (P1 Galerkin assembly with dolfin's symmetric per-element
Dirichlet elimination, PETSc's KSPCG with the preconditioned-norm default convergence
test) and is pinned against everything the reference's own tests hold for this path:

* mesh-quality known answers ``tests/test_geometry.py:138-269`` (0.4714045207910318,
  0.3162277660168379, skewness / maximum-angle from 90/45 degree angles) -> pinned, 1e-14;
* the analytic Poisson shape derivative ``tests/test_shape_optimization.py:229-299``
  and the Taylor-test rate > 1.9 ``:302-318`` (``cashocs/verification.py:182-280``) -> pinned;
* state / adjoint / control-gradient definitions ``tests/test_pde_problems.py:87-128`` -> pinned
  against scipy ``spsolve`` of the textbook systems;
* move / revert exactness and rejection of a +-1e3 random deformation
  ``tests/test_shape_optimization.py:157-179`` -> pinned;
* mesh-import invariants ``tests/test_geometry.py:51-99`` on the reference's gmsh fixtures -> pinned;
* optimiser iteration budgets: shape problems on the unit disc ``tests/test_shape_optimization.py:321-353``
  (gd <= 32, NCG FR / PR / HS / DY / HZ <= 21 / 16 / 18 / 18 / 18, L-BFGS <= 7), ``:934-942`` (angle_change,
  L-BFGS <= 15), ``:945-968`` (fixed dimensions) and control problems ``tests/test_optimal_control.py:147-195``
  (gd <= 46, L-BFGS <= 7, NCG <= 20 / 25 / 27 / 9 / 27, restarts <= 10 / 24) -> pinned: the oracle needs exactly
  one iteration less than each of these seventeen budgets, i.e. it reproduces the accept / reject sequences
  of the reference's line searches;
* the other line searches and the restarted L-BFGS: polynomial line search ``tests/test_optimal_control.py:588-596``
  (cubic <= 42, quadratic <= 44), constant stepsize ``:819-829`` (gd / ncg / bfgs <= 71 / 76 / 43) and
  ``tests/test_shape_optimization.py:1144-1160`` (<= 45 / 12 / 33), ``bfgs_periodic_restart = 2``
  ``tests/test_optimal_control.py:185-189`` (<= 17) -> pinned, again exactly budget - 1 for all nine;
* box-constrained controls 0 <= u <= 100 ``tests/test_optimal_control.py:196-225`` (gd <= 22, L-BFGS <= 11,
  NCG FR / PR / HS / DY / HZ <= 47 / 24 / 30 / 9 / 37) -> pinned, exactly budget - 1 for all seven
  (33 reference iteration counts in total);
* Newton's method for the semilinear state equation of the reference's tests (``oracle/nonlinear.py``;
  ``tests/test_nonlinear_solvers.py:28-61``, ``tests/test_optimal_control.py:364-376``): solution equal to an independent
  nonlinear solve, control-gradient Taylor rate > 1.9 -> pinned (the device layer does not have the nonlinear
  kernels yet: this is the checker they will be held to);
* P2 state / adjoint (``oracle/p2.py``): no reference test holds values -> validated by patch tests and
  Taylor tests only ("parity unpinned" for the P2 element tensors).

``oracle/amg.py`` (smoothed-aggregation multigrid in scipy) stands in for hypre BoomerAMG in the timed CPU arm of
``bench.py``; multigrid iteration counts are "parity unpinned" like all Krylov iteration counts.

Assembled matrix entries, dof numbering, BC-row diagonals and Krylov iteration counts are
NOT pinned by any reference test (the reference has no golden vectors): for those
quantities this oracle is "parity unpinned" and says so in DESIGN.md.
"""
