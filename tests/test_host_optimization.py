"""The shipped host optimisation layer (algorithms + line searches of ``cashocs_b200/_optimization``) executed on
the CPU stand-in backend of ``tests/cpu_backend.py`` and compared with the oracle's own loops: same reference
iteration budgets, same iterates.  The GPU tests run the identical host code on the CUDA layer."""

import os

import numpy as np
import pytest

import cashocs_b200 as cb
from oracle import meshes, pipeline
from cpu_backend import HostProblem

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def control_oracle(cc=None, **cfg):
    om = meshes.regular_mesh(10)
    yd = np.sin(2 * np.pi * om.coords[:, 0]) * np.sin(2 * np.pi * om.coords[:, 1])
    prob = pipeline.PoissonControlProblem(om, yd, np.zeros(om.num_vertices), alpha=1e-6, bc_markers=(1, 2, 3, 4),
                                          control_constraints=cc)
    prob._cfg = pipeline.ShapeConfig(**cfg)
    return prob


def shape_oracle(**cfg):
    om = meshes.load_npz(os.path.join(GOLDEN, "unit_circle.npz"))
    return pipeline.PoissonShapeProblem(om, config=pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, **cfg))


def host_config(shape: bool, **keys) -> cb.Config:
    """tests/config_sop.ini / tests/config_ocp.ini of the reference, as far as the loops read them."""
    cfg = cb.Config()
    cfg.set("Output", "save_results", "False")
    cfg.set("AlgoLBFGS", "bfgs_memory_size", "3")
    cfg.set("AlgoCG", "cg_method", "DY" if shape else "PR")
    if shape:
        cfg.set("MeshQuality", "tol_lower", "0.01")
        cfg.set("MeshQuality", "tol_upper", "0.02")
    for k, v in keys.items():
        section = {"method": "LineSearch", "polynomial_model": "LineSearch", "initial_stepsize": "LineSearch",
                   "bfgs_periodic_restart": "AlgoLBFGS", "damped": "AlgoLBFGS", "cg_method": "AlgoCG", "cg_periodic_restart": "AlgoCG",
                   "cg_periodic_its": "AlgoCG", "cg_relative_restart": "AlgoCG", "cg_restart_tol": "AlgoCG"}[k]
        cfg.set(section, k, str(v))
    cfg.validate_config()
    return cfg


# (problem, algorithm, host config keys, oracle loop, oracle kwargs, oracle config, reference budget, reference test)
CASES = [
    ("control", "gd", {}, "gradient_descent", {}, {}, 46, "test_optimal_control.py:147"),
    ("control", "bfgs", {}, "lbfgs", {"memory": 3}, {}, 7, "test_optimal_control.py:180"),
    ("control", "bfgs", {"bfgs_periodic_restart": 2}, "lbfgs", {"memory": 3, "periodic_restart": 2}, {}, 17,
     "test_optimal_control.py:185-189"),
    ("control", "ncg", {"cg_method": "FR"}, "ncg", {"method": "FR"}, {}, 20, "test_optimal_control.py:152"),
    ("control", "ncg", {"cg_method": "DY", "cg_periodic_restart": True, "cg_periodic_its": 5}, "ncg",
     {"method": "DY", "periodic_restart": True, "periodic_its": 5}, {}, 10, "test_optimal_control.py:166"),
    ("control", "gd", {"method": "polynomial", "polynomial_model": "cubic"}, "gradient_descent", {},
     {"line_search": "polynomial", "polynomial_model": "cubic"}, 42, "test_optimal_control.py:588-596"),
    ("control", "gd", {"method": "polynomial", "polynomial_model": "quadratic"}, "gradient_descent", {},
     {"line_search": "polynomial", "polynomial_model": "quadratic"}, 44, "test_optimal_control.py:588-596"),
    ("control", "gd", {"method": "basic", "initial_stepsize": 5e2}, "gradient_descent", {},
     {"line_search": "basic", "initial_stepsize": 5e2}, 71, "test_optimal_control.py:819-829"),
    ("control", "ncg", {"method": "basic", "initial_stepsize": 5e2}, "ncg", {"method": "PR"},
     {"line_search": "basic", "initial_stepsize": 5e2}, 76, "test_optimal_control.py:819-829"),
    ("control", "bfgs", {"method": "basic", "initial_stepsize": 0.1}, "lbfgs", {"memory": 3},
     {"line_search": "basic", "initial_stepsize": 0.1}, 43, "test_optimal_control.py:819-829"),
    # box constraints 0 <= u <= 100 (fixture ocp_cc)
    ("control", "gd", {}, "gradient_descent", {}, {"cc": (0, 100)}, 22, "test_optimal_control.py:196-200"),
    ("control", "bfgs", {}, "lbfgs", {"memory": 3}, {"cc": (0, 100)}, 11, "test_optimal_control.py:221-225"),
    ("control", "ncg", {"cg_method": "FR"}, "ncg", {"method": "FR"}, {"cc": (0, 100)}, 47, "test_optimal_control.py:203-218"),
    ("control", "ncg", {"cg_method": "PR"}, "ncg", {"method": "PR"}, {"cc": (0, 100)}, 24, "test_optimal_control.py:203-218"),
    ("control", "ncg", {"cg_method": "HS"}, "ncg", {"method": "HS"}, {"cc": (0, 100)}, 30, "test_optimal_control.py:203-218"),
    ("control", "ncg", {"cg_method": "DY"}, "ncg", {"method": "DY"}, {"cc": (0, 100)}, 9, "test_optimal_control.py:203-218"),
    ("control", "ncg", {"cg_method": "HZ"}, "ncg", {"method": "HZ"}, {"cc": (0, 100)}, 37, "test_optimal_control.py:203-218"),
    ("shape", "bfgs", {}, "lbfgs", {"memory": 3}, {}, 7, "test_shape_optimization.py:326"),
    ("shape", "ncg", {"cg_method": "PR"}, "ncg", {"method": "PR"}, {}, 16, "test_shape_optimization.py:336"),
    ("shape", "gd", {"method": "basic", "initial_stepsize": 0.25}, "gradient_descent", {},
     {"line_search": "basic", "initial_stepsize": 0.25}, 45, "test_shape_optimization.py:1144-1160"),
    ("shape", "ncg", {"method": "basic", "initial_stepsize": 0.25}, "ncg", {"method": "DY"},
     {"line_search": "basic", "initial_stepsize": 0.25}, 12, "test_shape_optimization.py:1144-1160"),
    ("shape", "bfgs", {"method": "basic", "initial_stepsize": 0.1}, "lbfgs", {"memory": 3},
     {"line_search": "basic", "initial_stepsize": 0.1}, 33, "test_shape_optimization.py:1144-1160"),
]


@pytest.mark.parametrize("kind,algorithm,keys,loop,loop_kw,ocfg,budget,ref", CASES,
                         ids=[f"{c[0]}-{c[1]}-{c[6]}{'-cc' if 'cc' in c[5] else ''}" for c in CASES])
def test_host_layer_reproduces_oracle_and_reference_budgets(kind, algorithm, keys, loop, loop_kw, ocfg, budget, ref):
    make = control_oracle if kind == "control" else shape_oracle
    # the oracle's own loop: exactly one iteration less than the reference's budget
    ok, hist = getattr(pipeline, loop)(make(**ocfg), rtol=1e-2, max_iter=budget, **loop_kw)
    assert ok and len(hist) - 1 == budget - 1, ref
    # the shipped host layer on the stand-in backend: converges within the reference's budget with the same iterates
    host = HostProblem(make(**ocfg), host_config(kind == "shape", **keys))
    host.solve(algorithm, rtol=1e-2, atol=0.0, max_iter=budget)
    s = host.solver
    assert s.converged and s.relative_norm <= 1e-2 and s.iteration == len(hist) - 1, ref
    got = np.array([[h["objective_value"], h["relative_norm"], h["stepsize"]] for h in s.history])
    want = np.array([[h[1], h[2], 0.0] for h in hist])
    assert np.allclose(got[:, :2], want[:, :2], rtol=1e-6, atol=0.0), ref  # round-off grows over ~40 mesh moves
    if "cc" in ocfg:  # the iterate is admissible (tests/test_optimal_control.py:199-200)
        u = host.controls[0].vector.numpy()
        assert u.min() >= ocfg["cc"][0] and u.max() <= ocfg["cc"][1] and (u == 0.0).sum() > 4 * 10


@pytest.mark.parametrize("inner", ["cg", "cr"])
@pytest.mark.parametrize("cc,budget,ref", [(None, 2, "test_optimal_control.py:193-195"), ((0, 100), 9, "test_optimal_control.py:228-232")])
def test_truncated_newton_host_layer(inner, cc, budget, ref):
    """The truncated Newton method (newton.py:57-101) with the inner CG / CR solvers of the Hessian problem
    (hessian_problems.py:267-431) and the settings of tests/config_ocp.ini:24-28: the oracle's loop meets the
    reference's budgets (unconstrained: converged to 1e-6 after ONE step; box constraints: 7 (cg) / 8 (cr) of the 9
    iterations), the shipped host layer walks the same iterates on the stand-in backend."""
    ocfg = {} if cc is None else {"cc": cc}
    ok, hist = pipeline.truncated_newton(control_oracle(**ocfg), inner=inner, rtol=1e-2, max_iter=budget)
    assert ok and len(hist) - 1 < budget, ref
    if cc is None:
        assert len(hist) - 1 == 1 and hist[-1][2] <= 1e-6, ref
    else:
        assert len(hist) - 1 == {"cg": 7, "cr": 8}[inner], ref
    cfg = host_config(False)
    for k, v in (("inner_newton", inner), ("max_it_inner_newton", 20), ("inner_newton_rtol", 1e-7), ("inner_newton_atol", 0.0)):
        cfg.set("AlgoTNM", k, str(v))
    host = HostProblem(control_oracle(**ocfg), cfg)
    host.solve("newton", rtol=1e-2, atol=0.0, max_iter=budget)
    s = host.solver
    assert s.converged and s.iteration == len(hist) - 1, ref
    got = np.array([[h["objective_value"], h["relative_norm"]] for h in s.history])
    want = np.array([[h[1], h[2]] for h in hist])
    # the inner solves stop at rtol 1e-7: iterates agree to a few times that, not to round-off
    assert np.allclose(got[:-1], want[:-1], rtol=1e-5, atol=0.0), ref
    # the last relative norm is at the level of the inner tolerance: compare it on that scale
    assert abs(got[-1, 1] - want[-1, 1]) <= 1e-6 and abs(got[-1, 0] - want[-1, 0]) <= 1e-6 * abs(want[-1, 0]), ref
    assert host.hessian_problem.no_sensitivity_solves > 0
    if cc is not None:
        u = host.controls[0].vector.numpy()
        assert u.min() >= cc[0] and u.max() <= cc[1]


def test_small_stepsize_fails_like_the_reference():
    """tests/test_optimal_control.py:562-567: initial_stepsize 1e-8 -> NotConvergedError "Armijo rule failed."."""
    cfg = host_config(False, initial_stepsize=1e-8)
    host = HostProblem(control_oracle(initial_stepsize=1e-8), cfg)
    with pytest.raises(cb.NotConvergedError, match="Armijo rule failed."):
        host.solve("gd", rtol=1e-2, atol=0.0, max_iter=2)
    ok, hist = pipeline.gradient_descent(control_oracle(initial_stepsize=1e-8), rtol=1e-2, max_iter=2)
    assert not ok and host.solver.converged_reason == -2


def test_box_constraints_are_validated():
    """box_constraints.py:106-120: the lower bound must be below the upper bound."""
    with pytest.raises(ValueError):
        control_oracle(cc=(1.0, 0.0))
    import torch
    import types

    from cashocs_b200._optimization import line_search as ls

    u = types.SimpleNamespace(vector=torch.zeros(4, dtype=torch.float64))
    with pytest.raises(cb.InputError, match="lower bound must always be smaller"):
        ls.BoxConstraints(u, [1.0, 0.0])
    assert not ls.BoxConstraints(u, None).require_control_constraints
    assert ls.BoxConstraints(u, [float("-inf"), 3.0]).require_control_constraints


@pytest.mark.parametrize("kind", ["control", "shape"])
def test_damped_bfgs_host_layer_equals_oracle(kind):
    """``[AlgoLBFGS] damped = True`` (Powell's damping, l_bfgs.py:284-304): the shipped host layer and the oracle walk the
    same iterates and converge.  (The reference's own test, tests/test_optimal_control.py:598-621, uses a cost
    functional outside the form family, so there is no iteration count to pin.)"""
    make = control_oracle if kind == "control" else shape_oracle
    ok, hist = pipeline.lbfgs(make(), rtol=1e-2, max_iter=40, memory=3, damped=True)
    assert ok
    host = HostProblem(make(), host_config(kind == "shape", damped=True))
    host.solve("bfgs", rtol=1e-2, atol=0.0, max_iter=40)
    s = host.solver
    assert s.converged and s.iteration == len(hist) - 1
    got = np.array([[h["objective_value"], h["relative_norm"]] for h in s.history])
    want = np.array([[h[1], h[2]] for h in hist])
    assert np.allclose(got, want, rtol=1e-6, atol=0.0)
    ok0, hist0 = pipeline.lbfgs(make(), rtol=1e-2, max_iter=40, memory=3)
    assert abs(len(hist) - len(hist0)) <= 3  # damping only acts when the curvature condition is nearly violated


def test_global_deformation_like_the_reference():
    """tests/test_shape_optimization.py:1015-1030: with ``global_deformation = True`` the sum of the accepted
    deformations moves the initial mesh onto the optimised one (1e-13)."""
    oracle = shape_oracle()
    x0 = oracle.coords.copy()
    cfg = host_config(True)
    cfg.set("ShapeGradient", "global_deformation", "True")
    cfg.validate_config()
    host = HostProblem(oracle, cfg)
    host.solve("bfgs", rtol=1e-2, atol=0.0, max_iter=7)
    assert host.solver.converged
    total = host.global_deformation_function.vector.numpy().reshape(-1, 2)
    assert np.max(np.abs(x0 + total - oracle.coords)) <= 1e-13
    assert np.max(np.abs(total)) > 1e-2
