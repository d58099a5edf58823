"""Diagnostic for VERDICT r01 weak #1: per-iteration / per-trial log of the degenerate NCG(DY) run
(``initial_stepsize 1e-3``, reference ``tests/test_shape_optimization.py:1003-1012``) on device and oracle, and
the first decision at which the two differ, with the margin of that decision.

    python tools/diag_armijo.py > gpurun_out/diag_armijo.json
"""

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import cashocs_b200 as cb  # noqa: E402
from oracle import meshes, pipeline  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.join(GOLDEN, ".."))
import test_gpu_pipeline as T  # noqa: E402


def device_run():
    dm, _ = T.circle()
    cfg = T.sop_config()
    cfg.set("LineSearch", "initial_stepsize", "1e-3")
    cfg.set("AlgoCG", "cg_method", "DY")
    sop = T.make_sop(dm, cfg)
    trials = []
    ls_cls = type(sop.line_search) if hasattr(sop, "line_search") else None
    try:
        sop.solve(algorithm="ncg", rtol=1e-3, atol=0.0, max_iter=1000)
    except cb.NotConvergedError as e:
        msg = str(e)
    else:
        msg = "converged"
    solver = sop.solver
    hist = [{k: (float(v) if v is not None else None) for k, v in h.items()} for h in solver.history]
    return dict(message=msg, iterations=solver.iteration, history=hist,
                log=[(a, float(b)) for a, b in solver.line_search.log], trials=trials)


def oracle_run():
    _, om = T.circle()
    ocfg = pipeline.ShapeConfig(tol_lower=0.01, tol_upper=0.02, initial_stepsize=1e-3)
    prob = pipeline.PoissonShapeProblem(om, config=ocfg)
    logs = []
    orig = pipeline.armijo_step

    def wrapped(*a, **k):
        r = orig(*a, **k)
        logs.extend(r[2])
        return r

    pipeline.armijo_step = wrapped
    try:
        ok, hist = pipeline.ncg(prob, method="DY", rtol=1e-3, max_iter=1000)
    finally:
        pipeline.armijo_step = orig
    return dict(ok=bool(ok), iterations=len(hist) - 1,
                history=[dict(iteration=h[0], objective_value=float(h[1]), relative_norm=float(h[2]),
                              mesh_quality=float(h[3]), stepsize=float(h[4])) for h in hist],
                log=[(a, float(b)) for a, b in logs])


if __name__ == "__main__":
    d, o = device_run(), oracle_run()
    first = None
    for i, (a, b) in enumerate(zip(d["log"], o["log"])):
        if a[0] != b[0] or abs(a[1] - b[1]) > 1e-15 * abs(b[1]):
            first = dict(index=i, device=a, oracle=b)
            break
    rows = []
    for hd, ho in zip(d["history"], o["history"]):
        rows.append(dict(it=hd["iteration"], J_dev=hd["objective_value"], J_orc=ho["objective_value"],
                         dJ=hd["objective_value"] - ho["objective_value"],
                         rel_dev=hd["relative_norm"], rel_orc=ho["relative_norm"],
                         q_dev=hd["mesh_quality"], q_orc=ho["mesh_quality"]))
    print(json.dumps(dict(device_iterations=d["iterations"], oracle_iterations=o["iterations"], message=d["message"],
                          n_log_device=len(d["log"]), n_log_oracle=len(o["log"]), first_difference=first,
                          device_log_tail=d["log"][-12:], oracle_log_tail=o["log"][-12:], rows=rows), indent=1))
