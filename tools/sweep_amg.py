#!/usr/bin/env python
"""Sweep AMG parameters (aggregation distance, smoother degree) on the bench problem: iterations,
setup and solve times of the state and the elasticity solve (diagnostic only)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cashocs_b200._utils import amg, linalg  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh-n", dest="n", type=int, default=119)
    ap.add_argument("--configs", default="1:2,2:1,2:2,2:3,2:4")
    args = ap.parse_args()
    sop = bench.build_problem(args.n)
    fh = sop.form_handler
    fh.update_scalar_product()
    A, b = sop.state_problem.assemble(0)
    sop.state_problem.solve()
    sop.adjoint_problem.solve()
    rhs = fh.assemble_shape_derivative()
    E = fh.scalar_product_matrix
    for cfg in args.configs.split(","):
        parts = cfg.split(":")
        dist, deg = int(parts[0]), int(parts[1])
        ratio = float(parts[2]) if len(parts) > 2 else 4.0
        amg.AMGSymbolic.DISTANCE = dist
        A.__dict__.pop("_amg", None); E.__dict__.pop("_amg", None)
        opts = dict(bench.KSP)
        opts.update({"pc_amg_smooth_degree": deg, "pc_amg_smooth_ratio": ratio})
        for rep in range(2):  # second repetition: symbolic cached, numeric redone
            A.touch(); E.touch()
            hA, tA = timed(lambda: linalg.LinearSolver.amg_hierarchy(A, opts))
            hE, tE = timed(lambda: linalg.LinearSolver.amg_hierarchy(E, opts))
            x = torch.zeros(A.ncols, dtype=torch.float64, device="cuda")
            rA, tsA = timed(lambda: linalg.LinearSolver().solve(x, A=A, b=b, ksp_options=opts))
            g = torch.zeros(E.ncols * 3, dtype=torch.float64, device="cuda")
            rE, tsE = timed(lambda: linalg.LinearSolver().solve(g, A=E, b=rhs, ksp_options=opts))
        print(f"dist {dist} degree {deg} ratio {ratio}: levels {hE.sizes} opc {hE.operator_complexity:.3f} | scalar numeric {tA:.1f} ms solve {tsA:.1f} ms its {rA.its}"
              f" | elasticity numeric {tE:.1f} ms solve {tsE:.1f} ms its {rE.its} ({tsE / rE.its:.2f} ms/it) | gnorm {float(g.norm()):.12e}", flush=True)


if __name__ == "__main__":
    main()
