#!/usr/bin/env python
"""Elasticity / state solve of the bench problem under the AMG precision variants (diagnostic only):
iterations and time per variant -- which operators of the V-cycle tolerate fp32 storage."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cashocs_b200._utils import linalg  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh-n", dest="n", type=int, default=119)
    ap.add_argument("--touch", type=int, default=0, help="extra numeric phases (matrix versions) before the timed solve")
    ap.add_argument("--variants", default="pc_amg_precision=double;pc_amg_precision=mixed:a0",
                    help="semicolon-separated option sets, each `key=value,key=value` on top of bench.KSP")
    args = ap.parse_args()
    sop = bench.build_problem(args.n)
    fh = sop.form_handler
    fh.update_scalar_product()
    A, b = sop.state_problem.assemble(0)
    sop.state_problem.solve()
    sop.adjoint_problem.solve()
    rhs = fh.assemble_shape_derivative()
    E = fh.scalar_product_matrix
    ref = {}
    for var in args.variants.split(";"):
        opts = dict(bench.KSP)
        opts["pc_amg_rigid_body_correction"] = True
        for kv in var.split(","):
            k, _, v = kv.partition("=")
            try:
                opts[k] = int(v)
            except ValueError:
                try:
                    opts[k] = float(v)
                except ValueError:
                    opts[k] = v
        for rep in range(2 + args.touch):
            if rep < args.touch:
                A.touch()
                E.touch()
            x = torch.zeros(A.ncols, dtype=torch.float64, device="cuda")
            rA, tA = timed(lambda: linalg.LinearSolver().solve(x, A=A, b=b, ksp_options={k: v for k, v in opts.items() if k != "pc_amg_rigid_body_correction"}))
            g = torch.zeros(E.ncols * 3, dtype=torch.float64, device="cuda")
            rE, tE = timed(lambda: linalg.LinearSolver().solve(g, A=E, b=rhs, ksp_options=opts))
        if not ref:
            ref = dict(x=x.clone(), g=g.clone())
        dx = float((x - ref["x"]).abs().max() / ref["x"].abs().max())
        dg = float((g - ref["g"]).abs().max() / ref["g"].abs().max())
        print(f"{var:60s} scalar its {rA.its:3d} {tA:7.2f} ms | elasticity its {rE.its:3d} {tE:7.2f} ms ({tE / rE.its:.3f} ms/it) "
              f"| vs first variant: du {dx:.1e} dg {dg:.1e}", flush=True)


if __name__ == "__main__":
    main()
