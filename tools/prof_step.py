#!/usr/bin/env python
"""A short slice of the bench step for ncu: same mesh (default n=119) and the same kernels as
`bench.py`, but the Krylov solves are cut after `--its` iterations so the whole run is a few
hundred launches (ncu replays every launch; the full step has ~12k).  The per-kernel durations and
the share of each kernel per *iteration* are what is read from the launch list; the share per step
follows from the iteration counts bench.py prints."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
import cashocs_b200 as cb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh-n", dest="n", type=int, default=119)
    ap.add_argument("--its", type=int, default=12)
    ap.add_argument("--pc", default="jacobi")
    args = ap.parse_args()
    bench.KSP["ksp_max_it"] = args.its
    bench.KSP["pc_type"] = args.pc
    sop = bench.build_problem(args.n)
    for name in ("state_problem", "adjoint_problem", "gradient_problem"):
        prob = getattr(sop, name)
        orig = prob.linear_solver.solve

        def tolerant(*a, _orig=orig, _prob=prob, **k):
            try:
                return _orig(*a, **k)
            except cb.PETScKSPError as e:  # -3 after `its` iterations is expected here
                assert e.error_code == -3
                return _prob.linear_solver.last_result

        prob.linear_solver.solve = tolerant
    for _ in range(2):
        out = bench.evaluation_step(sop, 1.0 / args.n)
    torch.cuda.synchronize()
    print("ok", out)


if __name__ == "__main__":
    main()
