#!/usr/bin/env python
"""Phase timing of one bench step (CUDA events + sync around each phase; diagnostic only)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from cashocs_b200._utils import linalg  # noqa: E402


class T:
    def __init__(self):
        self.acc = {}

    def __call__(self, name):
        t = self

        class Ctx:
            def __enter__(self):
                torch.cuda.synchronize()
                self.t0 = time.perf_counter()

            def __exit__(self, *a):
                torch.cuda.synchronize()
                t.acc[name] = t.acc.get(name, 0.0) + (time.perf_counter() - self.t0) * 1e3

        return Ctx()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh-n", dest="n", type=int, default=119)
    ap.add_argument("--pc", default="hypre")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--rbm", type=int, default=0, help="1: rigid-body coarse correction in the elasticity solve")
    ap.add_argument("--nullspace", default="translations")
    ap.add_argument("--precision", default="mixed")
    args = ap.parse_args()
    bench.KSP["pc_amg_precision"] = args.precision
    bench.GRADIENT_KSP_EXTRA["pc_amg_nullspace"] = args.nullspace
    bench.KSP["pc_type"] = args.pc
    bench.KSP["pc_amg_rigid_body_correction"] = bool(args.rbm)
    dist = None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:  # torchrun: the same phases on a partitioned mesh (times of rank 0)
        import torch.distributed as td

        from cashocs_b200 import distributed

        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
        dist = distributed.DistContext(int(os.environ["RANK"]), world)
    sop = bench.build_problem(args.n, dist)
    h = 1.0 / args.n
    bench.evaluation_step(sop, h)
    timer = T()
    fh = sop.form_handler
    for _ in range(args.reps):
        sop.state_problem.has_solution = sop.adjoint_problem.has_solution = sop.gradient_problem.has_solution = False
        with timer("mu + elasticity assembly"):
            fh.update_scalar_product()
        with timer("state assembly"):
            A, b = sop.state_problem.assemble(0)
        if args.pc == "hypre":
            with timer("amg numeric scalar"):
                linalg.LinearSolver.amg_hierarchy(A, bench.KSP)
            with timer("amg numeric elasticity"):
                if args.nullspace == "rigid_body":
                    hE = linalg.LinearSolver.rbm_hierarchy(fh.scalar_product_matrix)
                else:
                    hE = linalg.LinearSolver.amg_hierarchy(fh.scalar_product_matrix, bench.KSP)
        with timer("state solve"):
            sop.state_problem.linear_solver.solve(sop.states[0].vector, A=A, b=b, ksp_options=sop.ksp_options[0])
        sop.state_problem.has_solution = True
        with timer("adjoint assemble+solve"):
            sop.adjoint_problem.solve()
        with timer("shape derivative"):
            rhs = fh.assemble_shape_derivative()
        with timer("elasticity solve"):
            g = sop.gradient.vector
            g.zero_()
            res = sop.gradient_problem.linear_solver.solve(g, A=fh.scalar_product_matrix, b=rhs, ksp_options=sop.gradient_problem.ksp_options)
        with timer("gAg"):
            fh.scalar_product(g, g)
        with timer("move (det+move+intersection+quality) + revert"):
            ginf = float(g.abs().max())
            if dist is not None:
                ginf = dist.allreduce_scalar(ginf, "max")
            if sop.mesh_handler.move_mesh(g, step=-0.1 * h / ginf):
                sop.mesh_handler.revert_transformation()
    if dist is not None and dist.rank != 0:
        import torch.distributed as td

        td.barrier()
        dist.destroy()
        td.destroy_process_group()
        return
    for k, v in timer.acc.items():
        print(f"{k:50s} {v / args.reps:10.2f} ms")
    print("its", sop.state_problem.linear_solver.last_result.its, sop.adjoint_problem.iterations, res.its)
    if args.pc == "hypre":
        sym = getattr(hE, "symbolic", hE)
        print("levels", sym.sizes, "opc %.3f" % sym.operator_complexity, [L.nnzb for L in hE.levels], [round(D.lmax, 3) for D in hE.data])
        for L in hE.levels:
            if hasattr(L, "p_col") and L.p_col is not None and not getattr(L, "last", False):
                print("  level n", L.n, "nnzb", L.nnzb, "P entries", int(L.p_col.numel()), "AP entries", int(L.ap_col.numel()))
    if dist is not None:
        import torch.distributed as td

        td.barrier()
        dist.destroy()
        td.destroy_process_group()


if __name__ == "__main__":
    main()
