"""N-rank check of the P2 state / adjoint path on a partitioned mesh (torchrun --nproc-per-node N): cost
functional, P2 state / adjoint and the shape gradient must equal the one-GPU results on the same mesh."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as td

import cashocs_b200 as cb
from cashocs_b200 import distributed

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pc = sys.argv[2] if len(sys.argv) > 2 else "jacobi"


def problem(mesh):
    x = cb.spatial_coordinate(3)
    f = 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1
    cfg = cb.Config()
    cfg.set("StateSystem", "is_linear", "True")
    cfg.set("ShapeGradient", "shape_bdry_def", "[1, 2, 3, 4, 5, 6]")
    cfg.set("ShapeGradient", "lambda_lame", "1.428571428571429")
    cfg.set("ShapeGradient", "damping_factor", "0.2")
    cfg.set("ShapeGradient", "mu_fix", "0.35714285714285715")
    cfg.set("ShapeGradient", "mu_def", "0.35714285714285715")
    ksp = {"ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-13, "ksp_atol": 1e-30, "ksp_max_it": 20000}
    u, p = cb.Function(mesh, degree=2), cb.Function(mesh, degree=2)
    return cb.ShapeOptimizationProblem(cb.PoissonForm(source=f), cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4, 5, 6]),
                                       cb.IntegralFunctional(c_state=1.0), u, p, config=cfg, ksp_options=ksp,
                                       adjoint_ksp_options=ksp, gradient_ksp_options=ksp)


dist = distributed.DistContext(rank, world)
dmesh = distributed.distributed_regular_mesh(n, dist)
sop = problem(dmesh)
g = sop.compute_shape_gradient()[0].vector
J = sop.reduced_cost_functional.evaluate()
gn2 = sop.gradient_problem.gradient_norm_squared
# one-GPU reference on every rank (small mesh)
ref = problem(cb.regular_mesh(n, 1.0, 1.0, 1.0))
g0 = ref.compute_shape_gradient()[0].vector
J0 = ref.reduced_cost_functional.evaluate()
gid = dmesh.global_vertex_ids[: dmesh.n_owned_vertices]
sel = (gid[:, None] * 3 + torch.arange(3, device="cuda")[None, :]).reshape(-1)
err_g = float((g[: dmesh.n_owned_vertices * 3] - g0[sel]).abs().max() / g0.abs().max())
sp = sop.states[0].space
own_v = sp.vertex_dof[: dmesh.n_owned_vertices]
err_u = float((sop.states[0].vector[own_v] - ref.states[0].vector[gid]).abs().max() / ref.states[0].vector.abs().max())
errs = torch.tensor([err_g, err_u, abs(J - J0) / abs(J0), abs(gn2 - ref.gradient_problem.gradient_norm_squared) / gn2],
                    dtype=torch.float64, device="cuda")
dist.allreduce(errs, "max")
if rank == 0:
    print(f"n={n} world={world} pc={pc} P2 dofs owned {sp.n_owned_dofs} / local {sp.ndofs}: its state {sop.state_problem.iterations} "
          f"(1 GPU {ref.state_problem.iterations}); errors gradient {errs[0]:.2e} state {errs[1]:.2e} J {errs[2]:.2e} gnorm2 {errs[3]:.2e}",
          flush=True)
    print("DIST_P2_OK" if float(errs.max()) < 1e-8 else "DIST_P2_FAILED", flush=True)
td.barrier()
dist.destroy()
td.destroy_process_group()
