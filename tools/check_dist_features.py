"""N-rank check of round-2 features on a partitioned mesh (torchrun --nproc-per-node N): the distance-based Lame
parameter (eikonal distance: assembly, AMG-PCG, right-hand-side and residual kernels with halo exchanges / all-reduces)
and the re-extension of the gradient deformation (mode surface) must equal the one-GPU results on the same mesh."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def problem(mesh, **keys):
    x = cb.spatial_coordinate(3)
    f = 2.5 * (x[0] + 0.4 - x[1] ** 2) ** 2 + x[0] ** 2 + x[1] ** 2 - 1
    cfg = cb.Config()
    cfg.set("StateSystem", "is_linear", "True")
    cfg.set("ShapeGradient", "shape_bdry_def", "[1, 2, 3, 4, 5, 6]")
    cfg.set("ShapeGradient", "lambda_lame", "1.428571428571429")
    cfg.set("ShapeGradient", "damping_factor", "0.2")
    cfg.set("ShapeGradient", "mu_fix", "0.35714285714285715")
    cfg.set("ShapeGradient", "mu_def", "0.35714285714285715")
    for k, v in keys.items():
        cfg.set("ShapeGradient", k, str(v))
    ksp = {"ksp_type": "cg", "pc_type": "jacobi", "ksp_rtol": 1e-13, "ksp_atol": 1e-30, "ksp_max_it": 20000}
    u, p = cb.Function(mesh), cb.Function(mesh)
    return cb.ShapeOptimizationProblem(cb.PoissonForm(source=f), cb.create_dirichlet_bcs(0.0, [1, 2, 3, 4, 5, 6]),
                                       cb.IntegralFunctional(c_state=1.0), u, p, config=cfg, ksp_options=ksp,
                                       adjoint_ksp_options=ksp, gradient_ksp_options=ksp)


dist = distributed.DistContext(rank, world)
ok = True
cases = {
    "distance_mu": dict(use_distance_mu=True, dist_min=0.1, dist_max=0.25, mu_min=1.0, mu_max=10.0),
    "reextension": dict(reextend_from_boundary=True),
}
for name, keys in cases.items():
    dmesh = distributed.distributed_regular_mesh(n, dist)
    sop = problem(dmesh, **keys)
    g = sop.compute_shape_gradient()[0].vector
    ref = problem(cb.regular_mesh(n, 1.0, 1.0, 1.0), **keys)
    g0 = ref.compute_shape_gradient()[0].vector
    no = dmesh.n_owned_vertices
    gid = dmesh.global_vertex_ids[:no]
    sel = (gid[:, None] * 3 + torch.arange(3, device="cuda")[None, :]).reshape(-1)
    errs = [float((g[: no * 3] - g0[sel]).abs().max() / g0.abs().max())]
    mu, mu0 = sop.form_handler.mu_lame.vector, ref.form_handler.mu_lame.vector
    errs.append(float((mu[:no] - mu0[gid]).abs().max() / mu0.abs().max()))
    if name == "distance_mu":
        d, d0 = sop.form_handler.stiffness.distance, ref.form_handler.stiffness.distance
        errs.append(float((d[:no] - d0[gid]).abs().max() / d0.abs().max()))
    e = torch.tensor(errs, dtype=torch.float64, device="cuda")
    dist.allreduce(e, "max")
    if rank == 0:
        print(f"{name}: n={n} world={world} errors gradient / mu / distance {[f'{float(x):.2e}' for x in e]}", flush=True)
    ok = ok and float(e.max()) < 1e-8
if rank == 0:
    print("DIST_FEATURES_OK" if ok else "DIST_FEATURES_FAILED", flush=True)
td.barrier()
dist.destroy()
td.destroy_process_group()
