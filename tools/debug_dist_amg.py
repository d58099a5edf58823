"""N-rank check of the distributed AMG hierarchy (run under torchrun --nproc-per-node N):
PCG with the AMG V-cycle must converge to the same solution as Jacobi-PCG, in far fewer iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as td

import cashocs_b200 as cb
from cashocs_b200 import distributed
from cashocs_b200._utils import amg, linalg

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dist = distributed.DistContext(rank, world)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:
    amg.AMGSymbolic.AGGLOMERATE = int(sys.argv[2])
mesh = distributed.distributed_regular_mesh(n, dist)
mesh.build_pattern()
mu = torch.full((mesh.num_vertices,), 0.357, dtype=torch.float64, device="cuda")
E = linalg.assemble_elasticity_matrix(mesh, mu, 1.43, 0.2)
A, _ = linalg.assemble_scalar_system(mesh, 1.0, 1.0, rhs=False)
ok = True
for name, M in (("scalar", A), ("elasticity", E)):
    bs = M.bs
    b = torch.from_numpy(np.random.RandomState(rank).rand(M.nrows * bs)).cuda()
    sols = {}
    for pc in ("jacobi", "gamg"):
        x = torch.zeros(M.ncols * bs, dtype=torch.float64, device="cuda")
        try:
            res = cb.LinearSolver().solve(x, A=M, b=b, ksp_options={"ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-12, "ksp_max_it": 20000})
            sols[pc] = x[: M.nrows * bs].clone()
            if rank == 0:
                extra = ""
                if pc == "gamg":
                    h = list(M._amg.values())[-1]
                    kinds = ["rep" if L.rep_dist is not None else ("dist" if L.dist is not None else "local") for L in h.levels]
                    extra = f" levels {h.sizes} {kinds} lmax {[round(D.lmax, 3) for D in h.data]}"
                print(f"n={n} world={world} {name} {pc}: reason {res.reason} its {res.its}{extra}", flush=True)
        except cb.PETScKSPError as e:
            ok = False
            if rank == 0:
                print(f"n={n} {name} {pc} FAILED reason {e.error_code}", flush=True)
    if len(sols) == 2:
        d = torch.stack([(sols["jacobi"] - sols["gamg"]).abs().max(), sols["jacobi"].abs().max()])
        dist.allreduce(d, "max")
        if rank == 0:
            print(f"   max |x_jacobi - x_amg| / max|x| = {float(d[0] / d[1]):.3e}", flush=True)
        ok = ok and float(d[0] / d[1]) < 1e-8
if rank == 0:
    print("DIST_AMG_OK" if ok else "DIST_AMG_FAILED", flush=True)
td.barrier()
dist.destroy()
td.destroy_process_group()
