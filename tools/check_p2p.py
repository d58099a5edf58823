"""N-rank check of the one-sided NVLink transport (torchrun --nproc-per-node N): halo exchange and
small all-reduces through peer windows must reproduce the NCCL results bit for bit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as td

from cashocs_b200 import _lib, distributed

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dist = distributed.DistContext(rank, world)
mesh = distributed.distributed_regular_mesh(int(sys.argv[1]) if len(sys.argv) > 1 else 12, dist)
print(f"rank {rank}: p2p {dist.p2p} neigh {dist.neigh} owned {dist.n_owned} ghost {dist.n_ghost}", flush=True)
assert dist.p2p, "one-sided transport was not enabled"
lib = _lib.load()
ok = True
gid = mesh.global_vertex_ids.to(torch.float64)
for rep in range(5):
    for bs in (1, 3):
        x = torch.zeros(mesh.num_vertices * bs, dtype=torch.float64, device="cuda")
        ref = (gid[:, None] * 10 + torch.arange(bs, device="cuda")[None, :] + rep).reshape(-1)
        x[: mesh.n_owned_vertices * bs] = ref[: mesh.n_owned_vertices * bs]
        dist.halo_exchange(x, bs)
        torch.cuda.synchronize()
        good = torch.equal(x, ref)
        ok &= good
        if not good:
            print(f"rank {rank} rep {rep} bs {bs}: halo mismatch {(x != ref).sum().item()}", flush=True)
    for op, fn in (("sum", sum), ("max", max), ("min", min)):
        t = torch.tensor([rank + 1.5 + rep, -2.0 * rank], dtype=torch.float64, device="cuda")
        dist.allreduce(t, op)
        exp = [fn(r + 1.5 + rep for r in range(world)), fn(-2.0 * r for r in range(world))]
        good = t.tolist() == exp
        ok &= good
        if not good:
            print(f"rank {rank} rep {rep} {op}: {t.tolist()} != {exp}", flush=True)
dist.check_p2p()
print(f"rank {rank}: {'P2P_OK' if ok else 'P2P_FAILED'}", flush=True)
td.barrier()
dist.destroy()
td.destroy_process_group()
