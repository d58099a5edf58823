"""2-rank diagnostic of the distributed AMG hierarchy (run under torchrun --nproc-per-node 2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
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
mesh = distributed.distributed_regular_mesh(n, dist)
mesh.build_pattern()
mu = torch.full((mesh.num_vertices,), 0.357, dtype=torch.float64, device="cuda")
E = linalg.assemble_elasticity_matrix(mesh, mu, 1.43, 0.2)
A, _ = linalg.assemble_scalar_system(mesh, 1.0, 1.0, rhs=False)


def gather_level(L, D, bs, gids_fine, gids_coarse):
    """global COO triplets of A_l (owned rows) and P (owned rows)"""
    rp = L.rowptr.cpu().numpy(); col = L.col.cpu().numpy()
    rows = np.repeat(np.arange(L.n), np.diff(rp))
    val = D.cur_val.cpu().numpy().reshape(-1, bs, bs)
    Atrip = (gids_fine[rows], gids_fine[col], val)
    prp = L.p_rowptr.cpu().numpy()[: L.n + 1]; pcol = L.p_col.cpu().numpy()[: prp[-1]]
    prow = np.repeat(np.arange(L.n), np.diff(prp))
    Ptrip = (gids_fine[prow], gids_coarse[pcol], D.p_val.cpu().numpy()[: prp[-1]])
    return Atrip, Ptrip


for name, M in (("scalar", A), ("elasticity", E)):
    bs = M.bs
    h = amg.AMGHierarchy(M)
    h.numeric(M)
    if rank == 0:
        print(name, "levels", h.sizes, [L.ncols for L in h.levels], "lmax", [round(D.lmax, 3) for D in h.data], flush=True)
    # global ids: level 0 from the mesh, coarser levels = rank offset + local id, ghosts via halo exchange
    gids = mesh.global_vertex_ids.cpu().numpy()
    for li, L in enumerate(h.levels[:-1]):
        nxt = h.levels[li + 1]
        counts = [None] * world
        td.all_gather_object(counts, L.nc)
        off = int(np.sum(counts[:rank]))
        v = torch.zeros(nxt.ncols, dtype=torch.float64, device="cuda")
        v[: nxt.n] = torch.arange(off, off + nxt.n, dtype=torch.float64, device="cuda")
        nxt.dist.halo_exchange(v, 1)
        cg = v.cpu().numpy().round().astype(np.int64)
        Atrip, Ptrip = gather_level(L, h.data[li], bs, gids, cg)
        rp = nxt.rowptr.cpu().numpy(); col = nxt.col.cpu().numpy()
        rows = np.repeat(np.arange(nxt.n), np.diff(rp))
        Ctrip = (cg[rows], cg[col], h.data[li + 1].val.cpu().numpy().reshape(-1, bs, bs))
        allA, allP, allC = [None] * world, [None] * world, [None] * world
        td.all_gather_object(allA, Atrip); td.all_gather_object(allP, Ptrip); td.all_gather_object(allC, Ctrip)
        nf = int(max(t[0].max() for t in allA)) + 1
        ncg = int(sum(counts))
        if rank == 0:
            def blockmat(trips, nr, ncol):
                r = np.concatenate([t[0] for t in trips]); c = np.concatenate([t[1] for t in trips]); v_ = np.concatenate([t[2] for t in trips])
                return sp.bsr_matrix((v_, c, np.concatenate([[0], np.cumsum(np.bincount(r, minlength=nr))])), shape=(nr * bs, ncol * bs)) if False else \
                    sp.coo_matrix((np.concatenate([v_[:, i, j] for i in range(bs) for j in range(bs)]),
                                   (np.concatenate([r * bs + i for i in range(bs) for j in range(bs)]),
                                    np.concatenate([c * bs + j for i in range(bs) for j in range(bs)]))), shape=(nr * bs, ncol * bs)).tocsr()
            Ag = blockmat(allA, nf, nf)
            Cg = blockmat(allC, ncg, ncg)
            pr = np.concatenate([t[0] for t in allP]); pc = np.concatenate([t[1] for t in allP]); pv = np.concatenate([t[2] for t in allP])
            Pg = sp.coo_matrix((pv, (pr, pc)), shape=(nf, ncg)).tocsr()
            Pb = sp.kron(Pg, sp.identity(bs)).tocsr()
            ref = (Pb.T @ Ag @ Pb).tocsr()
            print(name, "level", li, "sym(A)", abs(Ag - Ag.T).max(), "sym(Ac)", abs(Cg - Cg.T).max(), "galerkin err", abs(Cg - ref).max() / abs(ref).max(),
                  "P rowsum min/max", Pg.sum(axis=1).min(), Pg.sum(axis=1).max(), flush=True)
        gids = cg
    x = torch.zeros(M.ncols * bs, dtype=torch.float64, device="cuda")
    b = torch.from_numpy(np.random.RandomState(rank).rand(M.nrows * bs)).cuda()
    for pc in ("jacobi", "gamg"):
        try:
            res = cb.LinearSolver().solve(x, A=M, b=b, ksp_options={"ksp_type": "cg", "pc_type": pc, "ksp_rtol": 1e-10, "ksp_max_it": 3000, "ksp_monitor": None})
            if rank == 0:
                print(name, pc, res, flush=True)
        except cb.PETScKSPError as e:
            if rank == 0:
                print(name, pc, "FAILED", e.error_code, flush=True)
td.barrier()
td.destroy_process_group()
