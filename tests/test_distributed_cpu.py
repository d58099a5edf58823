"""CPU, world_size 2 (gloo): mesh partitioning, owned/ghost numbering and the halo plan reproduce
the global operator -- the host-side logic of the multi-GPU path."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cashocs_b200 as cb
        from cashocs_b200 import distributed
        from oracle import fem, meshes

        dist = distributed.DistContext(rank, world, transport="torch")
        if kind == "cube":
            om = meshes.regular_mesh(4, 1.0, 1.0, 1.0)
            lm = distributed.distributed_regular_mesh(4, dist, device="cpu")
        else:
            om = meshes.regular_mesh(7)
            g = cb.regular_mesh(7, device="cpu")
            off = distributed.vertex_offsets(g.num_vertices, world)  # not layer aligned: ragged halos
            lm = distributed.partition_mesh(g.coords, g.cells, {k: g.boundary_vertices([k]) for k in g.boundary_markers}, off, dist)
        nv = om.num_vertices
        gid = lm.global_vertex_ids.numpy()
        lo, hi = int(lm.global_offsets[rank]), int(lm.global_offsets[rank + 1])
        assert np.array_equal(gid[: lm.n_owned_vertices], np.arange(lo, hi))
        assert np.array_equal(lm.coords.numpy(), om.coords[gid])
        # local cells == all global cells touching an owned vertex; owned cells == lowest vertex owned
        gcells = gid[lm.cells.numpy()]
        touch = ((om.cells >= lo) & (om.cells < hi)).any(axis=1)
        assert sorted(map(tuple, gcells)) == sorted(map(tuple, om.cells[touch]))
        own = (om.cells[:, 0] >= lo) & (om.cells[:, 0] < hi)
        assert sorted(map(tuple, gcells[: lm.n_owned_cells])) == sorted(map(tuple, om.cells[own]))
        ncells = torch.tensor([float(lm.n_owned_cells)], dtype=torch.float64)
        dist.allreduce(ncells, "sum")
        assert int(ncells.item()) == om.num_cells
        # boundary markers survive the renumbering
        for k in np.unique(om.facet_tags):
            assert set(gid[lm.boundary_vertices([int(k)]).numpy()]) == set(om.boundary_vertices([k])) & set(gid)
        # halo exchange: ghosts receive the owner's values, for scalar and blocked vectors
        for bs in (1, 3):
            xg = np.random.RandomState(7).rand(nv * bs)
            x = torch.zeros(lm.num_vertices * bs, dtype=torch.float64)
            x[: lm.n_owned_vertices * bs] = torch.from_numpy(xg.reshape(nv, bs)[lo:hi].reshape(-1))
            dist.halo_exchange(x, bs)
            assert np.array_equal(x.numpy(), xg.reshape(nv, bs)[gid].reshape(-1))
        # distributed SpMV (owned rows, local columns) == global SpMV
        _, vol, G = fem.cell_geometry(om.coords, om.cells)
        A = fem.assemble_matrix(fem.stiffness_elements(vol, G) + fem.mass_elements(vol, om.dim), om.cells, nv)
        _, lvol, lG = fem.cell_geometry(lm.coords.numpy(), lm.cells.numpy())
        Aloc = fem.assemble_matrix(fem.stiffness_elements(lvol, lG) + fem.mass_elements(lvol, om.dim), lm.cells.numpy(), lm.num_vertices)
        xg = np.random.RandomState(8).rand(nv)
        y = (Aloc @ xg[gid])[: lm.n_owned_vertices]
        assert np.allclose(y, (A @ xg)[lo:hi], rtol=1e-13, atol=1e-15)
        # distributed dot
        part = torch.tensor([float(xg[lo:hi] @ xg[lo:hi])], dtype=torch.float64)
        dist.allreduce(part, "sum")
        assert abs(part.item() - xg @ xg) < 1e-12
        assert abs(dist.allreduce_scalar(float(rank), "max") - (world - 1)) == 0
        results[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback

        results[rank] = traceback.format_exc() + str(exc)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("kind", ["cube", "square_ragged"])
def test_partition_and_halo_world2(kind):
    world = 2
    port = 29500 + (os.getpid() % 2000) + (0 if kind == "cube" else 1)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, kind, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)
