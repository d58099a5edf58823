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
        elif kind == "circle_rcb":
            # an imported, unstructured mesh (the reference's unit-disc fixture) through the general partitioner:
            # recursive coordinate bisection, then the same owned / ghost extraction
            z = np.load(os.path.join(ROOT, "tests", "golden", "unit_circle.npz"))
            lm = distributed.distribute_mesh(z["coords"], z["cells"], z["facets"], z["facet_tags"], dist, device="cpu")
            order = lm.rcb_order
            new_id = np.empty_like(order)
            new_id[order] = np.arange(order.size)
            om = meshes.Mesh(z["coords"][order], np.sort(new_id[z["cells"].astype(np.int64)], axis=1),
                             new_id[z["facets"].astype(np.int64)], z["facet_tags"])
            counts = np.diff(lm.global_offsets)
            assert counts.sum() == om.num_vertices and abs(int(counts[0]) - int(counts[1])) <= 1
            # a compact cut: the interface of a bisected disc is a few dozen vertices, not a band of them
            assert 0 < lm.num_vertices - lm.n_owned_vertices < 0.12 * om.num_vertices
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


@pytest.mark.parametrize("kind", ["cube", "square_ragged", "circle_rcb"])
def test_partition_and_halo_world2(kind):
    world = 2
    port = 29500 + (os.getpid() % 2000) + {"cube": 0, "square_ragged": 1, "circle_rcb": 2}[kind]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, kind, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)


# ----------------------------------------------------------------------------- distributed AMG (symbolic phase)
def _terms(seg, sl, sr, lhs, rhs):
    """numpy twin of cb_amg_spgemm_terms for bs = 1."""
    seg, sl, sr = seg.numpy(), sl.numpy().astype(np.int64), sr.numpy().astype(np.int64)
    prod = lhs[sl] * rhs[sr]
    ids = np.repeat(np.arange(seg.size - 1), np.diff(seg))
    return np.bincount(ids, weights=prod, minlength=seg.size - 1)


def _amg_worker(rank, world, port, n_mesh, agglomerate, rep_level, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp

        from cashocs_b200 import distributed
        from cashocs_b200._utils import amg
        from oracle import fem

        dist = distributed.DistContext(rank, world, transport="torch")
        lm = distributed.distributed_regular_mesh(n_mesh, dist, device="cpu")
        n, ncols = lm.n_owned_vertices, lm.num_vertices
        _, lvol, lG = fem.cell_geometry(lm.coords.numpy(), lm.cells.numpy())
        Aloc = fem.assemble_matrix(fem.stiffness_elements(lvol, lG) + fem.mass_elements(lvol, 3), lm.cells.numpy(), ncols).tocsr()[:n]
        Aloc.sort_indices()
        amg.AMGSymbolic.AGGLOMERATE = agglomerate
        sym = amg.AMGSymbolic(torch.from_numpy(Aloc.indptr.astype(np.int64)), torch.from_numpy(Aloc.indices.astype(np.int32)),
                              n, ncols, dist=dist)
        levels = sym.levels
        assert len(levels) >= 2
        kinds = ["rep" if L.rep_dist is not None else ("dist" if L.dist is not None else "local") for L in levels]
        assert kinds[0] == "dist" and "rep" in kinds and kinds.index("rep") == rep_level, (kinds, sym.sizes)
        assert all(k == "local" for k in kinds[kinds.index("rep") + 1:])
        rng = np.random.RandomState(5)
        val = Aloc.data.copy()
        gid = lm.global_vertex_ids.numpy()
        nglob = (n_mesh + 1) ** 3
        for li, L in enumerate(levels[:-1]):
            nxt = levels[li + 1]
            # prolongator values: arbitrary on owned rows, imported on ghost rows
            prp = L.p_rowptr.numpy()
            pval = np.ones(L.p_col.numel())
            pval[: L.nnzp_owned] = rng.rand(L.nnzp_owned) + 0.1
            if L.dist is not None:  # rows of the ghost nodes carry their owners' values
                pv = torch.from_numpy(pval)
                L.pplan.halo_exchange(pv, 1)
                pval = pv.numpy()
            ap = _terms(L.ap_seg, L.ap_sl, L.ap_sr, val, pval)
            ptval = pval[: L.nnzp_owned][L.pt_perm.numpy()]
            ac = _terms(L.ac_seg, L.ac_sl, L.ac_sr, ptval, ap)
            # global ids of the coarse nodes
            if L.dist is not None:
                counts = [None] * world
                td.all_gather_object(counts, L.nc)
                off = int(np.sum(counts[:rank]))
                ncg = int(np.sum(counts))
                if nxt.rep_dist is not None:
                    assert off == nxt.rep_off and nxt.n == ncg
                    cg = np.arange(off, off + L.nc)  # P of owned rows only references owned coarse nodes
                    glob = torch.zeros(nxt.nnzb, dtype=torch.float64)
                    glob[nxt.rep_pos] = torch.from_numpy(ac)
                    dist.allreduce(glob, "sum")
                    ac_next = glob.numpy()
                else:
                    v = torch.zeros(nxt.ncols, dtype=torch.float64)
                    v[: nxt.n] = torch.arange(off, off + nxt.n, dtype=torch.float64)
                    nxt.dist.halo_exchange(v, 1)
                    cg = v.numpy().round().astype(np.int64)
                    ac_next = ac
                rows = np.repeat(np.arange(L.n), np.diff(L.rowptr.numpy()))
                trip_a = (gid[rows], gid[L.col.numpy()], val)
                prow = np.repeat(np.arange(L.n), np.diff(prp[: L.n + 1]))
                assert L.p_col.numpy()[: L.nnzp_owned].max() < L.nc
                trip_p = (gid[prow], cg[L.p_col.numpy()[: L.nnzp_owned]], pval[: L.nnzp_owned])
                alla, allp = [None] * world, [None] * world
                td.all_gather_object(alla, trip_a)
                td.all_gather_object(allp, trip_p)
                cat = lambda ts, k: np.concatenate([t[k] for t in ts])
                Ag = sp.coo_matrix((cat(alla, 2), (cat(alla, 0), cat(alla, 1))), shape=(nglob, nglob)).tocsr()
                Pg = sp.coo_matrix((cat(allp, 2), (cat(allp, 0), cat(allp, 1))), shape=(nglob, ncg)).tocsr()
                ref = (Pg.T @ Ag @ Pg).tocsr()
                if nxt.rep_dist is not None:
                    got = sp.csr_matrix((ac_next, nxt.col.numpy(), nxt.rowptr.numpy()), shape=(ncg, ncg))
                else:
                    r2 = np.repeat(np.arange(nxt.n), np.diff(nxt.rowptr.numpy()))
                    allc = [None] * world
                    td.all_gather_object(allc, (cg[r2], cg[nxt.col.numpy()], ac_next))
                    got = sp.coo_matrix((cat(allc, 2), (cat(allc, 0), cat(allc, 1))), shape=(ncg, ncg)).tocsr()
                assert abs(got - ref).max() < 1e-13 * abs(ref).max()
                assert abs(got - got.T).max() < 1e-13 * abs(ref).max()
                gid, nglob = cg, ncg
            else:
                Al = sp.csr_matrix((val, L.col.numpy(), L.rowptr.numpy()), shape=(L.n, L.n))
                Pl = sp.csr_matrix((pval, L.p_col.numpy(), prp), shape=(L.n, L.nc))
                ref = (Pl.T @ Al @ Pl).tocsr()
                got = sp.csr_matrix((ac, nxt.col.numpy(), nxt.rowptr.numpy()), shape=(L.nc, L.nc))
                assert abs(got - ref).max() < 1e-13 * abs(ref).max()
                ac_next = ac
            val = ac_next
        results[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback

        results[rank] = traceback.format_exc() + str(exc)
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("n_mesh,agglomerate,rep_level", [(14, 20000, 1), (16, 50, 2)])
def test_distributed_amg_hierarchy_world2(n_mesh, agglomerate, rep_level):
    """Rank-local aggregates + imported prolongator rows of the ghost nodes + replicated coarse tail
    give the exact global Galerkin operators P^T A P (symmetric) on every level."""
    world = 2
    port = 31500 + (os.getpid() % 2000) + rep_level
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_amg_worker, args=(world, port, n_mesh, agglomerate, rep_level, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)


# ----------------------------------------------------------------------------- P2 dof numbering on a partition
def _p2_worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cashocs_b200 import distributed
        from cashocs_b200._forms import function_space
        from oracle import fem, meshes, p2

        dist = distributed.DistContext(rank, world, transport="torch")
        om = meshes.regular_mesh(4, 1.0, 1.0, 1.0)
        lm = distributed.distributed_regular_mesh(4, dist, device="cpu")
        num = function_space.p2_numbering(lm)
        nvg = om.num_vertices
        gdofs, gedges = p2.cell_dofs(om.cells, nvg)
        nd_glob = nvg + len(gedges)
        gid = lm.global_vertex_ids.numpy()
        # global dof id of every local dof
        gkey = gedges[:, 0] * nvg + gedges[:, 1]
        ek = num["edge_gkey"].numpy()
        eglob = nvg + np.searchsorted(gkey, ek[:, 0] * nvg + ek[:, 1])
        assert np.array_equal(gkey[eglob - nvg], ek[:, 0] * nvg + ek[:, 1])
        g_of = np.empty(num["ndofs"], dtype=np.int64)
        g_of[num["vertex_dof"].numpy()] = gid
        g_of[num["edge_dof"].numpy()] = eglob
        n_owned = num["n_owned"]
        # every global dof is owned by exactly one rank
        owned = [None] * world
        td.all_gather_object(owned, g_of[:n_owned])
        allowned = np.concatenate(owned)
        assert allowned.size == nd_glob and np.unique(allowned).size == nd_glob
        # local cell dofs == the oracle's global cell dofs of the same cells (as sets per cell, local edge order kept)
        gcells = gid[lm.cells.numpy()]
        order = {tuple(sorted(c)): i for i, c in enumerate(map(tuple, om.cells))}
        ci = np.array([order[tuple(sorted(c))] for c in map(tuple, gcells)])
        assert np.array_equal(np.sort(g_of[num["cell_dofs"].numpy()], axis=1), np.sort(gdofs[ci], axis=1))
        # halo exchange over dofs
        plan = num["dist"]
        xg = np.random.RandomState(3).rand(nd_glob)
        x = torch.zeros(num["ndofs"], dtype=torch.float64)
        x[:n_owned] = torch.from_numpy(xg[g_of[:n_owned]])
        plan.halo_exchange(x, 1)
        assert np.array_equal(x.numpy(), xg[g_of])
        # owned rows of the locally assembled P2 operator == rows of the global operator
        _, vol, G = fem.cell_geometry(om.coords, om.cells)
        A = fem.assemble_matrix(p2.stiffness_elements(vol, G) + p2.mass_elements(vol, 3), gdofs, nd_glob)
        _, lvol, lG = fem.cell_geometry(lm.coords.numpy(), lm.cells.numpy())
        Aloc = fem.assemble_matrix(p2.stiffness_elements(lvol, lG) + p2.mass_elements(lvol, 3), num["cell_dofs"].numpy().astype(np.int64),
                                   num["ndofs"])
        y = (Aloc @ xg[g_of])[:n_owned]
        assert np.allclose(y, (A @ xg)[g_of[:n_owned]], rtol=1e-12, atol=1e-14)
        results[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback

        results[rank] = traceback.format_exc() + str(exc)
    finally:
        td.destroy_process_group()


def test_p2_numbering_world2():
    """P2 dofs on a 2-rank partition: unique ownership, halo plan over vertex + edge dofs, owned rows of the
    locally assembled operator equal the global operator."""
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_p2_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)


# ----------------------------------------------------------------------------- output on a partitioned mesh
def _io_worker(rank, world, port, tmp, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types

        import cashocs_b200 as cb
        from cashocs_b200 import distributed
        from cashocs_b200.io import managers
        from cashocs_b200.io import mesh as iomesh

        dist = distributed.DistContext(rank, world, transport="torch")
        src = os.path.join(ROOT, "tests", "golden", "tiny_square.msh")
        g = cb.import_mesh(src, device="cpu")
        off = distributed.vertex_offsets(g.num_vertices, world)
        lm = distributed.partition_mesh(g.coords, g.cells, {k: g.boundary_vertices([k]) for k in g.boundary_markers}, off, dist)
        # every rank moves its own copy of the vertices by a function of the global position
        lm.coords.mul_(2.0).add_(0.125)
        want = g.coords.numpy() * 2.0 + 0.125
        pts = iomesh.gather_coordinates(lm)
        if rank == 0:
            assert np.array_equal(pts, want)
        else:
            assert pts is None
        out = os.path.join(tmp, "moved.msh")
        iomesh.write_out_mesh(lm, src, out)  # rank 0 writes, everybody waits
        coords = iomesh.read_gmsh(out)[0]
        assert np.array_equal(coords, want)
        # output managers: files are written by rank 0 only
        cfg = cb.Config()
        cfg.set("Output", "result_dir", os.path.join(tmp, "res"))
        cfg.set("Output", "save_txt", "True")
        om = managers.OutputManager(types.SimpleNamespace(config=cfg, mesh=lm))
        state = {"iteration": 0, "objective_value": 1.5, "relative_norm": 1.0, "gradient_norm": 2.0, "stepsize": 1.0}
        om.output(state)
        om.post_process(state)
        td.barrier()
        assert sorted(os.listdir(os.path.join(tmp, "res"))) == ["history.json", "history.txt"]
        assert open(os.path.join(tmp, "res", "history.txt")).read().count("1.500e+00") == 1
        results[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback

        results[rank] = traceback.format_exc() + str(exc)
    finally:
        td.destroy_process_group()


def test_mesh_export_and_output_world2(tmp_path):
    """``io.gather_coordinates`` / ``io.write_out_mesh`` / the output managers on a 2-rank partition: coordinates are
    gathered by global vertex id on rank 0, which alone writes the files (``cashocs/io/mesh.py:403-431,526-557``)."""
    world = 2
    port = 35500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_io_worker, args=(world, port, str(tmp_path), results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)


# ----------------------------------------------------------------------------- boundary facets on a partition
def _facet_worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cashocs_b200 as cb
        from cashocs_b200 import distributed
        from oracle import meshes, pipeline

        dist = distributed.DistContext(rank, world, transport="torch")
        for dim in (2, 3):
            if dim == 3:
                om = meshes.regular_mesh(4, 1.0, 1.0, 1.0)
                lm = distributed.distributed_regular_mesh(4, dist, device="cpu")
                markers = (1, 2, 3, 4, 5, 6)
            else:
                om = meshes.regular_mesh(7)
                g = cb.regular_mesh(7, device="cpu")
                off = distributed.vertex_offsets(g.num_vertices, world)
                lm = distributed.partition_mesh(g.coords, g.cells, {k: g.boundary_vertices([k]) for k in g.boundary_markers}, off, dist)
                markers = (1, 2, 3, 4)
            ref = pipeline.PoissonShapeProblem(om, bc_markers=markers, c_state=0.0,
                                               config=pipeline.ShapeConfig(shape_bdry_def=markers)).boundary_facets()
            want = {tuple(f) for f in ref.tolist()}
            gid = lm.global_vertex_ids.numpy()
            facets, counted = lm.boundary_facets_local()
            glob = np.sort(gid[facets.numpy()], axis=1)
            mine = {tuple(f) for f in glob.tolist()}
            lo, hi = int(lm.global_offsets[rank]), int(lm.global_offsets[rank + 1])
            # exactly the true boundary facets that touch an owned vertex
            assert mine == {f for f in want if any(lo <= v < hi for v in f)}
            # every boundary facet is counted by exactly one rank: the owner of its lowest vertex
            cnt = {tuple(f) for f in glob[counted.numpy()].tolist()}
            assert cnt == {f for f in want if lo <= min(f) < hi}
            total = torch.tensor([float(len(cnt))], dtype=torch.float64)
            dist.allreduce(total, "sum")
            assert int(total.item()) == len(want)
        results[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback

        results[rank] = traceback.format_exc() + str(exc)
    finally:
        td.destroy_process_group()


def test_boundary_facets_world2():
    """``Mesh.boundary_facets_local`` on a 2-rank partition (2-D ragged and 3-D slabs): the facets with exactly one
    local cell and an owned vertex are the true boundary facets touching this rank; the 'counted' subsets partition
    the global boundary."""
    world = 2
    port = 37500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_facet_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}, dict(results)
