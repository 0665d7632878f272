"""CPU tests (gloo, world_size 2 and 3) of the sharding host logic: partition, local
descriptors with ghost polytopes, exchange plan, ghost-value exchange, and the index
plumbing of the distributed vmult (emulated with the oracle's global matrix)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dim, n, shape, p, seed=3):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import pyoracle as po
    from pd_helpers import groups_for, oracle_handler, product_handler

    ogrid = po.Grid(dim, n, 0.0, 1.0, 1)
    groups = groups_for(shape, dim, n, ogrid, seed)
    _, oah = oracle_handler(dim, n, groups, p, p + 1, order=1)
    _, pah = product_handler(oah.grid, groups, p, p + 1)
    return po, oah, pah


def _worker(rank, world, port, dim, n, shape, p, results, partitioner="blocks"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from polydeal_b200 import distributed as pdd

        po, oah, pah = _build(dim, n, shape, p)
        # blocks: contiguous ranges of the DoF-block order; metis: METIS on the polytope adjacency graph,
        # ownership then is scattered over the numbering
        owner = pdd.partition_by_blocks(pah, world) if partitioner == "blocks" else pdd.partition_by_metis(pah, world)
        assert len(set(owner.tolist())) == world
        part = pdd.LocalPart(pah, owner, rank)
        nd = part.n
        # every polytope is owned exactly once; local block order = global order restricted
        assert (np.diff(part.owned_global_block) > 0).all()
        cnt = torch.tensor([part.n_owned], dtype=torch.int64)
        dist.all_reduce(cnt)
        assert int(cnt) == pah.n_polytopes
        # ghosts: exactly the neighbours owned elsewhere, grouped by owner
        expect = set()
        for lp in range(part.n_owned):
            pglob = int(part.local_poly_global[lp])
            for f in range(pah.n_faces(pglob)):
                q = pah.neighbor(pglob, f)
                if q >= 0 and owner[q] != rank:
                    expect.add(int(pah.get_dof_indices(q)[0]) // nd)
        assert set(part.ghost_global_block.tolist()) == expect
        assert (np.diff(part.ghost_owner) >= 0).all()
        # send/recv counts agree pairwise
        sc = torch.tensor(part.send_counts)
        rc_all = [torch.zeros_like(sc) for _ in range(world)]
        dist.all_gather(rc_all, torch.tensor(part.recv_counts))
        for s in range(world):
            assert int(rc_all[s][rank]) == int(sc[s]), (rank, s)
        # ghost exchange delivers the owners' values: fill owned DoFs with their global index
        x_full = torch.full((part.n_local_dofs,), -1.0, dtype=torch.float64)
        x_full[: part.n_owned_dofs] = torch.from_numpy(part.owned_global_dofs().astype(np.float64))
        pdd.exchange_ghost_values(part, x_full)
        np.testing.assert_array_equal(x_full[part.n_owned_dofs:].numpy(), part.ghost_global_dofs().astype(np.float64))
        # distributed vmult plumbing: rows of the oracle's global matrix restricted to my blocks,
        # columns renumbered (owned | ghost), times [x_owned ; x_ghost] == (A x)_owned
        A = po.assemble_dg_matrix(oah, degree=p).scipy().tocsr()
        xg = np.sin(0.37 * np.arange(A.shape[0]))
        rows = part.owned_global_dofs()
        cols = np.concatenate([rows, part.ghost_global_dofs()])
        A_loc = A[rows][:, cols]
        # nothing outside owned+ghost columns may be needed by my rows
        assert abs(A[rows]).sum() == pytest.approx(abs(A_loc).sum(), rel=1e-14)
        x_loc = torch.zeros(part.n_local_dofs, dtype=torch.float64)
        x_loc[: part.n_owned_dofs] = torch.from_numpy(xg[rows])
        pdd.exchange_ghost_values(part, x_loc)
        y = A_loc @ x_loc.numpy()
        np.testing.assert_allclose(y, (A @ xg)[rows], rtol=0, atol=1e-12 * np.abs(A @ xg).max())
        # the local block pattern equals the global pattern restricted and renumbered
        d = part.desc
        brow = np.ctypeslib.as_array(d.brow_ptr, (part.n_owned + 1,))
        bcol = np.ctypeslib.as_array(d.bcol_idx, (int(brow[-1]),))
        loc2glob = np.concatenate([part.owned_global_block, part.ghost_global_block])
        for r in range(part.n_owned):
            got = sorted(loc2glob[bcol[brow[r]:brow[r + 1]]].tolist())
            gb = int(part.owned_global_block[r])
            want = sorted(set((A[gb * nd].indices // nd).tolist()))
            assert got == want
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim,n,shape,p,partitioner", [(2, 2, 8, "blocks2", 1, "blocks"), (2, 3, 4, "random9", 1, "blocks"),
                                                             (3, 2, 8, "random11", 2, "blocks"), (2, 2, 8, "random12", 1, "metis")])
def test_sharding_host_logic_gloo(world, dim, n, shape, p, partitioner):
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dim, n, shape, p, results, partitioner), nprocs=world, join=True)
    assert dict(results) == {r: "ok" for r in range(world)}


def test_distributed_reference_goldens(goldens):
    """ghosted_bbox_01 / ghosted_dofs_01 / sparsity_distributed_tria (mpirun=3): [0,1]^2 refined twice on a
    parallel::distributed (p4est = Morton ranges) triangulation, agglomerates {0,1},{2,3} on rank 0,
    {4,5},{6,7},{8,9},{10,11} on rank 1, {12,13},{14,15} on rank 2.  What the reference prints about the ghosted
    neighbours -- bounding boxes seen from rank 1, DoF indices seen from rank 0, rank 0's rows of the
    distributed sparsity pattern -- must be what the local descriptor of that rank carries."""
    sys.path.insert(0, ROOT)
    import polydeal_b200 as pdl
    from polydeal_b200 import distributed as pdd

    owner = np.array([0, 0, 1, 1, 1, 1, 2, 2], dtype=np.int32)

    def handler(degree):
        grid = pdl.Grid.hyper_cube(2, 0.0, 1.0, 2)
        ah = pdl.AgglomerationHandler(grid)
        for k in range(8):
            ah.define_agglomerate([2 * k, 2 * k + 1])
        ah.initialize_fe_values(2 * degree + 1)
        ah.distribute_agglomerated_dofs(pdl.FE_DGQ, degree)
        return ah

    def ghost_walk(ah, part, rank):
        """(local index, [local index of the ghost polytope behind every face whose neighbour is remote])"""
        loc = {int(g): i for i, g in enumerate(part.local_poly_global)}
        first = int(np.nonzero(owner == rank)[0][0])
        out = []
        for p in np.nonzero(owner == rank)[0]:
            ghosts = [loc[ah.neighbor(int(p), f)] for f in range(ah.n_faces(int(p)))
                      if not ah.at_boundary(int(p), f) and owner[ah.neighbor(int(p), f)] != rank]
            out.append((int(p) - first, ghosts))
        return out

    # ghosted_bbox_01: FE_DGQ(0), seen from rank 1
    ah = handler(0)
    part = pdd.LocalPart(ah, owner, 1)
    bbox = np.ctypeslib.as_array(part.desc.bbox, (part.n_owned + part.n_ghost, 2, 2))
    gold = goldens["ghosted_bbox_01"]
    walk = ghost_walk(ah, part, 1)
    assert len(walk) == len(gold)
    for (li, ghosts), g in zip(walk, gold):
        assert (li, 1) == (g["local_index"], g["rank"]) and len(ghosts) == len(g["ghosts"])
        for lp, corners in zip(ghosts, g["ghosts"]):
            assert lp >= part.n_owned  # carried as a ghost
            assert bbox[lp].tolist() == corners
    # ghosted_dofs_01: FE_DGQ(1), seen from rank 0
    ah = handler(1)
    part = pdd.LocalPart(ah, owner, 0)
    gdofs = part.ghost_global_dofs().reshape(part.n_ghost, part.n)
    gold = goldens["ghosted_dofs_01"]
    walk = ghost_walk(ah, part, 0)
    assert len(walk) == len(gold)
    for (li, ghosts), g in zip(walk, gold):
        assert (li, 0) == (g["local_index"], g["rank"]) and len(ghosts) == len(g["ghosts"])
        for lp, dofs in zip(ghosts, g["ghosts"]):
            assert gdofs[lp - part.n_owned].tolist() == [int(d[0]) for d in dofs]
    # sparsity_distributed_tria: rank 0's rows of the distributed pattern, from the LOCAL descriptor
    d = part.desc
    brow = np.ctypeslib.as_array(d.brow_ptr, (d.n_block_rows + 1,))
    bcol = np.ctypeslib.as_array(d.bcol_idx, (int(brow[-1]),))
    gblock = np.concatenate([part.owned_global_block, part.ghost_global_block])
    got = set()
    for b in range(d.n_block_rows):
        for e in range(brow[b], brow[b + 1]):
            for i in range(part.n):
                for j in range(part.n):
                    got.add((int(gblock[b]) * part.n + i, int(gblock[bcol[e]]) * part.n + j))
    assert got == {tuple(rc) for rc in goldens["sparsity_distributed_tria"]}


def _subface_midpoints(desc, f):
    """geometric identity of the sub-faces of local interface f: the mid-points of their vertices, in list order"""
    dim = desc.dim
    vpc = 1 << dim
    verts = np.ctypeslib.as_array(desc.verts, (desc.n_verts, dim))
    cv = np.ctypeslib.as_array(desc.cell_verts, (desc.n_cells, vpc))
    ptr = np.ctypeslib.as_array(desc.iface_sub_ptr, (desc.n_ifaces + 1,))
    sc_ = np.ctypeslib.as_array(desc.sub_cell, (int(ptr[-1]),))
    sf = np.ctypeslib.as_array(desc.sub_face, (int(ptr[-1]),))
    out = []
    for s in range(ptr[f], ptr[f + 1]):
        axis, side = int(sf[s]) // 2, int(sf[s]) % 2  # deal.II face numbering: 2 * direction + side
        face_verts = [v for v in range(vpc) if ((v >> axis) & 1) == side]
        out.append(verts[cv[sc_[s]][face_verts]].mean(axis=0))
    return np.array(out)


@pytest.mark.parametrize("name,n_ref,groups,owner", [
    # reinit_ghosted_neighbor_01.cc:62-118: K0 | K1, K2 | K3 on ranks 0 | 1 | 2
    ("reinit_ghosted_neighbor_01", 2, [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]], [0, 1, 1, 2]),
    # reinit_ghosted_neighbor_02.cc:69-146 and locally_owned_polytope_05.cc:68-150: eight two-cell polytopes
    ("reinit_ghosted_neighbor_02", 2, [[2 * k, 2 * k + 1] for k in range(8)], [0, 0, 1, 1, 1, 1, 2, 2]),
    # locally_owned_polytope_01.cc:43-53 (mpirun=2): every cell its own polytope, two ranks of eight cells
    ("locally_owned_polytope_01", 2, [[c] for c in range(16)], [0] * 8 + [1] * 8),
])
def test_cut_interfaces_seen_from_both_ranks(name, n_ref, groups, owner):
    """test/polydeal/reinit_ghosted_neighbor_01/02 (mpirun=3, golden "Ok"): on an interface cut by the partition the
    face quadrature points and JxW a rank computes must be those its neighbour rank holds for the same interface
    (source/agglomeration_handler.cc:531-618 ships them; here BOTH ranks evaluate the interface from their local
    descriptor, so the check is that the two descriptors list the same sub-faces in the same order -- the
    quadrature is generated from exactly these lists).  locally_owned_polytope_01 (mpirun=2): every rank owns eight
    polytopes with local indices 0..7; locally_owned_polytope_05 (mpirun=3, golden "Ok"): owned volumes sum to 1 and
    owned boundary faces to 4 (:163-210)."""
    sys.path.insert(0, ROOT)
    import polydeal_b200 as pdl
    from polydeal_b200 import distributed as pdd

    owner = np.array(owner, dtype=np.int32)
    grid = pdl.Grid.hyper_cube(2, 0.0, 1.0, n_ref)
    ah = pdl.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(1)  # FE_DGQ(0), QGauss(1) as in the reference tests
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, 0)
    world = int(owner.max()) + 1
    # every cut interface, from both sides (the descriptor arrays belong to the handler and are rewritten by the
    # next flatten, so each rank's view is read before the next one is made)
    seen = {}
    volume = perimeter = 0.0
    for r in range(world):
        part = pdd.LocalPart(ah, owner, r)
        # ownership: every polytope exactly once, local indices 0 .. n_owned-1 in the global order
        assert part.local_poly_global[: part.n_owned].tolist() == np.nonzero(owner == r)[0].tolist()
        if name == "locally_owned_polytope_01":
            assert part.n_owned == 8
        d = part.desc
        A = np.ctypeslib.as_array(d.iface_polyA, (d.n_ifaces,))
        B = np.ctypeslib.as_array(d.iface_polyB, (d.n_ifaces,))
        glob = part.local_poly_global
        for f in range(d.n_ifaces):
            mid = _subface_midpoints(d, f)
            if B[f] < 0:
                if A[f] < part.n_owned:  # boundary face of an owned polytope: sum of JxW = its length
                    sub = mid.shape[0]
                    perimeter += sub * (0.5 ** n_ref)
                continue
            key = tuple(sorted((int(glob[A[f]]), int(glob[B[f]]))))  # whichever side a rank lists first
            if owner[key[0]] != owner[key[1]]:
                seen.setdefault(key, {})[r] = mid
        for lp in range(part.n_owned):
            lo, hi = ah.bbox(int(glob[lp]))
            volume += len(ah.get_agglomerate(int(glob[lp]))) * 0.25 ** n_ref
    n_cut = 0
    for key, views in seen.items():
        assert set(views) == {int(owner[key[0]]), int(owner[key[1]])}, key  # both owners carry the interface
        a, b = views.values()
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-15, key  # same sub-faces, same order
        n_cut += 1
    # every face between polytopes of different ranks was found
    want = {(p, ah.neighbor(p, f)) for p in range(ah.n_polytopes) for f in range(ah.n_faces(p))
            if not ah.at_boundary(p, f) and owner[p] != owner[ah.neighbor(p, f)]}
    assert {tuple(sorted(k)) for k in seen} == {tuple(sorted(k)) for k in want} and n_cut == len(want) // 2
    assert abs(volume - 1.0) < 1e-15 and abs(perimeter - 4.0) < 1e-15


def _fused_plan_worker(rank, world, port, n, results):
    """One rank of the fused sharded fine-mesh apply's HOST side (csrc/pd_peer.cu: peer_create / peer_connect,
    csrc/pd_finemesh.cu: setup_fine_fused), with the all-gather PeerExchange does over gloo: where this rank will read
    its ghost cells in the owners' export buffers is where the owners publish them, and the tile plan built from those
    addresses' 16-byte phases obeys the kernel's contract (tools/fused_plan_check.py holds the emulation)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import fused_plan_check as fpc
        import polydeal_b200 as pdl
        from pd_workloads import build_handler
        from polydeal_b200 import distributed as pdd

        dim, p = 3, 2
        nd = (p + 1) ** dim
        ah = build_handler(pdl, dict(dim=dim, n=n, b=1, p=p, nq=p + 1), world)  # n x n x (n world) cells, Morton numbered
        owner = pdd.partition_by_metis(ah, world)
        part = pdd.LocalPart(ah, owner, rank, penalty_constant=6.0, h_rule=pdl.H_NORMAL_EXTENT)
        n_own, n_total = part.n_owned, part.n_owned + part.n_ghost
        # what PeerExchange.__init__ gathers: every rank's send offsets; here also the send lists themselves (global cells)
        send_ptr = np.zeros(world + 1, dtype=np.int64)
        send_ptr[1:] = np.cumsum(part.send_counts)
        sent_global = np.concatenate([part.owned_global_block[b] for b in part.send_blocks]) if world > 1 else np.zeros(0)
        all_ptr, all_sent = [None] * world, [None] * world
        dist.all_gather_object(all_ptr, send_ptr.tolist())
        dist.all_gather_object(all_sent, sent_global.tolist())
        recv_ptr = np.zeros(world + 1, dtype=np.int64)
        recv_ptr[1:] = np.cumsum(part.recv_counts)
        # ghost g of owner s is block remote_off[s] + (g - recv_ptr[s]) of s's send list (peer_create): the same cell
        par = np.zeros(n_total, dtype=np.uint8)
        par[:n_own] = (np.arange(n_own, dtype=np.int64) * nd) & 1
        n_owners = 0
        for s in range(world):
            cnt = int(recv_ptr[s + 1] - recv_ptr[s])
            if cnt == 0:
                continue
            n_owners += 1
            assert all_ptr[s][rank + 1] - all_ptr[s][rank] == cnt
            b = all_ptr[s][rank] + np.arange(cnt)
            np.testing.assert_array_equal(np.asarray(all_sent[s])[b], part.ghost_global_block[recv_ptr[s]:recv_ptr[s + 1]])
            par[n_own + recv_ptr[s] + np.arange(cnt)] = b & 1  # export_at: block b sits at phase b & 1 (odd n)
        assert n_owners >= (2 if world > 2 and 0 < rank < world - 1 else 1)
        # the plan, with the row budget cut so that tiles are split
        lib = fpc.host_lib()
        nbr = fpc.neighbour_table(part.desc, dim)
        key = fpc.morton_keys(part.desc, dim)
        order = np.argsort(key, kind="stable").astype(np.int32)
        assert (order == np.arange(n_own)).all()  # a rank's share of a Morton-numbered mesh is in curve order
        blk = key >> np.uint64(2 * dim)
        is_outer = np.isin(blk[order], np.unique(blk[(nbr >= n_own).any(axis=1)]))
        inner, outer = order[~is_outer], order[is_outer]
        assert len(inner) and len(outer)
        tf1, _ = fpc.tile_first_of(lib, inner, key, nbr, n_total, nd, dim)
        tf2, _ = fpc.tile_first_of(lib, outer, key, nbr, n_total, nd, dim)
        for budget in (128, 48):
            out = fpc.fused_plan(lib, inner, outer, tf1, tf2, nbr, n_total, nd, dim, par, budget)
            assert out["rc"] == 0 and out["max_rows"] <= budget, out
            assert 0 < out["first_ghost_tile"] < out["n_tiles"]
            if budget == 48:
                assert out["n_tiles"] > out["unsplit_tiles"] and out["unsplit_max_rows"] > 48
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_fused_fine_mesh_plan_addresses_gloo(world):
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_fused_plan_worker, args=(world, port, 16, results), nprocs=world, join=True)
    assert dict(results) == {r: "ok" for r in range(world)}
