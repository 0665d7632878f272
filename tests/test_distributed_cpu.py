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


def _worker(rank, world, port, dim, n, shape, p, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from polydeal_b200 import distributed as pdd

        po, oah, pah = _build(dim, n, shape, p)
        owner = pdd.partition_by_blocks(pah, world)
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


@pytest.mark.parametrize("world,dim,n,shape,p", [(2, 2, 8, "blocks2", 1), (2, 3, 4, "random9", 1), (3, 2, 8, "random11", 2)])
def test_sharding_host_logic_gloo(world, dim, n, shape, p):
    mgr = mp.Manager()
    results = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dim, n, shape, p, results), nprocs=world, join=True)
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
