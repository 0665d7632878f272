#!/usr/bin/env python3
"""Weak scaling of the fine-mesh matrix-free SIP vmult (LaplaceOperatorDG semantics, the operator of
examples/matrix_free_agglo.cc: 64^3 cells, FE_DGQ(2) PER GPU), one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29520 tools/run_fine_mf_scaling.py [--cells 64] [--degree 2]

The mesh is [0,1]^2 x [0,N] with cells^2 x (cells N) hexes, every cell its own element, z-slabs as
shards; one ghost layer of cells per cut travels before every apply over NVLink peer memory
(pd_peer_*), NCCL timed beside it.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import polydeal_b200 as pdl
from polydeal_b200 import distributed as pdd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=64)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mesh", default="morton", choices=["morton", "lex"])
    ap.add_argument("--partition", default="metis", choices=["metis", "slabs"])
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c, p = args.cells, args.degree
    t0 = time.time()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from pd_workloads import morton_numbered_box

    if args.mesh == "morton" and world > 1:
        grid = morton_numbered_box(pdl, c, c, c * world, (1.0, 1.0, float(world)))
    elif args.mesh == "morton":
        grid = pdl.Grid.hyper_cube(3, 0.0, 1.0, c.bit_length() - 1)
    else:
        grid = pdl.Grid.structured(3, (c, c, c * world), 0.0, (1.0, 1.0, float(world)), order=1)
    ah = pdl.AgglomerationHandler(grid)
    ah.define_agglomerates(np.arange(grid.n_cells, dtype=np.int32).reshape(-1, 1))
    ah.initialize_fe_values(p + 1)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, p)
    C = max(p, 1) * (p + 1.0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    if world > 1:
        if args.partition == "metis":
            owner = pdd.partition_by_metis(ah, world)
        else:  # z-slabs
            v, cv, _ = grid.arrays()
            zc = v[cv].mean(axis=1)[:, 2]
            owner = np.minimum((zc * 1.0).astype(np.int32), world - 1)
        dop = pdd.DistributedSIPOperator(ah, owner, rank, penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT)
        op, part = dop.op, dop.part
    else:
        op = pdl.SIPOperator(ah.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT), keepalive=ah)
        part = None
    op.set_stream(stream.cuda_stream)
    assert op.matrix_free_available
    t_host = time.time() - t0
    n_own, n_src = op.m(), op.n_source_dofs
    x = torch.from_numpy(np.sin(0.37 * np.arange(n_src)) + 0.01 * (np.arange(n_src) % 7)).cuda()
    y = torch.empty(n_own, dtype=torch.float64, device="cuda")
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    peer = pdd.PeerExchange(part, op) if part is not None else None

    def apply(use_peer):
        if part is not None and use_peer:
            peer.vmult(y, x, pdl.VMULT_MATRIX_FREE)  # exchange overlapped with the cells that need no ghosts
            return
        if part is not None:
            pdd.exchange_ghost_values(part, x)
        op.vmult_ptr(y.data_ptr(), x.data_ptr(), pdl.VMULT_MATRIX_FREE)

    def timed(use_peer):
        # `steps` applies back to back, as inside a solver (no barrier between them: a barrier per apply
        # would put the ranks' arrival skew into every sample).  Source + destination + stencil records are
        # larger than the 126 MB L2, so consecutive applies do not find their inputs cached.
        for _ in range(3):
            apply(use_peer)
        flush.zero_()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            apply(use_peer)
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b) / args.steps

    t_peer = timed(True)
    t_nccl = timed(False) if world > 1 else t_peer
    breakdown = {}
    if world > 1:  # where the time goes: the exchange alone, the local apply alone (ghosts as they are)
        def batch(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(args.steps):
                fn()
            b.record(stream)
            b.synchronize()
            return a.elapsed_time(b) / args.steps
        breakdown = {"exchange_only_ms": batch(lambda: peer.exchange(x)),
                     "local_apply_only_ms": batch(lambda: op.vmult_ptr(y.data_ptr(), x.data_ptr(), pdl.VMULT_MATRIX_FREE))}
        for key, val in (("interior_part_only_ms", "1"), ("exchange_plus_boundary_part_ms", "2")):
            os.environ["PD_PEER_DEBUG_PART"] = val
            breakdown[key] = batch(lambda: peer.vmult(y, x, pdl.VMULT_MATRIX_FREE))
            del os.environ["PD_PEER_DEBUG_PART"]
        breakdown["owned_cells"] = int(part.n_owned)
        breakdown["ghost_cells"] = int(part.n_ghost)
    checksum = float(y.sum())
    t = torch.tensor([t_peer, t_nccl, checksum], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        t_peer, t_nccl, checksum = float(tm[0]), float(tm[1]), float(ts[2])
        assert peer.ok()
        dist.barrier()
        peer.close()
    if rank == 0:
        N = world * c**3 * (p + 1) ** 3
        print(json.dumps({
            "metric": "fine-mesh matrix-free SIP vmult GDoF/s (LaplaceOperatorDG, Cartesian hexes)", "n_gpus": world,
            "cells_per_gpu": c**3, "degree": p, "n_dofs": N, "ghost_cells_per_cut": c * c,
            "ms": t_peer, "value": N / (t_peer * 1e-3) / 1e9, "unit": "GDoF/s", "scaling": "weak",
            "ms_with_nccl_exchange": t_nccl, "value_with_nccl_exchange": N / (t_nccl * 1e-3) / 1e9,
            "exchange": "NVLink peer memory (pd_peer_*)" if world > 1 else "none", "host_setup_s": t_host,
            "breakdown_rank0": breakdown,
            "l2": "inputs (src + dst + stencil records, 163 MB per GPU at 64^3 DGQ2) larger than L2; applies back to back", "checksum": checksum}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
