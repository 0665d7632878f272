#!/usr/bin/env python3
"""Multi-process check of the sharded path on real GPUs (one rank per GPU, NCCL):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tests/run_distributed_check.py
Every rank assembles its share, then y = A x through DistributedSIPOperator.vmult with the
NCCL ghost exchange and with the NVLink peer-memory exchange; compared with the CPU oracle's
global result on the same input."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import polydeal_b200 as pdl
from oracle import pyoracle as po
from pd_helpers import groups_for, oracle_handler, product_handler, src_vector
from polydeal_b200 import distributed as pdd


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # PD_CHECK_SAME_DEVICE=1: all ranks share GPU 0 (CUDA IPC works between processes on one device;
    # their kernels are time-sliced), rendezvous over gloo -- lets a single-GPU test run exercise the
    # peer-memory collectives.  NCCL cannot put two ranks on one GPU, so its path is skipped there.
    same = os.environ.get("PD_CHECK_SAME_DEVICE") == "1"
    if same:
        local = 0
    torch.cuda.set_device(local)
    if same:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red_dev = "cpu" if same else "cuda"
    worst = 0.0
    solve_ref, cg_info = None, []
    for dim, n, shape, p in [(3, 8, "blocks2", 2), (3, 8, "random40", 1), (2, 16, "random23", 3)]:
        ogrid = po.Grid(dim, n, 0.0, 1.0, 1)
        groups = groups_for(shape, dim, n, ogrid, 5)
        _, oah = oracle_handler(dim, n, groups, p, p + 1, order=1)
        _, pah = product_handler(oah.grid, groups, p, p + 1)
        A = po.assemble_dg_matrix(oah, degree=p, n_threads=4).scipy().tocsr()
        x = src_vector(A.shape[0])
        y = A @ x
        owner = pdd.partition_by_blocks(pah, world)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            dop = pdd.DistributedSIPOperator(pah, owner, rank)
            dop.op.set_stream(stream.cuda_stream)
            dop.assemble()
            rows = dop.part.owned_global_dofs()
            xs = torch.from_numpy(x[rows]).cuda()
            yd = torch.empty_like(xs)
            err = 0.0
            if not same:
                dop.vmult(yd, xs)
                stream.synchronize()
                err = float(np.abs(yd.cpu().numpy() - y[rows]).max() / np.abs(y).max())
            # the same through NVLink peer memory (pd_peer_*): several applies back to back with a
            # changing source exercise the double-buffered epochs; ranks deliberately out of step
            dop.enable_peer_exchange()
            for k in range(1, 7):
                if (rank + k) % 3 == 0:
                    torch.cuda._sleep(2_000_000)
                dop.vmult(yd, xs * float(k))
                if k % 2 == 0:
                    stream.synchronize()
                    e2 = float(np.abs(yd.cpu().numpy() - k * y[rows]).max() / (k * np.abs(y).max()))
                    err = max(err, e2)
            stream.synchronize()
            assert dop.peer.ok(), "a neighbour did not publish in time"
            e2 = float(np.abs(yd.cpu().numpy() - 6 * y[rows]).max() / (6 * np.abs(y).max()))
            err = max(err, e2)
            # peer-memory all-reduce, then the sharded CG (exchange + all-reduce inside the CUDA graph)
            t4 = torch.tensor([1.0 + rank, 0.5 * rank, -2.0, 1e-3 * (rank + 1)], dtype=torch.float64, device="cuda")
            dop.peer.allreduce(t4)
            stream.synchronize()
            ranks = np.arange(world)
            ref4 = [np.sum(1.0 + ranks), np.sum(0.5 * ranks), -2.0 * world, np.sum(1e-3 * (ranks + 1))]
            assert np.allclose(t4.cpu().numpy(), ref4, rtol=1e-15, atol=0), (t4, ref4)
            if solve_ref is None:
                solve_ref = {}
            bs = torch.from_numpy(y[rows]).cuda()  # b = A x  =>  the solution is x
            xs0 = torch.zeros_like(bs)
            iters, relres = dop.peer.cg_solve(xs0, bs, max_iter=4000, rel_tol=1e-11)
            stream.synchronize()
            assert relres <= 1e-11, (iters, relres)
            e3 = float(np.abs(xs0.cpu().numpy() - x[rows]).max() / np.abs(x).max())
            assert e3 <= 1e-7, e3
            cg_info.append((iters, relres, e3))
            # sharded lambda_max (power iteration on D^-1 A) and Chebyshev smoother vs numpy on the global matrix
            dinv_g = 1.0 / A.diagonal()
            lam = dop.peer.estimate_lambda_max(40)
            import scipy.sparse as sp
            import scipy.sparse.linalg as spla
            Dh = sp.diags(np.sqrt(dinv_g))
            lam_true = float(spla.eigsh(Dh @ A @ Dh, k=1, which="LA", return_eigenvectors=False)[0])
            assert 0.8 * lam_true <= lam <= 1.0001 * lam_true, (lam, lam_true)
            lmax, rng_ = 1.2 * lam_true, 20.0
            lmin = lmax / rng_
            theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
            sigma1 = theta / delta
            rho = 1.0 / sigma1
            xg = np.zeros_like(y)
            dg = dinv_g * y / theta
            xg += dg
            for _ in range(1, 4):
                rho_new = 1.0 / (2 * sigma1 - rho)
                dg = rho_new * rho * dg + 2 * rho_new / delta * dinv_g * (y - A @ xg)
                xg += dg
                rho = rho_new
            xfull = torch.zeros(dop.part.n_local_dofs, dtype=torch.float64, device="cuda")
            dop.peer.chebyshev_smooth(xfull, bs, 4, lmax, rng_, zero_initial_guess=True)
            stream.synchronize()
            e4 = float(np.abs(xfull[: len(rows)].cpu().numpy() - xg[rows]).max() / np.abs(xg).max())
            assert e4 <= 1e-12, e4
            dist.barrier()
            dop.peer.close()
        t = torch.tensor([err], device=red_dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = max(worst, float(t))
        if rank == 0:
            print(f"case dim={dim} n={n} {shape} p={p}: world={world} max rel err {float(t):.2e}; sharded CG "
                  f"{cg_info[-1][0]} its, relres {cg_info[-1][1]:.1e}, solution err {cg_info[-1][2]:.1e}", flush=True)
    # fine-mesh matrix-free operator (every cell its own element), sharded: pd_peer_vmult applies the cells
    # without ghost neighbours while the ghost blocks travel on a second stream
    # (the last two cases: cells numbered along the Morton curve, as a p4est-distributed mesh has them on every rank,
    # uniform -> the FUSED apply: one kernel, boundary tiles wait for the owners' flags and read the ghost cells from
    # the owners' export buffers)
    # (order 3: METIS again with the fused plan's halo-row budget cut to 48 rows (PD_FINE_FUSED_MAX_ROWS), so that every
    # tile is split -- what happens to a few boundary tiles of a ragged METIS cut at bench.py's size, where unsplit they need
    # 129 / 130 of the gather's 128 rows)
    for dim, n, p, order in [(3, 8, 2, 1), (2, 16, 3, 1), (3, 16, 2, 0), (2, 64, 2, 0), (3, 16, 2, 2), (3, 16, 2, 3)]:
        metis, split_tiles = order >= 2, order == 3
        order = 0 if metis else order
        os.environ.pop("PD_FINE_FUSED_MAX_ROWS", None)
        if split_tiles:
            os.environ["PD_FINE_FUSED_MAX_ROWS"] = "48"
        ogrid = po.Grid(dim, n, 0.0, 1.0, order)
        groups = [[c] for c in range(ogrid.n_cells)]
        _, oah = oracle_handler(dim, n, groups, p, p + 1, order=order)
        _, pah = product_handler(oah.grid, groups, p, p + 1)
        C_ = max(p, 1) * (p + 1.0)
        A = po.assemble_dg_matrix(oah, penalty_constant=C_, h_rule=po.H_NORMAL_EXTENT, n_threads=4).scipy().tocsr()
        x = src_vector(A.shape[0])
        y = A @ x
        owner = pdd.partition_by_metis(pah, world) if metis else pdd.partition_by_blocks(pah, world)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            dop = pdd.DistributedSIPOperator(pah, owner, rank, penalty_constant=C_, h_rule=pdl.H_NORMAL_EXTENT)
            dop.op.set_stream(stream.cuda_stream)
            assert dop.op.matrix_free_available
            dop.enable_peer_exchange()
            rows = dop.part.owned_global_dofs()
            xs = torch.from_numpy(x[rows]).cuda()
            yd = torch.empty_like(xs)
            err = 0.0
            for k in range(1, 5):
                dop.vmult(yd, xs * float(k), mode=pdl.VMULT_MATRIX_FREE)
                stream.synchronize()
                err = max(err, float(np.abs(yd.cpu().numpy() - k * y[rows]).max() / (k * np.abs(y).max())))
            fused = dop.peer.fused
            if order == 0 and world == 2:
                assert fused and dop.op.fine_kernel_last == 3, (fused, dop.op.fine_kernel_last)
            if split_tiles and fused:  # (a whole 4^3 block needs 96 rows: every tile was split)
                assert dop.peer.fused_tiles >= 2 * (dop.part.n_owned // 64), (dop.peer.fused_tiles, dop.part.n_owned)
            # applies back to back with ranks out of step (double-buffered export slots, epochs), then vmult_add
            for k in range(5, 11):
                if (rank + k) % 3 == 0:
                    torch.cuda._sleep(2_000_000)
                dop.vmult(yd, xs * float(k), mode=pdl.VMULT_MATRIX_FREE)
            stream.synchronize()
            err = max(err, float(np.abs(yd.cpu().numpy() - 10 * y[rows]).max() / (10 * np.abs(y).max())))
            assert dop.peer.ok()
            dist.barrier()
            dop.peer.close()
        t = torch.tensor([err], device=red_dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = max(worst, float(t))
        if rank == 0:
            print(f"fine-mesh MF dim={dim} n={n} p={p} order={order} metis={metis} split_tiles={split_tiles}: world={world} "
                  f"fused={fused} max rel err {float(t):.2e}", flush=True)
    os.environ.pop("PD_FINE_FUSED_MAX_ROWS", None)
    dist.barrier()
    dist.destroy_process_group()
    assert worst <= 1e-12, worst
    if rank == 0:
        print("DISTRIBUTED CHECK OK")


if __name__ == "__main__":
    main()
