#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path on B200.

Workload (BASELINE.json configs[1]): unit cube, 64^3 hexes, 512
polyhedra of 8^3 cells (the `blocks` shape an R-tree of fan-out 8 extracts at level 3),
FE_DGQ<3>(2), QGauss(3), SIP assembly of stiffness + penalty + boundary terms with the
library penalty C = 10 (p+dim)(p+1) = 150 (include/poly_utils.h:2018-2019).
A step = agglomerated quadrature + volume + face + diagonal-gather kernels, from the
flattened agglomeration resident in HBM to the finished scalar-CSR values in HBM.

Metric: polytope DoFs assembled per second (whole job, all ranks).  The companion
metric of BASELINE.json, SIP vmult GDoF/s, is reported twice: "vmult" = the block-CSR
apply of the matrix just assembled (what the reference's solvers call on agglomerated
levels), "mf_vmult" (N = 1) = the matrix-free sum-factorised LaplaceOperatorDG on the
64^3 DGQ2 fine mesh of examples/matrix_free_agglo.cc, each with its own HBM roofline.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...   times the CPU oracle restating the reference's
                                         assembly on the host cores (the reference itself
                                         needs deal.II and cannot be built here)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

METRIC = "polytope DoFs assembled/s"
UNIT = "DoF/s"
DIM, N_CELLS_1D, BLOCK, DEGREE, NQ = 3, 64, 8, 2, 3
WORKLOAD = "B: unit cube 64^3 hexes -> 512 polyhedra (8^3 blocks, R-tree level 3), FE_DGQ(2), QGauss(3), SIP stiffness+penalty+boundary, C=150"


def block_groups(n, b):
    import pd_scenarios as sc

    return sc.block_partition(DIM, n, b)


def lex_block_groups(nx, ny, nz, b):
    """b^3 blocks of an nx x ny x nz lexicographic grid, cells of a block in active-cell order."""
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cell = (k * ny + j) * nx + i
    part = ((k // b) * (ny // b) + (j // b)) * (nx // b) + (i // b)
    order = np.lexsort((cell, part))
    return cell[order].astype(np.int32).reshape(-1, b**3)


def read_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            peaks["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            peaks["hbm_src"] = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    peaks["fp64_tflops"] = 37.1
    peaks["fp64_src"] = "fallback 37.1 (nominal 148 SM x 64 FMA/clk x 1.965 GHz)"
    p = os.path.join(ROOT, "profiles", "FP64_PEAK.json")
    if os.path.exists(p):
        try:
            peaks["fp64_tflops"] = float(json.load(open(p))["fp64_tflops"])
            peaks["fp64_src"] = "measured DMMA m8n8k4 (profiles/FP64_PEAK.json; MEASURED_PEAKS.json has no FP64 figure)"
        except Exception:
            pass
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        load = [v for v in sm if mx and v > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def pinned_copy_of_desc(desc):
    """Copy every descriptor array into pinned host memory; returns (new desc, keepalive, bytes)."""
    import torch

    from polydeal_b200 import _capi as K

    d = K.MeshDesc()
    keep, total = [], 0
    nsc = desc.poly_subcell_ptr[desc.n_polytopes]
    nsf = desc.iface_sub_ptr[desc.n_ifaces] if desc.n_ifaces else 0
    nblk = desc.brow_ptr[desc.n_block_rows]
    sizes = {
        "verts": desc.n_verts * desc.dim, "cell_verts": desc.n_cells << desc.dim,
        "poly_subcell_ptr": desc.n_polytopes + 1, "poly_subcell_idx": nsc, "bbox": desc.n_polytopes * 2 * desc.dim,
        "dof_block": desc.n_polytopes, "iface_polyA": desc.n_ifaces, "iface_polyB": desc.n_ifaces,
        "iface_sub_ptr": desc.n_ifaces + 1, "sub_cell": nsf, "sub_face": nsf, "sub_sigma": nsf,
        "brow_ptr": desc.n_block_rows + 1, "bcol_idx": nblk,
    }
    for name, ctype in K.MeshDesc._fields_:
        v = getattr(desc, name)
        if name in sizes:
            n = int(sizes[name])
            elem = ctype._type_
            dt = {C.c_double: torch.float64, C.c_int32: torch.int32, C.c_int64: torch.int64}[elem]
            t = torch.empty(max(n, 1), dtype=dt).pin_memory()
            src = np.ctypeslib.as_array(v, (n,)) if n else np.zeros(0)
            t[:n].copy_(torch.from_numpy(np.array(src)))
            keep.append(t)
            total += n * C.sizeof(elem)
            setattr(d, name, C.cast(t.data_ptr(), ctype))
        else:
            setattr(d, name, v)
    return d, keep, total


def run_gpu(args):
    import torch

    import polydeal_b200 as pdl
    from polydeal_b200 import _capi as K

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- build the agglomeration (host, untimed) --------------------------------------
    # Weak scaling: every rank owns one config-B box (64^3 cells, 512 polyhedra); the path
    # shards over polytopes with no data-path collective in assembly.
    part = None
    if world == 1:
        grid = pdl.Grid.hyper_cube(DIM, 0.0, 1.0, N_CELLS_1D.bit_length() - 1)
        ah = pdl.AgglomerationHandler(grid)
        for g in block_groups(N_CELLS_1D, BLOCK):
            ah.define_agglomerate(g)
        ah.initialize_fe_values(NQ)
        ah.distribute_agglomerated_dofs(pdl.FE_DGQ, DEGREE)
        desc0 = ah.flatten()  # library penalty, visit by id
    else:
        # N boxes stacked along z: [0,1]^2 x [0,N], 64 x 64 x 64N cubic cells, 8^3 blocks,
        # sharded into z-slabs of 512 polyhedra; the cut interfaces are evaluated by both
        # neighbours from one-time ghost geometry (owner-computes-rows)
        from polydeal_b200 import distributed as pdd

        nz = N_CELLS_1D * world
        grid = pdl.Grid.structured(DIM, (N_CELLS_1D, N_CELLS_1D, nz), 0.0, (1.0, 1.0, float(world)), order=1)
        ah = pdl.AgglomerationHandler(grid)
        for g in lex_block_groups(N_CELLS_1D, N_CELLS_1D, nz, BLOCK):
            ah.define_agglomerate(g)
        ah.initialize_fe_values(NQ)
        ah.distribute_agglomerated_dofs(pdl.FE_DGQ, DEGREE)
        owner = pdd.partition_by_blocks(ah, world)
        part = pdd.LocalPart(ah, owner, rank)
        desc0 = part.desc
    desc, keep, h2d_bytes = pinned_copy_of_desc(desc0)
    op = pdl.SIPOperator(desc, keepalive=(ah, keep, part))
    # all work and all timing events go to ONE explicit non-default stream (the legacy
    # default stream has handle 0, which pd_set_stream reads as "use the handle's own")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    op.set_stream(stream.cuda_stream)
    n_dofs = op.m()
    n = op.n_dofs_per_cell
    Q = int(desc.poly_subcell_ptr[desc.n_polytopes]) * NQ**DIM
    nnz = op.nnz
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")  # > 126 MB L2

    def step_device():
        # a step starts from the flattened agglomeration: the quadrature is rebuilt every step
        op.invalidate_quadrature()
        op.assemble()

    # device-resident timing -----------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = op.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kms = {"volume": [], "faces": [], "reduce": [], "quadrature": []}
    torch.cuda.synchronize()
    for s in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (outside the event pair)
        ev[s][0].record(stream)
        step_device()
        ev[s][1].record(stream)
        ev[s][1].synchronize()
        for k, v in op.last_kernel_ms().items():
            kms[k].append(v)
    torch.cuda.synchronize()
    launches = op.launch_count - launches0
    if dist:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    # end-to-end: host buffers in (pinned), matrix values out (pinned), every step.  A caller that
    # assembles a sequence of matrices double-buffers two handles on two streams, so that the
    # download of step i overlaps the upload and the kernels of step i+1 (PCIe is full duplex);
    # every step still uploads its whole descriptor and delivers its whole matrix inside the timed
    # region.  The strictly serial variant (one handle, synchronous download) is timed beside it.
    out_host = torch.empty(nnz, dtype=torch.float64).pin_memory()
    out_host2 = torch.empty(nnz, dtype=torch.float64).pin_memory()
    stream2 = torch.cuda.Stream()
    op2 = pdl.SIPOperator(desc, keepalive=(ah, keep, part))
    op2.set_stream(stream2.cuda_stream)
    pair = ((op, out_host), (op2, out_host2))

    def e2e_serial(k):
        for _ in range(k):
            op.upload()
            op.assemble()
            op.values_to_host_ptr(out_host.data_ptr())

    def e2e_pipelined(k):
        for i in range(k):
            o, buf = pair[i & 1]
            o.upload()
            o.assemble()
            o.values_to_host_ptr(buf.data_ptr(), wait=False)
        op.synchronize()
        op2.synchronize()

    def timed(fn):
        fn(min(3, args.warmup))
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        fn(args.steps)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    t_e2e_serial_ms = timed(e2e_serial)
    t_e2e_ms = timed(e2e_pipelined)
    assert torch.equal(out_host, out_host2)  # both handles delivered the same matrix
    checksum = float(out_host.sum())

    # vmult with the assembled matrix (device vectors), L2 flushed between applies
    n_src = op.n_source_dofs
    x = torch.from_numpy(np.sin(0.37 * np.arange(n_src)) + 0.01 * (np.arange(n_src) % 7)).cuda()
    y = torch.empty(n_dofs, dtype=torch.float64, device="cuda")

    # one exchange step per apply (ghost-polytope coefficients): over NVLink peer memory
    # (csrc/pd_peer.cu: publish + pull kernels, flag handshake) and, for comparison, NCCL
    peer = pdd.PeerExchange(part, op) if part is not None else None

    def apply(use_peer=True):
        if part is not None and use_peer:
            peer.vmult(y, x)  # pd_peer_vmult: block rows without ghost columns run while the ghost blocks travel
            return
        if part is not None:
            pdd.exchange_ghost_values(part, x)
        op.vmult_ptr(y.data_ptr(), x.data_ptr())

    def time_apply(use_peer):
        for _ in range(3):
            apply(use_peer)
        vm = []
        for _ in range(max(args.steps, 5)):
            flush.zero_()
            if dist:
                dist.barrier()  # ranks enter the timed apply together, as in a solver iteration
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            apply(use_peer)
            b.record(stream)
            b.synchronize()
            vm.append(a.elapsed_time(b))
        return statistics.mean(vm)

    t_vm_ms = time_apply(True)
    t_vm_nccl_ms = time_apply(False) if part is not None else t_vm_ms
    if peer is not None:
        assert peer.ok(), "peer exchange timed out"
    # the loop around vmult: Jacobi-preconditioned CG, device resident (CUDA graph); sharded: ghost exchange
    # and dot-product all-reduce over peer memory inside the graph.  96 iterations, no convergence test.
    bcg = torch.from_numpy(np.cos(0.23 * np.arange(n_dofs)) + 0.1).cuda()
    xcg = torch.zeros_like(bcg)
    solve = (lambda: peer.cg_solve(xcg, bcg, max_iter=96, rel_tol=0.0)) if peer is not None else \
            (lambda: op.cg_solve(xcg, bcg, max_iter=96, rel_tol=0.0))
    solve()
    xcg.zero_()
    if dist:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    cg_iters, _ = solve()
    b.record(stream)
    b.synchronize()
    t_cg_ms = a.elapsed_time(b) / max(cg_iters, 1)

    times = torch.tensor([t_ms, t_e2e_ms, t_vm_ms, t_vm_nccl_ms, t_cg_ms, t_e2e_serial_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_ms, t_e2e_ms, t_vm_ms, t_vm_nccl_ms, t_cg_ms, t_e2e_serial_ms = (float(v) for v in times.cpu())
    if peer is not None:
        dist.barrier()  # nobody unmaps while a neighbour may still pull
        peer.close()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks = read_peaks()
    ms_per_step = t_ms / args.steps
    value = world * n_dofs / (ms_per_step * 1e-3)
    e2e_value = world * n_dofs / (t_e2e_ms / args.steps * 1e-3)
    # dominant kernel: the volume contraction.  Algorithmic flops per launch =
    # 2 n^2 dim Q; bytes = 8 (dim+1) Q read + 8 n^2 per polytope written.
    vol_ms = statistics.mean(kms["volume"])
    flops = 2.0 * n * n * DIM * Q
    achieved = flops / (vol_ms * 1e-3) / 1e12
    vol_bytes = 8.0 * (DIM + 1) * Q + 8.0 * n * n * (n_dofs // n)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_volume_bytes_per_launch")
        except Exception:
            traffic = None
    nblocks = int(desc.brow_ptr[desc.n_block_rows])
    vm_bytes = 8.0 * n * n * nblocks + 4.0 * nblocks + 16.0 * n_dofs
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_dofs_per_gpu": n_dofs, "n_polytopes_per_gpu": int(desc.n_polytopes),
                   "volume_q_points_per_gpu": Q,
                   "sharding": "single GPU" if world == 1 else
                   f"[0,1]^2 x [0,{world}] in {world} z-slabs of 512 polyhedra, cut interfaces evaluated by both sides "
                   "from ghost bbox + DoF block (no assembly collective); vmult pulls ghost blocks over NVLink peer memory",
                   "ghost_polytopes_per_gpu": int(desc.n_polytopes - n_dofs // n),
                   "l2": "flushed (512 MiB memset) between timed steps; inputs 226 MB > L2 as well",
                   "step": "quadrature + volume + faces + diagonal gather, all device kernels"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": nnz * 8,
                "api": "pd_upload + pd_assemble + pd_matrix_values_to_host_async per step, two handles double-buffered on two streams (pinned host buffers); wall clock over all steps",
                "checksum": checksum, "serial_one_handle_value": world * n_dofs * args.steps / (t_e2e_serial_ms * 1e-3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "k_volume<3,2> (FP64 DMMA contraction)", "achieved": achieved,
                     "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["fp64_tflops"],
                     "traffic": traffic, "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": vol_bytes,
                     "kernel_ms": vol_ms, "peak_source": peaks["fp64_src"],
                     "hbm_frac_of_same_kernel": vol_bytes / (vol_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "kernel_ms": {k: statistics.mean(v) for k, v in kms.items()},
        "vmult": {"metric": "SIP vmult GDoF/s (block-CSR apply of the assembled operator"
                            + (", incl. the ghost exchange over NVLink peer memory)" if world > 1 else ")"),
                  "value": world * n_dofs / (t_vm_ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms": t_vm_ms,
                  "roofline": {"bound": "hbm", "achieved": vm_bytes / (t_vm_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": vm_bytes / (t_vm_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                               "peak_source": peaks["hbm_src"]}},
    }
    out["vmult"]["cg_ms_per_iteration"] = t_cg_ms
    out["vmult"]["cg"] = ("Jacobi-PCG around the block-CSR vmult, CUDA-graph replayed"
                          + ("; ghost exchange + dot-product all-reduce over NVLink peer memory inside the graph" if world > 1 else ""))
    if world > 1:
        out["vmult"]["ms_with_nccl_exchange"] = t_vm_nccl_ms
        out["vmult"]["exchange"] = "publish + pull kernels over CUDA-IPC peer memory, epoch-flag handshake (pd_peer_*); NCCL all_to_all_single timed beside it"
    if world == 1 and not args.no_mf_vmult:
        out["mf_vmult"] = mf_vmult_fine_mesh(stream, flush, peaks, max(args.steps, 5))
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(full=True)
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


def mf_vmult_fine_mesh(stream, flush, peaks, steps):
    """The second metric of BASELINE.json: the matrix-free sum-factorised SIP vmult of
    examples/matrix_free_agglo.cc -- Utils::MatrixFreeOperators::LaplaceOperatorDG (include/utils.h:819-925) on
    the fine hex mesh, hyper_cube refined 6x = 64^3 cells, FE_DGQ(2), 7.08 M DoFs -- through
    pd_vmult(PD_VMULT_MATRIX_FREE), device vectors, L2 flushed between applies, CUDA events on the launch
    stream.  Separate from the timed assembly steps above.  (Sharded: tools/run_fine_mf_scaling.py.)"""
    import torch

    import polydeal_b200 as pdl

    n, p = 64, 2
    grid = pdl.Grid.hyper_cube(3, 0.0, 1.0, 6)
    ah = pdl.AgglomerationHandler(grid)
    for c in range(n**3):
        ah.define_agglomerate([c])
    ah.initialize_fe_values(p + 1)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, p)
    op = pdl.SIPOperator(ah.flatten(penalty_constant=p * (p + 1.0), h_rule=pdl.H_NORMAL_EXTENT), keepalive=ah)
    op.set_stream(stream.cuda_stream)
    assert op.matrix_free_available
    op.set_operator(pdl.ASSEMBLE_ALL, 1.0, 0.0)
    N = op.m()
    x = torch.from_numpy(np.sin(0.37 * np.arange(N)) + 0.01 * (np.arange(N) % 7)).cuda()
    y = torch.empty_like(x)
    l0 = op.launch_count
    for _ in range(3):
        op.vmult(y, x, mode=pdl.VMULT_MATRIX_FREE)
    per_apply = (op.launch_count - l0 - 1) // 3  # the first apply also folds the stencil records
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.vmult(y, x, mode=pdl.VMULT_MATRIX_FREE)
        b.record(stream)
        b.synchronize()
        ms.append(a.elapsed_time(b))
    t = statistics.mean(ms)
    nbytes = 16.0 * N + 3 * 64.0 * n**3  # read src + write dst (16 B/DoF) + one 64-byte stencil record per (cell, direction)
    return {"metric": "matrix-free SIP vmult GDoF/s (LaplaceOperatorDG on the fine mesh, examples/matrix_free_agglo.cc: 64^3 hexes, FE_DGQ(2))",
            "value": N / (t * 1e-3) / 1e9, "unit": "GDoF/s", "ms": t, "n_dofs": N, "applies": steps, "gpu_launches_per_apply": int(per_apply),
            "kernel": "k_fine_tile<3,2> (one cell per thread pair, coefficients + halo staged in shared memory by TMA bulk copies)",
            "checksum": float(y.sum()),
            "roofline": {"bound": "hbm", "achieved": nbytes / (t * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": nbytes / (t * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["hbm_src"],
                         "algorithmic_bytes_per_launch": nbytes}}


def oracle_config_b():
    from oracle import pyoracle as po
    from pd_helpers import oracle_handler

    groups = block_groups(N_CELLS_1D, BLOCK)
    _, oah = oracle_handler(DIM, N_CELLS_1D, groups, DEGREE, NQ)
    return po, oah


def cpu_baseline(full, stride=None, oah=None, po=None):
    """The CPU oracle (port of include/poly_utils.h:2000-2195, faithful cost: tables
    re-evaluated per polytope and per face, scalar q*i*j loops) on all host cores."""
    if oah is None:
        po, oah = oracle_config_b()
    cores = os.cpu_count() or 1
    stride = stride or (1 if full else 4)
    m = po.assemble_dg_matrix(oah, degree=DEGREE, n_threads=cores, poly_stride=stride, poly_offset=0)
    n_poly = len(range(0, oah.n_polytopes, stride))
    dofs = n_poly * oah.n_dofs_per_cell
    return {"value": dofs / m.seconds, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_poly} of {oah.n_polytopes} polytopes of config B (every {stride}th), {m.seconds:.2f} s, "
                      "oracle CPU restatement of assemble_dg_matrix, not polyDEAL/deal.II itself", "seconds": m.seconds}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    po, oah = oracle_config_b()
    cores = os.cpu_count() or 1
    stride = 4
    for _ in range(args.warmup):
        cpu_baseline(False, stride, oah, po)
    t, dofs, last = 0.0, 0, None
    for _ in range(args.steps):
        last = cpu_baseline(False, stride, oah, po)
        t += last["seconds"]
        dofs += len(range(0, oah.n_polytopes, stride)) * oah.n_dofs_per_cell
    value = dofs / t
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU arm: oracle port of the reference assembly on the host cores; "
                   "polyDEAL itself needs deal.II/Trilinos/MPI and cannot be built in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mf-vmult", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
