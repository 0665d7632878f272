#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path on B200.

Workload (BASELINE.json configs[2], SURVEY 8d config C -- the largest configuration of `configs` whose matrix fits
one GPU): unit cube, 128^3 hexes agglomerated into 32 768 polyhedra of 4^3 cells, FE_DGQ<3>(3), QGauss(4), SIP
assembly of stiffness + penalty + boundary terms with the library penalty C = 10 (p+dim)(p+1) = 240
(include/poly_utils.h:2018-2019): 2 097 152 DoFs, 134 M volume and 26 M face quadrature points, 7.3 GB of matrix.
N > 1 (weak scaling): N such cubes stacked along z, the polytope adjacency graph partitioned into N parts by METIS
(vertex weight = sub-cells, edge weight = shared sub-faces); every rank assembles its own rows, cut
interfaces from one-time ghost geometry (no data-path collective in assembly); vmult exchanges ghost-polytope
coefficients over NVLink peer memory.

A step = agglomerated quadrature + volume + face + diagonal-gather kernels, from the flattened agglomeration
resident in HBM to the finished scalar-CSR values in HBM.  Metric: polytope DoFs assembled per second (whole job).
The companion metric of BASELINE.json, SIP vmult GDoF/s, rides in the same line at every N:
  "vmult"         block-CSR apply of the matrix just assembled (what the reference's solvers call on agglomerated
                  levels) incl. the ghost exchange, and the CG iteration around it,
  "poly_mf_vmult" the matrix-free apply on the same polytopes,
  "mf_vmult"      the matrix-free sum-factorised LaplaceOperatorDG on the fine mesh of examples/matrix_free_agglo.cc
                  (64^3 hexes, FE_DGQ(2), per GPU; N > 1: METIS partition of the cells, ghost exchange overlapped),
each with its HBM roofline; "rooflines" carries one block per kernel of the step.  N = 1 adds configs B and D.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...   times the CPU oracle restating the reference's assembly on the host cores
                                         (the reference itself needs deal.II and cannot be built here)
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

from pd_workloads import CONFIGS, build_handler

METRIC = "polytope DoFs assembled/s"
UNIT = "DoF/s"
CFG = CONFIGS["C"]
DIM, DEGREE, NQ = CFG["dim"], CFG["p"], CFG["nq"]
N_POLY_PER_GPU = (CFG["n"] // CFG["b"]) ** 3
N_DOFS_PER_GPU = N_POLY_PER_GPU * (DEGREE + 1) ** DIM
WORKLOAD = ("C: unit cube 128^3 hexes -> 32768 polyhedra (4^3 blocks) per GPU, FE_DGQ(3), QGauss(4), SIP "
            "stiffness+penalty+boundary, C=240 (BASELINE configs[2]; N>1: N cubes stacked along z, polytope graph "
            "METIS-partitioned into N parts)")
REF_STRIDE = 128  # --impl reference: every 128th polytope per step (256 polytopes, ~2 s on 16 cores)
CPU_STRIDE = 32   # cpu_baseline of the GPU arm: every 32nd polytope (1024 polytopes, ~10 s)


def config_dict():
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "n_dofs_per_gpu": N_DOFS_PER_GPU, "n_polytopes_per_gpu": N_POLY_PER_GPU,
            "volume_q_points_per_gpu": N_POLY_PER_GPU * CFG["b"] ** 3 * NQ**DIM,
            "l2": "flushed (512 MiB memset) between timed steps; the 7.3 GB matrix written per step does not fit L2 either",
            "step": "quadrature + volume + faces + diagonal gather, all device kernels"}


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


def pin_to_gpu_numa_node(local):
    """Run this rank (and first-touch its pinned buffers) on the CPUs of the GPU's NUMA node: eight ranks share the
    host links otherwise.  Returns a short description for the JSON line."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(local)
        bus = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return f"gpu {bus}: no NUMA affinity reported"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return f"gpu {bus} -> NUMA node {node}, {len(allowed)} cpus"
    except Exception as e:  # containers without sysfs topology
        return f"not pinned ({type(e).__name__})"


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


def src_values(n, offset=0):
    i = np.arange(offset, offset + n, dtype=np.float64)
    return np.sin(0.37 * i) + 0.01 * (np.arange(offset, offset + n) % 7)


class Timer:
    """CUDA-event timing on the launch stream; ranks enter together (barrier) when sharded."""

    def __init__(self, stream, dist):
        self.stream, self.dist = stream, dist

    def per_call_ms(self, fn, reps, warm=3):
        import torch

        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for _ in range(reps):
            fn()
        b.record(self.stream)
        b.synchronize()
        return a.elapsed_time(b) / reps


def hbm_block(name, nbytes, ms, peaks, **extra):
    ach = nbytes / (ms * 1e-3) / 1e9
    return dict(kernel=name, bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"],
                algorithmic_bytes_per_launch=nbytes, kernel_ms=ms, peak_source=peaks["hbm_src"], **extra)


def tensor_block(name, flops, ms, peaks, **extra):
    ach = flops / (ms * 1e-3) / 1e12
    return dict(kernel=name, bound="tensor", achieved=ach, peak=peaks["fp64_tflops"], unit="TFLOP/s",
                frac=ach / peaks["fp64_tflops"], algorithmic_flops_per_launch=flops, kernel_ms=ms,
                peak_source=peaks["fp64_src"], **extra)


def traffic_of(kernel_key):
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            return json.load(open(tp)).get(kernel_key)
        except Exception:
            return None
    return None


def face_point_counts(desc, nq):
    dim = desc.dim
    B = np.ctypeslib.as_array(desc.iface_polyB, (desc.n_ifaces,))
    sp = np.ctypeslib.as_array(desc.iface_sub_ptr, (desc.n_ifaces + 1,))
    nsub = np.diff(sp)
    return (int(nsub[B >= 0].sum()) * nq ** (dim - 1), int(nsub[B < 0].sum()) * nq ** (dim - 1), int((B >= 0).sum()),
            int((B < 0).sum()), int(nsub[B >= 0].sum()), int(nsub[B < 0].sum()))


def assembly_rooflines(desc, n, nq, kms, peaks, mass=False, path="dmma", stats=None):
    """One roofline block per kernel of the assembly step.  `survey_*` = SURVEY 8d's algorithmic count of the
    point-wise formulation (2 n^2 dim Q volume flops, 24 n^2 per interior face point); the tensor path computes the
    same matrix with far fewer operations, so its kernels are held to the bytes they must move (the matrix is
    written once) and to the FP64 pipe for the operations they actually issue."""
    dim = desc.dim
    n_own = desc.n_owned_polytopes or desc.n_polytopes
    n_sub = int(desc.poly_subcell_ptr[desc.n_polytopes])
    Q = n_sub * nq**dim
    qf_int, qf_bnd, n_int, n_bnd, sf_int, sf_bnd = face_point_counts(desc, nq)
    ncomp = dim + (1 if mass else 0)
    vol_flops = 2.0 * n * n * ncomp * Q
    face_flops = 24.0 * n * n * qf_int + 6.0 * n * n * qf_bnd
    out = []
    if path == "tensor":
        n1 = round(n ** (1.0 / dim))
        # per brick of a diagonal block two (cell) or one (face) Kronecker terms of n^2 multiply-adds; per face brick of
        # an interface one; the bricks' 1-D matrices are read (about dim * 2 * N1^2 doubles per item)
        st = stats or {"cell_bricks": n_sub, "face_bricks": sf_int + sf_bnd, "diag_items": n_sub + 2 * sf_int + sf_bnd}
        diag_fma = n * n * (st["diag_items"] + st["cell_bricks"])
        off_fma = n * n * 1.0 * st["face_bricks"] * (sf_int / max(sf_int + sf_bnd, 1))
        item_bytes = 8.0 * dim * 2 * n1 * n1
        diag_bytes = 8.0 * n * n * n_own + item_bytes * st["diag_items"]
        off_bytes = 8.0 * n * n * 2 * n_int + 0.5 * item_bytes * st["face_bricks"]
        for name, ms, nbytes, fma, survey in (
                (f"k_cart_diag<{dim},{desc.fe_degree}> (diagonal blocks: volume + own-side faces, Kronecker sums)", kms["volume"],
                 diag_bytes, diag_fma, vol_flops + (face_flops - 12.0 * n * n * qf_int if qf_int else face_flops)),
                (f"k_cart_offdiag<{dim},{desc.fe_degree}> (M12 and its transpose per interface)", kms["faces"], off_bytes,
                 off_fma, 12.0 * n * n * qf_int)):
            if ms <= 0:
                continue
            blk = hbm_block(name, nbytes, ms, peaks)
            blk["issued_tflops"] = 2.0 * fma / (ms * 1e-3) / 1e12
            blk["fp64_pipe_frac"] = blk["issued_tflops"] / peaks["fp64_tflops"]
            blk["survey_algorithmic_flops_per_launch"] = survey
            blk["survey_algorithmic_tflops"] = survey / (ms * 1e-3) / 1e12
            blk["bricks"] = st
            blk["note"] = ("tensor path: bytes = the blocks written once + the bricks' 1-D matrices read; issued_tflops = the Kronecker "
                           "multiply-adds actually issued; survey_algorithmic_* = the point-wise count of SURVEY 8d the same "
                           "result would cost (n1 = %d)" % n1)
            out.append(blk)
        return out
    out.append(tensor_block(f"k_volume<{dim},{desc.fe_degree}> (FP64 DMMA contraction, upper tiles only)",
                            vol_flops, kms["volume"], peaks,
                            algorithmic_bytes_per_launch=8.0 * (dim + 1) * Q + 8.0 * n * n * n_own,
                            note="algorithmic flops = 2 n^2 dim Q (SURVEY 8d); the kernel issues only the upper triangle of 8x8 "
                                 "tiles, so frac may exceed 1; frac_issued counts the DMMA flops actually issued"))
    n_tiles = (n + 7) // 8
    out[0]["frac_issued"] = out[0]["frac"] * (n_tiles * (n_tiles + 1) / 2) / (n_tiles * n_tiles)
    out.append(tensor_block(f"k_faces<{dim},{desc.fe_degree}> (T + T^T form, FP64 DMMA)", face_flops, kms["faces"], peaks,
                            algorithmic_bytes_per_launch=8.0 * (2 * dim + 1) * (qf_int + qf_bnd) + 8.0 * n * n * (4 * n_int + n_bnd),
                            note="algorithmic flops = 24 n^2 per interior face point + 6 n^2 per boundary point (SURVEY 8d); "
                                 "the T + T^T form issues a third of that"))
    out[1]["frac_issued"] = out[1]["frac"] / 3.0
    if kms.get("quadrature", 0) > 0:
        qbytes = 8.0 * (dim + 1) * Q + 8.0 * (2 * dim + 1) * (qf_int + qf_bnd) + (4.0 * 2**dim + 8.0 * dim) * desc.n_cells
        out.append(hbm_block("k_volume_quadrature + k_face_quadrature", qbytes, kms["quadrature"], peaks))
    rbytes = 8.0 * n * n * (2 * n_own + 2 * n_int + n_bnd)
    out.append(hbm_block("k_reduce_diag (gather of the diagonal blocks)", rbytes, kms["reduce"], peaks,
                         note="reads >= one volume partial per polytope + M11/M22/boundary parts, writes the diagonal block"))
    return out


def time_assembly_steps(op, stream, flush, steps, mass=0.0):
    """Device-timed assembly steps of the path the environment selects: (ms per step, mean kernel ms)."""
    import torch

    kms, tot = {"volume": [], "faces": [], "reduce": [], "quadrature": []}, []
    for s in range(steps + 2):
        flush.zero_()
        op.invalidate_quadrature()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.assemble(stiffness=1.0, mass=mass)
        e.record(stream)
        e.synchronize()
        if s >= 2:
            tot.append(a.elapsed_time(e))
            for k, v in op.last_kernel_ms().items():
                kms[k].append(v)
    return statistics.mean(tot), {k: statistics.mean(v) for k, v in kms.items()}


def serial_kernel_ms(op, stream, flush, steps, mass=0.0):
    """Per-kernel device times with the tensor path's two kernels run one after the other (PD_CART_SERIAL): what the
    rooflines of the individual kernels are computed from; the timed steps run them concurrently."""
    os.environ["PD_CART_SERIAL"] = "1"
    try:
        ms, kms = time_assembly_steps(op, stream, flush, steps, mass)
    finally:
        del os.environ["PD_CART_SERIAL"]
    return ms, kms


def generic_path_block(op, desc, n, nq, stream, flush, peaks, steps, mass=0.0):
    """The DMMA kernels on the agglomerated quadrature (what a distorted mesh runs), forced on the same handle."""
    os.environ["PD_ASSEMBLE_KERNELS"] = "generic"
    try:
        ms, kms = time_assembly_steps(op, stream, flush, steps, mass)
        assert op.assembly_path == "dmma"
    finally:
        del os.environ["PD_ASSEMBLE_KERNELS"]
    return {"what": "PD_ASSEMBLE_KERNELS=generic: rank-k updates over the agglomerated quadrature points on the FP64 tensor "
                    "cores (pd_assemble.cu), the path of distorted meshes, on the same input",
            "ms_per_step": ms, "dofs_per_s": op.m() / (ms * 1e-3), "kernel_ms": kms,
            "rooflines": assembly_rooflines(desc, n, nq, kms, peaks, mass=bool(mass), path="dmma")}


def run_gpu(args):
    # stdout carries exactly ONE JSON line: everything libraries print on fd 1 meanwhile (NCCL's version banner) goes to
    # stderr; the line itself is written to the saved descriptor at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch

    import polydeal_b200 as pdl
    from polydeal_b200 import distributed as pdd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    affinity0 = os.sched_getaffinity(0)
    numa = pin_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- build the agglomeration (host, untimed) --------------------------------------
    t_host0 = time.time()
    ah = build_handler(pdl, CFG, world)
    part = None
    if world == 1:
        desc0 = ah.flatten()  # library penalty, visit by id
    else:
        owner = pdd.partition_by_metis(ah, world)
        part = pdd.LocalPart(ah, owner, rank)
        desc0 = part.desc
    t_host = time.time() - t_host0
    desc, keep, h2d_bytes = pinned_copy_of_desc(desc0)
    op = pdl.SIPOperator(desc, keepalive=(ah, keep, part))
    # all work and all timing events go to ONE explicit non-default stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    op.set_stream(stream.cuda_stream)
    timer = Timer(stream, dist)
    n_dofs = op.m()
    n = op.n_dofs_per_cell
    nnz = op.nnz
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")  # > 126 MB L2

    def step_device():
        # a step starts from the flattened agglomeration: the quadrature is rebuilt every step
        op.invalidate_quadrature()
        op.assemble()

    # ---- device-resident timing --------------------------------------------------------
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
    kms = {k: statistics.mean(v) for k, v in kms.items()}

    # ---- end to end: host buffers in (pinned), matrix values out (pinned), every step ---------------------------
    # A caller that assembles a sequence of matrices double-buffers two handles on two streams, so that the download
    # of step i overlaps the upload and the kernels of step i+1 (PCIe is full duplex); every step still uploads its
    # whole descriptor and delivers its whole matrix inside the timed region.  The strictly serial variant (one
    # handle, synchronous download) is timed beside it.
    e2e_steps = max(2, min(args.steps, 6))
    # two pinned 7.3 GB result buffers per rank: only when the host has the memory for it (all ranks decide alike)
    try:
        import psutil

        avail = psutil.virtual_memory().available / max(world, 1)
    except Exception:
        avail = float("inf")
    double_buffered = avail > 3.5 * nnz * 8
    if dist:
        flag = torch.tensor([1.0 if double_buffered else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        double_buffered = bool(flag.item() > 0.5)
    out_host = torch.empty(nnz, dtype=torch.float64).pin_memory()
    out_host2 = torch.empty(nnz if double_buffered else 1, dtype=torch.float64).pin_memory()
    stream2 = torch.cuda.Stream()
    op2 = pdl.SIPOperator(desc, keepalive=(ah, keep, part)) if double_buffered else op
    if double_buffered:
        op2.set_stream(stream2.cuda_stream)
    pair = ((op, out_host), (op2, out_host2 if double_buffered else out_host))

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
        fn(2)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        fn(e2e_steps)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / e2e_steps

    t_e2e_serial_ms = timed(e2e_serial)
    t_e2e_ms = timed(e2e_pipelined) if double_buffered else t_e2e_serial_ms
    if double_buffered:
        assert torch.equal(out_host, out_host2)  # both handles delivered the same matrix
    checksum = float(out_host[:: max(1, nnz // 4_000_000)].sum())
    del out_host2, op2, pair
    gc.collect()

    # ---- vmult with the assembled matrix: back-to-back applies (the 7.3 GB matrix does not fit L2) ---------------
    n_src = op.n_source_dofs
    x = torch.from_numpy(src_values(n_src)).cuda()
    y = torch.empty(n_dofs, dtype=torch.float64, device="cuda")
    peer = pdd.PeerExchange(part, op) if part is not None else None
    reps = max(args.steps, 10)

    def apply_csr():
        if peer is not None:
            peer.vmult(y, x)  # pd_peer_vmult: block rows without ghost columns run while the ghost blocks travel
        else:
            op.vmult_ptr(y.data_ptr(), x.data_ptr())

    t_vm_ms = timer.per_call_ms(apply_csr, reps)
    t_vm_nccl_ms = t_vm_ms
    if part is not None:
        def apply_nccl():
            pdd.exchange_ghost_values(part, x)
            op.vmult_ptr(y.data_ptr(), x.data_ptr())
        t_vm_nccl_ms = timer.per_call_ms(apply_nccl, reps)
    # the polytopal matrix-free apply of the same operator (no matrix memory)
    op.set_operator(pdl.ASSEMBLE_ALL, 1.0, 0.0)

    def apply_pmf():
        if peer is not None:
            peer.vmult(y, x, mode=pdl.VMULT_MATRIX_FREE)
        else:
            op.vmult_ptr(y.data_ptr(), x.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)

    t_pmf_ms = timer.per_call_ms(apply_pmf, max(3, reps // 3), warm=2)
    pmf_checksum = float(y.sum())
    pmf_stats = op.tensor_path_stats()
    # the loop around vmult: Jacobi-preconditioned CG, device resident (CUDA graph); sharded: ghost exchange
    # and dot-product all-reduce over peer memory inside the graph.  24 iterations, no convergence test.
    bcg = torch.from_numpy(np.cos(0.23 * np.arange(n_dofs)) + 0.1).cuda()
    xcg = torch.zeros_like(bcg)
    solve = (lambda: peer.cg_solve(xcg, bcg, max_iter=24, rel_tol=0.0)) if peer is not None else \
            (lambda: op.cg_solve(xcg, bcg, max_iter=24, rel_tol=0.0))
    solve()
    xcg.zero_()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    cg_iters, _ = solve()
    b.record(stream)
    b.synchronize()
    t_cg_ms = a.elapsed_time(b) / max(cg_iters, 1)
    if peer is not None:
        assert peer.ok(), "peer exchange timed out"
        dist.barrier()  # nobody unmaps while a neighbour may still pull
        peer.close()
    nblocks = int(desc.brow_ptr[desc.n_block_rows])
    path = op.assembly_path
    serial_ms, kms_serial = serial_kernel_ms(op, stream, flush, 3) if path == "tensor" else (None, kms)
    roofs = assembly_rooflines(desc, n, NQ, kms_serial, read_peaks(), path=path, stats=op.tensor_path_stats()) if rank == 0 else None
    generic = None
    if world == 1 and not args.no_extra_configs:
        generic = generic_path_block(op, desc, n, NQ, stream, flush, read_peaks(), 3)
    n_ghost_poly = int(desc.n_polytopes - n_dofs // n)
    n_poly_own = n_dofs // n
    del op, x, y, bcg, xcg, keep, desc, desc0, part, out_host, ah, peer
    gc.collect()
    torch.cuda.empty_cache()

    # ---- the second metric's own path: matrix-free fine-mesh SIP vmult, sharded like the polytopes ---------------
    mf = None if args.no_mf_vmult else mf_vmult_fine_mesh(pdl, pdd, stream, timer, world, rank, dist, 2, reps)
    mfz = None if (args.no_mf_vmult or world == 1) else mf_vmult_fine_mesh(pdl, pdd, stream, timer, world, rank, dist, 2, reps, "zorder")
    if mfz is not None:  # (max over ranks, like every other time of the line)
        import torch

        t_ = torch.tensor([mfz["ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        mfz["ms"] = float(t_.cpu()[0])
    mf3 = None
    if world == 1 and not args.no_mf_vmult:
        mf3 = mf_vmult_fine_mesh(pdl, pdd, stream, timer, 1, 0, None, 3, reps)

    counts = torch.tensor([float(n_dofs)], dtype=torch.float64, device="cuda")
    times = torch.tensor([t_ms, t_e2e_ms, t_vm_ms, t_vm_nccl_ms, t_cg_ms, t_e2e_serial_ms, t_pmf_ms,
                          mf["ms"] if mf else 0.0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    t_ms, t_e2e_ms, t_vm_ms, t_vm_nccl_ms, t_cg_ms, t_e2e_serial_ms, t_pmf_ms, t_mf_ms = (float(v) for v in times.cpu())
    total_dofs = float(counts.cpu()[0])
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peaks = read_peaks()
    ms_per_step = t_ms / args.steps
    value = total_dofs / (ms_per_step * 1e-3)
    e2e_value = total_dofs / (t_e2e_ms * 1e-3)
    vol = max(roofs, key=lambda r: r["kernel_ms"])  # the dominant kernel of the step
    # DRAM bytes per launch from the committed ncu --set full capture of the same kernel on config C (1 GPU)
    for r in roofs:
        r["traffic"] = traffic_of(r["kernel"].split("<")[0] + "_C_bytes_per_launch") if world == 1 else None
    vm_bytes = 8.0 * n * n * nblocks + 4.0 * nblocks + 16.0 * n_dofs
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(),
        "sharding": {"how": "single GPU" if world == 1 else
                     f"[0,1]^2 x [0,{world}], {N_POLY_PER_GPU * world} polyhedra, METIS k-way on the polytope adjacency graph "
                     "(vertex weight sub-cells, edge weight shared sub-faces); cut interfaces evaluated by both sides from "
                     "ghost bbox + DoF block (no assembly collective); vmult ghost blocks over NVLink peer memory",
                     "rank0_owned_polytopes": n_poly_own, "rank0_ghost_polytopes": n_ghost_poly,
                     "total_dofs": total_dofs, "host_setup_s": t_host, "numa": numa},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": nnz * 8,
                "api": ("pd_upload + pd_assemble + pd_matrix_values_to_host_async per step, two handles double-buffered on two "
                        "streams (pinned host buffers, rank pinned to the GPU's NUMA node); wall clock, max over ranks")
                       if double_buffered else
                       "pd_upload + pd_assemble + pd_matrix_values_to_host per step, one handle (host memory too small for two "
                       "pinned result buffers per rank); wall clock, max over ranks",
                "steps": e2e_steps, "checksum": checksum,
                "serial_one_handle_value": total_dofs / (t_e2e_serial_ms * 1e-3)},
        "gpu_launches": int(launches),
        "assembly_path": path + (" (every sub-cell an axis-aligned box: per-sub-cell sum factorisation, pd_cartesian.cu)"
                                 if path == "tensor" else " (FP64 tensor-core contraction over the agglomerated quadrature)"),
        "clocks": clocks,
        "roofline": {k: vol[k] for k in vol if k != "kernel"} | {"kernel": vol["kernel"]},
        "rooflines": roofs,
        "kernel_ms": kms,
        "kernel_ms_serial": kms_serial,
        "kernel_ms_note": ("tensor path: k_cart_diag and k_cart_offdiag run CONCURRENTLY in a timed step (kernel_ms: volume = the "
                           "diagonal kernel, faces = what remains until the off-diagonal kernel ends); the rooflines use "
                           "kernel_ms_serial, measured with PD_CART_SERIAL=1 (one after the other, %s ms per step)"
                           % (("%.3f" % serial_ms) if serial_ms else "-")) if path == "tensor" else None,
        "vmult": {"metric": "SIP vmult GDoF/s (block-CSR apply of the assembled operator"
                            + (", incl. the ghost exchange over NVLink peer memory)" if world > 1 else ")"),
                  "value": total_dofs / (t_vm_ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms": t_vm_ms,
                  "timing": f"{reps} applies back to back (matrix 7.3 GB per GPU > L2), CUDA events, max over ranks",
                  "roofline": hbm_block("k_spmv_block_row", vm_bytes, t_vm_ms, peaks),
                  "cg_ms_per_iteration": t_cg_ms,
                  "cg": "Jacobi-PCG around the block-CSR vmult, CUDA-graph replayed"
                        + ("; ghost exchange + dot-product all-reduce over NVLink peer memory inside the graph" if world > 1 else "")},
        "poly_mf_vmult": {"metric": "matrix-free SIP vmult on the agglomerated polytopes, GDoF/s (sum-factorised per brick, "
                                    "k_cart_apply; coefficients = the bricks' 1-D matrices)",
                          "value": total_dofs / (t_pmf_ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms": t_pmf_ms,
                          "checksum_rank0": pmf_checksum, "bricks": pmf_stats,
                          "roofline": hbm_block("k_cart_apply", 16.0 * n_dofs + 8.0 * DIM * 2 * (DEGREE + 1) ** 2 * pmf_stats["apply_items"],
                                                t_pmf_ms, peaks,
                                                note="bytes = 16 B/DoF (src read, dst written) + the 1-D matrices of every item "
                                                     "(upper bound dim * 2 * N1^2 doubles each); the kernel is latency-bound, "
                                                     "not bandwidth-bound: about 11 k multiply-adds per polytope")},
    }
    if world > 1:
        out["vmult"]["ms_with_nccl_exchange"] = t_vm_nccl_ms
        out["vmult"]["exchange"] = ("publish + pull kernels over CUDA-IPC peer memory, epoch-flag handshake (pd_peer_*); "
                                    "NCCL all_to_all_single timed beside it")
    if mf:
        mf["ms"] = t_mf_ms
        mf["value"] = mf["total_dofs"] / (t_mf_ms * 1e-3) / 1e9
        mf["roofline"] = hbm_block(mf.pop("kernel"), 16.0 * mf["n_dofs_per_gpu"], t_mf_ms, peaks,
                                   note="algorithmic bytes = 16 B/DoF (read src, write dst; SURVEY 8d); per GPU")
        out["mf_vmult"] = mf
        if mfz is not None:
            tz = mfz["ms"]
            mfz["value"] = mfz["total_dofs"] / (tz * 1e-3) / 1e9
            mfz["roofline"] = hbm_block(mfz.pop("kernel"), 16.0 * mfz["n_dofs_per_gpu"], tz, peaks,
                                        note="algorithmic bytes = 16 B/DoF (read src, write dst; SURVEY 8d); per GPU")
            out["mf_vmult_zorder"] = mfz
    if mf3:
        mf3["roofline"] = hbm_block(mf3.pop("kernel"), 16.0 * mf3["n_dofs_per_gpu"], mf3["ms"], peaks)
        out["mf_vmult_dgq3"] = mf3
    if generic:
        out["generic_path"] = generic
    if world == 1 and not args.no_extra_configs:
        out["config_B"] = run_extra_config(pdl, "B", stream, peaks, steps=max(args.steps, 10))
        out["config_D"] = run_extra_config(pdl, "D", stream, peaks, steps=3)
    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, affinity0)  # the CPU arm gets every core of the box
        out["cpu_baseline"] = cpu_baseline(CPU_STRIDE)
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if dist:
        dist.destroy_process_group()


def mf_vmult_fine_mesh(pdl, pdd, stream, timer, world, rank, dist, p, reps, partition="metis"):
    """The second metric of BASELINE.json: the matrix-free sum-factorised SIP vmult of examples/matrix_free_agglo.cc --
    Utils::MatrixFreeOperators::LaplaceOperatorDG (include/utils.h:819-925) on the fine hex mesh, 64^3 cells of
    FE_DGQ(p) per GPU -- through pd_vmult(PD_VMULT_MATRIX_FREE) / pd_peer_vmult, device vectors, `reps` applies back to
    back (163 MB of vectors per GPU > L2), CUDA events on the launch stream.  N > 1: N cubes stacked along z, the
    cells METIS-partitioned, ghost cells pulled over NVLink while the interior cells are applied."""
    import torch

    cfg = dict(dim=3, n=64, b=1, p=p, nq=p + 1)
    ah = build_handler(pdl, cfg, world)
    pen = dict(penalty_constant=max(p, 1) * (p + 1.0), h_rule=pdl.H_NORMAL_EXTENT)
    part = peer = None
    if world == 1:
        op = pdl.SIPOperator(ah.flatten(**pen), keepalive=ah)
    else:
        # metis: GridTools::partition_triangulation as examples/matrix_free_agglo.cc:161 calls it (ragged cuts through the
        # blocks of the curve); zorder: equal ranges of the Morton curve, what p4est / partition_triangulation_zorder give
        owner = pdd.partition_by_metis(ah, world) if partition == "metis" else pdd.partition_by_blocks(ah, world)
        part = pdd.LocalPart(ah, owner, rank, **pen)
        op = pdl.SIPOperator(part.desc, keepalive=(ah, part))
    op.set_stream(stream.cuda_stream)
    assert op.matrix_free_available
    op.set_operator(pdl.ASSEMBLE_ALL, 1.0, 0.0)
    N, ns = op.m(), op.n_source_dofs
    x = torch.from_numpy(src_values(ns)).cuda()
    y = torch.empty(N, dtype=torch.float64, device="cuda")
    if part is not None:
        peer = pdd.PeerExchange(part, op)
    l0 = op.launch_count

    def apply():
        if peer is not None:
            peer.vmult(y, x, mode=pdl.VMULT_MATRIX_FREE)
        else:
            op.vmult_ptr(y.data_ptr(), x.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)

    apply()
    l1 = op.launch_count
    apply()
    per_apply = op.launch_count - l1
    ms = timer.per_call_ms(apply, reps)
    total = torch.tensor([float(N)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    # Sharded: the same apply once more through the independent path -- ghost section filled by the NCCL all_to_all,
    # then the plain local apply -- and the largest difference over all ranks (outside the timed region; a parity
    # property at the driver's scale: METIS cuts with several owners per rank, split tiles)
    check = None
    kernel_last = op.fine_kernel_last
    if peer is not None:
        try:
            y_peer = y.clone()
            pdd.exchange_ghost_values(part, x)
            op.vmult_ptr(y.data_ptr(), x.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)
            stream.synchronize()
            err = torch.stack([(y - y_peer).abs().max(), y.abs().max(),
                               torch.tensor(0.0 if peer.ok() else 1.0, dtype=torch.float64, device="cuda")])
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            err = [float(v) for v in err.cpu()]
            check = {"max_rel_diff_vs_nccl_exchange_path": err[0] / max(err[1], 1e-300), "peer_status_ok_on_all_ranks": err[2] == 0.0}
            check["ok"] = check["peer_status_ok_on_all_ranks"] and check["max_rel_diff_vs_nccl_exchange_path"] <= 1e-12
            y.copy_(y_peer)
        except Exception as e:  # the check must never cost the line
            check = {"ok": None, "not_run": f"{type(e).__name__}: {e}"}
    res = {"metric": f"matrix-free SIP vmult GDoF/s (LaplaceOperatorDG on the fine mesh, examples/matrix_free_agglo.cc: 64^3 hexes "
                     f"per GPU, FE_DGQ({p}))" + (", incl. the ghost exchange over NVLink peer memory" if world > 1 else ""),
           "value": float(total.cpu()[0]) / (ms * 1e-3) / 1e9, "unit": "GDoF/s", "ms": ms, "n_dofs_per_gpu": N,
           "total_dofs": float(total.cpu()[0]), "applies": reps, "gpu_launches_per_apply": int(per_apply),
           "kernel": os.environ.get("PD_FINE_KERNEL", "default") + f" fine-mesh kernel, FE_DGQ<3>({p})",
           "fine_kernel_last": {0: "none", 1: "k_fine_sip (line per thread)", 2: "k_fine_tile", 3: "k_fine_stream (pipelined tiles)"}[kernel_last],
           "partition": "single GPU" if world == 1 else partition,
           "fused_exchange": bool(peer.fused) if peer is not None else None,
           "fused_tiles": peer.fused_tiles if peer is not None else None,
           "sharded_check": check,
           "checksum_rank0": float(y.sum())}
    del l0
    if peer is not None:
        dist.barrier()
        peer.close()
        if check["ok"] is False:  # (from all-reduced values: the same on every rank) reported, not fatal to the headline
            res["error"] = "sharded apply differs from the NCCL-exchange path or a peer wait timed out: the value above is not valid"
    return res


def run_extra_config(pdl, name, stream, peaks, steps):
    """Assembly + block-CSR vmult of another SURVEY 8d configuration on one GPU (extra keys of the N = 1 line)."""
    import torch

    cfg = CONFIGS[name]
    t0 = time.time()
    ah = build_handler(pdl, cfg, 1)
    desc = ah.flatten(penalty_constant=-1.0 if cfg["C"] is None else cfg["C"])
    t_host = time.time() - t0
    op = pdl.SIPOperator(desc, keepalive=ah)
    op.set_stream(stream.cuda_stream)
    N, n = op.m(), op.n_dofs_per_cell
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    ms, kms = time_assembly_steps(op, stream, flush, steps, cfg["mass"])
    path = op.assembly_path
    kms_roof = serial_kernel_ms(op, stream, flush, min(steps, 3), cfg["mass"])[1] if path == "tensor" else kms
    x = torch.from_numpy(src_values(N)).cuda()
    y = torch.empty_like(x)
    vm = []
    for s in range(steps + 2):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.vmult_ptr(y.data_ptr(), x.data_ptr())
        e.record(stream)
        e.synchronize()
        if s >= 2:
            vm.append(a.elapsed_time(e))
    nblocks = int(desc.brow_ptr[desc.n_block_rows])
    vm_ms = statistics.mean(vm)
    res = {"workload": f"{name}: {cfg['n']}^{cfg['dim']} cells -> {desc.n_polytopes} polytopes ({cfg['b']}^{cfg['dim']} blocks), "
                       f"FE_DGQ({cfg['p']}), QGauss({cfg['nq']})" + (", + reaction c=0.5, C=40" if cfg["mass"] else ""),
           "n_dofs": N, "assembly_path": path, "assemble_ms": ms, "dofs_per_s": N / (ms * 1e-3), "kernel_ms": kms,
           "host_setup_s": t_host,
           "kernel_ms_serial": kms_roof,
           "rooflines": assembly_rooflines(desc, n, cfg["nq"], kms_roof, peaks, mass=bool(cfg["mass"]), path=path,
                                           stats=op.tensor_path_stats()),
           "generic_path": generic_path_block(op, desc, n, cfg["nq"], stream, flush, peaks, min(steps, 3), cfg["mass"]),
           "vmult_ms": vm_ms, "vmult_gdofs": N / (vm_ms * 1e-3) / 1e9,
           "vmult_roofline": hbm_block("k_spmv_block_row", 8.0 * n * n * nblocks + 4.0 * nblocks + 16.0 * N, vm_ms, peaks),
           "l2": "flushed between steps / applies"}
    del op, x, y, flush
    gc.collect()
    torch.cuda.empty_cache()
    return res


_ORACLE = {}


def oracle_config_c():
    """The oracle's handler of config C (built once per process)."""
    if "ah" not in _ORACLE:
        from oracle import pyoracle as po
        from pd_workloads import morton_block_groups

        grid = po.Grid(DIM, CFG["n"], 0.0, 1.0, 0)
        oah = po.AgglomerationHandler(grid)
        for g in morton_block_groups(DIM, CFG["n"], CFG["b"]):
            oah.define_agglomerate(g)
        oah.initialize_fe_values(NQ)
        oah.distribute_agglomerated_dofs(po.FE_DGQ, DEGREE)
        _ORACLE["ah"], _ORACLE["po"] = oah, po
    return _ORACLE["po"], _ORACLE["ah"]


def cpu_baseline(stride, offset=5):
    """The CPU oracle (port of include/poly_utils.h:2000-2195, faithful cost: tables re-evaluated per polytope and
    per face, scalar q*i*j loops) on all host cores, on every `stride`-th polytope of config C; the local matrices
    are computed in full and summed into a sink instead of a 7.3 GB global matrix."""
    po, oah = oracle_config_c()
    cores = len(os.sched_getaffinity(0)) or 1
    m = po.assemble_dg_matrix(oah, degree=DEGREE, n_threads=cores, poly_stride=stride, poly_offset=offset, discard_scatter=True)
    n_poly = len(range(offset, oah.n_polytopes, stride))
    dofs = n_poly * oah.n_dofs_per_cell
    return {"value": dofs / m.seconds, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_poly} of {oah.n_polytopes} polytopes of config C (every {stride}th: volume + boundary + visited "
                      f"interfaces), {m.seconds:.2f} s, oracle CPU restatement of assemble_dg_matrix without the global scatter, "
                      "not polyDEAL/deal.II itself", "seconds": m.seconds, "dofs": dofs, "checksum": float(m.values()[0])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0)) or 1
    for _ in range(args.warmup):
        cpu_baseline(REF_STRIDE)
    t, dofs, last = 0.0, 0, None
    for s in range(args.steps):
        last = cpu_baseline(REF_STRIDE, offset=5 + (s % 64))
        t += last["seconds"]
        dofs += last["dofs"]
    value = dofs / t
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(),
        "note": "CPU arm: oracle port of the reference assembly on the host cores, each step a bounded sample of the workload; "
                "polyDEAL itself needs deal.II/Trilinos/MPI and cannot be built in this image",
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
    ap.add_argument("--no-extra-configs", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
