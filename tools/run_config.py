#!/usr/bin/env python3
"""Run one SURVEY 8d configuration (blocks shape) on one GPU: assembly kernel times,
rooflines, size-independent parity properties and vmult timing.  Prints one JSON line.

  python tools/run_config.py A|B|C|D [--steps K] [--check]
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

CONFIGS = {
    # name: dim, cells per direction, block edge, degree, n_q1d, penalty constant (None = library), mass
    "A": dict(dim=2, n=256, b=16, p=1, nq=2, C=None, mass=0.0),
    "B": dict(dim=3, n=64, b=8, p=2, nq=3, C=None, mass=0.0),
    "C": dict(dim=3, n=128, b=4, p=3, nq=4, C=None, mass=0.0),
    "D": dict(dim=3, n=256, b=4, p=2, nq=3, C=40.0, mass=0.5),
    "D8": dict(dim=3, n=128, b=4, p=2, nq=3, C=40.0, mass=0.5),  # one eighth of D (per-GPU share at 8 GPUs)
    # fine-mesh matrix-free operators (every cell its own element), examples/matrix_free_agglo.cc: 64^3, DGQ2
    "F1": dict(dim=3, n=64, b=1, p=1, nq=2, C=2.0, mass=0.0, fine=True),
    "F2": dict(dim=3, n=64, b=1, p=2, nq=3, C=6.0, mass=0.0, fine=True),
    "F3": dict(dim=3, n=64, b=1, p=3, nq=4, C=12.0, mass=0.0, fine=True),
    "F2L": dict(dim=3, n=128, b=1, p=2, nq=3, C=6.0, mass=0.0, fine=True),
    # 2-D fine meshes (512^2 quads), for the tiled-vs-line kernel policy of pd_finemesh.cu
    "G2": dict(dim=2, n=512, b=1, p=2, nq=3, C=6.0, mass=0.0, fine=True),
    "G3": dict(dim=2, n=512, b=1, p=3, nq=4, C=12.0, mass=0.0, fine=True),
    "G4": dict(dim=2, n=512, b=1, p=4, nq=5, C=20.0, mass=0.0, fine=True),
    # E: distorted 64^3 hex box (stand-in for the missing LV mesh), MonodomainOperatorDG semantics
    # f M + sigma K without boundary terms, f = 1.5e4, sigma = 1e-4 (examples/parameters_monodomain.prm)
    "E1": dict(dim=3, n=64, b=1, p=1, nq=2, C=2.0, fine=True, mapped=True, distort=(0.2, 20251018),
               stiffness=1e-4, mass=1.5e4, boundary=False),
    "E2": dict(dim=3, n=64, b=1, p=2, nq=3, C=6.0, fine=True, mapped=True, distort=(0.2, 20251018),
               stiffness=1e-4, mass=1.5e4, boundary=False),
    # the same operators on the undistorted mesh through the general (mapped) kernel, to compare with F1/F2
    "M2": dict(dim=3, n=64, b=1, p=2, nq=3, C=6.0, mass=0.0, fine=True, mapped=True),
}


def fast_block_groups(dim, n, b):
    """block_partition without Python loops over cells: cells in Morton (active) order."""
    ax = np.arange(n, dtype=np.int64)
    if dim == 2:
        i, j = np.meshgrid(ax, ax, indexing="ij")
        k = np.zeros_like(i)
    else:
        i, j, k = np.meshgrid(ax, ax, ax, indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cell = np.zeros_like(i)
    for l in range(n.bit_length() - 1):
        cell |= ((i >> l) & 1) << (dim * l)
        cell |= ((j >> l) & 1) << (dim * l + 1)
        if dim == 3:
            cell |= ((k >> l) & 1) << (dim * l + 2)
    nb = n // b
    part = ((k // b) * nb + (j // b)) * nb + (i // b)
    order = np.lexsort((cell, part))
    cells_sorted = cell[order].astype(np.int32)
    per = b**dim
    return cells_sorted.reshape(nb**dim, per)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    import torch

    import polydeal_b200 as pdl

    dim, n, b, p, nq = cfg["dim"], cfg["n"], cfg["b"], cfg["p"], cfg["nq"]
    t0 = time.time()
    grid = pdl.Grid.hyper_cube(dim, 0.0, 1.0, n.bit_length() - 1)
    if cfg.get("distort"):
        grid.distort_random(*cfg["distort"])
    ah = pdl.AgglomerationHandler(grid)
    for g in fast_block_groups(dim, n, b):
        ah.define_agglomerate(g)
    ah.initialize_fe_values(nq)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, p)
    fine = cfg.get("fine", False)
    desc = ah.flatten(penalty_constant=-1.0 if cfg["C"] is None else cfg["C"],
                      h_rule=pdl.H_NORMAL_EXTENT if fine else pdl.H_DIAMETER_OF_VISITOR)
    t_host = time.time() - t0
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    op = pdl.SIPOperator(desc, keepalive=ah)
    op.set_stream(stream.cuda_stream)
    N, nd = op.m(), op.n_dofs_per_cell
    Q = int(desc.poly_subcell_ptr[desc.n_polytopes]) * nq**dim
    Qf = int(desc.iface_sub_ptr[desc.n_ifaces]) * nq ** (dim - 1)
    nblocks = int(desc.brow_ptr[desc.n_block_rows])
    if fine:
        x = torch.from_numpy(np.sin(0.37 * np.arange(N)) + 0.01 * (np.arange(N) % 7)).cuda()
        y = torch.empty_like(x)
        flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
        vm = []
        mapped = cfg.get("mapped", False)
        mode = pdl.VMULT_MAPPED_FINE if mapped else pdl.VMULT_MATRIX_FREE
        flags = pdl.ASSEMBLE_ALL if cfg.get("boundary", True) else (pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR)
        op.set_operator(flags, cfg.get("stiffness", 1.0), cfg.get("mass", 0.0))
        op.vmult(y, x, mode=mode)  # geometry set-up of the mapped operator happens on first use
        for s in range(args.steps + 3):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            op.vmult(y, x, mode=mode)
            e.record(stream)
            e.synchronize()
            if s >= 3:
                vm.append(a.elapsed_time(e))
        ms = statistics.mean(vm)
        geo_bytes = 0.0
        if mapped:  # streamed geometry: (d(d+1)/2 + 1) doubles per cell point, (2d + 1) per face point, + penalties
            geo_bytes = 8.0 * n**dim * ((dim * (dim + 1) // 2 + 1) * nq**dim + 2 * dim * (2 * dim + 1) * nq ** (dim - 1) + 2 * dim)
        print(json.dumps({"config": args.config, "cells": n**dim, "degree": p, "n_dofs": N, "host_setup_s": t_host,
                          "kernel": "mapped (general hexes)" if mapped else "Cartesian stencil",
                          "fine_kernel_last": op.fine_kernel_last,
                          "GBs_with_geometry": (16.0 * N + geo_bytes) / (ms * 1e-3) / 1e9,
                          "frac_hbm_with_geometry": (16.0 * N + geo_bytes) / (ms * 1e-3) / 1e9 / 6543.4,
                          "mf_vmult_ms": ms, "mf_vmult_gdofs": N / (ms * 1e-3) / 1e9,
                          "mf_GBs_algorithmic_16B_per_dof": 16.0 * N / (ms * 1e-3) / 1e9,
                          "frac_hbm": 16.0 * N / (ms * 1e-3) / 1e9 / 6543.4, "l2": "flushed between applies"}))
        return
    kms = {"volume": [], "faces": [], "reduce": [], "quadrature": []}
    tot = []
    for s in range(args.steps + 2):
        op.invalidate_quadrature()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.assemble(stiffness=1.0, mass=cfg["mass"])
        e.record(stream)
        e.synchronize()
        if s >= 2:
            tot.append(a.elapsed_time(e))
            for k, v in op.last_kernel_ms().items():
                kms[k].append(v)
    x = torch.from_numpy(np.sin(0.37 * np.arange(N)) + 0.01 * (np.arange(N) % 7)).cuda()
    y = torch.empty_like(x)
    vm = []
    for s in range(args.steps + 2):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.vmult(y, x)
        e.record(stream)
        e.synchronize()
        if s >= 2:
            vm.append(a.elapsed_time(e))
    ncomp = dim + (1 if cfg["mass"] else 0)
    vol_flops = 2.0 * nd * nd * ncomp * Q
    n_int = int(sum(1 for f in range(desc.n_ifaces) if desc.iface_polyB[f] >= 0)) if desc.n_ifaces < 4_000_000 else None
    vol_ms, vm_ms = statistics.mean(kms["volume"]), statistics.mean(vm)
    vm_bytes = 8.0 * nd * nd * nblocks + 4.0 * nblocks + 16.0 * N
    out = {
        "config": args.config, "dim": dim, "cells": n**dim, "polytopes": int(desc.n_polytopes), "degree": p, "n_dofs": N,
        "volume_points": Q, "face_points": Qf, "n_blocks": nblocks, "matrix_GB": nblocks * nd * nd * 8 / 1e9,
        "host_setup_s": t_host, "assemble_ms": statistics.mean(tot), "kernel_ms": {k: statistics.mean(v) for k, v in kms.items()},
        "dofs_per_s": N / (statistics.mean(tot) * 1e-3),
        "volume_tflops_algorithmic": vol_flops / (vol_ms * 1e-3) / 1e12, "volume_frac_fp64_peak": vol_flops / (vol_ms * 1e-3) / 1e12 / 37.1,
        "vmult_ms": vm_ms, "vmult_gdofs": N / (vm_ms * 1e-3) / 1e9, "vmult_GBs": vm_bytes / (vm_ms * 1e-3) / 1e9,
        "vmult_frac_hbm": vm_bytes / (vm_ms * 1e-3) / 1e9 / 6543.4, "interior_ifaces": n_int,
        "mem_GB": torch.cuda.max_memory_allocated() / 1e9,
    }
    # matrix-free polytopal apply of the same operator (no matrix memory)
    op.set_operator(pdl.ASSEMBLE_ALL, 1.0, cfg["mass"])
    mf = []
    for s in range(args.steps + 2):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.vmult(y, x, mode=pdl.VMULT_MATRIX_FREE)
        e.record(stream)
        e.synchronize()
        if s >= 2:
            mf.append(a.elapsed_time(e))
    out["mf_poly_vmult_ms"] = statistics.mean(mf)
    out["mf_poly_vmult_gdofs"] = N / (statistics.mean(mf) * 1e-3) / 1e9
    # CG iteration rate through the CUDA-graph loop (block-CSR)
    b_ = torch.ones_like(x)
    x_ = torch.zeros_like(x)
    op.cg_solve(x_, b_, max_iter=8, rel_tol=0.0)
    x_.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    it, rr = op.cg_solve(x_, b_, max_iter=96, rel_tol=0.0)
    e.record(stream)
    e.synchronize()
    out["cg_ms_per_iteration"] = a.elapsed_time(e) / max(it, 1)
    out["cg_iterations_timed"] = it
    if args.check:
        # size-independent properties: constants in the kernel of the boundary-free stiffness operator
        # (checked through vmult), symmetry through x'Ay = y'Ax, energy of u = x_0 equals |Omega| = 1
        op.assemble(flags=pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR, stiffness=1.0, mass=0.0)
        one = torch.ones_like(x)
        op.vmult(y, one)
        stream.synchronize()
        op.synchronize()
        scale = float(torch.max(torch.abs(torch.from_numpy(op.values()[: 10 * nd * nd]))))
        out["max_abs_A_one_over_scale"] = float(y.abs().max()) / scale
        z = torch.from_numpy(np.cos(0.11 * np.arange(N))).cuda()
        ax_, az_ = torch.empty_like(x), torch.empty_like(x)
        op.vmult(ax_, x)
        op.vmult(az_, z)
        stream.synchronize()
        out["sym_rel"] = abs(float(z @ ax_) - float(x @ az_)) / abs(float(z @ ax_))
        # u = x_0 interpolated at the DGQ support points of every bbox
        nodes = np.empty(p + 1)
        from polydeal_b200 import _capi as K

        K.check(K.lib().pd_dgq_nodes_1d(p, nodes.ctypes.data))
        bbox = np.ctypeslib.as_array(desc.bbox, (desc.n_polytopes, 2 * dim))
        blk = np.ctypeslib.as_array(desc.dof_block, (desc.n_polytopes,))
        i = np.arange(nd)
        u = np.empty(N)
        ux = bbox[:, 0:1] + nodes[i % (p + 1)][None, :] * (bbox[:, dim:dim + 1] - bbox[:, 0:1])
        u.reshape(-1, nd)[blk] = ux
        ud = torch.from_numpy(u).cuda()
        op.vmult(y, ud)
        stream.synchronize()
        out["energy_u_eq_x"] = float(ud @ y)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
