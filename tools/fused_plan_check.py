#!/usr/bin/env python3
"""Host-only check of the fused sharded fine-mesh plan (no GPU): builds the sharded fine mesh of bench.py's `mf_vmult`
leg (64^3 cells of FE_DGQ(2) per rank, `world` cubes stacked along z, METIS or z-order partition), repeats on the host
what setup_fine_operator / setup_fine_fused (csrc/pd_finemesh.cu) do with it -- neighbour table, Morton order,
interior / boundary split by blocks of the curve, tile plans, ghost phases as the owners' export buffers give them
(csrc/pd_peer.cu: export_at) -- through the same templates (csrc/pd_fine_cell.hpp, compiled by g++ from
tests/csrc/fine_cell_host.cpp), and reports per rank how many halo rows the tiles need, before and after the split of
the tiles that exceed the gather's budget (128 rows for N = 27).

  python tools/fused_plan_check.py [--world 2] [--n 64] [--partition metis|zorder]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

FINE_TILE = 64


def host_lib():
    out = os.path.join(tempfile.mkdtemp(prefix="fine_cell_"), "libfine_cell_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "csrc", "fine_cell_host.cpp")], stderr=subprocess.DEVNULL)
    return C.CDLL(out)


def neighbour_table(desc, dim):
    """nbr[block][face] as setup_fine_operator builds it from the flattened interface list (local block numbers,
    ghosts >= n_owned, -1 at the boundary)."""
    nfc = 2 * dim
    npoly, n_own = desc.n_polytopes, desc.n_owned_polytopes or desc.n_polytopes
    A = np.ctypeslib.as_array(desc.iface_polyA, (desc.n_ifaces,)).astype(np.int64)
    B = np.ctypeslib.as_array(desc.iface_polyB, (desc.n_ifaces,)).astype(np.int64)
    sp = np.ctypeslib.as_array(desc.iface_sub_ptr, (desc.n_ifaces + 1,))
    n_sub = int(sp[-1])
    sub_face = np.ctypeslib.as_array(desc.sub_face, (n_sub,)).astype(np.int64)
    blk = np.ctypeslib.as_array(desc.dof_block, (npoly,)).astype(np.int64)
    cnt = np.diff(sp)
    a, b = np.repeat(A, cnt), np.repeat(B, cnt)
    nbr = np.full((n_own, nfc), -1, dtype=np.int32)
    nbr[blk[a], sub_face] = np.where(b >= 0, blk[np.maximum(b, 0)], -1)
    m = (b >= 0) & (b < n_own)
    nbr[blk[b[m]], sub_face[m] ^ 1] = blk[a[m]]
    return nbr


def morton_keys(desc, dim):
    npoly, n_own = desc.n_polytopes, desc.n_owned_polytopes or desc.n_polytopes
    bbox = np.ctypeslib.as_array(desc.bbox, (npoly, 2 * dim))[:n_own]
    blk = np.ctypeslib.as_array(desc.dof_block, (npoly,))[:n_own].astype(np.int64)
    ctr = np.empty((n_own, dim))
    ctr[blk] = 0.5 * (bbox[:, :dim] + bbox[:, dim:])
    lo, hmin = bbox[:, :dim].min(axis=0), (bbox[:, dim:] - bbox[:, :dim]).min(axis=0)
    q0 = np.floor(np.floor(lo / hmin) / 1048576.0) * 1048576
    q = np.clip(np.floor(ctr / hmin) - q0, 0, 2097151).astype(np.uint64)
    key = np.zeros(n_own, dtype=np.uint64)
    for b in range(21):
        for k in range(dim):
            key |= ((q[:, k] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * dim + k)
    return key


def tile_first_of(lib, seq, key, nbr, n_total, n, dim):
    nfc = 2 * dim
    block_bits = 2 if dim == 3 else 3
    bkey = np.ascontiguousarray(key[seq] >> np.uint64(block_bits * dim))
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    n_seq = len(seq)
    tf, tp = np.zeros(n_seq + 1, np.int32), np.zeros(n_seq + 1, np.int32)
    noff = np.zeros(n_seq * nfc, np.uint16)
    cap = n_seq * nfc + 16
    halo = np.zeros(cap, np.int32)
    nt, mh, zoff, hr, nh = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    seqc, nbrc = np.ascontiguousarray(seq, dtype=np.int32), np.ascontiguousarray(nbr)
    rc = lib.fine_tile_plan_host(n_seq, dp(seqc), dp(bkey), dp(nbrc), nfc, n_total, FINE_TILE, n, C.byref(nt), C.byref(mh),
                                 C.byref(zoff), C.byref(hr), dp(tf), dp(tp), dp(halo), C.c_int64(cap), C.byref(nh), dp(noff))
    assert rc == 0, rc
    return tf[: nt.value + 1].copy(), mh.value


def fused_plan(lib, inner, outer, tf1, tf2, nbr, n_total, n, dim, par, max_rows):
    nfc = 2 * dim
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    ns = len(inner) + len(outer)
    seq, tf, base = np.zeros(ns, np.int32), np.zeros(ns + 1, np.int32), np.zeros(ns, np.int32)
    rows_cap = (ns + 1) * max(max_rows, 1)
    rows, noff = np.zeros(min(rows_cap, 1 << 28), np.int32), np.zeros(ns * nfc, np.uint16)
    nt, fg, mr, zoff, um = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    a = [np.ascontiguousarray(v, dtype=np.int32) for v in (inner, outer, tf1, tf2)]
    nbrc, parc = np.ascontiguousarray(nbr), np.ascontiguousarray(par, dtype=np.uint8)
    rc = lib.fine_fused_plan_host(len(inner), dp(a[0]), len(outer), dp(a[1]), len(tf1) - 1, dp(a[2]), len(tf2) - 1, dp(a[3]),
                                  dp(nbrc), nfc, n_total, FINE_TILE, n, dp(parc), max_rows, C.byref(nt), C.byref(fg), C.byref(mr),
                                  C.byref(zoff), C.byref(um), dp(seq), dp(tf), dp(base), dp(rows), C.c_int64(len(rows)), dp(noff))
    return dict(rc=rc, n_tiles=nt.value, first_ghost_tile=fg.value, max_rows=mr.value, unsplit_max_rows=um.value,
                unsplit_tiles=len(tf1) + len(tf2) - 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--partition", default="metis", choices=["metis", "zorder"])
    ap.add_argument("--max-rows", type=int, default=128)
    args = ap.parse_args()
    import polydeal_b200 as pdl
    from polydeal_b200 import distributed as pdd
    from pd_workloads import build_handler

    lib = host_lib()
    dim, p = 3, args.p
    n = (p + 1) ** dim
    ah = build_handler(pdl, dict(dim=dim, n=args.n, b=1, p=p, nq=p + 1), args.world)
    owner = pdd.partition_by_metis(ah, args.world) if args.partition == "metis" else pdd.partition_by_blocks(ah, args.world)
    pen = dict(penalty_constant=max(p, 1) * (p + 1.0), h_rule=pdl.H_NORMAL_EXTENT)
    # (a LocalPart's descriptor points into storage of the host handler that the next flatten reuses: one rank at a time)
    send_ptr = [np.concatenate([[0], np.cumsum(pdd.LocalPart(ah, owner, r, **pen).send_counts)]) for r in range(args.world)]
    for r in range(args.world):
        pt = pdd.LocalPart(ah, owner, r, **pen)
        desc = pt.desc
        n_own, n_total = pt.n_owned, pt.n_owned + pt.n_ghost
        nbr = neighbour_table(desc, dim)
        key = morton_keys(desc, dim)
        order = np.argsort(key, kind="stable").astype(np.int32)
        reads_ghost = (nbr >= n_own).any(axis=1)
        blk_key = key >> np.uint64(2 * dim if dim == 3 else 3 * dim)
        bnd_blocks = np.unique(blk_key[reads_ghost])
        is_outer = np.isin(blk_key[order], bnd_blocks)
        inner, outer = order[~is_outer], order[is_outer]
        tf1, mh1 = tile_first_of(lib, inner, key, nbr, n_total, n, dim)
        tf2, mh2 = tile_first_of(lib, outer, key, nbr, n_total, n, dim)
        # ghost g of owner s sits at position remote_off[s] + (g - recv_ptr[s]) of s's send list (pd_peer.cu: peer_create);
        # the export buffer gives block b the phase b & 1 for odd n (export_at)
        par = np.zeros(n_total, np.uint8)
        par[:n_own] = (np.arange(n_own, dtype=np.int64) * n) & 1
        recv_ptr = np.concatenate([[0], np.cumsum(pt.recv_counts)])
        for s in range(args.world):
            g = np.arange(recv_ptr[s], recv_ptr[s + 1])
            b = send_ptr[s][r] + (g - recv_ptr[s])
            par[n_own + g] = ((b & 1) if n & 1 else 0)
        out = fused_plan(lib, inner, outer, tf1, tf2, nbr, n_total, n, dim, par, args.max_rows)
        # the same with the ghost phases of the local vector (the split path's plan)
        par_local = ((np.arange(n_total, dtype=np.int64) * n) & 1).astype(np.uint8)
        loc = fused_plan(lib, inner, outer, tf1, tf2, nbr, n_total, n, dim, par_local, args.max_rows)
        print(json.dumps(dict(rank=r, n_owned=int(n_own), n_ghost=int(pt.n_ghost), interior_cells=int(len(inner)),
                              boundary_cells=int(len(outer)), tile_kernel_max_halo=[mh1, mh2], fused=out,
                              local_ghost_phases=loc)))


if __name__ == "__main__":
    main()
