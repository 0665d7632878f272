"""Host-side checks of the tiled fine-mesh SIP kernel (polydeal_b200/csrc/pd_fine_cell.hpp): the per-cell
arithmetic cell_apply (premultiplied stencils M^-1 L_d + mass passes, LaplaceOperatorDG / MonodomainOperatorDG
include/utils.h:819-925, 1565-1659) against a dense Kronecker restatement, and the tile plan (own / halo / zero
slots) against its contract.  The templates are compiled here with g++ (tests/csrc/fine_cell_host.cpp); the GPU
parity of the kernel itself is tests/test_gpu_parity.py::test_fine_mesh_matrix_free_vmult."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("fine_cell") / "libfine_cell_host.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                           os.path.join(HERE, "csrc", "fine_cell_host.cpp")])
    return C.CDLL(str(out))


def gll_nodes(p):
    return {1: [0.0, 1.0], 2: [0.0, 0.5, 1.0], 3: [0.0, 0.5 - np.sqrt(5) / 10, 0.5 + np.sqrt(5) / 10, 1.0],
            4: [0.0, 0.5 - np.sqrt(21) / 14, 0.5, 0.5 + np.sqrt(21) / 14, 1.0]}[p]


def tables_1d(p):
    """Mh, Sh, e[2], d[2] of FE_DGQ(p) on [0,1] with QGauss(p+1)."""
    nodes = np.array(gll_nodes(p))
    n1 = p + 1
    polys = []
    for a in range(n1):
        c = np.poly1d([1.0])
        for b in range(n1):
            if b != a:
                c = c * np.poly1d([1.0, -nodes[b]]) / (nodes[a] - nodes[b])
        polys.append(c)
    xg, wg = np.polynomial.legendre.leggauss(n1)
    xg, wg = 0.5 * (xg + 1.0), 0.5 * wg
    L = np.array([[pl(x) for x in xg] for pl in polys])
    dL = np.array([[pl.deriv()(x) for x in xg] for pl in polys])
    Mh = (L * wg) @ L.T
    Sh = (dL * wg) @ dL.T
    e = np.array([[pl(s) for pl in polys] for s in (0.0, 1.0)])
    d = np.array([[pl.deriv()(s) for pl in polys] for s in (0.0, 1.0)])
    return Mh, Sh, e, d


def kron_dir(mats):
    """x fastest: kron(A_z, A_y, A_x)"""
    out = np.array([[1.0]])
    for m in mats:
        out = np.kron(m, out)
    return out


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (2, 4), (3, 1), (3, 2)])
def test_cell_apply_equals_dense_kronecker_form(lib, dim, p):
    rng = np.random.default_rng(7 * dim + p)
    n1, N = p + 1, (p + 1) ** dim
    Mh, Sh, e, d = tables_1d(p)
    tab = np.concatenate([Mh.ravel(), np.linalg.solve(Mh, Sh).ravel(), np.linalg.solve(Mh, e.T).T.ravel(),
                          np.linalg.solve(Mh, d.T).T.ravel(), d.ravel()])
    for trial in range(4):
        u = rng.standard_normal(N)
        nb = rng.standard_normal((2 * dim, N))
        coef = rng.uniform(0.2, 2.0, (dim, 7))  # cVol, cD0, cD1, P0, P1, Q0, Q1
        if trial == 1:  # a boundary face: no neighbour, Q = 0
            nb[0] = 0.0
            coef[0, 5] = 0.0
        if trial == 2:  # face terms switched off in one direction
            coef[dim - 1, 1:] = 0.0
        mv = 0.0 if trial == 0 else 1.7
        # dense form of the line stencils (pd_finemesh.cu: a_s, b_s of k_fine_sip)
        ref = mv * kron_dir([Mh] * dim) @ u
        e0, e1, d0, d1 = e[0], e[1], d[0], d[1]
        for k in range(dim):
            cV, cD0, cD1, P0, P1, Q0, Q1 = coef[k]
            own = (cV * Sh + np.outer(e0, P0 * e0 + cD0 * d0) + np.outer(d0, cD0 * e0)
                   + np.outer(e1, P1 * e1 - cD1 * d1) + np.outer(d1, -cD1 * e1))
            n0 = np.outer(e0, -P0 * e1 + Q0 * d1) + np.outer(d0, -cD0 * e1)
            n1m = np.outer(e1, -P1 * e0 - Q1 * d0) + np.outer(d1, cD1 * e0)
            place = lambda m: kron_dir([m if kk == k else Mh for kk in range(dim)])
            ref = ref + place(own) @ u + place(n0) @ nb[2 * k] + place(n1m) @ nb[2 * k + 1]
        out = np.empty(N)
        dp = lambda a: a.ctypes.data_as(C.c_void_p)
        nbc, coefc = np.ascontiguousarray(nb), np.ascontiguousarray(coef)
        assert lib.fine_cell_host(dim, n1, dp(tab), dp(u), dp(nbc), dp(coefc), C.c_double(mv), dp(out)) == 0
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("dim,p", [(2, 2), (2, 4), (3, 2), (3, 3)])
def test_dense_uniform_form_equals_rank_one_form(lib, dim, p):
    """cell_apply_dense (the pipelined kernel's arithmetic on a uniform mesh: dense N1 x N1 line matrices built by
    build_dense_tables, boundary sides as a correction, mass term on the last direction's diagonal) equals
    cell_apply fed the per-cell records a uniform mesh would have, for every combination of boundary sides."""
    rng = np.random.default_rng(11 * dim + p)
    n1, N = p + 1, (p + 1) ** dim
    Mh, Sh, e, d = tables_1d(p)
    tab = np.concatenate([Mh.ravel(), np.linalg.solve(Mh, Sh).ravel(), np.linalg.solve(Mh, e.T).T.ravel(),
                          np.linalg.solve(Mh, d.T).T.ravel(), d.ravel()])
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    for trial in range(1 << (2 * dim)):
        bnd = np.array([(trial >> f) & 1 for f in range(2 * dim)], dtype=np.int32)
        uni = rng.uniform(0.2, 2.0, (dim, 7))  # cVol, cDi, Pi, Qi, cDb, Pb0, Pb1
        if trial % 5 == 3:
            uni[:, 4:] = 0.0  # boundary term switched off
        u = rng.standard_normal(N)
        nb = rng.standard_normal((2 * dim, N))
        coef = np.empty((dim, 7))  # cVol, cD0, cD1, P0, P1, Q0, Q1 as the records carry them
        for k in range(dim):
            b0, b1 = bnd[2 * k], bnd[2 * k + 1]
            coef[k] = [uni[k, 0], uni[k, 4] if b0 else uni[k, 1], uni[k, 4] if b1 else uni[k, 1],
                       uni[k, 5] if b0 else uni[k, 2], uni[k, 6] if b1 else uni[k, 2],
                       0.0 if b0 else uni[k, 3], 0.0 if b1 else uni[k, 3]]
            if b0:
                nb[2 * k] = 0.0
            if b1:
                nb[2 * k + 1] = 0.0
        mv = 0.0 if trial % 2 == 0 else 2.3
        ref, out = np.empty(N), np.empty(N)
        nbc, coefc, unic = np.ascontiguousarray(nb), np.ascontiguousarray(coef), np.ascontiguousarray(uni)
        assert lib.fine_cell_host(dim, n1, dp(tab), dp(u), dp(nbc), dp(coefc), C.c_double(mv), dp(ref)) == 0
        assert lib.fine_cell_dense_host(dim, n1, dp(tab), dp(u), dp(nbc), dp(unic), dp(bnd), C.c_double(mv), dp(out)) == 0
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()


def grid_neighbours(shape, order):
    """neighbour table [cell][2 dim] of a Cartesian grid, cells numbered lexicographically or in Morton order"""
    dim = len(shape)
    idx = np.arange(int(np.prod(shape))).reshape(shape[::-1])  # [z][y][x]
    coords = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1).reshape(-1, dim)
    if order == "morton":
        key = np.zeros(len(coords), dtype=np.int64)
        for b in range(10):
            for k in range(dim):
                key |= ((coords[:, k] >> b) & 1) << (b * dim + k)
        perm = np.argsort(key, kind="stable")
    else:
        perm = np.lexsort([coords[:, k] for k in range(dim)])
    coords = coords[perm]
    num = {tuple(c): i for i, c in enumerate(coords)}
    nbr = -np.ones((len(coords), 2 * dim), dtype=np.int32)
    for i, c in enumerate(coords):
        for k in range(dim):
            for s, step in enumerate((-1, 1)):
                cc = list(c)
                cc[k] += step
                if 0 <= cc[k] < shape[k]:
                    nbr[i, 2 * k + s] = num[tuple(cc)]
    return nbr


@pytest.mark.parametrize("shape,order,tile,with_list,n,blocks", [
    ((8, 8, 8), "morton", 64, False, 27, True), ((5, 7, 3), "lex", 64, False, 8, False), ((16, 16), "morton", 64, False, 16, True),
    ((6, 6, 6), "lex", 16, True, 27, False), ((3, 2), "lex", 64, False, 25, False), ((8, 8, 8), "morton", 64, True, 27, True),
    ((8, 8, 8), "morton", 64, "interior", 27, True),
])
def test_tile_plan_contract(lib, shape, order, tile, with_list, n, blocks):
    nbr = grid_neighbours(shape, order)
    n_cells, nfc = nbr.shape
    dim = len(shape)
    seq = None
    if with_list == "interior":  # what a sharded apply does first: the cells that touch no boundary, in curve order
        seq = np.ascontiguousarray(np.nonzero((nbr >= 0).all(axis=1))[0].astype(np.int32))
    elif with_list:  # an arbitrary sub-sequence
        seq = np.ascontiguousarray(np.arange(n_cells, dtype=np.int32)[::2])
    n_seq = n_cells if seq is None else len(seq)
    cells = np.arange(n_cells, dtype=np.int32) if seq is None else seq
    key = None
    if blocks:  # Morton-numbered cells: the aligned block of 64 = cell id >> 6
        key = np.ascontiguousarray((cells >> 6).astype(np.uint64))
    tile_first = np.zeros(n_seq + 1, dtype=np.int32)
    tile_ptr = np.zeros(n_seq + 1, dtype=np.int32)
    halo = np.zeros(n_seq * nfc + 1, dtype=np.int32)
    noff = np.zeros(n_seq * nfc, dtype=np.uint16)
    n_tiles, max_halo, zoff, rh, n_halo = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.fine_tile_plan_host(n_seq, dp(seq) if seq is not None else None, dp(key) if key is not None else None, dp(nbr), nfc,
                                 n_cells, tile, n, C.byref(n_tiles), C.byref(max_halo), C.byref(zoff), C.byref(rh), dp(tile_first),
                                 dp(tile_ptr), dp(halo), C.c_int64(len(halo)), C.byref(n_halo), dp(noff))
    assert rc == 0
    nt = n_tiles.value
    assert tile_first[0] == 0 and tile_first[nt] == n_seq and tile_ptr[nt] == n_halo.value
    rh, ro = rh.value, n | 1
    assert rh % 2 == 0 and rh >= n + 1 and rh % 4 == 2  # a 16-byte aligned row with room for an odd start, spread over the banks
    seen_max_halo = 0
    for k in range(nt):
        s0, s1 = tile_first[k], tile_first[k + 1]
        assert 0 < s1 - s0 <= tile
        own = cells[s0:s1]
        if key is None:
            assert s1 - s0 == tile or k == nt - 1
        else:  # a tile never straddles two blocks, and a block is not cut unless it is full
            assert len(set(key[s0:s1].tolist())) == 1
            assert s1 == n_seq or key[s1] != key[s0] or s1 - s0 == tile
        hl = halo[tile_ptr[k]:tile_ptr[k + 1]]
        assert len(set(own.tolist()) | set(hl.tolist())) == len(own) + len(hl)  # every cell staged once
        seen_max_halo = max(seen_max_halo, len(hl))
        where = {int(c): i * ro for i, c in enumerate(own)}  # own rows an odd number of doubles apart
        # a halo row starts at the 16-byte boundary below the cell's first coefficient
        where.update({int(c): tile * ro + r * rh + ((int(c) * n) & 1) for r, c in enumerate(hl)})
        for i, c in enumerate(own):
            for f in range(nfc):
                o = noff[(s0 + i) * nfc + f]
                assert o == (zoff.value if nbr[c, f] < 0 else where[int(nbr[c, f])])
        # no halo cell that nobody needs
        needed = {int(nbr[c, f]) for c in own for f in range(nfc) if nbr[c, f] >= 0} - set(own.tolist())
        assert needed == set(hl.tolist())
    assert max_halo.value == seen_max_halo and zoff.value == tile * ro + seen_max_halo * rh
    if order == "morton" and shape == (8, 8, 8) and with_list is False:
        assert nt == 8 and seen_max_halo == 48  # a 4x4x4 corner block and its three inner faces
    if with_list == "interior":  # 6^3 interior cells: the eight 3x3x3 corners of the blocks, compact halos
        assert nt == 8 and seen_max_halo <= 96


@pytest.mark.parametrize("n", [4, 8, 9, 16, 25, 27])
def test_halo_bulk_copies_are_aligned_and_in_bounds(n):
    """The arithmetic k_fine_tile uses to fill a halo row with ONE 16-byte-granular bulk copy (pd_finemesh.cu,
    stage of the halo: first = cell * n doubles, copy from the 16-byte boundary below it, a trailing 8-byte copy
    when 8 bytes are left), restated here and checked for every cell of a vector, the last one included:
    source and destination 16-byte aligned, size a positive multiple of 16, nothing read past the vector, nothing
    written past the row, and coefficient k of the cell lands where the tile plan's offset (odd start) says.
    (compute-sanitizer is not available on the GPU pool; this is the bounds check of our own.)"""
    tile, n_cells = 64, 37
    ro, rh = n | 1, None
    rh = (n + 2) // 2 * 2
    if rh % 4 == 0:
        rh += 2
    vector_bytes = n_cells * n * 8
    for hid in range(n_cells):
        first = hid * n
        a0 = first & ~1
        total = (n + first - a0) * 8
        sz = total & ~15
        tail = total - sz
        assert (a0 * 8) % 16 == 0 and sz > 0 and sz % 16 == 0 and tail in (0, 8)
        assert a0 * 8 + sz + tail <= vector_bytes  # the last cell does not read past the end of the vector
        for r in (0, 1, 95):
            dst = (tile * ro + r * rh) * 8
            assert dst % 16 == 0 and sz + tail <= rh * 8  # stays inside its row
        odd = (hid * n) & 1  # what build_tile_plan adds to the row's offset
        assert first - a0 == odd
        # row[j] = x[a0 + j]  =>  row[odd + k] = x[first + k]
        assert all(a0 + odd + k == first + k for k in range(n)) and (odd + n) * 8 <= sz + tail
    # the contiguous own range of a full tile: 16-byte aligned start and size whenever own rows are n doubles apart
    if ro == n:
        for t in range(5):
            assert (t * tile * n * 8) % 16 == 0 and (tile * n * 8) % 16 == 0


@pytest.mark.parametrize("shape,tile,n,ragged", [((8, 8, 8), 64, 27, False), ((16, 16), 64, 9, False), ((16, 16), 64, 25, True),
                                                  ((4, 4, 4), 64, 27, False), ((8, 8, 8), 64, 27, True)])
def test_stream_plan_contract(lib, shape, tile, n, ragged):
    """build_stream_plan (the pipelined kernel k_fine_stream): every neighbour offset points at the neighbour's own
    row (shifted by the tile's 16-byte phase), at the halo row that holds it, or at the zero row; halo rows are packed
    n doubles apart, a cell whose first coefficient is 16-byte aligned in the vector sits in an even row and the
    others in an odd row (source and destination of the 16-byte cp.async chunks then have the same phase), and no
    row is used twice.  ragged: tiles of any length and alignment (the blocks a METIS partition cuts)."""
    nbr = grid_neighbours(shape, "morton")
    n_cells, nfc = nbr.shape
    rows_cap = 128
    if ragged:
        rng = np.random.default_rng(5)
        tf = [0]
        while tf[-1] < n_cells:
            tf.append(min(n_cells, tf[-1] + int(rng.integers(1, tile + 1))))
    else:
        tf = list(range(0, n_cells + 1, tile))
    tf = np.array(tf, dtype=np.int32)
    n_tiles = len(tf) - 1
    rows = np.empty((n_tiles, rows_cap), dtype=np.int32)
    noff = np.empty((n_cells, nfc), dtype=np.uint16)
    max_rows, zoff, hb = C.c_int32(), C.c_int32(), C.c_int32()
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    nbrc = np.ascontiguousarray(nbr)
    rc = lib.fine_stream_plan_host(n_cells, None, dp(tf), n_tiles, dp(nbrc), nfc, n_cells, tile, n, rows_cap, C.byref(max_rows),
                                   C.byref(zoff), C.byref(hb), dp(rows), dp(noff))
    assert rc == 0
    assert hb.value == tile * n + 2 and zoff.value == hb.value + max_rows.value * n
    for k in range(n_tiles):
        lo, hi = int(tf[k]), int(tf[k + 1])
        par = (lo * n) % 2
        used = rows[k][rows[k] >= 0]
        assert len(set(used.tolist())) == len(used)
        for r in range(rows_cap):
            c = rows[k, r]
            if c >= 0:
                assert (c * n) % 2 == r % 2 and not (lo <= c < hi)
        for cell in range(lo, hi):
            for f in range(nfc):
                nb, o = nbr[cell, f], int(noff[cell, f])
                if nb < 0:
                    assert o == zoff.value
                elif lo <= nb < hi:
                    assert o == 2 - par + (nb - lo) * n
                else:
                    assert (o - hb.value) % n == 0 and 0 <= (o - hb.value) // n < max_rows.value and rows[k, (o - hb.value) // n] == nb


def morton_coords(shape):
    """cell coordinates in the Morton numbering of grid_neighbours(shape, "morton")"""
    dim = len(shape)
    coords = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1).reshape(-1, dim)
    key = np.zeros(len(coords), dtype=np.int64)
    for b in range(10):
        for k in range(dim):
            key |= ((coords[:, k] >> b) & 1) << (b * dim + k)
    return coords[np.argsort(key, kind="stable")]


@pytest.mark.parametrize("shape,n,max_rows,ghost_phase", [((8, 8, 16), 27, 128, "same"), ((8, 8, 16), 27, 128, "alternate"),
                                                          ((8, 8, 16), 27, 40, "same"), ((32, 32), 9, 30, "same"),
                                                          ((32, 32), 25, 24, "alternate")])
def test_fused_plan_splits_tiles_to_the_row_budget(lib, shape, n, max_rows, ghost_phase):
    """build_fused_plan (the fused sharded apply, csrc/pd_finemesh.cu: setup_fine_fused): one tile sequence, interior
    tiles then boundary tiles; a tile whose halo cells would need more than `max_rows` rows (ghost cells lie in their
    owners' export buffers at whatever 16-byte phase those give them; a METIS cut of bench.py's sharded 64^3 mesh
    leaves tiles with 129 / 130 rows against the gather's 128) is halved until it fits.  Contract: the tiles cover the
    sequence in order, each one a run of consecutive cells, no interior tile reads a ghost cell, every neighbour offset
    points at the own row / the halo row of matching phase / the zero row."""
    dim = len(shape)
    nbr_all = grid_neighbours(shape, "morton")
    n_all, nfc = nbr_all.shape
    # rank 0 owns a ragged half: cells of the curve whose (x + y [+ z]) coordinate sum lies below a staircase
    coords = morton_coords(shape)
    own_mask = coords.sum(axis=1) * 2 < sum(shape) + (coords[:, 0] % 3)
    own = np.flatnonzero(own_mask)
    local = np.full(n_all, -1, dtype=np.int64)
    local[own] = np.arange(len(own))
    ghosts = np.unique(nbr_all[own][(nbr_all[own] >= 0) & ~own_mask[np.maximum(nbr_all[own], 0)]])
    local[ghosts] = len(own) + np.arange(len(ghosts))
    n_own, n_total = len(own), len(own) + len(ghosts)
    nbr = np.where(nbr_all[own] >= 0, local[np.maximum(nbr_all[own], 0)], -1).astype(np.int32)
    assert (nbr[nbr_all[own] >= 0] >= 0).all()
    # interior / boundary split by whole blocks of the curve, tiles = the owned cells of a block (consecutive numbers)
    block = own // 64
    reads_ghost = (nbr >= n_own).any(axis=1)
    bnd = np.isin(block, np.unique(block[reads_ghost]))
    inner, outer = np.flatnonzero(~bnd).astype(np.int32), np.flatnonzero(bnd).astype(np.int32)
    assert len(inner) and len(outer)

    def tiles(seq):
        b = block[seq]
        return np.concatenate([[0], np.flatnonzero(np.diff(b)) + 1, [len(seq)]]).astype(np.int32)

    tf1, tf2 = tiles(inner), tiles(outer)
    par = np.zeros(n_total, dtype=np.uint8)
    par[:n_own] = (np.arange(n_own) * n) & 1
    par[n_own:] = 1 if ghost_phase == "same" else (np.arange(len(ghosts)) & 1)
    ns = n_own
    seq, tf, base = np.zeros(ns, np.int32), np.zeros(ns + 1, np.int32), np.zeros(ns, np.int32)
    rows, noff = np.full((ns + 1) * max_rows, -7, np.int32), np.zeros((ns, nfc), np.uint16)
    nt, fg, mr, zoff, um = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    nbrc = np.ascontiguousarray(nbr)
    rc = lib.fine_fused_plan_host(len(inner), dp(inner), len(outer), dp(outer), len(tf1) - 1, dp(tf1), len(tf2) - 1, dp(tf2), dp(nbrc),
                                  nfc, n_total, 64, n, dp(par), max_rows, C.byref(nt), C.byref(fg), C.byref(mr), C.byref(zoff),
                                  C.byref(um), dp(seq), dp(tf), dp(base), dp(rows), C.c_int64(len(rows)), dp(noff))
    assert rc == 0
    n_tiles, R = nt.value, mr.value
    assert R <= max_rows
    if max_rows < 96:
        assert um.value > max_rows and n_tiles > len(tf1) + len(tf2) - 2  # the case the split exists for
    else:
        assert um.value == R and n_tiles == len(tf1) + len(tf2) - 2  # nothing to split: the plan is the unsplit one
    tf = tf[: n_tiles + 1]
    assert tf[0] == 0 and tf[-1] == ns and (np.diff(tf) >= 1).all() and (np.diff(tf) <= 64).all()
    assert (seq == np.concatenate([inner, outer])).all()
    assert 0 < fg.value < n_tiles and tf[fg.value] == len(inner)
    hb = 64 * n + 2
    assert zoff.value == hb + R * n
    rows = rows[: n_tiles * max(R, 1)].reshape(n_tiles, max(R, 1))
    for k in range(n_tiles):
        lo, hi = int(tf[k]), int(tf[k + 1])
        c0 = int(seq[lo])
        assert base[k] == c0 and (seq[lo:hi] == c0 + np.arange(hi - lo)).all()
        tpar = (c0 * n) % 2
        used = rows[k][rows[k] >= 0]
        assert len(set(used.tolist())) == len(used)
        for r, c in enumerate(rows[k]):
            if c >= 0:
                assert par[c] == r % 2 and not (c0 <= c < c0 + hi - lo)
                assert k >= fg.value or c < n_own  # interior tiles never read a ghost cell
        for i in range(lo, hi):
            cell = int(seq[i])
            for f in range(nfc):
                nb, o = int(nbr[cell, f]), int(noff[i, f])
                if nb < 0:
                    assert o == zoff.value
                elif c0 <= nb < c0 + hi - lo:
                    assert o == 2 - tpar + (nb - c0) * n
                else:
                    assert (o - hb) % n == 0 and 0 <= (o - hb) // n < R and rows[k, (o - hb) // n] == nb
