"""The synthetic configurations of SURVEY 8d (BASELINE.json `configs`) as agglomerations of structured grids.
Shared by bench.py, tools/run_config.py and the full-size parity tests; host-side only (numpy)."""
from __future__ import annotations

import numpy as np

# name: dim, cells per direction, block edge, degree, n_q1d, penalty constant (None = library 10 (p+dim)(p+1)), mass
CONFIGS = {
    "A": dict(dim=2, n=256, b=16, p=1, nq=2, C=None, mass=0.0),
    "B": dict(dim=3, n=64, b=8, p=2, nq=3, C=None, mass=0.0),
    "C": dict(dim=3, n=128, b=4, p=3, nq=4, C=None, mass=0.0),
    "D": dict(dim=3, n=256, b=4, p=2, nq=3, C=40.0, mass=0.5),
    "D8": dict(dim=3, n=128, b=4, p=2, nq=3, C=40.0, mass=0.5),  # one eighth of D (per-GPU share at 8 GPUs)
}


def morton_index(dim, i, j, k, n):
    cell = np.zeros_like(i)
    for l in range(int(n).bit_length() - 1):
        cell |= ((i >> l) & 1) << (dim * l)
        cell |= ((j >> l) & 1) << (dim * l + 1)
        if dim == 3:
            cell |= ((k >> l) & 1) << (dim * l + 2)
    return cell


def morton_block_groups(dim, n, b):
    """b^dim blocks of the hyper_cube + refine_global grid (cells in Morton order): (n/b)^dim x b^dim int32 array,
    cells of a block in active-cell order (master = first)."""
    ax = np.arange(n, dtype=np.int64)
    if dim == 2:
        i, j = np.meshgrid(ax, ax, indexing="ij")
        k = np.zeros_like(i)
    else:
        i, j, k = np.meshgrid(ax, ax, ax, indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cell = morton_index(dim, i, j, k, n)
    nb = n // b
    part = ((k // b) * nb + (j // b)) * nb + (i // b)
    order = np.lexsort((cell, part))
    return cell[order].astype(np.int32).reshape(nb**dim, b**dim)


def lex_block_groups(nx, ny, nz, b):
    """b^3 blocks of an nx x ny x nz lexicographic grid (subdivided_hyper_rectangle), cells of a block ascending."""
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cell = (k * ny + j) * nx + i
    part = ((k // b) * (ny // b) + (j // b)) * (nx // b) + (i // b)
    order = np.lexsort((cell, part))
    return cell[order].astype(np.int32).reshape(-1, b**3)


def build_handler(pdl, cfg, world=1):
    """The agglomeration of one SURVEY configuration on `world` GPUs: world == 1 is the configuration itself
    (unit cube, Morton order); world > 1 stacks `world` such cubes along z (n x n x n*world cells, lexicographic),
    so that every GPU's share is one configuration's worth of polytopes (weak scaling)."""
    dim, n, b, p, nq = cfg["dim"], cfg["n"], cfg["b"], cfg["p"], cfg["nq"]
    if world == 1:
        grid = pdl.Grid.hyper_cube(dim, 0.0, 1.0, n.bit_length() - 1)
        groups = morton_block_groups(dim, n, b) if b > 1 else np.arange(n**dim, dtype=np.int32).reshape(-1, 1)
    elif b == 1:
        # fine mesh (every cell its own element): numbered along the Morton curve of the stacked box, as the cells of
        # a p4est-distributed mesh are on every rank (the tiled fine-mesh kernel stages contiguous runs of the curve)
        assert dim == 3 and (world & (world - 1)) == 0
        grid = morton_numbered_box(pdl, n, n, n * world, (1.0, 1.0, float(world)))
        groups = np.arange(n * n * n * world, dtype=np.int32).reshape(-1, 1)
    else:
        assert dim == 3
        grid = pdl.Grid.structured(dim, (n, n, n * world), 0.0, (1.0, 1.0, float(world)), order=1)
        groups = lex_block_groups(n, n, n * world, b)
    ah = pdl.AgglomerationHandler(grid)
    ah.define_agglomerates(groups)
    ah.initialize_fe_values(nq)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, p)
    return ah


def morton_numbered_box(pdl, nx, ny, nz, hi):
    """An nx x ny x nz Cartesian hex grid on [0, hi] whose cells are numbered along the Morton (z-order) curve of the
    whole box -- what a p4est-distributed deal.II mesh looks like on every rank -- as a pdl.Grid built from arrays
    (vertices lexicographic, deal.II vertex order in a cell, neighbours behind the six faces).  nx, ny, nz powers of two."""
    i, j, k = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64), np.arange(nz, dtype=np.int64), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    key = np.zeros_like(i)
    for b in range(max(nx, ny, nz).bit_length()):
        key |= ((i >> b) & 1) << (3 * b) | ((j >> b) & 1) << (3 * b + 1) | ((k >> b) & 1) << (3 * b + 2)
    order = np.argsort(key, kind="stable")
    i, j, k = i[order], j[order], k[order]
    n_cells = nx * ny * nz
    cell_at = np.empty((nx, ny, nz), dtype=np.int64)
    cell_at[i, j, k] = np.arange(n_cells)
    vid = lambda a, b, c: (c * (ny + 1) + b) * (nx + 1) + a
    cell_verts = np.stack([vid(i + (v & 1), j + ((v >> 1) & 1), k + ((v >> 2) & 1)) for v in range(8)], axis=1).astype(np.int32)
    nbr = -np.ones((n_cells, 6), dtype=np.int32)
    for f, (di, dj, dk) in enumerate([(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]):
        a, b, c = i + di, j + dj, k + dk
        ok = (a >= 0) & (a < nx) & (b >= 0) & (b < ny) & (c >= 0) & (c < nz)
        nbr[ok, f] = cell_at[a[ok], b[ok], c[ok]]
    x = np.linspace(0.0, hi[0], nx + 1)
    y = np.linspace(0.0, hi[1], ny + 1)
    z = np.linspace(0.0, hi[2], nz + 1)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    return pdl.Grid.from_arrays(verts, cell_verts, nbr)
