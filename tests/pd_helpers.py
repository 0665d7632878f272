"""Build the SAME scenario for the CPU oracle (checker) and for the product under test."""
from __future__ import annotations

import numpy as np

import pd_scenarios as sc
from oracle import pyoracle as po


def oracle_handler(dim, n, groups, p, nq, lo=0.0, hi=1.0, order=0, distort=None, nq_face=None, fe_kind=po.FE_DGQ):
    grid = po.Grid(dim, n, lo, hi, order)
    if distort:
        grid.distort_random(*distort)
    ah = po.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(nq, nq_face)
    ah.distribute_agglomerated_dofs(fe_kind, p)
    return grid, ah


def product_handler(ogrid, groups, p, nq, nq_face=None):
    """Product-side handler on the very same vertices / cells as the oracle grid."""
    import polydeal_b200 as pdl

    v, cv, nb = ogrid.arrays()
    grid = pdl.Grid.from_arrays(v, cv, nb)
    ah = pdl.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(nq, nq_face)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, p)
    return grid, ah


def groups_for(shape, dim, n, ogrid, seed=0):
    if shape == "singletons":
        return [[c] for c in range(ogrid.n_cells)]
    if shape.startswith("blocks"):
        return sc.block_partition(dim, n, int(shape[6:]), ogrid.order)
    if shape.startswith("random"):
        _, _, nbr = ogrid.arrays()
        return sc.random_partition(ogrid.n_cells, nbr, int(shape[6:]), seed)
    raise ValueError(shape)


def src_vector(n):
    """Deterministic source vector of SURVEY 8d: sin(0.37 i) + 0.01 (i mod 7)."""
    i = np.arange(n, dtype=np.float64)
    return np.sin(0.37 * i) + 0.01 * (np.arange(n) % 7)


def assert_blocks_close(got, ref, n, rowptr, tol=1e-12):
    """Parity bar: |got - ref| <= tol * max|ref block| for every n x n block entry
    (1e-12 relative to the block, north_star), reported with the worst offender."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    nrows = len(rowptr) - 1
    worst = 0.0
    for b in range(nrows // n):
        s, e = rowptr[b * n], rowptr[(b + 1) * n]
        nb = (rowptr[b * n + 1] - s) // n
        R = ref[s:e].reshape(n, nb, n)
        G = got[s:e].reshape(n, nb, n)
        scale = np.abs(R).max(axis=(0, 2))
        scale = np.where(scale > 0, scale, np.abs(ref).max())
        err = (np.abs(G - R).max(axis=(0, 2)) / scale).max()
        worst = max(worst, err)
    assert worst <= tol, f"worst block-relative error {worst:.3e} > {tol}"
    return worst
