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
    if shape.startswith("metis"):  # GridTools::partition_triangulation(k, tria, metis): the reference's main shape
        import polydeal_b200 as pdl

        v, cv, nbr = ogrid.arrays()
        return pdl.metis_agglomerates(pdl.Grid.from_arrays(v, cv, nbr), int(shape[5:]))
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


def check_continuous_face_scenario(ah, gold, reinit=True):
    """The walk of test/polydeal/continuous_face_0x.cc (perimeter_test, test_neighbors, test_face_qpoints)
    against one parsed golden scenario.  `ah` is the oracle handler or the product's host mirror (which
    has no reinit: reinit=False skips perimeter and q-point alignment)."""
    perimeter = 0.0
    assert ah.n_polytopes == len(gold["polytopes"])
    for p, g in enumerate(gold["polytopes"]):
        assert ah.master_cell(p) == g["master"]
        assert ah.n_faces(p) == g["n_faces"]
        for f, gf in enumerate(g["faces"]):
            if ah.at_boundary(p, f):
                assert gf["neighbor"] is None
                if reinit:
                    perimeter += ah.reinit(p, f).JxW.sum()
            else:
                nb = ah.neighbor(p, f)
                assert nb == gf["neighbor"]
                nofn = ah.neighbor_of_agglomerated_neighbor(p, f)
                assert nofn == gf["nofn"]
                assert ah.neighbor(nb, nofn) == p
                if gf["subfaces"]:  # continuous_face_03 prints neighbour / nofn only
                    assert [[c, lf, ah.master_cell(nb)] for c, lf in ah.interface(p, f)] == gf["subfaces"]
                if reinit:
                    f0, f1 = ah.reinit_interface(p, nb, f, nofn)
                    assert np.abs(f0.points - f1.points).max() < 1e-15
    if reinit:
        assert abs(perimeter - gold["perimeter"]) <= 1e-13


def check_neighbor_lists(ah, gold, by_master):
    """'<polytope> has n faces' + the neighbour of every non-boundary face in face order
    (reinit_cell_face_master_master: polytope indices; reinit_cell_face_quad_pts: master cell indices)."""
    assert ah.n_polytopes == len(gold)
    for p, g in enumerate(gold):
        assert (ah.master_cell(p) if by_master else p) == g["id"]
        assert ah.n_faces(p) == g["n_faces"]
        got = [ah.master_cell(ah.neighbor(p, f)) if by_master else ah.neighbor(p, f)
               for f in range(g["n_faces"]) if not ah.at_boundary(p, f)]
        assert got == g["neighbors"]
