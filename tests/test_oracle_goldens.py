"""Pins the CPU oracle (oracle/) against the reference's own test goldens
(tests/golden/reference_goldens.json, derived from /root/reference/test/polydeal/*.output).
Every test cites the reference test it re-creates."""
import numpy as np
import pytest

from oracle import pyoracle as po
import pd_scenarios as sc


def make_handler(dim, n_refine, groups, fe_degree=1, nq=None, lo=-1.0, hi=1.0, fe_kind=po.FE_DGQ):
    grid = po.Grid.hyper_cube(dim, lo, hi, n_refine)
    ah = po.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(nq if nq is not None else fe_degree + 1)
    ah.distribute_agglomerated_dofs(fe_kind, fe_degree)
    return grid, ah


# --- numbering + sparsity -----------------------------------------------------
def test_sparsity_agglomerated_tria(goldens):
    """test/polydeal/sparsity_agglomerated_tria.cc:30-107 -> 216 rows, bit exact."""
    _, ah = make_handler(2, 3, sc.standard_8x8_agglomerates())
    rp, cols = ah.create_agglomeration_sparsity_pattern()
    gold = goldens["sparsity_agglomerated_tria"]
    assert ah.n_dofs == len(gold) == 216
    for r, row in enumerate(gold):
        assert row[0] == r
        assert cols[rp[r]:rp[r + 1]].tolist() == row[1:], f"row {r}"


def test_hp_structure_01(goldens):
    """test/polydeal/hp_structure_01.cc:40-105: DoF indices + master-cell vertices."""
    grid, ah = make_handler(2, 2, sc.blocks_2x2_of_4x4())
    for p, g in enumerate(goldens["hp_structure_01"]):
        assert ah.master_cell(p) == g["master"]
        assert ah.get_dof_indices(p).tolist() == g["dofs"]
        np.testing.assert_array_equal(grid.cell_vertices(ah.master_cell(p)), np.array(g["vertices"]))


def test_polytope_iterator_dofs(goldens):
    """test/polydeal/polytope_iterator.cc: 64x64 grid, seven 2-cell agglomerates;
    DoF blocks follow the master's active-cell index, not the polytope index."""
    _, ah = make_handler(2, 6, sc.polytope_iterator_agglomerates())
    six = [p for p in range(ah.n_polytopes) if ah.n_faces(p) == 6]
    gold = goldens["polytope_iterator"]
    assert six == [g["index"] for g in gold]
    for g in gold:
        assert ah.get_dof_indices(g["index"]).tolist() == g["dofs"]


# --- face enumeration ----------------------------------------------------------
@pytest.mark.parametrize("name,n_refine,groups", [
    ("agglomerated_neighbors_01", 3, "std"),
    ("agglomerated_neighbors_02", 2, "blocks"),
])
def test_agglomerated_neighbors_faces(goldens, name, n_refine, groups):
    """test/polydeal/agglomerated_neighbors_01.cc / _02.cc: face count and the
    (deal.II cell, local face) list of every non-boundary polytope face."""
    groups = sc.standard_8x8_agglomerates() if groups == "std" else sc.blocks_2x2_of_4x4()
    _, ah = make_handler(2, n_refine, groups)
    gold = goldens[name]
    assert ah.n_polytopes == len(gold)
    for g in gold:
        p = g["index"]
        assert ah.n_faces(p) == g["n_faces"]
        for f in range(g["n_faces"]):
            if ah.at_boundary(p, f):
                assert str(f) not in g["faces"]
            else:
                assert [list(t) for t in ah.interface(p, f)] == g["faces"][str(f)], (p, f)


def test_agglomerated_neighbors_03(goldens):
    """test/polydeal/agglomerated_neighbors_03.cc: neighbor_of_agglomerated_neighbor."""
    _, ah = make_handler(2, 2, sc.blocks_2x2_of_4x4())
    for p, g in enumerate(goldens["agglomerated_neighbors_03"]):
        assert ah.master_cell(p) == g["master"]
        assert ah.n_faces(p) == g["n_faces"]
        assert [ah.neighbor_of_agglomerated_neighbor(p, f) for f in range(g["n_faces"])] == g["nofn"]


def test_continuous_face_01(goldens):
    """test/polydeal/continuous_face_01.cc test0/test1: neighbours, nofn, aligned
    sub-face lists, boundary perimeter (QGauss(1)) and two-sided q-point match."""
    cases = [[list(range(0, 8)), list(range(8, 16))], sc.blocks_2x2_of_4x4()]
    for groups, gold in zip(cases, goldens["continuous_face_01"]):
        _, ah = make_handler(2, 2, groups, nq=1)
        perimeter = 0.0
        assert ah.n_polytopes == len(gold["polytopes"])
        for p, g in enumerate(gold["polytopes"]):
            assert ah.master_cell(p) == g["master"]
            assert ah.n_faces(p) == g["n_faces"]
            for f, gf in enumerate(g["faces"]):
                if ah.at_boundary(p, f):
                    assert gf["neighbor"] is None
                    perimeter += ah.reinit(p, f).JxW.sum()
                else:
                    nb = ah.neighbor(p, f)
                    assert nb == gf["neighbor"]
                    nofn = ah.neighbor_of_agglomerated_neighbor(p, f)
                    assert nofn == gf["nofn"]
                    assert ah.neighbor(nb, nofn) == p
                    assert [[c, lf, ah.master_cell(nb)] for c, lf in ah.interface(p, f)] == gf["subfaces"]
                    f0, f1 = ah.reinit_interface(p, nb, f, nofn)
                    assert np.abs(f0.points - f1.points).max() < 1e-15
        assert perimeter == pytest.approx(gold["perimeter"], abs=1e-14)


def test_continuous_face_02_03_and_distorted(goldens):
    """continuous_face_02 (test0-2 + the METIS scenario, whose partition is recovered from the golden
    itself), continuous_face_03 (58 polytopes on 8x8) and continuous_face_distorted_grid (same topology on a
    distorted grid; boundary vertices stay put, so the perimeter is still 8)."""
    from pd_helpers import check_continuous_face_scenario

    for groups, gold in zip(sc.continuous_face_02_cases(), goldens["continuous_face_02"][:3]):
        _, ah = make_handler(2, 2, groups, nq=1)
        check_continuous_face_scenario(ah, gold)
    metis = goldens["continuous_face_02"][3]
    grid = po.Grid.hyper_cube(2, -1.0, 1.0, 3)
    groups = sc.partition_from_continuous_face_golden(metis, 64, grid.arrays()[2].tolist())
    assert groups is not None and len(groups) == 10 and sorted(c for g in groups for c in g) == list(range(64))
    _, ah = make_handler(2, 3, groups, nq=1)
    check_continuous_face_scenario(ah, metis)
    _, ah = make_handler(2, 3, sc.continuous_face_03_groups(), nq=1)
    check_continuous_face_scenario(ah, goldens["continuous_face_03"][0])
    for groups, gold in zip(sc.continuous_face_distorted_cases(), goldens["continuous_face_distorted_grid"]):
        grid = po.Grid.hyper_cube(2, -1.0, 1.0, 2)
        grid.distort_random(0.25, 7)
        ah = po.AgglomerationHandler(grid)
        for g in groups:
            ah.define_agglomerate(g)
        ah.initialize_fe_values(1)
        ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
        check_continuous_face_scenario(ah, gold)


def test_reinit_cell_face_master_master_and_quad_pts(goldens):
    """reinit_cell_face_master_master.cc (polytope indices of the neighbours per face) and
    reinit_cell_face_quad_pts.cc (master cells of the neighbours, q-points seen from both sides coincide)."""
    from pd_helpers import check_neighbor_lists

    _, ah = make_handler(2, 2, sc.reinit_cell_face_master_master_groups(), nq=1)
    check_neighbor_lists(ah, goldens["reinit_cell_face_master_master"], by_master=False)
    _, ah = make_handler(2, 3, sc.reinit_cell_face_quad_pts_groups(), nq=1)
    check_neighbor_lists(ah, goldens["reinit_cell_face_quad_pts"], by_master=True)
    for p in range(ah.n_polytopes):
        for f in range(ah.n_faces(p)):
            if not ah.at_boundary(p, f):
                nb = ah.neighbor(p, f)
                f0, f1 = ah.reinit_interface(p, nb, f, ah.neighbor_of_agglomerated_neighbor(p, f))
                assert np.abs(f0.points - f1.points).max() < 1e-15


def test_reinit_cell_face_02(goldens):
    """test/polydeal/reinit_cell_face_02.cc: {3,6,9},{15,36,37},{57,60,54},{25,19,22}
    + singletons; neighbour master indices per face, boundary faces flagged."""
    groups = sc._with_singletons([[3, 6, 9], [15, 36, 37], [57, 60, 54], [25, 19, 22]], 64)
    _, ah = make_handler(2, 3, groups, nq=1)
    gold = goldens["reinit_cell_face_02"]
    assert ah.n_polytopes == len(gold)
    for p, g in enumerate(gold):
        assert ah.master_cell(p) == g["master"]
        assert ah.n_faces(p) == g["n_faces"]
        got = [-1 if ah.at_boundary(p, f) else ah.master_cell(ah.neighbor(p, f)) for f in range(g["n_faces"])]
        assert got == g["faces"]


# --- geometry -------------------------------------------------------------------
def test_master_and_slaves_01(goldens):
    """test/polydeal/aggl_handler_master_and_slaves_01.cc: singletons first, then
    {3,6,9,12,13} re-agglomerated; prints master_slave_relationships."""
    grid = po.Grid.hyper_cube(2, -1, 1, 2)
    ah = po.AgglomerationHandler(grid)
    for c in range(16):
        ah.define_agglomerate([c])
    ah.define_agglomerate([3, 6, 9, 12, 13])
    assert [ah.master_slave_value(c) for c in range(16)] == goldens["aggl_handler_master_and_slaves_01"]


def test_agg_handler_bbox(goldens):
    """test/polydeal/agg_handler_bbox_test.cc (2-D {3,6,9,12,13}; 3-D {30,58} on 4^3)."""
    g = goldens["agg_handler_bbox_test"]
    for dim, cells, (lo, hi) in [(2, [3, 6, 9, 12, 13], g[0:2]), (3, [30, 58], g[2:4])]:
        grid = po.Grid.hyper_cube(dim, -1, 1, 2)
        ah = po.AgglomerationHandler(grid)
        p = ah.define_agglomerate(cells)
        blo, bhi = ah.bbox(p)
        assert blo.tolist() == lo and bhi.tolist() == hi


def test_fe_space_on_bbox(goldens):
    """test/polydeal/fe_space_on_bbox.cc: sum of JxW per polytope, QGauss(1)."""
    _, ah = make_handler(2, 3, sc.standard_8x8_agglomerates(), nq=1)
    sums = [ah.reinit(p).JxW.sum() for p in range(4)]
    groups3 = sc._with_singletons([[459, 463]], 512)
    _, ah3 = make_handler(3, 3, groups3, nq=1)
    sums.append(ah3.reinit(0).JxW.sum())
    np.testing.assert_allclose(sums, goldens["fe_space_on_bbox"], rtol=0, atol=1e-15)


def test_reinit_cell_face_01(goldens):
    """test/polydeal/reinit_cell_face_01.cc: perimeter of the first four polytopes."""
    _, ah = make_handler(2, 3, sc.standard_8x8_agglomerates(), nq=1)
    per = [sum(ah.reinit(p, f).JxW.sum() for f in range(ah.n_faces(p))) for p in range(4)]
    np.testing.assert_allclose(per, goldens["reinit_cell_face_01"], rtol=0, atol=1e-14)


# --- 1-D building blocks ----------------------------------------------------------
def test_gauss_and_lobatto_rules():
    for n in range(1, 9):
        x, w = po.gauss_1d(n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        np.testing.assert_allclose(x, (xr + 1) / 2, atol=2e-16)
        np.testing.assert_allclose(w, wr / 2, atol=2e-16)
    np.testing.assert_allclose(po.gauss_lobatto_nodes(3), [0, 0.5, 1], atol=0)
    np.testing.assert_allclose(po.gauss_lobatto_nodes(4), [0, 0.5 - np.sqrt(5) / 10, 0.5 + np.sqrt(5) / 10, 1], atol=1e-16)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("p", [0, 1, 2, 3, 4])
def test_dgq_basis_is_nodal_partition_of_unity(dim, p):
    rng = np.random.default_rng(p * 10 + dim)
    nodes = po.gauss_lobatto_nodes(p + 1) if p > 0 else np.array([0.5])
    n = (p + 1) ** dim
    for i in range(n):
        idx = [(i // (p + 1) ** d) % (p + 1) for d in range(dim)]
        v, _ = po.fe_evaluate(po.FE_DGQ, dim, p, nodes[idx])
        e = np.zeros(n)
        e[i] = 1
        np.testing.assert_allclose(v, e, atol=1e-14)
    x = rng.random(dim)
    v, g = po.fe_evaluate(po.FE_DGQ, dim, p, x)
    assert v.sum() == pytest.approx(1.0, abs=1e-13)
    np.testing.assert_allclose(g.sum(axis=0), 0, atol=1e-12)
    eps = 1e-6
    for d in range(dim):
        xp, xm = x.copy(), x.copy()
        xp[d] += eps
        xm[d] -= eps
        fd = (po.fe_evaluate(po.FE_DGQ, dim, p, xp)[0] - po.fe_evaluate(po.FE_DGQ, dim, p, xm)[0]) / (2 * eps)
        np.testing.assert_allclose(g[:, d], fd, atol=1e-6)


@pytest.mark.parametrize("dim,p,n", [(2, 1, 3), (2, 2, 6), (3, 1, 4), (3, 2, 10), (3, 3, 20)])
def test_agglodgp_is_orthonormal(dim, p, n):
    """source/fe_agglodgp.cc:28-57,91-101: C(p+d,d) L2-orthonormal Legendre modes, mode 0 constant."""
    assert po.lib().po_fe_n_dofs(po.FE_AGGLODGP, dim, p) == n
    x1, w1 = po.gauss_1d(p + 2)
    pts = np.stack(np.meshgrid(*([x1] * dim), indexing="ij"), -1).reshape(-1, dim)
    wts = np.prod(np.stack(np.meshgrid(*([w1] * dim), indexing="ij"), -1).reshape(-1, dim), axis=1)
    V = np.array([po.fe_evaluate(po.FE_AGGLODGP, dim, p, x)[0] for x in pts])
    np.testing.assert_allclose(V.T @ (wts[:, None] * V), np.eye(n), atol=1e-13)
    np.testing.assert_allclose(V[:, 0], 1.0, atol=1e-15)


@pytest.mark.parametrize("which", ["circle-grid.inp", "hyper_ball"])
def test_unstructured_grid_golden(goldens, which):
    """test/polydeal/unstructured_grid.cc: normals[0] of every face of the one polytope made of two cells, and its
    perimeter, on (i) circle-grid.inp read by GridIn::read_ucd and refined once, cells {25, 44} (:116-196, the block
    with six faces) and (ii) GridGenerator::hyper_ball refined once, cells {8, 5} (:28-110, five faces).  Pins, on
    meshes whose neighbours are rotated against each other: the vertex order GridIn hands over, the child order of
    refine_global, the new-vertex rules of the spherical manifold, the face enumeration and the outward normals."""
    g = goldens["unstructured_grid"]
    if which == "circle-grid.inp":
        gold, pair = g["blocks"][0], (25, 44)
        v, cv, nbr = sc.quad_mesh_from_gmsh(g["circle_grid"]["verts"], g["circle_grid"]["quads"], n_refine=1)
    else:
        gold, pair = g["blocks"][1], (5, 8)
        v, cv, nbr = sc.hyper_ball_2d_refined_once()
    ah = po.AgglomerationHandler(po.Grid.from_arrays(v, cv, nbr))
    ah.define_agglomerate(list(pair))
    for c in range(len(cv)):
        if c not in pair:
            ah.define_agglomerate([c])
    ah.initialize_fe_values(1, 1)  # QGauss<2>(1), faces QGauss<1>(1)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
    # the only polytope with that many faces is the agglomerated one
    assert [k for k in range(ah.n_polytopes) if ah.n_faces(k) == gold["n_faces"]] == [0]
    perimeter = 0.0
    for f, n_gold in enumerate(gold["normals"]):
        fv = ah.reinit(0, f)
        perimeter += fv.JxW.sum()
        assert np.abs(fv.normals[0] - np.array(n_gold)).max() < 5e-7 * max(1.0, np.abs(n_gold).max()), (f, fv.normals[0])
    assert perimeter == pytest.approx(gold["perimeter"], rel=5e-6)


@pytest.mark.parametrize("mesh", ["hyper_cube", "hyper_ball"])
def test_fe_collection_agglomeration(mesh):
    """test/polydeal/fe_collection_agglomeration.cc (golden: "Ok" twice): the agglomerates {3,6,9,12,13}, {15,36,37},
    {57,60,54}, {25,19,22} + singletons on (:30-116) hyper_cube(-1,1) refined 3 times with QGauss<2>(1) and (:119-203)
    hyper_ball(radius 2) refined 4 times (1280 cells, SphericalManifold on the boundary) with QGauss<2>(3): the JxW of
    reinit(polytope) over all polytopes sum to GridTools::volume(tria, MappingQ1) -- the sum of the cells' areas -- and
    every sub-cell is integrated exactly once (`total_sum == volume`, :110-113, 197-200)."""
    special = [[3, 6, 9, 12, 13], [15, 36, 37], [57, 60, 54], [25, 19, 22]]
    if mesh == "hyper_cube":
        grid, nq = po.Grid.hyper_cube(2, -1, 1, 3), 1
    else:
        v, cv, nbr = sc.hyper_ball_2d(2.0, 4)
        assert len(cv) == 5 * 4**4
        grid, nq = po.Grid.from_arrays(v, cv, nbr), 3
    v, cv, _ = grid.arrays()
    # GridTools::volume with MappingQ1: shoelace area of every (bilinear) quadrilateral, vertices 0, 1, 3, 2 around it
    q = v[cv][:, [0, 1, 3, 2], :]
    x, y = q[..., 0], q[..., 1]
    area = 0.5 * np.abs((x * np.roll(y, -1, axis=1) - np.roll(x, -1, axis=1) * y).sum(axis=1))
    ah = po.AgglomerationHandler(grid)
    flagged = {c for g in special for c in g}
    for g in special:
        ah.define_agglomerate(sorted(g))  # collect_cells_for_agglomeration: active-cell order (include/poly_utils.h:532-538)
    for c in range(len(cv)):
        if c not in flagged:
            ah.define_agglomerate([c])
    ah.initialize_fe_values(nq)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
    total = 0.0
    for p in range(ah.n_polytopes):
        w = ah.reinit(p).JxW
        cells = ah.get_agglomerate(p)
        assert len(w) == len(cells) * nq * nq
        assert w.sum() == pytest.approx(area[np.asarray(cells)].sum(), rel=1e-13)
        total += w.sum()
    assert ah.n_polytopes == len(cv) - len(flagged) + 4
    if mesh == "hyper_cube":
        assert total == 4.0  # (the reference compares with ==)
    else:
        assert total == pytest.approx(area.sum(), rel=1e-13) and 12.0 < total < 4 * np.pi
