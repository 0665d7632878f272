"""CPU-only tests of the product's host side: the C-ABI library loads and exports
every declared symbol, the host mirror of AgglomerationHandler reproduces the
reference goldens and agrees bit-exactly with the oracle on irregular partitions,
and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pd_scenarios as sc
import polydeal_b200 as pdl
from oracle import pyoracle as po
from pd_helpers import groups_for, oracle_handler, product_handler
from polydeal_b200 import _capi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "polydeal_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pdh?_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 50
    L = C.CDLL(K.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(K.SIGNATURES), declared ^ set(K.SIGNATURES)


def test_rule_tables_match_oracle():
    for n in range(1, 9):
        x, w = np.empty(n), np.empty(n)
        K.check(K.lib().pd_quadrature_rule_1d(n, x.ctypes.data, w.ctypes.data))
        xo, wo = po.gauss_1d(n)
        np.testing.assert_allclose(x, xo, rtol=0, atol=2e-16)
        np.testing.assert_allclose(w, wo, rtol=0, atol=2e-16)
    for p in range(1, 6):
        x = np.empty(p + 1)
        K.check(K.lib().pd_dgq_nodes_1d(p, x.ctypes.data))
        np.testing.assert_allclose(x, po.gauss_lobatto_nodes(p + 1), rtol=0, atol=2e-16)


def product_from_groups(dim, n_refine, groups, p=1, nq=2, lo=-1.0, hi=1.0):
    grid = pdl.Grid.hyper_cube(dim, lo, hi, n_refine)
    ah = pdl.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(nq)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, p)
    return grid, ah


def test_sparsity_agglomerated_tria_golden(goldens):
    """test/polydeal/sparsity_agglomerated_tria.cc through the product's host mirror."""
    _, ah = product_from_groups(2, 3, sc.standard_8x8_agglomerates())
    rp, cols = ah.create_agglomeration_sparsity_pattern()
    gold = goldens["sparsity_agglomerated_tria"]
    assert ah.n_dofs == len(gold)
    for r, row in enumerate(gold):
        assert cols[rp[r]:rp[r + 1]].tolist() == row[1:]


@pytest.mark.parametrize("name,n_refine,which", [("agglomerated_neighbors_01", 3, "std"), ("agglomerated_neighbors_02", 2, "blocks")])
def test_agglomerated_neighbors_goldens(goldens, name, n_refine, which):
    groups = sc.standard_8x8_agglomerates() if which == "std" else sc.blocks_2x2_of_4x4()
    _, ah = product_from_groups(2, n_refine, groups)
    for g in goldens[name]:
        p = g["index"]
        assert ah.n_faces(p) == g["n_faces"]
        for f in range(g["n_faces"]):
            if ah.at_boundary(p, f):
                assert str(f) not in g["faces"]
            else:
                assert [list(t) for t in ah.interface(p, f)] == g["faces"][str(f)]


def test_nofn_and_dofs_goldens(goldens):
    _, ah = product_from_groups(2, 2, sc.blocks_2x2_of_4x4())
    for p, g in enumerate(goldens["agglomerated_neighbors_03"]):
        assert ah.master_cell(p) == g["master"]
        assert [ah.neighbor_of_agglomerated_neighbor(p, f) for f in range(g["n_faces"])] == g["nofn"]
    for p, g in enumerate(goldens["hp_structure_01"]):
        assert ah.get_dof_indices(p).tolist() == g["dofs"]
    _, ah = product_from_groups(2, 6, sc.polytope_iterator_agglomerates())
    for g in goldens["polytope_iterator"]:
        assert ah.n_faces(g["index"]) == 6
        assert ah.get_dof_indices(g["index"]).tolist() == g["dofs"]


def test_bbox_golden(goldens):
    g = goldens["agg_handler_bbox_test"]
    for dim, cells, (lo, hi) in [(2, [3, 6, 9, 12, 13], g[0:2]), (3, [30, 58], g[2:4])]:
        grid = pdl.Grid.hyper_cube(dim, -1, 1, 2)
        ah = pdl.AgglomerationHandler(grid)
        p = ah.define_agglomerate(cells)
        blo, bhi = ah.bbox(p)
        assert blo.tolist() == lo and bhi.tolist() == hi


@pytest.mark.parametrize("dim,n,shape,order", [(2, 16, "random23", 0), (2, 12, "random17", 1), (3, 8, "random40", 0), (3, 8, "blocks2", 0), (3, 6, "random11", 1)])
def test_host_mirror_matches_oracle_bit_exactly(dim, n, shape, order):
    """Numbering, face enumeration, aligned sub-face lists, nofn, sparsity: identical to
    the oracle's literal restatement on irregular agglomerates."""
    ogrid = po.Grid(dim, n, 0.0, 1.0, order)
    groups = groups_for(shape, dim, n, ogrid, seed=dim * 100 + n)
    _, oah = oracle_handler(dim, n, groups, 1, 2, order=order)
    _, pah = product_handler(oah.grid, groups, 1, 2)
    assert pah.n_polytopes == oah.n_polytopes and pah.n_dofs == oah.n_dofs
    for p in range(oah.n_polytopes):
        assert pah.master_cell(p) == oah.master_cell(p)
        assert pah.get_agglomerate(p).tolist() == oah.get_agglomerate(p).tolist()
        assert pah.get_dof_indices(p).tolist() == oah.get_dof_indices(p).tolist()
        assert pah.n_faces(p) == oah.n_faces(p)
        assert pah.diameter(p) == oah.diameter(p) and pah.volume(p) == oah.volume(p)
        np.testing.assert_array_equal(np.concatenate(pah.bbox(p)), np.concatenate(oah.bbox(p)))
        for f in range(oah.n_faces(p)):
            assert pah.at_boundary(p, f) == oah.at_boundary(p, f)
            assert pah.neighbor(p, f) == oah.neighbor(p, f)
            assert pah.neighbor_of_agglomerated_neighbor(p, f) == oah.neighbor_of_agglomerated_neighbor(p, f)
            assert pah.interface(p, f) == oah.interface(p, f)
    rp, cols = pah.create_agglomeration_sparsity_pattern()
    orp, ocols = oah.create_agglomeration_sparsity_pattern()
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(cols, ocols)


def test_flatten_descriptor_is_consistent():
    ogrid = po.Grid(3, 4, 0.0, 1.0, 0)
    groups = groups_for("random9", 3, 4, ogrid, seed=3)
    _, oah = oracle_handler(3, 4, groups, 2, 3)
    _, pah = product_handler(oah.grid, groups, 2, 3)
    d = pah.flatten()
    assert (d.dim, d.fe_degree, d.n_q1d, d.n_q1d_face, d.n_polytopes) == (3, 2, 3, 3, 9)
    A = np.ctypeslib.as_array(d.iface_polyA, (d.n_ifaces,))
    B = np.ctypeslib.as_array(d.iface_polyB, (d.n_ifaces,))
    sub_ptr = np.ctypeslib.as_array(d.iface_sub_ptr, (d.n_ifaces + 1,))
    sigma = np.ctypeslib.as_array(d.sub_sigma, (sub_ptr[-1],))
    # every interior pair appears once, from the side with the smaller master id (poly_utils.h:2089)
    pairs = set()
    for a, b in zip(A, B):
        if b >= 0:
            assert pah.master_cell(a) < pah.master_cell(b)
            assert (a, b) not in pairs and (b, a) not in pairs
            pairs.add((a, b))
    n_int = sum(1 for p in range(9) for f in range(pah.n_faces(p)) if not pah.at_boundary(p, f))
    assert 2 * len(pairs) == n_int
    # library penalty: 10 (p+dim)(p+1) / diameter(visitor)
    for k, a in enumerate(A):
        np.testing.assert_allclose(sigma[sub_ptr[k]:sub_ptr[k + 1]], 10.0 * (2 + 3) * 3 / pah.diameter(a), rtol=1e-15)


def test_error_behaviour():
    grid = pdl.Grid.hyper_cube(2, 0, 1, 2)
    ah = pdl.AgglomerationHandler(grid)
    with pytest.raises(pdl.PolydealError, match="No cells to be agglomerated"):
        ah.define_agglomerate([])
    ah.define_agglomerate([0, 1])
    with pytest.raises(pdl.PolydealError, match="already belongs"):
        ah.define_agglomerate([1, 2])
    with pytest.raises(pdl.PolydealError, match="belongs to no agglomerate"):
        ah.distribute_agglomerated_dofs(pdl.FE_DGQ, 1)
    for c in range(2, 16):
        ah.define_agglomerate([c])
    with pytest.raises(pdl.PolydealError, match="forgot to distribute"):
        ah.n_faces(0)
    with pytest.raises(pdl.PolydealError, match="only DGQ and DGP"):
        ah.distribute_agglomerated_dofs(2, 1)  # FE_SimplexDGP is not on this path
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, 1)
    with pytest.raises(pdl.PolydealError, match="face index out of range"):
        ah.neighbor(0, 99)
    with pytest.raises(pdl.PolydealError, match="initialize_fe_values"):
        ah.flatten()
    with pytest.raises(pdl.PolydealError, match="Morton"):
        pdl.Grid.structured(2, 6, 0, 1, order=0)


@pytest.mark.skipif(K.lib().pd_device_count() > 0, reason="GPU present: the loud-failure path is for CPU-only hosts")
def test_compute_fails_loudly_without_gpu():
    _, ah = product_from_groups(2, 2, sc.blocks_2x2_of_4x4())
    with pytest.raises(pdl.PolydealError, match="no CPU fallback") as e:
        pdl.assemble_dg_matrix(ah)
    assert e.value.code == K.PD_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    """The product path must not route through oracle/ (task rule)."""
    pkg = os.path.join(ROOT, "polydeal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, os.path.join(dirpath, f)


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 3), (3, 1), (3, 2)])
def test_fe_agglodgp_numbering_and_sparsity_match_the_oracle(dim, p):
    """FE_AggloDGP<dim>(p): C(p+dim, dim) DoFs per polytope, same block numbering and pattern rules."""
    import math

    n = 4
    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    groups = groups_for("random3", dim, n, ogrid, 2)
    _, oah = oracle_handler(dim, n, groups, p, p + 1, fe_kind=po.FE_AGGLODGP)
    v, cv, nb = oah.grid.arrays()
    pah = pdl.AgglomerationHandler(pdl.Grid.from_arrays(v, cv, nb))
    for g in groups:
        pah.define_agglomerate(g)
    pah.initialize_fe_values(p + 1)
    pah.distribute_agglomerated_dofs(pdl.FE_AGGLODGP, p)
    assert pah.n_dofs_per_cell == oah.n_dofs_per_cell == math.comb(p + dim, dim)
    assert pah.n_dofs == oah.n_dofs
    for k in range(oah.n_polytopes):
        assert np.array_equal(pah.get_dof_indices(k), oah.get_dof_indices(k))
    rp, cols = pah.create_agglomeration_sparsity_pattern()
    orp, ocols = oah.create_agglomeration_sparsity_pattern()
    assert np.array_equal(rp, orp) and np.array_equal(cols, ocols)
    assert pah.flatten().fe_kind == pdl.FE_AGGLODGP


def test_c_abi_from_plain_c(tmp_path):
    """include/polydeal_b200.h is C99 and the library is callable from a plain C program:
    tests/c_abi/c_abi_smoke.c walks the host mirror and checks the no-GPU answer of pd_create."""
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "c_abi_smoke")
    lib_dir = os.path.dirname(K.LIB_PATH)
    subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.c"), "-L", lib_dir, "-lpolydeal_b200",
                    "-Wl,-rpath," + lib_dir, "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C ABI OK" in r.stdout, r.stdout + r.stderr


def build_shim_program(tmp_path):
    """g++ -std=c++17 on tests/c_abi/shim_sip_loop.cpp against include/polydeal_b200_shim.hpp (no CUDA headers)."""
    import shutil
    import subprocess

    cxx = shutil.which("g++") or shutil.which("c++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "shim_sip_loop")
    lib_dir = os.path.dirname(K.LIB_PATH)
    subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "shim_sip_loop.cpp"), "-L", lib_dir, "-lpolydeal_b200",
                    "-Wl,-rpath," + lib_dir, "-o", exe], check=True, capture_output=True, text=True)
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    """include/polydeal_b200_shim.hpp (the reference-named C++ surface) goes through a compiler; a reference-style
    loop written against it walks the host mirror, and its first device call is refused without a GPU."""
    import subprocess

    exe = build_shim_program(tmp_path)
    if K.lib().pd_device_count() > 0:
        pytest.skip("a GPU is visible: the full loop runs in the -m gpu suite")
    r = subprocess.run([exe, "--host"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "SHIM HOST OK" in r.stdout, r.stdout + r.stderr


def test_more_face_goldens_through_the_host_mirror(goldens):
    """continuous_face_02 (incl. the METIS scenario recovered from its golden), continuous_face_03,
    continuous_face_distorted_grid, reinit_cell_face_master_master, reinit_cell_face_quad_pts: face
    numbering, neighbours, nofn and aligned (cell, face) lists from the product's host mirror."""
    from pd_helpers import check_continuous_face_scenario, check_neighbor_lists

    for groups, gold in zip(sc.continuous_face_02_cases(), goldens["continuous_face_02"][:3]):
        _, ah = product_from_groups(2, 2, groups)
        check_continuous_face_scenario(ah, gold, reinit=False)
    metis = goldens["continuous_face_02"][3]
    nbr = pdl.Grid.hyper_cube(2, -1.0, 1.0, 3).arrays()[2].tolist()
    _, ah = product_from_groups(2, 3, sc.partition_from_continuous_face_golden(metis, 64, nbr))
    check_continuous_face_scenario(ah, metis, reinit=False)
    _, ah = product_from_groups(2, 3, sc.continuous_face_03_groups())
    check_continuous_face_scenario(ah, goldens["continuous_face_03"][0], reinit=False)
    for groups, gold in zip(sc.continuous_face_distorted_cases(), goldens["continuous_face_distorted_grid"]):
        grid = pdl.Grid.hyper_cube(2, -1.0, 1.0, 2)
        grid.distort_random(0.25, 7)
        ah = pdl.AgglomerationHandler(grid)
        for g in groups:
            ah.define_agglomerate(g)
        ah.initialize_fe_values(1)
        ah.distribute_agglomerated_dofs(pdl.FE_DGQ, 1)
        check_continuous_face_scenario(ah, gold, reinit=False)
    _, ah = product_from_groups(2, 2, sc.reinit_cell_face_master_master_groups())
    check_neighbor_lists(ah, goldens["reinit_cell_face_master_master"], by_master=False)
    _, ah = product_from_groups(2, 3, sc.reinit_cell_face_quad_pts_groups())
    check_neighbor_lists(ah, goldens["reinit_cell_face_quad_pts"], by_master=True)


def test_locally_owned_polytope_goldens(goldens):
    """locally_owned_polytope_02/03/04 (mpirun=3, p4est keeps refinement families together: rank 0 owns cells
    0-3 of the 4x4 grid): the (cell, local face) lists of every interface of the polytopes rank 0 owns, towards
    local and ghosted neighbours alike, in face order."""
    cases = {
        "02": [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]],
        "03": [[2 * k, 2 * k + 1] for k in range(8)],
        "04": [[c] for c in range(16)],
    }
    for key, groups in cases.items():
        _, ah = product_from_groups(2, 2, groups, lo=0.0, hi=1.0)
        got = []
        for p, g in enumerate(groups):
            if max(g) > 3:  # not on rank 0
                continue
            for f in range(ah.n_faces(p)):
                if not ah.at_boundary(p, f):
                    got += [[c, lf] for c, lf in ah.interface(p, f)]
        assert got == goldens["locally_owned_polytope"][key], key


def test_metis_agglomeration_and_sharding():
    """pdh_partition_graph (METIS of the CUDA toolkit, called like deal.II's SparsityTools::partition): the METIS
    agglomeration shape of the reference's examples and the METIS distribution of polytopes over ranks.  METIS
    versions differ, so partitions are inputs of the path, never outputs to match: checked here are the
    contracts (every cell / polytope exactly once, all parts used, balance, mostly face-connected parts)."""
    from polydeal_b200 import distributed as pdd

    for dim, n_ref, k in [(2, 4, 10), (2, 6, 256), (3, 3, 40)]:
        grid = pdl.Grid.hyper_cube(dim, 0.0, 1.0, n_ref)
        groups = pdl.metis_agglomerates(grid, k)
        assert len(groups) == k
        assert sorted(c for g in groups for c in g) == list(range(grid.n_cells))
        sizes = np.array([len(g) for g in groups])
        assert sizes.max() <= 1.35 * grid.n_cells / k + 1
        ah = pdl.AgglomerationHandler(grid)
        for g in groups:
            ah.define_agglomerate(g)
        ah.initialize_fe_values(2)
        ah.distribute_agglomerated_dofs(pdl.FE_DGQ, 1)
        for world in (2, 8):
            owner = pdd.partition_by_metis(ah, world)
            cnt = np.bincount(owner, minlength=world)
            assert owner.shape == (k,) and cnt.min() > 0 and cnt.max() <= 1.5 * k / world + 1
            parts = [pdd.LocalPart(ah, owner, r) for r in range(world)]
            assert sum(p.n_owned for p in parts) == k
            for r, p in enumerate(parts):  # what r expects from s is what s sends to r
                for s_ in range(world):
                    assert p.recv_counts[s_] == parts[s_].send_counts[r]


def test_host_mirror_on_an_unstructured_mesh(goldens):
    """The input grid of test/polydeal/fully_distributed_poisson_sanity_check_02.cc (input_grids/square.msh refined once:
    364 quadrilaterals, neighbours rotated against each other -- no opposite-face rule, cell->neighbor_of_neighbor
    looked up), 30 agglomerates: numbering, face enumeration, aligned sub-face lists, nofn and sparsity of the host
    mirror are those of the oracle's literal restatement, and the flattened descriptor lists every cut from both
    ranks with the same sub-faces in the same order."""
    from polydeal_b200 import distributed as pdd

    g = goldens["fully_distributed_poisson_sanity_check_02"]
    v, cv, nbr = sc.quad_mesh_from_gmsh(g["input_grid"]["verts"], g["input_grid"]["quads"], n_refine=1)
    assert len(cv) == int(g["n_cells"][0])
    rank_groups = sc.random_partition(len(nbr), nbr, 3, seed=0)
    groups, owner = [], []
    for r, cells in enumerate(rank_groups):
        cells = np.array(sorted(cells))
        local = -np.ones(len(nbr), dtype=np.int64)
        local[cells] = np.arange(len(cells))
        sub_nbr = np.where(nbr[cells] >= 0, local[np.maximum(nbr[cells], 0)], -1)
        for gr in sc.random_partition(len(cells), sub_nbr, 10, seed=1 + r):
            groups.append([int(cells[i]) for i in gr])
            owner.append(r)
    oah = po.AgglomerationHandler(po.Grid.from_arrays(v, cv, nbr))
    pah = pdl.AgglomerationHandler(pdl.Grid.from_arrays(v, cv, nbr))
    for ah, kind in ((oah, po.FE_DGQ), (pah, pdl.FE_DGQ)):
        for gr in groups:
            ah.define_agglomerate(gr)
        ah.initialize_fe_values(3)
        ah.distribute_agglomerated_dofs(kind, 1)
    assert pah.n_polytopes == oah.n_polytopes == 30 and pah.n_dofs == oah.n_dofs
    rotated = 0
    for p in range(oah.n_polytopes):
        assert pah.master_cell(p) == oah.master_cell(p)
        assert pah.get_agglomerate(p).tolist() == oah.get_agglomerate(p).tolist()
        assert pah.get_dof_indices(p).tolist() == oah.get_dof_indices(p).tolist()
        assert pah.n_faces(p) == oah.n_faces(p)
        np.testing.assert_array_equal(np.concatenate(pah.bbox(p)), np.concatenate(oah.bbox(p)))
        for f in range(oah.n_faces(p)):
            assert pah.at_boundary(p, f) == oah.at_boundary(p, f)
            assert pah.neighbor(p, f) == oah.neighbor(p, f)
            assert pah.neighbor_of_agglomerated_neighbor(p, f) == oah.neighbor_of_agglomerated_neighbor(p, f)
            assert pah.interface(p, f) == oah.interface(p, f)
            if not pah.at_boundary(p, f):  # the two sides list the same sub-faces, each through its own cell
                q, nofn = pah.neighbor(p, f), pah.neighbor_of_agglomerated_neighbor(p, f)
                mine, theirs = pah.interface(p, f), pah.interface(q, nofn)
                assert len(mine) == len(theirs)
                for (c, lf), (c2, lf2) in zip(mine, theirs):
                    assert nbr[c, lf] == c2 and nbr[c2, lf2] == c
                    rotated += lf2 != (lf ^ 1)
    assert rotated > 0  # the mesh does exercise the general neighbor_of_neighbor
    rp, cols = pah.create_agglomeration_sparsity_pattern()
    orp, ocols = oah.create_agglomeration_sparsity_pattern()
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(cols, ocols)
    # local descriptors of the three ranks: every cut interface from both owners, same sub-faces in the same order
    owner = np.array(owner, dtype=np.int32)
    seen = {}
    for r in range(3):
        part = pdd.LocalPart(pah, owner, r, penalty_constant=1.0, h_rule=pdl.H_CONSTANT, h_const=1.0)
        d = part.desc
        A = np.ctypeslib.as_array(d.iface_polyA, (d.n_ifaces,))
        B = np.ctypeslib.as_array(d.iface_polyB, (d.n_ifaces,))
        ptr = np.ctypeslib.as_array(d.iface_sub_ptr, (d.n_ifaces + 1,))
        sub_c = np.ctypeslib.as_array(d.sub_cell, (int(ptr[-1]),))
        sub_f = np.ctypeslib.as_array(d.sub_face, (int(ptr[-1]),))
        verts = np.ctypeslib.as_array(d.verts, (d.n_verts, 2))
        lcv = np.ctypeslib.as_array(d.cell_verts, (d.n_cells, 4))
        glob = part.local_poly_global
        for f in range(d.n_ifaces):
            if B[f] < 0 or owner[glob[A[f]]] == owner[glob[B[f]]]:
                continue
            mids = []
            for s_ in range(ptr[f], ptr[f + 1]):
                axis, side = int(sub_f[s_]) // 2, int(sub_f[s_]) % 2
                mids.append(verts[lcv[sub_c[s_]][[k for k in range(4) if ((k >> axis) & 1) == side]]].mean(axis=0))
            seen.setdefault(tuple(sorted((int(glob[A[f]]), int(glob[B[f]])))), {})[r] = np.array(mids)
    assert seen
    for key, views in seen.items():
        assert set(views) == {int(owner[key[0]]), int(owner[key[1]])}, key
        a, b = views.values()
        assert a.shape == b.shape and np.abs(a - b).max() < 1e-15, key


@pytest.mark.parametrize("which", ["circle-grid.inp", "hyper_ball"])
def test_unstructured_grid_golden_through_the_host_mirror(goldens, which):
    """test/polydeal/unstructured_grid.cc through the product's host mirror: the two-cell polytope has the golden
    number of faces, and numbering, face enumeration, neighbours, nofn, aligned sub-face lists and sparsity are those
    of the oracle (which reproduces the golden's normals and perimeter, tests/test_oracle_goldens.py)."""
    g = goldens["unstructured_grid"]
    if which == "circle-grid.inp":
        gold, pair = g["blocks"][0], (25, 44)
        v, cv, nbr = sc.quad_mesh_from_gmsh(g["circle_grid"]["verts"], g["circle_grid"]["quads"], n_refine=1)
    else:
        gold, pair = g["blocks"][1], (5, 8)
        v, cv, nbr = sc.hyper_ball_2d_refined_once()
    oah = po.AgglomerationHandler(po.Grid.from_arrays(v, cv, nbr))
    pah = pdl.AgglomerationHandler(pdl.Grid.from_arrays(v, cv, nbr))
    for ah, kind in ((oah, po.FE_DGQ), (pah, pdl.FE_DGQ)):
        ah.define_agglomerate(list(pair))
        for c in range(len(cv)):
            if c not in pair:
                ah.define_agglomerate([c])
        ah.initialize_fe_values(1, 1)
        ah.distribute_agglomerated_dofs(kind, 1)
    assert pah.n_faces(0) == gold["n_faces"] and pah.get_agglomerate(0).tolist() == [pair[1], pair[0]]  # slave, master
    assert pah.n_polytopes == oah.n_polytopes and pah.n_dofs == oah.n_dofs
    for p in range(oah.n_polytopes):
        assert pah.master_cell(p) == oah.master_cell(p)
        assert pah.get_dof_indices(p).tolist() == oah.get_dof_indices(p).tolist()
        assert pah.n_faces(p) == oah.n_faces(p)
        np.testing.assert_array_equal(np.concatenate(pah.bbox(p)), np.concatenate(oah.bbox(p)))
        for f in range(oah.n_faces(p)):
            assert pah.at_boundary(p, f) == oah.at_boundary(p, f)
            assert pah.neighbor(p, f) == oah.neighbor(p, f)
            assert pah.neighbor_of_agglomerated_neighbor(p, f) == oah.neighbor_of_agglomerated_neighbor(p, f)
            assert pah.interface(p, f) == oah.interface(p, f)
    rp, cols = pah.create_agglomeration_sparsity_pattern()
    orp, ocols = oah.create_agglomeration_sparsity_pattern()
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(cols, ocols)
    d = pah.flatten()
    assert d.n_polytopes == oah.n_polytopes and d.n_ifaces > 0


@pytest.mark.parametrize("mesh,n_refine", [("square.msh", 0), ("square.msh", 1), ("circle-grid.inp", 0), ("circle-grid.inp", 2), ("hyper_ball", 1)])
def test_host_mirror_matches_oracle_on_random_agglomerations_of_unstructured_meshes(goldens, mesh, n_refine):
    """face enumeration, neighbours, nofn, aligned sub-face lists and sparsity on random connected agglomerations of
    the reference tests' unstructured input grids (neighbours rotated against each other), host mirror vs oracle;
    and the oracle's matrix keeps the constants in its kernel."""
    if mesh == "square.msh":
        src = goldens["fully_distributed_poisson_sanity_check_02"]["input_grid"]
        v, cv, nbr = sc.quad_mesh_from_gmsh(src["verts"], src["quads"], n_refine=n_refine)
    elif mesh == "circle-grid.inp":
        src = goldens["unstructured_grid"]["circle_grid"]
        v, cv, nbr = sc.quad_mesh_from_gmsh(src["verts"], src["quads"], n_refine=n_refine)
    else:
        v, cv, nbr = sc.hyper_ball_2d_refined_once()
    for seed in range(3):
        groups = sc.random_partition(len(cv), nbr, max(2, len(cv) // (3 + 2 * seed)), seed=seed)
        oah = po.AgglomerationHandler(po.Grid.from_arrays(v, cv, nbr))
        pah = pdl.AgglomerationHandler(pdl.Grid.from_arrays(v, cv, nbr))
        for ah, kind in ((oah, po.FE_DGQ), (pah, pdl.FE_DGQ)):
            for gr in groups:
                ah.define_agglomerate(gr)
            ah.initialize_fe_values(2)
            ah.distribute_agglomerated_dofs(kind, 1)
        for p in range(oah.n_polytopes):
            assert pah.n_faces(p) == oah.n_faces(p)
            for f in range(oah.n_faces(p)):
                assert pah.neighbor(p, f) == oah.neighbor(p, f) and pah.interface(p, f) == oah.interface(p, f)
                assert pah.neighbor_of_agglomerated_neighbor(p, f) == oah.neighbor_of_agglomerated_neighbor(p, f)
        rp, cols = pah.create_agglomeration_sparsity_pattern()
        orp, ocols = oah.create_agglomeration_sparsity_pattern()
        np.testing.assert_array_equal(rp, orp)
        np.testing.assert_array_equal(cols, ocols)
        A = po.assemble_dg_matrix(oah, penalty_constant=10.0, h_rule=po.H_MAX_INVERSE_DIAMETER, with_boundary=False).scipy()
        one = np.ones(oah.n_dofs)
        assert abs(A - A.T).max() < 1e-12 and abs(one @ (A @ one)) < 1e-10
        assert pah.flatten().n_polytopes == oah.n_polytopes
