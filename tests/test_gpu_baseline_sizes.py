"""GPU parity at the FULL sizes BASELINE.json names, through the C ABI against the CPU oracle.

 * A (2-D 256^2 cells -> 256 polygons, DGQ1) and B (64^3 -> 512 polyhedra, DGQ2): the whole matrix per block entry
   and the vmult per output, `blocks` and `metis` agglomeration shapes.
 * C (128^3 -> 32 768 polyhedra, DGQ3) and D/8 (128^3 -> 32 768 polyhedra, DGQ2 + reaction, the per-GPU share of D):
   the scalar oracle cannot assemble 1e9 matrix entries in seconds, so the COMPLETE block rows of every 64th polytope
   (oracle `assemble_block_rows`: every interface evaluated from its visiting side exactly as
   include/poly_utils.h:2086-2132 does) are compared entry by entry with the same rows of the GPU matrix, and the
   vmult outputs of those rows with the oracle rows times x.
 * METIS shapes (the reference's main agglomeration shape, examples/poisson.cc) in 2-D and 3-D, and an 8-rank
   emulation of the METIS distribution of polytopes over GPUs.

Tolerance (north_star): 1e-12 relative per matrix block entry (to the block's largest entry) and per vmult output
(to max|y|)."""
import numpy as np
import pytest

import pd_scenarios as sc
from oracle import pyoracle as po
from pd_helpers import assert_blocks_close, groups_for, oracle_handler, product_handler, src_vector

pytestmark = pytest.mark.gpu
TOL = 1e-12


def fast_block_groups(dim, n, b):
    """b^dim blocks of the Morton-ordered n^dim grid without Python loops over cells (tools/pd_workloads.py)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from pd_workloads import morton_block_groups

    return morton_block_groups(dim, n, b)


def handlers(dim, n, groups, p, nq):
    import polydeal_b200 as pdl

    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    oah = po.AgglomerationHandler(ogrid)
    grid = pdl.Grid.hyper_cube(dim, 0.0, 1.0, n.bit_length() - 1)
    pah = pdl.AgglomerationHandler(grid)
    for g in groups:
        oah.define_agglomerate(g)
        pah.define_agglomerate(g)
    for ah, fe in ((oah, po.FE_DGQ), (pah, pdl.FE_DGQ)):
        ah.initialize_fe_values(nq)
        ah.distribute_agglomerated_dofs(fe, p)
    return oah, pah


def groups_of_shape(shape, dim, n):
    if shape.startswith("blocks"):
        return fast_block_groups(dim, n, int(shape[6:]))
    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    return groups_for(shape, dim, n, ogrid)


# ----------------------------------------------------------------------------------
# A and B at full size: the whole matrix per block entry + vmult per output
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name,dim,n,shape,p,nq", [
    ("A-blocks", 2, 256, "blocks16", 1, 2),
    ("A-metis", 2, 256, "metis256", 1, 2),
    ("B-blocks", 3, 64, "blocks8", 2, 3),
    ("B-metis", 3, 64, "metis512", 2, 3),
])
@pytest.mark.parametrize("kernels", ["tensor", "dmma"])
def test_full_matrix_at_baseline_size(name, dim, n, shape, p, nq, kernels, monkeypatch):
    import torch

    import polydeal_b200 as pdl

    if kernels == "dmma":
        monkeypatch.setenv("PD_ASSEMBLE_KERNELS", "generic")

    oah, pah = handlers(dim, n, groups_of_shape(shape, dim, n), p, nq)
    ref = po.assemble_dg_matrix(oah, degree=p, n_threads=16)
    op = pdl.assemble_dg_matrix(pah)
    assert op.assembly_path == kernels
    rp, cols = op.pattern()
    orp, ocols, ovals = ref.csr()
    np.testing.assert_array_equal(rp, orp)  # sparsity bit exact
    np.testing.assert_array_equal(cols, ocols)
    for k in range(oah.n_polytopes):  # DoF numbering bit exact
        np.testing.assert_array_equal(pah.get_dof_indices(k), oah.get_dof_indices(k))
    vals = op.values()
    assert np.isfinite(vals).all()
    worst = assert_blocks_close(vals, ovals, oah.n_dofs_per_cell, rp, TOL)
    x = src_vector(op.m())
    yref = ref.vmult(x, n_threads=8)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    op.vmult(yd, xd)
    op.synchronize()
    scale = np.abs(yref).max()
    err = np.abs(yd.cpu().numpy() - yref).max() / scale
    assert err <= TOL
    # the polytopal matrix-free apply of the same operator
    op.set_operator(pdl.ASSEMBLE_ALL, 1.0, 0.0)
    op.vmult(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    err_mf = np.abs(yd.cpu().numpy() - yref).max() / scale
    assert err_mf <= TOL
    print(f"{name}: {op.m()} DoFs, worst block-relative matrix error {worst:.2e}, vmult {err:.2e}, matrix-free {err_mf:.2e}")


# ----------------------------------------------------------------------------------
# C and D/8 at full size: complete block rows of every 64th polytope + their vmult outputs
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,b,p,nq,kw", [
    ("C", 128, 4, 3, 4, {}),
    ("D/8", 128, 4, 2, 3, dict(penalty_constant=40.0, mass_coeff=0.5)),
])
@pytest.mark.parametrize("kernels", ["tensor", "dmma"])
def test_sampled_block_rows_at_baseline_size(name, n, b, p, nq, kw, kernels, monkeypatch):
    import torch

    import polydeal_b200 as pdl

    if kernels == "dmma":
        monkeypatch.setenv("PD_ASSEMBLE_KERNELS", "generic")

    dim = 3
    oah, pah = handlers(dim, n, fast_block_groups(dim, n, b), p, nq)
    nd = oah.n_dofs_per_cell
    polys = np.arange(7, oah.n_polytopes, 64)  # 512 polytopes: interior, face, edge and corner ones
    okw = dict(kw)
    okw.setdefault("penalty_constant", None)
    ptr, bcol, rows, seconds = po.assemble_block_rows(oah, polys, degree=p, n_threads=16, **okw)
    pkw = dict(kw)
    pkw.setdefault("penalty_constant", -1.0)
    op = pdl.assemble_dg_matrix(pah, **pkw)
    assert op.m() == oah.n_dofs and op.assembly_path == kernels
    dvals = torch.as_tensor(_DevView(op.values_device_ptr(), op.nnz), device="cuda")
    # pattern of the sampled rows from the host mirror (bit exact vs the oracle's block columns)
    d = op.desc
    brow = np.ctypeslib.as_array(d.brow_ptr, (d.n_block_rows + 1,))
    bcols = np.ctypeslib.as_array(d.bcol_idx, (int(brow[-1]),))
    x = src_vector(op.m())
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    op.vmult(yd, xd)
    op.synchronize()
    y = yd.cpu().numpy()
    op.set_operator(pdl.ASSEMBLE_ALL, 1.0, kw.get("mass_coeff", 0.0))
    op.vmult(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    ymf = yd.cpu().numpy()
    worst, worst_y, worst_mf, yscale = 0.0, 0.0, 0.0, np.abs(y).max()
    for s, poly in enumerate(polys):
        np.testing.assert_array_equal(pah.get_dof_indices(int(poly)), oah.get_dof_indices(int(poly)))
        blk = int(oah.get_dof_indices(int(poly))[0]) // nd
        cols_s = bcol[ptr[s]:ptr[s + 1]]
        np.testing.assert_array_equal(bcols[brow[blk]:brow[blk + 1]], cols_s)  # sparsity of the row bit exact
        nb = len(cols_s)
        got = dvals[brow[blk] * nd * nd:brow[blk + 1] * nd * nd].cpu().numpy().reshape(nd, nb, nd)
        ref = rows[s].reshape(nd, nb, nd)
        scale = np.abs(ref).max(axis=(0, 2))
        worst = max(worst, (np.abs(got - ref).max(axis=(0, 2)) / scale).max())
        xs = x.reshape(-1, nd)[cols_s].ravel()
        yref = rows[s] @ xs
        worst_y = max(worst_y, np.abs(y[blk * nd:(blk + 1) * nd] - yref).max() / yscale)
        worst_mf = max(worst_mf, np.abs(ymf[blk * nd:(blk + 1) * nd] - yref).max() / yscale)
    print(f"{name}: {op.m()} DoFs, {len(polys)} sampled block rows ({seconds:.1f} s of oracle), worst block-relative "
          f"matrix error {worst:.2e}, vmult {worst_y:.2e}, matrix-free {worst_mf:.2e}")
    assert worst <= TOL and worst_y <= TOL and worst_mf <= TOL


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


# ----------------------------------------------------------------------------------
# METIS shapes (small enough for the whole matrix) and the METIS distribution over 8 emulated ranks
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n,shape,p,nq,kw", [
    (2, 32, "metis24", 1, 2, {}),
    (2, 16, "metis9", 3, 4, {}),
    (3, 8, "metis11", 2, 3, {}),
    (3, 8, "metis6", 3, 4, dict(mass_coeff=0.5, penalty_constant=90.0)),
    (3, 16, "metis64", 1, 2, dict(visit_rule=1)),
])
def test_metis_shapes_match_oracle(dim, n, shape, p, nq, kw):
    import polydeal_b200 as pdl

    oah, pah = handlers(dim, n, groups_of_shape(shape, dim, n), p, nq)
    okw = dict(kw)
    okw.setdefault("penalty_constant", None)
    ref = po.assemble_dg_matrix(oah, degree=p, n_threads=8, **okw)
    pkw = dict(kw)
    pkw.setdefault("penalty_constant", -1.0)
    op = pdl.assemble_dg_matrix(pah, **pkw)
    rp, cols = op.pattern()
    orp, ocols, ovals = ref.csr()
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(cols, ocols)
    assert_blocks_close(op.values(), ovals, oah.n_dofs_per_cell, rp, TOL)


@pytest.mark.parametrize("dim,n,shape,p,world", [(3, 16, "blocks2", 2, 8), (2, 64, "metis96", 1, 8)])
def test_eight_ranks_metis_distribution_emulated(dim, n, shape, p, world):
    """polytopes distributed over 8 GPUs by METIS on the polytope adjacency graph; every rank's local
    assembly (owner computes rows, cut interfaces from ghost bbox + DoF block) and vmult against the serial oracle."""
    import torch

    import polydeal_b200 as pdl
    from polydeal_b200 import distributed as pdd

    oah, pah = handlers(dim, n, groups_of_shape(shape, dim, n), p, p + 1)
    A = po.assemble_dg_matrix(oah, degree=p, n_threads=8).scipy().tocsr()
    nd = oah.n_dofs_per_cell
    x = src_vector(A.shape[0])
    y = A @ x
    owner = pdd.partition_by_metis(pah, world)
    assert sorted(set(owner.tolist())) == list(range(world))
    seen = 0
    for rank in range(world):
        part = pdd.LocalPart(pah, owner, rank)
        op = pdl.SIPOperator(part.desc, keepalive=(pah, part))
        op.assemble()
        rows = part.owned_global_dofs()
        cols = np.concatenate([rows, part.ghost_global_dofs()])
        ref = A[rows][:, cols].tocsr()
        ref.sort_indices()
        got = op.scipy()
        got.sort_indices()
        np.testing.assert_array_equal(got.indptr, ref.indptr)
        np.testing.assert_array_equal(got.indices, ref.indices)
        rp, _ = op.pattern()
        assert_blocks_close(op.values(), ref.data, nd, rp, TOL)
        xd = torch.from_numpy(x[cols]).cuda()
        yd = torch.empty(len(rows), dtype=torch.float64, device="cuda")
        op.vmult_ptr(yd.data_ptr(), xd.data_ptr())
        op.synchronize()
        assert np.abs(yd.cpu().numpy() - y[rows]).max() <= TOL * np.abs(y).max()
        op.set_operator()
        op.vmult_ptr(yd.data_ptr(), xd.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)
        op.synchronize()
        assert np.abs(yd.cpu().numpy() - y[rows]).max() <= TOL * np.abs(y).max()
        seen += len(rows)
    assert seen == A.shape[0]
