"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs, plus the reference tests' invariants evaluated on the GPU result.

Tolerance (north_star): 1e-12 relative in fp64, applied per matrix block entry relative
to the block's largest entry and per vmult output relative to max|y|."""
import os

import numpy as np
import pytest

import pd_scenarios as sc
from oracle import pyoracle as po
from pd_helpers import assert_blocks_close, groups_for, oracle_handler, product_handler, src_vector

pytestmark = pytest.mark.gpu

TOL = 1e-12


def gpu():
    import polydeal_b200 as pdl

    return pdl


def both(dim, n, shape, p, nq=None, order=0, distort=None, lo=0.0, hi=1.0, seed=1):
    nq = nq if nq is not None else p + 1
    ogrid = po.Grid(dim, n, lo, hi, order)
    groups = groups_for(shape, dim, n, ogrid, seed)
    _, oah = oracle_handler(dim, n, groups, p, nq, lo=lo, hi=hi, order=order, distort=distort)
    _, pah = product_handler(oah.grid, groups, p, nq)
    return oah, pah


# ----------------------------------------------------------------------------------
# row 1: agglomerated quadrature (source/agglomeration_handler.cc:622-707, 1139-1165)
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n,shape,nq,distort", [
    (2, 8, "random7", 2, None), (2, 8, "random7", 3, (0.25, 5)),
    (3, 4, "random6", 2, None), (3, 4, "blocks2", 3, (0.2, 9)),
])
def test_quadrature_matches_oracle(dim, n, shape, nq, distort):
    pdl = gpu()
    oah, pah = both(dim, n, shape, 1, nq=nq, distort=distort)
    op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    op.build_quadrature()
    Q = sum(len(oah.get_agglomerate(p)) for p in range(oah.n_polytopes)) * nq**dim
    x = op.copy_array("vol_qpt").reshape(dim, Q)
    w = op.copy_array("vol_jxw")
    k = 0
    for p in range(oah.n_polytopes):
        fev = oah.reinit(p)
        np.testing.assert_allclose(x[:, k:k + fev.n_q].T, fev.points, rtol=0, atol=1e-15)
        np.testing.assert_allclose(w[k:k + fev.n_q], fev.JxW, rtol=1e-14, atol=0)
        k += fev.n_q
    assert k == Q
    # faces: the flattened work list, visited from A
    d = op.desc
    A = np.ctypeslib.as_array(d.iface_polyA, (d.n_ifaces,))
    B = np.ctypeslib.as_array(d.iface_polyB, (d.n_ifaces,))
    Qf = int(np.ctypeslib.as_array(d.iface_sub_ptr, (d.n_ifaces + 1,))[-1]) * nq ** (dim - 1)
    fx = op.copy_array("face_qpt").reshape(dim, Qf)
    fn = op.copy_array("face_normal").reshape(dim, Qf)
    fw = op.copy_array("face_jxw")
    k = 0
    for a, b in zip(A, B):
        f = next(f for f in range(oah.n_faces(a)) if oah.neighbor(a, f) == b)
        ff = oah.reinit(a, f)
        np.testing.assert_allclose(fx[:, k:k + ff.n_q].T, ff.points, rtol=0, atol=1e-15)
        np.testing.assert_allclose(fn[:, k:k + ff.n_q].T, ff.normals, rtol=0, atol=1e-15)
        np.testing.assert_allclose(fw[k:k + ff.n_q], ff.JxW, rtol=1e-14, atol=0)
        k += ff.n_q
    assert k == Qf


# ----------------------------------------------------------------------------------
# rows 2-9: assembled matrix, per block entry
# ----------------------------------------------------------------------------------
CASES = [
    # dim, n, shape, p, nq, distort, kwargs
    (2, 16, "blocks4", 1, 2, None, {}),                       # config A shape (reduced)
    (2, 16, "random13", 1, 3, None, {}),
    (2, 8, "random5", 2, 3, (0.2, 3), {}),
    (2, 8, "blocks2", 3, 4, None, {}),
    (2, 8, "random4", 4, 5, None, {}),
    (3, 8, "blocks4", 1, 2, None, {}),
    (3, 8, "blocks4", 2, 3, None, {}),                        # config B / D shape (reduced)
    (3, 8, "random12", 2, 3, None, {}),
    (3, 4, "random5", 2, 3, (0.15, 4), {}),
    (3, 8, "blocks4", 3, 4, None, {}),                        # config C shape (reduced)
    (3, 4, "random3", 3, 4, None, {}),
    (3, 4, "singletons", 2, 3, None, dict(penalty_constant=6.0, h_rule=3)),  # fine-mesh LaplaceOperatorDG penalty
    (3, 8, "blocks4", 2, 3, None, dict(mass_coeff=0.5, penalty_constant=40.0)),  # config D: reaction c=0.5, C=10p^2
    (2, 8, "random6", 1, 3, None, dict(penalty_constant=10.0, h_rule=1, with_boundary=False)),  # sanity-check rule
    (3, 4, "random4", 1, 2, None, dict(stiffness_coeff=1e-4, mass_coeff=1.5e4, with_boundary=False)),  # monodomain f M + sigma K
    (2, 8, "random6", 2, 3, None, dict(visit_rule=1)),        # examples/poisson.cc: visit by index()
    (3, 4, "random4", 2, 3, None, dict(mass_coeff=-0.3)),     # negative coefficient: signed-weight kernel variant
    (2, 8, "random5", 1, 2, None, dict(mass_coeff=-2.0)),
]


@pytest.mark.parametrize("kernels", ["default", "generic"])
@pytest.mark.parametrize("dim,n,shape,p,nq,distort,kw", CASES)
def test_assembled_matrix_matches_oracle(dim, n, shape, p, nq, distort, kw, kernels, monkeypatch):
    """Both kernel families against the oracle: by default axis-aligned meshes take the tensor path
    (pd_cartesian.cu: per-sub-cell sum factorisation) and distorted ones the DMMA kernels on the agglomerated
    quadrature (pd_assemble.cu); PD_ASSEMBLE_KERNELS=generic runs the DMMA kernels everywhere."""
    pdl = gpu()
    if kernels == "generic":
        if distort is not None:
            pytest.skip("distorted meshes take the DMMA kernels by default")
        monkeypatch.setenv("PD_ASSEMBLE_KERNELS", "generic")
    oah, pah = both(dim, n, shape, p, nq=nq, distort=distort)
    okw = dict(kw)
    okw.setdefault("penalty_constant", None)
    ref = po.assemble_dg_matrix(oah, degree=p, n_threads=4, **okw)
    pkw = dict(kw)
    pkw.setdefault("penalty_constant", -1.0)
    op = pdl.assemble_dg_matrix(pah, **pkw)
    assert op.assembly_path == ("tensor" if distort is None and kernels == "default" else "dmma")
    rp, cols = op.pattern()
    orp, ocols, ovals = ref.csr()
    np.testing.assert_array_equal(rp, orp)      # sparsity bit exact
    np.testing.assert_array_equal(cols, ocols)
    vals = op.values()
    assert np.isfinite(vals).all()
    assert_blocks_close(vals, ovals, oah.n_dofs_per_cell, rp, TOL)
    # vmult with the assembled matrix, device vectors and host vectors
    import torch

    x = src_vector(op.m())
    yref = ref.vmult(x)
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    op.vmult(yd, xd)
    torch.cuda.synchronize()
    op.synchronize()
    scale = np.abs(yref).max()
    assert np.abs(yd.cpu().numpy() - yref).max() <= TOL * scale
    yh = np.empty_like(x)
    op.vmult(yh, x)
    assert np.abs(yh - yref).max() <= TOL * scale
    op.vmult_add(yd, xd)
    op.synchronize()
    assert np.abs(yd.cpu().numpy() - 2 * yref).max() <= 2 * TOL * scale
    dinv = torch.empty_like(xd)
    op.get_matrix_diagonal_inverse(dinv)
    op.synchronize()
    diag = ref.scipy().diagonal()
    np.testing.assert_allclose(dinv.cpu().numpy(), np.where(np.abs(diag) > 1e-10, 1.0 / diag, diag), rtol=1e-11)


@pytest.mark.parametrize("kernels", ["default", "generic"])
def test_assemble_flags_split_the_matrix(kernels, monkeypatch):
    """volume + boundary + interior parts add up to the full matrix."""
    pdl = gpu()
    if kernels == "generic":
        monkeypatch.setenv("PD_ASSEMBLE_KERNELS", "generic")
    from polydeal_b200 import ASSEMBLE_BOUNDARY, ASSEMBLE_INTERIOR, ASSEMBLE_VOLUME

    oah, pah = both(3, 4, "random5", 2, nq=3)
    op = pdl.assemble_dg_matrix(pah)
    full = op.values().copy()
    parts = np.zeros_like(full)
    for fl in (ASSEMBLE_VOLUME, ASSEMBLE_BOUNDARY, ASSEMBLE_INTERIOR):
        op.assemble(fl)
        parts += op.values()
    assert np.abs(parts - full).max() <= 1e-13 * np.abs(full).max()


# ----------------------------------------------------------------------------------
# the reference tests' own invariants, evaluated on the GPU result
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [2, 3])
def test_minimal_sip_poisson_on_gpu(dim):
    """test/polydeal/minimal_SIP_Poisson.cc: agglomerated == standard matrix to 1e-13."""
    pdl = gpu()
    kw = dict(penalty_constant=20.0, h_rule=pdl.H_CONSTANT, h_const=1.0, visit_rule=pdl.VISIT_BY_INDEX)
    if dim == 2:
        ga, gs = sc.blocks_2x2_of_4x4(), [[c] for c in range(4)]
        na, ns = 4, 2
    else:
        ga, gs = [list(range(8))], [[0]]
        na, ns = 2, 1
    mats = []
    for n, groups in ((na, ga), (ns, gs)):
        ogrid = po.Grid(dim, n, -1.0, 1.0, 0)
        _, pah = product_handler(ogrid, groups, 1, 3)
        mats.append(pdl.assemble_dg_matrix(pah, **kw).scipy().toarray())
    assert np.abs(mats[0] - mats[1]).max() < 1e-13


def interpolate(pah, fun, p):
    nodes = po.gauss_lobatto_nodes(p + 1)
    u = np.zeros(pah.n_dofs)
    for k in range(pah.n_polytopes):
        lo, hi = pah.bbox(k)
        for i, dof in enumerate(pah.get_dof_indices(k)):
            idx = [(i // (p + 1) ** d) % (p + 1) for d in range(pah.dim)]
            u[dof] = fun(lo + nodes[idx] * (hi - lo))
    return u


@pytest.mark.parametrize("n_parts", [50, 120])
def test_poisson_sanity_check_on_gpu(n_parts, goldens):
    """test/polydeal/poisson_sanity_check_01.cc: x'Ax = 1, (x+y)'A(x+y) = 2, 1'A1 ~ 1e-14."""
    pdl = gpu()
    ogrid = po.Grid(2, 64, 0.0, 1.0, 0)
    groups = groups_for(f"random{n_parts}", 2, 64, ogrid, seed=n_parts)
    _, pah = product_handler(ogrid, groups, 1, 3)
    A = pdl.assemble_dg_matrix(pah, penalty_constant=10.0, h_rule=pdl.H_MAX_INVERSE_DIAMETER, with_boundary=False).scipy()
    ux, uxy, one = interpolate(pah, lambda x: x[0], 1), interpolate(pah, lambda x: x[0] + x[1], 1), np.ones(pah.n_dofs)
    g = goldens["poisson_sanity_check_01"]
    assert ux @ (A @ ux) == pytest.approx(g["x"][0], abs=1e-11)
    assert uxy @ (A @ uxy) == pytest.approx(g["xplusy"][0], abs=1e-11)
    assert abs(one @ (A @ one)) < 1e-11
    assert abs(A - A.T).max() < 1e-12


# ----------------------------------------------------------------------------------
# full BASELINE sizes through size-independent properties
# ----------------------------------------------------------------------------------
def test_config_b_full_size_properties():
    """Config B (64^3 hexes, 512 polyhedra of 8^3 cells, DGQ2, QGauss(3)): too big for the
    scalar oracle in seconds, so: symmetry, constants in the kernel of the boundary-free
    operator, energy of u = x equals |Omega| = 1, and translation invariance (all interior
    polytopes of the uniform `blocks` shape carry identical diagonal blocks)."""
    pdl = gpu()
    import torch

    grid = pdl.Grid.hyper_cube(3, 0.0, 1.0, 6)
    ah = pdl.AgglomerationHandler(grid)
    for g in sc.block_partition(3, 64, 8):
        ah.define_agglomerate(g)
    ah.initialize_fe_values(3)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ, 2)
    op = pdl.assemble_dg_matrix(ah, with_boundary=False)
    A = op.scipy()
    assert A.shape == (13824, 13824)
    scale = abs(A).max()
    assert abs(A - A.T).max() <= 1e-12 * scale
    one = np.ones(op.m())
    assert np.abs(A @ one).max() <= 1e-10 * scale
    ux = interpolate(ah, lambda x: x[0], 2)
    assert ux @ (A @ ux) == pytest.approx(1.0, abs=1e-10)
    # interior polytopes: identical diagonal blocks
    n = 27
    D = A.toarray().reshape(512, n, 512, n)
    blocks = sc.block_partition(3, 64, 8)
    interior = [p for p in range(512) if all(not ah.at_boundary(p, f) for f in range(ah.n_faces(p)))]
    assert len(interior) == 216
    b0 = ah.get_dof_indices(interior[0])[0] // n
    ref_block = D[b0, :, b0, :]
    for p in interior[1:]:
        b = ah.get_dof_indices(p)[0] // n
        assert np.abs(D[b, :, b, :] - ref_block).max() <= 1e-11 * np.abs(ref_block).max()
    del blocks
    # vmult agrees with the assembled matrix applied on the host
    x = src_vector(op.m())
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    op.vmult(yd, xd)
    op.synchronize()
    y = A @ x
    assert np.abs(yd.cpu().numpy() - y).max() <= 1e-12 * np.abs(y).max()


def test_edge_cases():
    """single polytope covering the mesh (no interior faces), singleton polytopes only
    (every face a 1-sub-face interface), and a two-polytope split."""
    pdl = gpu()
    for dim, n, shape, p in [(2, 4, "blocks4", 2), (3, 2, "singletons", 1), (2, 4, "random2", 1)]:
        oah, pah = both(dim, n, shape, p)
        ref = po.assemble_dg_matrix(oah, degree=p)
        op = pdl.assemble_dg_matrix(pah)
        rp, _ = op.pattern()
        assert_blocks_close(op.values(), ref.values(), oah.n_dofs_per_cell, rp, TOL)


def test_unsupported_degree_fails_loudly():
    pdl = gpu()
    ogrid = po.Grid(3, 2, 0.0, 1.0, 0)
    _, pah = product_handler(ogrid, [[c] for c in range(8)], 4, 5)
    with pytest.raises(pdl.PolydealError, match="no sm_100a kernel"):
        pdl.assemble_dg_matrix(pah)


# ----------------------------------------------------------------------------------
# rows 10-11: matrix-free sum-factorised SIP on the fine mesh
# (LaplaceOperatorDG include/utils.h:819-925, MonodomainOperatorDG :1565-1659).
# Checker: the oracle's matrix of the same form on singleton polytopes (the emulation
# the reference itself documents at examples/monodomain_DG3D.cc:1470-1498) times x.
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n,p,order,hi,kw", [
    (2, (8, 8), 1, 0, 1.0, {}),
    (2, (6, 5), 2, 1, (1.0, 0.7), {}),
    (2, (4, 4), 3, 0, 1.0, {}),
    (2, (4, 3), 4, 1, (2.0, 1.0), {}),
    (3, (4, 4, 4), 1, 0, 1.0, {}),
    (3, (4, 4, 4), 2, 0, 1.0, {}),
    (3, (3, 4, 5), 2, 1, (1.0, 0.8, 1.3), {}),
    (3, (2, 2, 2), 3, 0, 1.0, {}),
    (3, (8, 8, 8), 2, 0, 1.0, {}),                  # tiled kernel: eight 4x4x4 Morton tiles with halos
    (3, (6, 5, 7), 1, 1, (1.0, 0.9, 1.2), {}),      # ... lexicographic order: pencils with large halos, ragged last tile
    (2, (16, 16), 3, 0, 1.0, {}),
    (2, (20, 13), 2, 1, (1.0, 0.6), dict(with_boundary=False, stiffness_coeff=0.3, mass_coeff=2.0)),
    (3, (4, 4, 4), 1, 0, 1.0, dict(with_boundary=False, stiffness_coeff=1e-4, mass_coeff=1.5e4)),   # monodomain, BDF2
    (3, (3, 3, 3), 2, 1, 1.0, dict(with_boundary=False, stiffness_coeff=1e-4, mass_coeff=1.0e4)),  # monodomain, BDF1
])
@pytest.mark.parametrize("kernel", ["default", "tile"])
def test_fine_mesh_matrix_free_vmult(dim, n, p, order, hi, kw, kernel, monkeypatch):
    pdl = gpu()
    import torch

    if kernel == "tile":  # the tiled kernel wherever it exists (by default only where it is the faster one)
        if dim == 3 and p == 3:
            pytest.skip("3-D DGQ3 has the line-per-thread kernel only")
        monkeypatch.setenv("PD_FINE_KERNEL", "tile")

    ogrid = po.Grid(dim, n, 0.0, hi, order)
    groups = [[c] for c in range(ogrid.n_cells)]
    oah = po.AgglomerationHandler(ogrid)
    for g in groups:
        oah.define_agglomerate(g)
    oah.initialize_fe_values(p + 1)
    oah.distribute_agglomerated_dofs(po.FE_DGQ, p)
    _, pah = product_handler(ogrid, groups, p, p + 1)
    C = max(p, 1) * (p + 1.0)  # include/utils.h:866
    ref = po.assemble_dg_matrix(oah, penalty_constant=C, h_rule=po.H_NORMAL_EXTENT, n_threads=4, **kw)
    op = pdl.SIPOperator(pah.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
    assert op.matrix_free_available
    flags = pdl.ASSEMBLE_ALL if kw.get("with_boundary", True) else (pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR)
    op.set_operator(flags, kw.get("stiffness_coeff", 1.0), kw.get("mass_coeff", 0.0))
    x = src_vector(op.m())
    yref = ref.vmult(x)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros_like(xd)
    op.vmult(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    scale = np.abs(yref).max()
    assert np.abs(yd.cpu().numpy() - yref).max() <= TOL * scale
    op.vmult_add(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    assert np.abs(yd.cpu().numpy() - 2 * yref).max() <= 2 * TOL * scale
    # and it is the same operator as the assembled one on the GPU
    op.assemble(flags, kw.get("stiffness_coeff", 1.0), kw.get("mass_coeff", 0.0))
    yb = torch.empty_like(xd)
    op.vmult(yb, xd, mode=pdl.VMULT_BLOCK_CSR)
    op.synchronize()
    assert np.abs(yb.cpu().numpy() - yref).max() <= TOL * scale


@pytest.mark.parametrize("dim,n,p,order", [(3, (16, 16, 16), 2, 0), (3, (12, 10, 9), 1, 1), (2, (40, 33), 4, 1), (2, (32, 32), 1, 0)])
def test_fine_mesh_tiled_kernel_equals_line_kernel(dim, n, p, order, monkeypatch):
    """The two kernels behind PD_VMULT_MATRIX_FREE on a fine Cartesian mesh -- k_fine_tile (one thread per cell,
    coefficients staged in shared memory; the default) and k_fine_sip (one thread per line; PD_FINE_KERNEL=line,
    and the only one for 3-D DGQ3) -- on meshes of many tiles, Laplace and monodomain coefficients, vmult and vmult_add."""
    pdl = gpu()
    import torch

    ogrid = po.Grid(dim, n, 0.0, 1.0, order)
    groups = [[c] for c in range(ogrid.n_cells)]
    C = max(p, 1) * (p + 1.0)
    ops = {}
    for kernel in ("tile", "line"):
        monkeypatch.setenv("PD_FINE_KERNEL", kernel)
        _, pah = product_handler(ogrid, groups, p, p + 1)
        ops[kernel] = pdl.SIPOperator(pah.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
        assert ops[kernel].matrix_free_available
    x = torch.from_numpy(src_vector(ops["tile"].m())).cuda()
    for flags, sc, mc in [(pdl.ASSEMBLE_ALL, 1.0, 0.0), (pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR, 1e-4, 1.5e4)]:
        y = {}
        for kernel, op in ops.items():
            op.set_operator(flags, sc, mc)
            y[kernel] = torch.full_like(x, 0.25)
            op.vmult(y[kernel], x, mode=pdl.VMULT_MATRIX_FREE)
            op.vmult_add(y[kernel], x, mode=pdl.VMULT_MATRIX_FREE)
            op.synchronize()
        scale = y["line"].abs().max().item()
        assert scale > 0 and (y["tile"] - y["line"]).abs().max().item() <= TOL * scale


@pytest.mark.parametrize("dim,n,p,hi", [
    (3, (8, 8, 8), 2, 1.0),             # eight tiles, every one with boundary faces
    (3, (16, 16, 16), 2, (1.0, 0.5, 2.0)),  # 64 tiles (more than a CTA's ring of stages), anisotropic cells
    (3, (4, 4, 4), 2, (0.5, 1.0, 1.0)),  # one tile without halo
    (2, (32, 32), 2, 1.0),
    (2, (32, 32), 4, (1.0, 3.0)),
])
def test_fine_mesh_pipelined_kernel(dim, n, p, hi, monkeypatch):
    """k_fine_stream (uniform meshes: persistent CTAs, producer warp + ring of shared-memory stages, results
    by bulk store / bulk reduce) is the kernel that runs on these meshes, and it agrees with the oracle's
    matrix of the same form (include/utils.h:819-925, 1565-1659) per output, with the line kernel, for
    Laplace and monodomain coefficients, vmult and vmult_add, whatever the number of stages."""
    pdl = gpu()
    import torch

    ogrid = po.Grid(dim, n, 0.0, hi, 0)
    groups = [[c] for c in range(ogrid.n_cells)]
    oah = po.AgglomerationHandler(ogrid)
    for g in groups:
        oah.define_agglomerate(g)
    oah.initialize_fe_values(p + 1)
    oah.distribute_agglomerated_dofs(po.FE_DGQ, p)
    C = max(p, 1) * (p + 1.0)
    ops = {}
    for kernel in ("stream", "line"):
        monkeypatch.setenv("PD_FINE_KERNEL", kernel)
        _, pah = product_handler(ogrid, groups, p, p + 1)
        ops[kernel] = pdl.SIPOperator(pah.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
    x = torch.from_numpy(src_vector(ops["stream"].m())).cuda()
    small = ogrid.n_cells <= 1024
    for flags, kw in [(pdl.ASSEMBLE_ALL, {}),
                      (pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR, dict(with_boundary=False, stiffness_coeff=1e-4, mass_coeff=1.5e4)),
                      (pdl.ASSEMBLE_INTERIOR | pdl.ASSEMBLE_BOUNDARY, None)]:
        y = {}
        for kernel, op in ops.items():
            sc, mc = (kw or {}).get("stiffness_coeff", 1.0), (kw or {}).get("mass_coeff", 0.0)
            op.set_operator(flags, sc, mc)
            y[kernel] = torch.full_like(x, 0.25)
            op.vmult(y[kernel], x, mode=pdl.VMULT_MATRIX_FREE)
            op.vmult_add(y[kernel], x, mode=pdl.VMULT_MATRIX_FREE)
            op.synchronize()
        assert ops["stream"].fine_kernel_last == 3 and ops["line"].fine_kernel_last == 1
        scale = y["line"].abs().max().item()
        assert scale > 0 and (y["stream"] - y["line"]).abs().max().item() <= TOL * scale
        if small and kw is not None:
            ref = po.assemble_dg_matrix(oah, penalty_constant=C, h_rule=po.H_NORMAL_EXTENT, n_threads=4, **kw)
            yref = 2 * ref.vmult(x.cpu().numpy())
            assert np.abs(y["stream"].cpu().numpy() - yref).max() <= 2 * TOL * np.abs(yref).max()


@pytest.mark.parametrize("dim,n,shape,p,nq,distort", [
    (2, 8, "random5", 1, 2, None),
    (2, 8, "blocks4", 2, 4, (0.2, 3)),
    (2, 8, "random4", 4, 5, None),
    (3, 4, "random6", 1, 3, None),
    (3, 4, "blocks2", 2, 3, (0.15, 8)),
    (3, 4, "random3", 3, 4, None),
])
def test_rhs_and_error_functionals(dim, n, shape, p, nq, distort):
    """pd_assemble_rhs (examples/poisson.cc:745-761 + Dirichlet terms of diffusion_reaction.cc:550-556)
    and pd_error_norms (PolyUtils::compute_global_error, include/poly_utils.h:1647-1750) against
    the oracle's reinit() tables."""
    pdl = gpu()
    import torch

    oah, pah = both(dim, n, shape, p, nq=nq, distort=distort)
    C = 10.0 * (p + dim) * (p + 1)
    op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    f = lambda x: np.sin(1.3 * x[..., 0] + 0.2) * np.cos(0.7 * x[..., 1]) + (x[..., 2] ** 2 if dim == 3 else 0.0)
    g = lambda x: 1.0 + x[..., 0] - 0.5 * x[..., 1] ** 2
    gradg = lambda x: np.stack([np.ones_like(x[..., 0]), -x[..., 1]] + ([np.zeros_like(x[..., 0])] if dim == 3 else []), axis=-1)
    # ---- oracle side
    N = oah.n_dofs
    b_ref = np.zeros(N)
    uh = src_vector(N) * 0.1
    l2_ref = h1_ref = 0.0
    for k in range(oah.n_polytopes):
        fev = oah.reinit(k)
        dofs = oah.get_dof_indices(k)
        b_ref[dofs] += fev.values @ (f(fev.points) * fev.JxW)
        uq = uh[dofs] @ fev.values
        gq = np.einsum("i,iqd->qd", uh[dofs], fev.grads)
        l2_ref += np.sum((uq - g(fev.points)) ** 2 * fev.JxW)
        h1_ref += np.sum(np.sum((gq - gradg(fev.points)) ** 2, axis=1) * fev.JxW)
        for fc in range(oah.n_faces(k)):
            if oah.at_boundary(k, fc):
                fv = oah.reinit(k, fc)
                sig = C / oah.diameter(k)
                gn = np.einsum("iqd,qd->iq", fv.grads, fv.normals)
                b_ref[dofs] += 0.7 * ((sig * fv.values - gn) @ (g(fv.points) * fv.JxW))
    # ---- device side: data evaluated at the device quadrature points
    q = op.quadrature()
    vx = q["vol_x"].T.cpu().numpy()
    fx = q["face_x"].T.cpu().numpy()
    assert abs(q["vol_jxw"].sum().item() - 1.0) < 1e-12
    fq = torch.from_numpy(np.ascontiguousarray(f(vx))).cuda()
    gq = torch.from_numpy(np.ascontiguousarray(g(fx))).cuda()
    rhs = torch.empty(N, dtype=torch.float64, device="cuda")
    op.assemble_rhs(rhs, fq, gq, stiffness=0.7)
    op.synchronize()
    assert np.abs(rhs.cpu().numpy() - b_ref).max() <= TOL * np.abs(b_ref).max()
    ex = torch.from_numpy(np.ascontiguousarray(g(vx))).cuda()
    exg = torch.from_numpy(np.ascontiguousarray(gradg(vx).T)).cuda()
    l2, h1 = op.error_norms(torch.from_numpy(uh).cuda(), ex, exg)
    assert abs(l2 - np.sqrt(l2_ref)) <= 1e-12 * np.sqrt(l2_ref)
    assert abs(h1 - np.sqrt(h1_ref)) <= 1e-12 * np.sqrt(h1_ref)
    l2_only, none = op.error_norms(torch.from_numpy(uh).cuda(), ex)
    assert none is None and l2_only == l2


@pytest.mark.parametrize("dim,n,p,order,distort,kw", [
    (2, 8, 1, 0, (0.25, 11), {}),
    (2, 6, 2, 1, (0.2, 5), dict(stiffness=1.3, mass=0.7)),
    (2, 5, 3, 1, (0.2, 7), {}),
    (2, 4, 4, 0, (0.15, 2), dict(boundary=False)),
    (3, 4, 1, 0, (0.2, 20251018), dict(stiffness=1e-4, mass=1.5e4, boundary=False)),  # monodomain, config E
    (3, 4, 2, 0, (0.2, 20251018), {}),
    (3, 3, 2, 1, (0.25, 1), dict(interior=False)),
    (3, 3, 3, 1, (0.2, 9), dict(mass=2.0)),
    (3, 4, 2, 0, None, {}),  # Cartesian: also equals the stencil kernel
])
def test_mapped_fine_mesh_vmult(dim, n, p, order, distort, kw):
    """PD_VMULT_MAPPED_FINE: LaplaceOperatorDG / MonodomainOperatorDG semantics with the mapped
    FE_DGQ basis on distorted cells (include/utils.h:819-925, 1565-1659); checker: the oracle's
    restatement of the matrix-based twin (examples/monodomain_DG3D.cc:1374-1622)."""
    pdl = gpu()
    import torch

    ogrid = po.Grid(dim, n, 0.0, 1.0, order)
    if distort:
        ogrid.distort_random(*distort)
    groups = [[c] for c in range(ogrid.n_cells)]
    _, pah = product_handler(ogrid, groups, p, p + 1)
    op = pdl.SIPOperator(pah.flatten(penalty_constant=max(p, 1) * (p + 1.0), h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
    assert op.mapped_fine_available
    flags = pdl.ASSEMBLE_VOLUME
    if kw.get("boundary", True):
        flags |= pdl.ASSEMBLE_BOUNDARY
    if kw.get("interior", True):
        flags |= pdl.ASSEMBLE_INTERIOR
    op.set_operator(flags, kw.get("stiffness", 1.0), kw.get("mass", 0.0))
    x = src_vector(op.m())
    yref = po.mapped_fine_vmult(ogrid, p, p + 1, x, **kw)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros_like(xd)
    op.vmult(yd, xd, mode=pdl.VMULT_MAPPED_FINE)
    op.synchronize()
    scale = np.abs(yref).max()
    assert np.abs(yd.cpu().numpy() - yref).max() <= TOL * scale
    op.vmult_add(yd, xd, mode=pdl.VMULT_MAPPED_FINE)
    op.synchronize()
    assert np.abs(yd.cpu().numpy() - 2 * yref).max() <= 2 * TOL * scale
    if distort is None:
        ys = torch.empty_like(xd)
        op.vmult(ys, xd, mode=pdl.VMULT_MATRIX_FREE)
        op.synchronize()
        assert np.abs(ys.cpu().numpy() - yref).max() <= TOL * scale


def test_mapped_fine_mesh_unavailable_on_agglomerates():
    pdl = gpu()
    import torch

    oah, pah = both(2, 8, "blocks2", 1, nq=2)
    op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    assert not op.mapped_fine_available
    x = torch.zeros(op.m(), dtype=torch.float64, device="cuda")
    with pytest.raises(pdl.PolydealError):
        op.vmult(torch.empty_like(x), x, mode=pdl.VMULT_MAPPED_FINE)


@pytest.mark.parametrize("dim,n,shape,p,nq,distort,kw", [
    (2, 16, "blocks4", 1, 2, None, {}),
    (2, 8, "random5", 2, 3, (0.2, 3), {}),
    (2, 8, "random4", 4, 5, None, {}),
    (3, 8, "blocks4", 2, 3, None, {}),
    (3, 8, "random12", 2, 3, None, dict(mass_coeff=0.5, penalty_constant=40.0)),
    (3, 4, "random3", 3, 4, None, {}),
    (3, 4, "random4", 1, 2, None, dict(stiffness_coeff=1e-4, mass_coeff=1.5e4, with_boundary=False)),
    (3, 4, "singletons", 2, 3, None, dict(penalty_constant=6.0, h_rule=3)),
])
@pytest.mark.parametrize("kernels", ["default", "pointwise"])
def test_polytopal_matrix_free_vmult(dim, n, shape, p, nq, distort, kw, kernels, monkeypatch):
    """PD_VMULT_MATRIX_FREE on genuine agglomerates, same operator as the assembled one (checker: oracle matrix).
    Axis-aligned sub-cells: sum factorisation per sub-cell / sub-face (k_cart_apply); otherwise, and with
    PD_POLY_APPLY=pointwise, the basis is regenerated at the agglomerated quadrature points (k_pw_*)."""
    pdl = gpu()
    import torch

    if kernels == "pointwise":
        if distort is not None:
            pytest.skip("distorted meshes take the point-wise kernels by default")
        monkeypatch.setenv("PD_POLY_APPLY", "pointwise")

    oah, pah = both(dim, n, shape, p, nq=nq, distort=distort)
    okw = dict(kw)
    okw.setdefault("penalty_constant", None)
    ref = po.assemble_dg_matrix(oah, degree=p, n_threads=4, **okw)
    fkw = {k: v for k, v in kw.items() if k in ("penalty_constant", "h_rule", "visit_rule")}
    fkw.setdefault("penalty_constant", -1.0)
    op = pdl.SIPOperator(pah.flatten(**fkw), keepalive=pah)
    op.force_generic_matrix_free(True)
    flags = pdl.ASSEMBLE_ALL if kw.get("with_boundary", True) else (pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR)
    op.set_operator(flags, kw.get("stiffness_coeff", 1.0), kw.get("mass_coeff", 0.0))
    x = src_vector(op.m())
    yref = ref.vmult(x)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros_like(xd)
    op.vmult(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    scale = np.abs(yref).max()
    assert np.abs(yd.cpu().numpy() - yref).max() <= TOL * scale
    op.vmult_add(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    assert np.abs(yd.cpu().numpy() - 2 * yref).max() <= 2 * TOL * scale


# ----------------------------------------------------------------------------------
# row (e): sharded assembly + vmult, the ranks emulated one after the other on ONE GPU
# (B200_PROFILING.md: with fewer GPUs than ranks, emulate; the real multi-process path is
# tests/run_distributed_check.py under torchrun and the gloo tests on CPU).
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("world,dim,n,shape,p,kw", [
    (2, 2, 8, "blocks2", 1, {}),
    (3, 2, 8, "random9", 2, {}),
    (2, 3, 4, "random7", 2, dict(visit_rule=1)),
    (4, 3, 4, "blocks2", 1, dict(penalty_constant=10.0, h_rule=1)),
    (2, 3, 4, "singletons", 2, dict(penalty_constant=6.0, h_rule=3)),
])
def test_sharded_assembly_and_vmult_match_serial(world, dim, n, shape, p, kw):
    pdl = gpu()
    import torch

    from polydeal_b200 import distributed as pdd

    oah, pah = both(dim, n, shape, p, order=1)
    okw = dict(kw)
    okw.setdefault("penalty_constant", None)
    A = po.assemble_dg_matrix(oah, degree=p, n_threads=4, **okw).scipy().tocsr()
    nd = oah.n_dofs_per_cell
    x = src_vector(A.shape[0])
    y = A @ x
    owner = pdd.partition_by_blocks(pah, world)
    pkw = dict(kw)
    pkw.setdefault("penalty_constant", -1.0)
    seen_rows = 0
    for rank in range(world):
        part = pdd.LocalPart(pah, owner, rank, **pkw)
        op = pdl.SIPOperator(part.desc, keepalive=(pah, part))
        op.assemble()
        rows = part.owned_global_dofs()
        cols = np.concatenate([rows, part.ghost_global_dofs()])
        ref = A[rows][:, cols].tocsr()
        ref.sort_indices()
        got = op.scipy()
        assert got.shape == (len(rows), len(cols))
        got.sort_indices()
        np.testing.assert_array_equal(got.indptr, ref.indptr)
        np.testing.assert_array_equal(got.indices, ref.indices)
        rp, _ = op.pattern()
        assert_blocks_close(op.values(), ref.data, nd, rp, TOL)
        # vmult with the ghost section filled from the global vector (= what the exchange delivers)
        xd = torch.from_numpy(x[cols]).cuda()
        yd = torch.empty(len(rows), dtype=torch.float64, device="cuda")
        op.vmult_ptr(yd.data_ptr(), xd.data_ptr())
        op.synchronize()
        assert np.abs(yd.cpu().numpy() - y[rows]).max() <= TOL * np.abs(y).max()
        op.set_operator()
        for force in ([False, True] if op.matrix_free_available else [True]):
            op.force_generic_matrix_free(force)
            ym = torch.empty_like(yd)
            op.vmult_ptr(ym.data_ptr(), xd.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)
            op.synchronize()
            assert np.abs(ym.cpu().numpy() - y[rows]).max() <= TOL * np.abs(y).max()
        seen_rows += len(rows)
    assert seen_rows == A.shape[0]


@pytest.mark.parametrize("dim,n,p,world,how", [(3, 16, 2, 3, "metis"), (3, 8, 2, 2, "diagonal"), (2, 32, 2, 3, "metis"), (2, 32, 4, 2, "diagonal")])
def test_sharded_fine_mesh_ragged_partition(dim, n, p, world, how):
    """The fine-mesh matrix-free operator on the ranks of a partition that cuts through the 4x4x4 / 8x8 blocks of the
    Morton curve (METIS, or a diagonal cut): the pipelined kernel k_fine_stream takes tiles of any length and 16-byte
    phase (own runs by one bulk copy + single doubles, results likewise), ghost cells from the ghost section.
    Ranks emulated one after the other; checker: the oracle's matrix of the same form (include/utils.h:819-925)."""
    pdl = gpu()
    import torch

    from polydeal_b200 import distributed as pdd

    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    groups = [[c] for c in range(ogrid.n_cells)]
    _, oah = oracle_handler(dim, n, groups, p, p + 1, order=0)
    _, pah = product_handler(oah.grid, groups, p, p + 1)
    C_ = max(p, 1) * (p + 1.0)
    A = po.assemble_dg_matrix(oah, penalty_constant=C_, h_rule=po.H_NORMAL_EXTENT, n_threads=4).scipy().tocsr()
    x = src_vector(A.shape[0])
    y = A @ x
    if how == "metis":
        owner = pdd.partition_by_metis(pah, world)
    else:  # cells by the sum of their centre coordinates: every block near the cut is split
        v, cv, _ = oah.grid.arrays()
        ctr = v[cv].mean(axis=1).sum(axis=1)  # (polytope k is the cell k)
        owner = np.minimum((ctr / ctr.max() * world * 0.999).astype(np.int32), world - 1)
    streamed = 0
    for rank in range(world):
        part = pdd.LocalPart(pah, owner, rank, penalty_constant=C_, h_rule=pdl.H_NORMAL_EXTENT)
        op = pdl.SIPOperator(part.desc, keepalive=(pah, part))
        assert op.matrix_free_available
        rows = part.owned_global_dofs()
        cols = np.concatenate([rows, part.ghost_global_dofs()])
        xd = torch.from_numpy(x[cols]).cuda()
        yd = torch.full((len(rows),), 0.5, dtype=torch.float64, device="cuda")
        op.vmult_ptr(yd.data_ptr(), xd.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)
        op.synchronize()
        assert np.abs(yd.cpu().numpy() - y[rows]).max() <= TOL * np.abs(y).max()
        op.vmult_ptr(yd.data_ptr(), xd.data_ptr(), mode=pdl.VMULT_MATRIX_FREE, add=True)
        op.synchronize()
        assert np.abs(yd.cpu().numpy() - 2 * y[rows]).max() <= 2 * TOL * np.abs(y).max()
        streamed += op.fine_kernel_last == 3
    assert streamed == world  # the pipelined kernel ran on every rank


def test_poisson_golden_l2_error_on_gpu(goldens):
    """test/polydeal/poisson.cc / poisson.output: L2 error 0.00647702 with the matrix assembled
    by the CUDA path (RHS and error functional from the checker's tables, direct solve on the
    host), and the same system solved by CG running on the GPU through pd_vmult."""
    pdl = gpu()
    import scipy.sparse.linalg as spla
    import torch

    from test_oracle_sip import poisson_golden_problem, poisson_rhs_and_error

    grid, oah, kw = poisson_golden_problem()
    groups = [oah.get_agglomerate(p).tolist()[-1:] + oah.get_agglomerate(p).tolist()[:-1] for p in range(oah.n_polytopes)]
    _, pah = product_handler(grid, groups, 1, 3)
    op = pdl.assemble_dg_matrix(pah, penalty_constant=20.0, h_rule=pdl.H_CONSTANT, h_const=2.0 / 64,
                                visit_rule=pdl.VISIT_BY_INDEX)
    b, l2_error = poisson_rhs_and_error(grid, oah)
    u = spla.spsolve(op.scipy().tocsc(), b)
    assert l2_error(u) == pytest.approx(goldens["poisson"][0], abs=5e-9)
    # conjugate gradients on the device: the loop SolverCG runs around vmult in the reference
    # (examples/diffusion_reaction.cc:721-724), Jacobi-preconditioned with pd_diagonal_inverse
    bd = torch.from_numpy(b).cuda()
    x = torch.zeros_like(bd)
    dinv = torch.empty_like(bd)
    op.get_matrix_diagonal_inverse(dinv)
    op.synchronize()
    r = bd.clone()
    z = dinv * r
    p_ = z.clone()
    Ap = torch.empty_like(bd)
    rz = torch.dot(r, z)
    for it in range(5000):
        torch.cuda.synchronize()
        op.vmult(Ap, p_)
        op.synchronize()
        alpha = rz / torch.dot(p_, Ap)
        x += alpha * p_
        r -= alpha * Ap
        if float(torch.linalg.norm(r)) < 1e-12 * float(torch.linalg.norm(bd)):
            break
        z = dinv * r
        rz_new = torch.dot(r, z)
        p_ = z + (rz_new / rz) * p_
        rz = rz_new
    assert it < 4999
    assert l2_error(x.cpu().numpy()) == pytest.approx(goldens["poisson"][0], abs=5e-9)
    # and end to end on the device: right-hand side by pd_assemble_rhs from f evaluated at the
    # device quadrature points (examples/poisson.cc:745-761), pd_cg_solve, golden functional
    q = op.quadrature()
    vx = q["vol_x"]
    fq = (8 * np.pi**2) * torch.sin(2 * np.pi * vx[0]) * torch.sin(2 * np.pi * vx[1])
    rhs = torch.empty_like(bd)
    op.assemble_rhs(rhs, fq.contiguous())
    op.synchronize()
    assert np.abs(rhs.cpu().numpy() - b).max() <= TOL * np.abs(b).max()
    xs = torch.zeros_like(bd)
    iters, relres = op.cg_solve(xs, rhs, max_iter=5000, rel_tol=1e-12)
    assert relres <= 1e-12
    assert l2_error(xs.cpu().numpy()) == pytest.approx(goldens["poisson"][0], abs=5e-9)
    # PolyUtils::compute_global_error of the same solution (full agglomerated quadrature)
    exact = (torch.sin(2 * np.pi * vx[0]) * torch.sin(2 * np.pi * vx[1])).contiguous()
    l2_full, _ = op.error_norms(xs, exact)
    assert 0.5 * goldens["poisson"][0] < l2_full < 2.0 * goldens["poisson"][0]


# ----------------------------------------------------------------------------------
# the callers of vmult on the device (CG, Chebyshev smoother, lambda_max)
# ----------------------------------------------------------------------------------
def numpy_chebyshev(A, dinv, b, x0, degree, lam_max, rng):
    """Saad Alg. 12.1 on [lam_max/rng, lam_max] with Jacobi inner preconditioner (the recurrence
    deal.II's PreconditionChebyshev runs)."""
    lmin = lam_max / rng
    theta, delta = 0.5 * (lam_max + lmin), 0.5 * (lam_max - lmin)
    sigma1 = theta / delta
    rho = 1.0 / sigma1
    x = x0.copy()
    d = dinv * (b - A @ x) / theta
    x += d
    for _ in range(1, degree):
        rho_new = 1.0 / (2 * sigma1 - rho)
        d = rho_new * rho * d + 2 * rho_new / delta * dinv * (b - A @ x)
        x += d
        rho = rho_new
    return x


def test_device_cg_chebyshev_and_lambda_max(goldens):
    pdl = gpu()
    import scipy.sparse.linalg as spla
    import torch

    from test_oracle_sip import poisson_golden_problem, poisson_rhs_and_error

    grid, oah, kw = poisson_golden_problem()
    groups = [oah.get_agglomerate(p).tolist()[-1:] + oah.get_agglomerate(p).tolist()[:-1] for p in range(oah.n_polytopes)]
    _, pah = product_handler(grid, groups, 1, 3)
    op = pdl.assemble_dg_matrix(pah, penalty_constant=20.0, h_rule=pdl.H_CONSTANT, h_const=2.0 / 64,
                                visit_rule=pdl.VISIT_BY_INDEX)
    A = op.scipy().tocsr()
    b, l2_error = poisson_rhs_and_error(grid, oah)
    bd = torch.from_numpy(b).cuda()
    # --- CG (Jacobi), CUDA-graph replayed, against the golden and a numpy PCG iteration count
    x = torch.zeros_like(bd)
    iters, relres = op.cg_solve(x, bd, max_iter=4000, rel_tol=1e-12, jacobi=True)
    op.synchronize()
    assert relres <= 1e-12 and iters < 4000
    assert l2_error(x.cpu().numpy()) == pytest.approx(goldens["poisson"][0], abs=5e-9)
    dinv = 1.0 / A.diagonal()
    xr, r = np.zeros_like(b), b.copy()
    z = dinv * r
    p_, rz, k = z.copy(), r @ z, 0
    while np.linalg.norm(r) > 1e-12 * np.linalg.norm(b):
        Ap = A @ p_
        alpha = rz / (p_ @ Ap)
        xr += alpha * p_
        r -= alpha * Ap
        z = dinv * r
        rz, rz_old = r @ z, rz
        p_ = z + (rz / rz_old) * p_
        k += 1
    # same algorithm => same iteration count up to rounding in the dot products (a few per
    # cent at 1e-12) and the 8-iteration check interval of the device loop
    assert abs(iters - k) <= 0.05 * k + 8
    # second solve with the same vectors replays the cached graph
    x.zero_()
    iters2, _ = op.cg_solve(x, bd, max_iter=4000, rel_tol=1e-12, jacobi=True)
    assert iters2 == iters
    # unpreconditioned variant
    x2 = torch.zeros_like(bd)
    it3, rr3 = op.cg_solve(x2, bd, max_iter=8000, rel_tol=1e-10, jacobi=False)
    assert rr3 <= 1e-10 and np.abs(x2.cpu().numpy() - xr).max() <= 1e-7 * np.abs(xr).max()
    # --- lambda_max(D^-1 A) by power iteration: a lower bound converging to the true value
    lam_true = float(spla.eigs(spla.aslinearoperator(A.multiply(dinv[:, None]).tocsr()), k=1, which="LM",
                               return_eigenvectors=False)[0].real)
    lam = op.estimate_lambda_max(60)
    assert 0.9 * lam_true <= lam <= lam_true * (1 + 1e-10)
    # --- Chebyshev smoother (degree 3, range 20 as in examples/matrix_free_agglo.cc:306-308)
    for zero_guess in (True, False):
        x0 = np.zeros_like(b) if zero_guess else np.cos(0.1 * np.arange(len(b)))
        want = numpy_chebyshev(A, dinv, b, x0, 3, 1.2 * lam, 20.0)
        xs = torch.from_numpy(x0.copy()).cuda()
        op.chebyshev_smooth(xs, bd, 3, 1.2 * lam, 20.0, zero_initial_guess=zero_guess)
        op.synchronize()
        assert np.abs(xs.cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()


# ----------------------------------------------------------------------------------
# "next" row N2: level transfers (include/utils.h:95-270, include/poly_utils.h:1469-1634,
# source/multigrid_amg.cc:66-110)
# ----------------------------------------------------------------------------------
def _dgq_unit_support_points(dim, p):
    g = po.gauss_lobatto_nodes(p + 1)
    return np.array([[g[(i // (p + 1) ** d) % (p + 1)] for d in range(dim)] for i in range((p + 1) ** dim)])  # x fastest


def _nested_levels(dim, n, p, coarse_shape, fine_shape, distort=None, seed=1):
    """Two agglomeration levels of the same mesh, the fine one nested in the coarse one."""
    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    if distort:
        ogrid.distort_random(*distort)
    fine_groups = groups_for(fine_shape, dim, n, ogrid, seed)
    if coarse_shape.startswith("blocks"):
        coarse_groups = groups_for(coarse_shape, dim, n, ogrid, seed)
    else:  # unions of consecutive fine groups: irregular, still nested
        k = int(coarse_shape[5:])
        coarse_groups = [sum((list(g) for g in fine_groups[i:i + k]), []) for i in range(0, len(fine_groups), k)]
    cell_to_coarse = {}
    for K_, g in enumerate(coarse_groups):
        for c in g:
            cell_to_coarse[int(c)] = K_
    parent = np.array([cell_to_coarse[int(g[0])] for g in fine_groups], dtype=np.int32)
    for q, g in enumerate(fine_groups):
        assert all(cell_to_coarse[int(c)] == parent[q] for c in g)
    return ogrid, coarse_groups, fine_groups, parent


@pytest.mark.parametrize("dim,n,p,coarse,fine,distort", [
    (2, 8, 1, "blocks4", "blocks2", None),
    (2, 8, 2, "union3", "random4", (0.2, 3)),
    (2, 8, 4, "blocks4", "blocks2", None),
    (3, 4, 1, "union2", "random5", None),
    (3, 4, 2, "blocks4", "blocks2", (0.15, 2)),
    (3, 4, 3, "blocks2", "singletons", None),
])
def test_injection_between_agglomeration_levels(dim, n, p, coarse, fine, distort):
    pdl = gpu()
    import torch

    ogrid, cg, fg, parent = _nested_levels(dim, n, p, coarse, fine, distort)
    ops, oahs = [], []
    for groups in (cg, fg):
        _, oah = oracle_handler(dim, n, groups, p, p + 1, distort=distort)
        _, pah = product_handler(oah.grid, groups, p, p + 1)
        ops.append(pdl.SIPOperator(pah.flatten(), keepalive=pah))
        oahs.append(oah)
    (cop, fop), (coah, foah) = ops, oahs
    T = pdl.Transfer(cop, fop, parent)
    nd = (p + 1) ** dim
    assert (T.m(), T.n()) == (foah.n_dofs, coah.n_dofs)
    # checker: the injection matrix of the reference, entry by entry
    usp = _dgq_unit_support_points(dim, p)
    P = np.zeros((foah.n_dofs, coah.n_dofs))
    for q in range(foah.n_polytopes):
        flo, fhi = foah.bbox(q)
        clo, chi = coah.bbox(int(parent[q]))
        rows, cols = foah.get_dof_indices(q), coah.get_dof_indices(int(parent[q]))
        for i in range(nd):
            real = flo + usp[i] * (fhi - flo)  # fine_bbox.unit_to_real
            phi, _ = po.fe_evaluate(po.FE_DGQ, dim, p, (real - clo) / (chi - clo))  # coarse_bbox.real_to_unit
            P[rows[i], cols] = phi
    x = src_vector(coah.n_dofs)
    y = np.cos(0.11 * np.arange(foah.n_dofs)) + 0.3
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out_f = torch.full((foah.n_dofs,), 7.0, dtype=torch.float64, device="cuda")
    T.prolongate(out_f, xd)
    T.synchronize()
    ref = P @ x
    assert np.abs(out_f.cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
    T.prolongate_and_add(out_f, xd)
    T.synchronize()
    assert np.abs(out_f.cpu().numpy() - 2 * ref).max() <= 2 * TOL * np.abs(ref).max()
    out_c = torch.full((coah.n_dofs,), -3.0, dtype=torch.float64, device="cuda")
    T.restrict(out_c, yd)
    T.synchronize()
    ref = P.T @ y
    assert np.abs(out_c.cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
    T.restrict_and_add(out_c, yd)
    T.synchronize()
    assert np.abs(out_c.cpu().numpy() - 2 * ref).max() <= 2 * TOL * np.abs(ref).max()
    # test/polydeal/distributed_injection_01: a polynomial of the space is reproduced exactly
    # (interpolate on the coarse level, inject, compare with the interpolant on the fine level)
    f = lambda X: X.sum(axis=-1) - 1.0 if p == 1 else (X**2).sum(axis=-1) - 1.0
    def interpolant(oah):
        v = np.zeros(oah.n_dofs)
        for k in range(oah.n_polytopes):
            lo, hi = oah.bbox(k)
            v[oah.get_dof_indices(k)] = f(lo + usp * (hi - lo))
        return v
    T.prolongate(out_f, torch.from_numpy(interpolant(coah)).cuda())
    T.synchronize()
    assert np.abs(out_f.cpu().numpy() - interpolant(foah)).max() <= 1e-13


@pytest.mark.parametrize("dim,n,p,shape,distort", [
    (2, 8, 1, "random5", (0.25, 4)),
    (2, 8, 3, "blocks4", None),
    (3, 4, 2, "random6", (0.2, 6)),
    (3, 4, 3, "blocks2", None),
])
def test_interpolation_to_the_fine_mesh_space(dim, n, p, shape, distort):
    pdl = gpu()
    import torch

    oah, pah = both(dim, n, shape, p, distort=distort)
    op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    T = pdl.Transfer.to_cells(op)
    nd = (p + 1) ** dim
    n_cells = oah.grid.n_cells
    assert (T.m(), T.n()) == (n_cells * nd, oah.n_dofs)
    usp = _dgq_unit_support_points(dim, p)
    P = np.zeros((n_cells * nd, oah.n_dofs))
    for k in range(oah.n_polytopes):
        lo, hi = oah.bbox(k)
        cols = oah.get_dof_indices(k)
        for c in oah.get_agglomerate(k):
            V = oah.grid.cell_vertices(int(c))
            for i in range(nd):
                w = np.array([np.prod([usp[i][d] if (v >> d) & 1 else 1 - usp[i][d] for d in range(dim)]) for v in range(1 << dim)])
                real = w @ V  # the Q1-mapped support point of the fine cell
                phi, _ = po.fe_evaluate(po.FE_DGQ, dim, p, (real - lo) / (hi - lo))
                P[int(c) * nd + i, cols] = phi
    x = src_vector(oah.n_dofs)
    y = np.cos(0.13 * np.arange(n_cells * nd)) - 0.2
    out_f = torch.empty(n_cells * nd, dtype=torch.float64, device="cuda")
    T.prolongate(out_f, torch.from_numpy(x).cuda())
    T.synchronize()
    ref = P @ x
    assert np.abs(out_f.cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()
    out_c = torch.empty(oah.n_dofs, dtype=torch.float64, device="cuda")
    T.restrict(out_c, torch.from_numpy(y).cuda())
    T.synchronize()
    ref = P.T @ y
    assert np.abs(out_c.cpu().numpy() - ref).max() <= TOL * np.abs(ref).max()


# ----------------------------------------------------------------------------------
# row 5 / "next" row N4: FE_AggloDGP on the bounding box (source/fe_agglodgp.cc:28-57)
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n,shape,p,nq,distort,kw", [
    (2, 8, "blocks4", 1, 2, None, {}),
    (2, 8, "random5", 2, 3, (0.2, 3), {}),
    (2, 8, "random4", 3, 4, None, dict(mass_coeff=0.5)),
    (2, 8, "blocks2", 4, 5, None, {}),
    (3, 4, "random6", 1, 2, None, {}),
    (3, 4, "blocks2", 2, 3, (0.15, 2), dict(mass_coeff=2.0, stiffness_coeff=0.3)),
    (3, 4, "random3", 3, 4, None, {}),
    (3, 4, "singletons", 2, 3, None, dict(with_boundary=False)),
])
@pytest.mark.parametrize("kernels", ["default", "generic"])
def test_assembly_with_fe_agglodgp(dim, n, shape, p, nq, distort, kw, kernels, monkeypatch):
    """assemble_dg_matrix with FE_AggloDGP<dim>(p): C(p+dim, dim) Legendre products per polytope."""
    if kernels == "generic":
        if distort is not None:
            pytest.skip("distorted meshes take the DMMA kernels by default")
        monkeypatch.setenv("PD_ASSEMBLE_KERNELS", "generic")
    pdl = gpu()
    import math
    import torch

    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    groups = groups_for(shape, dim, n, ogrid, 1)
    _, oah = oracle_handler(dim, n, groups, p, nq, distort=distort, fe_kind=po.FE_AGGLODGP)
    v, cv, nb = oah.grid.arrays()
    pah = pdl.AgglomerationHandler(pdl.Grid.from_arrays(v, cv, nb))
    for g_ in groups:
        pah.define_agglomerate(g_)
    pah.initialize_fe_values(nq)
    pah.distribute_agglomerated_dofs(pdl.FE_AGGLODGP, p)
    nd = math.comb(p + dim, dim)
    assert pah.n_dofs_per_cell == oah.n_dofs_per_cell == nd
    ref = po.assemble_dg_matrix(oah, degree=p, n_threads=4, **kw)
    op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    flags = pdl.ASSEMBLE_ALL if kw.get("with_boundary", True) else (pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR)
    op.assemble(flags, kw.get("stiffness_coeff", 1.0), kw.get("mass_coeff", 0.0))
    rowptr, cols = op.pattern()
    rp, rc, rv = ref.csr()
    assert np.array_equal(rowptr, rp) and np.array_equal(cols, rc)  # sparsity bit-exact
    assert_blocks_close(op.values(), rv, nd, rowptr)
    x = src_vector(op.m())
    y = torch.empty(op.m(), dtype=torch.float64, device="cuda")
    op.vmult(y, torch.from_numpy(x).cuda())
    op.synchronize()
    yref = ref.vmult(x)
    assert np.abs(y.cpu().numpy() - yref).max() <= TOL * np.abs(yref).max()
    # the point-wise kernels on the Legendre basis: matrix-free apply, right-hand side, error norms
    op.set_operator(flags, kw.get("stiffness_coeff", 1.0), kw.get("mass_coeff", 0.0))
    op.vmult(y, torch.from_numpy(x).cuda(), mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    assert np.abs(y.cpu().numpy() - yref).max() <= TOL * np.abs(yref).max()
    f = lambda X: np.sin(1.1 * X[..., 0]) + X[..., 1] ** 2
    b_ref, l2_ref = np.zeros(oah.n_dofs), 0.0
    uh = 0.1 * x
    for k in range(oah.n_polytopes):
        fev = oah.reinit(k)
        dofs = oah.get_dof_indices(k)
        b_ref[dofs] += fev.values @ (f(fev.points) * fev.JxW)
        l2_ref += np.sum((uh[dofs] @ fev.values - f(fev.points)) ** 2 * fev.JxW)
    q = op.quadrature()
    fq = torch.from_numpy(np.ascontiguousarray(f(q["vol_x"].T.cpu().numpy()))).cuda()
    rhs = torch.empty(op.m(), dtype=torch.float64, device="cuda")
    op.assemble_rhs(rhs, fq)
    op.synchronize()
    assert np.abs(rhs.cpu().numpy() - b_ref).max() <= TOL * np.abs(b_ref).max()
    l2, _ = op.error_norms(torch.from_numpy(uh).cuda(), fq)
    assert abs(l2 - np.sqrt(l2_ref)) <= 1e-12 * np.sqrt(l2_ref)
    with pytest.raises(pdl.PolydealError):
        op.vmult(y, torch.from_numpy(x).cuda(), mode=pdl.VMULT_MAPPED_FINE)


# ----------------------------------------------------------------------------------
# multi-rank collectives over peer memory, on ONE GPU: two processes share the device
# ----------------------------------------------------------------------------------
def test_peer_memory_collectives_two_processes_one_gpu():
    """pd_peer_* (ghost exchange, all-reduce) and pd_cg_solve_sharded between two PROCESSES that share
    cuda:0 (CUDA IPC + epoch flags; the kernels of the two contexts are time-sliced).  The same script
    runs one rank per GPU under torchrun with NCCL (tests/run_distributed_check.py)."""
    gpu()
    import socket
    import subprocess
    import sys

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # PD_PEER_CSR_SPLIT_MIN_NNZ=0: also take the interior / boundary split of the block-CSR apply, which is
    # reserved for large matrices by default
    env = dict(os.environ, PD_CHECK_SAME_DEVICE="1", OMP_NUM_THREADS="2", PD_PEER_CSR_SPLIT_MIN_NNZ="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tests", "run_distributed_check.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DISTRIBUTED CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("kind", ["polytopal", "fine", "mapped"])
def test_matrix_free_inverse_diagonal_and_jacobi_cg(kind):
    """get_matrix_diagonal_inverse() of the matrix-free operators without an assembled matrix
    (MatrixFreeTools::compute_diagonal, include/utils.h:929-1100): unit vectors over independent sets of
    polytopes; then Jacobi-CG and the Chebyshev machinery run on the matrix-free operator alone."""
    pdl = gpu()
    import torch

    if kind == "polytopal":
        oah, pah = both(3, 4, "random6", 2, nq=3)
        op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
        mode = pdl.VMULT_MATRIX_FREE
        A = po.assemble_dg_matrix(oah, degree=2, n_threads=4).scipy().tocsr()
        apply_ref = lambda v: A @ v
        diag_ref = A.diagonal()
    else:
        dim, n, p = (3, 3, 2) if kind == "fine" else (2, 5, 2)
        ogrid = po.Grid(dim, n, 0.0, 1.0, 1)
        if kind == "mapped":
            ogrid.distort_random(0.2, 3)
        groups = [[c] for c in range(ogrid.n_cells)]
        _, pah = product_handler(ogrid, groups, p, p + 1)
        C_ = max(p, 1) * (p + 1.0)
        op = pdl.SIPOperator(pah.flatten(penalty_constant=C_, h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
        mode = pdl.VMULT_MATRIX_FREE if kind == "fine" else pdl.VMULT_MAPPED_FINE
        apply_ref = lambda v: po.mapped_fine_vmult(ogrid, p, p + 1, v)
        N_ = ogrid.n_cells * (p + 1) ** dim
        diag_ref = np.array([apply_ref(np.eye(1, N_, k).ravel())[k] for k in range(N_)])
    N = op.m()
    dinv = torch.empty(N, dtype=torch.float64, device="cuda")
    op.get_matrix_diagonal_inverse(dinv, mode=mode)
    op.synchronize()
    got = 1.0 / dinv.cpu().numpy()
    assert np.abs(got - diag_ref).max() <= TOL * np.abs(diag_ref).max()
    # Jacobi-preconditioned CG on the matrix-free operator: b = A x*  =>  x*
    xs = src_vector(N)
    b = torch.from_numpy(apply_ref(xs)).cuda()
    x = torch.zeros_like(b)
    iters, relres = op.cg_solve(x, b, max_iter=3000, rel_tol=1e-11, jacobi=True, mode=mode)
    assert relres <= 1e-11
    assert np.abs(x.cpu().numpy() - xs).max() <= 1e-7 * np.abs(xs).max()
    lam = op.estimate_lambda_max(30, mode=mode)
    assert lam > 0


# ----------------------------------------------------------------------------------
# the reference's exactness tests, entirely on the device: exact_solutions.cc, exact_solutions_dgp.cc,
# continuous_face_exact_solution.cc, disconnected_exact_solution.cc
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("fe,kind,shape", [
    ("dgq", "linear", "blocks"), ("dgq", "quadratic", "blocks"),
    ("dgp", "linear", "blocks"), ("dgp", "quadratic", "blocks"),
    ("dgq", "quadratic", "disconnected"), ("dgp", "quadratic", "disconnected"),
])
def test_exact_solutions_on_the_device(fe, kind, shape):
    """Distorted 4x4 grid of [0,1]^2, penalty 10 / diameter, Dirichlet data from the exact solution:
    assembly, right-hand side (volume + boundary terms), Jacobi-CG and the error norms all run on the GPU;
    "Linear: OK" / "Quadratic: OK" = volume 1, perimeter 4, L2 error and H1 seminorm at round-off."""
    pdl = gpu()
    import torch

    p = 1 if kind == "linear" else 2
    groups = sc.blocks_2x2_of_4x4() if shape == "blocks" else [[0, 1, 2, 3], [12, 13, 14, 15], [4, 5, 6, 7], [8, 9, 10], [11]]
    grid = pdl.Grid.hyper_cube(2, 0.0, 1.0, 2)
    grid.distort_random(0.25, 5)
    ah = pdl.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(2 * p + 1)
    ah.distribute_agglomerated_dofs(pdl.FE_DGQ if fe == "dgq" else pdl.FE_AGGLODGP, p)
    op = pdl.SIPOperator(ah.flatten(penalty_constant=10.0, visit_rule=pdl.VISIT_BY_INDEX), keepalive=ah)
    op.assemble()
    q = op.quadrature()
    assert abs(q["vol_jxw"].sum().item() - 1.0) < 1e-14
    vx, fx = q["vol_x"], q["face_x"]
    if p == 1:
        exact = lambda X: X[0] + X[1] - 1.0
        grad = lambda X: torch.stack([torch.ones_like(X[0]), torch.ones_like(X[0])])
        f = torch.zeros_like(vx[0])
    else:
        exact = lambda X: X[0] ** 2 + X[1] ** 2 - 1.0
        grad = lambda X: torch.stack([2 * X[0], 2 * X[1]])
        f = torch.full_like(vx[0], -4.0)
    rhs = torch.empty(op.m(), dtype=torch.float64, device="cuda")
    op.assemble_rhs(rhs, f.contiguous(), exact(fx).contiguous())
    x = torch.zeros_like(rhs)
    iters, relres = op.cg_solve(x, rhs, max_iter=5000, rel_tol=1e-14)
    l2, h1 = op.error_norms(x, exact(vx).contiguous(), grad(vx).contiguous())
    assert l2 < 1e-11 and h1 < 1e-10, (l2, h1, iters, relres)


def test_coarse_operator_from_matrix_free(goldens):
    """test/polydeal/coarse_operator_from_matrix_free.cc: the Galerkin coarse operator P^T A P of the fine-mesh
    MATRIX-FREE cell Laplacian (MatrixFreeProjector::compute_level_matrices), P = fill_interpolation_matrix.
    <v, A v> on the fine mesh and <v_c, P^T A P v_c> on the agglomerates agree and equal 0, 1, 2 for
    v = 1, x, x + y (the reference prints them for an R-tree and a gmsh agglomeration; the numbers are
    partition independent -- 66 random connected agglomerates of the 64x64 grid here).  Everything on the GPU:
    prolongate, matrix-free apply (volume term only), restrict."""
    pdl = gpu()
    import torch

    g = goldens["coarse_operator_from_matrix_free"]
    p = 1
    ogrid = po.Grid(2, 64, 0.0, 1.0, 0)
    _, pah_fine = product_handler(ogrid, [[c] for c in range(ogrid.n_cells)], p, p + 1)
    fine = pdl.SIPOperator(pah_fine.flatten(penalty_constant=2.0, h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah_fine)
    fine.set_operator(pdl.ASSEMBLE_VOLUME, 1.0, 0.0)  # the test's operator has no face terms
    assert fine.matrix_free_available
    groups = groups_for("random66", 2, 64, ogrid, 4)
    _, pah = product_handler(ogrid, groups, p, p + 1)
    coarse = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    P = pdl.Transfer.to_cells(coarse)
    assert (P.m(), P.n()) == (16384, 264) and fine.m() == 16384  # "Injection matrix has size: (16384,264)"
    usp = _dgq_unit_support_points(2, p)
    verts, cv, _ = ogrid.arrays()

    def interpolate_fine(f):
        out = np.empty(fine.m())
        for c in range(ogrid.n_cells):
            lo, hi = verts[cv[c]].min(axis=0), verts[cv[c]].max(axis=0)
            out[c * 4:(c + 1) * 4] = f(lo + usp * (hi - lo))
        return out

    def interpolate_coarse(f):
        out = np.empty(coarse.m())
        for k in range(pah.n_polytopes):
            lo, hi = pah.bbox(k)
            out[pah.get_dof_indices(k)] = f(lo + usp * (hi - lo))
        return out

    funcs = [lambda X: np.ones(len(X)), lambda X: X[:, 0], lambda X: X[:, 0] + X[:, 1]]
    for f, gf, gc in zip(funcs, g["fine"][::2], g["agglomerated"][::2]):
        vf = torch.from_numpy(interpolate_fine(f)).cuda()
        Av = torch.empty_like(vf)
        fine.vmult(Av, vf, mode=pdl.VMULT_MATRIX_FREE)
        fine.synchronize()
        assert abs(float(vf @ Av) - gf) <= 1e-11
        vc = torch.from_numpy(interpolate_coarse(f)).cuda()
        Pv = torch.empty_like(vf)
        P.prolongate(Pv, vc)
        P.synchronize()
        APv = torch.empty_like(vf)
        fine.vmult(APv, Pv, mode=pdl.VMULT_MATRIX_FREE)
        fine.synchronize()
        PtAPv = torch.empty_like(vc)
        P.restrict(PtAPv, APv)
        P.synchronize()
        assert abs(float(vc @ PtAPv) - round(gc)) <= 1e-11 and abs(gc - round(gc)) < 1e-12


def _sharded_energies(pdl, pah, owner, funcs, **penalty):
    """sum over the ranks of u_owned . (A_rank u) for a SHARDED assembly (owner-computes-rows, cut interfaces
    evaluated from the ghost polytope's bounding box), the ranks emulated one after the other on one GPU;
    volume + interior-face terms only, as in the reference's sanity checks."""
    import torch

    from polydeal_b200 import distributed as pdd

    usp = _dgq_unit_support_points(2, 1)
    u = {k: np.empty(pah.n_dofs) for k in funcs}
    for k_ in range(pah.n_polytopes):
        lo, hi = pah.bbox(k_)
        for name, f in funcs.items():
            u[name][pah.get_dof_indices(k_)] = f(lo + usp * (hi - lo))
    energy = {k: 0.0 for k in funcs}
    for rank in range(int(owner.max()) + 1):
        part = pdd.LocalPart(pah, owner, rank, **penalty)
        op = pdl.SIPOperator(part.desc, keepalive=(pah, part))
        op.assemble(pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR)
        rows = part.owned_global_dofs()
        cols = np.concatenate([rows, part.ghost_global_dofs()])
        for name in funcs:
            xd = torch.from_numpy(u[name][cols]).cuda()
            yd = torch.empty(len(rows), dtype=torch.float64, device="cuda")
            op.vmult_ptr(yd.data_ptr(), xd.data_ptr())
            op.synchronize()
            energy[name] += float(u[name][rows] @ yd.cpu().numpy())
    return energy


_SANITY_FUNCS = {"x": lambda X: X[:, 0], "xplusy": lambda X: X[:, 0] + X[:, 1], "one": lambda X: np.ones(len(X))}


def test_distributed_poisson_sanity_check(goldens):
    """test/polydeal/distributed_poisson_sanity_check_01 (mpirun=3): the energies 1 and 2 of x and x + y summed
    over the ranks of a sharded assembly, three ranks; boundary terms dropped, penalty 10 max(1/hA, 1/hB)."""
    pdl = gpu()
    from polydeal_b200 import distributed as pdd

    g = goldens["distributed_poisson_sanity_check_01"]
    ogrid = po.Grid(2, 32, 0.0, 1.0, 0)
    groups = groups_for("random40", 2, 32, ogrid, 9)
    _, pah = product_handler(ogrid, groups, 1, 3)
    energy = _sharded_energies(pdl, pah, pdd.partition_by_blocks(pah, 3), _SANITY_FUNCS, penalty_constant=10.0,
                               h_rule=pdl.H_MAX_INVERSE_DIAMETER)
    assert abs(energy["x"] - g["x"][0]) <= 1e-11 and abs(energy["xplusy"] - g["xplusy"][0]) <= 1e-11
    assert abs(energy["one"]) <= 1e-11


def test_distributed_poisson_sanity_check_02(goldens):
    """test/polydeal/distributed_poisson_sanity_check_02 (mpirun=3): [0,1]^2 refined 6x, every rank agglomerates ALL
    of its locally owned cells into one polytope (p4est: three contiguous pieces of the Morton curve, 1365-1366
    cells each), DGQ1, QGauss(3), penalty 1/1; three polytopes, one per rank (:115-156, 230-231)."""
    pdl = gpu()
    g = goldens["distributed_poisson_sanity_check_02"]
    ogrid = po.Grid(2, 64, 0.0, 1.0, 0)
    n = ogrid.n_cells
    cuts = [n * r // 3 for r in range(4)]
    groups = [list(range(cuts[r], cuts[r + 1])) for r in range(3)]
    _, pah = product_handler(ogrid, groups, 1, 3)
    energy = _sharded_energies(pdl, pah, np.arange(3, dtype=np.int32), _SANITY_FUNCS, penalty_constant=1.0,
                               h_rule=pdl.H_CONSTANT, h_const=1.0)
    assert abs(energy["x"] - g["x"][0]) <= 1e-11 and abs(energy["xplusy"] - g["xplusy"][0]) <= 1e-11
    assert abs(energy["one"]) <= 1e-11


def test_fully_distributed_poisson_sanity_check_01(goldens):
    """test/polydeal/fully_distributed_poisson_sanity_check_01 (mpirun=3): [0,1]^2 refined 4x (256 cells), ranks =
    strips floor(3 x_centre) (:48-66), METIS into 10 agglomerates inside every rank's region
    (PolyUtils::partition_locally_owned_regions, include/poly_utils.h:629-704), DGQ1, QGauss(3), penalty 1/1.
    METIS versions differ, so the partition is an input; the golden energies do not depend on it."""
    pdl = gpu()
    g = goldens["fully_distributed_poisson_sanity_check_01"]
    ogrid = po.Grid(2, 16, 0.0, 1.0, 0)
    assert ogrid.n_cells == int(g["n_cells"][0])
    v, cv, nbr = ogrid.arrays()
    centre_x = v[cv].mean(axis=1)[:, 0]
    rank_of_cell = np.floor(centre_x * 3).astype(np.int32)
    groups, owner = [], []
    for r in range(3):
        cells = np.nonzero(rank_of_cell == r)[0]
        local = -np.ones(ogrid.n_cells, dtype=np.int64)
        local[cells] = np.arange(len(cells))
        adj = [[local[q] for q in nbr[c] if q >= 0 and local[q] >= 0] for c in cells]  # face graph of the owned region
        xadj = np.concatenate([[0], np.cumsum([len(a) for a in adj])])
        part = pdl.partition_graph(xadj, np.concatenate(adj), 10)
        for k in range(10):
            if (part == k).any():
                groups.append(cells[part == k].tolist())
                owner.append(r)
    _, pah = product_handler(ogrid, groups, 1, 3)
    energy = _sharded_energies(pdl, pah, np.array(owner, dtype=np.int32), _SANITY_FUNCS, penalty_constant=1.0,
                               h_rule=pdl.H_CONSTANT, h_const=1.0)
    assert abs(energy["x"] - g["x"][0]) <= 1e-11 and abs(energy["xplusy"] - g["xplusy"][0]) <= 1e-11
    assert abs(energy["one"]) <= 1e-11


def test_assembly_on_an_unstructured_mesh(goldens):
    """The input grid of test/polydeal/fully_distributed_poisson_sanity_check_02.cc (input_grids/square.msh refined
    once, 364 quadrilaterals read the way GridIn hands them over: neighbours rotated against each other), 30
    agglomerates, DGQ2: the GPU-assembled SIP matrix against the oracle per block entry, and the golden energies of
    x and x + y (1 and 2, no boundary terms, penalty 1/1) with the GPU matrix."""
    pdl = gpu()
    import torch

    g = goldens["fully_distributed_poisson_sanity_check_02"]
    v, cv, nbr = sc.quad_mesh_from_gmsh(g["input_grid"]["verts"], g["input_grid"]["quads"], n_refine=1)
    groups = sc.random_partition(len(nbr), nbr, 30, seed=5)
    handlers = []
    for mod, grid in ((po, po.Grid.from_arrays(v, cv, nbr)), (pdl, pdl.Grid.from_arrays(v, cv, nbr))):
        ah = mod.AgglomerationHandler(grid)
        for gr in groups:
            ah.define_agglomerate(gr)
        ah.initialize_fe_values(3)
        ah.distribute_agglomerated_dofs(mod.FE_DGQ, 2)
        handlers.append(ah)
    oah, pah = handlers
    ref = po.assemble_dg_matrix(oah, degree=2, n_threads=4)
    op = pdl.assemble_dg_matrix(pah)
    rp, cols = op.pattern()
    orp, ocols, ovals = ref.csr()
    np.testing.assert_array_equal(rp, orp)
    np.testing.assert_array_equal(cols, ocols)
    assert_blocks_close(op.values(), ovals, oah.n_dofs_per_cell, rp, TOL)
    # the sanity-check invariants with the GPU operator
    kw = dict(penalty_constant=1.0, h_rule=pdl.H_CONSTANT, h_const=1.0)
    op = pdl.SIPOperator(pah.flatten(**kw), keepalive=pah)
    op.assemble(pdl.ASSEMBLE_VOLUME | pdl.ASSEMBLE_INTERIOR)
    usp = _dgq_unit_support_points(2, 2)
    for name, f in (("x", lambda X: X[:, 0]), ("xplusy", lambda X: X[:, 0] + X[:, 1])):
        u = np.empty(pah.n_dofs)
        for k_ in range(pah.n_polytopes):
            lo, hi = pah.bbox(k_)
            u[pah.get_dof_indices(k_)] = f(lo + usp * (hi - lo))
        xd = torch.from_numpy(u).cuda()
        yd = torch.empty_like(xd)
        op.vmult(yd, xd)
        op.synchronize()
        assert abs(float(u @ yd.cpu().numpy()) - g[name][0]) <= 1e-11


# ----------------------------------------------------------------------------------
# state the handle caches around the operator: solver graphs, uploads, streams
# ----------------------------------------------------------------------------------
def test_cg_after_set_operator_uses_the_new_operator():
    """The replayed CG graph bakes the matrix-free operator in (coefficients are kernel parameters): a second
    solve on the SAME x / b after pd_set_operator must iterate with the new operator, and max_iter is exact."""
    pdl = gpu()
    import torch

    n, p = 4, 2
    ogrid = po.Grid(3, n, 0.0, 1.0, 0)
    groups = [[c] for c in range(ogrid.n_cells)]
    oah = po.AgglomerationHandler(ogrid)
    for g in groups:
        oah.define_agglomerate(g)
    oah.initialize_fe_values(p + 1)
    oah.distribute_agglomerated_dofs(po.FE_DGQ, p)
    _, pah = product_handler(ogrid, groups, p, p + 1)
    C = p * (p + 1.0)
    op = pdl.SIPOperator(pah.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
    b = torch.from_numpy(np.cos(0.3 * np.arange(op.m())) + 0.2).cuda()
    x = torch.zeros_like(b)
    sols = {}
    for mass in (0.0, 50.0, 0.0):
        op.set_operator(pdl.ASSEMBLE_ALL, 1.0, mass)
        x.zero_()
        it, rr = op.cg_solve(x, b, max_iter=2000, rel_tol=1e-11, jacobi=True, mode=pdl.VMULT_MATRIX_FREE)
        assert op.last_cg_converged and rr <= 1e-11
        A = po.assemble_dg_matrix(oah, penalty_constant=C, h_rule=po.H_NORMAL_EXTENT, mass_coeff=mass).scipy().tocsc()
        import scipy.sparse.linalg as spla

        want = spla.spsolve(A, b.cpu().numpy())
        torch.cuda.synchronize()
        assert np.abs(x.cpu().numpy() - want).max() <= 1e-8 * np.abs(want).max(), mass
        sols[mass] = it
    # the iteration budget is exact and a missed tolerance is reported, not silently accepted
    x.zero_()
    it, rr = op.cg_solve(x, b, max_iter=5, rel_tol=1e-14, jacobi=True, mode=pdl.VMULT_MATRIX_FREE)
    assert it == 5 and not op.last_cg_converged and rr > 1e-14
    x.zero_()
    it, rr = op.cg_solve(x, b, max_iter=19, rel_tol=0.0, jacobi=True, mode=pdl.VMULT_MATRIX_FREE)
    assert it == 19 and op.last_cg_converged


def test_upload_rederives_the_fine_mesh_operator_and_rejects_new_topology():
    """pd_upload with moved vertices / bounding boxes: the matrix-free fine-mesh operator (stencil records,
    tile plan, cached inverse diagonal) follows; a descriptor with another topology is refused."""
    pdl = gpu()
    import torch

    p = 2
    C = p * (p + 1.0)

    def build(hi):
        ogrid = po.Grid(3, (4, 4, 4), 0.0, hi, 1)
        groups = [[c] for c in range(ogrid.n_cells)]
        oah = po.AgglomerationHandler(ogrid)
        for g in groups:
            oah.define_agglomerate(g)
        oah.initialize_fe_values(p + 1)
        oah.distribute_agglomerated_dofs(po.FE_DGQ, p)
        _, pah = product_handler(ogrid, groups, p, p + 1)
        return oah, pah, pah.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT)

    oah1, pah1, d1 = build((1.0, 1.0, 1.0))
    oah2, pah2, d2 = build((1.0, 0.7, 1.9))  # same topology, other coordinates / boxes / penalties
    op = pdl.SIPOperator(d1, keepalive=(pah1, pah2))
    x = src_vector(op.m())
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    dinv = torch.empty_like(xd)
    for oah, desc in ((oah1, d1), (oah2, d2), (oah1, d1)):
        op.upload(desc)
        ref = po.assemble_dg_matrix(oah, penalty_constant=C, h_rule=po.H_NORMAL_EXTENT)
        op.vmult(yd, xd, mode=pdl.VMULT_MATRIX_FREE)
        torch.cuda.synchronize()
        yref = ref.vmult(x)
        assert np.abs(yd.cpu().numpy() - yref).max() <= TOL * np.abs(yref).max()
        op.get_matrix_diagonal_inverse(dinv, mode=pdl.VMULT_MATRIX_FREE)
        torch.cuda.synchronize()
        np.testing.assert_allclose(dinv.cpu().numpy(), 1.0 / ref.scipy().diagonal(), rtol=1e-11)
        op.assemble()
        rp, _ = op.pattern()
        assert_blocks_close(op.values(), ref.values(), 27, rp, TOL)
    # another topology: refused, the handle keeps working
    ogrid = po.Grid(3, (4, 4, 4), 0.0, 1.0, 1)
    groups = [[c] for c in range(ogrid.n_cells)]
    groups[0], groups[5] = groups[5], groups[0]  # polytope order (hence dof_block / interface list) differs
    _, pah3 = product_handler(ogrid, groups, p, p + 1)
    d3 = pah3.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT)
    with pytest.raises(pdl.PolydealError, match="topology"):
        op.upload(d3)
    bad = pah1.flatten(penalty_constant=C, h_rule=pdl.H_NORMAL_EXTENT)
    cv = np.ctypeslib.as_array(bad.cell_verts, (bad.n_cells * 8,)).copy()
    cv[17] = bad.n_verts + 3
    import ctypes as Cc

    bad.cell_verts = cv.ctypes.data_as(Cc.POINTER(Cc.c_int32))
    with pytest.raises(pdl.PolydealError, match="cell_verts out of range"):
        op.upload(bad)
    with pytest.raises(pdl.PolydealError, match="cell_verts out of range"):
        pdl.SIPOperator(bad, keepalive=pah1)


def test_duplicate_interfaces_are_rejected():
    pdl = gpu()
    import ctypes as Cc

    oah, pah = both(2, 4, "blocks2", 1)
    d = pah.flatten()
    A = np.ctypeslib.as_array(d.iface_polyA, (d.n_ifaces,)).copy()
    B = np.ctypeslib.as_array(d.iface_polyB, (d.n_ifaces,)).copy()
    interior = np.nonzero(B >= 0)[0]
    A[interior[1]], B[interior[1]] = A[interior[0]], B[interior[0]]  # the same pair listed twice
    d.iface_polyA = A.ctypes.data_as(Cc.POINTER(Cc.c_int32))
    d.iface_polyB = B.ctypes.data_as(Cc.POINTER(Cc.c_int32))
    with pytest.raises(pdl.PolydealError, match="two interfaces join the same pair"):
        pdl.SIPOperator(d, keepalive=pah)


def test_torch_entry_points_follow_the_current_stream():
    """vmult with torch tensors is ordered with the caller's tensor work on torch's current stream (the handle's
    own stream is non-blocking): produce the source on a side stream right before the apply, no synchronisation."""
    pdl = gpu()
    import torch

    oah, pah = both(3, 8, "blocks4", 2)
    ref = po.assemble_dg_matrix(oah, degree=2, n_threads=4)
    op = pdl.assemble_dg_matrix(pah)
    x = src_vector(op.m())
    yref = ref.vmult(x)
    xh = torch.from_numpy(x).pin_memory()
    for stream in (torch.cuda.Stream(), torch.cuda.default_stream(), torch.cuda.Stream()):
        with torch.cuda.stream(stream):
            big = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")  # keeps the stream busy
            big += 1.0
            xd = torch.zeros(op.m(), dtype=torch.float64, device="cuda")
            xd.copy_(xh, non_blocking=True)
            yd = torch.empty_like(xd)
            op.vmult(yd, xd)
            yh = yd.cpu()  # stream-ordered D2H on the same stream
        assert np.abs(yh.numpy() - yref).max() <= TOL * np.abs(yref).max()


# ----------------------------------------------------------------------------------
# row (b): the reinit() family through the C ABI (include/agglomeration_handler.h:431-452)
# ----------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n,shape,p,nq,distort,fe", [
    (2, 8, "random7", 1, 2, None, "dgq"), (2, 8, "random5", 3, 4, (0.2, 5), "dgq"), (3, 4, "random6", 2, 3, None, "dgq"),
    (3, 4, "blocks2", 3, 4, (0.15, 9), "dgq"), (2, 8, "random6", 2, 3, None, "dgp"), (3, 4, "random5", 2, 3, None, "dgp"),
])
def test_reinit_tables_match_oracle(dim, n, shape, p, nq, distort, fe):
    """reinit(polytope), reinit(polytope, f) and reinit_interface: values, gradients, JxW, points and normals of
    every polytope and every polytope face against the oracle's tables; two-sided alignment of the interface points
    (test/polydeal/reinit_cell_face_quad_pts); agglomerated_quadrature in bbox unit coordinates."""
    pdl = gpu()
    kind = po.FE_AGGLODGP if fe == "dgp" else po.FE_DGQ
    ogrid = po.Grid(dim, n, 0.0, 1.0, 0)
    groups = groups_for(shape, dim, n, ogrid, 3)
    _, oah = oracle_handler(dim, n, groups, p, nq, distort=distort, fe_kind=kind)
    v, cv, nb = oah.grid.arrays()
    pah = pdl.AgglomerationHandler(pdl.Grid.from_arrays(v, cv, nb))
    for g in groups:
        pah.define_agglomerate(g)
    pah.initialize_fe_values(nq)
    pah.distribute_agglomerated_dofs(pdl.FE_AGGLODGP if fe == "dgp" else pdl.FE_DGQ, p)
    op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    tol = 1e-12
    for poly in range(oah.n_polytopes):
        ref, got = oah.reinit(poly), op.reinit(poly)
        assert got.n_q == ref.n_q
        np.testing.assert_allclose(got.points, ref.points, rtol=0, atol=1e-15)
        np.testing.assert_allclose(got.JxW, ref.JxW, rtol=1e-14, atol=0)
        np.testing.assert_allclose(got.values, ref.values, rtol=0, atol=tol * np.abs(ref.values).max())
        np.testing.assert_allclose(got.grads, ref.grads, rtol=0, atol=tol * np.abs(ref.grads).max())
        u, w, x = op.agglomerated_quadrature(poly)
        np.testing.assert_allclose(x, ref.points, rtol=0, atol=1e-15)
        np.testing.assert_allclose(u, pah.real_to_unit(poly, ref.points), rtol=0, atol=1e-15)
        np.testing.assert_allclose(pah.unit_to_real(poly, u), ref.points, rtol=0, atol=1e-14)
        lo, hi = oah.bbox(poly)
        np.testing.assert_allclose(u, (ref.points - lo) / (hi - lo), rtol=0, atol=1e-15)
        np.testing.assert_allclose(w, ref.JxW, rtol=1e-14, atol=0)
        # the element on the unit cell at the unit points reproduces the values (FE_DGQ / FE_AggloDGP evaluators)
        vals, grads = pdl.fe_evaluate(pdl.FE_AGGLODGP if fe == "dgp" else pdl.FE_DGQ, dim, p, u)
        np.testing.assert_allclose(vals, ref.values, rtol=0, atol=tol * np.abs(ref.values).max())
        np.testing.assert_allclose(grads / (hi - lo), ref.grads, rtol=0, atol=tol * np.abs(ref.grads).max())
        for f in range(oah.n_faces(poly)):
            iface, side = pah.face_work_item(poly, f)
            rf, gf = oah.reinit(poly, f), op.reinit_face(iface, side)
            assert gf.n_q == rf.n_q
            if side == 0:  # listed from this polytope: its own sub-face order
                order = np.arange(rf.n_q)
            else:  # the neighbour's list, aligned point by point with the listing side
                nbp, nofn = oah.neighbor(poly, f), oah.neighbor_of_agglomerated_neighbor(poly, f)
                assert pah.face_work_item(nbp, nofn) == (iface, 0)
                other = oah.reinit(nbp, nofn)
                assert np.abs(other.points - rf.points).max() < 1e-15  # reinit_cell_face_quad_pts
                order = np.arange(rf.n_q)
            np.testing.assert_allclose(gf.points, rf.points[order], rtol=0, atol=1e-15)
            np.testing.assert_allclose(gf.normals, rf.normals[order], rtol=0, atol=1e-15)
            np.testing.assert_allclose(gf.JxW, rf.JxW[order], rtol=1e-14, atol=0)
            np.testing.assert_allclose(gf.values, rf.values[:, order], rtol=0, atol=tol * np.abs(rf.values).max())
            np.testing.assert_allclose(gf.grads, rf.grads[:, order], rtol=0, atol=tol * np.abs(rf.grads).max())
            if not oah.at_boundary(poly, f) and side == 0:
                f0, f1 = op.reinit_interface(iface)
                assert np.abs(f0.points - f1.points).max() == 0.0
                assert np.abs(f0.normals + f1.normals).max() == 0.0


def test_reference_style_loop_over_reinit_reproduces_the_assembled_matrix():
    """The loop of PolyUtils::assemble_dg_matrix / examples/poisson.cc:745-905 written against the reinit() tables of
    the library (host side, numpy) gives the matrix pd_assemble computes on the device."""
    pdl = gpu()
    dim, n, p, nq = 2, 8, 2, 3
    oah, pah = both(dim, n, "random6", p, nq=nq, seed=4)
    desc = pah.flatten()
    op = pdl.SIPOperator(desc, keepalive=pah)
    nd, N = pah.n_dofs_per_cell, pah.n_dofs
    A = np.zeros((N, N))
    C = 10.0 * (p + dim) * (p + 1)
    for poly in range(pah.n_polytopes):
        fev = op.reinit(poly)
        dofs = pah.get_dof_indices(poly).astype(np.int64)
        cell = np.einsum("iqd,jqd,q->ij", fev.grads, fev.grads, fev.JxW)
        pen = C / pah.diameter(poly)
        for f in range(pah.n_faces(poly)):
            iface, side = pah.face_work_item(poly, f)
            if pah.at_boundary(poly, f):
                ff = op.reinit_face(iface, side)
                gn = np.einsum("iqd,qd->iq", ff.grads, ff.normals)
                cell += np.einsum("iq,jq,q->ij", -ff.values, gn, ff.JxW) + np.einsum("iq,jq,q->ij", -gn, ff.values, ff.JxW) \
                    + pen * np.einsum("iq,jq,q->ij", ff.values, ff.values, ff.JxW)
                continue
            nbp = pah.neighbor(poly, f)
            if not pah.master_cell(poly) < pah.master_cell(nbp):
                continue
            assert side == 0
            f0, f1 = op.reinit_interface(iface)
            nrm, w = f0.normals, f0.JxW
            g0, g1 = np.einsum("iqd,qd->iq", f0.grads, nrm), np.einsum("iqd,qd->iq", f1.grads, nrm)
            v0, v1 = f0.values, f1.values
            E = lambda a, b: np.einsum("iq,jq,q->ij", a, b, w)
            ndofs = pah.get_dof_indices(nbp).astype(np.int64)
            A[np.ix_(dofs, dofs)] += -0.5 * E(g0, v0) - 0.5 * E(v0, g0) + pen * E(v0, v0)
            A[np.ix_(dofs, ndofs)] += 0.5 * E(g0, v1) - 0.5 * E(v0, g1) - pen * E(v0, v1)
            A[np.ix_(ndofs, dofs)] += -0.5 * E(g1, v0) + 0.5 * E(v1, g0) - pen * E(v1, v0)
            A[np.ix_(ndofs, ndofs)] += 0.5 * E(g1, v1) + 0.5 * E(v1, g1) + pen * E(v1, v1)
        A[np.ix_(dofs, dofs)] += cell
    op.assemble()
    got = op.scipy().toarray()
    assert np.abs(got - A).max() <= 1e-12 * np.abs(A).max()
    ref = po.assemble_dg_matrix(oah, degree=p).scipy().toarray()
    assert np.abs(ref - A).max() <= 1e-12 * np.abs(A).max()


def test_reference_style_cpp_loop_through_the_shim(tmp_path):
    """tests/c_abi/shim_sip_loop.cpp: the SIP loop of examples/poisson.cc written against the header-only C++ shim
    (reinit / reinit_interface / get_dof_indices / LinearOperatorMG-style vmult), compiled here and run on the GPU:
    the host-assembled matrix equals pd_assemble's per entry (1e-12), in 2-D and 3-D."""
    import subprocess

    from test_host_mirror import build_shim_program

    exe = build_shim_program(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHIM LOOP OK" in r.stdout, r.stdout + r.stderr
    print(r.stdout)
