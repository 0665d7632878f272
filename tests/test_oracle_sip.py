"""SIP arithmetic of the CPU oracle, pinned through the reference tests' own
invariants: minimal_SIP_Poisson equality, poisson_sanity_check
energies, exact_solutions exactness on distorted grids (2-D and 3-D)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import pd_scenarios as sc
from oracle import pyoracle as po
from pd_helpers import oracle_handler, src_vector


def handler(dim, n_refine, groups, p, nq, lo=-1.0, hi=1.0, distort=None, fe_kind=po.FE_DGQ):
    grid = po.Grid.hyper_cube(dim, lo, hi, n_refine)
    if distort:
        grid.distort_random(*distort)
    ah = po.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(nq)
    ah.distribute_agglomerated_dofs(fe_kind, p)
    return grid, ah


def interpolate(ah, fun):
    """VectorTools::interpolate on the box mapping: nodal values at the DGQ support
    points of each polytope's bounding box."""
    p = round(ah.n_dofs_per_cell ** (1.0 / ah.dim)) - 1
    nodes = po.gauss_lobatto_nodes(p + 1)
    u = np.zeros(ah.n_dofs)
    for k in range(ah.n_polytopes):
        lo, hi = ah.bbox(k)
        dofs = ah.get_dof_indices(k)
        for i, dof in enumerate(dofs):
            idx = [(i // (p + 1) ** d) % (p + 1) for d in range(ah.dim)]
            x = lo + nodes[idx] * (hi - lo)
            u[dof] = fun(x)
    return u


@pytest.mark.parametrize("dim", [2, 3])
def test_minimal_sip_poisson(dim):
    """test/polydeal/minimal_SIP_Poisson.cc:101,247-253,308,490-508: the matrix on
    2x2-cell (2^3-cell) agglomerates equals the one on the coarser standard grid to 1e-13.
    DGQ1, QGauss(3), penalty 20, h_f = 1, visit by index()."""
    if dim == 2:
        _, ah_a = handler(2, 2, sc.blocks_2x2_of_4x4(), 1, 3)
        _, ah_s = handler(2, 1, [[c] for c in range(4)], 1, 3)
    else:
        _, ah_a = handler(3, 1, [list(range(8))], 1, 3)
        _, ah_s = handler(3, 0, [[0]], 1, 3)
    kw = dict(penalty_constant=20.0, h_rule=po.H_CONSTANT, h_const=1.0, visit_rule=po.VISIT_BY_INDEX)
    A = po.assemble_dg_matrix(ah_a, **kw).scipy().toarray()
    S = po.assemble_dg_matrix(ah_s, **kw).scipy().toarray()
    assert A.shape == S.shape
    assert np.abs(A - S).max() < 1e-13
    assert np.abs(A - A.T).max() < 1e-13


@pytest.mark.parametrize("n_parts", [50, 100, 120])
def test_poisson_sanity_check(n_parts, goldens):
    """test/polydeal/poisson_sanity_check_01.cc:158-164,261-266,420-450: on ANY
    agglomeration of [0,1]^2 64x64 (the reference uses METIS; partition independent),
    boundary terms dropped, penalty 10 max(1/hA,1/hB): x'Ax = 1, (x+y)'A(x+y) = 2, 1'A1 ~ 0."""
    grid = po.Grid.hyper_cube(2, 0.0, 1.0, 6)
    _, _, nbr = grid.arrays()
    groups = sc.random_partition(grid.n_cells, nbr, n_parts, seed=n_parts)
    ah = po.AgglomerationHandler(grid)
    for g in groups:
        ah.define_agglomerate(g)
    ah.initialize_fe_values(3)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
    A = po.assemble_dg_matrix(ah, penalty_constant=10.0, h_rule=po.H_MAX_INVERSE_DIAMETER, with_boundary=False).scipy()
    ux = interpolate(ah, lambda x: x[0])
    uxy = interpolate(ah, lambda x: x[0] + x[1])
    one = np.ones(ah.n_dofs)
    g = goldens["poisson_sanity_check_01"]
    assert ux @ (A @ ux) == pytest.approx(g["x"][0], abs=1e-11)
    assert uxy @ (A @ uxy) == pytest.approx(g["xplusy"][0], abs=1e-11)
    assert abs(one @ (A @ one)) < 1e-11
    assert abs(A - A.T).max() < 1e-12


def test_poisson_sanity_check_02(goldens):
    """test/polydeal/poisson_sanity_check_02.cc (input fully specified): [0,1]^2 in 2x2 cells, agglomerates
    {0,2} and {1,3} (two columns), DGQ1, QGauss(3), boundary terms zeroed, penalty 10 max(1/hA, 1/hB):
    the interpolated step function (a ramp on the left polytope) has energy 2, |x - 1/2| has energy 1."""
    _, ah = handler(2, 1, [[0, 2], [1, 3]], 1, 3, lo=0.0, hi=1.0)
    A = po.assemble_dg_matrix(ah, penalty_constant=10.0, h_rule=po.H_MAX_INVERSE_DIAMETER, with_boundary=False).scipy()
    step = interpolate(ah, lambda x: 0.0 if x[0] < 0.5 else 1.0)
    vfun = interpolate(ah, lambda x: abs(x[0] - 0.5))
    g = goldens["poisson_sanity_check_02"]
    assert step @ (A @ step) == pytest.approx(g["step"][0], abs=1e-13)
    assert vfun @ (A @ vfun) == pytest.approx(g["v"][0], abs=1e-13)


@pytest.mark.parametrize("name", ["distributed_poisson_sanity_check_02", "fully_distributed_poisson_sanity_check_01"])
def test_distributed_poisson_sanity_checks(name, goldens):
    """test/polydeal/distributed_poisson_sanity_check_02.cc (mpirun=3; [0,1]^2 refined 6x, one agglomerate per rank =
    all of its cells, i.e. three contiguous pieces of the Morton curve, :115-156) and
    fully_distributed_poisson_sanity_check_01.cc (256 cells, strips floor(3 x_centre), 10 agglomerates inside every
    strip, :48-66,166-177): DGQ1, QGauss(3), no boundary terms, penalty 1/1 (:230-231 / :249-250).  The golden
    energies 1 and 2 are sums over the ranks, so the serial oracle on the same agglomeration must print them too;
    the sharded GPU assembly is checked against the same goldens in tests/test_gpu_parity.py."""
    g = goldens[name]
    if name == "distributed_poisson_sanity_check_02":
        grid = po.Grid.hyper_cube(2, 0.0, 1.0, 6)
        n = grid.n_cells
        groups = [list(range(n * r // 3, n * (r + 1) // 3)) for r in range(3)]
    else:
        grid = po.Grid.hyper_cube(2, 0.0, 1.0, 4)
        assert grid.n_cells == int(g["n_cells"][0])
        v, cv, nbr = grid.arrays()
        strip = np.floor(v[cv].mean(axis=1)[:, 0] * 3).astype(int)
        groups = []
        for r in range(3):
            cells = np.nonzero(strip == r)[0]
            local = -np.ones(grid.n_cells, dtype=np.int64)
            local[cells] = np.arange(len(cells))
            sub_nbr = np.where(nbr[cells] >= 0, local[np.maximum(nbr[cells], 0)], -1)
            groups += [[int(cells[i]) for i in gr] for gr in sc.random_partition(len(cells), sub_nbr, 10, seed=r)]
        assert len(groups) == 30
    ah = po.AgglomerationHandler(grid)
    for gr in groups:
        ah.define_agglomerate(gr)
    ah.initialize_fe_values(3)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
    A = po.assemble_dg_matrix(ah, penalty_constant=1.0, h_rule=po.H_CONSTANT, h_const=1.0, with_boundary=False).scipy()
    ux = interpolate(ah, lambda x: x[0])
    uxy = interpolate(ah, lambda x: x[0] + x[1])
    one = np.ones(ah.n_dofs)
    assert ux @ (A @ ux) == pytest.approx(g["x"][0], abs=1e-11)
    assert uxy @ (A @ uxy) == pytest.approx(g["xplusy"][0], abs=1e-11)
    assert abs(one @ (A @ one)) < 1e-11


@pytest.mark.parametrize("k", range(6))
def test_poisson_sanity_check_03(k, goldens):
    """test/polydeal/poisson_sanity_check_03.cc:105-116: the same invariants on the UNSTRUCTURED mesh t3.msh (identical
    to test/polydeal/input_grids/square.msh, parsed into tests/golden), read by GridIn and refined three times (5824
    quadrilaterals), split into 50 ... 800 agglomerates (METIS in the reference: the partition is an input, the numbers
    1, 2, ~1e-14 do not depend on it)."""
    g = goldens["poisson_sanity_check_03"]
    n_parts = int(g["n_subdomains"][k])
    mesh = goldens["fully_distributed_poisson_sanity_check_02"]["input_grid"]
    v, cv, nbr = sc.quad_mesh_from_gmsh(mesh["verts"], mesh["quads"], n_refine=3)
    assert len(cv) == 91 * 64
    grid = po.Grid.from_arrays(v, cv, nbr)
    groups = sc.random_partition(grid.n_cells, nbr, n_parts, seed=n_parts)
    ah = po.AgglomerationHandler(grid)
    for gr in groups:
        ah.define_agglomerate(gr)
    ah.initialize_fe_values(3)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
    A = po.assemble_dg_matrix(ah, penalty_constant=10.0, h_rule=po.H_MAX_INVERSE_DIAMETER, with_boundary=False).scipy()
    ux = interpolate(ah, lambda x: x[0])
    uxy = interpolate(ah, lambda x: x[0] + x[1])
    one = np.ones(ah.n_dofs)
    assert ux @ (A @ ux) == pytest.approx(g["x"][k], abs=1e-11)
    assert uxy @ (A @ uxy) == pytest.approx(g["xplusy"][k], abs=1e-11)
    assert abs(one @ (A @ one)) < 1e-10 and abs(g["one"][k]) < 1e-12


def solve_poisson(ah, p, exact, rhs_f, penalty_constant=10.0):
    """The assembly loop of test/polydeal/exact_solutions.cc:400-640 written against the
    oracle's reinit tables (matrix from the oracle's assembler; Dirichlet data and
    forcing added here): penalty = C/diameter(visitor)."""
    A = po.assemble_dg_matrix(ah, penalty_constant=penalty_constant, visit_rule=po.VISIT_BY_INDEX).scipy()
    b = np.zeros(ah.n_dofs)
    vol = bdry = 0.0
    for k in range(ah.n_polytopes):
        fev = ah.reinit(k)
        dofs = ah.get_dof_indices(k)
        vol += fev.JxW.sum()
        f = np.array([rhs_f(x) for x in fev.points])
        b[dofs] += fev.values @ (f * fev.JxW)
        for fc in range(ah.n_faces(k)):
            if ah.at_boundary(k, fc):
                ff = ah.reinit(k, fc)
                bdry += ff.JxW.sum()
                g = np.array([exact(x) for x in ff.points])
                pen = penalty_constant / ah.diameter(k)
                gn = np.einsum("iqd,qd->iq", ff.grads, ff.normals)
                b[dofs] += (pen * ff.values - gn) @ (g * ff.JxW)
    u = spla.spsolve(A.tocsc(), b)
    # L2 and H1-semi errors with the same agglomerated quadrature
    l2 = h1 = 0.0
    eps = 1e-6
    for k in range(ah.n_polytopes):
        fev = ah.reinit(k)
        uk = u[ah.get_dof_indices(k)]
        uh = uk @ fev.values
        gh = np.einsum("i,iqd->qd", uk, fev.grads)
        ue = np.array([exact(x) for x in fev.points])
        ge = np.array([[(exact(x + eps * e) - exact(x - eps * e)) / (2 * eps) for e in np.eye(ah.dim)] for x in fev.points])
        l2 += ((uh - ue) ** 2 * fev.JxW).sum()
        h1 += (((gh - ge) ** 2).sum(axis=1) * fev.JxW).sum()
    return vol, bdry, np.sqrt(l2), np.sqrt(h1)


@pytest.mark.parametrize("kind", ["linear", "quadratic"])
def test_exact_solutions_distorted_2d(kind):
    """test/polydeal/exact_solutions.cc:31,296-311,470-476,552-555,638-648: randomly
    distorted 4x4 grid in 2x2 blocks; u = x+y-1 (DGQ1) and x^2+y^2-1 (DGQ2) are reproduced
    to 1e-14-ish; sum of volume JxW = 1, boundary JxW = 4.  (deal.II's distort_random
    stream is not reproducible, so a seeded distortion of our own is used.)"""
    p = 1 if kind == "linear" else 2
    exact = (lambda x: x[0] + x[1] - 1.0) if p == 1 else (lambda x: x[0] ** 2 + x[1] ** 2 - 1.0)
    rhs = (lambda x: 0.0) if p == 1 else (lambda x: -4.0)
    _, ah = handler(2, 2, sc.blocks_2x2_of_4x4(), p, 2 * p + 1, lo=0.0, hi=1.0, distort=(0.25, 7))
    vol, bdry, l2, h1 = solve_poisson(ah, p, exact, rhs)
    assert vol == pytest.approx(1.0, abs=1e-14)
    assert bdry == pytest.approx(4.0, abs=1e-14)
    assert l2 < 1e-13
    assert h1 < 1e-7  # finite-difference gradient of the exact solution limits this one


@pytest.mark.parametrize("kind", ["linear", "quadratic"])
def test_exact_solutions_dgp(kind):
    """test/polydeal/exact_solutions_dgp.cc:26,278-283,304-306,321-348: the same problem with
    FE_AggloDGP(1) / FE_AggloDGP(2) (Legendre basis of total degree p on the bounding boxes): "Linear: OK",
    "Quadratic: OK" = volume 1, boundary 4, L2 and H1 errors below 1e-14."""
    p = 1 if kind == "linear" else 2
    exact = (lambda x: x[0] + x[1] - 1.0) if p == 1 else (lambda x: x[0] ** 2 + x[1] ** 2 - 1.0)
    rhs = (lambda x: 0.0) if p == 1 else (lambda x: -4.0)
    _, ah = handler(2, 2, sc.blocks_2x2_of_4x4(), p, 2 * p + 1, lo=0.0, hi=1.0, distort=(0.25, 7), fe_kind=po.FE_AGGLODGP)
    assert ah.n_dofs_per_cell == (3 if p == 1 else 6)
    vol, bdry, l2, h1 = solve_poisson(ah, p, exact, rhs)
    assert vol == pytest.approx(1.0, abs=1e-14) and bdry == pytest.approx(4.0, abs=1e-14)
    assert l2 < 1e-13 and h1 < 1e-7


@pytest.mark.parametrize("kind", ["linear", "quadratic"])
def test_disconnected_exact_solution(kind):
    """test/polydeal/disconnected_exact_solution.cc:343-377: define_agglomerate_with_check splits
    {0,1,2,3,12,13,14,15} and {4,...,10} into their face-connected components (the two blocks of the second
    set touch in a corner only), {11} stays: "Number of generated agglomerates: 5", one of them L-shaped;
    linear / quadratic solutions are still reproduced (continuous_face_exact_solution.cc is the same problem
    on the four 2x2 blocks, covered by test_exact_solutions_distorted_2d)."""
    grid = po.Grid.hyper_cube(2, 0.0, 1.0, 2)
    nbr = grid.arrays()[2]

    def components(cells):  # source/agglomeration_handler.cc:174-207
        cells, seen, out = list(cells), set(), []
        for c in cells:
            if c in seen:
                continue
            comp, stack = [], [c]
            seen.add(c)
            while stack:
                a = stack.pop()
                comp.append(a)
                for b in nbr[a]:
                    if b in cells and b not in seen:
                        seen.add(int(b))
                        stack.append(int(b))
            out.append(sorted(comp))
        return out

    groups = components([0, 1, 2, 3, 12, 13, 14, 15]) + components([4, 5, 6, 7, 8, 9, 10]) + components([11])
    assert len(groups) == 5 and [8, 9, 10] in groups
    p = 1 if kind == "linear" else 2
    exact = (lambda x: x[0] + x[1] - 1.0) if p == 1 else (lambda x: x[0] ** 2 + x[1] ** 2 - 1.0)
    rhs = (lambda x: 0.0) if p == 1 else (lambda x: -4.0)
    _, ah = handler(2, 2, groups, p, 2 * p + 1, lo=0.0, hi=1.0, distort=(0.25, 11))
    vol, bdry, l2, h1 = solve_poisson(ah, p, exact, rhs)
    assert vol == pytest.approx(1.0, abs=1e-14) and bdry == pytest.approx(4.0, abs=1e-14)
    assert l2 < 1e-13 and h1 < 1e-7


@pytest.mark.parametrize("dim", [2, 3])
def test_exact_solutions_3d_invariants(dim):
    """test/polydeal/exact_solutions_distributed.cc:45,182-211,299,669 run serially:
    [0,1]^d refined 3x, DGQ1, penalty 10/h: vol = 1, boundary measure = 2d, linear
    solution exact."""
    n = 8
    groups = sc.block_partition(dim, n, 2)
    _, ah = handler(dim, 3, groups, 1, 3, lo=0.0, hi=1.0)
    exact = (lambda x: x.sum() - 1.0)
    vol, bdry, l2, _ = solve_poisson(ah, 1, exact, lambda x: 0.0)
    assert vol == pytest.approx(1.0, abs=1e-13)
    assert bdry == pytest.approx(2.0 * dim, abs=1e-13)
    assert l2 < 1e-12


def test_threaded_assembly_is_consistent():
    grid = po.Grid.hyper_cube(3, 0.0, 1.0, 3)
    ah = po.AgglomerationHandler(grid)
    for g in sc.block_partition(3, 8, 2):
        ah.define_agglomerate(g)
    ah.initialize_fe_values(3)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 2)
    a1 = po.assemble_dg_matrix(ah, degree=2, n_threads=1).values()
    a4 = po.assemble_dg_matrix(ah, degree=2, n_threads=4).values()
    assert np.abs(a1 - a4).max() <= 1e-12 * np.abs(a1).max()


def test_fine_mesh_penalty_rule_matches_monodomain_emulation():
    """include/utils.h:861-866,906-909 on Cartesian singletons: sigma_F = p(p+1)(1/h_m+1/h_p),
    boundary 4 p(p+1)/h -- same SIP form, so constants are in the null space without boundary
    terms and the matrix is SPD with them."""
    p = 2
    grid = po.Grid.hyper_cube(3, 0.0, 1.0, 2)
    ah = po.AgglomerationHandler(grid)
    for c in range(grid.n_cells):
        ah.define_agglomerate([c])
    ah.initialize_fe_values(p + 1)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, p)
    A0 = po.assemble_dg_matrix(ah, penalty_constant=p * (p + 1.0), h_rule=po.H_NORMAL_EXTENT, with_boundary=False).scipy()
    assert np.abs(A0 @ np.ones(ah.n_dofs)).max() < 1e-11
    A = po.assemble_dg_matrix(ah, penalty_constant=p * (p + 1.0), h_rule=po.H_NORMAL_EXTENT).scipy()
    assert abs(A - A.T).max() < 1e-12
    w = np.linalg.eigvalsh(A.toarray())
    assert w.min() > 0


def poisson_golden_problem():
    """test/polydeal/poisson.cc:107,122-123,153-232,316,357: [-1,1]^2 refined 6x, seven 2-cell
    agglomerates + singletons, DGQ1, QGauss(3), penalty 20 / h_f with h_f = measure of a cell
    face = 2/64, visited by index(), f = 8 pi^2 sin(2 pi x) sin(2 pi y), homogeneous Dirichlet."""
    groups = sc.polytope_iterator_agglomerates()
    grid, ah = handler(2, 6, groups, 1, 3, lo=-1.0, hi=1.0)
    kw = dict(penalty_constant=20.0, h_rule=po.H_CONSTANT, h_const=2.0 / 64, visit_rule=po.VISIT_BY_INDEX)
    return grid, ah, kw


def poisson_rhs_and_error(grid, ah):
    """RHS vector and the error functional of the reference test: the polytopal solution is
    interpolated to the fine DGQ1 space (exact, it is bilinear on every sub-cell) and
    integrate_difference runs with QGauss(degree) = ONE point per cell (poisson.cc:output_results)."""
    f = lambda x: 8 * np.pi**2 * np.sin(2 * np.pi * x[:, 0]) * np.sin(2 * np.pi * x[:, 1])
    b = np.zeros(ah.n_dofs)
    for k in range(ah.n_polytopes):
        fev = ah.reinit(k)
        b[ah.get_dof_indices(k)] += fev.values @ (f(fev.points) * fev.JxW)

    def l2_error(u):
        err2 = 0.0
        for k in range(ah.n_polytopes):
            lo, hi = ah.bbox(k)
            uk = u[ah.get_dof_indices(k)]
            for c in ah.get_agglomerate(k):
                v = grid.cell_vertices(int(c))
                xc = v.mean(axis=0)
                area = (v[:, 0].max() - v[:, 0].min()) * (v[:, 1].max() - v[:, 1].min())
                phi, _ = po.fe_evaluate(po.FE_DGQ, 2, 1, (xc - lo) / (hi - lo))
                exact = np.sin(2 * np.pi * xc[0]) * np.sin(2 * np.pi * xc[1])
                err2 += (uk @ phi - exact) ** 2 * area
        return np.sqrt(err2)

    return b, l2_error


def test_poisson_golden_l2_error(goldens):
    """The one end-to-end NUMBER the reference pins for this path: L2 error 0.00647702
    (test/polydeal/poisson.output).  Matrix from the oracle, direct solve."""
    grid, ah, kw = poisson_golden_problem()
    A = po.assemble_dg_matrix(ah, n_threads=4, **kw).scipy().tocsc()
    b, l2_error = poisson_rhs_and_error(grid, ah)
    u = spla.spsolve(A, b)
    assert l2_error(u) == pytest.approx(goldens["poisson"][0], abs=5e-9)  # golden printed with 6 digits


@pytest.mark.parametrize("dim,n,p", [(2, 4, 1), (2, 4, 2), (2, 3, 3), (3, 2, 1), (3, 3, 2)])
def test_mapped_fine_operator_equals_polytope_assembly_on_cartesian_cells(dim, n, p):
    """The oracle's mapped-basis fine-mesh operator (unpinned by reference tests) against the
    pinned polytope assembly: on Cartesian cells with one cell per polytope and the
    normal-extent penalty max(p,1)(p+1)(1/h_m + 1/h_p) they are the same matrix."""
    grid, ah = oracle_handler(dim, n, [[c] for c in range(n**dim)], p, p + 1, order=1)
    x = src_vector(ah.n_dofs)
    pc = max(p, 1) * (p + 1.0)
    y = po.mapped_fine_vmult(grid, p, p + 1, x, stiffness=1.3, mass=0.7)
    M = po.assemble_dg_matrix(ah, penalty_constant=pc, h_rule=po.H_NORMAL_EXTENT, stiffness_coeff=1.3, mass_coeff=0.7)
    y2 = M.vmult(x)
    assert np.abs(y - y2).max() <= 1e-13 * np.abs(y2).max()


@pytest.mark.parametrize("dim,n,p", [(2, 5, 2), (3, 3, 1), (3, 3, 2)])
def test_mapped_fine_operator_invariants_on_distorted_cells(dim, n, p):
    """Symmetry, constants in the kernel of the interior part, positivity (SIP with the
    reference's penalty is coercive), mass = volume."""
    grid = po.Grid(dim, n, 0.0, 1.0, 1)
    grid.distort_random(0.25, 4)
    N = grid.n_cells * (p + 1) ** dim
    rng = np.random.default_rng(0)
    x, z = rng.standard_normal(N), rng.standard_normal(N)
    A = lambda v, **kw: po.mapped_fine_vmult(grid, p, p + 1, v, **kw)
    assert abs(z @ A(x) - x @ A(z)) <= 1e-12 * abs(z @ A(x))
    one = np.ones(N)
    assert np.abs(A(one, boundary=False)).max() <= 1e-12
    assert x @ A(x) > 0
    vol = one @ po.mapped_fine_vmult(grid, p, p + 1, one, stiffness=0.0, mass=1.0, boundary=False, interior=False)
    assert abs(vol - 1.0) <= 1e-13


def _unstructured_square(goldens):
    """the input grid of fully_distributed_poisson_sanity_check_02.cc:124-129 (input_grids/square.msh, 91 quads,
    parsed into tests/golden by make_golden.py), refined once: 364 cells, neighbours rotated against each other"""
    g = goldens["fully_distributed_poisson_sanity_check_02"]
    v, cv, nbr = sc.quad_mesh_from_gmsh(g["input_grid"]["verts"], g["input_grid"]["quads"], n_refine=1)
    assert len(cv) == int(g["n_cells"][0])
    # it IS unstructured: some neighbour does not see the cell through the opposite face
    assert any(nbr[nbr[c, f], f ^ 1] != c for c in range(len(cv)) for f in range(4) if nbr[c, f] >= 0)
    return v, cv, nbr


def _rank_then_agglomerates(nbr, n_ranks, n_local, seed=0):
    """GridTools::partition_triangulation(n_ranks) then PolyUtils::partition_locally_owned_regions(n_local): both
    METIS in the reference (partition = input of the path); here seeded graph growing on the same graphs"""
    rank_groups = sc.random_partition(len(nbr), nbr, n_ranks, seed=seed)
    groups, owner = [], []
    for r, cells in enumerate(rank_groups):
        cells = np.array(sorted(cells))
        local = -np.ones(len(nbr), dtype=np.int64)
        local[cells] = np.arange(len(cells))
        sub_nbr = np.where(nbr[cells] >= 0, local[np.maximum(nbr[cells], 0)], -1)
        for gr in sc.random_partition(len(cells), sub_nbr, n_local, seed=seed + 1 + r):
            groups.append([int(cells[i]) for i in gr])
            owner.append(r)
    return groups, owner


def test_fully_distributed_poisson_sanity_check_02_unstructured(goldens):
    """test/polydeal/fully_distributed_poisson_sanity_check_02.cc (mpirun=3): the sanity-check energies on an
    UNSTRUCTURED quadrilateral mesh read by GridIn (square.msh refined once = 364 cells), three ranks, ten
    agglomerates per rank, DGQ1, QGauss(3), no boundary terms, penalty 1/1: x'Ax = 1, (x+y)'A(x+y) = 2."""
    g = goldens["fully_distributed_poisson_sanity_check_02"]
    v, cv, nbr = _unstructured_square(goldens)
    groups, owner = _rank_then_agglomerates(nbr, 3, 10)
    assert len(groups) == 30
    grid = po.Grid.from_arrays(v, cv, nbr)
    ah = po.AgglomerationHandler(grid)
    for gr in groups:
        ah.define_agglomerate(gr)
    ah.initialize_fe_values(3)
    ah.distribute_agglomerated_dofs(po.FE_DGQ, 1)
    # volume and boundary quadrature see the unit square
    assert sum(ah.reinit(k).JxW.sum() for k in range(ah.n_polytopes)) == pytest.approx(1.0, abs=1e-14)
    assert sum(ah.reinit(k, f).JxW.sum() for k in range(ah.n_polytopes) for f in range(ah.n_faces(k))
               if ah.at_boundary(k, f)) == pytest.approx(4.0, abs=1e-14)
    # the two sides of every interface meet in the same quadrature points (test/polydeal/reinit_cell_face_quad_pts.cc)
    for k in range(ah.n_polytopes):
        for f in range(ah.n_faces(k)):
            if not ah.at_boundary(k, f) and k < ah.neighbor(k, f):
                q = ah.neighbor(k, f)
                a, b = ah.reinit_interface(k, q, f, ah.neighbor_of_agglomerated_neighbor(k, f))
                assert np.abs(a.points - b.points).max() < 1e-15 and np.abs(a.normals + b.normals).max() < 1e-14
    A = po.assemble_dg_matrix(ah, penalty_constant=1.0, h_rule=po.H_CONSTANT, h_const=1.0, with_boundary=False).scipy()
    ux = interpolate(ah, lambda x: x[0])
    uxy = interpolate(ah, lambda x: x[0] + x[1])
    one = np.ones(ah.n_dofs)
    assert ux @ (A @ ux) == pytest.approx(g["x"][0], abs=1e-11)
    assert uxy @ (A @ uxy) == pytest.approx(g["xplusy"][0], abs=1e-11)
    assert abs(one @ (A @ one)) < 1e-11 and abs(A - A.T).max() < 1e-12
