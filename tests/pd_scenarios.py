"""Mesh + agglomeration scenarios taken from the reference's own tests
(/root/reference/test/polydeal/*.cc); shared by the oracle tests and the
product parity tests so both are driven with identical inputs."""
from __future__ import annotations

import numpy as np


def standard_8x8_agglomerates():
    """[-1,1]^2 refined 3x; {3,6,9,12,13},{15,36,37},{57,60,54},{25,19,22} + singletons
    (test/polydeal/agglomerated_neighbors_01.cc:30-107, sparsity_agglomerated_tria.cc:30-107,
    fe_space_on_bbox.cc, reinit_cell_face_01.cc).  collect_cells_for_agglomeration sorts
    into active-cell order (include/poly_utils.h:532-538)."""
    groups = [[3, 6, 9, 12, 13], [15, 36, 37], [57, 60, 54], [25, 19, 22]]
    return _with_singletons(groups, 64)


def _with_singletons(groups, n_cells):
    flagged = {c for g in groups for c in g}
    out = [sorted(g) for g in groups]
    out += [[c] for c in range(n_cells) if c not in flagged]
    return out


def blocks_2x2_of_4x4():
    """[-1,1]^2 refined 2x, {0..3},{4..7},{8..11},{12..15}
    (agglomerated_neighbors_02.cc, hp_structure_01.cc, minimal_SIP_Poisson.cc)."""
    return [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12, 13, 14, 15]]


def polytope_iterator_agglomerates():
    """[-1,1]^2 refined 6x with seven 2-cell agglomerates (polytope_iterator.cc:142-230,
    same sets as test/polydeal/poisson.cc:153-232)."""
    groups = [[3235, 3238], [831, 874], [1226, 1227], [2279, 2278], [3760, 3761], [3648, 3306], [3765, 3764]]
    return _with_singletons(groups, 4096)


def block_partition(dim, n, b, order=0):
    """Uniform b^dim blocks of an n^dim structured grid (the `blocks` shape of
    SURVEY 8d): list of cell lists, cells in active-cell order."""
    import numpy as np

    idx = np.arange(n**dim)
    if dim == 2:
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        k = np.zeros_like(i)
    else:
        i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    if order == 0:
        levels = n.bit_length() - 1
        cell = np.zeros_like(i)
        for l in range(levels):
            cell |= ((i >> l) & 1) << (dim * l)
            cell |= ((j >> l) & 1) << (dim * l + 1)
            if dim == 3:
                cell |= ((k >> l) & 1) << (dim * l + 2)
    else:
        cell = (k * n + j) * n + i
    nb = n // b
    part = ((k // b) * nb + (j // b)) * nb + (i // b)
    groups = [[] for _ in range(nb**dim)]
    order_ = np.argsort(cell)
    for c, p in zip(cell[order_], part[order_]):
        groups[p].append(int(c))
    del idx
    return groups


def random_partition(n_cells, nbr, n_parts, seed):
    """Connected random agglomerates by seeded region growing on the face-adjacency
    graph -- stand-in for METIS partitions (inputs, never outputs)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    part = -np.ones(n_cells, dtype=np.int64)
    seeds = rng.choice(n_cells, size=n_parts, replace=False)
    frontier = [[int(s)] for s in seeds]
    for p, s in enumerate(seeds):
        part[s] = p
    remaining = n_cells - n_parts
    while remaining > 0:
        progressed = False
        for p in rng.permutation(n_parts):
            fr = frontier[p]
            while fr:
                c = fr[int(rng.integers(len(fr)))]
                free = [int(x) for x in nbr[c] if x >= 0 and part[x] < 0]
                if not free:
                    fr.remove(c)
                    continue
                x = free[int(rng.integers(len(free)))]
                part[x] = p
                fr.append(x)
                remaining -= 1
                progressed = True
                break
        if not progressed:
            break
    assert remaining == 0
    groups = [[] for _ in range(n_parts)]
    for c in range(n_cells):
        groups[part[c]].append(c)
    return groups


# ---- more fixtures of the reference's face tests (inputs fully specified in the .cc files) ----
def continuous_face_02_cases():
    """test/polydeal/continuous_face_02.cc test0/test1/test2 on [-1,1]^2 refined twice."""
    return [
        [list(range(0, 8)), list(range(8, 12)), [12], [13], [14], [15]],
        [list(range(0, 8)), list(range(8, 12)), list(range(12, 16))],
        [[3, 6, 9, 12], [0, 1, 4, 5], [2, 8, 10], [11, 14, 15], [7, 13]],
    ]


def continuous_face_03_groups():
    """test/polydeal/continuous_face_03.cc: 8x8 cells; every unflagged cell first (active order), then
    {36,37,38,39}, {18,24,25}, {3,6}."""
    agg = [[36, 37, 38, 39], [18, 24, 25], [3, 6]]
    flagged = {c for g in agg for c in g}
    return [[c] for c in range(64) if c not in flagged] + agg


def continuous_face_distorted_cases():
    """test/polydeal/continuous_face_distorted_grid.cc test0/test1 (4x4 cells, distort_random(0.25))."""
    return [[list(range(0, 8)), list(range(8, 16))], [list(range(0, 4)), list(range(4, 8)), list(range(8, 12)), list(range(12, 16))]]


def reinit_cell_face_master_master_groups():
    return [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11], [12], [13], [14], [15]]


def reinit_cell_face_quad_pts_groups():
    """{3,6,9}, {36,37}, {25,19} first, then every other cell of the 8x8 grid in active order."""
    agg = [[3, 6, 9], [36, 37], [19, 25]]
    flagged = {c for g in agg for c in g}
    return agg + [[c] for c in range(64) if c not in flagged]


def partition_from_continuous_face_golden(scenario, n_cells, nbr):
    """Recover the (METIS) partition a continuous_face scenario was run on from its own golden: every
    cell that touches an interface is listed on its polytope's side; the remaining cells have all their
    neighbours in their own polytope, so a flood fill assigns them.  Returns groups in polytope order
    (cells ascending = active order), or None if the golden does not determine the partition."""
    owner = [-1] * n_cells
    for p, poly in enumerate(scenario["polytopes"]):
        owner[poly["master"]] = p
        for face in poly["faces"]:
            for cell, _lf, _nm in face["subfaces"]:
                if owner[cell] not in (-1, p):
                    return None
                owner[cell] = p
    changed = True
    while changed:
        changed = False
        for c in range(n_cells):
            if owner[c] >= 0:
                continue
            cand = {owner[n] for n in nbr[c] if n >= 0 and owner[n] >= 0}
            # an unlisted cell is interior to its polytope: any assigned neighbour that is itself unlisted-or-
            # same-polytope decides; if two different polytopes touch it, it would have been listed
            if len(cand) == 1:
                owner[c] = cand.pop()
                changed = True
    if min(owner) < 0:
        return None
    groups = [[] for _ in scenario["polytopes"]]
    for c, p in enumerate(owner):
        groups[p].append(c)
    return groups


def quad_mesh_from_gmsh(verts, quads_ccw, n_refine=0):
    """(verts, cell_verts, nbr) of an unstructured quadrilateral mesh the way deal.II's GridIn + refine_global hand
    it over: gmsh lists a quad counter-clockwise, deal.II lexicographically with consistently directed edges; faces
    0: x-, 1: x+, 2: y-, 3: y+ of the reference cell; children of a refined cell in the order of its vertices.
    Neighbouring cells are in general rotated against each other (no opposite-face rule), but an edge has the same
    direction seen from both sides, as in every 2-D deal.II triangulation."""
    verts = [tuple(map(float, v)) for v in verts]
    # GridTools::consistently_order_cells (what GridIn does in 2-D): every edge gets ONE direction, the same seen
    # from both of its cells, and the opposite edges of a quad are parallel; found by walking the chains of
    # opposite edges.  direction[(a, b)] = True means a -> b for the edge {a, b}, a < b.
    cells_of_edge = {}
    for qi, q in enumerate(quads_ccw):
        for k in range(4):
            a, b = q[k], q[(k + 1) % 4]
            cells_of_edge.setdefault((min(a, b), max(a, b)), []).append(qi)
    direction = {}
    for start in cells_of_edge:
        if start in direction:
            continue
        direction[start] = True
        stack = [start]
        while stack:
            e = stack.pop()
            tail, head = e if direction[e] else e[::-1]
            for qi in cells_of_edge[e]:
                q = quads_ccw[qi]
                k = next(k for k in range(4) if {q[k], q[(k + 1) % 4]} == set(e))
                # the opposite edge runs q[k+3] -> q[k+2] when this one runs q[k] -> q[k+1]
                o_tail, o_head = (q[(k + 3) % 4], q[(k + 2) % 4]) if q[k] == tail else (q[(k + 2) % 4], q[(k + 3) % 4])
                o = (min(o_tail, o_head), max(o_tail, o_head))
                if o not in direction:
                    direction[o] = o_tail < o_head
                    stack.append(o)
                else:
                    assert direction[o] == (o_tail < o_head), "mesh is not orientable"
    starts_at = lambda a, b: direction[(min(a, b), max(a, b))] == (a < b)
    # a file whose cells already are consistently directed keeps its vertex order (deal.II reorders only if it must)
    as_given, consistent = {}, True
    for q in quads_ccw:
        for a, b in ((q[0], q[3]), (q[1], q[2]), (q[0], q[1]), (q[3], q[2])):  # lines 0..3 of (n0 n1 n3 n2)
            consistent = consistent and as_given.setdefault((min(a, b), max(a, b)), a < b) == (a < b)
    if consistent:
        starts_at = lambda a, b: as_given[(min(a, b), max(a, b))] == (a < b)
    cells = []
    for q in quads_ccw:
        k = next(k for k in range(4) if starts_at(q[k], q[(k + 1) % 4]) and starts_at(q[k], q[(k + 3) % 4]))
        cells.append([q[k], q[(k + 1) % 4], q[(k + 3) % 4], q[(k + 2) % 4]])  # lexicographic, counter-clockwise = positive
    for _ in range(n_refine):
        mid, new_cells = {}, []

        def midpoint(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                verts.append(tuple(0.5 * (verts[a][k] + verts[b][k]) for k in range(2)))
                mid[key] = len(verts) - 1
            return mid[key]

        for v0, v1, v2, v3 in cells:
            m01, m02, m13, m23 = midpoint(v0, v1), midpoint(v0, v2), midpoint(v1, v3), midpoint(v2, v3)
            verts.append(tuple(0.25 * (verts[v0][k] + verts[v1][k] + verts[v2][k] + verts[v3][k]) for k in range(2)))
            c = len(verts) - 1
            new_cells += [[v0, m01, m02, c], [m01, v1, c, m13], [m02, c, v2, m23], [c, m13, m23, v3]]
        cells = new_cells
    face_verts = [(0, 2), (1, 3), (0, 1), (2, 3)]
    edge = {}
    for ci, cv in enumerate(cells):
        for f, (a, b) in enumerate(face_verts):
            edge.setdefault((min(cv[a], cv[b]), max(cv[a], cv[b])), []).append((ci, f))
    nbr = -np.ones((len(cells), 4), dtype=np.int32)
    for sides in edge.values():
        assert len(sides) <= 2
        if len(sides) == 2:
            (c0, f0), (c1, f1) = sides
            nbr[c0, f0], nbr[c1, f1] = c1, c0
    return np.array(verts), np.array(cells, dtype=np.int32), nbr


def hyper_ball_2d(radius=1.0, n_refine=1):
    """GridGenerator::hyper_ball<2>(tria, {}, radius) (the five-cell disc, inner square of half width
    radius / (sqrt 2 (1 + sqrt 2)), SphericalManifold on the boundary) followed by refine_global(n_refine), as in
    test/polydeal/unstructured_grid.cc:30-32 and fe_collection_agglomeration.cc:119-122.  deal.II behaviours restated
    (pinned by the first test's golden): the new vertex of a boundary line lies on the circle, the one of an interior
    line at its mid-point, the one of a cell at the transfinite interpolation of its 4 vertices (weight -1/4) and 4
    new line vertices (+1/2); children in the order of their parents (all leaves on one level: the active-cell order)."""
    a, r = 1.0 / (1.0 + np.sqrt(2.0)), radius / np.sqrt(2.0)
    verts = [(-r, -r), (r, -r), (-r * a, -r * a), (r * a, -r * a), (-r * a, r * a), (r * a, r * a), (-r, r), (r, r)]
    cells = [[0, 1, 2, 3], [0, 2, 6, 4], [2, 3, 4, 5], [1, 7, 3, 5], [6, 4, 7, 5]]
    face_verts = [(0, 2), (1, 3), (0, 1), (2, 3)]
    for _ in range(n_refine):
        n_cells_at = {}
        for c in cells:
            for i, j in face_verts:
                key = (min(c[i], c[j]), max(c[i], c[j]))
                n_cells_at[key] = n_cells_at.get(key, 0) + 1
        mid = {}

        def midpoint(p, q):
            key = (min(p, q), max(p, q))
            if key not in mid:
                m = 0.5 * (np.array(verts[p]) + np.array(verts[q]))
                if n_cells_at[key] == 1:  # boundary line: SphericalManifold
                    m = m * (radius / np.linalg.norm(m))
                verts.append(tuple(m))
                mid[key] = len(verts) - 1
            return mid[key]

        children = []
        for v0, v1, v2, v3 in cells:
            m01, m02, m13, m23 = midpoint(v0, v1), midpoint(v0, v2), midpoint(v1, v3), midpoint(v2, v3)
            P = lambda i: np.array(verts[i])
            verts.append(tuple(-0.25 * (P(v0) + P(v1) + P(v2) + P(v3)) + 0.5 * (P(m01) + P(m02) + P(m13) + P(m23))))
            c = len(verts) - 1
            children += [[v0, m01, m02, c], [m01, v1, c, m13], [m02, c, v2, m23], [c, m13, m23, v3]]
        cells = children
    edge = {}
    for ci, cv in enumerate(cells):
        for f, (i, j) in enumerate(face_verts):
            edge.setdefault((min(cv[i], cv[j]), max(cv[i], cv[j])), []).append((ci, f))
    nbr = -np.ones((len(cells), 4), dtype=np.int32)
    for sides in edge.values():
        if len(sides) == 2:
            (c0, f0), (c1, f1) = sides
            nbr[c0, f0], nbr[c1, f1] = c1, c0
    return np.array(verts), np.array(cells, dtype=np.int32), nbr


def hyper_ball_2d_refined_once():
    return hyper_ball_2d(1.0, 1)
