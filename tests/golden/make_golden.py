#!/usr/bin/env python3
"""Derive tests/golden/reference_goldens.json from the reference's own test goldens.

Run in the build container only (it reads /root/reference/test/polydeal/*.output,
which does not exist on the GPU box).  The .output files are parsed into
structured records (polytopes, faces, (cell, face) lists, DoF indices, sparsity
rows, invariants); tests/test_oracle_goldens.py rebuilds each scenario with the
CPU oracle and compares against these records.  Nothing from the reference's
sources is copied; only the numbers its tests print.
"""
import json
import os
import re
import sys

REF = "/root/reference/test/polydeal"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")


def lines(name):
    with open(os.path.join(REF, name)) as f:
        return [l.rstrip("\n") for l in f]


def parse_sparsity(name):
    rows = []
    for l in lines(name):
        nums = [int(t) for t in l.strip("[]").split(",")]
        rows.append(nums)  # [row, col0, col1, ...]
    return rows


def parse_neighbors_faces(name):
    polys, cur, face = [], None, None
    pend_cell = None
    for l in lines(name):
        if m := re.match(r"Polytope with idx: (\d+)", l):
            cur = {"index": int(m[1]), "n_faces": None, "faces": {}}
            polys.append(cur)
        elif m := re.match(r"Number of faces for the agglomeration: (\d+)", l):
            cur["n_faces"] = int(m[1])
        elif m := re.match(r"Agglomerated face with idx: (\d+)", l):
            face = cur["faces"].setdefault(m[1], [])
        elif m := re.match(r"deal.II cell idx: (\d+)", l):
            pend_cell = int(m[1])
        elif m := re.match(r"deal.II face idx: (\d+)", l):
            face.append([pend_cell, int(m[1])])
    return polys


def parse_nofn(name):
    out, cur = [], None
    for l in lines(name):
        if m := re.match(r"Cell with idx: (\d+)", l):
            cur = {"master": int(m[1]), "nofn": []}
            out.append(cur)
        elif m := re.match(r"Number of faces for this cell: (\d+)", l):
            cur["n_faces"] = int(m[1])
        elif m := re.match(r"Neighbor of neighbor for \((\d+),(\d+)\) = (\d+)", l):
            cur["nofn"].append(int(m[3]))
    return out


def parse_continuous_face(name):
    scenarios, polys, cur, face = [], [], None, None
    pend = {}
    for l in lines(name):
        if m := re.match(r"Master cell index = (\d+)", l):
            cur = {"master": int(m[1]), "faces": []}
            polys.append(cur)
        elif m := re.match(r"Number of agglomerated faces = (\d+)", l):
            cur["n_faces"] = int(m[1])
        elif m := re.match(r"Agglomerate face index = (\d+)", l):
            face = {"f": int(m[1]), "neighbor": None, "nofn": None, "subfaces": []}
            cur["faces"].append(face)
        elif m := re.match(r"Neighbor(?: polytope index)? = (\d+)", l):
            face["neighbor"] = int(m[1])
        elif m := re.match(r"Neighbor of neighbor = (\d+)", l):
            face["nofn"] = int(m[1])
        elif m := re.match(r"deal.II cell index = (\d+)", l):
            pend = {"cell": int(m[1])}
        elif m := re.match(r"Local face idx = (\d+)", l):
            pend["face"] = int(m[1])
        elif m := re.match(r"Neighboring master cell index = (\d+)", l):
            face["subfaces"].append([pend["cell"], pend["face"], int(m[1])])
        elif m := re.match(r"Perimeter = (\S+)", l):
            scenarios.append({"polytopes": polys, "perimeter": float(m[1])})
            polys = []
    return scenarios


def parse_hp_structure(name):
    out, cur, mode = [], None, None
    for l in lines(name):
        if m := re.match(r"Cell with global index: (\d+) has global DoF indices:", l):
            cur = {"master": int(m[1]), "dofs": [], "vertices": []}
            out.append(cur)
            mode = "dofs"
        elif l.strip() == "and vertices:":
            mode = "verts"
        elif l.strip():
            if mode == "dofs":
                cur["dofs"].append(int(l))
            else:
                cur["vertices"].append([float(t) for t in l.split()])
    return out


def parse_reinit_cell_face_02(name):
    out, cur = [], None
    for l in lines(name):
        if m := re.match(r"Cell with index (\d+) has (\d+) faces", l):
            cur = {"master": int(m[1]), "n_faces": int(m[2]), "faces": []}
            out.append(cur)
        elif m := re.match(r"Neighbor index= (\d+)", l):
            cur["faces"].append(int(m[1]))  # master cell index of neighbour
        elif re.match(r"Face with idx: (\d+) is a boundary face", l):
            cur["faces"].append(-1)
    return out


def parse_polytope_iterator(name):
    out, cur = [], None
    for l in lines(name):
        if l.startswith("Looping backwards"):
            break
        if m := re.match(r"Global DoF indices for polytope (\d+)", l):
            cur = {"index": int(m[1]), "dofs": []}
            out.append(cur)
        elif re.match(r"^\d+$", l.strip()) and cur is not None:
            cur["dofs"].append(int(l))
    return out


def parse_neighbor_lists(name, header):
    """'<header> i has n faces' followed by 'Neighbor is: k' for every non-boundary face, in face order."""
    out, cur = [], None
    for l in lines(name):
        if m := re.match(header + r" (\d+) has (\d+) faces", l):
            cur = {"id": int(m[1]), "n_faces": int(m[2]), "neighbors": []}
            out.append(cur)
        elif m := re.match(r"Neighbor is: (\d+)", l):
            cur["neighbors"].append(int(m[1]))
    return out


def parse_ghosted(name):
    """'Polytope with local index i from rank r' followed, for every face whose neighbour lives on another
    rank, by a block of numbers (bbox corners or DoF indices)."""
    out, cur, block = [], None, None
    for l in lines(name):
        if m := re.match(r"Polytope with local index (\d+) from rank (\d+)", l):
            cur = {"local_index": int(m[1]), "rank": int(m[2]), "ghosts": []}
            out.append(cur)
        elif l.startswith("Neighboring bbox") or l.startswith("DoFs indices"):
            block = []
            cur["ghosts"].append(block)
        elif cur is not None and re.match(r"^[-0-9. e]+$", l.strip()) and l.strip():
            block.append([float(t) for t in l.split()])
    return out


def parse_cell_face_pairs(name):
    out, cell = [], None
    for l in lines(name):
        if m := re.match(r"deal.II cell index = (\d+)", l):
            cell = int(m[1])
        elif m := re.match(r"Local face idx = (\d+)", l):
            out.append([cell, int(m[1])])
    return out


def floats_after(name, pat):
    return [float(m[1]) for l in lines(name) if (m := re.search(pat, l))]


def parse_gmsh_quads(path):
    """vertices (x, y) and 4-node quadrilaterals (gmsh's counter-clockwise node order, 0-based, renumbered
    compactly) of a gmsh 4.1 ASCII file -- the input grid of fully_distributed_poisson_sanity_check_02.cc."""
    toks = open(path).read().split("\n")
    i = toks.index("$Nodes") + 1
    n_blocks, n_nodes = int(toks[i].split()[0]), int(toks[i].split()[1])
    i += 1
    coords = {}
    for _ in range(n_blocks):
        nb = int(toks[i].split()[3])
        tags = [int(toks[i + 1 + k]) for k in range(nb)]
        for k, t in enumerate(tags):
            coords[t] = [float(v) for v in toks[i + 1 + nb + k].split()[:2]]
        i += 1 + 2 * nb
    assert len(coords) == n_nodes
    i = toks.index("$Elements") + 1
    n_blocks = int(toks[i].split()[0])
    i += 1
    quads = []
    for _ in range(n_blocks):
        _, _, etype, nb = (int(v) for v in toks[i].split())
        if etype == 3:
            quads += [[int(v) for v in toks[i + 1 + k].split()[1:5]] for k in range(nb)]
        i += 1 + nb
    used = sorted({t for q in quads for t in q})
    new = {t: k for k, t in enumerate(used)}
    return {"verts": [coords[t] for t in used], "quads": [[new[t] for t in q] for q in quads]}


def parse_ucd_quads(path):
    """vertices and quadrilaterals (counter-clockwise node order, 0-based) of a UCD (.inp) file -- circle-grid.inp of
    unstructured_grid.cc."""
    rows = [l.split() for l in open(path).read().split("\n") if l.strip() and not l.startswith("#")]
    n_verts, n_cells = int(rows[0][0]), int(rows[0][1])
    tag = {int(r[0]): k for k, r in enumerate(rows[1:1 + n_verts])}
    verts = [[float(r[1]), float(r[2])] for r in rows[1:1 + n_verts]]
    quads = [[tag[int(t)] for t in r[3:7]] for r in rows[1 + n_verts:1 + n_verts + n_cells] if r[2] == "quad"]
    return {"verts": verts, "quads": quads}


def parse_unstructured_grid(name):
    """blocks 'Number of faces ... normals ... Perimeter' of unstructured_grid.output"""
    out, cur = [], None
    for l in lines(name):
        if m := re.match(r"Number of faces of this cell: (\d+)", l):
            cur = {"n_faces": int(m.group(1)), "normals": []}
            out.append(cur)
        elif m := re.match(r"For face with index f =(\d+) the normal is (\S+) (\S+)", l):
            cur["normals"].append([float(m.group(2)), float(m.group(3))])
        elif m := re.match(r"Perimeter of agglomeration.* is (\S+)", l):
            cur["perimeter"] = float(m.group(1))
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; goldens can only be regenerated in the build container")
    g = {
        "_provenance": "parsed from /root/reference/test/polydeal/*.output by tests/golden/make_golden.py",
        "sparsity_agglomerated_tria": parse_sparsity("sparsity_agglomerated_tria.output"),
        "agglomerated_neighbors_01": parse_neighbors_faces("agglomerated_neighbors_01.output"),
        "agglomerated_neighbors_02": parse_neighbors_faces("agglomerated_neighbors_02.output"),
        "agglomerated_neighbors_03": parse_nofn("agglomerated_neighbors_03.output"),
        "continuous_face_01": parse_continuous_face("continuous_face_01.output"),
        "continuous_face_02": parse_continuous_face("continuous_face_02.output"),
        "continuous_face_03": parse_continuous_face("continuous_face_03.output"),
        "continuous_face_distorted_grid": parse_continuous_face("continuous_face_distorted_grid.output"),
        "reinit_cell_face_master_master": parse_neighbor_lists("reinit_cell_face_master_master.output", "Polytope with index"),
        "reinit_cell_face_quad_pts": parse_neighbor_lists("reinit_cell_face_quad_pts.output", "Cell with index"),
        "ghosted_bbox_01": parse_ghosted("ghosted_bbox_01.with_mpi=true.with_p4est=true.mpirun=3.output"),
        "ghosted_dofs_01": parse_ghosted("ghosted_dofs_01.with_mpi=true.with_p4est=true.mpirun=3.output"),
        "sparsity_distributed_tria": [
            [int(t) for t in re.match(r"\((\d+),(\d+)\)", l).groups()]
            for l in lines("sparsity_distributed_tria.with_mpi=true.with_p4est=true.mpirun=3.output") if l.startswith("(")
        ],
        "locally_owned_polytope": {
            k: parse_cell_face_pairs(f"locally_owned_polytope_{k}.with_mpi=true.with_p4est=true.mpirun=3.output")
            for k in ("02", "03", "04")
        },
        "hp_structure_01": parse_hp_structure("hp_structure_01.output"),
        "reinit_cell_face_02": parse_reinit_cell_face_02("reinit_cell_face_02.output"),
        "polytope_iterator": parse_polytope_iterator("polytope_iterator.output"),
        "fe_space_on_bbox": floats_after("fe_space_on_bbox.output", r"Sum is: (\S+)"),
        "reinit_cell_face_01": floats_after("reinit_cell_face_01.output", r" is (\S+)$"),
        "agg_handler_bbox_test": [
            [float(t) for t in l.split("=")[1].split()] for l in lines("agg_handler_bbox_test.output")
        ],
        "aggl_handler_master_and_slaves_01": floats_after(
            "aggl_handler_master_and_slaves_01.output", r"associated value: (\S+)"
        ),
        "poisson_sanity_check_01": {
            "x": floats_after("poisson_sanity_check_01.output", r"f\(x,y\)=x:(\S+)"),
            "xplusy": floats_after("poisson_sanity_check_01.output", r"f\(x,y\)=x\+y:(\S+)"),
            "one": floats_after("poisson_sanity_check_01.output", r"Test with 1: (\S+)"),
        },
        "poisson_sanity_check_02": {
            "step": floats_after("poisson_sanity_check_02.output", r"Step function = (\S+)"),
            "v": floats_after("poisson_sanity_check_02.output", r"V function = (\S+)"),
        },
        "poisson_sanity_check_03": {
            "n_subdomains": floats_after("poisson_sanity_check_03.output", r"N subdomains: (\S+)"),
            "x": floats_after("poisson_sanity_check_03.output", r"f\(x,y\)=x:(\S+)"),
            "xplusy": floats_after("poisson_sanity_check_03.output", r"f\(x,y\)=x\+y:(\S+)"),
            "one": floats_after("poisson_sanity_check_03.output", r"Test with 1: (\S+)"),
        },
        "coarse_operator_from_matrix_free": {
            "fine": floats_after("coarse_operator_from_matrix_free.with_mpi=true.with_p4est=true.mpirun=3.output",
                                 r"induced by fine operator: (\S+)"),
            "agglomerated": floats_after("coarse_operator_from_matrix_free.with_mpi=true.with_p4est=true.mpirun=3.output",
                                         r"induced by agglomerated operator: (\S+)"),
        },
        "distributed_poisson_sanity_check_01": {
            "x": floats_after("distributed_poisson_sanity_check_01.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x: (\S+)"),
            "xplusy": floats_after("distributed_poisson_sanity_check_01.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x\+y: (\S+)"),
        },
        "distributed_poisson_sanity_check_02": {
            "x": floats_after("distributed_poisson_sanity_check_02.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x: (\S+)"),
            "xplusy": floats_after("distributed_poisson_sanity_check_02.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x\+y: (\S+)"),
        },
        "fully_distributed_poisson_sanity_check_01": {
            "n_cells": floats_after("fully_distributed_poisson_sanity_check_01.with_mpi=true.with_p4est=true.mpirun=3.output", r"Number of cells: (\S+)"),
            "x": floats_after("fully_distributed_poisson_sanity_check_01.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x: (\S+)"),
            "xplusy": floats_after("fully_distributed_poisson_sanity_check_01.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x\+y: (\S+)"),
        },
        "fully_distributed_poisson_sanity_check_02": {
            "n_cells": floats_after("fully_distributed_poisson_sanity_check_02.with_mpi=true.with_p4est=true.mpirun=3.output", r"Number of cells: (\S+)"),
            "x": floats_after("fully_distributed_poisson_sanity_check_02.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x: (\S+)"),
            "xplusy": floats_after("fully_distributed_poisson_sanity_check_02.with_mpi=true.with_p4est=true.mpirun=3.output", r"f\(x,y\)=x\+y: (\S+)"),
            "input_grid": parse_gmsh_quads(os.path.join(REF, "input_grids", "square.msh")),
        },
        "unstructured_grid": {
            "blocks": parse_unstructured_grid("unstructured_grid.output"),
            "circle_grid": parse_ucd_quads(os.path.join(REF, "circle-grid.inp")),
        },
        "poisson": floats_after("poisson.output", r"(\d\.\d+)"),
    }
    with open(OUT, "w") as f:
        json.dump(g, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
