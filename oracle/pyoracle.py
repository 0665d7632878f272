"""ctypes front-end of the CPU oracle (oracle/polyoracle.hpp).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The
class and method names mirror the reference surface they restate
(/root/reference/include/agglomeration_handler.h:171-575,
include/agglomeration_accessor.h:41-299) so tests read like the reference's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpolyoracle.so")

FE_DGQ, FE_AGGLODGP = 0, 1
H_DIAMETER_OF_VISITOR, H_MAX_INVERSE_DIAMETER, H_CONSTANT, H_NORMAL_EXTENT = 0, 1, 2, 3
VISIT_BY_ID, VISIT_BY_INDEX = 0, 1
INVALID_UINT = 0xFFFFFFFF


def build(force: bool = False) -> str:
    """Compile oracle/libpolyoracle.so with the committed Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("polyoracle.hpp", "polyoracle_capi.cpp")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpolyoracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("penalty_constant", C.c_double),
        ("h_rule", C.c_int),
        ("h_const", C.c_double),
        ("visit_rule", C.c_int),
        ("with_boundary", C.c_int),
        ("stiffness_coeff", C.c_double),
        ("mass_coeff", C.c_double),
        ("n_threads", C.c_int),
        ("poly_stride", C.c_int),
        ("poly_offset", C.c_int),
        ("discard_scatter", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, u32, f64, i64 = C.c_void_p, C.c_int, C.c_uint, C.c_double, C.c_int64
        P = C.POINTER
        sig = {
            "po_last_error": (C.c_char_p, []),
            "po_gauss_1d": (None, [i32, P(f64), P(f64)]),
            "po_gauss_lobatto_nodes": (None, [i32, P(f64)]),
            "po_fe_n_dofs": (i32, [i32, i32, i32]),
            "po_fe_evaluate": (None, [i32, i32, i32, P(f64), P(f64), P(f64)]),
            "po_grid_structured": (vp, [i32, P(i32), P(f64), P(f64), i32]),
            "po_grid_from_arrays": (vp, [i32, i32, P(f64), i32, P(i32), P(i32)]),
            "po_grid_free": (None, [vp]),
            "po_grid_distort_random": (None, [vp, f64, C.c_uint64]),
            "po_grid_dim": (i32, [vp]),
            "po_grid_n_cells": (i32, [vp]),
            "po_grid_n_verts": (i32, [vp]),
            "po_grid_cell_vertices": (None, [vp, i32, P(f64)]),
            "po_grid_neighbor": (i32, [vp, i32, i32]),
            "po_grid_copy_arrays": (None, [vp, P(f64), P(i32), P(i32)]),
            "po_ah_create": (vp, [vp]),
            "po_ah_free": (None, [vp]),
            "po_ah_define_agglomerate": (i32, [vp, P(i32), i32]),
            "po_ah_initialize_fe_values": (None, [vp, i32, i32]),
            "po_ah_distribute_agglomerated_dofs": (i32, [vp, i32, i32]),
            "po_ah_n_polytopes": (i32, [vp]),
            "po_ah_n_dofs": (i32, [vp]),
            "po_ah_n_dofs_per_cell": (i32, [vp]),
            "po_ah_master_cell": (i32, [vp, i32]),
            "po_ah_master_slave_value": (f64, [vp, i32]),
            "po_ah_n_subcells": (i32, [vp, i32]),
            "po_ah_get_agglomerate": (None, [vp, i32, P(i32)]),
            "po_ah_bbox": (None, [vp, i32, P(f64), P(f64)]),
            "po_ah_n_faces": (u32, [vp, i32]),
            "po_ah_at_boundary": (i32, [vp, i32, u32]),
            "po_ah_neighbor": (i32, [vp, i32, u32]),
            "po_ah_neighbor_of_agglomerated_neighbor": (u32, [vp, i32, u32]),
            "po_ah_interface": (i32, [vp, i32, u32, P(i32), P(i32), i32]),
            "po_ah_get_dof_indices": (None, [vp, i32, P(u32)]),
            "po_ah_diameter": (f64, [vp, i32]),
            "po_ah_volume": (f64, [vp, i32]),
            "po_ah_reinit": (i32, [vp, i32, i32, u32, P(f64), P(f64), P(f64), P(f64), P(f64)]),
            "po_ah_sparsity_nnz": (i64, [vp]),
            "po_ah_sparsity": (None, [vp, P(i64), P(i32)]),
            "po_assemble_dg_matrix": (vp, [vp, P(_Params), P(f64)]),
            "po_assemble_block_rows": (i64, [vp, P(_Params), P(i32), i32, P(i64), P(i32), P(f64), P(f64)]),
            "po_matrix_free": (None, [vp]),
            "po_matrix_n_rows": (i64, [vp]),
            "po_matrix_nnz": (i64, [vp]),
            "po_matrix_copy": (None, [vp, P(i64), P(i32), P(f64)]),
            "po_matrix_vmult": (None, [vp, P(f64), P(f64), i32]),
            "po_mapped_fine_vmult": (i32, [vp, i32, i32, f64, f64, i32, i32, i32, P(f64), P(f64)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _err():
    return lib().po_last_error().decode()


def gauss_1d(n):
    x, w = np.empty(n), np.empty(n)
    lib().po_gauss_1d(n, _p(x, C.c_double), _p(w, C.c_double))
    return x, w


def gauss_lobatto_nodes(n):
    x = np.empty(n)
    lib().po_gauss_lobatto_nodes(n, _p(x, C.c_double))
    return x


def fe_evaluate(kind, dim, degree, xhat):
    n = lib().po_fe_n_dofs(kind, dim, degree)
    xhat = np.ascontiguousarray(xhat, dtype=np.float64)
    v, g = np.empty(n), np.empty((n, dim))
    lib().po_fe_evaluate(kind, dim, degree, _p(xhat, C.c_double), _p(v, C.c_double), _p(g, C.c_double))
    return v, g


class Grid:
    """Hypercube-cell mesh with deal.II conventions.

    order=0: GridGenerator::hyper_cube + refine_global (Morton order);
    order=1: subdivided_hyper_rectangle (lexicographic)."""

    def __init__(self, dim, n, lo, hi, order=0):
        n = np.ascontiguousarray(np.broadcast_to(n, (dim,)), dtype=np.int32)
        lo = np.ascontiguousarray(np.broadcast_to(lo, (dim,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(hi, (dim,)), dtype=np.float64)
        self.h = lib().po_grid_structured(dim, _p(n, C.c_int), _p(lo, C.c_double), _p(hi, C.c_double), order)
        if not self.h:
            raise RuntimeError(_err())
        self.dim = dim
        self.n = tuple(int(v) for v in n)
        self.order = order

    @staticmethod
    def hyper_cube(dim, a, b, n_refine):
        return Grid(dim, 1 << n_refine, a, b, order=0)

    @staticmethod
    def from_arrays(verts, cell_verts, nbr):
        """Any conforming quad / hex mesh (what GridIn would read): vertices of a cell in deal.II's lexicographic
        order, nbr[c, 2 * direction + side] = neighbour or -1.  Neighbours may be rotated against each other."""
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        cell_verts = np.ascontiguousarray(cell_verts, dtype=np.int32)
        nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        g = Grid.__new__(Grid)
        g.dim = verts.shape[1]
        g.h = lib().po_grid_from_arrays(g.dim, len(verts), _p(verts, C.c_double), len(cell_verts), _p(cell_verts, C.c_int),
                                        _p(nbr, C.c_int))
        if not g.h:
            raise RuntimeError(_err())
        g.n, g.order = None, None
        return g

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.po_grid_free(self.h)
            self.h = None

    def distort_random(self, factor, seed):
        lib().po_grid_distort_random(self.h, factor, seed)

    @property
    def n_cells(self):
        return lib().po_grid_n_cells(self.h)

    @property
    def n_verts(self):
        return lib().po_grid_n_verts(self.h)

    def cell_vertices(self, c):
        out = np.empty((1 << self.dim, self.dim))
        lib().po_grid_cell_vertices(self.h, c, _p(out, C.c_double))
        return out

    def neighbor(self, c, f):
        return lib().po_grid_neighbor(self.h, c, f)

    def arrays(self):
        """(verts[nv,dim], cell_verts[nc,2^dim], nbr[nc,2*dim]) -- the same mesh
        for the product under test."""
        nv, nc, d = self.n_verts, self.n_cells, self.dim
        v = np.empty((nv, d))
        cv = np.empty((nc, 1 << d), dtype=np.int32)
        nb = np.empty((nc, 2 * d), dtype=np.int32)
        lib().po_grid_copy_arrays(self.h, _p(v, C.c_double), _p(cv, C.c_int), _p(nb, C.c_int))
        return v, cv, nb


class FEValues:
    """What `ah.reinit(...)` returns in the reference (FEValues / FEImmersedSurfaceValues)."""

    def __init__(self, points, jxw, normals, values, grads):
        self.points, self.JxW, self.normals, self.values, self.grads = points, jxw, normals, values, grads
        self.n_q = len(jxw)

    def shape_value(self, i, q):
        return self.values[i, q]

    def shape_grad(self, i, q):
        return self.grads[i, q]


class AgglomerationHandler:
    def __init__(self, grid: Grid):
        self.grid = grid
        self.h = lib().po_ah_create(grid.h)
        self.dim = grid.dim

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.po_ah_free(self.h)
            self.h = None

    # PolyUtils::collect_cells_for_agglomeration returns cells in ACTIVE-CELL
    # order (include/poly_utils.h:532-538), so the master is min(idxs).
    def define_agglomerate(self, cells, collect=False):
        cells = sorted(cells) if collect else list(cells)
        a = np.ascontiguousarray(cells, dtype=np.int32)
        r = lib().po_ah_define_agglomerate(self.h, _p(a, C.c_int), len(a))
        if r < 0:
            raise RuntimeError(_err())
        return r

    def initialize_fe_values(self, nq_cell, nq_face=None):
        lib().po_ah_initialize_fe_values(self.h, nq_cell, nq_face if nq_face is not None else nq_cell)

    def distribute_agglomerated_dofs(self, fe_kind, degree):
        if lib().po_ah_distribute_agglomerated_dofs(self.h, fe_kind, degree):
            raise RuntimeError(_err())

    @property
    def n_polytopes(self):
        return lib().po_ah_n_polytopes(self.h)

    @property
    def n_dofs(self):
        return lib().po_ah_n_dofs(self.h)

    @property
    def n_dofs_per_cell(self):
        return lib().po_ah_n_dofs_per_cell(self.h)

    def master_cell(self, p):
        return lib().po_ah_master_cell(self.h, p)

    def master_slave_value(self, c):
        return lib().po_ah_master_slave_value(self.h, c)

    def get_agglomerate(self, p):
        n = lib().po_ah_n_subcells(self.h, p)
        out = np.empty(n, dtype=np.int32)
        lib().po_ah_get_agglomerate(self.h, p, _p(out, C.c_int))
        return out

    def bbox(self, p):
        lo, hi = np.empty(self.dim), np.empty(self.dim)
        lib().po_ah_bbox(self.h, p, _p(lo, C.c_double), _p(hi, C.c_double))
        return lo, hi

    def n_faces(self, p):
        return lib().po_ah_n_faces(self.h, p)

    def at_boundary(self, p, f):
        return bool(lib().po_ah_at_boundary(self.h, p, f))

    def neighbor(self, p, f):
        return lib().po_ah_neighbor(self.h, p, f)

    def neighbor_of_agglomerated_neighbor(self, p, f):
        return lib().po_ah_neighbor_of_agglomerated_neighbor(self.h, p, f)

    def interface(self, p, f):
        n = lib().po_ah_interface(self.h, p, f, None, None, 0)
        c, fa = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        lib().po_ah_interface(self.h, p, f, _p(c, C.c_int), _p(fa, C.c_int), n)
        return list(zip(c.tolist(), fa.tolist()))

    def get_dof_indices(self, p):
        out = np.empty(self.n_dofs_per_cell, dtype=np.uint32)
        lib().po_ah_get_dof_indices(self.h, p, _p(out, C.c_uint))
        return out

    def diameter(self, p):
        return lib().po_ah_diameter(self.h, p)

    def volume(self, p):
        return lib().po_ah_volume(self.h, p)

    def _reinit(self, kind, p, f):
        nq = lib().po_ah_reinit(self.h, kind, p, f, None, None, None, None, None)
        if nq < 0:
            raise RuntimeError(_err())
        n, d = self.n_dofs_per_cell, self.dim
        pts, jxw = np.empty((nq, d)), np.empty(nq)
        nrm = np.empty((nq, d)) if kind == 1 else None
        val, grd = np.empty((n, nq)), np.empty((n, nq, d))
        lib().po_ah_reinit(
            self.h, kind, p, f, _p(pts, C.c_double), _p(jxw, C.c_double), _p(nrm, C.c_double),
            _p(val, C.c_double), _p(grd, C.c_double),
        )
        return FEValues(pts, jxw, nrm, val, grd)

    def reinit(self, p, f=None):
        return self._reinit(0, p, 0) if f is None else self._reinit(1, p, f)

    def reinit_interface(self, p_in, p_out, f_in, f_out):
        return self._reinit(1, p_in, f_in), self._reinit(1, p_out, f_out)

    def create_agglomeration_sparsity_pattern(self):
        nnz = lib().po_ah_sparsity_nnz(self.h)
        rp = np.empty(self.n_dofs + 1, dtype=np.int64)
        cols = np.empty(nnz, dtype=np.int32)
        lib().po_ah_sparsity(self.h, _p(rp, C.c_int64), _p(cols, C.c_int))
        return rp, cols


class Matrix:
    def __init__(self, h, seconds):
        self.h = h
        self.seconds = seconds

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.po_matrix_free(self.h)
            self.h = None

    def csr(self):
        nr, nnz = lib().po_matrix_n_rows(self.h), lib().po_matrix_nnz(self.h)
        rp, cols, vals = np.empty(nr + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        lib().po_matrix_copy(self.h, _p(rp, C.c_int64), _p(cols, C.c_int), _p(vals, C.c_double))
        return rp, cols, vals

    def values(self):
        vals = np.empty(lib().po_matrix_nnz(self.h))
        lib().po_matrix_copy(self.h, None, None, _p(vals, C.c_double))
        return vals

    def scipy(self):
        import scipy.sparse as sp

        rp, cols, vals = self.csr()
        n = len(rp) - 1
        return sp.csr_matrix((vals, cols, rp), shape=(n, n))

    def vmult(self, x, n_threads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        lib().po_matrix_vmult(self.h, _p(x, C.c_double), _p(y, C.c_double), n_threads)
        return y


def assemble_dg_matrix(
    ah: AgglomerationHandler,
    penalty_constant=None,
    h_rule=H_DIAMETER_OF_VISITOR,
    h_const=1.0,
    visit_rule=VISIT_BY_ID,
    with_boundary=True,
    stiffness_coeff=1.0,
    mass_coeff=0.0,
    n_threads=1,
    degree=None,
    poly_stride=1,
    poly_offset=0,
    discard_scatter=False,
) -> Matrix:
    """PolyUtils::assemble_dg_matrix (include/poly_utils.h:2000-2195); the
    default penalty is the library's 10 (p+dim)(p+1) (:2018-2019)."""
    if penalty_constant is None:
        assert degree is not None
        penalty_constant = 10.0 * (degree + ah.dim) * (degree + 1)
    prm = _Params(penalty_constant, h_rule, h_const, visit_rule, int(with_boundary), stiffness_coeff, mass_coeff, n_threads,
                  poly_stride, poly_offset, int(discard_scatter))
    sec = C.c_double(0.0)
    h = lib().po_assemble_dg_matrix(ah.h, C.byref(prm), C.byref(sec))
    if not h:
        raise RuntimeError(_err())
    return Matrix(h, sec.value)


def assemble_block_rows(ah: AgglomerationHandler, polys, penalty_constant=None, h_rule=H_DIAMETER_OF_VISITOR, h_const=1.0,
                        visit_rule=VISIT_BY_ID, with_boundary=True, stiffness_coeff=1.0, mass_coeff=0.0, n_threads=1,
                        degree=None):
    """The complete block rows of the polytopes `polys` (every interface evaluated from its visiting side as
    PolyUtils::assemble_dg_matrix does, include/poly_utils.h:2086-2132): parity checks at sizes where the whole
    matrix is out of reach for the scalar oracle.  Returns (ptr, bcol, rows, seconds): polytope s has block columns
    bcol[ptr[s]:ptr[s+1]] (ascending) and rows[s] of shape (n, nb_s * n) = its scalar CSR rows."""
    if penalty_constant is None:
        assert degree is not None
        penalty_constant = 10.0 * (degree + ah.dim) * (degree + 1)
    prm = _Params(penalty_constant, h_rule, h_const, visit_rule, int(with_boundary), stiffness_coeff, mass_coeff, n_threads, 1, 0, 0)
    polys = np.ascontiguousarray(polys, dtype=np.int32)
    n = ah.n_dofs_per_cell
    ptr = np.empty(len(polys) + 1, dtype=np.int64)
    nblk = sum(1 + sum(1 for f in range(ah.n_faces(int(p))) if not ah.at_boundary(int(p), f)) for p in polys)
    bcol = np.empty(nblk, dtype=np.int32)
    vals = np.empty(nblk * n * n)
    sec = C.c_double(0.0)
    got = lib().po_assemble_block_rows(ah.h, C.byref(prm), _p(polys, C.c_int), len(polys), _p(ptr, C.c_int64), _p(bcol, C.c_int),
                                       _p(vals, C.c_double), C.byref(sec))
    if got < 0:
        raise RuntimeError(_err())
    assert got == nblk
    rows = [vals[ptr[s] * n * n: ptr[s + 1] * n * n].reshape(n, -1) for s in range(len(polys))]
    return ptr, bcol, rows, sec.value


def mapped_fine_vmult(grid: Grid, degree, nq, x, stiffness=1.0, mass=0.0, volume=True, boundary=True, interior=True):
    """y = (mass M + stiffness K_SIP) x of the fine-mesh operator with the mapped FE_DGQ basis
    (LaplaceOperatorDG / MonodomainOperatorDG semantics, include/utils.h:819-925, 1565-1659;
    matrix-based twin examples/monodomain_DG3D.cc:1374-1622)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    if lib().po_mapped_fine_vmult(grid.h, degree, nq, stiffness, mass, int(volume), int(boundary), int(interior),
                                  _p(x, C.c_double), _p(y, C.c_double)) != 0:
        raise RuntimeError(_err())
    return y
