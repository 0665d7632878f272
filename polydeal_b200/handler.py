"""Python mirror of the reference surface for the hot path.

Names, argument meaning and error behaviour follow
  AgglomerationHandler   /root/reference/include/agglomeration_handler.h:171-575
  AgglomerationAccessor  /root/reference/include/agglomeration_accessor.h:41-299
  PolyUtils::assemble_dg_matrix  /root/reference/include/poly_utils.h:2000-2195
  LinearOperatorMG::vmult / m() / n()  /root/reference/include/linear_operator_for_mg.h:295-357
Everything forwards to the C ABI; polytopes are addressed by `polytope->index()`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Grid:
    """Background quad/hex mesh in deal.II conventions."""

    def __init__(self, handle, dim):
        self._h, self.dim = handle, dim

    @staticmethod
    def structured(dim, n, lo, hi, order=0):
        n = np.ascontiguousarray(np.broadcast_to(n, (dim,)), dtype=np.int32)
        lo = np.ascontiguousarray(np.broadcast_to(lo, (dim,)), dtype=np.float64)
        hi = np.ascontiguousarray(np.broadcast_to(hi, (dim,)), dtype=np.float64)
        h = C.c_void_p()
        K.check(K.lib().pdh_grid_create_structured(dim, _ptr(n), _ptr(lo), _ptr(hi), order, C.byref(h)))
        return Grid(h, dim)

    @staticmethod
    def hyper_cube(dim, a, b, n_refine):
        """GridGenerator::hyper_cube(tria, a, b); tria.refine_global(n_refine)."""
        return Grid.structured(dim, 1 << n_refine, a, b, order=0)

    @staticmethod
    def from_arrays(verts, cell_verts, nbr):
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        cell_verts = np.ascontiguousarray(cell_verts, dtype=np.int32)
        nbr = np.ascontiguousarray(nbr, dtype=np.int32)
        dim = verts.shape[1]
        h = C.c_void_p()
        K.check(K.lib().pdh_grid_create(dim, verts.shape[0], _ptr(verts), cell_verts.shape[0], _ptr(cell_verts), _ptr(nbr), C.byref(h)))
        return Grid(h, dim)

    def __del__(self):
        if getattr(self, "_h", None) and K._lib is not None:
            K._lib.pdh_grid_destroy(self._h)
            self._h = None

    @property
    def n_cells(self):
        return K.lib().pdh_grid_n_cells(self._h)

    @property
    def n_verts(self):
        return K.lib().pdh_grid_n_verts(self._h)

    def arrays(self):
        v = np.empty((self.n_verts, self.dim))
        cv = np.empty((self.n_cells, 1 << self.dim), dtype=np.int32)
        nb = np.empty((self.n_cells, 2 * self.dim), dtype=np.int32)
        K.check(K.lib().pdh_grid_get_arrays(self._h, _ptr(v), _ptr(cv), _ptr(nb)))
        return v, cv, nb

    def set_vertices(self, verts):
        verts = np.ascontiguousarray(verts, dtype=np.float64)
        assert verts.shape == (self.n_verts, self.dim)
        K.check(K.lib().pdh_grid_set_vertices(self._h, _ptr(verts)))

    def distort_random(self, factor, seed):
        """GridTools::distort_random(factor, tria, keep_boundary = true): every interior vertex
        moves by a uniform random vector of at most factor * (shortest incident edge) per
        coordinate.  (numpy's PCG64 stream: like deal.II's own boost RNG, not reproducible
        elsewhere; distorted-mesh checks are therefore made on the same vertex array.)"""
        v, cv, nb = self.arrays()
        dim, vpc = self.dim, 1 << self.dim
        minlen = np.full(self.n_verts, np.inf)
        for d in range(dim):
            for a in range(vpc):
                b = a ^ (1 << d)
                if b < a:
                    continue
                ln = np.linalg.norm(v[cv[:, a]] - v[cv[:, b]], axis=1)
                np.minimum.at(minlen, cv[:, a], ln)
                np.minimum.at(minlen, cv[:, b], ln)
        on_boundary = np.zeros(self.n_verts, dtype=bool)
        for f in range(2 * dim):
            cells = np.nonzero(nb[:, f] < 0)[0]
            for a in range(vpc):
                if ((a >> (f // 2)) & 1) == f % 2:
                    on_boundary[cv[cells, a]] = True
        u = np.random.default_rng(seed).uniform(-1.0, 1.0, size=v.shape)
        v[~on_boundary] += (u * (factor * minlen)[:, None])[~on_boundary]
        self.set_vertices(v)


def partition_graph(xadj, adjncy, n_parts, vertex_weights=None, edge_weights=None):
    """METIS through the C ABI (pdh_partition_graph): part[v] in [0, n_parts)."""
    xadj = np.ascontiguousarray(xadj, dtype=np.int64)
    adjncy = np.ascontiguousarray(adjncy, dtype=np.int64)
    vw = None if vertex_weights is None else np.ascontiguousarray(vertex_weights, dtype=np.int64)
    ew = None if edge_weights is None else np.ascontiguousarray(edge_weights, dtype=np.int64)
    part = np.empty(len(xadj) - 1, dtype=np.int32)
    K.check(K.lib().pdh_partition_graph(len(xadj) - 1, _ptr(xadj), _ptr(adjncy), _ptr(vw) if vw is not None else None,
                                        _ptr(ew) if ew is not None else None, int(n_parts), _ptr(part)))
    return part


def metis_agglomerates(grid: "Grid", n_parts: int):
    """GridTools::partition_triangulation(n_parts, tria, SparsityTools::Partitioner::metis) followed by one
    agglomerate per subdomain (examples/poisson.cc, test/polydeal/continuous_face_02.cc test3): METIS on the
    face-adjacency graph of the cells.  Returns the cell lists in subdomain order, cells ascending."""
    _, _, nbr = grid.arrays()
    mask = nbr >= 0
    xadj = np.zeros(grid.n_cells + 1, dtype=np.int64)
    xadj[1:] = np.cumsum(mask.sum(axis=1))
    part = partition_graph(xadj, nbr[mask], n_parts)
    order = np.argsort(part, kind="stable")
    counts = np.bincount(part, minlength=n_parts)
    return [g.tolist() for g in np.split(order.astype(np.int32), np.cumsum(counts)[:-1]) if len(g)]


class AgglomerationHandler:
    def __init__(self, grid: Grid):
        self.grid, self.dim = grid, grid.dim
        self._h = C.c_void_p()
        K.check(K.lib().pdh_handler_create(grid._h, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and K._lib is not None:
            K._lib.pdh_handler_destroy(self._h)
            self._h = None

    def define_agglomerate(self, cells):
        a = np.ascontiguousarray(cells, dtype=np.int32)
        r = K.lib().pdh_define_agglomerate(self._h, _ptr(a), len(a))
        if r < 0:
            K.check(r)
        return r

    def define_agglomerates(self, groups):
        """define_agglomerate for every cell list of `groups` (a 2-D integer array or a list of lists), in order."""
        if isinstance(groups, np.ndarray) and groups.ndim == 2:
            ptr = np.arange(groups.shape[0] + 1, dtype=np.int64) * groups.shape[1]
            cells = np.ascontiguousarray(groups, dtype=np.int32).ravel()
        else:
            ptr = np.zeros(len(groups) + 1, dtype=np.int64)
            ptr[1:] = np.cumsum([len(g) for g in groups])
            cells = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.int32) for g in groups]), dtype=np.int32)
        K.check(K.lib().pdh_define_agglomerates(self._h, len(ptr) - 1, _ptr(ptr), _ptr(cells)))

    def polytope_graph(self):
        """(xadj, adjncy, vertex_weights, edge_weights) of the polytope adjacency graph (pdh_polytope_graph)."""
        ne = K.lib().pdh_polytope_graph(self._h, None, None, None, None)
        if ne < 0:
            K.check(int(ne))
        np_ = self.n_polytopes
        xadj, adj = np.empty(np_ + 1, dtype=np.int64), np.empty(ne, dtype=np.int64)
        vw, ew = np.empty(np_, dtype=np.int64), np.empty(ne, dtype=np.int64)
        K.lib().pdh_polytope_graph(self._h, _ptr(xadj), _ptr(adj), _ptr(vw), _ptr(ew))
        return xadj, adj, vw, ew

    def initialize_fe_values(self, nq_cell, nq_face=None):
        K.check(K.lib().pdh_initialize_fe_values(self._h, nq_cell, nq_face if nq_face is not None else nq_cell))

    def distribute_agglomerated_dofs(self, fe_kind, degree):
        K.check(K.lib().pdh_distribute_agglomerated_dofs(self._h, fe_kind, degree))

    n_polytopes = property(lambda s: K.lib().pdh_n_polytopes(s._h))
    n_dofs = property(lambda s: K.lib().pdh_n_dofs(s._h))
    n_dofs_per_cell = property(lambda s: K.lib().pdh_n_dofs_per_cell(s._h))

    def _int(self, r):
        if r < -1:
            K.check(r)
        return r

    def master_cell(self, p):
        r = K.lib().pdh_master_cell(self._h, p)
        if r < 0:
            K.check(r)
        return r

    def get_agglomerate(self, p):
        n = K.lib().pdh_n_background_cells(self._h, p)
        if n < 0:
            K.check(n)
        out = np.empty(n, dtype=np.int32)
        K.check(K.lib().pdh_get_agglomerate(self._h, p, _ptr(out)))
        return out

    def n_faces(self, p):
        r = K.lib().pdh_n_faces(self._h, p)
        if r == K.INVALID_UINT:
            raise K.PolydealError(K.PD_ERR_INVALID, K.lib().pd_last_error().decode())
        return r

    def at_boundary(self, p, f):
        r = K.lib().pdh_at_boundary(self._h, p, f)
        if r < 0:
            K.check(r)
        return bool(r)

    def neighbor(self, p, f):
        if f >= self.n_faces(p):
            raise K.PolydealError(K.PD_ERR_INVALID, "face index out of range")
        return K.lib().pdh_neighbor(self._h, p, f)

    def neighbor_of_agglomerated_neighbor(self, p, f):
        if f >= self.n_faces(p):
            raise K.PolydealError(K.PD_ERR_INVALID, "face index out of range")
        return K.lib().pdh_neighbor_of_agglomerated_neighbor(self._h, p, f)

    def interface(self, p, f):
        n = K.lib().pdh_interface(self._h, p, f, None, None, 0)
        if n < 0:
            K.check(n)
        c, fa = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        K.lib().pdh_interface(self._h, p, f, _ptr(c), _ptr(fa), n)
        return list(zip(c.tolist(), fa.tolist()))

    def get_dof_indices(self, p):
        out = np.empty(self.n_dofs_per_cell, dtype=np.uint32)
        K.check(K.lib().pdh_get_dof_indices(self._h, p, _ptr(out)))
        return out

    def bbox(self, p):
        lo, hi = np.empty(self.dim), np.empty(self.dim)
        K.check(K.lib().pdh_bounding_box(self._h, p, _ptr(lo), _ptr(hi)))
        return lo, hi

    def diameter(self, p):
        return K.lib().pdh_diameter(self._h, p)

    def volume(self, p):
        return K.lib().pdh_volume(self._h, p)

    def face_work_item(self, p, f):
        """(iface, side) of face f of polytope p in the work list of the last flatten() (pdh_face_work_item)."""
        i, s = C.c_int32(-1), C.c_int32(-1)
        K.check(K.lib().pdh_face_work_item(self._h, p, f, C.byref(i), C.byref(s)))
        return i.value, s.value

    def real_to_unit(self, p, points):
        """BoundingBox::real_to_unit of polytope p's box (MappingBox::transform_real_to_unit_cell)."""
        x = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        out = np.empty_like(x)
        K.check(K.lib().pdh_real_to_unit(self._h, p, len(x), _ptr(x), _ptr(out)))
        return out

    def unit_to_real(self, p, points):
        x = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        out = np.empty_like(x)
        K.check(K.lib().pdh_unit_to_real(self._h, p, len(x), _ptr(x), _ptr(out)))
        return out

    def create_agglomeration_sparsity_pattern(self):
        nnz = K.lib().pdh_sparsity_nnz(self._h)
        if nnz < 0:
            raise K.PolydealError(K.PD_ERR_STATE, K.lib().pd_last_error().decode())
        rp = np.empty(self.n_dofs + 1, dtype=np.int64)
        cols = np.empty(nnz, dtype=np.int32)
        K.check(K.lib().pdh_create_agglomeration_sparsity_pattern(self._h, _ptr(rp), _ptr(cols)))
        return rp, cols

    def flatten(self, penalty_constant=-1.0, h_rule=K.H_DIAMETER_OF_VISITOR, h_const=1.0, visit_rule=K.VISIT_BY_ID):
        """The flattened agglomeration as a pd_mesh_desc (arrays owned by this handler)."""
        prm = K.FlattenParams(penalty_constant, h_rule, h_const, visit_rule)
        d = K.MeshDesc()
        K.check(K.lib().pdh_flatten(self._h, C.byref(prm), C.byref(d)))
        return d


class FEValuesTables:
    """What `reinit*` returns in the reference (FEValues / FEImmersedSurfaceValues), as arrays."""

    def __init__(self, values, grads, jxw, points, normals):
        self.values, self.grads, self.JxW, self.points, self.normals = values, grads, jxw, points, normals
        self.n_q = len(jxw)

    def shape_value(self, i, q):
        return self.values[i, q]

    def shape_grad(self, i, q):
        return self.grads[i, q]


def fe_evaluate(fe_kind, dim, degree, unit_points):
    """FE_DGQ / FE_AggloDGP on the unit cell: (values [n, Q], grads [n, Q, dim]) at unit_points [Q, dim] (pd_fe_evaluate)."""
    x = np.ascontiguousarray(unit_points, dtype=np.float64).reshape(-1, dim)
    n = (degree + 1) ** dim if fe_kind == 0 else int(np.prod([degree + k for k in range(1, dim + 1)]) // np.prod(range(1, dim + 1)))
    v, g = np.empty((n, len(x))), np.empty((n, len(x), dim))
    K.check(K.lib().pd_fe_evaluate(fe_kind, dim, degree, len(x), _ptr(x), _ptr(v), _ptr(g)))
    return v, g


class SIPOperator:
    """Device-resident SIP operator: assembled block-CSR matrix + apply.

    Plays the role of the matrix/LinearOperatorMG the reference's solvers call
    (vmult / vmult_add / Tvmult / m / n / get_matrix_diagonal_inverse)."""

    def __init__(self, desc: K.MeshDesc, keepalive=None):
        self._keep = keepalive
        self.desc = desc
        self._h = C.c_void_p()
        K.check(K.lib().pd_create(C.byref(desc), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and K._lib is not None:
            K._lib.pd_destroy(self._h)
            self._h = None

    def upload(self, desc=None):
        K.check(K.lib().pd_upload(self._h, C.byref(desc if desc is not None else self.desc)))

    def set_stream(self, cuda_stream_ptr):
        """Enqueue every device call of this operator on the given cudaStream_t (an int).  0 is read as the
        legacy default stream (cudaStreamLegacy), which is what torch's default stream is."""
        ptr = int(cuda_stream_ptr) or 1  # cudaStreamLegacy == (cudaStream_t)0x1
        K.check(K.lib().pd_set_stream(self._h, C.c_void_p(ptr)))
        self._bound_stream = ptr

    def _bind_torch_stream(self):
        """The torch-tensor entry points run on torch's CURRENT stream, so that they are ordered with the
        caller's tensor work like any torch op (the handle's own stream is cudaStreamNonBlocking and would
        race with it).  Re-bound only when the current stream changed (pd_set_stream synchronises)."""
        import torch

        ptr = int(torch.cuda.current_stream().cuda_stream) or 1
        if getattr(self, "_bound_stream", None) != ptr:
            self.set_stream(ptr)

    def synchronize(self):
        K.check(K.lib().pd_synchronize(self._h))

    def build_quadrature(self):
        K.check(K.lib().pd_build_quadrature(self._h))

    def invalidate_quadrature(self):
        K.check(K.lib().pd_invalidate_quadrature(self._h))

    def assemble(self, flags=K.ASSEMBLE_ALL, stiffness=1.0, mass=0.0):
        c = K.Coefficients(stiffness, mass)
        K.check(K.lib().pd_assemble(self._h, flags, C.byref(c)))

    def set_operator(self, flags=K.ASSEMBLE_ALL, stiffness=1.0, mass=0.0):
        """Operator of the matrix-free apply (pd_set_operator)."""
        c = K.Coefficients(stiffness, mass)
        K.check(K.lib().pd_set_operator(self._h, flags, C.byref(c)))

    def force_generic_matrix_free(self, on=True):
        K.check(K.lib().pd_force_generic_matrix_free(self._h, int(on)))

    @property
    def matrix_free_available(self):
        return bool(K.lib().pd_matrix_free_available(self._h))

    @property
    def fine_kernel_last(self):
        """Which fine-mesh kernel the last matrix-free apply launched: 0 none, 1 line, 2 tile, 3 stream."""
        return int(K.lib().pd_fine_kernel_last(self._h))

    @property
    def mapped_fine_available(self):
        """VMULT_MAPPED_FINE (mapped FE_DGQ basis on general hexes) can be applied."""
        return bool(K.lib().pd_mapped_fine_available(self._h))

    def m(self):
        return K.lib().pd_n_dofs(self._h)

    n = m

    @property
    def n_source_dofs(self):
        """Length of vmult source vectors: owned + ghost DoFs (== m() without sharding)."""
        return K.lib().pd_n_source_dofs(self._h)

    @property
    def nnz(self):
        return K.lib().pd_nnz(self._h)

    @property
    def n_dofs_per_cell(self):
        return K.lib().pd_n_dofs_per_cell(self._h)

    def values_device_ptr(self):
        p = C.c_void_p()
        K.check(K.lib().pd_matrix_values_device(self._h, C.byref(p)))
        return p.value

    def values(self, out=None):
        if out is None:
            out = np.empty(self.nnz)
        K.check(K.lib().pd_matrix_values_to_host(self._h, _ptr(out)))
        return out

    def values_to_host_ptr(self, host_ptr, wait=True):
        """D2H of the CSR values into (pinned) host memory; wait=False returns without synchronising."""
        fn = K.lib().pd_matrix_values_to_host if wait else K.lib().pd_matrix_values_to_host_async
        K.check(fn(self._h, C.c_void_p(host_ptr)))

    def pattern(self):
        rp = np.empty(self.m() + 1, dtype=np.int64)
        cols = np.empty(self.nnz, dtype=np.int32)
        K.check(K.lib().pd_matrix_pattern_to_host(self._h, _ptr(rp), _ptr(cols)))
        return rp, cols

    def scipy(self):
        import scipy.sparse as sp

        rp, cols = self.pattern()
        return sp.csr_matrix((self.values(), cols, rp), shape=(self.m(), self.n_source_dofs))

    # --- apply: device pointers (ints, e.g. torch.Tensor.data_ptr()) or host numpy arrays
    def vmult_ptr(self, dst_ptr, src_ptr, mode=K.VMULT_BLOCK_CSR, add=False):
        fn = K.lib().pd_vmult_add if add else K.lib().pd_vmult
        K.check(fn(self._h, mode, C.c_void_p(src_ptr), C.c_void_p(dst_ptr)))

    def vmult(self, dst, src, mode=K.VMULT_BLOCK_CSR):
        """dst = A src.  torch CUDA tensors (float64, contiguous) stay on the device;
        numpy arrays go through pd_vmult_host (H2D, apply, D2H)."""
        if isinstance(src, np.ndarray):
            assert src.dtype == np.float64 and dst.dtype == np.float64 and src.flags.c_contiguous and dst.flags.c_contiguous
            assert src.size == self.m() and dst.size == self.m()
            K.check(K.lib().pd_vmult_host(self._h, mode, _ptr(src), _ptr(dst)))
        else:
            self._check_tensor(src), self._check_tensor(dst)
            self._bind_torch_stream()
            self.vmult_ptr(dst.data_ptr(), src.data_ptr(), mode)
        return dst

    def vmult_add(self, dst, src, mode=K.VMULT_BLOCK_CSR):
        self._check_tensor(src), self._check_tensor(dst)
        self._bind_torch_stream()
        self.vmult_ptr(dst.data_ptr(), src.data_ptr(), mode, add=True)
        return dst

    # the SIP operator is symmetric (include/utils.h:431-445 does the same)
    Tvmult = vmult
    Tvmult_add = vmult_add

    def vmult_host_ptr(self, dst_host_ptr, src_host_ptr, mode=K.VMULT_BLOCK_CSR):
        K.check(K.lib().pd_vmult_host(self._h, mode, C.c_void_p(src_host_ptr), C.c_void_p(dst_host_ptr)))

    def get_matrix_diagonal_inverse(self, out, mode=K.VMULT_BLOCK_CSR):
        """Inverse diagonal of the operator `mode` applies (pd_diagonal_inverse_of); matrix-free modes
        need no assembled matrix."""
        self._check_tensor(out)
        self._bind_torch_stream()
        K.check(K.lib().pd_diagonal_inverse_of(self._h, mode, C.c_void_p(out.data_ptr())))
        return out

    def cg_solve(self, x, b, max_iter=1000, rel_tol=1e-10, jacobi=True, mode=K.VMULT_BLOCK_CSR):
        """SolverCG around vmult, device resident (pd_cg_solve).  Returns (iterations, relative residual);
        `last_cg_converged` is False when the tolerance was not reached within max_iter (PD_NOT_CONVERGED)."""
        self._check_tensor(x), self._check_tensor(b)
        self._bind_torch_stream()
        it, rr = C.c_int(0), C.c_double(0.0)
        rc = K.check(K.lib().pd_cg_solve(self._h, mode, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), max_iter, rel_tol,
                                         int(jacobi), C.byref(it), C.byref(rr)))
        self.last_cg_converged = rc != K.PD_NOT_CONVERGED
        return it.value, rr.value

    # ---- data at the agglomerated quadrature points: right-hand side and error norms ----
    def n_quadrature_points(self, faces=False):
        return int(K.lib().pd_n_quadrature_points(self._h, int(faces)))

    def quadrature(self):
        """Device views (torch, no copy) of the agglomerated quadrature:
        dict(vol_x [dim, Q], vol_jxw [Q], face_x [dim, Qf], face_n [dim, Qf], face_jxw [Qf])."""
        import torch

        ptrs = [C.c_void_p() for _ in range(5)]
        self._bind_torch_stream()
        K.check(K.lib().pd_quadrature_device(self._h, *[C.byref(p) for p in ptrs]))
        Q, Qf, dim = self.n_quadrature_points(), self.n_quadrature_points(True), self.desc.dim

        def view(ptr, shape):
            n = int(np.prod(shape))
            if n == 0:
                return torch.zeros(shape, dtype=torch.float64, device="cuda")
            iface = {"shape": tuple(shape), "typestr": "<f8", "data": (ptr.value, False), "version": 2}
            holder = type("_DevView", (), {"__cuda_array_interface__": iface})()
            return torch.as_tensor(holder, device="cuda")

        names = ["vol_x", "vol_jxw", "face_x", "face_n", "face_jxw"]
        shapes = [(dim, Q), (Q,), (dim, Qf), (dim, Qf), (Qf,)]
        return {k: view(p, sh) for k, p, sh in zip(names, ptrs, shapes)}

    def assemble_rhs(self, rhs, f_vol=None, g_face=None, stiffness=1.0):
        """rhs_i = sum_q f_q phi_i w_q + stiffness * boundary Dirichlet terms (pd_assemble_rhs);
        f_vol [Q] / g_face [Qf] are CUDA tensors of values at the quadrature points."""
        self._check_tensor(rhs)
        if f_vol is not None:
            self._check_tensor(f_vol, self.n_quadrature_points())
        if g_face is not None:
            self._check_tensor(g_face, self.n_quadrature_points(True))
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._bind_torch_stream()
        K.check(K.lib().pd_assemble_rhs(self._h, ptr(f_vol), ptr(g_face), float(stiffness), ptr(rhs)))
        return rhs

    def error_norms(self, u, exact, exact_grad=None):
        """PolyUtils::compute_global_error: (L2, H1-seminorm or None) of u_h - u (pd_error_norms)."""
        Q = self.n_quadrature_points()
        self._check_tensor(u, K.lib().pd_n_source_dofs(self._h)), self._check_tensor(exact, Q)
        l2, h1 = C.c_double(0.0), C.c_double(0.0)
        if exact_grad is not None:
            self._check_tensor(exact_grad, Q * self.desc.dim)
        self._bind_torch_stream()
        K.check(K.lib().pd_error_norms(self._h, C.c_void_p(u.data_ptr()), C.c_void_p(exact.data_ptr()),
                                       C.c_void_p(exact_grad.data_ptr()) if exact_grad is not None else None,
                                       C.byref(l2), C.byref(h1) if exact_grad is not None else None))
        return l2.value, (h1.value if exact_grad is not None else None)

    # ---- the reinit() family: FEValues tables of one polytope / face, as numpy arrays on the host ----
    def reinit(self, poly):
        """ah.reinit(polytope): FEValuesTables(values [n, Q], grads [n, Q, dim], JxW [Q], points [Q, dim])."""
        Q = K.lib().pd_reinit_n_points(self._h, poly)
        if Q < 0:
            K.check(int(Q))
        n, dim = self.n_dofs_per_cell, self.desc.dim
        t = FEValuesTables(np.empty((n, Q)), np.empty((n, Q, dim)), np.empty(Q), np.empty((Q, dim)), None)
        K.check(K.lib().pd_reinit_polytope(self._h, poly, _ptr(t.values), _ptr(t.grads), _ptr(t.JxW), _ptr(t.points)))
        return t

    def reinit_face(self, iface, side=0):
        """ah.reinit(polytope, f) with (iface, side) = AgglomerationHandler.face_work_item(polytope, f)."""
        Q = K.lib().pd_reinit_iface_n_points(self._h, iface)
        if Q < 0:
            K.check(int(Q))
        n, dim = self.n_dofs_per_cell, self.desc.dim
        t = FEValuesTables(np.empty((n, Q)), np.empty((n, Q, dim)), np.empty(Q), np.empty((Q, dim)), np.empty((Q, dim)))
        K.check(K.lib().pd_reinit_face(self._h, iface, side, _ptr(t.values), _ptr(t.grads), _ptr(t.JxW), _ptr(t.points),
                                       _ptr(t.normals)))
        return t

    def reinit_interface(self, iface):
        """ah.reinit_interface(...): the pair (side 0, side 1) at aligned points, each with its own outward normal."""
        return self.reinit_face(iface, 0), self.reinit_face(iface, 1)

    def agglomerated_quadrature(self, poly):
        """(unit points [Q, dim], JxW [Q], real points [Q, dim]) of polytope `poly` (pd_agglomerated_quadrature)."""
        Q = K.lib().pd_reinit_n_points(self._h, poly)
        if Q < 0:
            K.check(int(Q))
        dim = self.desc.dim
        u, w, x = np.empty((Q, dim)), np.empty(Q), np.empty((Q, dim))
        K.check(K.lib().pd_agglomerated_quadrature(self._h, poly, _ptr(u), _ptr(w), _ptr(x)))
        return u, w, x

    def estimate_lambda_max(self, n_iterations=20, mode=K.VMULT_BLOCK_CSR):
        lam = C.c_double(0.0)
        K.check(K.lib().pd_estimate_lambda_max(self._h, mode, n_iterations, C.byref(lam)))
        return lam.value

    def chebyshev_smooth(self, x, b, degree, lambda_max, smoothing_range=20.0, zero_initial_guess=True,
                         mode=K.VMULT_BLOCK_CSR):
        """PreconditionChebyshev (Jacobi inner preconditioner) as a smoother (pd_chebyshev_smooth)."""
        self._check_tensor(x), self._check_tensor(b)
        self._bind_torch_stream()
        K.check(K.lib().pd_chebyshev_smooth(self._h, mode, degree, lambda_max, smoothing_range, C.c_void_p(b.data_ptr()),
                                            C.c_void_p(x.data_ptr()), int(zero_initial_guess)))
        return x

    def _check_tensor(self, t, numel=None):
        import torch

        numel = self.m() if numel is None else numel
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == numel):
            raise K.PolydealError(K.PD_ERR_INVALID, f"expected a contiguous float64 CUDA tensor of {numel} entries")

    def copy_array(self, name):
        cnt = C.c_int64()
        K.check(K.lib().pd_copy_array(self._h, name.encode(), None, C.byref(cnt)))
        out = np.empty(cnt.value)
        K.check(K.lib().pd_copy_array(self._h, name.encode(), _ptr(out), C.byref(cnt)))
        return out

    @property
    def assembly_path(self):
        """'tensor' (axis-aligned sub-cells, pd_cartesian.cu) or 'dmma' (pd_assemble.cu) for the last assemble()."""
        return {0: "dmma", 1: "tensor"}.get(K.lib().pd_assembly_path(self._h))

    def tensor_path_stats(self):
        """dict(axis_aligned, cell_bricks, face_bricks, diag_items) of the tensor path (pd_tensor_path_stats)."""
        st = (C.c_int64 * 5)()
        K.check(K.lib().pd_tensor_path_stats(self._h, st))
        return {"axis_aligned": bool(st[0]), "cell_bricks": int(st[1]), "face_bricks": int(st[2]), "diag_items": int(st[3]),
                "apply_items": int(st[4])}

    @property
    def launch_count(self):
        return K.lib().pd_launch_count(self._h)

    def last_kernel_ms(self):
        ms = (C.c_float * 4)()
        K.check(K.lib().pd_last_kernel_ms(self._h, ms))
        return {"volume": ms[0], "faces": ms[1], "reduce": ms[2], "quadrature": ms[3]}


def assemble_dg_matrix(ah: AgglomerationHandler, penalty_constant=-1.0, h_rule=K.H_DIAMETER_OF_VISITOR, h_const=1.0,
                       visit_rule=K.VISIT_BY_ID, with_boundary=True, stiffness_coeff=1.0, mass_coeff=0.0) -> SIPOperator:
    """PolyUtils::assemble_dg_matrix(system_matrix, fe_dg, ah): flattens the handler,
    creates the device copy and assembles.  penalty_constant < 0 selects the library's
    10 (p+dim)(p+1)."""
    desc = ah.flatten(penalty_constant, h_rule, h_const, visit_rule)
    op = SIPOperator(desc, keepalive=ah)
    flags = K.ASSEMBLE_ALL if with_boundary else (K.ASSEMBLE_VOLUME | K.ASSEMBLE_INTERIOR)
    op.assemble(flags, stiffness_coeff, mass_coeff)
    return op


class Transfer:
    """Level transfer applied on the fly (pd_transfer_*): `Transfer(coarse_op, fine_op, parent)` is
    Utils::fill_injection_matrix between two agglomeration levels (include/utils.h:95-270),
    `Transfer.to_cells(op)` is PolyUtils::fill_interpolation_matrix onto the mesh's FE_DGQ space
    (include/poly_utils.h:1469-1634).  prolongate / restrict_and_add as in MGTransferAgglomeration
    (source/multigrid_amg.cc:66-110)."""

    def __init__(self, coarse: SIPOperator, fine: SIPOperator = None, parent_of_fine=None, _to_cells=False):
        self._h = C.c_void_p()
        self._keep = (coarse, fine)
        if _to_cells:
            K.check(K.lib().pd_transfer_create_to_cells(coarse._h, C.byref(self._h)))
        else:
            parent = np.ascontiguousarray(parent_of_fine, dtype=np.int32)
            assert parent.shape == (fine.desc.n_owned_polytopes or fine.desc.n_polytopes,)
            K.check(K.lib().pd_transfer_create(coarse._h, fine._h, _ptr(parent), C.byref(self._h)))

    @staticmethod
    def to_cells(op: SIPOperator):
        return Transfer(op, _to_cells=True)

    def __del__(self):
        if getattr(self, "_h", None) and K._lib is not None:
            K._lib.pd_transfer_destroy(self._h)
            self._h = None

    def m(self):
        return int(K.lib().pd_transfer_m(self._h))

    def n(self):
        return int(K.lib().pd_transfer_n(self._h))

    def _check(self, t, numel):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == numel):
            raise K.PolydealError(K.PD_ERR_INVALID, f"expected a contiguous float64 CUDA tensor of {numel} entries")

    def prolongate(self, dst_fine, src_coarse, add=False):
        self._check(dst_fine, self.m()), self._check(src_coarse, self.n())
        self._keep[0]._bind_torch_stream()
        K.check(K.lib().pd_transfer_prolongate(self._h, C.c_void_p(src_coarse.data_ptr()), C.c_void_p(dst_fine.data_ptr()), int(add)))
        return dst_fine

    def prolongate_and_add(self, dst_fine, src_coarse):
        return self.prolongate(dst_fine, src_coarse, add=True)

    def restrict_and_add(self, dst_coarse, src_fine):
        return self.restrict(dst_coarse, src_fine, add=True)

    def restrict(self, dst_coarse, src_fine, add=False):
        self._check(dst_coarse, self.n()), self._check(src_fine, self.m())
        self._keep[0]._bind_torch_stream()
        K.check(K.lib().pd_transfer_restrict(self._h, C.c_void_p(src_fine.data_ptr()), C.c_void_p(dst_coarse.data_ptr()), int(add)))
        return dst_coarse

    def synchronize(self):
        """Transfers run on the coarse operator's stream."""
        self._keep[0].synchronize()

    vmult = prolongate
    Tvmult = restrict
