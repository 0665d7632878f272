"""ctypes binding of include/polydeal_b200.h (the C-ABI shared library built from
polydeal_b200/csrc).  Fails loudly when the library is missing: there is no
Python / CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpolydeal_b200.so")

PD_OK, PD_ERR_INVALID, PD_ERR_CUDA, PD_ERR_UNSUPPORTED, PD_ERR_NO_DEVICE, PD_ERR_STATE = 0, -1, -2, -3, -4, -5
PD_NOT_CONVERGED = 1
INVALID_UINT = 0xFFFFFFFF
ASSEMBLE_VOLUME, ASSEMBLE_BOUNDARY, ASSEMBLE_INTERIOR, ASSEMBLE_ALL = 1, 2, 4, 7
VMULT_BLOCK_CSR, VMULT_MATRIX_FREE, VMULT_MAPPED_FINE = 0, 1, 2
H_DIAMETER_OF_VISITOR, H_MAX_INVERSE_DIAMETER, H_CONSTANT, H_NORMAL_EXTENT = 0, 1, 2, 3
VISIT_BY_ID, VISIT_BY_INDEX = 0, 1

i32, i64, u32, f64, vp = C.c_int32, C.c_int64, C.c_uint32, C.c_double, C.c_void_p
P = C.POINTER


class MeshDesc(C.Structure):
    _fields_ = [
        ("dim", i32), ("fe_degree", i32), ("n_q1d", i32), ("n_q1d_face", i32),
        ("n_verts", i64), ("verts", P(f64)), ("n_cells", i64), ("cell_verts", P(i32)),
        ("n_polytopes", i32), ("poly_subcell_ptr", P(i64)), ("poly_subcell_idx", P(i32)),
        ("bbox", P(f64)), ("dof_block", P(i32)),
        ("n_ifaces", i32), ("iface_polyA", P(i32)), ("iface_polyB", P(i32)), ("iface_sub_ptr", P(i64)),
        ("sub_cell", P(i32)), ("sub_face", P(i32)), ("sub_sigma", P(f64)),
        ("n_block_rows", i32), ("brow_ptr", P(i64)), ("bcol_idx", P(i32)),
        ("n_owned_polytopes", i32), ("fe_kind", i32),
    ]


class Coefficients(C.Structure):
    _fields_ = [("stiffness", f64), ("mass", f64)]


class LocalInfo(C.Structure):
    _fields_ = [("n_owned", i32), ("n_ghost", i32), ("owned_global_block", P(i32)), ("ghost_global_block", P(i32)),
                ("ghost_owner", P(i32)), ("local_poly_global", P(i32))]


class FlattenParams(C.Structure):
    _fields_ = [("penalty_constant", f64), ("h_rule", i32), ("h_const", f64), ("visit_rule", i32)]


SIGNATURES = {
    "pd_last_error": (C.c_char_p, []),
    "pd_device_count": (C.c_int, []),
    "pd_create": (C.c_int, [P(MeshDesc), P(vp)]),
    "pd_destroy": (C.c_int, [vp]),
    "pd_upload": (C.c_int, [vp, P(MeshDesc)]),
    "pd_set_stream": (C.c_int, [vp, vp]),
    "pd_synchronize": (C.c_int, [vp]),
    "pd_build_quadrature": (C.c_int, [vp]),
    "pd_invalidate_quadrature": (C.c_int, [vp]),
    "pd_assemble": (C.c_int, [vp, u32, P(Coefficients)]),
    "pd_n_dofs": (i64, [vp]),
    "pd_n_source_dofs": (i64, [vp]),
    "pd_nnz": (i64, [vp]),
    "pd_n_dofs_per_cell": (i32, [vp]),
    "pd_matrix_values_device": (C.c_int, [vp, P(vp)]),
    "pd_matrix_values_to_host": (C.c_int, [vp, vp]),
    "pd_matrix_values_to_host_async": (C.c_int, [vp, vp]),
    "pd_matrix_pattern_to_host": (C.c_int, [vp, vp, vp]),
    "pd_vmult": (C.c_int, [vp, C.c_int, vp, vp]),
    "pd_set_operator": (C.c_int, [vp, u32, P(Coefficients)]),
    "pd_matrix_free_available": (C.c_int, [vp]),
    "pd_force_generic_matrix_free": (C.c_int, [vp, C.c_int]),
    "pd_mapped_fine_available": (C.c_int, [vp]),
    "pd_fine_kernel_last": (C.c_int, [vp]),
    "pd_transfer_create": (C.c_int, [vp, vp, C.c_void_p, C.POINTER(vp)]),
    "pd_transfer_create_to_cells": (C.c_int, [vp, C.POINTER(vp)]),
    "pd_transfer_destroy": (None, [vp]),
    "pd_transfer_m": (C.c_int64, [vp]),
    "pd_transfer_n": (C.c_int64, [vp]),
    "pd_transfer_prolongate": (C.c_int, [vp, vp, vp, C.c_int]),
    "pd_transfer_restrict": (C.c_int, [vp, vp, vp, C.c_int]),
    "pd_peer_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(vp)]),
    "pd_peer_handle_bytes": (C.c_int, []),
    "pd_peer_export": (C.c_int, [vp, C.c_void_p]),
    "pd_peer_connect": (C.c_int, [vp, C.c_void_p]),
    "pd_peer_exchange": (C.c_int, [vp, C.c_void_p]),
    "pd_peer_vmult": (C.c_int, [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "pd_peer_status": (C.c_int, [vp]),
    "pd_peer_fused": (C.c_int, [vp]),
    "pd_peer_allreduce": (C.c_int, [vp, C.c_void_p, C.c_int]),
    "pd_estimate_lambda_max_sharded": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "pd_chebyshev_smooth_sharded": (C.c_int, [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int]),
    "pd_cg_solve_sharded": (C.c_int, [vp, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "pd_peer_destroy": (None, [vp]),
    "pd_n_quadrature_points": (C.c_int64, [vp, C.c_int]),
    "pd_quadrature_device": (C.c_int, [vp] + [C.POINTER(C.c_void_p)] * 5),
    "pd_quadrature_to_host": (C.c_int, [vp] + [C.c_void_p] * 5),
    "pd_assemble_rhs": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "pd_error_norms": (C.c_int, [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pd_vmult_add": (C.c_int, [vp, C.c_int, vp, vp]),
    "pd_vmult_host": (C.c_int, [vp, C.c_int, vp, vp]),
    "pd_diagonal_inverse": (C.c_int, [vp, vp]),
    "pd_diagonal_inverse_of": (C.c_int, [vp, C.c_int, vp]),
    "pd_cg_solve": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, f64, C.c_int, P(C.c_int), P(f64)]),
    "pd_estimate_lambda_max": (C.c_int, [vp, C.c_int, C.c_int, P(f64)]),
    "pd_chebyshev_smooth": (C.c_int, [vp, C.c_int, C.c_int, f64, f64, vp, vp, C.c_int]),
    "pd_reinit_n_points": (i64, [vp, i32]),
    "pd_reinit_iface_n_points": (i64, [vp, i32]),
    "pd_reinit_polytope": (C.c_int, [vp, i32, vp, vp, vp, vp]),
    "pd_reinit_face": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, vp]),
    "pd_reinit_interface": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
    "pd_agglomerated_quadrature": (C.c_int, [vp, i32, vp, vp, vp]),
    "pd_fe_evaluate": (C.c_int, [i32, i32, i32, i64, vp, vp, vp]),
    "pd_copy_array": (C.c_int, [vp, C.c_char_p, vp, P(i64)]),
    "pd_assembly_path": (C.c_int, [vp]),
    "pd_tensor_path_stats": (C.c_int, [vp, vp]),
    "pd_launch_count": (i64, [vp]),
    "pd_last_kernel_ms": (C.c_int, [vp, P(C.c_float)]),
    "pd_quadrature_rule_1d": (C.c_int, [C.c_int, vp, vp]),
    "pd_dgq_nodes_1d": (C.c_int, [C.c_int, vp]),
    "pdh_grid_create_structured": (C.c_int, [i32, vp, vp, vp, i32, P(vp)]),
    "pdh_grid_create": (C.c_int, [i32, i64, vp, i64, vp, vp, P(vp)]),
    "pdh_grid_destroy": (C.c_int, [vp]),
    "pdh_grid_n_cells": (i64, [vp]),
    "pdh_grid_n_verts": (i64, [vp]),
    "pdh_grid_set_vertices": (C.c_int, [vp, vp]),
    "pdh_grid_get_arrays": (C.c_int, [vp, vp, vp, vp]),
    "pdh_handler_create": (C.c_int, [vp, P(vp)]),
    "pdh_handler_destroy": (C.c_int, [vp]),
    "pdh_define_agglomerate": (i32, [vp, vp, i32]),
    "pdh_initialize_fe_values": (C.c_int, [vp, i32, i32]),
    "pdh_distribute_agglomerated_dofs": (C.c_int, [vp, i32, i32]),
    "pdh_n_polytopes": (i32, [vp]),
    "pdh_n_dofs": (i64, [vp]),
    "pdh_n_dofs_per_cell": (i32, [vp]),
    "pdh_master_cell": (i32, [vp, i32]),
    "pdh_n_background_cells": (i32, [vp, i32]),
    "pdh_get_agglomerate": (C.c_int, [vp, i32, vp]),
    "pdh_n_faces": (u32, [vp, i32]),
    "pdh_at_boundary": (i32, [vp, i32, u32]),
    "pdh_neighbor": (i32, [vp, i32, u32]),
    "pdh_neighbor_of_agglomerated_neighbor": (u32, [vp, i32, u32]),
    "pdh_interface": (i32, [vp, i32, u32, vp, vp, i32]),
    "pdh_get_dof_indices": (C.c_int, [vp, i32, vp]),
    "pdh_bounding_box": (C.c_int, [vp, i32, vp, vp]),
    "pdh_diameter": (f64, [vp, i32]),
    "pdh_volume": (f64, [vp, i32]),
    "pdh_sparsity_nnz": (i64, [vp]),
    "pdh_create_agglomeration_sparsity_pattern": (C.c_int, [vp, vp, vp]),
    "pdh_flatten": (C.c_int, [vp, P(FlattenParams), P(MeshDesc)]),
    "pdh_face_work_item": (C.c_int, [vp, i32, u32, P(i32), P(i32)]),
    "pdh_real_to_unit": (C.c_int, [vp, i32, i64, vp, vp]),
    "pdh_unit_to_real": (C.c_int, [vp, i32, i64, vp, vp]),
    "pdh_define_agglomerates": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "pdh_polytope_graph": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdh_partition_graph": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "pdh_flatten_local": (C.c_int, [vp, P(FlattenParams), vp, i32, P(MeshDesc), P(LocalInfo)]),
    "pdh_create_device": (C.c_int, [vp, P(FlattenParams), P(vp)]),
}

_lib = None


class PolydealError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def lib():
    """Load the C-ABI library.  Raises if it has not been built (run
    `python -c 'import __graft_entry__ as g; g.build()'` or `make -C polydeal_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the polydeal_b200 CUDA library has not been built; "
                "there is no fallback path (make -C polydeal_b200/csrc)"
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(code):
    """Negative = error (raises); positive = a status the caller interprets (PD_NOT_CONVERGED)."""
    if code < 0:
        raise PolydealError(code, lib().pd_last_error().decode())
    return code
