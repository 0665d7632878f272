"""polydeal_b200 -- B200-native SIP-DG assembly and operator apply over
agglomerated polytopes, behind the reference's AgglomerationHandler-shaped API.

The compute path is the CUDA shared library (polydeal_b200/lib/libpolydeal_b200.so,
C ABI in include/polydeal_b200.h).  This package is only the Python-side mirror of
the reference interface; it never computes anything itself and raises if the
library or a GPU is missing.
"""
from . import _capi
from ._capi import (ASSEMBLE_ALL, ASSEMBLE_BOUNDARY, ASSEMBLE_INTERIOR, ASSEMBLE_VOLUME, H_CONSTANT,
                    H_DIAMETER_OF_VISITOR, H_MAX_INVERSE_DIAMETER, H_NORMAL_EXTENT, INVALID_UINT, VISIT_BY_ID,
                    VISIT_BY_INDEX, VMULT_BLOCK_CSR, VMULT_MATRIX_FREE, VMULT_MAPPED_FINE, PolydealError)
from .handler import (AgglomerationHandler, FEValuesTables, Grid, SIPOperator, Transfer, assemble_dg_matrix, fe_evaluate,
                      metis_agglomerates, partition_graph)

FE_DGQ, FE_AGGLODGP = 0, 1

__all__ = [
    "AgglomerationHandler", "FEValuesTables", "fe_evaluate", "Grid", "SIPOperator", "Transfer", "assemble_dg_matrix", "metis_agglomerates", "partition_graph", "PolydealError", "FE_DGQ", "FE_AGGLODGP",
    "ASSEMBLE_ALL", "ASSEMBLE_BOUNDARY", "ASSEMBLE_INTERIOR", "ASSEMBLE_VOLUME",
    "H_CONSTANT", "H_DIAMETER_OF_VISITOR", "H_MAX_INVERSE_DIAMETER", "H_NORMAL_EXTENT",
    "VISIT_BY_ID", "VISIT_BY_INDEX", "VMULT_BLOCK_CSR", "VMULT_MATRIX_FREE", "VMULT_MAPPED_FINE", "INVALID_UINT",
]
