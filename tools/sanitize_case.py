#!/usr/bin/env python3
"""Small end-to-end invocation of every kernel for `compute-sanitizer --tool memcheck`:
assembly (2-D p=3 random partition, 3-D p=2 with mass, 3-D p=3), block-CSR vmult, diagonal inverse,
fine-mesh matrix-free vmult (line and tiled kernels), and one rank of a sharded problem (ghost interfaces)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import polydeal_b200 as pdl
from oracle import pyoracle as po
from pd_helpers import groups_for, product_handler
from polydeal_b200 import distributed as pdd


def run(dim, n, shape, p, mass=0.0, fine=False, shard=False, order=1, kernel=None):
    if kernel:  # fine-mesh operator: force the tiled / line kernel (pd_finemesh.cu)
        os.environ["PD_FINE_KERNEL"] = kernel
    ogrid = po.Grid(dim, n, 0.0, 1.0, order)
    groups = [[c] for c in range(ogrid.n_cells)] if shape == "singletons" else groups_for(shape, dim, n, ogrid, 3)
    _, pah = product_handler(ogrid, groups, p, p + 1)
    if shard:
        kw = dict(penalty_constant=p * (p + 1.0), h_rule=pdl.H_NORMAL_EXTENT) if fine else {}
        part = pdd.LocalPart(pah, pdd.partition_by_blocks(pah, 2), 1, **kw)
        op = pdl.SIPOperator(part.desc, keepalive=(pah, part))
    elif fine:
        op = pdl.SIPOperator(pah.flatten(penalty_constant=p * (p + 1.0), h_rule=pdl.H_NORMAL_EXTENT), keepalive=pah)
    else:
        op = pdl.SIPOperator(pah.flatten(), keepalive=pah)
    op.assemble(mass=mass)
    x = torch.from_numpy(np.sin(0.37 * np.arange(op.n_source_dofs))).cuda()
    y = torch.empty(op.m(), dtype=torch.float64, device="cuda")
    op.vmult_ptr(y.data_ptr(), x.data_ptr())
    d = torch.empty(op.m(), dtype=torch.float64, device="cuda")
    op.get_matrix_diagonal_inverse(d)
    if op.matrix_free_available:
        op.set_operator()
        op.vmult_ptr(y.data_ptr(), x.data_ptr(), mode=pdl.VMULT_MATRIX_FREE)
    op.synchronize()
    assert torch.isfinite(y).all() and np.isfinite(op.values()).all()
    print("ok", dim, n, shape, p, "mass" if mass else "", "fine" if fine else "", "shard" if shard else "", flush=True)


run(2, 8, "random6", 3)
run(3, 4, "random5", 2, mass=0.5)
run(3, 4, "blocks2", 3)
run(3, 4, "singletons", 2, fine=True)
run(2, 6, "singletons", 1, fine=True)
run(3, 4, "random6", 1, shard=True)
# the tiled fine-mesh kernel: TMA bulk copies of halo rows (cells of odd and even alignment, the last cell of the
# vector), the contiguous own range (Morton numbering) and the cp.async path (other numberings, ragged tiles, ghosts)
run(3, 8, "singletons", 2, fine=True, order=0, kernel="tile")
run(3, (5, 4, 3), "singletons", 2, fine=True, order=1, kernel="tile")
run(2, (9, 7), "singletons", 4, fine=True, order=1, kernel="tile")
run(3, 4, "singletons", 1, fine=True, order=0, kernel="tile")
run(3, (4, 4, 8), "singletons", 2, fine=True, shard=True, order=1, kernel="tile")
print("SANITIZE CASE DONE")
