"""Sharding of the hot path over GPUs: one process per GPU, polytopes as the unit.

Reference: MPI domain decomposition in which an agglomerate never straddles ranks
(/root/reference/source/agglomeration_handler.cc:83-87); ghost metadata travels once
(setup_ghost_polytopes / exchange_interface_values, :531-618, 1026-1091) and the vector
ghost values travel in every vmult (LinearAlgebra::distributed::Vector inside
MatrixFree::loop, include/utils.h:466-472; Trilinos import for the matrix-based vmult).

Here:
 * assembly is owner-computes-rows: every rank evaluates its cut interfaces itself from the
   ghost polytope's bounding box + DoF block (one-time ghost geometry), so there is NO
   data-path collective in assembly and no reverse matrix communication;
 * vmult has one exchange step per apply: the coefficient blocks of the ghost polytopes,
   grouped send/recv per peer (torch.distributed P2P batch = ncclGroupStart/ncclSend/ncclRecv/
   ncclGroupEnd over NVLink), received straight into the ghost section of the source vector.

Every rank holds the (host-side, replicated) AgglomerationHandler of the whole mesh, so the
exchange plan is computed locally without any setup communication.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K
from .handler import AgglomerationHandler, SIPOperator, _ptr


def partition_by_blocks(ah: AgglomerationHandler, n_ranks: int) -> np.ndarray:
    """owner[p]: contiguous, equally sized ranges of the global DoF-block order (a slab
    decomposition on lexicographic grids, a space-filling-curve one on Morton grids)."""
    np_ = ah.n_polytopes
    n = ah.n_dofs_per_cell
    block = np.array([int(ah.get_dof_indices(p)[0]) // n for p in range(np_)], dtype=np.int64)
    owner = np.empty(np_, dtype=np.int32)
    owner[:] = (block * n_ranks) // np_
    return owner


def partition_by_coordinate(ah: AgglomerationHandler, n_ranks: int, axis: int) -> np.ndarray:
    """owner[p] by the bounding-box centre along `axis`: equal-count slabs."""
    np_ = ah.n_polytopes
    centre = np.array([0.5 * (ah.bbox(p)[0][axis] + ah.bbox(p)[1][axis]) for p in range(np_)])
    order = np.argsort(centre, kind="stable")
    owner = np.empty(np_, dtype=np.int32)
    owner[order] = (np.arange(np_) * n_ranks) // np_
    return owner


def partition_by_metis(ah: AgglomerationHandler, n_ranks: int) -> np.ndarray:
    """owner[p] from a METIS partition of the polytope adjacency graph, vertex weight = number of sub-cells
    (volume work), edge weight = number of shared sub-faces (ghost traffic): SURVEY 8e."""
    from .handler import partition_graph

    xadj, adjncy, vw, ew = ah.polytope_graph()
    return partition_graph(xadj, adjncy, n_ranks, vw, ew).astype(np.int32)


class LocalPart:
    """The local descriptor of one rank + the index maps of the halo exchange.

    `desc` points into storage of the host handler that the next flatten / LocalPart on the same handler reuses
    (pdh_flatten_local): create the SIPOperator (pd_create copies to the device) before building another rank's part
    from the same handler.  The index maps (owned_global_block, ghost_*, send_blocks, counts) are copies."""

    def __init__(self, ah: AgglomerationHandler, owner: np.ndarray, rank: int, penalty_constant=-1.0,
                 h_rule=K.H_DIAMETER_OF_VISITOR, h_const=1.0, visit_rule=K.VISIT_BY_ID):
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        assert owner.shape == (ah.n_polytopes,)
        self.ah, self.owner, self.rank = ah, owner, rank
        self.n = ah.n_dofs_per_cell
        prm = K.FlattenParams(penalty_constant, h_rule, h_const, visit_rule)
        self.desc, info = K.MeshDesc(), K.LocalInfo()
        K.check(K.lib().pdh_flatten_local(ah._h, C.byref(prm), _ptr(owner), rank, C.byref(self.desc), C.byref(info)))
        self.n_owned, self.n_ghost = info.n_owned, info.n_ghost
        as_np = lambda p, m: np.ctypeslib.as_array(p, (m,)).copy() if m else np.zeros(0, dtype=np.int32)
        self.owned_global_block = as_np(info.owned_global_block, self.n_owned)
        self.ghost_global_block = as_np(info.ghost_global_block, self.n_ghost)
        self.ghost_owner = as_np(info.ghost_owner, self.n_ghost)
        self.local_poly_global = as_np(info.local_poly_global, self.n_owned + self.n_ghost)
        # ---- exchange plan, computed without communication from the replicated handler ----
        n_ranks = int(owner.max()) + 1
        self.n_ranks = n_ranks
        # what I receive: my ghost section is grouped by owner rank
        self.recv_counts = np.bincount(self.ghost_owner, minlength=n_ranks).astype(np.int64)
        # what I send to peer s: my owned polytopes adjacent to a polytope owned by s, by global block
        xadj, adjncy, _, _ = ah.polytope_graph()
        own = self.local_poly_global[: self.n_owned].astype(np.int64)
        deg = xadj[own + 1] - xadj[own]
        local_block = np.ctypeslib.as_array(self.desc.dof_block, (self.n_owned + self.n_ghost,))[: self.n_owned].astype(np.int64)
        src_local = np.repeat(local_block, deg)  # local block (row) of the owned end of every edge
        idx = np.repeat(xadj[own], deg) + (np.arange(int(deg.sum()), dtype=np.int64) - np.repeat(np.cumsum(deg) - deg, deg))
        nb_owner = owner[adjncy[idx]]
        self.send_blocks = []  # per peer: local owned block indices, in the order the peer stores them (global block)
        for s in range(n_ranks):
            if s == rank:
                self.send_blocks.append(np.zeros(0, dtype=np.int64))
                continue
            self.send_blocks.append(np.unique(src_local[nb_owner == s]))  # owned_global_block ascends with the local block
        self.send_counts = np.array([len(b) for b in self.send_blocks], dtype=np.int64)

    @property
    def n_owned_dofs(self):
        return self.n_owned * self.n

    @property
    def n_local_dofs(self):
        return (self.n_owned + self.n_ghost) * self.n

    def owned_global_dofs(self):
        return (self.owned_global_block[:, None] * self.n + np.arange(self.n)[None, :]).ravel()

    def ghost_global_dofs(self):
        return (self.ghost_global_block[:, None] * self.n + np.arange(self.n)[None, :]).ravel()


def exchange_ghost_values(part: LocalPart, x_full, group=None):
    """update_ghost_values(): fill the ghost section of `x_full` (torch tensor, length
    (n_owned + n_ghost) * n, CPU for gloo or CUDA for nccl) from the owning ranks.

    One pack kernel (index_select with the concatenated per-peer send lists, cached on the
    device) and one exchange that receives straight into the ghost section: a single
    all_to_all_single on NCCL (one grouped ncclSend/ncclRecv per peer), P2P batches on gloo."""
    import torch
    import torch.distributed as dist

    n = part.n
    xo = x_full[: part.n_owned * n].view(part.n_owned, n)
    xg = x_full[part.n_owned * n:].view(part.n_ghost, n)
    cache = part.__dict__.setdefault("_xchg", {})
    key = (x_full.device.type, x_full.device.index)
    if key not in cache:
        idx = np.concatenate(part.send_blocks) if part.n_ranks else np.zeros(0, dtype=np.int64)
        cache[key] = (torch.as_tensor(idx, device=x_full.device),
                      torch.empty((len(idx), n), dtype=x_full.dtype, device=x_full.device))
    idx, sendbuf = cache[key]
    if len(idx):
        torch.index_select(xo, 0, idx, out=sendbuf)  # pack
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(xg, sendbuf, output_split_sizes=part.recv_counts.tolist(),
                               input_split_sizes=part.send_counts.tolist(), group=group)
        return x_full
    ops, off_r, off_s = [], 0, 0
    for s in range(part.n_ranks):
        cnt = int(part.recv_counts[s])
        if cnt:
            ops.append(dist.P2POp(dist.irecv, xg[off_r:off_r + cnt], s, group))
        off_r += cnt
    for s in range(part.n_ranks):
        cnt = int(part.send_counts[s])
        if cnt:
            ops.append(dist.P2POp(dist.isend, sendbuf[off_s:off_s + cnt], s, group))
        off_s += cnt
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return x_full


class PeerExchange:
    """update_ghost_values() over NVLink peer memory (pd_peer_*, csrc/pd_peer.cu): every rank
    publishes the blocks its neighbours need into an IPC-mapped buffer and pulls its ghost blocks
    with plain loads, ordered by an epoch-flag handshake -- two small kernels on the operator's
    stream per apply, no NCCL call on the data path.  Setup (once) all-gathers the send counts
    and the CUDA IPC handles through `torch.distributed`."""

    def __init__(self, part: LocalPart, op: SIPOperator, group=None):
        import torch.distributed as dist

        self.part, self.op = part, op
        world, rank = part.n_ranks, part.rank
        if dist.get_world_size(group) != world:
            raise ValueError("the partition has a different number of ranks than the process group")
        send_ptr = np.zeros(world + 1, dtype=np.int64)
        send_ptr[1:] = np.cumsum(part.send_counts)
        recv_ptr = np.zeros(world + 1, dtype=np.int64)
        recv_ptr[1:] = np.cumsum(part.recv_counts)
        send_blocks = np.ascontiguousarray(np.concatenate(part.send_blocks) if world else np.zeros(0), dtype=np.int32)
        all_ptr = [None] * world
        dist.all_gather_object(all_ptr, send_ptr.tolist(), group=group)
        remote_off = np.array([all_ptr[s][rank] for s in range(world)], dtype=np.int64)
        for s in range(world):  # what s sends me is what I expect from s
            assert all_ptr[s][rank + 1] - all_ptr[s][rank] == part.recv_counts[s]
        self._h = C.c_void_p()
        K.check(K.lib().pd_peer_create(op._h, rank, world, _ptr(send_ptr), _ptr(send_blocks), _ptr(recv_ptr),
                                       _ptr(remote_off), C.byref(self._h)))
        nb = K.lib().pd_peer_handle_bytes()
        mine = (C.c_char * nb)()
        K.check(K.lib().pd_peer_export(self._h, mine))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine), group=group)
        blob = b"".join(handles)
        K.check(K.lib().pd_peer_connect(self._h, blob))
        dist.barrier(group=group)  # every rank has mapped its neighbours before anyone publishes

    def exchange(self, x_full):
        self.op._bind_torch_stream()
        K.check(K.lib().pd_peer_exchange(self._h, C.c_void_p(x_full.data_ptr())))
        return x_full

    def vmult(self, dst, x_full, mode=K.VMULT_BLOCK_CSR, add=False):
        """exchange + vmult in one call (pd_peer_vmult): overlaps the exchange with the cells that need
        no ghost data when the fine-mesh stencil kernel applies."""
        self.op._bind_torch_stream()
        K.check(K.lib().pd_peer_vmult(self._h, mode, C.c_void_p(x_full.data_ptr()), C.c_void_p(dst.data_ptr()), int(add)))
        return dst

    def ok(self):
        return K.lib().pd_peer_status(self._h) == 0

    @property
    def fused(self):
        """pd_peer_vmult(MATRIX_FREE) runs the fused fine-mesh apply (one kernel, ghost cells read from the peers)."""
        return bool(K.lib().pd_peer_fused(self._h))

    @property
    def fused_tiles(self):
        """Tiles of the fused apply's plan (blocks of the curve, split where their halo exceeds the kernel's gather); 0: not fused."""
        return int(K.lib().pd_peer_fused(self._h))

    def allreduce(self, scalars):
        """In-place sum over the ranks of a CUDA float64 tensor of at most 4 entries."""
        assert scalars.is_cuda and scalars.numel() <= 4
        self.op._bind_torch_stream()
        K.check(K.lib().pd_peer_allreduce(self._h, C.c_void_p(scalars.data_ptr()), scalars.numel()))
        return scalars

    def estimate_lambda_max(self, n_iterations=20, mode=K.VMULT_BLOCK_CSR):
        lam = C.c_double(0.0)
        K.check(K.lib().pd_estimate_lambda_max_sharded(self._h, mode, n_iterations, C.byref(lam)))
        return lam.value

    def chebyshev_smooth(self, x_full, b, degree, lambda_max, smoothing_range=20.0, zero_initial_guess=True,
                         mode=K.VMULT_BLOCK_CSR):
        """PreconditionChebyshev on the sharded operator; x_full has the (owned + ghost) length."""
        self.op._bind_torch_stream()
        K.check(K.lib().pd_chebyshev_smooth_sharded(self._h, mode, degree, lambda_max, smoothing_range,
                                                    C.c_void_p(b.data_ptr()), C.c_void_p(x_full.data_ptr()),
                                                    int(zero_initial_guess)))
        return x_full

    def cg_solve(self, x, b, max_iter=1000, rel_tol=1e-10, jacobi=True, mode=K.VMULT_BLOCK_CSR):
        """SolverCG on the sharded operator, device resident on every rank (pd_cg_solve_sharded);
        x, b hold the owned DoFs.  Returns (iterations, global relative residual)."""
        it, rr = C.c_int(0), C.c_double(0.0)
        self.op._bind_torch_stream()
        rc = K.check(K.lib().pd_cg_solve_sharded(self._h, mode, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), max_iter,
                                                 rel_tol, int(jacobi), C.byref(it), C.byref(rr)))
        self.last_cg_converged = rc != K.PD_NOT_CONVERGED
        return it.value, rr.value

    def close(self):
        if getattr(self, "_h", None) and K._lib is not None:
            K._lib.pd_peer_destroy(self._h)
            self._h = None

    __del__ = close


class DistributedSIPOperator:
    """One rank's share of the SIP operator: local assembly with ghost interfaces and a
    vmult that exchanges the ghost-polytope coefficients first."""

    def __init__(self, ah: AgglomerationHandler, owner, rank: int, group=None, **flatten_kw):
        self.part = LocalPart(ah, owner, rank, **flatten_kw)
        self.op = SIPOperator(self.part.desc, keepalive=(ah, self.part))
        self.group = group
        self._x_full = None
        self.peer = None

    def enable_peer_exchange(self):
        """Switch the ghost exchange of vmult from NCCL to NVLink peer memory (collective call)."""
        self.peer = PeerExchange(self.part, self.op, self.group)
        return self

    def assemble(self, flags=K.ASSEMBLE_ALL, stiffness=1.0, mass=0.0):
        self.op.assemble(flags, stiffness, mass)

    def m(self):
        return self.part.n_owned_dofs

    def vmult(self, dst, src, mode=K.VMULT_BLOCK_CSR, exchange=True):
        """dst (owned DoFs) = A_local [src ; ghosts].  `src` holds the owned DoFs."""
        import torch

        p = self.part
        if self._x_full is None or self._x_full.device != src.device:
            self._x_full = torch.empty(p.n_local_dofs, dtype=torch.float64, device=src.device)
        self.op._bind_torch_stream()  # the copy above, the exchange and the apply all on torch's current stream
        self._x_full[: p.n_owned_dofs].copy_(src)
        if exchange and p.n_ranks > 1:
            if self.peer is not None:
                return self.peer.vmult(dst, self._x_full, mode)
            exchange_ghost_values(p, self._x_full, self.group)
        self.op.vmult_ptr(dst.data_ptr(), self._x_full.data_ptr(), mode)
        return dst
