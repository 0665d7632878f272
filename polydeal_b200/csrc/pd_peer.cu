// -----------------------------------------------------------------------------
// pd_peer.cu -- the ghost exchange of the sharded vmult over NVLink peer memory.
//
// Reference: LinearAlgebra::distributed::Vector::update_ghost_values() inside
// MatrixFree::loop (include/utils.h:466-472) / the Trilinos import of the matrix-based
// vmult: the coefficient blocks of the ghost polytopes travel before every apply.
//
// With one process per GPU on an NVSwitch node there is no need for a message: every rank
// PUBLISHES the blocks its neighbours need into a buffer that they have mapped through CUDA
// IPC, and every rank PULLS its ghost blocks with ordinary loads over NVLink.  Ordering is a
// flag handshake in the same peer memory (no host, no NCCL call on the data path):
//   publish(e)  pack my send blocks into export[e & 1]; the last CTA to finish fences
//               (system scope) and stores the epoch e into flags[my rank] ON every neighbour
//   pull(e)     every CTA spins (bounded) until flags[owner] >= e for the ranks I receive from,
//               then the grid copies their segments of export[e & 1] into my ghost section
// Double buffering removes the write-after-read hazard: a rank cannot publish epoch e + 2 into
// the buffer a neighbour still reads for epoch e, because its own pull of epoch e + 1 waited
// for that neighbour's publish(e + 1), which follows the neighbour's pull(e) in stream order
// (adjacency is symmetric).
// Two small kernels per apply on the handle's stream, a few microseconds, instead of a NCCL
// all-to-all whose latency dominates a latency-bound vmult.
// A spin that does not see its flag within ~2 s raises the error word (reported by
// pd_peer_status) and moves on instead of hanging the GPU.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"
#include "pd_host.hpp"

#include <cstring>
#include <memory>
#include <vector>

struct pd_peer
{
  pd_handle *h     = nullptr;
  int        rank  = 0, world = 1;
  int64_t    epoch = 0;
  // plan
  std::vector<int64_t> send_ptr, recv_ptr, remote_off; // in blocks
  pd::DevBuf<int32_t>  send_blocks, send_peer_of_block, recv_owner_of_block;
  pd::DevBuf<int64_t>  recv_src_block; // position of every ghost block inside its owner's export buffer
  int64_t              n_send = 0, n_recv = 0;
  // IPC memory of this rank
  double             *export_buf = nullptr; // [n_send][2 parities][n]
  unsigned long long *flags      = nullptr; // [world] epochs written by the peers; [world] = error word
  unsigned int       *counter    = nullptr; // last-CTA ticket
  // mapped memory of the peers
  std::vector<double *>             peer_export;
  std::vector<unsigned long long *> peer_flags;
  pd::DevBuf<double *>              d_peer_export;
  pd::DevBuf<unsigned long long *>  d_peer_flags;
  pd::DevBuf<int32_t>               d_neighbours, d_owners; // ranks I send to / receive from
  int                               n_neighbours = 0, n_owners = 0;
  bool                              connected    = false;
};

namespace pd
{
  namespace
  {
    __global__ void
    k_peer_publish(const double *x, const int32_t *send_blocks, const int64_t n_send, const int n, const int parity, double *exp,
                   unsigned long long *const *peer_flags, const int32_t *neighbours, const int n_neighbours, const int rank,
                   const unsigned long long epoch, unsigned int *counter)
    {
      const int64_t total = n_send * n;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t b = i / n;
          exp[(b * 2 + parity) * n + (i - b * n)] = x[(int64_t)send_blocks[b] * n + (i - b * n)];
        }
      // last CTA to finish announces the epoch to every neighbour
      __shared__ bool last;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0)
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
      __syncthreads();
      if (last)
        {
          __threadfence_system();
          for (int k = threadIdx.x; k < n_neighbours; k += blockDim.x)
            {
              volatile unsigned long long *f = peer_flags[neighbours[k]] + rank;
              *f                             = epoch;
            }
          if (threadIdx.x == 0)
            *counter = 0;
        }
    }

    __global__ void
    k_peer_pull(double *ghost, const int32_t *owner_of_block, const int64_t *src_block, const int64_t n_recv, const int n,
                double *const *peer_export, const int parity, volatile unsigned long long *flags, const int world,
                const unsigned long long epoch, const int32_t *owners, const int n_owners)
    {
      // every CTA waits (bounded) until all the ranks I receive from have published this epoch,
      // then the whole grid copies element-wise: wide, coalesced loads over NVLink
      if ((int)threadIdx.x < n_owners)
        {
          const int       owner = owners[threadIdx.x];
          const long long t0    = clock64();
          while (flags[owner] < epoch)
            {
              if (clock64() - t0 > 4000000000ll) // ~2 s: the peer is gone; report, do not hang
                {
                  flags[world] = 1ull;
                  break;
                }
              __nanosleep(64);
            }
          __threadfence_system();
        }
      __syncthreads();
      const int64_t total = n_recv * n;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t b = i / n;
          ghost[i]        = __ldcg(peer_export[owner_of_block[b]] + (src_block[b] * 2 + parity) * n + (i - b * n));
        }
    }
  } // namespace
} // namespace pd

namespace pd
{
  pd_peer *
  peer_create(pd_handle *h, int rank, int world, const int64_t *send_ptr, const int32_t *send_blocks,
              const int64_t *recv_ptr, const int64_t *remote_offset)
  {
    {
      if (!h || !send_ptr || !recv_ptr || !remote_offset || rank < 0 || rank >= world)
        throw Error(PD_ERR_INVALID, "pd_peer_create: bad argument");
      std::unique_ptr<pd_peer> p(new pd_peer);
      p->h     = h;
      p->rank  = rank;
      p->world = world;
      p->send_ptr.assign(send_ptr, send_ptr + world + 1);
      p->recv_ptr.assign(recv_ptr, recv_ptr + world + 1);
      p->remote_off.assign(remote_offset, remote_offset + world);
      p->n_send = send_ptr[world];
      p->n_recv = recv_ptr[world];
      if (p->n_recv != (int64_t)(h->np - h->np_own))
        throw Error(PD_ERR_INVALID, "pd_peer_create: the receive plan does not cover the ghost polytopes of the handle");
      auto put = [](auto &buf, const auto &v) {
        buf.alloc(v.size());
        if (!v.empty())
          PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
      };
      std::vector<int32_t> sb(send_blocks, send_blocks + p->n_send), owner((size_t)p->n_recv), nbrs, owners;
      std::vector<int64_t> src((size_t)p->n_recv);
      for (int s = 0; s < world; ++s)
        {
          for (int64_t k = recv_ptr[s]; k < recv_ptr[s + 1]; ++k)
            {
              owner[(size_t)k] = s;
              src[(size_t)k]   = remote_offset[s] + (k - recv_ptr[s]);
            }
          if (send_ptr[s + 1] > send_ptr[s])
            nbrs.push_back(s);
          if (recv_ptr[s + 1] > recv_ptr[s])
            owners.push_back(s);
        }
      put(p->send_blocks, sb);
      put(p->recv_owner_of_block, owner);
      put(p->recv_src_block, src);
      put(p->d_neighbours, nbrs);
      put(p->d_owners, owners);
      p->n_neighbours = (int)nbrs.size();
      p->n_owners     = (int)owners.size();
      // IPC-exportable allocations (plain cudaMalloc)
      const size_t exp_count = (size_t)std::max<int64_t>(1, 2 * p->n_send * h->n);
      PD_CUDA(cudaMalloc((void **)&p->export_buf, exp_count * sizeof(double)));
      PD_CUDA(cudaMemset(p->export_buf, 0, exp_count * sizeof(double)));
      PD_CUDA(cudaMalloc((void **)&p->flags, (size_t)(world + 1) * sizeof(unsigned long long)));
      PD_CUDA(cudaMemset(p->flags, 0, (size_t)(world + 1) * sizeof(unsigned long long)));
      PD_CUDA(cudaMalloc((void **)&p->counter, sizeof(unsigned int)));
      PD_CUDA(cudaMemset(p->counter, 0, sizeof(unsigned int)));
      return p.release();
    }
  }

  int
  peer_handle_bytes()
  {
    return 2 * (int)sizeof(cudaIpcMemHandle_t);
  }

  void
  peer_export(pd_peer *p, void *handles_out)
  {
    {
      if (!p || !handles_out)
        throw Error(PD_ERR_INVALID, "pd_peer_export: null argument");
      cudaIpcMemHandle_t hd[2];
      PD_CUDA(cudaIpcGetMemHandle(&hd[0], p->export_buf));
      PD_CUDA(cudaIpcGetMemHandle(&hd[1], p->flags));
      std::memcpy(handles_out, hd, sizeof hd);
    }
  }

  void
  peer_connect(pd_peer *p, const void *all_handles)
  {
    {
      if (!p || !all_handles)
        throw Error(PD_ERR_INVALID, "pd_peer_connect: null argument");
      const auto *hd = static_cast<const cudaIpcMemHandle_t *>(all_handles);
      p->peer_export.assign(p->world, nullptr);
      p->peer_flags.assign(p->world, nullptr);
      for (int s = 0; s < p->world; ++s)
        {
          if (s == p->rank)
            {
              p->peer_export[s] = p->export_buf;
              p->peer_flags[s]  = p->flags;
              continue;
            }
          const bool need = p->send_ptr[s + 1] > p->send_ptr[s] || p->recv_ptr[s + 1] > p->recv_ptr[s];
          if (!need)
            continue;
          void *a = nullptr, *b = nullptr;
          PD_CUDA(cudaIpcOpenMemHandle(&a, hd[2 * s], cudaIpcMemLazyEnablePeerAccess));
          PD_CUDA(cudaIpcOpenMemHandle(&b, hd[2 * s + 1], cudaIpcMemLazyEnablePeerAccess));
          p->peer_export[s] = static_cast<double *>(a);
          p->peer_flags[s]  = static_cast<unsigned long long *>(b);
        }
      p->d_peer_export.alloc(p->world);
      p->d_peer_flags.alloc(p->world);
      PD_CUDA(cudaMemcpy(p->d_peer_export.p, p->peer_export.data(), sizeof(double *) * p->world, cudaMemcpyHostToDevice));
      PD_CUDA(cudaMemcpy(p->d_peer_flags.p, p->peer_flags.data(), sizeof(unsigned long long *) * p->world,
                         cudaMemcpyHostToDevice));
      p->connected = true;
    }
  }

  void
  peer_exchange(pd_peer *p, double *x_full_dev)
  {
    {
      if (!p || !x_full_dev)
        throw Error(PD_ERR_INVALID, "pd_peer_exchange: null argument");
      if (!p->connected)
        throw Error(PD_ERR_STATE, "pd_peer_exchange: pd_peer_connect has not been called");
      pd_handle    *h      = p->h;
      const int     n      = h->n;
      const int64_t parity = (++p->epoch) & 1;
      if (p->n_send > 0 || p->n_neighbours > 0)
        {
          const int grid = (int)std::min<int64_t>(std::max<int64_t>(1, (p->n_send * n + 255) / 256), h->sm_count);
          k_peer_publish<<<grid, 256, 0, h->stream>>>(x_full_dev, p->send_blocks.p, p->n_send, n, (int)parity,
                                                      p->export_buf, p->d_peer_flags.p,
                                                      p->d_neighbours.p, p->n_neighbours, p->rank,
                                                      (unsigned long long)p->epoch, p->counter);
          ++h->launches;
        }
      if (p->n_recv > 0)
        {
          const int grid = (int)std::min<int64_t>((p->n_recv * n + 255) / 256, (int64_t)h->sm_count * 4);
          k_peer_pull<<<grid, 256, 0, h->stream>>>(x_full_dev + (int64_t)h->np_own * n, p->recv_owner_of_block.p,
                                                   p->recv_src_block.p, p->n_recv, n, p->d_peer_export.p, (int)parity, p->flags,
                                                   p->world, (unsigned long long)p->epoch, p->d_owners.p, p->n_owners);
          ++h->launches;
        }
      PD_CUDA(cudaGetLastError());
    }
  }

  int
  peer_status(pd_peer *p)
  {
    if (!p)
      return PD_ERR_INVALID;
    unsigned long long e = 0;
    if (cudaMemcpy(&e, p->flags + p->world, sizeof e, cudaMemcpyDeviceToHost) != cudaSuccess)
      return PD_ERR_CUDA;
    return e ? PD_ERR_STATE : PD_OK;
  }

  void
  peer_destroy(pd_peer *p)
  {
    if (!p)
      return;
    for (int s = 0; s < p->world; ++s)
      if (s != p->rank)
        {
          if (s < (int)p->peer_export.size() && p->peer_export[s])
            cudaIpcCloseMemHandle(p->peer_export[s]);
          if (s < (int)p->peer_flags.size() && p->peer_flags[s])
            cudaIpcCloseMemHandle(p->peer_flags[s]);
        }
    cudaFree(p->export_buf);
    cudaFree(p->flags);
    cudaFree(p->counter);
    delete p;
  }
} // namespace pd
