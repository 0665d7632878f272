// -----------------------------------------------------------------------------
// pd_peer.cu -- the collectives of the sharded operator over NVLink peer memory:
// the ghost exchange of vmult and the all-reduce of the solver's dot products.
//
// Reference: LinearAlgebra::distributed::Vector::update_ghost_values() inside
// MatrixFree::loop (include/utils.h:466-472) / the Trilinos import of the matrix-based
// vmult -- the coefficient blocks of the ghost polytopes travel before every apply -- and the
// MPI_Allreduce inside every dot product of SolverCG on distributed vectors.
//
// With one process per GPU on an NVSwitch node there is no need for a message.  Every rank
// owns ONE buffer allocated with cudaMalloc that all the others map through CUDA IPC:
//   [ flags[world] | error word | red_flags[world] | red_vals[world][2][4] | export[n_send][2][n] ]
// Ghost exchange, two small kernels on the operator's stream:
//   publish  e = epoch + 1; pack my send blocks into export[.][e & 1]; the last CTA to finish
//            fences (system scope) and stores e into flags[my rank] ON every neighbour
//   pull     every CTA spins (bounded) until flags[owner] >= e for the ranks I receive from,
//            then the grid copies their segments into my ghost section with loads over NVLink
// All-reduce of up to 4 scalars, one warp: store my values into red_vals[my rank][e & 1] ON every
// rank, fence, store e into red_flags[my rank] there; spin until every rank's flag reached e;
// sum in rank order (every rank gets bitwise the same result).
// Double buffering removes the write-after-read hazard: a rank cannot write epoch e + 2 into
// the slot a peer still reads for epoch e, because its own wait of epoch e + 1 needed that
// peer's store of e + 1, which follows the peer's read of e in stream order.
// The epochs live in device memory, so the kernels carry no host state and a whole solver
// iteration (exchange + SpMV + dot products + all-reduce) replays from a CUDA graph.
// A spin that does not see its flag within ~2 s raises the error word (pd_peer_status) and
// moves on instead of hanging the GPU.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"
#include "pd_host.hpp"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace
{
  using u64 = unsigned long long;
  constexpr int RED_MAX = 4; // scalars per all-reduce

  // byte offsets of the fixed-size regions inside every rank's IPC buffer
  __host__ __device__ inline size_t
  off_flags(int)
  {
    return 0;
  }
  __host__ __device__ inline size_t
  off_red_flags(const int world)
  {
    return (size_t)(world + 1) * sizeof(u64);
  }
  __host__ __device__ inline size_t
  off_red_vals(const int world)
  {
    return off_red_flags(world) + (size_t)world * sizeof(u64);
  }
  __host__ __device__ inline size_t
  off_export(const int world)
  {
    return (off_red_vals(world) + (size_t)world * 2 * RED_MAX * sizeof(double) + 15) / 16 * 16;
  }
  // The export region: send block b, epoch parity pe, at double (2 b + pe) slot + shift(b) with slot = n rounded up
  // to even and shift(b) = b & 1 for odd n.  Both epoch copies of a block then have the SAME 16-byte alignment, and
  // consecutive blocks alternate -- what the fused fine-mesh apply needs to read ghost cells from here with 16-byte
  // cp.async chunks into rows of matching alignment (pd_fine_cell.hpp: StreamPlan).
  __host__ __device__ inline int64_t
  export_slot(const int n)
  {
    return n + (n & 1);
  }
  __host__ __device__ inline int64_t
  export_at(const int64_t b, const int parity, const int n)
  {
    return (2 * b + parity) * export_slot(n) + ((n & 1) ? (b & 1) : 0);
  }
} // namespace

struct pd_peer
{
  pd_handle *h    = nullptr;
  int        rank = 0, world = 1;
  // plan (in polytope blocks)
  std::vector<int64_t> send_ptr, recv_ptr;
  pd::DevBuf<int32_t>  send_blocks, recv_owner_of_block, d_neighbours, d_owners;
  pd::DevBuf<int64_t>  recv_src_block; // position of every ghost block inside its owner's send list
  std::vector<int32_t> h_recv_owner;
  std::vector<int64_t> h_recv_src;
  bool                 fused = false; // the fused fine-mesh apply is set up (pd_finemesh.cu: setup_fine_fused)
  int64_t              n_send = 0, n_recv = 0;
  int                  n_neighbours = 0, n_owners = 0;
  // this rank's IPC buffer and device-side state
  char         *ipc     = nullptr;
  u64          *epochs  = nullptr; // [0] exchange epoch, [1] all-reduce epoch
  unsigned int *counter = nullptr; // last-CTA ticket of publish
  // the peers' buffers as mapped here ([rank] = own)
  std::vector<char *> peer_base;
  pd::DevBuf<char *>  d_peer_base;
  bool                connected = false;
  // a second stream for the exchange, so that work which needs no ghost data overlaps it
  cudaStream_t aux     = nullptr;
  cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
};

namespace pd
{
  namespace
  {
    constexpr long long SPIN_CYCLES = 4000000000ll; // ~2 s

    __device__ __forceinline__ void
    wait_for(volatile u64 *flag, const u64 epoch, volatile u64 *error_word)
    {
      const long long t0 = clock64();
      while (*flag < epoch)
        {
          if (clock64() - t0 > SPIN_CYCLES) // the peer is gone: report, do not hang
            {
              *error_word = 1ull;
              break;
            }
          __nanosleep(64);
        }
    }

    __global__ void
    k_peer_publish(const double *x, const int32_t *send_blocks, const int64_t n_send, const int n, char *const *peer_base,
                   const int32_t *neighbours, const int n_neighbours, const int rank, const int world, u64 *epochs,
                   unsigned int *counter)
    {
      const u64     e      = *(volatile u64 *)epochs + 1;
      const int     parity = (int)(e & 1);
      double       *exp    = reinterpret_cast<double *>(peer_base[rank] + off_export(world));
      const int64_t total  = n_send * n;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t b                         = i / n;
          exp[export_at(b, parity, n) + (i - b * n)] = x[(int64_t)send_blocks[b] * n + (i - b * n)];
        }
      // the last CTA to finish announces the epoch to every neighbour (all CTAs have read `epochs`
      // before taking their ticket, so bumping it here is safe)
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
              volatile u64 *f = reinterpret_cast<u64 *>(peer_base[neighbours[k]] + off_flags(world)) + rank;
              *f              = e;
            }
          if (threadIdx.x == 0)
            {
              *counter  = 0;
              epochs[0] = e;
            }
        }
    }

    __global__ void
    k_peer_pull(double *ghost, const int32_t *owner_of_block, const int64_t *src_block, const int64_t n_recv, const int n,
                char *const *peer_base, const int rank, const int world, const u64 *epochs, const int32_t *owners,
                const int n_owners)
    {
      const u64     e      = *(volatile const u64 *)epochs; // bumped by the publish in front of me
      const int     parity = (int)(e & 1);
      volatile u64 *flags  = reinterpret_cast<u64 *>(peer_base[rank] + off_flags(world));
      if ((int)threadIdx.x < n_owners)
        {
          wait_for(flags + owners[threadIdx.x], e, flags + world);
          __threadfence_system();
        }
      __syncthreads();
      const int64_t total = n_recv * n;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        {
          const int64_t b   = i / n;
          const double *exp = reinterpret_cast<const double *>(peer_base[owner_of_block[b]] + off_export(world));
          ghost[i]          = __ldcg(exp + export_at(src_block[b], parity, n) + (i - b * n));
        }
    }

    // scal[dst0 .. dst0+nk) <- sum over ranks, one warp
    __global__ void
    k_peer_allreduce(double *scal, const int dst0, const int nk, char *const *peer_base, const int rank, const int world,
                     u64 *epochs)
    {
      const int lane   = threadIdx.x;
      const u64 e      = *(volatile u64 *)(epochs + 1) + 1;
      const int parity = (int)(e & 1);
      double    v[RED_MAX];
#pragma unroll
      for (int k = 0; k < RED_MAX; ++k)
        v[k] = k < nk ? scal[dst0 + k] : 0.;
      for (int r = lane; r < world; r += 32)
        {
          volatile double *dst =
            reinterpret_cast<double *>(peer_base[r] + off_red_vals(world)) + ((size_t)rank * 2 + parity) * RED_MAX;
#pragma unroll
          for (int k = 0; k < RED_MAX; ++k)
            dst[k] = v[k];
          __threadfence_system();
          volatile u64 *f = reinterpret_cast<u64 *>(peer_base[r] + off_red_flags(world)) + rank;
          *f              = e;
        }
      __syncwarp();
      volatile u64 *my_flags = reinterpret_cast<u64 *>(peer_base[rank] + off_red_flags(world));
      volatile u64 *err      = reinterpret_cast<u64 *>(peer_base[rank] + off_flags(world)) + world;
      for (int s = lane; s < world; s += 32)
        wait_for(my_flags + s, e, err);
      __threadfence_system();
      __syncwarp();
      if (lane < nk)
        {
          const volatile double *vals = reinterpret_cast<const double *>(peer_base[rank] + off_red_vals(world));
          double                 t    = 0.;
          for (int s = 0; s < world; ++s) // rank order: the same bits on every rank
            t += vals[((size_t)s * 2 + parity) * RED_MAX + lane];
          scal[dst0 + lane] = t;
        }
      __syncwarp();
      if (lane == 0)
        epochs[1] = e;
    }
  } // namespace

  pd_peer *
  peer_create(pd_handle *h, int rank, int world, const int64_t *send_ptr, const int32_t *send_blocks, const int64_t *recv_ptr,
              const int64_t *remote_offset)
  {
    if (!h || !send_ptr || !recv_ptr || !remote_offset || rank < 0 || rank >= world || world > 32)
      throw Error(PD_ERR_INVALID, "pd_peer_create: bad argument (1 <= world <= 32)");
    std::unique_ptr<pd_peer> p(new pd_peer);
    p->h     = h;
    p->rank  = rank;
    p->world = world;
    p->send_ptr.assign(send_ptr, send_ptr + world + 1);
    p->recv_ptr.assign(recv_ptr, recv_ptr + world + 1);
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
    p->h_recv_owner = owner, p->h_recv_src = src;
    put(p->send_blocks, sb);
    put(p->recv_owner_of_block, owner);
    put(p->recv_src_block, src);
    put(p->d_neighbours, nbrs);
    put(p->d_owners, owners);
    p->n_neighbours = (int)nbrs.size();
    p->n_owners     = (int)owners.size();
    // the IPC-exportable buffer (plain cudaMalloc) and the local device state
    const size_t bytes = off_export(world) + (size_t)std::max<int64_t>(2, 2 * p->n_send * export_slot(h->n) + 2) * sizeof(double);
    PD_CUDA(cudaMalloc((void **)&p->ipc, bytes));
    PD_CUDA(cudaMemset(p->ipc, 0, bytes));
    PD_CUDA(cudaMalloc((void **)&p->epochs, 2 * sizeof(u64)));
    PD_CUDA(cudaMemset(p->epochs, 0, 2 * sizeof(u64)));
    PD_CUDA(cudaMalloc((void **)&p->counter, sizeof(unsigned int)));
    PD_CUDA(cudaMemset(p->counter, 0, sizeof(unsigned int)));
    {
      // highest priority: the small exchange kernels must get SM slots next to a resident grid-stride kernel
      int lo = 0, hi = 0;
      PD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      PD_CUDA(cudaStreamCreateWithPriority(&p->aux, cudaStreamNonBlocking, hi));
    }
    PD_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    PD_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    return p.release();
  }

  int
  peer_handle_bytes()
  {
    return (int)sizeof(cudaIpcMemHandle_t);
  }

  void
  peer_export(pd_peer *p, void *handles_out)
  {
    if (!p || !handles_out)
      throw Error(PD_ERR_INVALID, "pd_peer_export: null argument");
    cudaIpcMemHandle_t hd;
    PD_CUDA(cudaIpcGetMemHandle(&hd, p->ipc));
    std::memcpy(handles_out, &hd, sizeof hd);
  }

  void
  peer_connect(pd_peer *p, const void *all_handles)
  {
    if (!p || !all_handles)
      throw Error(PD_ERR_INVALID, "pd_peer_connect: null argument");
    const auto *hd = static_cast<const cudaIpcMemHandle_t *>(all_handles);
    p->peer_base.assign(p->world, nullptr);
    for (int s = 0; s < p->world; ++s)
      {
        if (s == p->rank)
          {
            p->peer_base[s] = p->ipc;
            continue;
          }
        void *a = nullptr; // every rank is mapped: the all-reduce talks to all of them
        PD_CUDA(cudaIpcOpenMemHandle(&a, hd[s], cudaIpcMemLazyEnablePeerAccess));
        p->peer_base[s] = static_cast<char *>(a);
      }
    p->d_peer_base.alloc(p->world);
    PD_CUDA(cudaMemcpy(p->d_peer_base.p, p->peer_base.data(), sizeof(char *) * p->world, cudaMemcpyHostToDevice));
    p->connected = true;
    // the fused fine-mesh apply reads ghost cells where their owners publish them
    {
      pd_handle                  *h = p->h;
      std::vector<const double *> gs((size_t)p->n_recv);
      for (int64_t g = 0; g < p->n_recv; ++g)
        gs[(size_t)g] = reinterpret_cast<const double *>(p->peer_base[p->h_recv_owner[(size_t)g]] + off_export(p->world)) +
                        export_at(p->h_recv_src[(size_t)g], 0, h->n);
      u64 *flags = reinterpret_cast<u64 *>(p->ipc + off_flags(p->world));
      p->fused   = p->n_recv > 0 && setup_fine_fused(h, gs.data(), export_slot(h->n), p->epochs, flags, flags + p->world,
                                                   p->d_owners.p, p->n_owners);
    }
  }

  static void
  require_connected(const pd_peer *p, const char *who)
  {
    if (!p)
      throw Error(PD_ERR_INVALID, std::string(who) + ": null argument");
    if (!p->connected)
      throw Error(PD_ERR_STATE, std::string(who) + ": pd_peer_connect has not been called");
  }

  void vmult_dispatch(pd_handle *h, int mode, const double *src, double *dst, bool add); // pd_api.cu

  static void
  exchange_on(pd_peer *p, double *x_full_dev, cudaStream_t stream, const bool publish_only = false)
  {
    pd_handle *h = p->h;
    const int  n = h->n;
    // always launched: it is also what advances the epoch the pull reads (a rank that only receives
    // must still step its parity)
    {
        const int grid = (int)std::min<int64_t>(std::max<int64_t>(1, (p->n_send * n + 255) / 256), h->sm_count);
        k_peer_publish<<<grid, 256, 0, stream>>>(x_full_dev, p->send_blocks.p, p->n_send, n, p->d_peer_base.p,
                                                    p->d_neighbours.p, p->n_neighbours, p->rank, p->world, p->epochs,
                                                    p->counter);
        ++h->launches;
      }
    if (p->n_recv > 0 && !publish_only)
      {
        const int grid = (int)std::min<int64_t>((p->n_recv * n + 255) / 256, (int64_t)h->sm_count * 4);
        k_peer_pull<<<grid, 256, 0, stream>>>(x_full_dev + (int64_t)h->np_own * n, p->recv_owner_of_block.p,
                                                 p->recv_src_block.p, p->n_recv, n, p->d_peer_base.p, p->rank, p->world,
                                                 p->epochs, p->d_owners.p, p->n_owners);
        ++h->launches;
      }
    PD_CUDA(cudaGetLastError());
  }

  void
  peer_exchange(pd_peer *p, double *x_full_dev)
  {
    require_connected(p, "pd_peer_exchange");
    if (!x_full_dev)
      throw Error(PD_ERR_INVALID, "pd_peer_exchange: null argument");
    exchange_on(p, x_full_dev, p->h->stream);
  }

  // update_ghost_values() + vmult.  On a fine Cartesian mesh (stencil kernel) the cells whose
  // neighbours are all owned are applied on the operator's stream WHILE the ghost blocks travel
  // on a second stream; the cells next to a cut follow once the pull has finished.
  void
  peer_vmult(pd_peer *p, const int mode, double *x_full_dev, double *dst, const bool add)
  {
    require_connected(p, "pd_peer_vmult");
    if (!x_full_dev || !dst)
      throw Error(PD_ERR_INVALID, "pd_peer_vmult: null argument");
    pd_handle *h = p->h;
    static const bool no_split = getenv("PD_PEER_NO_SPLIT") != nullptr; // A/B switch for measurements
    const bool split = !no_split && mode == PD_VMULT_MATRIX_FREE && h->mf_ready && !h->force_generic_mf && h->fe_kind == PD_FE_DGQ &&
                       h->np != h->np_own && h->mf_list_interior.n > 0 && h->mf_list_boundary.n > 0;
    // block-CSR: worth the extra launch + event only when the apply is long enough to hide the exchange behind
    // (measured on config B, a 16 us apply: 37 us unsplit, 39 us split)
    static const int64_t csr_min = getenv("PD_PEER_CSR_SPLIT_MIN_NNZ") ? atoll(getenv("PD_PEER_CSR_SPLIT_MIN_NNZ")) : (int64_t)16 * 1024 * 1024;
    const bool split_csr = !no_split && mode == PD_VMULT_BLOCK_CSR && h->nnz > csr_min && spmv_can_split(h);
    // the fused path (uniform fine mesh, pipelined kernel): publish, then ONE kernel over all tiles whose boundary
    // tiles wait for the owners' flags and read the ghost cells from the owners' export buffers.  The ghost section
    // of x_full_dev is not written on this path.  PD_PEER_NO_FUSED keeps the split path (measurements, tests).
    static const bool no_fused = getenv("PD_PEER_NO_FUSED") != nullptr;
    if (split && p->fused && !no_fused && reinterpret_cast<uintptr_t>(x_full_dev) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0)
      {
        exchange_on(p, x_full_dev, h->stream, true);
        if (launch_fine_fused(h, x_full_dev, dst, add))
          return;
        // (not launched after all: the ghost section still needs the pull)
        const int grid = (int)std::min<int64_t>((p->n_recv * h->n + 255) / 256, (int64_t)h->sm_count * 4);
        k_peer_pull<<<grid, 256, 0, h->stream>>>(x_full_dev + (int64_t)h->np_own * h->n, p->recv_owner_of_block.p,
                                                    p->recv_src_block.p, p->n_recv, h->n, p->d_peer_base.p, p->rank, p->world,
                                                    p->epochs, p->d_owners.p, p->n_owners);
        ++h->launches;
        vmult_dispatch(h, mode, x_full_dev, dst, add);
        return;
      }
    if (!split && !split_csr)
      {
        exchange_on(p, x_full_dev, h->stream);
        vmult_dispatch(h, mode, x_full_dev, dst, add);
        return;
      }
    // measurements: PD_PEER_DEBUG_PART = 1 runs the interior part only, 2 the exchange + the boundary part only
    const char *dbg = getenv("PD_PEER_DEBUG_PART");
    if (dbg && dbg[0] == '1')
      {
        split ? launch_fine_operator(h, x_full_dev, dst, add, 1) : launch_spmv(h, x_full_dev, dst, add, 1);
        return;
      }
    if (dbg && dbg[0] == '2')
      {
        exchange_on(p, x_full_dev, h->stream);
        split ? launch_fine_operator(h, x_full_dev, dst, add, 2) : launch_spmv(h, x_full_dev, dst, add, 2);
        return;
      }
    // aux (high priority): publish -> pull -> the cells / block rows that read ghost data;
    // main: the cells / block rows that do not -- both run concurrently and meet at the end, so the exchange and
    // the small boundary kernel hide behind the interior work instead of following it
    PD_CUDA(cudaEventRecord(p->ev_fork, h->stream));
    PD_CUDA(cudaStreamWaitEvent(p->aux, p->ev_fork, 0));
    exchange_on(p, x_full_dev, p->aux);
    {
      cudaStream_t main_stream = h->stream;
      h->stream                = p->aux; // the launch helpers enqueue on the handle's stream
      try
        {
          if (split)
            launch_fine_operator(h, x_full_dev, dst, add, 2);
          else
            launch_spmv(h, x_full_dev, dst, add, 2);
        }
      catch (...)
        {
          h->stream = main_stream;
          throw;
        }
      h->stream = main_stream;
    }
    PD_CUDA(cudaEventRecord(p->ev_join, p->aux));
    if (split)
      launch_fine_operator(h, x_full_dev, dst, add, 1);
    else
      launch_spmv(h, x_full_dev, dst, add, 1); // block rows without ghost columns
    PD_CUDA(cudaStreamWaitEvent(h->stream, p->ev_join, 0));
  }

  void
  peer_allreduce(pd_peer *p, double *scal_dev, const int dst0, const int nk)
  {
    require_connected(p, "pd_peer_allreduce");
    if (!scal_dev || nk < 1 || nk > RED_MAX)
      throw Error(PD_ERR_INVALID, "pd_peer_allreduce: 1..4 scalars");
    k_peer_allreduce<<<1, 32, 0, p->h->stream>>>(scal_dev, dst0, nk, p->d_peer_base.p, p->rank, p->world, p->epochs);
    ++p->h->launches;
    PD_CUDA(cudaGetLastError());
  }

  pd_handle *
  peer_handle(pd_peer *p)
  {
    return p ? p->h : nullptr;
  }

  int
  peer_fused(pd_peer *p)
  {
    static const bool no_fused = getenv("PD_PEER_NO_FUSED") != nullptr;
    return (p && p->fused && !no_fused) ? std::max(1, (int)p->h->mf_tiles[3].n_tiles) : 0;
  }

  int
  peer_status(pd_peer *p)
  {
    if (!p)
      return PD_ERR_INVALID;
    u64 e = 0;
    if (cudaMemcpy(&e, p->ipc + off_flags(p->world) + (size_t)p->world * sizeof(u64), sizeof e, cudaMemcpyDeviceToHost) !=
        cudaSuccess)
      return PD_ERR_CUDA;
    return e ? PD_ERR_STATE : PD_OK;
  }

  void
  peer_destroy(pd_peer *p)
  {
    if (!p)
      return;
    for (int s = 0; s < (int)p->peer_base.size(); ++s)
      if (s != p->rank && p->peer_base[s])
        cudaIpcCloseMemHandle(p->peer_base[s]);
    if (p->aux)
      cudaStreamDestroy(p->aux);
    if (p->ev_fork)
      cudaEventDestroy(p->ev_fork);
    if (p->ev_join)
      cudaEventDestroy(p->ev_join);
    cudaFree(p->ipc);
    cudaFree(p->epochs);
    cudaFree(p->counter);
    delete p;
  }
} // namespace pd
