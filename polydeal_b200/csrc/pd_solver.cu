// -----------------------------------------------------------------------------
// pd_solver.cu -- the immediate callers of vmult, device resident.
//
// Reference: SolverCG around the operator (examples/diffusion_reaction.cc:
// 709-724, examples/matrix_free_agglo.cc:377-384) and PreconditionChebyshev
// with a Jacobi (inverse-diagonal) inner preconditioner as multigrid smoother
// (examples/matrix_free_agglo.cc:264-319; inverse diagonal include/utils.h:
// 797-814).  Both are deal.II classes; what is restated here is the textbook
// algorithm they implement:
//   * preconditioned conjugate gradients (Hestenes-Stiefel),
//   * the Chebyshev three-term recurrence on [lambda_max / range, lambda_max]
//     (Saad, Iterative Methods, Alg. 12.1; the form used by deal.II),
//   * a power iteration on D^-1 A for lambda_max (deal.II estimates it with a
//     few CG/Lanczos steps; either way it is an input of the smoother).
// Parity unpinned against deal.II's own iterates (no golden in the reference);
// tests compare with a numpy restatement and with direct solves.
//
// Everything stays on the device: scalars (alpha, beta, residual norms) live in
// device memory, the vector updates are fused around the two dot products, the
// dot products are deterministic two-stage reductions, and the body of a CG
// iteration is captured once in a CUDA graph and replayed (the loop is
// launch-bound at config-B sizes: one SpMV + 3 small kernels; the dot products are finalised by the
// last CTA of the kernel that produces them, so there is no separate reduction launch).
// -----------------------------------------------------------------------------
#include "pd_host.hpp"
#include "pd_internal.hpp"

#include <algorithm>
#include <cmath>
#include <vector>

namespace pd
{
  void vmult_dispatch(pd_handle *h, int mode, const double *src, double *dst, bool add); // pd_api.cu

  namespace
  {
    constexpr int RB = 256;  // reduction block
    constexpr int RG = 296;  // reduction grid (2 per SM)

    // partial[k][block] = sum over this block's grid-stride share of a_k[i] * b_k[i], k < NK; the LAST CTA
    // to arrive (ticket) then sums the RG partials of every k in a fixed order -- lane l adds b = l, l + 32,
    // ..., fixed butterfly -- into scal[dst0 + k]: deterministic, and no separate finalize launch.
    template <int NK>
    __device__ __forceinline__ void
    block_reduce_store(double (&s)[NK], double *partial, double *scal, const int dst0, unsigned int *ticket)
    {
      __shared__ double sh[NK][RB / 32];
      __shared__ bool   last;
#pragma unroll
      for (int k = 0; k < NK; ++k)
        {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
          if ((threadIdx.x & 31) == 0)
            sh[k][threadIdx.x >> 5] = s[k];
        }
      __syncthreads();
      if (threadIdx.x < NK)
        {
          double t = 0.;
#pragma unroll
          for (int w = 0; w < RB / 32; ++w)
            t += sh[threadIdx.x][w];
          partial[threadIdx.x * RG + blockIdx.x] = t;
          __threadfence();
        }
      __syncthreads();
      if (threadIdx.x == 0)
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
      __syncthreads();
      if (last && threadIdx.x < 32)
        {
          __threadfence();
          for (int k = 0; k < NK; ++k)
            {
              double t = 0.;
              for (int b = threadIdx.x; b < RG; b += 32)
                t += __ldcg(partial + k * RG + b);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
                t += __shfl_xor_sync(0xffffffffu, t, o);
              if (threadIdx.x == 0)
                scal[dst0 + k] = t;
            }
          if (threadIdx.x == 0)
            *ticket = 0;
        }
    }

    // scalars: [0] rz  [1] pAp  [2] rz_new  [3] rr  [4] bb
    __global__ void __launch_bounds__(RB)
    k_dot_pAp(const double *__restrict__ p, const double *__restrict__ Ap, const int64_t n, double *partial, double *scal,
              unsigned int *ticket)
    {
      double s[1] = {0.};
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)RG * RB)
        s[0] += p[i] * Ap[i];
      block_reduce_store<1>(s, partial, scal, 1, ticket); // scal[1] = pAp
    }

    // x += alpha p ; r -= alpha Ap ; z = dinv r (or r) ; partial <- (r.z, r.r);  alpha = rz / pAp
    __global__ void __launch_bounds__(RB)
    k_cg_update(double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, const double *__restrict__ p,
                const double *__restrict__ Ap, const double *__restrict__ dinv, double *scal, const int64_t n,
                double *partial, unsigned int *ticket)
    {
      const double alpha = scal[1] != 0. ? scal[0] / scal[1] : 0.; // converged exactly: stay put
      double       s[2]  = {0., 0.};
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)RG * RB)
        {
          x[i] += alpha * p[i];
          const double ri = r[i] - alpha * Ap[i];
          r[i]            = ri;
          const double zi = dinv ? dinv[i] * ri : ri;
          z[i]            = zi;
          s[0] += ri * zi;
          s[1] += ri * ri;
        }
      block_reduce_store<2>(s, partial, scal, 2, ticket); // scal[2] = rz_new, scal[3] = rr
    }

    // p = z + beta p, beta = rz_new / rz; the last CTA to finish rolls rz <- rz_new (every CTA has read
    // both before it takes its ticket)
    __global__ void __launch_bounds__(RB)
    k_cg_direction(double *__restrict__ p, const double *__restrict__ z, double *scal, const int64_t n, unsigned int *ticket)
    {
      const double rz = scal[0], rz_new = scal[2];
      const double beta = rz != 0. ? rz_new / rz : 0.;
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)RG * RB)
        p[i] = z[i] + beta * p[i];
      __syncthreads();
      if (threadIdx.x == 0 && atomicAdd(ticket, 1u) == gridDim.x - 1)
        {
          scal[0] = rz_new;
          *ticket = 0;
        }
    }
    __global__ void
    k_cg_roll(double *scal)
    {
      scal[0] = scal[2];
    }

    // r = b - Ax (Ax given) ; z = dinv r ; p = z ; partial <- (r.z, r.r, b.b)
    __global__ void __launch_bounds__(RB)
    k_cg_init(const double *__restrict__ b, const double *__restrict__ Ax, double *__restrict__ r, double *__restrict__ z,
              double *__restrict__ p, const double *__restrict__ dinv, const int64_t n, double *partial, double *scal,
              unsigned int *ticket)
    {
      double s[3] = {0., 0., 0.};
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)RG * RB)
        {
          const double ri = b[i] - Ax[i];
          r[i]            = ri;
          const double zi = dinv ? dinv[i] * ri : ri;
          z[i]            = zi;
          p[i]            = zi;
          s[0] += ri * zi;
          s[1] += ri * ri;
          s[2] += b[i] * b[i];
        }
      block_reduce_store<3>(s, partial, scal, 2, ticket); // scal[2] = rz, [3] = rr, [4] = bb
    }

    // Chebyshev step: d = c1 d + c2 dinv (b - Ax) ; x += d     (Ax == nullptr: zero start, Ax = 0)
    __global__ void __launch_bounds__(RB)
    k_cheb_step(double *__restrict__ x, double *__restrict__ d, const double *__restrict__ b, const double *__restrict__ Ax,
                const double *__restrict__ dinv, const double c1, const double c2, const int64_t n, const int first)
    {
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB)
        {
          const double res = b[i] - (Ax ? Ax[i] : 0.);
          const double di  = (first ? 0. : c1 * d[i]) + c2 * dinv[i] * res;
          d[i]             = di;
          x[i]             = (first && !Ax ? 0. : x[i]) + di;
        }
    }

    // power iteration on D^-1 A:  w = dinv (A v) ; partial <- (w.w, v.w)
    __global__ void __launch_bounds__(RB)
    k_power(const double *__restrict__ Av, const double *__restrict__ dinv, const double *__restrict__ v,
            double *__restrict__ w, const int64_t n, double *partial, double *scal, unsigned int *ticket)
    {
      double s[2] = {0., 0.};
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)RG * RB)
        {
          const double wi = dinv[i] * Av[i];
          w[i]            = wi;
          s[0] += wi * wi;
          s[1] += v[i] * wi;
        }
      block_reduce_store<2>(s, partial, scal, 0, ticket);
    }
    __global__ void __launch_bounds__(RB)
    k_scale_from(double *__restrict__ v, const double *__restrict__ w, const double *__restrict__ scal, const int64_t n)
    {
      const double inv = rsqrt(scal[0]);
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB)
        v[i] = w[i] * inv;
    }
    __global__ void __launch_bounds__(RB)
    k_fill_guess(double *__restrict__ v, const int64_t n)
    {
      // deterministic, non-constant start vector with zero mean tendency
      for (int64_t i = (int64_t)blockIdx.x * RB + threadIdx.x; i < n; i += (int64_t)gridDim.x * RB)
        v[i] = (double)((i % 11) - 5) + 0.5;
    }

    // x = sum over the blocks of one colour of the unit vector of local DoF i
    __global__ void __launch_bounds__(RB)
    k_unit_of_colour(double *__restrict__ x, const int32_t *__restrict__ colour, const int c, const int i, const int n,
                     const int64_t n_own, const int64_t n_src)
    {
      for (int64_t k = (int64_t)blockIdx.x * RB + threadIdx.x; k < n_src; k += (int64_t)gridDim.x * RB)
        {
          const int64_t b = k / n;
          x[k]            = (k < n_own && colour[b] == c && k - b * n == i) ? 1. : 0.;
        }
    }
    // dinv[b n + i] = 1 / y[b n + i] for the blocks of that colour (entries below 1e-10 kept,
    // include/utils.h:797-814)
    __global__ void __launch_bounds__(RB)
    k_take_diagonal(const double *__restrict__ y, const int32_t *__restrict__ colour, const int c, const int i, const int n,
                    const int64_t n_blocks, double *__restrict__ dinv)
    {
      for (int64_t b = (int64_t)blockIdx.x * RB + threadIdx.x; b < n_blocks; b += (int64_t)gridDim.x * RB)
        if (colour[b] == c)
          {
            const double d  = y[b * n + i];
            dinv[b * n + i] = fabs(d) > 1e-10 ? 1. / d : d;
          }
    }

    void
    ensure_work(pd_handle *h)
    {
      const size_t n = (size_t)h->n_dofs, ns = (size_t)h->np * h->n;
      if (h->sv_r.n != n)
        {
          h->sv_r.alloc(n);
          h->sv_z.alloc(n);
          h->sv_Ap.alloc(n);
          h->sv_dinv.alloc(n);
          h->sv_p.alloc(ns); // search direction is a vmult source: owned + ghost length
          h->sv_partial.alloc(4 * RG);
          h->sv_scal.alloc(8);
          h->sv_ticket.alloc(1);
          ++h->op_generation; // captured graphs point into the old buffers
          PD_CUDA(cudaMemsetAsync(h->sv_ticket.p, 0, sizeof(unsigned int), h->stream));
          PD_CUDA(cudaMemsetAsync(h->sv_p.p, 0, sizeof(double) * ns, h->stream));
        }
    }
  } // namespace

  // With `peer` (a sharded handle): the search direction's ghost blocks are pulled over NVLink
  // before every apply and every dot product is summed over the ranks by the peer-memory
  // all-reduce (pd_peer.cu) -- all inside the replayed graph, no NCCL and no host in the loop.
  // The reduced scalars are bitwise identical on all ranks, so every rank takes the same path.
  // get_matrix_diagonal_inverse() of whatever operator `mode` applies (include/utils.h:797-814, 929-1100).
  // BLOCK_CSR: read off the assembled matrix.  Matrix-free modes: the operator is applied to sums of unit
  // vectors over an independent set of polytopes (greedy colouring of the block pattern, host, once), so
  // every diagonal entry is isolated exactly: n * n_colours applies, cached per (mode, flags, coefficients).
  void
  solver_diagonal_inverse(pd_handle *h, const int mode, double *dst)
  {
    if (mode == PD_VMULT_BLOCK_CSR || h->np != h->np_own)
      {
        // sharded handles: the assembled diagonal (equal to the matrix-free operator's on Cartesian cells)
        if (!h->assembled)
          throw Error(PD_ERR_STATE, "the inverse diagonal needs pd_assemble (block-CSR mode or a sharded handle)");
        launch_diagonal_inverse(h, dst);
        return;
      }
    ensure_work(h);
    const int64_t n_own = h->n_dofs, n_src = (int64_t)h->np * h->n;
    cudaStream_t  s = h->stream;
    if (!(h->mfd_valid && h->mfd_mode == mode && h->mfd_flags == h->op_flags && h->mfd_coef[0] == h->op_coef.stiffness &&
          h->mfd_coef[1] == h->op_coef.mass))
      {
        if (h->mfd_colour.n != (size_t)h->np_own)
          {
            std::vector<int32_t> colour((size_t)h->np_own, -1);
            int                  n_colours = 0;
            std::vector<char>    used;
            for (int32_t b = 0; b < h->np_own; ++b)
              {
                used.assign((size_t)n_colours + 1, 0);
                for (int64_t e = h->h_brow_ptr[b]; e < h->h_brow_ptr[b + 1]; ++e)
                  {
                    const int32_t nb = h->h_bcol[e];
                    if (nb != b && nb < h->np_own && colour[nb] >= 0)
                      used[colour[nb]] = 1;
                  }
                int c = 0;
                while (used[c])
                  ++c;
                colour[b] = c;
                n_colours = std::max(n_colours, c + 1);
              }
            h->mfd_colour.alloc(colour.size());
            PD_CUDA(cudaMemcpy(h->mfd_colour.p, colour.data(), colour.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            h->mfd_n_colours = n_colours;
            h->mfd_dinv.alloc((size_t)n_own);
          }
        double   *x = h->sv_p.p, *y = h->sv_Ap.p;
        const int grid = (int)std::min<int64_t>((n_src + RB - 1) / RB, 148 * 8);
        for (int c = 0; c < h->mfd_n_colours; ++c)
          for (int i = 0; i < h->n; ++i)
            {
              k_unit_of_colour<<<grid, RB, 0, s>>>(x, h->mfd_colour.p, c, i, h->n, n_own, n_src);
              vmult_dispatch(h, mode, x, y, false);
              k_take_diagonal<<<grid, RB, 0, s>>>(y, h->mfd_colour.p, c, i, h->n, h->np_own, h->mfd_dinv.p);
              h->launches += 2;
            }
        PD_CUDA(cudaGetLastError());
        h->mfd_valid   = true;
        h->mfd_mode    = mode;
        h->mfd_flags   = h->op_flags;
        h->mfd_coef[0] = h->op_coef.stiffness;
        h->mfd_coef[1] = h->op_coef.mass;
      }
    PD_CUDA(cudaMemcpyAsync(dst, h->mfd_dinv.p, sizeof(double) * n_own, cudaMemcpyDeviceToDevice, s));
  }

  // returns false when the tolerance was not reached within max_iter iterations (SolverCG throws
  // SolverControl::NoConvergence there; the C ABI reports PD_NOT_CONVERGED with the iterate kept)
  bool
  solver_cg(pd_handle *h, const int mode, const double *b, double *x, const int max_iter, const double rel_tol,
            const int jacobi, int *iters_out, double *relres_out, pd_peer *peer)
  {
    if (h->np != h->np_own && !peer)
      throw Error(PD_ERR_UNSUPPORTED, "pd_cg_solve: a handle with ghost polytopes needs pd_cg_solve_sharded");
    if (peer && peer_handle(peer) != h)
      throw Error(PD_ERR_INVALID, "pd_cg_solve_sharded: the peer object belongs to another handle");
    // stream capture is not available on the legacy / per-thread default streams: run the solve on the handle's
    // own stream, ordered behind the caller's work by an event (the solve ends with a synchronisation of its
    // stream, so the caller's stream needs no dependency in the other direction)
    struct OwnStreamScope
    {
      pd_handle   *h;
      cudaStream_t saved;
      explicit OwnStreamScope(pd_handle *h_)
        : h(h_)
        , saved(h_->stream)
      {
        if (saved == cudaStreamLegacy || saved == cudaStreamPerThread || saved == nullptr)
          {
            if (!h->ev_order)
              PD_CUDA(cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming));
            PD_CUDA(cudaEventRecord(h->ev_order, saved));
            PD_CUDA(cudaStreamWaitEvent(h->own_stream, h->ev_order, 0));
            h->stream = h->own_stream;
          }
      }
      ~OwnStreamScope()
      {
        h->stream = saved;
      }
    } own_stream_scope(h);
    ensure_work(h);
    auto apply = [&](double *src_full, double *dst) {
      if (peer)
        peer_vmult(peer, mode, src_full, dst, false);
      else
        vmult_dispatch(h, mode, src_full, dst, false);
    };
    auto reduce = [&](const int dst0, const int nk) {
      if (peer)
        peer_allreduce(peer, h->sv_scal.p, dst0, nk);
    };
    const int64_t n = h->n_dofs;
    cudaStream_t  s = h->stream;
    double       *r = h->sv_r.p, *z = h->sv_z.p, *p = h->sv_p.p, *Ap = h->sv_Ap.p, *partial = h->sv_partial.p,
           *scal = h->sv_scal.p;
    const double *dinv = nullptr;
    if (jacobi)
      {
        solver_diagonal_inverse(h, mode, h->sv_dinv.p);
        dinv = h->sv_dinv.p;
      }
    if (peer)
      {
        // x holds the owned DoFs only: stage it in the (owned + ghost) direction buffer
        PD_CUDA(cudaMemcpyAsync(p, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        apply(p, Ap);
      }
    else
      vmult_dispatch(h, mode, x, Ap, false);
    unsigned int *ticket = h->sv_ticket.p;
    k_cg_init<<<RG, RB, 0, s>>>(b, Ap, r, z, p, dinv, n, partial, scal, ticket);
    reduce(2, 3);
    k_cg_roll<<<1, 1, 0, s>>>(scal);
    h->launches += 2;
    double hs[5];
    PD_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, s));
    PD_CUDA(cudaStreamSynchronize(s));
    const double bnorm = std::sqrt(hs[4] > 0 ? hs[4] : 1.);
    double       relres = std::sqrt(hs[3]) / bnorm;
    int          it     = 0;
    // one iteration = SpMV + dot, update, direction (3 kernels); captured once and replayed: a graph of CHUNK
    // iterations and a graph of one iteration for the tail, so that exactly min(max_iter, needed rounded up to
    // the check interval) iterations run.  The graphs bake in the operator (terms, coefficients, kernel choice
    // are kernel parameters / template choices of the captured launches), the stream, the work buffers and
    // the x / b pointers: all of that is in the cache key (op_generation is bumped by pd_set_operator,
    // pd_force_generic_matrix_free, pd_upload, pd_set_stream and every reallocation of the work buffers).
    constexpr int CHUNK = 8;
    const bool    stale = !h->cg_graph_exec || h->cg_graph_mode != mode || h->cg_graph_jacobi != jacobi ||
                       h->cg_graph_peer != peer || h->cg_graph_x != x || h->cg_graph_b != b ||
                       h->cg_graph_generation != h->op_generation;
    if (stale)
      {
        for (cudaGraphExec_t *g : {&h->cg_graph_exec, &h->cg_graph1_exec})
          if (*g)
            {
              cudaGraphExecDestroy(*g);
              *g = nullptr;
            }
        auto capture = [&](const int n_iter, cudaGraphExec_t *exec) {
          cudaGraph_t   graph = nullptr;
          const int64_t l0    = h->launches;
          PD_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
          try
            {
              for (int k = 0; k < n_iter; ++k)
                {
                  apply(p, Ap);
                  k_dot_pAp<<<RG, RB, 0, s>>>(p, Ap, n, partial, scal, ticket);
                  reduce(1, 1);
                  k_cg_update<<<RG, RB, 0, s>>>(x, r, z, p, Ap, dinv, scal, n, partial, ticket);
                  reduce(2, 2);
                  k_cg_direction<<<RG, RB, 0, s>>>(p, z, scal, n, ticket);
                }
            }
          catch (...)
            {
              cudaStreamEndCapture(s, &graph); // never leave the stream in capture mode
              if (graph)
                cudaGraphDestroy(graph);
              h->launches = l0;
              throw;
            }
          const int64_t per_iter = (h->launches - l0) / n_iter + 3;
          h->launches            = l0;
          PD_CUDA(cudaStreamEndCapture(s, &graph));
          const cudaError_t e = cudaGraphInstantiate(exec, graph, 0);
          cudaGraphDestroy(graph);
          PD_CUDA(e);
          return per_iter;
        };
        const int64_t per_iter   = capture(CHUNK, &h->cg_graph_exec);
        h->cg_launches_per_chunk = per_iter * CHUNK;
        capture(1, &h->cg_graph1_exec);
        h->cg_graph_mode       = mode;
        h->cg_graph_jacobi     = jacobi;
        h->cg_graph_peer       = peer;
        h->cg_graph_x          = x;
        h->cg_graph_b          = b;
        h->cg_graph_generation = h->op_generation;
      }
    // rel_tol <= 0: no convergence test, exactly max_iter iterations (timing runs)
    while (it < max_iter && (rel_tol <= 0. || relres > rel_tol))
      {
        const bool chunk = max_iter - it >= CHUNK;
        PD_CUDA(cudaGraphLaunch(chunk ? h->cg_graph_exec : h->cg_graph1_exec, s));
        h->launches += chunk ? h->cg_launches_per_chunk : h->cg_launches_per_chunk / CHUNK;
        it += chunk ? CHUNK : 1;
        if (rel_tol <= 0. && it < max_iter)
          continue;
        PD_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, s));
        PD_CUDA(cudaStreamSynchronize(s));
        relres = std::sqrt(hs[3]) / bnorm;
        if (!(relres == relres))
          throw Error(PD_ERR_CUDA, "pd_cg_solve: residual became NaN (operator not positive definite?)");
      }
    if (iters_out)
      *iters_out = it;
    if (relres_out)
      *relres_out = relres;
    return rel_tol <= 0. || relres <= rel_tol;
  }

  double
  solver_lambda_max(pd_handle *h, const int mode, const int n_iter, pd_peer *peer)
  {
    if (h->np != h->np_own && !peer)
      throw Error(PD_ERR_UNSUPPORTED, "pd_estimate_lambda_max: a handle with ghost polytopes needs the sharded call");
    if (peer && peer_handle(peer) != h)
      throw Error(PD_ERR_INVALID, "pd_estimate_lambda_max_sharded: the peer object belongs to another handle");
    ensure_work(h);
    const int64_t n = h->n_dofs;
    cudaStream_t  s = h->stream;
    double       *v = h->sv_p.p, *w = h->sv_z.p, *Av = h->sv_Ap.p, *partial = h->sv_partial.p, *scal = h->sv_scal.p;
    solver_diagonal_inverse(h, mode, h->sv_dinv.p);
    k_fill_guess<<<RG, RB, 0, s>>>(v, n);
    double lambda = 0.;
    for (int it = 0; it < n_iter; ++it)
      {
        if (peer)
          peer_exchange(peer, v);
        vmult_dispatch(h, mode, v, Av, false);
        k_power<<<RG, RB, 0, s>>>(Av, h->sv_dinv.p, v, w, n, partial, scal, h->sv_ticket.p);
        if (peer)
          peer_allreduce(peer, scal, 0, 2);
        k_scale_from<<<RG, RB, 0, s>>>(v, w, scal, n);
        h->launches += 2;
      }
    double hs[2];
    PD_CUDA(cudaMemcpyAsync(hs, scal, sizeof(hs), cudaMemcpyDeviceToHost, s));
    PD_CUDA(cudaStreamSynchronize(s));
    // last iterate: v was normalised before the product, so |D^-1 A v| estimates lambda_max
    lambda = std::sqrt(hs[0]);
    return lambda;
  }

  // sharded (peer != null): x is a vmult source, i.e. it has the (owned + ghost) length
  // pd_n_source_dofs; its ghost section is refreshed over peer memory before every apply.
  void
  solver_chebyshev(pd_handle *h, const int mode, const int degree, const double lambda_max, const double smoothing_range,
                   const double *b, double *x, const int zero_initial_guess, pd_peer *peer)
  {
    if (h->np != h->np_own && !peer)
      throw Error(PD_ERR_UNSUPPORTED, "pd_chebyshev_smooth: a handle with ghost polytopes needs the sharded call");
    if (peer && peer_handle(peer) != h)
      throw Error(PD_ERR_INVALID, "pd_chebyshev_smooth_sharded: the peer object belongs to another handle");
    if (degree < 1 || !(lambda_max > 0.) || !(smoothing_range > 1.))
      throw Error(PD_ERR_INVALID, "pd_chebyshev_smooth: need degree >= 1, lambda_max > 0, smoothing_range > 1");
    ensure_work(h);
    const int64_t n = h->n_dofs;
    cudaStream_t  s = h->stream;
    double       *d = h->sv_z.p, *Ax = h->sv_Ap.p;
    solver_diagonal_inverse(h, mode, h->sv_dinv.p);
    const double lmin = lambda_max / smoothing_range;
    const double theta = 0.5 * (lambda_max + lmin), delta = 0.5 * (lambda_max - lmin);
    const double sigma1 = theta / delta;
    double       rho    = 1. / sigma1;
    const int    grid   = (int)std::min<int64_t>((n + RB - 1) / RB, 148 * 8);
    // step 0: d = 1/theta D^-1 (b - A x0), x = x0 + d
    auto apply = [&] {
      if (peer)
        peer_exchange(peer, x);
      vmult_dispatch(h, mode, x, Ax, false);
    };
    if (!zero_initial_guess)
      apply();
    k_cheb_step<<<grid, RB, 0, s>>>(x, d, b, zero_initial_guess ? nullptr : Ax, h->sv_dinv.p, 0., 1. / theta, n, 1);
    ++h->launches;
    for (int k = 1; k < degree; ++k)
      {
        const double rho_new = 1. / (2. * sigma1 - rho);
        apply();
        k_cheb_step<<<grid, RB, 0, s>>>(x, d, b, Ax, h->sv_dinv.p, rho_new * rho, 2. * rho_new / delta, n, 0);
        ++h->launches;
        rho = rho_new;
      }
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
