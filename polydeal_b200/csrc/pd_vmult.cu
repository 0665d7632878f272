// -----------------------------------------------------------------------------
// pd_vmult.cu -- operator apply with the assembled matrix.
//
// Reference semantics: the matrix-based vmult used on agglomerated levels,
// LinearOperatorMG::vmult -> TrilinosWrappers::SparseMatrix::vmult
// (include/multigrid_amg.h:345-355, include/linear_operator_for_mg.h:295;
// CG at examples/diffusion_reaction.cc:721-724), and the inverse diagonal of
// include/utils.h:797-814.  The reference's SpMV is scalar CRS and ignores the
// block structure.
//
// HBM-bound: 8 n^2 B per block + 4 B per block index + 16 B per DoF.  The value
// array is the scalar CSR of the reference pattern, so a block row is ONE
// contiguous run of n * (nb*n) doubles.  A CTA takes a block row: it gathers the
// nb source blocks once into shared memory laid out exactly like a matrix row
// (xs[k*n + j] = x[bcol[k]*n + j]), then every warp streams whole rows with
// unit-stride loads and multiplies against xs -- no index arithmetic and no
// gather in the inner loop, two rows in flight per warp for memory-level
// parallelism.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

#include <algorithm>

namespace pd
{
  namespace
  {
    constexpr int SPMV_THREADS = 256;
    constexpr int SPMV_MAX_ROW = 4096; // doubles of shared memory for the gathered source row

    template <bool ADD>
    __global__ void __launch_bounds__(SPMV_THREADS)
    k_spmv_block_row(const double *__restrict__ vals,
                     const int64_t *__restrict__ brow_ptr,
                     const int32_t *__restrict__ bcol,
                     const int      n,
                     const int32_t  n_block_rows,
                     const int32_t *__restrict__ row_list, // optional: the block rows to process
                     const double *__restrict__ x,
                     double *__restrict__ y)
    {
      __shared__ double xs[SPMV_MAX_ROW];
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = SPMV_THREADS / 32;
      for (int bi = blockIdx.x; bi < n_block_rows; bi += gridDim.x)
        {
          const int     b   = row_list ? row_list[bi] : bi;
          const int64_t kb  = brow_ptr[b];
          const int     nb  = (int)(brow_ptr[b + 1] - kb);
          const int     len = nb * n;
          __syncthreads(); // previous block row is done with xs
          for (int e = threadIdx.x; e < len; e += SPMV_THREADS)
            {
              const int k = e / n, j = e - k * n;
              xs[e]       = __ldg(&x[(int64_t)bcol[kb + k] * n + j]);
            }
          __syncthreads();
          const double *rows = vals + kb * n * n;
          // RPW rows in flight per warp: short rows (n = 27: 189 doubles) need the
          // memory-level parallelism, long rows (n = 64) are fine either way
          constexpr int RPW = 4;
          for (int i = warp; i < n; i += RPW * nwarp)
            {
              const double *r[RPW];
              double        s[RPW];
#pragma unroll
              for (int u = 0; u < RPW; ++u)
                {
                  const int iu = i + u * nwarp;
                  r[u]         = rows + (int64_t)(iu < n ? iu : i) * len;
                  s[u]         = 0.;
                }
              for (int e = lane; e < len; e += 32)
                {
                  const double xe = xs[e];
#pragma unroll
                  for (int u = 0; u < RPW; ++u)
                    s[u] += r[u][e] * xe;
                }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < RPW; ++u)
                  s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
              if (lane == 0)
                {
#pragma unroll
                  for (int u = 0; u < RPW; ++u)
                    if (i + u * nwarp < n)
                      {
                        const int64_t row = (int64_t)b * n + i + u * nwarp;
                        y[row]            = ADD ? y[row] + s[u] : s[u];
                      }
                }
            }
        }
    }

    // fallback for block rows that do not fit the shared-memory row buffer
    template <bool ADD>
    __global__ void __launch_bounds__(256)
    k_spmv_warp_per_row(const double *__restrict__ vals,
                        const int64_t *__restrict__ brow_ptr,
                        const int32_t *__restrict__ bcol,
                        const int32_t *__restrict__ row_stride,
                        const int      n,
                        const int64_t  n_rows,
                        const double *__restrict__ x,
                        double *__restrict__ y)
    {
      const int     lane  = threadIdx.x & 31;
      const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
      const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
      for (int64_t r = warp0; r < n_rows; r += nwarp)
        {
          const int64_t b   = r / n;
          const int     i   = (int)(r - b * n);
          const int64_t kb  = brow_ptr[b];
          const int     len = row_stride[b];
          const double *row = vals + kb * n * n + (int64_t)i * len;
          double        s   = 0.;
          for (int e = lane; e < len; e += 32)
            {
              const int k = e / n, j = e - k * n;
              s += row[e] * __ldg(&x[(int64_t)bcol[kb + k] * n + j]);
            }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane == 0)
            y[r] = ADD ? y[r] + s : s;
        }
    }

    __global__ void __launch_bounds__(256)
    k_diag_inverse(const double *__restrict__ vals,
                   const int64_t *__restrict__ diag_base,
                   const int32_t *__restrict__ dof_block,
                   const int32_t *__restrict__ row_stride,
                   const int     n,
                   const int32_t np,
                   double *__restrict__ out)
    {
      const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (int64_t)np * n)
        return;
      const int    p = (int)(idx / n), i = (int)(idx - (int64_t)p * n);
      const int    b = dof_block[p];
      const double d = vals[diag_base[p] + (int64_t)i * row_stride[b] + i];
      // include/utils.h:808-813: invert only entries above 1e-10
      out[(int64_t)b * n + i] = fabs(d) > 1e-10 ? 1. / d : d;
    }
  } // namespace

  bool
  spmv_can_split(pd_handle *h)
  {
    if (h->max_row_len < 0)
      {
        int64_t m = 0;
        for (int32_t b = 0; b < h->np_own; ++b)
          m = std::max<int64_t>(m, (h->h_brow_ptr[b + 1] - h->h_brow_ptr[b]) * h->n);
        h->max_row_len = m;
      }
    return h->assembled && h->np != h->np_own && h->max_row_len <= SPMV_MAX_ROW;
  }

  void
  launch_spmv(pd_handle *h, const double *src, double *dst, const bool add, const int part)
  {
    // part 0: all block rows; 1: rows without ghost columns; 2: rows with ghost columns (sharded handles)
    if (part != 0 && h->spmv_list_interior.n + h->spmv_list_boundary.n != (size_t)h->np_own)
      {
        std::vector<int32_t> inner, outer;
        for (int32_t b = 0; b < h->np_own; ++b)
          {
            bool ghost = false;
            for (int64_t e = h->h_brow_ptr[b]; e < h->h_brow_ptr[b + 1]; ++e)
              ghost = ghost || h->h_bcol[e] >= h->np_own;
            (ghost ? outer : inner).push_back(b);
          }
        auto put = [](auto &buf, const auto &v) {
          buf.alloc(v.size());
          if (!v.empty())
            PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
        };
        put(h->spmv_list_interior, inner);
        put(h->spmv_list_boundary, outer);
      }
    const int32_t *list   = part == 0 ? nullptr : (part == 1 ? h->spmv_list_interior.p : h->spmv_list_boundary.p);
    const int32_t  n_rows = part == 0 ? h->np_own : (int32_t)(part == 1 ? h->spmv_list_interior.n : h->spmv_list_boundary.n);
    if (n_rows == 0)
      return;
    if (h->max_row_len < 0)
      {
        int64_t m = 0;
        for (int32_t b = 0; b < h->np_own; ++b)
          m = std::max<int64_t>(m, (h->h_brow_ptr[b + 1] - h->h_brow_ptr[b]) * h->n);
        h->max_row_len = m;
      }
    if (h->max_row_len <= SPMV_MAX_ROW)
      {
        const int grid = (int)std::min<int64_t>(n_rows, (int64_t)h->sm_count * 8);
        if (add)
          k_spmv_block_row<true><<<grid, SPMV_THREADS, 0, h->stream>>>(h->values.p, h->brow_ptr.p, h->bcol.p, h->n, n_rows,
                                                                      list, src, dst);
        else
          k_spmv_block_row<false><<<grid, SPMV_THREADS, 0, h->stream>>>(h->values.p, h->brow_ptr.p, h->bcol.p, h->n, n_rows,
                                                                       list, src, dst);
      }
    else
      {
        if (part == 2)
          return; // the fallback kernel has no row list: part 1 did all rows (callers order 1 after the exchange)
        const int     tb   = 256;
        const int64_t rows = h->n_dofs;
        const int64_t want = (rows * 32 + tb - 1) / tb;
        const int     grid = (int)std::min<int64_t>(want, (int64_t)h->sm_count * 16);
        if (add)
          k_spmv_warp_per_row<true><<<grid, tb, 0, h->stream>>>(h->values.p, h->brow_ptr.p, h->bcol.p, h->row_stride.p,
                                                               h->n, rows, src, dst);
        else
          k_spmv_warp_per_row<false><<<grid, tb, 0, h->stream>>>(h->values.p, h->brow_ptr.p, h->bcol.p, h->row_stride.p,
                                                                h->n, rows, src, dst);
      }
    ++h->launches;
    PD_CUDA(cudaGetLastError());
  }

  void
  launch_diagonal_inverse(pd_handle *h, double *dst)
  {
    const int64_t nd = (int64_t)h->np_own * h->n;
    k_diag_inverse<<<(unsigned)((nd + 255) / 256), 256, 0, h->stream>>>(h->values.p, h->diag_base.p, h->dof_block.p,
                                                                      h->row_stride.p, h->n, h->np_own, dst);
    ++h->launches;
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
