// -----------------------------------------------------------------------------
// pd_vmult.cu -- operator apply with the assembled matrix.
//
// Reference semantics: the matrix-based vmult used on agglomerated levels,
// LinearOperatorMG::vmult -> TrilinosWrappers::SparseMatrix::vmult
// (include/multigrid_amg.h:345-355, include/linear_operator_for_mg.h:295;
// CG at examples/diffusion_reaction.cc:721-724), and the inverse diagonal of
// include/utils.h:797-814.
//
// HBM-bound: 8 n^2 B per block + 4 B per block index + 16 B per DoF.  The value
// array is the scalar CSR of the reference pattern, so a scalar row is one
// contiguous run of nb*n doubles: one warp streams one row with unit-stride
// loads; the source vector is gathered per block (n contiguous doubles).
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

namespace pd
{
  namespace
  {
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

  void
  launch_spmv(pd_handle *h, const double *src, double *dst, const bool add)
  {
    const int     tb   = 256;
    const int64_t rows = h->n_dofs;
    const int64_t want = (rows * 32 + tb - 1) / tb;
    const int     grid = (int)std::min<int64_t>(want, (int64_t)h->sm_count * 16);
    if (add)
      k_spmv_warp_per_row<true><<<grid, tb, 0, h->stream>>>(h->values.p, h->brow_ptr.p, h->bcol.p, h->row_stride.p, h->n,
                                                           rows, src, dst);
    else
      k_spmv_warp_per_row<false><<<grid, tb, 0, h->stream>>>(h->values.p, h->brow_ptr.p, h->bcol.p, h->row_stride.p,
                                                            h->n, rows, src, dst);
    ++h->launches;
    PD_CUDA(cudaGetLastError());
  }

  void
  launch_diagonal_inverse(pd_handle *h, double *dst)
  {
    const int64_t nd = (int64_t)h->np * h->n;
    k_diag_inverse<<<(unsigned)((nd + 255) / 256), 256, 0, h->stream>>>(h->values.p, h->diag_base.p, h->dof_block.p,
                                                                      h->row_stride.p, h->n, h->np, dst);
    ++h->launches;
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
