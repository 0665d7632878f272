// -----------------------------------------------------------------------------
// pd_cartesian.cu -- SIP-DG assembly over agglomerates of AXIS-ALIGNED cells by
// per-sub-cell sum factorisation (the tensor path of pd_assemble).
//
// Same result as pd_assemble.cu (reference: PolyUtils::assemble_dg_matrix,
// include/poly_utils.h:2000-2195, kernels :1870-1926, through reinit / MappingBox,
// source/agglomeration_handler.cc:729-906, source/mapping_box.cc:393-439), another
// algorithm.  On a sub-cell [a, a + eta] that is an axis-aligned box, the agglomerated
// quadrature QGauss<dim>(n_q) is a TENSOR grid in the bounding-box coordinates and the basis
// is a tensor product, so every integral the reference sums point by point factorises into
// 1-D sums:
//   int_cell  d_x phi_i d_x phi_j = K^x[a,a'] M^y[b,b'] M^z[c,c'],   M^d[a,a'] = eta_d sum_q w_q l_a l_a'(xhat_q),
//                                                                    K^d[a,a'] = eta_d sum_q w_q l_a' l_a''(xhat_q) / h_d^2
//   int_face(normal x) (-1/2 dn phi_i phi_j - 1/2 phi_i dn phi_j + pen phi_i phi_j)
//                                  = F^x[a,a'] M^y[b,b'] M^z[c,c'],  F^x from l, l' at the face coordinate
// and the block of a polytope is  sum over its sub-cells / sub-faces of  X (x) Y (x) Z  -- a
// rank-structured sum with n^2 (2 + 1/S) multiply-adds per sub-cell instead of the 2 n^2 dim n_q^dim
// of the point-wise contraction (p = 3: 8.2 k instead of 1.57 M flops per sub-cell).  The sums over the
// quadrature points are the SAME sums the reference forms, regrouped; results agree to rounding
// (tests: per block entry <= 1e-12 against the oracle, and against the DMMA kernels).
//
// Applies when every owned sub-cell is an axis-aligned box in standard orientation (checked on the
// device by k_check_axis_aligned at pd_create / pd_upload); distorted meshes take the DMMA kernels
// of pd_assemble.cu.  FE_DGQ and FE_AggloDGP (the latter is a sub-set of the same tensor index set).
//
//   k_cart_diag     one CTA per polytope: diagonal block = volume + boundary faces + own side
//                   (M11 / M22) of every interior interface, accumulated in registers in the fixed
//                   order sub-cells, then adjacency list -- deterministic, no atomics, no partials
//   k_cart_offdiag  one CTA per interior interface: M12 from cross matrices of the two bounding-box
//                   bases, written to (A,B) and, transposed (M21 = M12^T), to (B,A)
// Thread layout: thread (col, rg) owns the column col = (b,b',c,c') of the factorised block and
// RPT = N1^2 / RG rows (a,a'); per item it forms its Y (x) Z entry from two shared-memory loads and
// sweeps its rows with broadcast loads of X -- the 1-D matrices are the only shared-memory state.
// Roofline: the matrix is written once (8 n^2 bytes per block: config C 7.3 GB) => HBM-bound.
// -----------------------------------------------------------------------------
#include "pd_host.hpp"
#include "pd_internal.hpp"
#include "pd_device.cuh"

#include <cstdlib>

namespace pd
{
  namespace
  {
    struct CartArgs
    {
      const double  *verts;
      const int32_t *cell_verts;
      const int64_t *subcell_ptr;
      const int32_t *subcell_idx;
      const double  *bbox;
      const int32_t *ifA, *ifB;
      const int64_t *if_sub_ptr;
      const int32_t *sub_cell, *sub_face;
      const double  *sub_sigma;
      const int64_t *padj_ptr, *padj;
      const int64_t *diag_base, *if_baseAB, *if_baseBA;
      const int32_t *dof_block, *row_stride;
      double        *values;
      double         stiffness, mass;
      uint32_t       flags;
      int32_t        np_own, n_ifaces, nq, nqf;
      Basis1D        basis;
      Quad1D         quad, quadf;
      unsigned char  dof_abc[64][3]; // (a, b, c) of every DoF of the element
    };

    // cells of the mesh that are NOT axis-aligned boxes in standard orientation (vertex v at lo + bit_k(v) (hi - lo))
    __global__ void __launch_bounds__(256)
    k_check_axis_aligned(const double *verts, const int32_t *cell_verts, const int32_t *subcell_idx, const int64_t n_sub,
                         const int dim, int *n_bad)
    {
      const int vpc = 1 << dim;
      int       bad = 0;
      for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sub; s += (int64_t)gridDim.x * blockDim.x)
        {
          const int32_t *cv = cell_verts + (int64_t)subcell_idx[s] * vpc;
          const double  *lo = verts + (int64_t)cv[0] * dim, *hi = verts + (int64_t)cv[vpc - 1] * dim;
          for (int k = 0; k < dim; ++k)
            bad |= !(hi[k] > lo[k]);
          for (int v = 1; v < vpc - 1; ++v)
            {
              const double *x = verts + (int64_t)cv[v] * dim;
              for (int k = 0; k < dim; ++k)
                bad |= x[k] != (((v >> k) & 1) ? hi[k] : lo[k]);
            }
        }
      if (__syncthreads_or(bad) && threadIdx.x == 0)
        atomicAdd(n_bad, 1);
    }

    template <int DIM, int DEGX>
    struct CartCfg
    {
      using C                 = Cfg<DIM, DEGX>;
      static constexpr int N1 = C::N1, NX = N1 * N1, NYZ = ipow(NX, DIM - 1), NF = ipow(N1, DIM);
      // row groups: threads = NYZ * RG, RPT rows each
      static constexpr int RG   = NYZ >= 256 ? 1 : (NYZ * NX <= 1024 && NYZ < 64 ? NX : (NX % 3 == 0 ? 3 : (NX % 2 == 0 ? 2 : 1)));
      static constexpr int RPT  = NX / RG;
      static constexpr int NTHR = ((NYZ * RG + 31) / 32) * 32;
      static constexpr int CH   = 16; // items per chunk
      static_assert(NX % RG == 0, "row groups");
    };

    // rows a of the 1-D matrices of one item along one axis, [a'] = 0..N1-1:
    //  cell / tangential axis:  M = eta sum_q w l_a l_a',  K = eta sum_q w l_a' l_a'' (derivatives per real length)
    template <class C>
    __device__ __forceinline__ void
    axis_rows_mass(const Basis1D &B, const double *qx, const double *qw, const int nq, const double c_lo, const double eta,
                   const double b_lo, const double inv_h, const int a, double *M, double *K)
    {
      constexpr int N1 = C::N1;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        M[k] = K[k] = 0.;
      for (int q = 0; q < nq; ++q)
        {
          double       L[N1], dL[N1];
          const double xh = (c_lo + eta * qx[q] - b_lo) * inv_h;
          basis_1d<C>(B, xh, inv_h, L, dL);
          const double w = qw[q] * eta;
          double       la = 0., da = 0.;
#pragma unroll
          for (int k = 0; k < N1; ++k) // runtime row index without dynamic register indexing
            if (k == a)
              {
                la = L[k];
                da = dL[k];
              }
#pragma unroll
          for (int k = 0; k < N1; ++k)
            {
              M[k] += w * la * L[k];
              K[k] += w * da * dL[k];
            }
        }
    }

    // two bases (A rows, B columns) over the same interval: cross mass rows
    template <class C>
    __device__ __forceinline__ void
    axis_rows_cross(const Basis1D &B, const double *qx, const double *qw, const int nq, const double c_lo, const double eta,
                    const double a_lo, const double a_inv_h, const double b_lo, const double b_inv_h, const int a, double *M)
    {
      constexpr int N1 = C::N1;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        M[k] = 0.;
      for (int q = 0; q < nq; ++q)
        {
          double       LA[N1], dLA[N1], LB[N1], dLB[N1];
          const double x = c_lo + eta * qx[q];
          basis_1d<C>(B, (x - a_lo) * a_inv_h, a_inv_h, LA, dLA);
          basis_1d<C>(B, (x - b_lo) * b_inv_h, b_inv_h, LB, dLB);
          const double w = qw[q] * eta;
          double       la = 0.;
#pragma unroll
          for (int k = 0; k < N1; ++k)
            if (k == a)
              la = LA[k];
#pragma unroll
          for (int k = 0; k < N1; ++k)
            M[k] += w * la * LB[k];
        }
    }

    template <int N1>
    __device__ __forceinline__ double
    pick(const double *v, const int a)
    {
      double r = 0.;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        if (k == a)
          r = v[k];
      return r;
    }

    // box of a cell: lo corner = vertex 0, hi corner = vertex 2^dim - 1
    template <int DIM>
    __device__ __forceinline__ void
    cell_box(const CartArgs &A, const int32_t c, const int d, double &lo, double &hi)
    {
      const int32_t *cv = A.cell_verts + (int64_t)c * (1 << DIM);
      lo                = A.verts[(int64_t)cv[0] * DIM + d];
      hi                = A.verts[(int64_t)cv[(1 << DIM) - 1] * DIM + d];
    }

    // acc[r] += X1[r] yz1 + X2[r] yz2 over the items of a chunk.  SL[item][d][2][NX]: slot 0 = M-like, slot 1 = K-like.
    // kind[item]: 1 = cell (two terms), 2 = face (one term, F sits in slot 0 of its normal axis), 0 = empty.
    template <int DIM, int DEGX>
    __device__ __forceinline__ void
    accumulate_chunk(const double *SL, const int *kind, const int cnt, const int col, const int rg, const double stiffness,
                     const double mass, double *acc)
    {
      using CC            = CartCfg<DIM, DEGX>;
      constexpr int NX    = CC::NX, RPT = CC::RPT;
      constexpr int ISTR  = DIM * 2 * NX;
      const int     cb    = DIM == 3 ? col % NX : col; // (b,b')
      const int     cc    = DIM == 3 ? col / NX : 0;   // (c,c')
      for (int it = 0; it < cnt; ++it)
        {
          const double *s  = SL + it * ISTR;
          const int     kd = kind[it];
          if (kd == 0)
            continue;
          const double m1 = s[(1 * 2 + 0) * NX + cb];
          const double m2 = DIM == 3 ? s[(2 * 2 + 0) * NX + cc] : 1.;
          const double yz1 = m1 * m2;
          if (kd == 1)
            {
              const double k1  = s[(1 * 2 + 1) * NX + cb];
              const double k2  = DIM == 3 ? s[(2 * 2 + 1) * NX + cc] : 0.;
              const double yz2 = stiffness * (k1 * m2 + m1 * k2) + mass * yz1;
              const double sy1 = stiffness * yz1;
              const double *x1 = s + (0 * 2 + 1) * NX + rg * RPT; // K along x
              const double *x0 = s + (0 * 2 + 0) * NX + rg * RPT; // M along x
#pragma unroll
              for (int r = 0; r < RPT; ++r)
                acc[r] += x1[r] * sy1 + x0[r] * yz2;
            }
          else
            {
              const double *x0 = s + (0 * 2 + 0) * NX + rg * RPT;
#pragma unroll
              for (int r = 0; r < RPT; ++r)
                acc[r] += x0[r] * yz1;
            }
        }
    }

    // registers -> OUT[full row index][full column index] (shared, NF x (NF + 1))
    template <int DIM, int DEGX>
    __device__ __forceinline__ void
    stage_block(double *OUT, const int col, const int rg, const double *acc)
    {
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, RPT = CC::RPT, NF = CC::NF;
      const int     cb = DIM == 3 ? col % NX : col, cc = DIM == 3 ? col / NX : 0;
      const int     b = cb / N1, bp = cb % N1, c = cc / N1, cp = cc % N1;
#pragma unroll
      for (int r = 0; r < RPT; ++r)
        {
          const int ra = rg * RPT + r;
          const int a = ra / N1, ap = ra % N1;
          const int i = a + N1 * (b + N1 * c), j = ap + N1 * (bp + N1 * cp);
          OUT[i * (NF + 1) + j] = acc[r];
        }
    }

    template <int DIM, int DEGX>
    __global__ void __launch_bounds__(CartCfg<DIM, DEGX>::NTHR)
    k_cart_diag(const CartArgs A)
    {
      using C          = Cfg<DIM, DEGX>;
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, NYZ = CC::NYZ, RG = CC::RG, RPT = CC::RPT, NF = CC::NF, CH = CC::CH, N = C::N;
      constexpr int ISTR = DIM * 2 * NX;
      __shared__ double SL[CH * ISTR > NF * (NF + 1) ? CH * ISTR : NF * (NF + 1)];
      __shared__ int    kind[CH];
      const int tid = threadIdx.x;
      const int col = tid % NYZ, rg = tid / NYZ;
      const bool active = tid < NYZ * RG;

      for (int p = blockIdx.x; p < A.np_own; p += gridDim.x)
        {
          double acc[RPT];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
            acc[r] = 0.;
          const double *bb = A.bbox + (int64_t)p * 2 * DIM;
          const int64_t s0 = A.subcell_ptr[p], s1 = A.subcell_ptr[p + 1];
          const int64_t k0 = A.padj_ptr[p], k1 = A.padj_ptr[p + 1];
          // segment -1: the sub-cells; segments k0..k1-1: the faces of the adjacency list
          for (int64_t seg = k0 - 1; seg < k1; ++seg)
            {
              int64_t i0, i1;
              int     side = 0;
              int64_t f    = -1;
              bool    bnd  = false;
              if (seg < k0)
                {
                  if (!(A.flags & PD_ASSEMBLE_VOLUME))
                    continue;
                  i0 = s0;
                  i1 = s1;
                }
              else
                {
                  const int64_t e = A.padj[seg];
                  f               = e >> 1;
                  side            = (int)(e & 1);
                  bnd             = A.ifB[f] < 0;
                  if (!(bnd ? (A.flags & PD_ASSEMBLE_BOUNDARY) : (A.flags & PD_ASSEMBLE_INTERIOR)))
                    continue;
                  i0 = A.if_sub_ptr[f];
                  i1 = A.if_sub_ptr[f + 1];
                }
              for (int64_t c0 = i0; c0 < i1; c0 += CH)
                {
                  const int cnt = (int)(i1 - c0 < CH ? i1 - c0 : CH);
                  __syncthreads(); // the previous chunk has been consumed
                  // ---- 1-D matrices of the chunk's items: thread = (item, axis, row a)
                  for (int w = tid; w < cnt * DIM * N1; w += CC::NTHR)
                    {
                      const int it = w / (DIM * N1), d = (w / N1) % DIM, a = w % N1;
                      double    M[N1], K[N1];
                      const double b_lo = bb[d], inv_h = 1. / (bb[DIM + d] - bb[d]);
                      if (seg < k0)
                        {
                          double lo, hi;
                          cell_box<DIM>(A, A.subcell_idx[c0 + it], d, lo, hi);
                          axis_rows_mass<C>(A.basis, A.quad.x, A.quad.w, A.nq, lo, hi - lo, b_lo, inv_h, a, M, K);
                          if (d == 0 && a == 0)
                            kind[it] = 1;
                        }
                      else
                        {
                          const int64_t s  = c0 + it;
                          const int     lf = A.sub_face[s], fd = lf >> 1, fs = lf & 1;
                          double        lo, hi;
                          cell_box<DIM>(A, A.sub_cell[s], d, lo, hi);
                          if (d == fd)
                            {
                              // own-side face term at the face coordinate: outward normal of THIS polytope
                              const double nrm = (fs ? 1. : -1.) * (side ? -1. : 1.);
                              const double cf  = bnd ? 1. : 0.5;
                              double       L[N1], dL[N1];
                              basis_1d<C>(A.basis, ((fs ? hi : lo) - b_lo) * inv_h, inv_h, L, dL);
                              const double la = pick<N1>(L, a), da = pick<N1>(dL, a), pen = A.sub_sigma[s];
#pragma unroll
                              for (int k = 0; k < N1; ++k)
                                {
                                  M[k] = A.stiffness * (-cf * nrm * (da * L[k] + la * dL[k]) + pen * la * L[k]);
                                  K[k] = 0.;
                                }
                            }
                          else
                            axis_rows_mass<C>(A.basis, A.quadf.x, A.quadf.w, A.nqf, lo, hi - lo, b_lo, inv_h, a, M, K);
                          if (d == 0 && a == 0)
                            kind[it] = 2;
                        }
                      double *dst = SL + it * ISTR + d * 2 * NX + a * N1;
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        {
                          dst[k]      = M[k];
                          dst[NX + k] = K[k];
                        }
                    }
                  __syncthreads();
                  if (active)
                    accumulate_chunk<DIM, DEGX>(SL, kind, cnt, col, rg, A.stiffness, A.mass, acc);
                }
            }
          // ---- epilogue: registers -> shared tile -> the diagonal block of the CSR rows
          __syncthreads();
          if (active)
            stage_block<DIM, DEGX>(SL, col, rg, acc);
          __syncthreads();
          const int64_t base   = A.diag_base[p];
          const int     stride = A.row_stride[A.dof_block[p]];
          for (int idx = tid; idx < N * N; idx += CC::NTHR)
            {
              const int i = idx / N, j = idx - i * N;
              const int fi = A.dof_abc[i][0] + N1 * (A.dof_abc[i][1] + N1 * A.dof_abc[i][2]);
              const int fj = A.dof_abc[j][0] + N1 * (A.dof_abc[j][1] + N1 * A.dof_abc[j][2]);
              A.values[base + (int64_t)i * stride + j] = SL[fi * (NF + 1) + fj];
            }
        }
    }

    template <int DIM, int DEGX>
    __global__ void __launch_bounds__(CartCfg<DIM, DEGX>::NTHR)
    k_cart_offdiag(const CartArgs A)
    {
      using C          = Cfg<DIM, DEGX>;
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, NYZ = CC::NYZ, RG = CC::RG, RPT = CC::RPT, NF = CC::NF, CH = CC::CH, N = C::N;
      constexpr int ISTR = DIM * 2 * NX;
      __shared__ double SL[CH * ISTR > NF * (NF + 1) ? CH * ISTR : NF * (NF + 1)];
      __shared__ int    kind[CH];
      const int tid = threadIdx.x;
      const int col = tid % NYZ, rg = tid / NYZ;
      const bool active = tid < NYZ * RG;

      for (int f = blockIdx.x; f < A.n_ifaces; f += gridDim.x)
        {
          const int32_t pa = A.ifA[f], pb = A.ifB[f];
          if (pb < 0)
            continue;
          double acc[RPT];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
            acc[r] = 0.;
          const double *ba = A.bbox + (int64_t)pa * 2 * DIM, *bbx = A.bbox + (int64_t)pb * 2 * DIM;
          const int64_t i0 = A.if_sub_ptr[f], i1 = A.if_sub_ptr[f + 1];
          for (int64_t c0 = i0; c0 < i1; c0 += CH)
            {
              const int cnt = (int)(i1 - c0 < CH ? i1 - c0 : CH);
              __syncthreads();
              for (int w = tid; w < cnt * DIM * N1; w += CC::NTHR)
                {
                  const int     it = w / (DIM * N1), d = (w / N1) % DIM, a = w % N1;
                  const int64_t s  = c0 + it;
                  const int     lf = A.sub_face[s], fd = lf >> 1, fs = lf & 1;
                  double        lo, hi, M[N1];
                  cell_box<DIM>(A, A.sub_cell[s], d, lo, hi);
                  const double a_lo = ba[d], a_ih = 1. / (ba[DIM + d] - ba[d]);
                  const double b_lo = bbx[d], b_ih = 1. / (bbx[DIM + d] - bbx[d]);
                  if (d == fd)
                    {
                      // M12 = sum w [ 1/2 (dn phi0_i) phi1_j - 1/2 phi0_i (dn phi1_j) - pen phi0_i phi1_j ], n = A's normal
                      const double nrm = fs ? 1. : -1., x = fs ? hi : lo, pen = A.sub_sigma[s];
                      double       LA[N1], dLA[N1], LB[N1], dLB[N1];
                      basis_1d<C>(A.basis, (x - a_lo) * a_ih, a_ih, LA, dLA);
                      basis_1d<C>(A.basis, (x - b_lo) * b_ih, b_ih, LB, dLB);
                      const double la = pick<N1>(LA, a), da = pick<N1>(dLA, a);
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        M[k] = A.stiffness * (0.5 * nrm * (da * LB[k] - la * dLB[k]) - pen * la * LB[k]);
                    }
                  else
                    axis_rows_cross<C>(A.basis, A.quadf.x, A.quadf.w, A.nqf, lo, hi - lo, a_lo, a_ih, b_lo, b_ih, a, M);
                  if (d == 0 && a == 0)
                    kind[it] = 2;
                  double *dst = SL + it * ISTR + d * 2 * NX + a * N1;
#pragma unroll
                  for (int k = 0; k < N1; ++k)
                    dst[k] = M[k];
                }
              __syncthreads();
              if (active)
                accumulate_chunk<DIM, DEGX>(SL, kind, cnt, col, rg, 1., 0., acc);
            }
          __syncthreads();
          if (active)
            stage_block<DIM, DEGX>(SL, col, rg, acc);
          __syncthreads();
          const int64_t baseAB = A.if_baseAB[f], baseBA = A.if_baseBA[f];
          const int     strideA = A.row_stride[A.dof_block[pa]];
          const int     strideB = baseBA >= 0 ? A.row_stride[A.dof_block[pb]] : 0;
          for (int idx = tid; idx < N * N; idx += CC::NTHR)
            {
              const int i = idx / N, j = idx - i * N;
              const int fi = A.dof_abc[i][0] + N1 * (A.dof_abc[i][1] + N1 * A.dof_abc[i][2]);
              const int fj = A.dof_abc[j][0] + N1 * (A.dof_abc[j][1] + N1 * A.dof_abc[j][2]);
              A.values[baseAB + (int64_t)i * strideA + j] = SL[fi * (NF + 1) + fj];
              if (baseBA >= 0) // M21 = M12^T; coalesced over j as well: entry (i, j) of (B,A) is M12(j, i)
                A.values[baseBA + (int64_t)i * strideB + j] = SL[fj * (NF + 1) + fi];
            }
          __syncthreads();
        }
    }

    template <int DIM, int DEGX>
    void
    run_cart(pd_handle *h, const CartArgs &a)
    {
      using CC = CartCfg<DIM, DEGX>;
      PD_CUDA(cudaEventRecord(h->ev[0], h->stream));
      {
        const int grid = (int)std::min<int64_t>(h->np_own, (int64_t)h->sm_count * 16);
        k_cart_diag<DIM, DEGX><<<grid, CC::NTHR, 0, h->stream>>>(a);
        ++h->launches;
      }
      PD_CUDA(cudaEventRecord(h->ev[1], h->stream));
      if ((a.flags & PD_ASSEMBLE_INTERIOR) && h->n_ifaces > 0)
        {
          const int grid = (int)std::min<int64_t>(h->n_ifaces, (int64_t)h->sm_count * 16);
          k_cart_offdiag<DIM, DEGX><<<grid, CC::NTHR, 0, h->stream>>>(a);
          ++h->launches;
        }
      PD_CUDA(cudaEventRecord(h->ev[2], h->stream));
      PD_CUDA(cudaEventRecord(h->ev[3], h->stream));
      PD_CUDA(cudaGetLastError());
    }
  } // namespace

  // every owned sub-cell an axis-aligned box?  (device scan of the uploaded mesh; one 4-byte read-back)
  bool
  check_axis_aligned(pd_handle *h)
  {
    if (h->n_subcells == 0)
      return false;
    if (!h->cart_flag.p)
      h->cart_flag.alloc(1);
    PD_CUDA(cudaMemsetAsync(h->cart_flag.p, 0, sizeof(int), h->stream));
    const int grid = (int)std::min<int64_t>((h->n_subcells + 255) / 256, (int64_t)h->sm_count * 8);
    k_check_axis_aligned<<<grid, 256, 0, h->stream>>>(h->verts.p, h->cell_verts.p, h->subcell_idx.p, h->n_subcells, h->dim,
                                                       h->cart_flag.p);
    int bad = 1;
    PD_CUDA(cudaMemcpyAsync(&bad, h->cart_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PD_CUDA(cudaStreamSynchronize(h->stream));
    return bad == 0;
  }

  bool
  cartesian_assembly_selected(const pd_handle *h)
  {
    const char *e = getenv("PD_ASSEMBLE_KERNELS"); // "generic": the DMMA kernels everywhere (A/B measurements, tests)
    if (e && e[0] == 'g')
      return false;
    return h->cartesian;
  }

  void
  launch_assemble_cartesian(pd_handle *h, const uint32_t flags, const pd_coefficients &coef)
  {
    CartArgs a{};
    a.verts       = h->verts.p;
    a.cell_verts  = h->cell_verts.p;
    a.subcell_ptr = h->subcell_ptr.p;
    a.subcell_idx = h->subcell_idx.p;
    a.bbox        = h->bbox.p;
    a.ifA         = h->ifA.p;
    a.ifB         = h->ifB.p;
    a.if_sub_ptr  = h->if_sub_ptr.p;
    a.sub_cell    = h->sub_cell.p;
    a.sub_face    = h->sub_face.p;
    a.sub_sigma   = h->sub_sigma.p;
    a.padj_ptr    = h->padj_ptr.p;
    a.padj        = h->padj.p;
    a.diag_base   = h->diag_base.p;
    a.if_baseAB   = h->if_baseAB.p;
    a.if_baseBA   = h->if_baseBA.p;
    a.dof_block   = h->dof_block.p;
    a.row_stride  = h->row_stride.p;
    a.values      = h->values.p;
    a.stiffness   = coef.stiffness;
    a.mass        = coef.mass;
    a.flags       = flags;
    a.np_own      = h->np_own;
    a.n_ifaces    = h->n_ifaces;
    a.nq          = h->nq1;
    a.nqf         = h->nq1f;
    a.basis       = h->basis;
    a.quad        = h->quad;
    a.quadf       = h->quadf;
    {
      // (a, b, c) of DoF i: FE_DGQ lexicographic; FE_AggloDGP in PolynomialSpace order (last coordinate outermost,
      // first fastest, total degree <= p: source/fe_agglodgp.cc:28-57)
      const int n1 = h->n1, p = h->degree;
      int       i  = 0;
      for (int c = 0; c < (h->dim == 3 ? n1 : 1); ++c)
        for (int b = 0; b < n1; ++b)
          for (int aa = 0; aa < n1; ++aa)
            {
              if (h->fe_kind == PD_FE_AGGLODGP && aa + b + c > p)
                continue;
              a.dof_abc[i][0] = (unsigned char)aa;
              a.dof_abc[i][1] = (unsigned char)b;
              a.dof_abc[i][2] = (unsigned char)c;
              ++i;
            }
    }
    const int key = h->fe_kind * 100 + h->dim * 10 + h->degree;
    switch (key)
      {
        case 121: run_cart<2, DGP_BASE + 1>(h, a); break;
        case 122: run_cart<2, DGP_BASE + 2>(h, a); break;
        case 123: run_cart<2, DGP_BASE + 3>(h, a); break;
        case 124: run_cart<2, DGP_BASE + 4>(h, a); break;
        case 131: run_cart<3, DGP_BASE + 1>(h, a); break;
        case 132: run_cart<3, DGP_BASE + 2>(h, a); break;
        case 133: run_cart<3, DGP_BASE + 3>(h, a); break;
        case 21: run_cart<2, 1>(h, a); break;
        case 22: run_cart<2, 2>(h, a); break;
        case 23: run_cart<2, 3>(h, a); break;
        case 24: run_cart<2, 4>(h, a); break;
        case 31: run_cart<3, 1>(h, a); break;
        case 32: run_cart<3, 2>(h, a); break;
        case 33: run_cart<3, 3>(h, a); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no assembly kernel for this (dim, degree)", __LINE__};
      }
  }
} // namespace pd
