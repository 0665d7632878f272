// -----------------------------------------------------------------------------
// pd_polyapply.cu -- matrix-free SIP operator apply on AGGLOMERATED polytopes.
//
// y = A x without the matrix: the operand rows that pd_assemble.cu contracts
// into blocks are generated the same way (MappingBox + FE_DGQ on the bounding
// box at the agglomerated quadrature points, source/agglomeration_handler.cc:
// 729-906, source/mapping_box.cc:393-532) and applied to the polytope's
// coefficient vector instead:
//   volume   y_P  += G^T diag(w c) G u_P              (include/poly_utils.h:2038-2052)
//   faces    [y_A; y_B] += Z (V'^T u) + V' (Z^T u),   u = [u_A; u_B]
//            (the T + T^T form of poly_utils.h:1870-1926, see pd_assemble.cu)
// The reference has no matrix-free operator on the agglomerated space
// (SURVEY.md, fact 3); this is the memory-free alternative to the block-CSR
// apply: no 8 n^2 bytes per block, at the price of regenerating the basis
// (12 n flops per volume point).  On the configurations of SURVEY 8d the
// block-CSR apply, which runs at the HBM roofline, is faster whenever the matrix
// fits; this path exists for the cases where it does not.
//
// Deterministic: every work item (volume item, interface) writes its own partial
// result vector; k_poly_gather sums them per polytope in a fixed order.
// Shared-memory layout as in the assembly kernels (DoF-major panels, row
// stride = 4 mod 16, lane = quadrature point): conflict free.
// -----------------------------------------------------------------------------
#include "pd_device.cuh"
#include "pd_host.hpp"

#include <algorithm>

namespace pd
{
  void plan_volume_items(pd_handle *h, int tq, int grid); // pd_assemble.cu

  namespace
  {
    constexpr int TQ = 32;
    constexpr int NW = 8; // warps per CTA

    struct VolApplyArgs
    {
      const double  *vq_x, *vq_w;
      int64_t        Q;
      const double  *bbox;
      const int32_t *dof_block;
      const int32_t *item_poly;
      const int64_t *item_q0, *item_q1;
      int32_t        n_items;
      const double  *x;
      double        *partial; // [n_items][N]
      double         stiffness, mass;
      Basis1D        basis;
    };

    template <int DIM, int DEG, bool MASS>
    __global__ void __launch_bounds__(NW * 32)
    k_poly_apply_volume(const VolApplyArgs A)
    {
      using C             = Cfg<DIM, DEG>;
      constexpr int N1    = C::N1, N = C::N, NU = C::NU;
      constexpr int NC    = DIM + (MASS ? 1 : 0);
      constexpr int R     = TQ * NC;
      constexpr int RS    = ((R + 11) / 16) * 16 + 4;
      constexpr int NPW   = (N + NW - 1) / NW; // DoFs per warp in the integrate phase
      __shared__ double T[DIM * 2 * N1 * TQ];
      __shared__ double WC[R];
      __shared__ double TV[R];
      __shared__ double U[N];
      extern __shared__ double G[]; // [N][RS]

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      for (int item = blockIdx.x; item < A.n_items; item += gridDim.x)
        {
          const int     poly = A.item_poly[item];
          const int64_t q0 = A.item_q0[item], q1 = A.item_q1[item];
          const double *bb = A.bbox + (int64_t)poly * 2 * DIM;
          __syncthreads();
          if (tid < N)
            U[tid] = A.x[(int64_t)A.dof_block[poly] * N + tid];
          double acc[NPW];
#pragma unroll
          for (int k = 0; k < NPW; ++k)
            acc[k] = 0.;
          for (int64_t qt = q0; qt < q1; qt += TQ)
            {
              __syncthreads();
              // ---- 1-D tables: warp d < DIM handles direction d, lane = point
              if (warp < DIM)
                {
                  const int     d  = warp;
                  const int64_t gq = qt + lane;
                  const double  lo = bb[d], hi = bb[DIM + d];
                  const bool    ok = gq < q1;
                  const double  x  = ok ? A.vq_x[(int64_t)d * A.Q + gq] : lo;
                  double        L[N1], dL[N1];
                  lagrange<N1>(A.basis, (x - lo) / (hi - lo), 1. / (hi - lo), L, dL);
#pragma unroll
                  for (int a = 0; a < N1; ++a)
                    {
                      T[(d * 2 * N1 + a) * TQ + lane]      = L[a];
                      T[(d * 2 * N1 + N1 + a) * TQ + lane] = dL[a];
                    }
                  if (d == 0)
                    {
                      const double w = ok ? A.vq_w[gq] : 0.;
#pragma unroll
                      for (int c = 0; c < DIM; ++c)
                        WC[c * TQ + lane] = w * A.stiffness;
                      if (MASS)
                        WC[DIM * TQ + lane] = w * A.mass;
                    }
                }
              __syncthreads();
              // ---- operand rows G[i][c*TQ + q]
              for (int wu = warp; wu < NU; wu += NW)
                {
                  const double *Tq = T + lane;
                  double       *Gq = G + lane;
                  if constexpr (DIM == 2)
                    {
                      const double ly = Tq[(2 * N1 + wu) * TQ], dy = Tq[(3 * N1 + wu) * TQ];
#pragma unroll
                      for (int a = 0; a < N1; ++a)
                        {
                          const double lx = Tq[a * TQ], dx = Tq[(N1 + a) * TQ];
                          double      *o  = Gq + (wu * N1 + a) * RS;
                          o[0]            = dx * ly;
                          o[TQ]           = lx * dy;
                          if (MASS)
                            o[2 * TQ] = lx * ly;
                        }
                    }
                  else
                    {
                      const int    b = wu % N1, c = wu / N1;
                      const double ly = Tq[(2 * N1 + b) * TQ], dy = Tq[(3 * N1 + b) * TQ];
                      const double lz = Tq[(4 * N1 + c) * TQ], dz = Tq[(5 * N1 + c) * TQ];
                      const double yz = ly * lz, dyz = dy * lz, ydz = ly * dz;
#pragma unroll
                      for (int a = 0; a < N1; ++a)
                        {
                          const double lx = Tq[a * TQ], dx = Tq[(N1 + a) * TQ];
                          double      *o  = Gq + (wu * N1 + a) * RS;
                          o[0]            = dx * yz;
                          o[TQ]           = lx * dyz;
                          o[2 * TQ]       = lx * ydz;
                          if (MASS)
                            o[3 * TQ] = lx * yz;
                        }
                    }
                }
              __syncthreads();
              // ---- evaluate: t_r = w_r c_r sum_i G[i][r] u_i
              for (int r = tid; r < R; r += NW * 32)
                {
                  double t = 0.;
#pragma unroll 9
                  for (int i = 0; i < N; ++i)
                    t += G[i * RS + r] * U[i];
                  TV[r] = t * WC[r];
                }
              __syncthreads();
              // ---- integrate: y_i += sum_r G[i][r] t_r ; warp w owns DoFs w, w+NW, ...
#pragma unroll
              for (int k = 0; k < NPW; ++k)
                {
                  const int i = warp + k * NW;
                  if (i < N)
                    {
                      double s = 0.;
                      for (int r = lane; r < R; r += 32)
                        s += G[i * RS + r] * TV[r];
#pragma unroll
                      for (int o = 16; o > 0; o >>= 1)
                        s += __shfl_xor_sync(0xffffffffu, s, o);
                      acc[k] += s;
                    }
                }
            }
          if (lane == 0)
            {
#pragma unroll
              for (int k = 0; k < NPW; ++k)
                if (warp + k * NW < N)
                  A.partial[(int64_t)item * N + warp + k * NW] = acc[k];
            }
        }
    }

    struct FaceApplyArgs
    {
      const double  *fq_x, *fq_n, *fq_w;
      int64_t        Qf;
      int            nqf;
      const double  *bbox;
      const int32_t *ifA, *ifB, *dof_block;
      const int64_t *if_sub_ptr;
      const double  *sub_sigma;
      int32_t        n_ifaces;
      const double  *x;
      double        *partial; // [n_ifaces][2][N]
      double         stiffness;
      uint32_t       flags;
      Basis1D        basis;
    };

    template <int DIM, int DEG>
    __global__ void __launch_bounds__(NW * 32)
    k_poly_apply_faces(const FaceApplyArgs A)
    {
      using C            = Cfg<DIM, DEG>;
      constexpr int N1   = C::N1, N = C::N, NU = C::NU;
      constexpr int N2   = 2 * N;
      constexpr int RS   = 36;
      constexpr int NPW  = (N2 + NW - 1) / NW;
      constexpr int NTAB = 2 * DIM;
      static_assert(NTAB <= NW, "one warp per table task");
      __shared__ double T[2 * DIM * 2 * N1 * TQ];
      __shared__ double WQ[TQ], SG[TQ], TVZ[2][TQ];
      __shared__ double U[N2];
      extern __shared__ double P[]; // Z[N2][RS] then V'[N2][RS]
      double *Zp = P, *Vp = P + N2 * RS;

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      for (int f = blockIdx.x; f < A.n_ifaces; f += gridDim.x)
        {
          const int  pa = A.ifA[f], pb = A.ifB[f];
          const bool interior = pb >= 0;
          double    *out = A.partial + (int64_t)f * 2 * N;
          if (interior ? !(A.flags & PD_ASSEMBLE_INTERIOR) : !(A.flags & PD_ASSEMBLE_BOUNDARY))
            {
              for (int i = tid; i < N2; i += NW * 32)
                out[i] = 0.;
              continue;
            }
          const int64_t q0 = A.if_sub_ptr[f] * A.nqf, q1 = A.if_sub_ptr[f + 1] * A.nqf;
          const int     nside = interior ? 2 : 1, ncol = nside * N;
          const double  dscale = interior ? 0.5 : 1.;
          __syncthreads();
          if (tid < N)
            U[tid] = A.x[(int64_t)A.dof_block[pa] * N + tid];
          else if (tid < N2)
            U[tid] = interior ? A.x[(int64_t)A.dof_block[pb] * N + tid - N] : 0.;
          double acc[NPW];
#pragma unroll
          for (int k = 0; k < NPW; ++k)
            acc[k] = 0.;
          for (int64_t qt = q0; qt < q1; qt += TQ)
            {
              __syncthreads();
              if (warp < NTAB && (interior || warp < DIM))
                {
                  const int     side = warp / DIM, d = warp % DIM;
                  const double *bb   = A.bbox + (int64_t)(side ? pb : pa) * 2 * DIM;
                  const double  lo = bb[d], hi = bb[DIM + d];
                  const int64_t gq = qt + lane;
                  const bool    ok = gq < q1;
                  const double  x  = ok ? A.fq_x[(int64_t)d * A.Qf + gq] : lo;
                  const double  nd = ok ? A.fq_n[(int64_t)d * A.Qf + gq] : 0.;
                  double        L[N1], dL[N1];
                  lagrange<N1>(A.basis, (x - lo) / (hi - lo), nd / (hi - lo), L, dL);
                  double *Tq = T + ((side * DIM + d) * 2 * N1) * TQ + lane;
#pragma unroll
                  for (int a = 0; a < N1; ++a)
                    {
                      Tq[a * TQ]        = L[a];
                      Tq[(N1 + a) * TQ] = dL[a];
                    }
                  if (warp == 0)
                    {
                      WQ[lane] = ok ? A.fq_w[gq] * A.stiffness : 0.;
                      SG[lane] = ok ? 0.5 * A.sub_sigma[gq / A.nqf] : 0.;
                    }
                }
              __syncthreads();
              const double hs = SG[lane], wq = WQ[lane];
              for (int wu = warp; wu < nside * NU; wu += NW)
                {
                  const int     side = wu / NU, bc = wu - side * NU;
                  const double *Tq   = T + (side * DIM * 2 * N1) * TQ + lane;
                  const double  sgn  = side ? -1. : 1.;
                  double        s1, s2;
                  if constexpr (DIM == 2)
                    {
                      s1 = Tq[(2 * N1 + bc) * TQ];
                      s2 = Tq[(3 * N1 + bc) * TQ];
                    }
                  else
                    {
                      const int    b = bc % N1, c = bc / N1;
                      const double ly = Tq[(2 * N1 + b) * TQ], dy = Tq[(3 * N1 + b) * TQ];
                      const double lz = Tq[(4 * N1 + c) * TQ], dz = Tq[(5 * N1 + c) * TQ];
                      s1 = ly * lz;
                      s2 = dy * lz + ly * dz;
                    }
#pragma unroll
                  for (int a = 0; a < N1; ++a)
                    {
                      const double lx = Tq[a * TQ], dx = Tq[(N1 + a) * TQ];
                      const double V  = sgn * lx * s1;
                      const double dn = dx * s1 + lx * s2;
                      const int    col = side * N + bc * N1 + a;
                      Vp[col * RS + lane] = V * wq;
                      Zp[col * RS + lane] = hs * V - dscale * dn;
                    }
                }
              __syncthreads();
              // ---- tv_q = V'_q . u ,  tz_q = Z_q . u
              if (tid < 2 * TQ)
                {
                  const double *M = tid < TQ ? Vp : Zp;
                  double        t = 0.;
                  for (int col = 0; col < ncol; ++col)
                    t += M[col * RS + lane] * U[col];
                  TVZ[tid / TQ][lane] = t;
                }
              __syncthreads();
              // ---- y_col += sum_q Z[col][q] tv_q + V'[col][q] tz_q
#pragma unroll
              for (int k = 0; k < NPW; ++k)
                {
                  const int col = warp + k * NW;
                  if (col < ncol)
                    {
                      double s = Zp[col * RS + lane] * TVZ[0][lane] + Vp[col * RS + lane] * TVZ[1][lane];
#pragma unroll
                      for (int o = 16; o > 0; o >>= 1)
                        s += __shfl_xor_sync(0xffffffffu, s, o);
                      acc[k] += s;
                    }
                }
            }
          if (lane == 0)
            {
#pragma unroll
              for (int k = 0; k < NPW; ++k)
                if (warp + k * NW < N2)
                  out[warp + k * NW] = (warp + k * NW < ncol) ? acc[k] : 0.;
            }
        }
    }

    struct GatherArgs
    {
      const int64_t *poly_vitem_ptr;
      const double  *vol_partial;
      const int64_t *padj_ptr, *padj;
      const double  *face_partial;
      const int32_t *dof_block;
      double        *y;
      int32_t        np_own, n;
      uint32_t       flags;
      int            add;
    };

    __global__ void __launch_bounds__(256)
    k_poly_gather(const GatherArgs A)
    {
      const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (int64_t)A.np_own * A.n)
        return;
      const int p = (int)(idx / A.n), i = (int)(idx - (int64_t)p * A.n);
      double    s = 0.;
      if (A.flags & PD_ASSEMBLE_VOLUME)
        for (int64_t it = A.poly_vitem_ptr[p]; it < A.poly_vitem_ptr[p + 1]; ++it)
          s += A.vol_partial[it * A.n + i];
      for (int64_t k = A.padj_ptr[p]; k < A.padj_ptr[p + 1]; ++k)
        {
          const int64_t e = A.padj[k];
          s += A.face_partial[((e >> 1) * 2 + (e & 1)) * A.n + i];
        }
      double *yp = A.y + (int64_t)A.dof_block[p] * A.n + i;
      *yp        = A.add ? *yp + s : s;
    }

    template <int DIM, int DEG>
    void
    run(pd_handle *h, const double *src, double *dst, const bool add)
    {
      using C         = Cfg<DIM, DEG>;
      const bool mass = h->op_coef.mass != 0.;
      // volume
      // reuse the assembly's volume schedule when there is one (items are just independent runs
      // of stages here), else build one with a few items per SM
      if (h->vol_plan_tq == 0)
        plan_volume_items(h, TQ, (int)std::min<int64_t>(std::max<int64_t>(1, (h->Q + TQ - 1) / TQ / 16 + 1),
                                                        (int64_t)h->sm_count * 4));
      if (h->mf_vol_partial.n != (size_t)h->n_vitems * h->n)
        h->mf_vol_partial.alloc((size_t)h->n_vitems * h->n);
      if (h->mf_face_partial.n != (size_t)h->n_ifaces * 2 * h->n)
        h->mf_face_partial.alloc((size_t)h->n_ifaces * 2 * h->n);
      if ((h->op_flags & PD_ASSEMBLE_VOLUME) && h->n_vitems > 0)
        {
          VolApplyArgs a;
          a.vq_x      = h->vq_x.p;
          a.vq_w      = h->vq_w.p;
          a.Q         = h->Q;
          a.bbox      = h->bbox.p;
          a.dof_block = h->dof_block.p;
          a.item_poly = h->vitem_poly.p;
          a.item_q0   = h->vitem_q0.p;
          a.item_q1   = h->vitem_q1.p;
          a.n_items   = h->n_vitems;
          a.x         = src;
          a.partial   = h->mf_vol_partial.p;
          a.stiffness = h->op_coef.stiffness;
          a.mass      = h->op_coef.mass;
          a.basis     = h->basis;
          const int    nc   = DIM + (mass ? 1 : 0);
          const int    rs   = ((TQ * nc + 11) / 16) * 16 + 4;
          const size_t smem = sizeof(double) * C::N * rs;
          const int    grid = (int)std::min<int64_t>(h->n_vitems, (int64_t)h->sm_count * 4);
          if (mass)
            {
              if (smem > 40 * 1024)
                PD_CUDA(cudaFuncSetAttribute(k_poly_apply_volume<DIM, DEG, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              k_poly_apply_volume<DIM, DEG, true><<<grid, NW * 32, smem, h->stream>>>(a);
            }
          else
            {
              if (smem > 40 * 1024)
                PD_CUDA(cudaFuncSetAttribute(k_poly_apply_volume<DIM, DEG, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              k_poly_apply_volume<DIM, DEG, false><<<grid, NW * 32, smem, h->stream>>>(a);
            }
          ++h->launches;
        }
      if (h->n_ifaces > 0)
        {
          FaceApplyArgs a;
          a.fq_x       = h->fq_x.p;
          a.fq_n       = h->fq_n.p;
          a.fq_w       = h->fq_w.p;
          a.Qf         = h->Qf;
          a.nqf        = h->nqf;
          a.bbox       = h->bbox.p;
          a.ifA        = h->ifA.p;
          a.ifB        = h->ifB.p;
          a.dof_block  = h->dof_block.p;
          a.if_sub_ptr = h->if_sub_ptr.p;
          a.sub_sigma  = h->sub_sigma.p;
          a.n_ifaces   = h->n_ifaces;
          a.x          = src;
          a.partial    = h->mf_face_partial.p;
          a.stiffness  = h->op_coef.stiffness;
          a.flags      = h->op_flags;
          a.basis      = h->basis;
          const size_t smem = sizeof(double) * 2 * (2 * C::N) * 36;
          if (smem > 40 * 1024)
            PD_CUDA(cudaFuncSetAttribute(k_poly_apply_faces<DIM, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
          const int grid = (int)std::min<int64_t>(h->n_ifaces, (int64_t)h->sm_count * 4);
          k_poly_apply_faces<DIM, DEG><<<grid, NW * 32, smem, h->stream>>>(a);
          ++h->launches;
        }
      GatherArgs g;
      g.poly_vitem_ptr = h->poly_vitem_ptr.p;
      g.vol_partial    = h->mf_vol_partial.p;
      g.padj_ptr       = h->padj_ptr.p;
      g.padj           = h->padj.p;
      g.face_partial   = h->mf_face_partial.p;
      g.dof_block      = h->dof_block.p;
      g.y              = dst;
      g.np_own         = h->np_own;
      g.n              = h->n;
      g.flags          = h->op_flags;
      g.add            = add ? 1 : 0;
      const int64_t nd = (int64_t)h->np_own * h->n;
      k_poly_gather<<<(unsigned)((nd + 255) / 256), 256, 0, h->stream>>>(g);
      ++h->launches;
      PD_CUDA(cudaGetLastError());
    }
  } // namespace

  void
  launch_poly_apply(pd_handle *h, const double *src, double *dst, const bool add)
  {
    const int key = h->dim * 10 + h->degree;
    switch (key)
      {
        case 21: run<2, 1>(h, src, dst, add); break;
        case 22: run<2, 2>(h, src, dst, add); break;
        case 23: run<2, 3>(h, src, dst, add); break;
        case 24: run<2, 4>(h, src, dst, add); break;
        case 31: run<3, 1>(h, src, dst, add); break;
        case 32: run<3, 2>(h, src, dst, add); break;
        case 33: run<3, 3>(h, src, dst, add); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no matrix-free polytopal kernel for this (dim, degree)", __LINE__};
      }
  }
} // namespace pd
