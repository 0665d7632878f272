// -----------------------------------------------------------------------------
// pd_assemble.cu -- SIP-DG assembly over agglomerated polytopes on sm_100a.
//
// Reference semantics: PolyUtils::assemble_dg_matrix
// (include/poly_utils.h:2000-2195) with the jump/average kernel of
// :1870-1926, evaluated through AgglomerationHandler::reinit /
// reinit(polytope,f) / reinit_interface (source/agglomeration_handler.cc:
// 729-906) and MappingBox (source/mapping_box.cc:393-439,465-532).
//
// The reference builds, per polytope and per face, heap-allocated tables of all
// basis values/gradients at all quadrature points and runs scalar q*i*j loops.
// Here every term is a rank-k update  C += sum_r a_r b_r^T  whose operand rows
// are generated on the fly in shared memory from the (x, JxW[, n]) streams and
// contracted with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 -> DMMA.8x8x4,
// which profiles/FP64_PEAK.json shows to reach the full 37.1 TFLOP/s of the
// FP64 pipe, plain DFMA 34.1):
//
//   volume   K   = sum_{q,d} w_q g_d(q) g_d(q)^T (+ f w_q phi phi^T)
//            symmetric: only tiles on/above the diagonal are computed.
//   faces    with V = [phi0 ; -phi1], D = [dn phi0 ; dn phi1]/2, Z = sigma/2 V - D:
//            [M11 M12; M21 M22] = T + T^T,  T = sum_q w_q Z_q V_q^T
//            (algebraically identical to the four formulas of poly_utils.h:
//            1891-1922; boundary: Z = sigma/2 phi - dn phi, one side only).
//
// Determinism / no atomics: every off-diagonal block (A,B) is written by the
// one interface {A,B}; the diagonal block of P is gathered by k_reduce_diag
// from P's volume partials and the M11/M22 parts of P's faces in a fixed order.
//
// Result layout: scalar CSR values of the reference sparsity pattern (ascending
// columns): block row b, local row i, block k, local column j at
//   brow_ptr[b]*n*n + i*(nb_b*n) + k*n + j.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

namespace pd
{
  namespace
  {
    constexpr int
    ipow(const int b, const int e)
    {
      return e == 0 ? 1 : b * ipow(b, e - 1);
    }

    template <int DIM, int DEG>
    struct Cfg
    {
      static constexpr int N1  = DEG + 1;
      static constexpr int N   = ipow(N1, DIM);
      static constexpr int NT8 = (N + 7) / 8; // 8x8 MMA tiles per side
      static constexpr int NP  = NT8 * 8;
      // Row stride (doubles) of operand panels.  A fragment load touches 4 rows x
      // 8 consecutive doubles; with stride = 4, 8 or 12 (mod 16) the four rows
      // cover every bank pair exactly twice => 2 wavefronts, the minimum for 256 B.
      static constexpr int STRIDE = (NP % 16 == 0) ? NP + 8 : NP;
    };

    __device__ __forceinline__ void
    dmma884(double &c0, double &c1, const double a, const double b)
    {
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
          : "+d"(c0), "+d"(c1)
          : "d"(a), "d"(b));
    }

    __device__ __forceinline__ void
    cta_sync()
    {
      // plain barrier 0; issued from role-specialised code paths, hence PTX
      asm volatile("bar.sync 0;\n" ::: "memory");
    }

    // l_a(x), l_a'(x) * scale for a < N1
    template <int N1>
    __device__ __forceinline__ void
    lagrange(const Basis1D &B, const double x, const double scale, double *L, double *dL)
    {
#pragma unroll
      for (int a = 0; a < N1; ++a)
        {
          double val = 1., der = 0.;
#pragma unroll
          for (int b = 0; b < N1; ++b)
            if (b != a)
              {
                const double t = x - B.node[b];
                der            = der * t + val;
                val            = val * t;
              }
          L[a]  = val * B.wprod[a];
          dL[a] = der * B.wprod[a] * scale;
        }
    }

    // -------------------------------------------------------------------------
    // volume kernel
    // -------------------------------------------------------------------------
    struct VolArgs
    {
      const double  *vq_x;
      const double  *vq_w;
      int64_t        Q; // stride between coordinate streams
      const double  *bbox;
      const int32_t *item_poly;
      const int64_t *item_q0, *item_q1;
      int32_t        n_items;
      double        *partial; // [n_items][N][N]
      double         stiffness, mass;
      Basis1D        basis;
    };

    template <int NT8, int NROLE, int ROLE>
    __device__ __forceinline__ constexpr bool
    vol_role_owns_row(const int I)
    {
      // NROLE == 2: rows {0,3,4,7} vs {1,2,5,6}: 18 upper-triangle tiles each for NT8 == 8
      return NROLE == 1 ? true : ((((I & 3) == 0) || ((I & 3) == 3)) == (ROLE == 0));
    }

    template <int DIM, int DEG, bool MASS, int TQ, int NWARPS, int NROLE, int ROLE>
    __device__ __forceinline__ void
    volume_body(const VolArgs &A, double *smem)
    {
      using C              = Cfg<DIM, DEG>;
      constexpr int N1     = C::N1, N = C::N, NT8 = C::NT8, NP = C::NP, STRIDE = C::STRIDE;
      constexpr int NC     = DIM + (MASS ? 1 : 0);
      constexpr int R      = TQ * NC;
      constexpr int KSPLIT = NWARPS / NROLE;
      constexpr int NTHR   = NWARPS * 32;
      static_assert(R % 4 == 0, "panel rows must be a multiple of the MMA k");
      static_assert(TQ * DIM <= NTHR, "one thread per (point, direction) in the table phase");

      double *G  = smem;                        // [R][STRIDE] operand panel
      double *WC = G + R * STRIDE;              // [R] JxW * coefficient of the row
      double *T  = WC + R;                      // [TQ][DIM][2][N1] 1-D tables
      double *S  = smem;                        // [NP][NP+1] aliases G after the main loop
      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int g = lane >> 2, t = lane & 3;
      const int kpart = warp / NROLE;

      // columns >= N of the panel are never written by the generator: clear once
      for (int i = tid; i < R * STRIDE; i += NTHR)
        G[i] = 0.;

      for (int item = blockIdx.x; item < A.n_items; item += gridDim.x)
        {
          const int      poly = A.item_poly[item];
          const int64_t  q0 = A.item_q0[item], q1 = A.item_q1[item];
          const double  *bb = A.bbox + (int64_t)poly * 2 * DIM;
          double         acc[NT8][NT8][2];
#pragma unroll
          for (int I = 0; I < NT8; ++I)
#pragma unroll
            for (int J = 0; J < NT8; ++J)
              acc[I][J][0] = acc[I][J][1] = 0.;

          for (int64_t qt = q0; qt < q1; qt += TQ)
            {
              cta_sync(); // previous MMA phase done with G / T / WC
              // ---- phase 1: 1-D tables, one thread per (direction, point)
              if (tid < TQ * DIM)
                {
                  const int     d = tid / TQ, q = tid - d * TQ;
                  const int64_t gq = qt + q;
                  const double  lo = bb[d], hi = bb[DIM + d];
                  const double  x  = gq < q1 ? A.vq_x[(int64_t)d * A.Q + gq] : lo;
                  // BoundingBox::real_to_unit, covariant scaling by 1/h
                  const double xhat = (x - lo) / (hi - lo);
                  double       L[N1], dL[N1];
                  lagrange<N1>(A.basis, xhat, 1. / (hi - lo), L, dL);
                  double *Tq = T + (q * DIM + d) * 2 * N1;
#pragma unroll
                  for (int a = 0; a < N1; ++a)
                    {
                      Tq[a]      = L[a];
                      Tq[N1 + a] = dL[a];
                    }
                  if (d == 0)
                    {
                      const double w = gq < q1 ? A.vq_w[gq] : 0.;
#pragma unroll
                      for (int c = 0; c < DIM; ++c)
                        WC[q * NC + c] = w * A.stiffness;
                      if (MASS)
                        WC[q * NC + DIM] = w * A.mass;
                    }
                }
              cta_sync();
              // ---- phase 2: rows g_d(q)[i] (and phi(q)[i]) of the operand panel
              for (int idx = tid; idx < TQ * N; idx += NTHR)
                {
                  const int     q  = idx / N, i = idx - q * N;
                  const double *Tq = T + q * DIM * 2 * N1;
                  double       *Gq = G + (q * NC) * STRIDE + i;
                  if constexpr (DIM == 2)
                    {
                      const int    a = i % N1, b = i / N1;
                      const double lx = Tq[a], dx = Tq[N1 + a], ly = Tq[2 * N1 + b], dy = Tq[3 * N1 + b];
                      Gq[0]          = dx * ly;
                      Gq[STRIDE]     = lx * dy;
                      if (MASS)
                        Gq[2 * STRIDE] = lx * ly;
                    }
                  else
                    {
                      const int    a = i % N1, b = (i / N1) % N1, c = i / (N1 * N1);
                      const double lx = Tq[a], dx = Tq[N1 + a], ly = Tq[2 * N1 + b], dy = Tq[3 * N1 + b],
                                   lz = Tq[4 * N1 + c], dz = Tq[5 * N1 + c];
                      const double lxy = lx * ly;
                      Gq[0]            = dx * ly * lz;
                      Gq[STRIDE]       = lx * dy * lz;
                      Gq[2 * STRIDE]   = lxy * dz;
                      if (MASS)
                        Gq[3 * STRIDE] = lxy * lz;
                    }
                }
              cta_sync();
              // ---- phase 3: C += sum_r wc_r G_r G_r^T on the upper triangle of tiles
#pragma unroll 2
              for (int ks = kpart; ks < R / 4; ks += KSPLIT)
                {
                  const int     r   = ks * 4 + t;
                  const double *row = G + r * STRIDE + g;
                  const double  wc  = WC[r];
                  double        a[NT8], b[NT8];
#pragma unroll
                  for (int I = 0; I < NT8; ++I)
                    {
                      a[I] = row[8 * I];
                      b[I] = a[I] * wc;
                    }
#pragma unroll
                  for (int I = 0; I < NT8; ++I)
                    if (vol_role_owns_row<NT8, NROLE, ROLE>(I))
                      {
#pragma unroll
                        for (int J = I; J < NT8; ++J)
                          dmma884(acc[I][J][0], acc[I][J][1], a[I], b[J]);
                      }
                }
            }
          // ---- reduce the k-split partial tiles through shared memory
          constexpr int SS = NP + 1;
          cta_sync();
          for (int round = 0; round < KSPLIT; ++round)
            {
              if (kpart == round)
                {
#pragma unroll
                  for (int I = 0; I < NT8; ++I)
                    if (vol_role_owns_row<NT8, NROLE, ROLE>(I))
                      {
#pragma unroll
                        for (int J = I; J < NT8; ++J)
                          {
                            double *s = S + (8 * I + g) * SS + 8 * J + 2 * t;
                            if (round == 0)
                              {
                                s[0] = acc[I][J][0];
                                s[1] = acc[I][J][1];
                              }
                            else
                              {
                                s[0] += acc[I][J][0];
                                s[1] += acc[I][J][1];
                              }
                          }
                      }
                }
              cta_sync();
            }
          double *out = A.partial + (int64_t)item * N * N;
          for (int idx = tid; idx < N * N; idx += NTHR)
            {
              const int i = idx / N, j = idx - i * N;
              out[idx]    = (i >> 3) <= (j >> 3) ? S[i * SS + j] : S[j * SS + i];
            }
          cta_sync();
          // S aliased the panel: restore the zero padding columns
          for (int i = tid; i < R * STRIDE; i += NTHR)
            G[i] = 0.;
        }
    }

    template <int DIM, int DEG, bool MASS, int TQ, int NWARPS>
    __global__ void __launch_bounds__(NWARPS * 32, 2)
    k_volume(const VolArgs A)
    {
      extern __shared__ double smem[];
      constexpr int NROLE = Cfg<DIM, DEG>::NT8 >= 8 ? 2 : 1;
      if constexpr (NROLE == 1)
        volume_body<DIM, DEG, MASS, TQ, NWARPS, 1, 0>(A, smem);
      else
        {
          if (((threadIdx.x >> 5) & 1) == 0)
            volume_body<DIM, DEG, MASS, TQ, NWARPS, 2, 0>(A, smem);
          else
            volume_body<DIM, DEG, MASS, TQ, NWARPS, 2, 1>(A, smem);
        }
    }

    template <int DIM, int DEG, bool MASS, int TQ>
    constexpr size_t
    volume_smem_bytes()
    {
      using C          = Cfg<DIM, DEG>;
      constexpr int NC = DIM + (MASS ? 1 : 0);
      constexpr int R  = TQ * NC;
      size_t        a  = (size_t)R * C::STRIDE + R + (size_t)TQ * DIM * 2 * C::N1;
      size_t        s  = (size_t)C::NP * (C::NP + 1);
      return sizeof(double) * (a > s ? a : s);
    }

    // -------------------------------------------------------------------------
    // face kernel
    // -------------------------------------------------------------------------
    struct FaceArgs
    {
      const double  *fq_x, *fq_n, *fq_w;
      int64_t        Qf;
      int            nqf;
      const double  *bbox;
      const int32_t *ifA, *ifB;
      const int64_t *if_sub_ptr;
      const double  *sub_sigma;
      const int64_t *if_baseAB, *if_baseBA;
      const int32_t *dof_block, *row_stride;
      int32_t        n_ifaces;
      double        *face_diag; // [n_ifaces][2][N][N]
      double        *values;
      double         stiffness;
      uint32_t       flags;
      Basis1D        basis;
    };

    template <int DIM, int DEG, int TQ, int ISPLIT, int KSPLIT>
    __global__ void __launch_bounds__(4 * ISPLIT * KSPLIT * 32, 1)
    k_faces(const FaceArgs A)
    {
      using C              = Cfg<DIM, DEG>;
      constexpr int N1     = C::N1, N = C::N, NT8 = C::NT8, NP = C::NP;
      constexpr int NP2    = 2 * NP;
      constexpr int STRIDE = (NP2 % 16 == 0) ? NP2 + 8 : NP2;
      constexpr int NWARPS = 4 * ISPLIT * KSPLIT;
      constexpr int NTHR   = NWARPS * 32;
      constexpr int MI     = NT8 / ISPLIT;
      constexpr int SS     = NP2 + 1;
      static_assert(NT8 % ISPLIT == 0, "row split must divide the tile count");
      static_assert(TQ % 4 == 0, "panel rows must be a multiple of the MMA k");
      static_assert(2 * TQ * DIM <= NTHR, "one thread per (side, direction, point)");

      extern __shared__ double smem[];
      double *Zp = smem;                  // [TQ][STRIDE]  Z rows (side 0 | side 1)
      double *Vp = Zp + TQ * STRIDE;      // [TQ][STRIDE]  V rows
      double *WC = Vp + TQ * STRIDE;      // [TQ] JxW * sigma_coefficient
      double *SG = WC + TQ;               // [TQ] penalty per point
      double *T  = SG + TQ;               // [TQ][2 sides][DIM][2][N1]
      double *S  = smem;                  // [NP2][SS] aliases the panels in the epilogue

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int g = lane >> 2, t = lane & 3;
      const int quad = warp & 3, qa = quad >> 1, qb = quad & 1;
      const int isub  = (warp >> 2) % ISPLIT;
      const int kpart = (warp >> 2) / ISPLIT;

      for (int i = tid; i < 2 * TQ * STRIDE; i += NTHR)
        smem[i] = 0.;

      for (int f = blockIdx.x; f < A.n_ifaces; f += gridDim.x)
        {
          const int  pa = A.ifA[f], pb = A.ifB[f];
          const bool interior = pb >= 0;
          if (interior ? !(A.flags & PD_ASSEMBLE_INTERIOR) : !(A.flags & PD_ASSEMBLE_BOUNDARY))
            continue;
          const int64_t s0 = A.if_sub_ptr[f], s1 = A.if_sub_ptr[f + 1];
          const int64_t q0 = s0 * A.nqf, q1 = s1 * A.nqf;
          const double *bba = A.bbox + (int64_t)pa * 2 * DIM;
          const double *bbb = A.bbox + (int64_t)(interior ? pb : pa) * 2 * DIM;
          const bool    active = interior || quad == 0; // boundary: only the (0,0) quadrant

          double acc[MI][NT8][2];
#pragma unroll
          for (int I = 0; I < MI; ++I)
#pragma unroll
            for (int J = 0; J < NT8; ++J)
              acc[I][J][0] = acc[I][J][1] = 0.;

          for (int64_t qt = q0; qt < q1; qt += TQ)
            {
              __syncthreads();
              // ---- phase 1: per (side, direction, point): l_a, l_a' * n_d / h_d
              if (tid < 2 * TQ * DIM)
                {
                  const int     side = tid / (TQ * DIM);
                  const int     rem  = tid - side * TQ * DIM;
                  const int     d = rem / TQ, q = rem - d * TQ;
                  const int64_t gq = qt + q;
                  const double *bb = side ? bbb : bba;
                  const double  lo = bb[d], hi = bb[DIM + d];
                  const bool    ok = gq < q1;
                  const double  x  = ok ? A.fq_x[(int64_t)d * A.Qf + gq] : lo;
                  const double  nd = ok ? A.fq_n[(int64_t)d * A.Qf + gq] : 0.;
                  const double  xhat = (x - lo) / (hi - lo);
                  double        L[N1], dL[N1];
                  lagrange<N1>(A.basis, xhat, nd * (1. / (hi - lo)), L, dL);
                  double *Tq = T + ((q * 2 + side) * DIM + d) * 2 * N1;
#pragma unroll
                  for (int a = 0; a < N1; ++a)
                    {
                      Tq[a]      = L[a];
                      Tq[N1 + a] = dL[a];
                    }
                  if (side == 0 && d == 0)
                    {
                      WC[q] = ok ? A.fq_w[gq] * A.stiffness : 0.;
                      SG[q] = ok ? A.sub_sigma[gq / A.nqf] : 0.;
                    }
                }
              __syncthreads();
              // ---- phase 2: V and Z rows.  interior: V = [phi0 ; -phi1],
              //      Z = sigma/2 V - [dn0 ; dn1]/2.  boundary: V = phi0, Z = sigma/2 phi0 - dn0.
              for (int idx = tid; idx < TQ * 2 * N; idx += NTHR)
                {
                  const int q = idx / (2 * N), rem = idx - q * 2 * N;
                  const int side = rem / N, i = rem - side * N;
                  if (side == 1 && !interior)
                    continue;
                  const double *Tq = T + (q * 2 + side) * DIM * 2 * N1;
                  double        v, dn;
                  if constexpr (DIM == 2)
                    {
                      const int a = i % N1, b = i / N1;
                      v  = Tq[a] * Tq[2 * N1 + b];
                      dn = Tq[N1 + a] * Tq[2 * N1 + b] + Tq[a] * Tq[3 * N1 + b];
                    }
                  else
                    {
                      const int    a = i % N1, b = (i / N1) % N1, c = i / (N1 * N1);
                      const double lx = Tq[a], ly = Tq[2 * N1 + b], lz = Tq[4 * N1 + c];
                      v  = lx * ly * lz;
                      dn = Tq[N1 + a] * ly * lz + lx * (Tq[3 * N1 + b] * lz + ly * Tq[5 * N1 + c]);
                    }
                  const double hs = 0.5 * SG[q];
                  const int    col = side * NP + i;
                  if (interior)
                    {
                      const double V = side ? -v : v;
                      Vp[q * STRIDE + col] = V;
                      Zp[q * STRIDE + col] = hs * V - 0.5 * dn;
                    }
                  else
                    {
                      Vp[q * STRIDE + col] = v;
                      Zp[q * STRIDE + col] = hs * v - dn;
                    }
                }
              __syncthreads();
              // ---- phase 3: T_(qa,qb) += sum_q wc Z_qa V_qb^T
              if (active)
                {
                  // boundary faces: the four quadrant warps of an (isub,kpart) group would be
                  // idle except quadrant 0; acceptable, boundary faces are a surface term
                  for (int ks = kpart; ks < TQ / 4; ks += KSPLIT)
                    {
                      const int     r  = ks * 4 + t;
                      const double  wc = WC[r];
                      const double *zr = Zp + r * STRIDE + qa * NP + isub * MI * 8 + g;
                      const double *vr = Vp + r * STRIDE + qb * NP + g;
                      double        a[MI], b[NT8];
#pragma unroll
                      for (int I = 0; I < MI; ++I)
                        a[I] = zr[8 * I];
#pragma unroll
                      for (int J = 0; J < NT8; ++J)
                        b[J] = vr[8 * J] * wc;
#pragma unroll
                      for (int I = 0; I < MI; ++I)
#pragma unroll
                        for (int J = 0; J < NT8; ++J)
                          dmma884(acc[I][J][0], acc[I][J][1], a[I], b[J]);
                    }
                }
            }
          // ---- epilogue: T -> shared, M = T + T^T -> global
          __syncthreads();
          for (int round = 0; round < KSPLIT; ++round)
            {
              if (kpart == round && active)
                {
#pragma unroll
                  for (int I = 0; I < MI; ++I)
#pragma unroll
                    for (int J = 0; J < NT8; ++J)
                      {
                        double *s = S + (qa * NP + (isub * MI + I) * 8 + g) * SS + qb * NP + 8 * J + 2 * t;
                        if (round == 0)
                          {
                            s[0] = acc[I][J][0];
                            s[1] = acc[I][J][1];
                          }
                        else
                          {
                            s[0] += acc[I][J][0];
                            s[1] += acc[I][J][1];
                          }
                      }
                }
              __syncthreads();
            }
          double *fd = A.face_diag + (int64_t)f * 2 * N * N;
          if (interior)
            {
              const int64_t baseAB = A.if_baseAB[f], baseBA = A.if_baseBA[f];
              const int     strideA = A.row_stride[A.dof_block[pa]], strideB = A.row_stride[A.dof_block[pb]];
              for (int idx = tid; idx < 4 * N * N; idx += NTHR)
                {
                  const int which = idx / (N * N), rem = idx - which * N * N;
                  const int i = rem / N, j = rem - i * N;
                  const int ri = (which >> 1) * NP + i, cj = (which & 1) * NP + j;
                  const double m = S[ri * SS + cj] + S[cj * SS + ri];
                  if (which == 0)
                    fd[rem] = m;
                  else if (which == 3)
                    fd[N * N + rem] = m;
                  else if (which == 1)
                    A.values[baseAB + (int64_t)i * strideA + j] = m;
                  else
                    A.values[baseBA + (int64_t)i * strideB + j] = m;
                }
            }
          else
            {
              for (int idx = tid; idx < N * N; idx += NTHR)
                {
                  const int i = idx / N, j = idx - i * N;
                  fd[idx]     = S[i * SS + j] + S[j * SS + i];
                }
            }
          __syncthreads();
          for (int i = tid; i < 2 * TQ * STRIDE; i += NTHR)
            smem[i] = 0.;
        }
    }

    template <int DIM, int DEG, int TQ>
    constexpr size_t
    face_smem_bytes()
    {
      using C              = Cfg<DIM, DEG>;
      constexpr int NP2    = 2 * C::NP;
      constexpr int STRIDE = (NP2 % 16 == 0) ? NP2 + 8 : NP2;
      size_t        a      = (size_t)2 * TQ * STRIDE + 2 * TQ + (size_t)TQ * 2 * DIM * 2 * C::N1;
      size_t        s      = (size_t)NP2 * (NP2 + 1);
      return sizeof(double) * (a > s ? a : s);
    }

    // -------------------------------------------------------------------------
    // diagonal gather: block (P,P) = sum of P's volume partials + the M11 / M22 /
    // boundary parts of P's faces, in the fixed order of the adjacency list
    // -------------------------------------------------------------------------
    struct ReduceArgs
    {
      const int64_t *poly_vitem_ptr;
      const double  *partial;
      const int64_t *padj_ptr, *padj; // entries: iface*2 + side
      const int32_t *ifB;
      const double  *face_diag;
      const int64_t *diag_base;
      const int32_t *dof_block, *row_stride;
      double        *values;
      int32_t        np, n;
      uint32_t       flags;
    };

    __global__ void __launch_bounds__(256)
    k_reduce_diag(const ReduceArgs A)
    {
      const int nn = A.n * A.n;
      for (int p = blockIdx.x; p < A.np; p += gridDim.x)
        {
          const int64_t base   = A.diag_base[p];
          const int     stride = A.row_stride[A.dof_block[p]];
          for (int idx = threadIdx.x; idx < nn; idx += blockDim.x)
            {
              double s = 0.;
              if (A.flags & PD_ASSEMBLE_VOLUME)
                for (int64_t it = A.poly_vitem_ptr[p]; it < A.poly_vitem_ptr[p + 1]; ++it)
                  s += A.partial[it * nn + idx];
              for (int64_t k = A.padj_ptr[p]; k < A.padj_ptr[p + 1]; ++k)
                {
                  const int64_t e  = A.padj[k];
                  const int64_t f  = e >> 1;
                  const bool    in = A.ifB[f] >= 0;
                  if (in ? (A.flags & PD_ASSEMBLE_INTERIOR) : (A.flags & PD_ASSEMBLE_BOUNDARY))
                    s += A.face_diag[(f * 2 + (e & 1)) * nn + idx];
                }
              const int i = idx / A.n, j = idx - i * A.n;
              A.values[base + (int64_t)i * stride + j] = s;
            }
        }
    }

    // -------------------------------------------------------------------------
    // dispatch
    // -------------------------------------------------------------------------
    template <class K>
    void
    set_smem(K kernel, const size_t bytes)
    {
      if (bytes > 48 * 1024)
        PD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }

    template <int DIM, int DEG, bool MASS, int TQ>
    void
    run_volume(pd_handle *h, const pd_coefficients &coef)
    {
      constexpr int NWARPS = 8;
      VolArgs       a;
      a.vq_x      = h->vq_x.p;
      a.vq_w      = h->vq_w.p;
      a.Q         = h->Q;
      a.bbox      = h->bbox.p;
      a.item_poly = h->vitem_poly.p;
      a.item_q0   = h->vitem_q0.p;
      a.item_q1   = h->vitem_q1.p;
      a.n_items   = h->n_vitems;
      a.partial   = h->vol_partial.p;
      a.stiffness = coef.stiffness;
      a.mass      = coef.mass;
      a.basis     = h->basis;
      auto           kern  = k_volume<DIM, DEG, MASS, TQ, NWARPS>;
      constexpr size_t smem = volume_smem_bytes<DIM, DEG, MASS, TQ>();
      set_smem(kern, smem);
      int per_sm = 1;
      PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NWARPS * 32, smem));
      if (per_sm < 1)
        per_sm = 1;
      const int grid = std::min<int64_t>(h->n_vitems, (int64_t)h->sm_count * per_sm);
      kern<<<grid, NWARPS * 32, smem, h->stream>>>(a);
      ++h->launches;
    }

    template <int DIM, int DEG, int TQ, int ISPLIT, int KSPLIT>
    void
    run_faces(pd_handle *h, const pd_coefficients &coef, const uint32_t flags)
    {
      FaceArgs a;
      a.fq_x       = h->fq_x.p;
      a.fq_n       = h->fq_n.p;
      a.fq_w       = h->fq_w.p;
      a.Qf         = h->Qf;
      a.nqf        = h->nqf;
      a.bbox       = h->bbox.p;
      a.ifA        = h->ifA.p;
      a.ifB        = h->ifB.p;
      a.if_sub_ptr = h->if_sub_ptr.p;
      a.sub_sigma  = h->sub_sigma.p;
      a.if_baseAB  = h->if_baseAB.p;
      a.if_baseBA  = h->if_baseBA.p;
      a.dof_block  = h->dof_block.p;
      a.row_stride = h->row_stride.p;
      a.n_ifaces   = h->n_ifaces;
      a.face_diag  = h->face_diag.p;
      a.values     = h->values.p;
      a.stiffness  = coef.stiffness;
      a.flags      = flags;
      a.basis      = h->basis;
      auto             kern  = k_faces<DIM, DEG, TQ, ISPLIT, KSPLIT>;
      constexpr size_t smem  = face_smem_bytes<DIM, DEG, TQ>();
      constexpr int    nthr  = 4 * ISPLIT * KSPLIT * 32;
      set_smem(kern, smem);
      int per_sm = 1;
      PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthr, smem));
      if (per_sm < 1)
        per_sm = 1;
      const int grid = std::min<int64_t>(h->n_ifaces, (int64_t)h->sm_count * per_sm);
      kern<<<grid, nthr, smem, h->stream>>>(a);
      ++h->launches;
    }

    template <int DIM, int DEG, int TQV, int TQF, int ISPLIT, int KSPLIT>
    void
    run_all(pd_handle *h, const uint32_t flags, const pd_coefficients &coef)
    {
      PD_CUDA(cudaEventRecord(h->ev[0], h->stream));
      if ((flags & PD_ASSEMBLE_VOLUME) && h->n_vitems > 0)
        {
          if (coef.mass != 0.)
            run_volume<DIM, DEG, true, TQV>(h, coef);
          else
            run_volume<DIM, DEG, false, TQV>(h, coef);
        }
      PD_CUDA(cudaEventRecord(h->ev[1], h->stream));
      if ((flags & (PD_ASSEMBLE_BOUNDARY | PD_ASSEMBLE_INTERIOR)) && h->n_ifaces > 0)
        run_faces<DIM, DEG, TQF, ISPLIT, KSPLIT>(h, coef, flags);
      PD_CUDA(cudaEventRecord(h->ev[2], h->stream));
      ReduceArgs r;
      r.poly_vitem_ptr = h->poly_vitem_ptr.p;
      r.partial        = h->vol_partial.p;
      r.padj_ptr       = h->padj_ptr.p;
      r.padj           = h->padj.p;
      r.ifB            = h->ifB.p;
      r.face_diag      = h->face_diag.p;
      r.diag_base      = h->diag_base.p;
      r.dof_block      = h->dof_block.p;
      r.row_stride     = h->row_stride.p;
      r.values         = h->values.p;
      r.np             = h->np;
      r.n              = h->n;
      r.flags          = flags;
      const int grid   = std::min<int64_t>(h->np, (int64_t)h->sm_count * 8);
      k_reduce_diag<<<grid, 256, 0, h->stream>>>(r);
      ++h->launches;
      PD_CUDA(cudaEventRecord(h->ev[3], h->stream));
      PD_CUDA(cudaGetLastError());
    }
  } // namespace

  bool
  assemble_supported(const int dim, const int degree)
  {
    if (dim == 2)
      return degree >= 1 && degree <= 4;
    if (dim == 3)
      return degree >= 1 && degree <= 3;
    return false;
  }

  void
  launch_assemble(pd_handle *h, const uint32_t flags, const pd_coefficients &coef)
  {
    // off-diagonal blocks are only written by interior interfaces: clear when skipped
    if (!(flags & PD_ASSEMBLE_INTERIOR))
      PD_CUDA(cudaMemsetAsync(h->values.p, 0, sizeof(double) * h->nnz, h->stream));
    const int key = h->dim * 10 + h->degree;
    switch (key)
      {
        //                 DIM DEG TQV TQF ISPLIT KSPLIT
        case 21: run_all<2, 1, 64, 32, 1, 2>(h, flags, coef); break;
        case 22: run_all<2, 2, 64, 32, 1, 2>(h, flags, coef); break;
        case 23: run_all<2, 3, 64, 32, 1, 2>(h, flags, coef); break;
        case 24: run_all<2, 4, 64, 32, 1, 2>(h, flags, coef); break;
        case 31: run_all<3, 1, 64, 32, 1, 2>(h, flags, coef); break;
        case 32: run_all<3, 2, 64, 32, 1, 2>(h, flags, coef); break;
        case 33: run_all<3, 3, 32, 16, 4, 1>(h, flags, coef); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no assembly kernel for this (dim, degree)", __LINE__};
      }
  }
} // namespace pd
