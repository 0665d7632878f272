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
#include "pd_host.hpp"
#include "pd_internal.hpp"
#include "pd_device.cuh"

namespace pd
{
  namespace
  {
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
      const int32_t *cta_item_ptr; // [grid+1] items of every persistent CTA
      double        *partial;      // [n_items][N][N]
      double         stiffness, mass;
      Basis1D        basis;
    };

    template <int NROLE, int ROLE>
    __device__ __forceinline__ constexpr bool
    vol_role_owns_row(const int I)
    {
      // NROLE == 2: rows {0,3,4,7} vs {1,2,5,6}: 18 upper-triangle tiles each for NT8 == 8
      return NROLE == 1 ? true : ((((I & 3) == 0) || ((I & 3) == 3)) == (ROLE == 0));
    }

    __device__ __forceinline__ constexpr int
    tri_index(const int NT8, const int I, const int J) // I <= J
    {
      return I * NT8 - (I * (I - 1)) / 2 + (J - I);
    }

    // Volume kernel, software-pipelined.  Per stage of TQ = 32 points (lane = point):
    //   table warps : 1-D Lagrange tables of stage s+2 (+ prefetch of the raw x, JxW of s+3)
    //   gen warps   : operand rows of stage s+1 from the tables of s+1
    //   all warps   : DMMA contraction of stage s
    // then ONE barrier.  The phases of different warps overlap on the SM's pipes
    // (DMMA vs LSU/FP64), so the tensor pipe does not wait for the generator.
    //
    // Shared-memory layout (all accesses bank-conflict free):
    //   T[buf][d][2][N1][TQ]   tables, point fastest
    //   G[buf][i][RS]          operand panel, DOF-major, row index r = c*TQ + q fastest,
    //                          RS = 4 (mod 16): a fragment load (8 dofs x 4 rows) touches
    //                          every bank pair once per half-warp
    //   WC[3][R]               JxW * coefficient per row
    // SQ: the row weight w_q * coefficient (>= 0) is folded into the operand rows as its square
    // root by the generator, so the contraction is pure LDS + DMMA (the profile of the first
    // version showed a DMUL b = a * wc in front of almost every DMMA, on the same FP64 pipe).
    // !SQ (a negative coefficient): rows unscaled, b = a * wc in the contraction.
    template <int DIM, int DEG, bool MASS, bool SQ, int NWARPS, int NROLE, int ROLE>
    __device__ __forceinline__ void
    volume_body(const VolArgs &A, double *smem)
    {
      using C              = Cfg<DIM, DEG>;
      constexpr int TQ     = 32;
      constexpr int N1     = C::N1, N = C::N, NT8 = C::NT8, NP = C::NP, NU = C::NU, NTRI = C::NTRI;
      constexpr int NC     = DIM + (MASS ? 1 : 0);
      constexpr int R      = TQ * NC;
      constexpr int RS     = ((R + 11) / 16) * 16 + 4; // >= R, = 4 (mod 16)
      constexpr int KSPLIT = NWARPS / NROLE;
      constexpr int NTHR   = NWARPS * 32;
      constexpr int NTABW  = DIM;            // table warps: one per direction
      constexpr int NGENW  = NWARPS - NTABW; // generator warps
      constexpr int TSZ    = DIM * 2 * N1 * TQ;
      constexpr int GSZ    = NP * RS;
      static_assert(RS >= R && RS % 16 == 4, "panel row stride");
      static_assert(NGENW >= 1, "need generator warps");

      double *Gb = smem;               // [2][GSZ]
      double *Tb = Gb + 2 * GSZ;       // [2][TSZ]
      double *WC = Tb + 2 * TSZ;       // [3][R]
      double *S  = smem;               // [KSPLIT][NTRI][64] epilogue slabs alias the panels
      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int g = lane >> 2, t = lane & 3;
      const int kpart = warp / NROLE;
      const bool tabw = warp >= NGENW; // warp-uniform role
      const int  td   = tabw ? warp - NGENW : 0;

      const int item_begin = A.cta_item_ptr[blockIdx.x], item_end = A.cta_item_ptr[blockIdx.x + 1];

      for (int item = item_begin; item < item_end; ++item)
        {
          const int     poly = A.item_poly[item];
          const int64_t q0 = A.item_q0[item], q1 = A.item_q1[item];
          const int     nst = (int)((q1 - q0 + TQ - 1) / TQ);
          const double *bb  = A.bbox + (int64_t)poly * 2 * DIM;
          double        acc[NT8][NT8][2];
#pragma unroll
          for (int I = 0; I < NT8; ++I)
#pragma unroll
            for (int J = 0; J < NT8; ++J)
              acc[I][J][0] = acc[I][J][1] = 0.;

          const double lo = bb[td], hi = bb[DIM + td];
          const double inv_h = 1. / (hi - lo);
          double       x_pre = lo, w_pre = 0.; // raw data of the next stage to tabulate
          int          s_pre = 0;              // stage x_pre belongs to

          auto prefetch = [&](const int s) {
            const int64_t gq = q0 + (int64_t)s * TQ + lane;
            x_pre = lo;
            w_pre = 0.;
            if (s < nst && gq < q1)
              {
                x_pre = A.vq_x[(int64_t)td * A.Q + gq];
                if (td == 0)
                  w_pre = A.vq_w[gq];
              }
            s_pre = s;
          };
          // tables of stage s from the prefetched registers (table warps only)
          auto tables = [&](const int s) {
            const double x = x_pre, w = w_pre;
            prefetch(s + 1);
            if (s >= nst)
              return;
            // BoundingBox::real_to_unit, covariant scaling by 1/h
            const double xhat = (x - lo) / (hi - lo);
            double       L[N1], dL[N1];
            basis_1d<C>(A.basis, xhat, inv_h, L, dL);
            double *Tq = Tb + (s & 1) * TSZ + td * 2 * N1 * TQ + lane;
#pragma unroll
            for (int a = 0; a < N1; ++a)
              {
                Tq[a * TQ]        = L[a];
                Tq[(N1 + a) * TQ] = dL[a];
              }
            if (td == 0)
              {
                double *wc = WC + (s % 3) * R + lane;
#pragma unroll
                for (int c = 0; c < DIM; ++c)
                  wc[c * TQ] = SQ ? sqrt(w * A.stiffness) : w * A.stiffness;
                if (MASS)
                  wc[DIM * TQ] = SQ ? sqrt(w * A.mass) : w * A.mass;
              }
          };
          // operand rows of stage s (generator warps only): one warp-unit = all 32 points
          // for one (b[,c]) pair; the y/z products are formed once and swept over a
          auto generate = [&](const int s) {
            if (s >= nst)
              return;
            const double *Tq = Tb + (s & 1) * TSZ + lane;
            double       *Gq = Gb + (s & 1) * GSZ + lane;
            // sqrt(w * coefficient) of this lane's point: gradient rows / mass row
            const double sg = SQ ? WC[(s % 3) * R + lane] : 1.;
            const double sm = (SQ && MASS) ? WC[(s % 3) * R + DIM * TQ + lane] : 1.;
            for (int wu = warp; wu < NU; wu += NGENW)
              {
                if constexpr (DIM == 2)
                  {
                    const double ly0 = Tq[(2 * N1 + wu) * TQ];
                    const double ly = ly0 * sg, dy = Tq[(3 * N1 + wu) * TQ] * sg, lym = ly0 * sm;
                    const int    row0 = C::unit_base(wu), cnt = C::unit_count(wu);
#pragma unroll
                    for (int a = 0; a < N1; ++a)
                      {
                        if (a >= cnt)
                          break;
                        const double lx = Tq[a * TQ], dx = Tq[(N1 + a) * TQ];
                        double      *o  = Gq + (row0 + a) * RS;
                        o[0]            = dx * ly;
                        o[TQ]           = lx * dy;
                        if (MASS)
                          o[2 * TQ] = lx * lym;
                      }
                  }
                else
                  {
                    const int    b = wu % N1, c = wu / N1;
                    const double ly = Tq[(2 * N1 + b) * TQ], dy = Tq[(3 * N1 + b) * TQ];
                    const double lz = Tq[(4 * N1 + c) * TQ], dz = Tq[(5 * N1 + c) * TQ];
                    const double yz0 = ly * lz;
                    const double yz = yz0 * sg, dyz = dy * lz * sg, ydz = ly * dz * sg, yzm = yz0 * sm;
                    const int    row0 = C::unit_base(wu), cnt = C::unit_count(wu);
#pragma unroll
                    for (int a = 0; a < N1; ++a)
                      {
                        if (a >= cnt)
                          break;
                        const double lx = Tq[a * TQ], dx = Tq[(N1 + a) * TQ];
                        double      *o  = Gq + (row0 + a) * RS;
                        o[0]            = dx * yz;
                        o[TQ]           = lx * dyz;
                        o[2 * TQ]       = lx * ydz;
                        if (MASS)
                          o[3 * TQ] = lx * yzm;
                      }
                  }
              }
          };
          // C += sum_r wc_r G_r G_r^T on the upper triangle of tiles, rows of this k-part
          auto contract = [&](const int s) {
            const double *Gs = Gb + (s & 1) * GSZ + g * RS + t;
            const double *ws = WC + (s % 3) * R + t;
#pragma unroll 3
            for (int ks = kpart; ks < R / 4; ks += KSPLIT)
              {
                double a[NT8];
#pragma unroll
                for (int I = 0; I < NT8; ++I)
                  a[I] = Gs[8 * I * RS + ks * 4];
                if constexpr (SQ)
                  {
#pragma unroll
                    for (int I = 0; I < NT8; ++I)
                      if (vol_role_owns_row<NROLE, ROLE>(I))
                        {
#pragma unroll
                          for (int J = I; J < NT8; ++J)
                            dmma884(acc[I][J][0], acc[I][J][1], a[I], a[J]);
                        }
                  }
                else
                  {
                    const double wc = ws[ks * 4];
                    double       b[NT8];
#pragma unroll
                    for (int I = 0; I < NT8; ++I)
                      b[I] = a[I] * wc;
#pragma unroll
                    for (int I = 0; I < NT8; ++I)
                      if (vol_role_owns_row<NROLE, ROLE>(I))
                        {
#pragma unroll
                          for (int J = I; J < NT8; ++J)
                            dmma884(acc[I][J][0], acc[I][J][1], a[I], b[J]);
                        }
                  }
              }
          };

          // ---- pipeline fill
          cta_sync(); // previous item's epilogue is done with the aliased slabs
          if (tabw)
            {
              prefetch(0);
              tables(0);
            }
          cta_sync();
          if (tabw)
            tables(1);
          else
            generate(0);
          cta_sync();
          // ---- steady state: one barrier per stage
          for (int s = 0; s < nst; ++s)
            {
              if (tabw)
                tables(s + 2);
              else
                generate(s + 1);
              contract(s);
              cta_sync();
            }
          (void)s_pre;
          // ---- every k-part writes its partial tiles into its own slab, one barrier,
          //      then all threads sum the slabs while writing the n x n partial block
#pragma unroll
          for (int I = 0; I < NT8; ++I)
            if (vol_role_owns_row<NROLE, ROLE>(I))
              {
#pragma unroll
                for (int J = I; J < NT8; ++J)
                  {
                    double *sl = S + (kpart * NTRI + tri_index(NT8, I, J)) * 64 + g * 8 + 2 * t;
                    sl[0]      = acc[I][J][0];
                    sl[1]      = acc[I][J][1];
                  }
              }
          cta_sync();
          double *out = A.partial + (int64_t)item * N * N;
          for (int idx = tid; idx < N * N; idx += NTHR)
            {
              int i = idx / N, j = idx - i * N;
              if ((i >> 3) > (j >> 3))
                {
                  const int sw = i;
                  i            = j;
                  j            = sw;
                }
              const double *sl = S + tri_index(NT8, i >> 3, j >> 3) * 64 + (i & 7) * 8 + (j & 7);
              double        v  = 0.;
#pragma unroll
              for (int k = 0; k < KSPLIT; ++k)
                v += sl[k * NTRI * 64];
              out[idx] = v;
            }
        }
    }

    template <int DIM, int DEG, bool MASS, bool SQ, int NWARPS>
    __global__ void __launch_bounds__(NWARPS * 32, NWARPS == 8 ? 3 : 1)
    k_volume(const VolArgs A)
    {
      extern __shared__ double smem[];
      constexpr int NROLE = Cfg<DIM, DEG>::NT8 >= 8 ? 2 : 1;
      if constexpr (NROLE == 1)
        volume_body<DIM, DEG, MASS, SQ, NWARPS, 1, 0>(A, smem);
      else
        {
          if (((threadIdx.x >> 5) & 1) == 0)
            volume_body<DIM, DEG, MASS, SQ, NWARPS, 2, 0>(A, smem);
          else
            volume_body<DIM, DEG, MASS, SQ, NWARPS, 2, 1>(A, smem);
        }
    }

    template <int DIM, int DEG, bool MASS, int NWARPS>
    constexpr size_t
    volume_smem_bytes()
    {
      using C             = Cfg<DIM, DEG>;
      constexpr int TQ    = 32;
      constexpr int NC    = DIM + (MASS ? 1 : 0);
      constexpr int R     = TQ * NC;
      constexpr int RS    = ((R + 11) / 16) * 16 + 4;
      constexpr int NROLE = C::NT8 >= 8 ? 2 : 1;
      size_t        a     = (size_t)2 * C::NP * RS + (size_t)2 * DIM * 2 * C::N1 * TQ + 3 * R;
      size_t        s     = (size_t)(NWARPS / NROLE) * C::NTRI * 64;
      return sizeof(double) * (a > s ? a : s);
    }

    // -------------------------------------------------------------------------
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

    // Face kernel, software-pipelined like the volume kernel (stage = 32 points, lane =
    // point, one barrier per stage):
    //   table tasks (side, direction): l_a and l_a' n_d / h_d of stage s+2
    //   generator units (side, b[,c]): V and Z columns of stage s+1
    //   all warps: DMMA contraction of stage s
    // Warp roles in the contraction.  Interior interface: 4 quadrants (Z side, V side) x
    // ISPLIT row groups x KSPLIT k-parts, each warp MI x NT8 tiles.  Boundary face: only
    // the (0,0) quadrant exists, so the quadrant index becomes an extra 4-way k-split.
    // Panels are DOF-major with row stride 36 = 4 (mod 16): conflict-free stores
    // (lane = point) and fragment loads.
    template <int DIM, int DEG, int ISPLIT, int KSPLIT, int MINB>
    __global__ void __launch_bounds__(4 * ISPLIT * KSPLIT * 32, MINB)
    k_faces(const FaceArgs A)
    {
      using C              = Cfg<DIM, DEG>;
      constexpr int TQ     = 32;
      constexpr int RS     = 36;
      constexpr int N1     = C::N1, N = C::N, NT8 = C::NT8, NP = C::NP, NU = C::NU;
      constexpr int NP2    = 2 * NP;
      constexpr int NWARPS = 4 * ISPLIT * KSPLIT;
      constexpr int NTHR   = NWARPS * 32;
      constexpr int MI     = NT8 / ISPLIT;
      constexpr int SS     = NP2 + 1;    // slab row stride
      constexpr int SLAB   = NP2 * SS;   // doubles per k-part slab (interior)
      constexpr int KB     = 4 * KSPLIT; // k-parts of a boundary face
      constexpr int SSB    = NP + 1;
      constexpr int SLABB  = NP * SSB;
      constexpr int NTAB   = 2 * DIM;    // table tasks
      constexpr int PSZ    = NP2 * RS;   // one panel (Z or V) of one buffer
      constexpr int TSZ    = 2 * DIM * 2 * N1 * TQ;
      static_assert(NT8 % ISPLIT == 0, "row split must divide the tile count");
      static_assert(NTAB <= NWARPS, "one warp per table task");

      extern __shared__ double smem[];
      double *Pb = smem;             // [2 bufs][Z,V][NP2][RS]
      double *Tb = Pb + 4 * PSZ;     // [2 bufs][side][d][2][N1][TQ]
      double *WC = Tb + 2 * TSZ;     // [3][TQ] JxW * coefficient
      double *SG = WC + 3 * TQ;      // [2][TQ] sigma / 2
      double *S  = smem;             // epilogue slabs alias the panels

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int g = lane >> 2, t = lane & 3;
      const int quad  = warp & 3;
      const int isub  = (warp >> 2) % ISPLIT;
      const int kpart = (warp >> 2) / ISPLIT;
      const bool tabw  = warp < NTAB;
      const int  tside = tabw ? warp / DIM : 0, td = tabw ? warp % DIM : 0;
      const bool lead  = tabw && warp == 0; // task (side 0, d 0) also carries JxW and sigma
      const int  first_unit = (warp - NTAB + NWARPS) % NWARPS;

      for (int f = blockIdx.x; f < A.n_ifaces; f += gridDim.x)
        {
          const int  pa = A.ifA[f], pb = A.ifB[f];
          const bool interior = pb >= 0;
          if (interior ? !(A.flags & PD_ASSEMBLE_INTERIOR) : !(A.flags & PD_ASSEMBLE_BOUNDARY))
            continue;
          const int64_t s0 = A.if_sub_ptr[f], s1 = A.if_sub_ptr[f + 1];
          const int64_t q0 = s0 * A.nqf, q1 = s1 * A.nqf;
          const int     nst = (int)((q1 - q0 + TQ - 1) / TQ);
          const double *bbt = A.bbox + (int64_t)((tside && interior) ? pb : pa) * 2 * DIM;
          const double  lo = bbt[td], hi = bbt[DIM + td];
          const double  inv_h = 1. / (hi - lo);
          const bool    tab_active = tabw && (interior || tside == 0);
          // interior: (qa, qb) quadrant, k-part kpart of KSPLIT; boundary: quadrant (0,0),
          // k-part kpart*4 + quad of KB
          const int qa = interior ? (quad >> 1) : 0, qb = interior ? (quad & 1) : 0;
          const int kp = interior ? kpart : kpart * 4 + quad;
          const int kn = interior ? KSPLIT : KB;
          const int nunits = (interior ? 2 : 1) * NU;
          const double dscale = interior ? 0.5 : 1.;

          double acc[MI][NT8][2];
#pragma unroll
          for (int I = 0; I < MI; ++I)
#pragma unroll
            for (int J = 0; J < NT8; ++J)
              acc[I][J][0] = acc[I][J][1] = 0.;

          double x_pre = lo, n_pre = 0., w_pre = 0., sg_pre = 0.;
          auto   prefetch = [&](const int s) {
            const int64_t gq = q0 + (int64_t)s * TQ + lane;
            x_pre = lo;
            n_pre = w_pre = sg_pre = 0.;
            if (s < nst && gq < q1)
              {
                x_pre = A.fq_x[(int64_t)td * A.Qf + gq];
                n_pre = A.fq_n[(int64_t)td * A.Qf + gq];
                if (lead)
                  {
                    w_pre  = A.fq_w[gq];
                    sg_pre = A.sub_sigma[gq / A.nqf];
                  }
              }
          };
          auto tables = [&](const int s) {
            const double x = x_pre, nd = n_pre, w = w_pre, sg = sg_pre;
            prefetch(s + 1);
            if (s >= nst)
              return;
            const double xhat = (x - lo) / (hi - lo);
            double       L[N1], dL[N1];
            basis_1d<C>(A.basis, xhat, nd * inv_h, L, dL);
            double *Tq = Tb + (s & 1) * TSZ + (tside * DIM + td) * 2 * N1 * TQ + lane;
#pragma unroll
            for (int a = 0; a < N1; ++a)
              {
                Tq[a * TQ]        = L[a];
                Tq[(N1 + a) * TQ] = dL[a];
              }
            if (lead)
              {
                WC[(s % 3) * TQ + lane] = w * A.stiffness;
                SG[(s & 1) * TQ + lane] = 0.5 * sg;
              }
          };
          // V and Z columns.  interior: V = [phi0 ; -phi1], Z = sigma/2 V - [dn0 ; dn1]/2.
          // boundary: V = phi0, Z = sigma/2 phi0 - dn0.
          auto generate = [&](const int s) {
            if (s >= nst)
              return;
            const double hs = SG[(s & 1) * TQ + lane];
            const double wq = WC[(s % 3) * TQ + lane];
            double      *Zp = Pb + (s & 1) * 2 * PSZ + lane;
            double      *Vp = Zp + PSZ;
            for (int wu = first_unit; wu < nunits; wu += NWARPS)
              {
                const int     side = wu / NU, bc = wu - side * NU;
                const double *Tq   = Tb + (s & 1) * TSZ + side * DIM * 2 * N1 * TQ + lane;
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
                const int row0 = C::unit_base(bc), cnt = C::unit_count(bc);
#pragma unroll
                for (int a = 0; a < N1; ++a)
                  {
                    if (a >= cnt)
                      break;
                    const double lx = Tq[a * TQ], dx = Tq[(N1 + a) * TQ];
                    const double v  = lx * s1;
                    const double dn = dx * s1 + lx * s2;
                    const double V  = sgn * v;
                    const int    col = side * NP + row0 + a;
                    Vp[col * RS]     = V * wq; // JxW * coefficient folded into the V operand
                    Zp[col * RS]     = hs * V - dscale * dn;
                  }
              }
          };
          // T_(qa,qb) += sum_q wc Z_qa V_qb^T
          auto contract = [&](const int s) {
            const double *Zs = Pb + (s & 1) * 2 * PSZ + (qa * NP + isub * MI * 8 + g) * RS + t;
            const double *Vs = Pb + (s & 1) * 2 * PSZ + PSZ + (qb * NP + g) * RS + t;
            // the last stage of an interface is usually partial (e.g. 144 points = 4.5 stages):
            // skip the k-steps that only hold padding rows
            const int64_t left = q1 - (q0 + (int64_t)s * TQ);
            const int     nk   = left >= TQ ? TQ / 4 : (int)((left + 3) / 4);
            for (int ks = kp; ks < nk; ks += kn)
              {
                double a[MI], b[NT8];
#pragma unroll
                for (int I = 0; I < MI; ++I)
                  a[I] = Zs[8 * I * RS + ks * 4];
#pragma unroll
                for (int J = 0; J < NT8; ++J)
                  b[J] = Vs[8 * J * RS + ks * 4];
#pragma unroll
                for (int I = 0; I < MI; ++I)
#pragma unroll
                  for (int J = 0; J < NT8; ++J)
                    dmma884(acc[I][J][0], acc[I][J][1], a[I], b[J]);
              }
          };

          // ---- pipeline fill
          __syncthreads(); // previous interface's epilogue is done with the aliased slabs
          if (tab_active)
            {
              prefetch(0);
              tables(0);
            }
          __syncthreads();
          if (tab_active)
            tables(1);
          generate(0);
          __syncthreads();
          // ---- steady state
          for (int s = 0; s < nst; ++s)
            {
              if (tab_active)
                tables(s + 2);
              generate(s + 1);
              contract(s);
              __syncthreads();
            }
          // ---- epilogue: k-part slabs -> M = T + T^T -> global
          double *fd = A.face_diag + (int64_t)f * 2 * N * N;
          if (interior)
            {
#pragma unroll
              for (int I = 0; I < MI; ++I)
#pragma unroll
                for (int J = 0; J < NT8; ++J)
                  {
                    double *sl = S + kp * SLAB + (qa * NP + (isub * MI + I) * 8 + g) * SS + qb * NP + 8 * J + 2 * t;
                    sl[0]      = acc[I][J][0];
                    sl[1]      = acc[I][J][1];
                  }
              __syncthreads();
              const int64_t baseAB = A.if_baseAB[f], baseBA = A.if_baseBA[f];
              const int     strideA = A.row_stride[A.dof_block[pa]];
              const int     strideB = baseBA >= 0 ? A.row_stride[A.dof_block[pb]] : 0;
              for (int idx = tid; idx < 4 * N * N; idx += NTHR)
                {
                  const int which = idx / (N * N), rem = idx - which * N * N;
                  const int i = rem / N, j = rem - i * N;
                  const int ri = (which >> 1) * NP + i, cj = (which & 1) * NP + j;
                  double    m = 0.;
#pragma unroll
                  for (int k = 0; k < KSPLIT; ++k)
                    m += S[k * SLAB + ri * SS + cj] + S[k * SLAB + cj * SS + ri];
                  if (which == 0)
                    fd[rem] = m;
                  else if (which == 3)
                    fd[N * N + rem] = m;
                  else if (which == 1)
                    A.values[baseAB + (int64_t)i * strideA + j] = m;
                  else if (baseBA >= 0) // B is a ghost polytope: its rows live on another rank
                    A.values[baseBA + (int64_t)i * strideB + j] = m;
                }
            }
          else
            {
#pragma unroll
              for (int I = 0; I < MI; ++I)
#pragma unroll
                for (int J = 0; J < NT8; ++J)
                  {
                    double *sl = S + kp * SLABB + ((isub * MI + I) * 8 + g) * SSB + 8 * J + 2 * t;
                    sl[0]      = acc[I][J][0];
                    sl[1]      = acc[I][J][1];
                  }
              __syncthreads();
              for (int idx = tid; idx < N * N; idx += NTHR)
                {
                  const int i = idx / N, j = idx - i * N;
                  double    m = 0.;
#pragma unroll
                  for (int k = 0; k < KB; ++k)
                    m += S[k * SLABB + i * SSB + j] + S[k * SLABB + j * SSB + i];
                  fd[idx] = m;
                }
            }
        }
    }

    template <int DIM, int DEG, int KSPLIT>
    constexpr size_t
    face_smem_bytes()
    {
      using C           = Cfg<DIM, DEG>;
      constexpr int TQ  = 32, RS = 36;
      constexpr int NP2 = 2 * C::NP;
      size_t        a   = (size_t)4 * NP2 * RS + (size_t)2 * 2 * DIM * 2 * C::N1 * TQ + 5 * TQ;
      size_t        s   = (size_t)KSPLIT * NP2 * (NP2 + 1);
      size_t        sb  = (size_t)4 * KSPLIT * C::NP * (C::NP + 1);
      s                 = s > sb ? s : sb;
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

    // Stage-balanced persistent schedule: the stages (TQ points of one polytope) of all
    // polytopes are dealt to `grid` CTAs in equal contiguous shares; an item is a maximal
    // run of one polytope's stages inside one share.  Every CTA does the same amount of
    // MMA work whatever the polytope sizes are (METIS shapes vary in #sub-cells).
    void
    plan_volume(pd_handle *h, const int tq, const int grid)
    {
      if (h->vol_plan_tq == tq && h->vol_plan_grid == grid)
        return;
      const std::vector<int64_t> &sub = h->h_subcell_ptr;
      std::vector<int64_t> stage_ptr(h->np + 1, 0);
      for (int32_t p = 0; p < h->np; ++p)
        stage_ptr[p + 1] = stage_ptr[p] + ((sub[p + 1] - sub[p]) * h->nqc + tq - 1) / tq;
      const int64_t        S = stage_ptr[h->np];
      std::vector<int32_t> item_poly, cta_ptr(grid + 1, 0);
      std::vector<int64_t> q0, q1, poly_item_ptr(h->np + 1, 0);
      // items are emitted polytope-major so that a polytope's partials are contiguous;
      // a polytope's stage range is cut wherever a CTA share boundary falls inside it
      std::vector<int32_t> item_cta;
      for (int32_t p = 0; p < h->np; ++p)
        {
          int64_t s = stage_ptr[p];
          while (s < stage_ptr[p + 1])
            {
              const int64_t c   = std::min<int64_t>(grid - 1, (s * grid) / std::max<int64_t>(S, 1));
              // first stage of CTA c+1 = smallest s' with floor(s' grid / S) >= c+1
              int64_t end = ((c + 1) * S + grid - 1) / grid;
              end         = std::min(end, stage_ptr[p + 1]);
              if (end <= s)
                end = s + 1;
              const int64_t base = sub[p] * h->nqc;
              item_poly.push_back(p);
              item_cta.push_back((int32_t)c);
              q0.push_back(base + (s - stage_ptr[p]) * tq);
              q1.push_back(std::min(base + (end - stage_ptr[p]) * tq, sub[p + 1] * h->nqc));
              s = end;
            }
          poly_item_ptr[p + 1] = (int64_t)item_poly.size();
        }
      // items are already sorted by CTA (shares are contiguous in stage order)
      for (size_t i = 0; i < item_cta.size(); ++i)
        ++cta_ptr[item_cta[i] + 1];
      for (int c = 0; c < grid; ++c)
        cta_ptr[c + 1] += cta_ptr[c];
      h->n_vitems = (int32_t)item_poly.size();
      auto put = [](auto &buf, const auto &v) {
        buf.alloc(v.size());
        if (!v.empty())
          PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
      };
      put(h->vitem_poly, item_poly);
      put(h->vitem_q0, q0);
      put(h->vitem_q1, q1);
      put(h->poly_vitem_ptr, poly_item_ptr);
      put(h->cta_item_ptr, cta_ptr);
      h->vol_partial.alloc((size_t)h->n_vitems * h->n * h->n);
      h->vol_plan_tq   = tq;
      h->vol_plan_grid = grid;
    }

    template <int DIM, int DEG, bool MASS, bool SQ>
    void
    run_volume(pd_handle *h, const pd_coefficients &coef)
    {
      constexpr int    TQ     = 32;
      constexpr int    NWARPS = Cfg<DIM, DEG>::NT8 >= 8 ? 16 : 8;
      auto             kern   = k_volume<DIM, DEG, MASS, SQ, NWARPS>;
      constexpr size_t smem   = volume_smem_bytes<DIM, DEG, MASS, NWARPS>();
      set_smem(kern, smem);
      int per_sm = 1;
      PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NWARPS * 32, smem));
      if (per_sm < 1)
        per_sm = 1;
      const int64_t total_stages = (h->Q + TQ - 1) / TQ + h->np;
      const int     grid         = (int)std::max<int64_t>(1, std::min<int64_t>(total_stages / 4 + 1, (int64_t)h->sm_count * per_sm));
      plan_volume(h, TQ, grid);
      VolArgs a;
      a.vq_x         = h->vq_x.p;
      a.vq_w         = h->vq_w.p;
      a.Q            = h->Q;
      a.bbox         = h->bbox.p;
      a.item_poly    = h->vitem_poly.p;
      a.item_q0      = h->vitem_q0.p;
      a.item_q1      = h->vitem_q1.p;
      a.cta_item_ptr = h->cta_item_ptr.p;
      a.partial      = h->vol_partial.p;
      a.stiffness    = coef.stiffness;
      a.mass         = coef.mass;
      a.basis        = h->basis;
      kern<<<grid, NWARPS * 32, smem, h->stream>>>(a);
      ++h->launches;
    }

    template <int DIM, int DEG, int ISPLIT, int KSPLIT, int MINB>
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
      auto             kern = k_faces<DIM, DEG, ISPLIT, KSPLIT, MINB>;
      constexpr size_t smem = face_smem_bytes<DIM, DEG, KSPLIT>();
      constexpr int    nthr = 4 * ISPLIT * KSPLIT * 32;
      set_smem(kern, smem);
      int per_sm = 1;
      PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthr, smem));
      if (per_sm < 1)
        per_sm = 1;
      const int grid = std::min<int64_t>(h->n_ifaces, (int64_t)h->sm_count * per_sm);
      kern<<<grid, nthr, smem, h->stream>>>(a);
      ++h->launches;
    }

    template <int DIM, int DEG, int ISPLIT, int KSPLIT, int MINB>
    void
    run_all(pd_handle *h, const uint32_t flags, const pd_coefficients &coef)
    {
      PD_CUDA(cudaEventRecord(h->ev[0], h->stream));
      if ((flags & PD_ASSEMBLE_VOLUME) && h->Q > 0)
        {
          // non-negative coefficients (every use in the reference): weights folded into the
          // operands as square roots; otherwise the signed variant
          const bool sq = coef.stiffness >= 0. && coef.mass >= 0.;
          if (coef.mass != 0.)
            sq ? run_volume<DIM, DEG, true, true>(h, coef) : run_volume<DIM, DEG, true, false>(h, coef);
          else
            sq ? run_volume<DIM, DEG, false, true>(h, coef) : run_volume<DIM, DEG, false, false>(h, coef);
        }
      PD_CUDA(cudaEventRecord(h->ev[1], h->stream));
      if ((flags & (PD_ASSEMBLE_BOUNDARY | PD_ASSEMBLE_INTERIOR)) && h->n_ifaces > 0)
        run_faces<DIM, DEG, ISPLIT, KSPLIT, MINB>(h, coef, flags);
      PD_CUDA(cudaEventRecord(h->ev[2], h->stream));
      if (h->vol_plan_tq == 0) // volume never planned (flags without VOLUME): empty item lists
        plan_volume(h, 32, 1);
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
      r.np             = h->np_own;
      r.n              = h->n;
      r.flags          = flags;
      const int grid   = std::min<int64_t>(h->np_own, (int64_t)h->sm_count * 8);
      k_reduce_diag<<<grid, 256, 0, h->stream>>>(r);
      ++h->launches;
      PD_CUDA(cudaEventRecord(h->ev[3], h->stream));
      PD_CUDA(cudaGetLastError());
    }
  } // namespace

  void
  plan_volume_items(pd_handle *h, const int tq, const int grid)
  {
    plan_volume(h, tq, grid);
  }

  bool
  assemble_supported(const int dim, const int degree, const int fe_kind)
  {
    if (fe_kind != PD_FE_DGQ && fe_kind != PD_FE_AGGLODGP)
      return false;
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
    const int key = h->fe_kind * 100 + h->dim * 10 + h->degree;
    switch (key)
      {
        //                 DIM DEG ISPLIT KSPLIT MINB
        case 121: run_all<2, DGP_BASE + 1, 1, 2, 2>(h, flags, coef); break; // FE_AggloDGP
        case 122: run_all<2, DGP_BASE + 2, 1, 2, 2>(h, flags, coef); break;
        case 123: run_all<2, DGP_BASE + 3, 1, 2, 2>(h, flags, coef); break;
        case 124: run_all<2, DGP_BASE + 4, 1, 2, 2>(h, flags, coef); break;
        case 131: run_all<3, DGP_BASE + 1, 1, 2, 2>(h, flags, coef); break;
        case 132: run_all<3, DGP_BASE + 2, 1, 2, 2>(h, flags, coef); break;
        case 133: run_all<3, DGP_BASE + 3, 1, 2, 2>(h, flags, coef); break;
        case 21: run_all<2, 1, 1, 2, 2>(h, flags, coef); break;
        case 22: run_all<2, 2, 1, 2, 2>(h, flags, coef); break;
        case 23: run_all<2, 3, 1, 2, 2>(h, flags, coef); break;
        case 24: run_all<2, 4, 1, 2, 2>(h, flags, coef); break;
        case 31: run_all<3, 1, 1, 2, 2>(h, flags, coef); break;
        case 32: run_all<3, 2, 1, 2, 2>(h, flags, coef); break;
        case 33: run_all<3, 3, 4, 1, 1>(h, flags, coef); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no assembly kernel for this (dim, degree)", __LINE__};
      }
  }
} // namespace pd
