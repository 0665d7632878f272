// -----------------------------------------------------------------------------
// pd_polyapply.cu -- point-wise evaluate / integrate kernels on AGGLOMERATED
// polytopes: the matrix-free SIP apply, the right-hand side and the error norms.
//
// All three walk the agglomerated quadrature of a polytope (MappingBox + FE_DGQ on
// the bounding box, source/agglomeration_handler.cc:729-906, source/mapping_box.cc:
// 393-532) and differ only in what happens at a point:
//   apply  y_P += G^T diag(w c) G u_P                       (include/poly_utils.h:2038-2052)
//          faces: the M11/M12/M21/M22 rows of poly_utils.h:1870-1926 applied to [u_A; u_B],
//          boundary rows of :2060-2085
//   rhs    b_i = sum_q f(x_q) phi_i w_q                      (examples/poisson.cc:745-761)
//          + sum_bdry [ sigma g phi_i - (grad phi_i . n) g ] w (examples/diffusion_reaction.cc:550-556)
//   error  sum_q (u_h - u)^2 w,  sum_q |grad u_h - grad u|^2 w (include/poly_utils.h:1647-1750)
// The reference has no matrix-free operator on the agglomerated space;
// the apply is the memory-free alternative to the block-CSR one (no 8 n^2 bytes per block).
//
// One LANE per quadrature point: the lane builds the 1-D Lagrange tables of its point in
// registers, evaluates u_h / grad u_h by sum factorisation against the polytope's
// coefficients (broadcast reads from shared memory), and integrates back into a PRIVATE
// accumulator per DoF -- no barrier inside the point loop.  For n = 64 two lanes share a
// point (each owns half of the z-slices).  Accumulators meet once per work item (warp
// shuffles, then shared memory), every item writes its own partial vector and
// k_poly_gather sums a polytope's partials in a fixed order: deterministic.
// (The first version staged operand panels for 32 points in shared memory with four block
// barriers per tile and ran 20-30x slower.)
// -----------------------------------------------------------------------------
#include "pd_device.cuh"
#include "pd_host.hpp"

#include <algorithm>
#include <cmath>

namespace pd
{
  namespace
  {
    constexpr int NW = 8; // warps per CTA
    enum
    {
      MODE_APPLY = 0,
      MODE_RHS   = 1,
      MODE_ERROR = 2
    };

    template <int DIM, int N1>
    struct PointTab
    {
      double L[DIM][N1], dL[DIM][N1]; // l_a(xhat_d), l_a'(xhat_d) / h_d
    };

    template <int DIM, int N1, bool DGP = false>
    __device__ __forceinline__ void
    point_tables(const Basis1D &B, const double *bb, const double (&x)[DIM], PointTab<DIM, N1> &T)
    {
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          const double ih = 1. / (bb[DIM + d] - bb[d]);
          if constexpr (DGP) // FE_AggloDGP: orthonormal Legendre factors
            legendre01<N1>((x[d] - bb[d]) * ih, ih, T.L[d], T.dL[d]);
          else
            lagrange<N1>(B, (x[d] - bb[d]) * ih, ih, T.L[d], T.dL[d]);
        }
    }

    // the slices of the last tensor index this lane owns (part of GZ lanes per point)
    template <int DIM, int N1, int GZ>
    __device__ __forceinline__ void
    my_slices(const PointTab<DIM, N1> &T, const int part, double (&ls)[N1 / GZ], double (&dls)[N1 / GZ])
    {
      constexpr int NS = N1 / GZ;
#pragma unroll
      for (int cc = 0; cc < NS; ++cc)
        {
          ls[cc]  = T.L[DIM - 1][cc];
          dls[cc] = T.dL[DIM - 1][cc];
#pragma unroll
          for (int g = 1; g < GZ; ++g)
            if (part == g)
              {
                ls[cc]  = T.L[DIM - 1][g * NS + cc];
                dls[cc] = T.dL[DIM - 1][g * NS + cc];
              }
        }
    }

    // u_h and grad u_h at the point, summed over this lane's slices; Us = first owned slice
    template <int DIM, int N1, int NS, bool DGP = false>
    __device__ __forceinline__ void
    eval_slices(const PointTab<DIM, N1> &T, const double (&ls)[NS], const double (&dls)[NS], const double *Us, double &u,
                double (&g)[DIM])
    {
      u = 0.;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        g[d] = 0.;
      if constexpr (DGP)
        {
          // FE_AggloDGP: the DoFs (a, b[, c]) with a + b [+ c] <= p, last index outermost, first fastest
          constexpr int P   = N1 - 1;
          int           idx = 0; // compile-time after unrolling
          if constexpr (DIM == 3)
            {
#pragma unroll
              for (int c = 0; c <= P; ++c)
                {
                  double Yv = 0., Yd = 0., Yx = 0.;
#pragma unroll
                  for (int b = 0; b <= P - c; ++b)
                    {
                      double X0 = 0., X1 = 0.;
#pragma unroll
                      for (int a = 0; a <= P - b - c; ++a)
                        {
                          const double v = *(const volatile double *)&Us[idx++];
                          X0 += T.L[0][a] * v;
                          X1 += T.dL[0][a] * v;
                        }
                      Yv += T.L[1][b] * X0;
                      Yd += T.dL[1][b] * X0;
                      Yx += T.L[1][b] * X1;
                    }
                  u += T.L[2][c] * Yv;
                  g[0] += T.L[2][c] * Yx;
                  g[1] += T.L[2][c] * Yd;
                  g[2] += T.dL[2][c] * Yv;
                }
            }
          else
            {
#pragma unroll
              for (int b = 0; b <= P; ++b)
                {
                  double X0 = 0., X1 = 0.;
#pragma unroll
                  for (int a = 0; a <= P - b; ++a)
                    {
                      const double v = *(const volatile double *)&Us[idx++];
                      X0 += T.L[0][a] * v;
                      X1 += T.dL[0][a] * v;
                    }
                  u += T.L[1][b] * X0;
                  g[0] += T.L[1][b] * X1;
                  g[1] += T.dL[1][b] * X0;
                }
            }
        }
      else if constexpr (DIM == 3)
        {
#pragma unroll
          for (int cc = 0; cc < NS; ++cc)
            {
              double Yv = 0., Yd = 0., Yx = 0.;
#pragma unroll
              for (int b = 0; b < N1; ++b)
                {
                  double X0 = 0., X1 = 0.;
#pragma unroll
                  for (int a = 0; a < N1; ++a)
                    {
                      const double v = *(const volatile double *)&Us[(cc * N1 + b) * N1 + a]; // re-read: keeps the coefficients out of registers
                      X0 += T.L[0][a] * v;
                      X1 += T.dL[0][a] * v;
                    }
                  Yv += T.L[1][b] * X0;
                  Yd += T.dL[1][b] * X0;
                  Yx += T.L[1][b] * X1;
                }
              u += ls[cc] * Yv;
              g[0] += ls[cc] * Yx;
              g[1] += ls[cc] * Yd;
              g[2] += dls[cc] * Yv;
            }
        }
      else
        {
#pragma unroll
          for (int bb = 0; bb < NS; ++bb)
            {
              double X0 = 0., X1 = 0.;
#pragma unroll
              for (int a = 0; a < N1; ++a)
                {
                  const double v = *(const volatile double *)&Us[bb * N1 + a];
                  X0 += T.L[0][a] * v;
                  X1 += T.dL[0][a] * v;
                }
              u += ls[bb] * X0;
              g[0] += ls[bb] * X1;
              g[1] += dls[bb] * X0;
            }
        }
    }

    // acc_i += fm phi_i + f . grad phi_i for the DoFs of this lane's slices
    template <int DIM, int N1, int NS, bool GRAD, bool DGP = false>
    __device__ __forceinline__ void
    integrate_slices(const PointTab<DIM, N1> &T, const double (&ls)[NS], const double (&dls)[NS], const double (&f)[DIM],
                     const double fm, double *acc)
    {
      if constexpr (DGP)
        {
          constexpr int P   = N1 - 1;
          int           idx = 0;
          if constexpr (DIM == 3)
            {
#pragma unroll
              for (int c = 0; c <= P; ++c)
#pragma unroll
                for (int b = 0; b <= P - c; ++b)
                  {
                    const double yz = T.L[1][b] * T.L[2][c];
                    double       A = 0., Bv = fm * yz;
                    if constexpr (GRAD)
                      {
                        A = f[0] * yz;
                        Bv += f[1] * (T.dL[1][b] * T.L[2][c]) + f[2] * (T.L[1][b] * T.dL[2][c]);
                      }
#pragma unroll
                    for (int a = 0; a <= P - b - c; ++a, ++idx)
                      {
                        double t = acc[idx] + T.L[0][a] * Bv;
                        if constexpr (GRAD)
                          t += T.dL[0][a] * A;
                        acc[idx] = t;
                      }
                  }
            }
          else
            {
#pragma unroll
              for (int b = 0; b <= P; ++b)
                {
                  double A = 0., Bv = fm * T.L[1][b];
                  if constexpr (GRAD)
                    {
                      A = f[0] * T.L[1][b];
                      Bv += f[1] * T.dL[1][b];
                    }
#pragma unroll
                  for (int a = 0; a <= P - b; ++a, ++idx)
                    {
                      double t = acc[idx] + T.L[0][a] * Bv;
                      if constexpr (GRAD)
                        t += T.dL[0][a] * A;
                      acc[idx] = t;
                    }
                }
            }
        }
      else if constexpr (DIM == 3)
        {
#pragma unroll
          for (int cc = 0; cc < NS; ++cc)
#pragma unroll
            for (int b = 0; b < N1; ++b)
              {
                const double yz = T.L[1][b] * ls[cc];
                double       A = 0., Bv = fm * yz;
                if constexpr (GRAD)
                  {
                    A = f[0] * yz;
                    Bv += f[1] * (T.dL[1][b] * ls[cc]) + f[2] * (T.L[1][b] * dls[cc]);
                  }
#pragma unroll
                for (int a = 0; a < N1; ++a)
                  {
                    double t = acc[(cc * N1 + b) * N1 + a] + T.L[0][a] * Bv;
                    if constexpr (GRAD)
                      t += T.dL[0][a] * A;
                    acc[(cc * N1 + b) * N1 + a] = t;
                  }
              }
        }
      else
        {
#pragma unroll
          for (int bb = 0; bb < NS; ++bb)
            {
              double A = 0., Bv = fm * ls[bb];
              if constexpr (GRAD)
                {
                  A = f[0] * ls[bb];
                  Bv += f[1] * dls[bb];
                }
#pragma unroll
              for (int a = 0; a < N1; ++a)
                {
                  double t = acc[bb * N1 + a] + T.L[0][a] * Bv;
                  if constexpr (GRAD)
                    t += T.dL[0][a] * A;
                  acc[bb * N1 + a] = t;
                }
            }
        }
    }

    // sum the accumulators of one warp's work item over its points (lanes of the same
    // part / side group are PTS apart) and write the item's partial vector
    template <int NA, int PTS>
    __device__ __forceinline__ void
    warp_reduce_store(const double *acc, double *out /*[NGRP*NA]*/, const int lane)
    {
      const int pt = lane % PTS, grp = lane / PTS;
#pragma unroll
      for (int k = 0; k < NA; ++k)
        {
          double v = acc[k];
#pragma unroll
          for (int o = PTS / 2; o > 0; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
          if (pt == (k % PTS))
            out[grp * NA + k] = v;
        }
    }

    struct VolArgs
    {
      const double  *vq_x, *vq_w;
      int64_t        Q;
      const double  *bbox;
      const int32_t *dof_block;
      const int32_t *item_poly;
      const int64_t *item_q0, *item_q1;
      int32_t        n_items;
      const double  *x;       // coefficients (apply, error)
      const double  *data0;   // rhs: f at the points; error: exact values
      const double  *data1;   // error: exact gradient, SoA [dim][Q], or null
      double        *partial; // [n_items][N] (apply, rhs) or [n_items][2] (error)
      double         stiffness, mass;
      Basis1D        basis;
    };

    // one WARP per work item (a run of quadrature points of one polytope)
    template <int DIM, int DEG, int MODE, bool MASS, int GZ>
    __global__ void __launch_bounds__(NW * 32, 2) k_pw_volume(const VolArgs A)
    {
      using C           = Cfg<DIM, DEG>;
      constexpr int N1  = C::N1, N = C::N;
      constexpr int NS  = N1 / GZ, NA = C::DGP ? N : NS * (N / N1);
      constexpr int PTS = 32 / GZ;
      static_assert(N1 % GZ == 0 && GZ * NA == N && (!C::DGP || GZ == 1), "slices must tile the last index");
      __shared__ double Us[NW][N];

      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pt = lane % PTS, part = lane / PTS;
      double   *U = Us[warp];
      for (int item = blockIdx.x * NW + warp; item < A.n_items; item += gridDim.x * NW)
        {
          const int     poly = A.item_poly[item];
          const int64_t q0 = A.item_q0[item], q1 = A.item_q1[item];
          const double *bb = A.bbox + (int64_t)poly * 2 * DIM;
          __syncwarp();
          if (MODE != MODE_RHS)
            for (int i = lane; i < N; i += 32)
              U[i] = A.x[(int64_t)A.dof_block[poly] * N + i];
          __syncwarp();
          double acc[NA];
#pragma unroll
          for (int k = 0; k < NA; ++k)
            acc[k] = 0.;
          double e0 = 0., e1 = 0.;
          for (int64_t qb = q0; qb < q1; qb += PTS)
            {
              const int64_t q  = qb + pt;
              const bool    ok = q < q1;
              double        x[DIM];
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                x[d] = ok ? A.vq_x[(int64_t)d * A.Q + q] : bb[d];
              const double      w = ok ? A.vq_w[q] : 0.;
              PointTab<DIM, N1> T;
              point_tables<DIM, N1, C::DGP>(A.basis, bb, x, T);
              double ls[NS], dls[NS];
              my_slices<DIM, N1, GZ>(T, part, ls, dls);
              double u = 0., g[DIM];
              if constexpr (MODE != MODE_RHS)
                {
                  eval_slices<DIM, N1, NS, C::DGP>(T, ls, dls, U + part * NA, u, g);
#pragma unroll
                  for (int o = PTS; o < 32; o <<= 1)
                    {
                      u += __shfl_xor_sync(0xffffffffu, u, o);
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        g[d] += __shfl_xor_sync(0xffffffffu, g[d], o);
                    }
                }
              if constexpr (MODE == MODE_APPLY)
                {
                  double f[DIM];
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    f[d] = w * A.stiffness * g[d];
                  integrate_slices<DIM, N1, NS, true, C::DGP>(T, ls, dls, f, MASS ? w * A.mass * u : 0., acc);
                }
              else if constexpr (MODE == MODE_RHS)
                {
                  double f[DIM] = {};
                  integrate_slices<DIM, N1, NS, false, C::DGP>(T, ls, dls, f, ok ? w * A.data0[q] : 0., acc);
                }
              else if (ok && part == 0)
                {
                  const double du = u - A.data0[q];
                  e0 += w * du * du;
                  if (A.data1)
                    {
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        {
                          const double dg = g[d] - A.data1[(int64_t)d * A.Q + q];
                          e1 += w * dg * dg;
                        }
                    }
                }
            }
          if constexpr (MODE == MODE_ERROR)
            {
              const double e[2] = {e0, e1};
              warp_reduce_store<2, 32>(e, A.partial + (int64_t)item * 2, lane);
            }
          else
            warp_reduce_store<NA, PTS>(acc, A.partial + (int64_t)item * N, lane);
        }
    }

    struct FaceArgs
    {
      const double  *fq_x, *fq_n, *fq_w;
      int64_t        Qf;
      int            nqf;
      const double  *bbox;
      const int32_t *ifA, *ifB, *dof_block;
      const double  *sub_sigma;
      const int32_t *item_iface; // work items: runs of face points of one interface
      const int64_t *item_q0, *item_q1;
      int32_t        n_items;
      const double  *x;       // apply
      const double  *gdata;   // rhs: Dirichlet data at every face point (boundary ones are used)
      double        *partial; // [n_items][2][N]
      double         stiffness;
      uint32_t       flags;
      Basis1D        basis;
    };

    // one WARP per work item; 2 GZ lanes per point: side A / side B, each split into GZ slice groups
    template <int DIM, int DEG, int MODE, int GZ>
    __global__ void __launch_bounds__(NW * 32, 2) k_pw_faces(const FaceArgs A)
    {
      using C           = Cfg<DIM, DEG>;
      constexpr int N1  = C::N1, N = C::N;
      constexpr int NS  = N1 / GZ, NA = C::DGP ? N : NS * (N / N1);
      constexpr int PTS = 32 / (2 * GZ);
      __shared__ double Us[NW][2 * N];

      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pt = lane % PTS, grp = lane / PTS;
      const int side = grp / GZ, part = grp % GZ;
      double   *U = Us[warp];
      for (int item = blockIdx.x * NW + warp; item < A.n_items; item += gridDim.x * NW)
        {
          const int  f  = A.item_iface[item];
          const int  pa = A.ifA[f], pb = A.ifB[f];
          const bool interior = pb >= 0;
          double    *out      = A.partial + (int64_t)item * 2 * N;
          bool       skip     = interior ? !(A.flags & PD_ASSEMBLE_INTERIOR) : !(A.flags & PD_ASSEMBLE_BOUNDARY);
          if (MODE == MODE_RHS && interior)
            skip = true;
          if (skip)
            {
              for (int i = lane; i < 2 * N; i += 32)
                out[i] = 0.;
              continue;
            }
          const int64_t q0 = A.item_q0[item], q1 = A.item_q1[item];
          const bool    active = side == 0 || interior;
          const double *bb     = A.bbox + (int64_t)(active && side ? pb : pa) * 2 * DIM;
          __syncwarp();
          if (MODE == MODE_APPLY)
            for (int i = lane; i < 2 * N; i += 32)
              U[i] = i < N ? A.x[(int64_t)A.dof_block[pa] * N + i] :
                             (interior ? A.x[(int64_t)A.dof_block[pb] * N + i - N] : 0.);
          __syncwarp();
          double acc[NA];
#pragma unroll
          for (int k = 0; k < NA; ++k)
            acc[k] = 0.;
          for (int64_t qb = q0; qb < q1; qb += PTS)
            {
              const int64_t q  = qb + pt;
              const bool    ok = q < q1 && active;
              double        x[DIM], n[DIM];
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  x[d] = ok ? A.fq_x[(int64_t)d * A.Qf + q] : bb[d];
                  n[d] = ok ? A.fq_n[(int64_t)d * A.Qf + q] : 0.;
                }
              const double      w  = ok ? A.fq_w[q] * A.stiffness : 0.;
              const double      sg = ok ? A.sub_sigma[q / A.nqf] : 0.;
              PointTab<DIM, N1> T;
              point_tables<DIM, N1, C::DGP>(A.basis, bb, x, T);
              double ls[NS], dls[NS];
              my_slices<DIM, N1, GZ>(T, part, ls, dls);
              double fv[DIM], fm;
              if constexpr (MODE == MODE_APPLY)
                {
                  double u, g[DIM];
                  eval_slices<DIM, N1, NS, C::DGP>(T, ls, dls, U + side * N + part * NA, u, g);
                  double dn = 0.;
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    dn += n[d] * g[d];
#pragma unroll
                  for (int o = PTS; o < PTS * GZ; o <<= 1)
                    {
                      u += __shfl_xor_sync(0xffffffffu, u, o);
                      dn += __shfl_xor_sync(0xffffffffu, dn, o);
                    }
                  const double uo  = __shfl_xor_sync(0xffffffffu, u, PTS * GZ);
                  const double dno = __shfl_xor_sync(0xffffffffu, dn, PTS * GZ);
                  if (interior)
                    {
                      // rows of M11 u_A + M12 u_B (side A) and of M21 u_A + M22 u_B (side B)
                      const double jump = side ? uo - u : u - uo, avg = 0.5 * (dn + dno);
                      fm                = side ? avg - sg * jump : sg * jump - avg;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        fv[d] = -0.5 * jump * n[d];
                    }
                  else
                    {
                      fm = sg * u - dn;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        fv[d] = -u * n[d];
                    }
                }
              else
                {
                  const double gq = ok ? A.gdata[q] : 0.;
                  fm              = sg * gq;
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    fv[d] = -gq * n[d];
                }
              fm *= w;
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                fv[d] *= w;
              integrate_slices<DIM, N1, NS, true, C::DGP>(T, ls, dls, fv, fm, acc);
            }
          // lane groups are (side, part): group g holds the DoFs side * N + part * NA + [0, NA)
          warp_reduce_store<NA, PTS>(acc, out, lane);
        }
    }

    struct GatherArgs
    {
      const int64_t *poly_vitem_ptr;
      const double  *vol_partial;
      const int64_t *padj_ptr, *padj, *iface_item_ptr;
      const double  *face_partial;
      const int32_t *dof_block;
      double        *y;
      int32_t        np_own, n;
      int            with_volume, with_faces, add;
    };

    __global__ void __launch_bounds__(256)
    k_poly_gather(const GatherArgs A)
    {
      const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (int64_t)A.np_own * A.n)
        return;
      const int p = (int)(idx / A.n), i = (int)(idx - (int64_t)p * A.n);
      double    s = 0.;
      if (A.with_volume)
        for (int64_t it = A.poly_vitem_ptr[p]; it < A.poly_vitem_ptr[p + 1]; ++it)
          s += A.vol_partial[it * A.n + i];
      if (A.with_faces)
        for (int64_t k = A.padj_ptr[p]; k < A.padj_ptr[p + 1]; ++k)
          {
            const int64_t e = A.padj[k];
            for (int64_t it = A.iface_item_ptr[e >> 1]; it < A.iface_item_ptr[(e >> 1) + 1]; ++it)
              s += A.face_partial[(it * 2 + (e & 1)) * A.n + i];
          }
      double *yp = A.y + (int64_t)A.dof_block[p] * A.n + i;
      *yp        = A.add ? *yp + s : s;
    }

    // fixed-order sum of the per-item error contributions
    __global__ void __launch_bounds__(256)
    k_sum_pairs(const double *partial, const int32_t n_items, double *out)
    {
      __shared__ double s[2][256];
      double            a = 0., b = 0.;
      for (int i = threadIdx.x; i < n_items; i += 256)
        {
          a += partial[2 * (int64_t)i];
          b += partial[2 * (int64_t)i + 1];
        }
      s[0][threadIdx.x] = a;
      s[1][threadIdx.x] = b;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1)
        {
          if ((int)threadIdx.x < o)
            {
              s[0][threadIdx.x] += s[0][threadIdx.x + o];
              s[1][threadIdx.x] += s[1][threadIdx.x + o];
            }
          __syncthreads();
        }
      if (threadIdx.x == 0)
        {
          out[0] = s[0][0];
          out[1] = s[1][0];
        }
    }

    // Work items of the volume kernels: runs of at most `chunk` quadrature points of one polytope,
    // sized so that every resident warp gets a few items whatever the polytope sizes are.
    void
    ensure_plan(pd_handle *h)
    {
      if (h->pw_plan_valid)
        return;
      const std::vector<int64_t> &sub     = h->h_subcell_ptr;
      const int64_t               n_warps = (int64_t)h->sm_count * 2 * NW;
      const int64_t chunk = std::max<int64_t>(512, ((h->Q / (4 * n_warps) + 31) / 32) * 32);
      std::vector<int32_t> item_poly;
      std::vector<int64_t> q0, q1, poly_item_ptr(h->np + 1, 0);
      for (int32_t p = 0; p < h->np; ++p)
        {
          const int64_t b = sub[p] * h->nqc, e = sub[p + 1] * h->nqc;
          for (int64_t s = b; s < e; s += chunk)
            {
              item_poly.push_back(p);
              q0.push_back(s);
              q1.push_back(std::min(e, s + chunk));
            }
          poly_item_ptr[p + 1] = (int64_t)item_poly.size();
        }
      auto put = [](auto &buf, const auto &v) {
        buf.alloc(v.size());
        if (!v.empty())
          PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
      };
      h->pw_n_items = (int32_t)item_poly.size();
      put(h->pw_item_poly, item_poly);
      put(h->pw_item_q0, q0);
      put(h->pw_item_q1, q1);
      put(h->pw_poly_item_ptr, poly_item_ptr);
      h->mf_vol_partial.alloc((size_t)h->pw_n_items * h->n);
      // interfaces likewise: runs of at most fchunk face points
      const int64_t        fchunk = std::max<int64_t>(128, ((h->Qf / (4 * n_warps) + 15) / 16) * 16);
      std::vector<int32_t> fitem_iface;
      std::vector<int64_t> fq0, fq1, iface_item_ptr(h->n_ifaces + 1, 0);
      for (int32_t f = 0; f < h->n_ifaces; ++f)
        {
          const int64_t b = h->h_if_sub_ptr[f] * h->nqf, e = h->h_if_sub_ptr[f + 1] * h->nqf;
          for (int64_t s = b; s < e; s += fchunk)
            {
              fitem_iface.push_back(f);
              fq0.push_back(s);
              fq1.push_back(std::min(e, s + fchunk));
            }
          iface_item_ptr[f + 1] = (int64_t)fitem_iface.size();
        }
      h->pw_n_fitems = (int32_t)fitem_iface.size();
      put(h->pw_fitem_iface, fitem_iface);
      put(h->pw_fitem_q0, fq0);
      put(h->pw_fitem_q1, fq1);
      put(h->pw_iface_item_ptr, iface_item_ptr);
      h->mf_face_partial.alloc((size_t)h->pw_n_fitems * 2 * h->n);
      h->pw_plan_valid = true;
    }

    VolArgs
    vol_args(pd_handle *h)
    {
      VolArgs a;
      a.vq_x      = h->vq_x.p;
      a.vq_w      = h->vq_w.p;
      a.Q         = h->Q;
      a.bbox      = h->bbox.p;
      a.dof_block = h->dof_block.p;
      a.item_poly = h->pw_item_poly.p;
      a.item_q0   = h->pw_item_q0.p;
      a.item_q1   = h->pw_item_q1.p;
      a.n_items   = h->pw_n_items;
      a.x         = nullptr;
      a.data0     = nullptr;
      a.data1     = nullptr;
      a.partial   = h->mf_vol_partial.p;
      a.stiffness = 1.;
      a.mass      = 0.;
      a.basis     = h->basis;
      return a;
    }

    FaceArgs
    face_args(pd_handle *h)
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
      a.dof_block  = h->dof_block.p;
      a.sub_sigma  = h->sub_sigma.p;
      a.item_iface = h->pw_fitem_iface.p;
      a.item_q0    = h->pw_fitem_q0.p;
      a.item_q1    = h->pw_fitem_q1.p;
      a.n_items    = h->pw_n_fitems;
      a.x          = nullptr;
      a.gdata      = nullptr;
      a.partial    = h->mf_face_partial.p;
      a.stiffness  = 1.;
      a.flags      = PD_ASSEMBLE_ALL;
      a.basis      = h->basis;
      return a;
    }

    void
    gather(pd_handle *h, double *dst, const bool with_volume, const bool with_faces, const bool add)
    {
      GatherArgs g;
      g.poly_vitem_ptr = h->pw_poly_item_ptr.p;
      g.vol_partial    = h->mf_vol_partial.p;
      g.padj_ptr       = h->padj_ptr.p;
      g.padj           = h->padj.p;
      g.iface_item_ptr = h->pw_iface_item_ptr.p;
      g.face_partial   = h->mf_face_partial.p;
      g.dof_block      = h->dof_block.p;
      g.y              = dst;
      g.np_own         = h->np_own;
      g.n              = h->n;
      g.with_volume    = with_volume ? 1 : 0;
      g.with_faces     = with_faces ? 1 : 0;
      g.add            = add ? 1 : 0;
      const int64_t nd = (int64_t)h->np_own * h->n;
      k_poly_gather<<<(unsigned)((nd + 255) / 256), 256, 0, h->stream>>>(g);
      ++h->launches;
    }

    template <int DIM, int DEG>
    constexpr int
    lanes_per_point()
    {
      return !Cfg<DIM, DEG>::DGP && Cfg<DIM, DEG>::N1 % 2 == 0 && Cfg<DIM, DEG>::N > 32 ? 2 : 1;
    }

    template <int DIM, int DEG>
    void
    run_apply(pd_handle *h, const double *src, double *dst, const bool add)
    {
      constexpr int GZ = lanes_per_point<DIM, DEG>();
      ensure_plan(h);
      const bool vol_on = (h->op_flags & PD_ASSEMBLE_VOLUME) != 0 && h->pw_n_items > 0;
      if (vol_on)
        {
          VolArgs a      = vol_args(h);
          a.x            = src;
          a.stiffness    = h->op_coef.stiffness;
          a.mass         = h->op_coef.mass;
          const int grid = (int)std::min<int64_t>((h->pw_n_items + NW - 1) / NW, (int64_t)h->sm_count * 2);
          if (h->op_coef.mass != 0.)
            k_pw_volume<DIM, DEG, MODE_APPLY, true, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          else
            k_pw_volume<DIM, DEG, MODE_APPLY, false, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          ++h->launches;
        }
      if (h->n_ifaces > 0)
        {
          FaceArgs a     = face_args(h);
          a.x            = src;
          a.stiffness    = h->op_coef.stiffness;
          a.flags        = h->op_flags;
          const int grid = (int)std::min<int64_t>((h->pw_n_fitems + NW - 1) / NW, (int64_t)h->sm_count * 2);
          k_pw_faces<DIM, DEG, MODE_APPLY, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          ++h->launches;
        }
      gather(h, dst, vol_on, h->n_ifaces > 0, add);
      PD_CUDA(cudaGetLastError());
    }

    template <int DIM, int DEG>
    void
    run_rhs(pd_handle *h, const double *f_vol, const double *g_face, const double stiffness, double *rhs)
    {
      constexpr int GZ = lanes_per_point<DIM, DEG>();
      ensure_plan(h);
      const bool vol_on = f_vol != nullptr && h->pw_n_items > 0, face_on = g_face != nullptr && h->n_ifaces > 0;
      if (vol_on)
        {
          VolArgs a      = vol_args(h);
          a.data0        = f_vol;
          const int grid = (int)std::min<int64_t>((h->pw_n_items + NW - 1) / NW, (int64_t)h->sm_count * 2);
          k_pw_volume<DIM, DEG, MODE_RHS, false, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          ++h->launches;
        }
      if (face_on)
        {
          FaceArgs a     = face_args(h);
          a.gdata        = g_face;
          a.stiffness    = stiffness;
          a.flags        = PD_ASSEMBLE_BOUNDARY;
          const int grid = (int)std::min<int64_t>((h->pw_n_fitems + NW - 1) / NW, (int64_t)h->sm_count * 2);
          k_pw_faces<DIM, DEG, MODE_RHS, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          ++h->launches;
        }
      gather(h, rhs, vol_on, face_on, false);
      PD_CUDA(cudaGetLastError());
    }

    template <int DIM, int DEG>
    void
    run_error(pd_handle *h, const double *u, const double *exact, const double *exact_grad, double *out2_dev)
    {
      constexpr int GZ = lanes_per_point<DIM, DEG>();
      ensure_plan(h);
      if (h->pw_n_items > 0)
        {
          VolArgs a      = vol_args(h);
          a.x            = u;
          a.data0        = exact;
          a.data1        = exact_grad;
          const int grid = (int)std::min<int64_t>((h->pw_n_items + NW - 1) / NW, (int64_t)h->sm_count * 2);
          k_pw_volume<DIM, DEG, MODE_ERROR, false, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          ++h->launches;
        }
      k_sum_pairs<<<1, 256, 0, h->stream>>>(h->mf_vol_partial.p, h->pw_n_items, out2_dev);
      ++h->launches;
      PD_CUDA(cudaGetLastError());
    }

    // ---- level transfers ------------------------------------------------
    // P evaluates the parent polytope's function at the support points of a child element:
    //   kind 0  child = polytope of a finer agglomeration level, points = the FE_DGQ nodes of
    //           ITS bounding box                     (fill_injection_matrix, include/utils.h:95-270)
    //   kind 1  child = one mesh cell, points = its Q1-mapped FE_DGQ nodes
    //           (fill_interpolation_matrix, include/poly_utils.h:1469-1634)
    // local_matrix(i, j) = phi^parent_j(p_i): rows = child DoFs.  Applied on the fly, one warp
    // per child, lane = support point: prolongation is a pure evaluation (no reduction),
    // restriction (P^T, MGTransferAgglomeration::restrict_and_add, source/multigrid_amg.cc:
    // 92-108) integrates the child's values as point weights and sums a parent's children in
    // list order.
    struct TransferArgs
    {
      int            kind;
      int32_t        n_children;
      const int32_t *parent;     // [child] parent polytope (coarse handle numbering)
      const int32_t *child_blk;  // [child] DoF block in the fine vector
      const double  *child_bbox; // kind 0: [child][2 dim]
      const int32_t *child_cv;   // kind 1: the mesh's cell_verts, [cell][2^dim] vertex ids
      const double  *verts;
      const double  *bbox;       // coarse bounding boxes
      const int32_t *dof_block;  // coarse DoF blocks
      const double  *src;
      double        *dst;     // prolongation: fine vector
      double        *partial; // restriction: [child][N]
      int            add;
      Basis1D        basis;
    };

    template <int DIM, int DEG, bool TRANSPOSE, int GZ>
    __global__ void __launch_bounds__(NW * 32, 2) k_pw_transfer(const TransferArgs A)
    {
      using C           = Cfg<DIM, DEG>;
      constexpr int N1  = C::N1, N = C::N;
      constexpr int NS  = N1 / GZ, NA = C::DGP ? N : NS * (N / N1);
      constexpr int PTS = 32 / GZ;
      __shared__ double Us[NW][N];

      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pt = lane % PTS, part = lane / PTS;
      double   *U = Us[warp];
      for (int c = blockIdx.x * NW + warp; c < A.n_children; c += gridDim.x * NW)
        {
          const int     K  = A.parent[c];
          const double *bb = A.bbox + (int64_t)K * 2 * DIM;
          __syncwarp();
          if (!TRANSPOSE)
            for (int i = lane; i < N; i += 32)
              U[i] = A.src[(int64_t)A.dof_block[K] * N + i];
          __syncwarp();
          double acc[NA];
#pragma unroll
          for (int k = 0; k < NA; ++k)
            acc[k] = 0.;
          for (int ib = 0; ib < N; ib += PTS)
            {
              const int  i  = ib + pt;
              const bool ok = i < N;
              // support point i of the child
              double xi[DIM], x[DIM];
              {
                int r = ok ? i : 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                  {
                    const int a = r % N1;
                    r /= N1;
                    xi[d] = A.basis.node[0];
#pragma unroll
                    for (int t = 1; t < N1; ++t)
                      if (a == t)
                        xi[d] = A.basis.node[t];
                  }
              }
              if (A.kind == 0)
                {
                  const double *cb = A.child_bbox + (int64_t)c * 2 * DIM;
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    x[d] = cb[d] + xi[d] * (cb[DIM + d] - cb[d]);
                }
              else
                {
                  const int32_t *cv = A.child_cv + (int64_t)A.child_blk[c] * (1 << DIM); // the child IS mesh cell child_blk[c]
#pragma unroll
                  for (int d = 0; d < DIM; ++d)
                    x[d] = 0.;
#pragma unroll
                  for (int v = 0; v < (1 << DIM); ++v)
                    {
                      double w = 1.;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        w *= ((v >> d) & 1) ? xi[d] : 1. - xi[d];
                      const double *X = A.verts + (int64_t)cv[v] * DIM;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        x[d] += w * X[d];
                    }
                }
              PointTab<DIM, N1> T;
              point_tables<DIM, N1, C::DGP>(A.basis, bb, x, T);
              double ls[NS], dls[NS];
              my_slices<DIM, N1, GZ>(T, part, ls, dls);
              if constexpr (!TRANSPOSE)
                {
                  double u, g[DIM];
                  eval_slices<DIM, N1, NS, C::DGP>(T, ls, dls, U + part * NA, u, g);
#pragma unroll
                  for (int o = PTS; o < 32; o <<= 1)
                    u += __shfl_xor_sync(0xffffffffu, u, o);
                  if (ok && part == 0)
                    {
                      double *yp = A.dst + (int64_t)A.child_blk[c] * N + i;
                      *yp        = A.add ? *yp + u : u;
                    }
                }
              else
                {
                  const double yv    = ok ? A.src[(int64_t)A.child_blk[c] * N + i] : 0.;
                  double       f[DIM] = {};
                  integrate_slices<DIM, N1, NS, false, C::DGP>(T, ls, dls, f, yv, acc);
                }
            }
          if constexpr (TRANSPOSE)
            warp_reduce_store<NA, PTS>(acc, A.partial + (int64_t)c * N, lane);
        }
    }

    struct TransferGatherArgs
    {
      const int64_t *pc_ptr; // children of every coarse polytope
      const int32_t *pc_idx;
      const double  *partial;
      const int32_t *dof_block;
      double        *y;
      int32_t        np, n;
      int            add;
    };

    __global__ void __launch_bounds__(256)
    k_transfer_gather(const TransferGatherArgs A)
    {
      const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= (int64_t)A.np * A.n)
        return;
      const int p = (int)(idx / A.n), i = (int)(idx - (int64_t)p * A.n);
      double    s = 0.;
      for (int64_t k = A.pc_ptr[p]; k < A.pc_ptr[p + 1]; ++k)
        s += A.partial[(int64_t)A.pc_idx[k] * A.n + i];
      double *yp = A.y + (int64_t)A.dof_block[p] * A.n + i;
      *yp        = A.add ? *yp + s : s;
    }

    template <int DIM, int DEG>
    void
    run_transfer(pd_handle *h, const pd_transfer &t, const bool transpose, const double *src, double *dst, const bool add)
    {
      constexpr int GZ = lanes_per_point<DIM, DEG>();
      TransferArgs  a;
      a.kind       = t.kind;
      a.n_children = t.n_children;
      a.parent     = t.parent.p;
      a.child_blk  = t.child_blk.p;
      a.child_bbox = t.child_bbox;
      a.child_cv   = h->cell_verts.p;
      a.verts      = h->verts.p;
      a.bbox       = h->bbox.p;
      a.dof_block  = h->dof_block.p;
      a.src        = src;
      a.dst        = dst;
      a.partial    = t.partial.p;
      a.add        = add ? 1 : 0;
      a.basis      = h->basis;
      const int grid = (int)std::min<int64_t>((t.n_children + NW - 1) / NW, (int64_t)h->sm_count * 2);
      if (t.n_children > 0)
        {
          if (transpose)
            k_pw_transfer<DIM, DEG, true, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          else
            k_pw_transfer<DIM, DEG, false, GZ><<<grid, NW * 32, 0, h->stream>>>(a);
          ++h->launches;
        }
      if (transpose)
        {
          TransferGatherArgs g;
          g.pc_ptr    = t.pc_ptr.p;
          g.pc_idx    = t.pc_idx.p;
          g.partial   = t.partial.p;
          g.dof_block = h->dof_block.p;
          g.y         = dst;
          g.np        = h->np_own;
          g.n         = h->n;
          g.add       = add ? 1 : 0;
          const int64_t nd = (int64_t)h->np_own * h->n;
          k_transfer_gather<<<(unsigned)((nd + 255) / 256), 256, 0, h->stream>>>(g);
          ++h->launches;
        }
      PD_CUDA(cudaGetLastError());
    }

#define PD_DISPATCH(FN, ...)                                                                                      \
  switch (h->dim * 10 + h->degree)                                                                                 \
    {                                                                                                              \
      case 21: FN<2, 1>(__VA_ARGS__); break;                                                                       \
      case 22: FN<2, 2>(__VA_ARGS__); break;                                                                       \
      case 23: FN<2, 3>(__VA_ARGS__); break;                                                                       \
      case 24: FN<2, 4>(__VA_ARGS__); break;                                                                       \
      case 31: FN<3, 1>(__VA_ARGS__); break;                                                                       \
      case 32: FN<3, 2>(__VA_ARGS__); break;                                                                       \
      case 33: FN<3, 3>(__VA_ARGS__); break;                                                                       \
      default:                                                                                                     \
        throw CudaError{cudaErrorNotSupported, "no point-wise polytopal kernel for this (dim, degree)", __LINE__}; \
    }
  } // namespace

#define PD_DISPATCH_FE(FN, ...)                                                                                    \
  if (h->fe_kind == PD_FE_AGGLODGP)                                                                                \
    switch (h->dim * 10 + h->degree)                                                                               \
      {                                                                                                            \
        case 21: FN<2, DGP_BASE + 1>(__VA_ARGS__); break;                                                          \
        case 22: FN<2, DGP_BASE + 2>(__VA_ARGS__); break;                                                          \
        case 23: FN<2, DGP_BASE + 3>(__VA_ARGS__); break;                                                          \
        case 24: FN<2, DGP_BASE + 4>(__VA_ARGS__); break;                                                          \
        case 31: FN<3, DGP_BASE + 1>(__VA_ARGS__); break;                                                          \
        case 32: FN<3, DGP_BASE + 2>(__VA_ARGS__); break;                                                          \
        case 33: FN<3, DGP_BASE + 3>(__VA_ARGS__); break;                                                          \
        default:                                                                                                   \
          throw CudaError{cudaErrorNotSupported, "no point-wise FE_AggloDGP kernel for this (dim, degree)", __LINE__}; \
      }                                                                                                            \
  else                                                                                                             \
    PD_DISPATCH(FN, __VA_ARGS__)

  void
  launch_poly_apply(pd_handle *h, const double *src, double *dst, const bool add)
  {
    PD_DISPATCH_FE(run_apply, h, src, dst, add);
  }

  void
  launch_poly_rhs(pd_handle *h, const double *f_vol, const double *g_face, const double stiffness, double *rhs)
  {
    PD_DISPATCH_FE(run_rhs, h, f_vol, g_face, stiffness, rhs);
  }

  void
  launch_transfer(pd_handle *h, const pd_transfer &t, const bool transpose, const double *src, double *dst, const bool add)
  {
    PD_DISPATCH(run_transfer, h, t, transpose, src, dst, add);
  }

  void
  launch_poly_error(pd_handle *h, const double *u, const double *exact, const double *exact_grad, double *out2_dev)
  {
    PD_DISPATCH_FE(run_error, h, u, exact, exact_grad, out2_dev);
  }
} // namespace pd
