// -----------------------------------------------------------------------------
// pd_mappedfine.cu -- matrix-free SIP operator apply on a fine (non-agglomerated)
// mesh of GENERAL (Q1-mapped, distorted) hexahedra / quadrilaterals with the standard
// mapped FE_DGQ(p) basis, phi_i(x) = phihat_i(F_K^{-1}(x)).
//
// Reference semantics:
//   Utils::MatrixFreeOperators::LaplaceOperatorDG::{local_apply, local_apply_face,
//     local_apply_boundary}                              include/utils.h:819-925
//   MonodomainOperatorDG (f M + sigma K, no boundary)     include/utils.h:1565-1659
//   matrix-based twin of the same operator                examples/monodomain_DG3D.cc:1374-1622
// with n_q_points_1d = p+1 and the face penalty
//   sigma_F = max(p,1)(p+1) (|n . J_m^{-1}|_normal + |n . J_p^{-1}|_normal)  at face point 0,
//   boundary 2 * 2 * max(p,1)(p+1) |n . J^{-1}|_normal.
// On Cartesian cells this is the operator of pd_finemesh.cu; there the geometry collapses
// to five scalars per (cell, direction), here it is per quadrature point:
//   cell point : G = w |J| J^{-1} J^{-T} (symmetric) and m = w |J|
//   face point : W = w |J| |J^{-T} e_d| (surface element), c- = J_m^{-1} n, c+ = J_p^{-1} n
// precomputed once per mesh by k_mapped_geometry (as MatrixFree's MappingInfo does) and
// streamed: (d(d+1)/2 + 1) * 8 B per cell point + (2d+1) * 8 B per face point.
//
// The kernel works in the Lagrange basis ON the Gauss points (collocation): the vector is
// changed to that basis once per apply (k_basis_change, V (x) V (x) V), where the mass is
// diagonal, gradients are the collocation derivative D~ along lines, and face traces are
// 1-D extrapolations e^_s / d^_s along the normal line.  Cell-centric, own rows only (every
// interior face is evaluated from both sides): no atomics, deterministic.
//
// One thread per LINE (d, j) as in pd_finemesh.cu.  Neighbouring cells must be in standard
// orientation (opposite local face, aligned tangential axes) -- checked on the vertex ids
// at set-up; meshes that violate it get "not available", never a wrong answer.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace pd
{
  namespace
  {
    constexpr int
    ipow_(const int b, const int e)
    {
      return e == 0 ? 1 : b * ipow_(b, e - 1);
    }
    constexpr int
    pow2_at_least(const int v)
    {
      int g = 4;
      while (g < v)
        g *= 2;
      return g;
    }

    // 1-D tables in the Gauss-point Lagrange basis l~_q (kernel parameters; compile-time
    // indices become constant-bank operands)
    template <int N1>
    struct MappedTables
    {
      double Dt[N1 * N1]; // Dt[q][r] = l~_r'(x_q)        collocation derivative
      double e[2][N1];    // e[s][q]  = l~_q(s)           trace at the face s = 0 / 1
      double d[2][N1];    // d[s][q]  = l~_q'(s)
      double Vt[N1 * N1]; // Vt[i][q] = l_i(x_q): back to the nodal test functions
    };

    template <int N1>
    struct MappedArgs
    {
      MappedTables<N1> T;
      const double    *dt_rows; // the same Dt in global memory (rows picked by a per-thread index)
      const double    *cgeo;    // [cell][NG + 1][N]
      const double    *fgeo;    // [cell][DIM][2][1 + 2 DIM][NF]
      const double    *sigma;   // [cell][2 DIM]
      const int32_t   *nbr;     // [cell][2 DIM]
      const double    *zero;
      const double    *xg;
      double          *y; // nodal result (the basis change back is fused into this kernel)
      int              add;
      int32_t          n_cells;
      double           stiffness, mass;
      uint32_t         flags;
    };

    template <int DIM, int DEG, int MINB>
    __global__ void __launch_bounds__(256, MINB) k_mapped_sip(const __grid_constant__ MappedArgs<DEG + 1> A)
    {
      constexpr int N1  = DEG + 1;
      constexpr int N   = ipow_(N1, DIM);
      constexpr int NF  = N / N1;   // lines per direction = points per face
      constexpr int NT  = DIM * NF; // line tasks per cell
      constexpr int GS  = pow2_at_least(N > NT ? N : NT);
      constexpr int CPB = 256 / GS;
      constexpr int NG  = DIM * (DIM + 1) / 2;
      constexpr int NFD = 1 + 2 * DIM;
      constexpr int TD  = DIM - 1;
      constexpr int NP  = N | 1;

      __shared__ double sDt[N1 * N1];
      __shared__ double sG[CPB][DIM][NP];         // reference gradient components; later the work arrays W_d
      __shared__ double sT[CPB][DIM][2][2][NF];   // face value traces: own, neighbour
      __shared__ double sA[CPB][DIM][2][TD][NF];  // -[u] W c-_t : tangential-derivative tests

      for (int i = threadIdx.x; i < N1 * N1; i += blockDim.x)
        sDt[i] = A.dt_rows[i];
      __syncthreads();

      const int slot = threadIdx.x / GS, l = threadIdx.x % GS;
      auto      group_sync = [slot] {
        if constexpr (GS <= 32)
          __syncwarp();
        else
          asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(GS) : "memory");
      };
      const bool task_ok = l < NT;
      const int  d = task_ok ? l / NF : 0, j = l % NF;
      int        off[N1];
      {
        const int stride = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
        const int base   = d == 0 ? j * N1 : (d == 1 ? (j % N1) + (j / N1) * N1 * N1 : j);
#pragma unroll
        for (int t = 0; t < N1; ++t)
          off[t] = base + t * stride;
      }
      // transverse directions (increasing) and the position of this line in them
      int tdim[TD > 0 ? TD : 1], jt[TD > 0 ? TD : 1], tstr[TD > 0 ? TD : 1];
#pragma unroll
      for (int a = 0; a < TD; ++a)
        {
          tdim[a] = a < d ? a : a + 1;
          jt[a]   = a == 0 ? j % N1 : j / N1;
          tstr[a] = a == 0 ? 1 : N1; // stride of transverse direction a inside a face array
        }
      int gidx[DIM]; // row d of the symmetric G
#pragma unroll
      for (int b = 0; b < DIM; ++b)
        {
          const int lo = d < b ? d : b, hi = d < b ? b : d;
          gidx[b]      = lo * DIM - lo * (lo - 1) / 2 + (hi - lo);
        }
      const bool vol_on = (A.flags & PD_ASSEMBLE_VOLUME) != 0, int_on = (A.flags & PD_ASSEMBLE_INTERIOR) != 0,
                 bnd_on = (A.flags & PD_ASSEMBLE_BOUNDARY) != 0;
      const double mass = vol_on ? A.mass : 0.;

      for (int c0 = blockIdx.x * CPB; c0 < A.n_cells; c0 += gridDim.x * CPB)
        {
          const int  cell    = c0 + slot;
          const bool cell_ok = cell < A.n_cells;
          const bool work    = cell_ok && task_ok;
          double     u[N1], tu[2] = {0., 0.}, du[2] = {0., 0.}, tn[2] = {0., 0.}, dn[2] = {0., 0.};
          int        nb[2] = {-1, -1};
          group_sync(); // the previous cell's final sum has read this slot's arrays
          if (work)
            {
              nb[0]             = A.nbr[(int64_t)cell * 2 * DIM + 2 * d];
              nb[1]             = A.nbr[(int64_t)cell * 2 * DIM + 2 * d + 1];
              const double *xc  = A.xg + (int64_t)cell * N;
              const double *xn0 = nb[0] >= 0 ? A.xg + (int64_t)nb[0] * N : A.zero;
              const double *xn1 = nb[1] >= 0 ? A.xg + (int64_t)nb[1] * N : A.zero;
              double        nv[2][N1];
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
                  u[t]     = xc[off[t]];
                  nv[0][t] = xn0[off[t]];
                  nv[1][t] = xn1[off[t]];
                }
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
                  double g = 0.;
#pragma unroll
                  for (int r = 0; r < N1; ++r)
                    g += A.T.Dt[t * N1 + r] * u[r];
                  sG[slot][d][off[t]] = g;
#pragma unroll
                  for (int s = 0; s < 2; ++s)
                    {
                      tu[s] += A.T.e[s][t] * u[t];
                      du[s] += A.T.d[s][t] * u[t];
                      tn[s] += A.T.e[1 - s][t] * nv[s][t];
                      dn[s] += A.T.d[1 - s][t] * nv[s][t];
                    }
                }
#pragma unroll
              for (int s = 0; s < 2; ++s)
                {
                  sT[slot][d][s][0][j] = tu[s];
                  sT[slot][d][s][1][j] = tn[s];
                }
            }
          group_sync();
          double out[N1], Av[2] = {0., 0.}, Bn[2] = {0., 0.};
#pragma unroll
          for (int i = 0; i < N1; ++i)
            out[i] = 0.;
          if (work)
            {
              // ---- cell term: flux_d = sum_b G[d][b] dhat_b u at the points of this line, then D~^T
              if (vol_on)
                {
                  const double *cg = A.cgeo + (int64_t)cell * (NG + 1) * N;
                  double        fl[N1];
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    {
                      double f = 0.;
#pragma unroll
                      for (int b = 0; b < DIM; ++b)
                        f += cg[gidx[b] * N + off[t]] * sG[slot][b][off[t]];
                      fl[t] = f;
                    }
#pragma unroll
                  for (int i = 0; i < N1; ++i)
#pragma unroll
                    for (int t = 0; t < N1; ++t)
                      out[i] += A.T.Dt[t * N1 + i] * fl[t];
                }
              // ---- the two faces of this line's direction, at this line's transverse point
#pragma unroll
              for (int s = 0; s < 2; ++s)
                {
                  const bool inner = nb[s] >= 0;
                  const bool on    = inner ? int_on : bnd_on;
                  double     B     = 0.;
                  double     cmt[TD > 0 ? TD : 1];
#pragma unroll
                  for (int a = 0; a < TD; ++a)
                    cmt[a] = 0.;
                  if (on)
                    {
                      const double *fg  = A.fgeo + (((int64_t)cell * DIM + d) * 2 + s) * NFD * NF + j;
                      const double  W   = fg[0];
                      const double  sg  = A.sigma[(int64_t)cell * 2 * DIM + 2 * d + s];
                      double        dnm = fg[(1 + d) * NF] * du[s], dnp = fg[(1 + DIM + d) * NF] * dn[s];
#pragma unroll
                      for (int a = 0; a < TD; ++a)
                        {
                          // tangential reference derivatives of the two traces (collocation)
                          const double *row = sDt + jt[a] * N1;
                          const double *tl  = &sT[slot][d][s][0][j - jt[a] * tstr[a]];
                          double        dtu = 0., dtn = 0.;
#pragma unroll
                          for (int k = 0; k < N1; ++k)
                            {
                              dtu += row[k] * tl[k * tstr[a]];
                              dtn += row[k] * tl[NF + k * tstr[a]];
                            }
                          cmt[a] = fg[(1 + tdim[a]) * NF];
                          dnm += cmt[a] * dtu;
                          dnp += fg[(1 + DIM + tdim[a]) * NF] * dtn;
                        }
                      // interior: j = (u- - u+)/2, a = 2 sigma j - (dn u- + dn u+)/2;  boundary: j = u-, a = sigma_b u- - dn u-
                      const double jv = inner ? 0.5 * (tu[s] - tn[s]) : tu[s];
                      const double av = inner ? 2. * sg * jv - 0.5 * (dnm + dnp) : sg * jv - dnm;
                      Av[s]           = av * W;
                      B               = -jv * W;
                      Bn[s]           = B * fg[(1 + d) * NF];
                    }
#pragma unroll
                  for (int a = 0; a < TD; ++a)
                    sA[slot][d][s][a][j] = B * cmt[a];
                }
            }
          group_sync();
          if (work)
            {
#pragma unroll
              for (int s = 0; s < 2; ++s)
                {
                  double Fv = Av[s];
#pragma unroll
                  for (int a = 0; a < TD; ++a)
                    {
                      const double *col = sDt + jt[a];
                      const double *al  = &sA[slot][d][s][a][j - jt[a] * tstr[a]];
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        Fv += col[k * N1] * al[k * tstr[a]];
                    }
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                    out[i] += A.T.e[s][i] * Fv + A.T.d[s][i] * Bn[s];
                }
#pragma unroll
              for (int i = 0; i < N1; ++i)
                sG[slot][d][off[i]] = out[i];
            }
          group_sync();
          if (cell_ok && l < N)
            {
              double acc = 0.;
#pragma unroll
              for (int b = 0; b < DIM; ++b)
                acc += sG[slot][b][l];
              acc *= A.stiffness;
              if (mass != 0.)
                acc += mass * A.cgeo[(int64_t)cell * (NG + 1) * N + NG * N + l] * A.xg[(int64_t)cell * N + l];
              sT[slot][0][0][0][l] = acc; // the face-trace arrays are free now: N <= DIM * 4 * NF doubles
            }
          group_sync();
          // ---- back to the nodal basis: y = (V^T (x) ... (x) V^T) y~, one line per thread and direction pass
          double *yq = &sT[slot][0][0][0][0];
#pragma unroll
          for (int pass = 0; pass < DIM; ++pass)
            {
              if (cell_ok && l < NF)
                {
                  const int stride = pass == 0 ? 1 : (pass == 1 ? N1 : N1 * N1);
                  const int base   = pass == 0 ? l * N1 : (pass == 1 ? (l % N1) + (l / N1) * N1 * N1 : l);
                  double    v[N1];
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    v[t] = yq[base + t * stride];
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                    {
                      double sm = 0.;
#pragma unroll
                      for (int t = 0; t < N1; ++t)
                        sm += A.T.Vt[i * N1 + t] * v[t];
                      yq[base + i * stride] = sm;
                    }
                }
              group_sync();
            }
          if (cell_ok && l < N)
            {
              double *yp = A.y + (int64_t)cell * N + l;
              *yp        = A.add ? *yp + yq[l] : yq[l];
            }
        }
    }

    // dst = (M (x) ... (x) M) src per cell, M = V (to the Gauss basis) or V^T (back, tested
    // residual), one thread per line and direction pass; `add` accumulates into dst.
    template <int N1>
    struct ChangeArgs
    {
      double        M[N1 * N1]; // M[i][t]: out_i = sum_t M[i][t] in_t
      const double *src;
      double       *dst;
      int32_t       n_cells;
      int           add;
    };

    template <int DIM, int DEG>
    __global__ void __launch_bounds__(256) k_basis_change(const __grid_constant__ ChangeArgs<DEG + 1> A)
    {
      constexpr int N1 = DEG + 1, N = ipow_(N1, DIM), NL = N / N1;
      constexpr int GS  = pow2_at_least(N);
      constexpr int CPB = 256 / GS;
      __shared__ double sW[CPB][N | 1];
      const int slot = threadIdx.x / GS, l = threadIdx.x % GS;
      auto      group_sync = [slot] {
        if constexpr (GS <= 32)
          __syncwarp();
        else
          asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(GS) : "memory");
      };
      for (int c0 = blockIdx.x * CPB; c0 < A.n_cells; c0 += gridDim.x * CPB)
        {
          const int  cell = c0 + slot;
          const bool ok   = cell < A.n_cells;
          group_sync();
          if (ok && l < N)
            sW[slot][l] = A.src[(int64_t)cell * N + l];
          group_sync();
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              if (ok && l < NL)
                {
                  const int stride = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
                  const int base   = d == 0 ? l * N1 : (d == 1 ? (l % N1) + (l / N1) * N1 * N1 : l);
                  double    v[N1];
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    v[t] = sW[slot][base + t * stride];
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                    {
                      double s = 0.;
#pragma unroll
                      for (int t = 0; t < N1; ++t)
                        s += A.M[i * N1 + t] * v[t];
                      sW[slot][base + i * stride] = s;
                    }
                }
              group_sync();
            }
          if (ok && l < N)
            {
              double *yp = A.dst + (int64_t)cell * N + l;
              *yp        = A.add ? *yp + sW[slot][l] : sW[slot][l];
            }
        }
    }

    // ---- geometry, once per mesh -------------------------------------------------------
    struct GeoArgs
    {
      const double  *verts;  // [n_verts][dim]
      const int32_t *cellv;  // [cell][2^dim] vertex ids, block order
      const int32_t *nbr;    // [cell][2 dim]
      double         gx[8], gw[8];
      int            dim, n1;
      int32_t        n_cells;
      double         pc; // max(p,1)(p+1)
      double        *cgeo, *fgeo, *sigma;
    };

    template <int DIM>
    __device__ void
    q1_inverse_jacobian(const double *verts, const int32_t *cv, const double *xi, double Ji[DIM][DIM], double &det)
    {
      double J[DIM][DIM];
      for (int a = 0; a < DIM; ++a)
        for (int b = 0; b < DIM; ++b)
          J[a][b] = 0.;
      for (int v = 0; v < (1 << DIM); ++v)
        {
          double f[DIM], df[DIM];
          for (int k = 0; k < DIM; ++k)
            {
              const int bit = (v >> k) & 1;
              f[k]          = bit ? xi[k] : 1. - xi[k];
              df[k]         = bit ? 1. : -1.;
            }
          const double *X = verts + (int64_t)cv[v] * DIM;
          for (int b = 0; b < DIM; ++b)
            {
              double g = df[b];
              for (int k = 0; k < DIM; ++k)
                if (k != b)
                  g *= f[k];
              for (int a = 0; a < DIM; ++a)
                J[a][b] += X[a] * g;
            }
        }
      if constexpr (DIM == 2)
        {
          det      = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          Ji[0][0] = J[1][1] / det;
          Ji[0][1] = -J[0][1] / det;
          Ji[1][0] = -J[1][0] / det;
          Ji[1][1] = J[0][0] / det;
        }
      else
        {
          const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                       c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
          det              = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
          Ji[0][0]         = c00 / det;
          Ji[1][0]         = c01 / det;
          Ji[2][0]         = c02 / det;
          Ji[0][1]         = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
          Ji[1][1]         = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
          Ji[2][1]         = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
          Ji[0][2]         = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
          Ji[1][2]         = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
          Ji[2][2]         = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
        }
    }

    template <int DIM>
    __global__ void
    k_mapped_geometry(const GeoArgs A)
    {
      const int     n1 = A.n1, N = DIM == 2 ? n1 * n1 : n1 * n1 * n1, NF = N / n1;
      constexpr int NG = DIM * (DIM + 1) / 2, NFD = 1 + 2 * DIM;
      const int     per_cell = N + 2 * DIM * NF;
      const int64_t gid      = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (gid >= (int64_t)A.n_cells * per_cell)
        return;
      const int32_t  cell = (int32_t)(gid / per_cell);
      const int      item = (int)(gid % per_cell);
      const int32_t *cv   = A.cellv + (int64_t)cell * (1 << DIM);
      double         xi[DIM], Ji[DIM][DIM], det;
      if (item < N)
        {
          double w = 1.;
          int    r = item;
          for (int k = 0; k < DIM; ++k)
            {
              xi[k] = A.gx[r % n1];
              w *= A.gw[r % n1];
              r /= n1;
            }
          q1_inverse_jacobian<DIM>(A.verts, cv, xi, Ji, det);
          double   *cg = A.cgeo + (int64_t)cell * (NG + 1) * N + item;
          const double m = w * fabs(det);
          int          g = 0;
          for (int a = 0; a < DIM; ++a)
            for (int b = a; b < DIM; ++b, ++g)
              {
                double s = 0.;
                for (int i = 0; i < DIM; ++i)
                  s += Ji[a][i] * Ji[b][i];
                cg[g * N] = m * s;
              }
          cg[NG * N] = m;
          return;
        }
      // face point: direction d, side s, transverse point j (transverse directions increasing, first fastest)
      const int fi = item - N, f = fi / NF, j = fi % NF, d = f / 2, s = f % 2;
      double    w = 1.;
      {
        int r = j;
        for (int k = 0; k < DIM; ++k)
          if (k != d)
            {
              xi[k] = A.gx[r % n1];
              w *= A.gw[r % n1];
              r /= n1;
            }
      }
      xi[d] = s ? 1. : 0.;
      q1_inverse_jacobian<DIM>(A.verts, cv, xi, Ji, det);
      double len = 0.;
      for (int i = 0; i < DIM; ++i)
        len += Ji[d][i] * Ji[d][i];
      len = sqrt(len);
      double       nrm[DIM];
      const double sn = (s ? 1. : -1.) * (det > 0. ? 1. : -1.);
      for (int i = 0; i < DIM; ++i)
        nrm[i] = sn * Ji[d][i] / len;
      double *fg = A.fgeo + (((int64_t)cell * DIM + d) * 2 + s) * NFD * NF + j;
      fg[0]      = w * fabs(det) * len;
      double cmn = 0.;
      for (int k = 0; k < DIM; ++k)
        {
          double c = 0.;
          for (int i = 0; i < DIM; ++i)
            c += Ji[k][i] * nrm[i];
          fg[(1 + k) * NF] = c;
          if (k == d)
            cmn = c;
        }
      const int32_t nb  = A.nbr[(int64_t)cell * 2 * DIM + f];
      double        cpn = 0.;
      if (nb >= 0)
        {
          double Jn[DIM][DIM], detn;
          xi[d] = s ? 0. : 1.;
          q1_inverse_jacobian<DIM>(A.verts, A.cellv + (int64_t)nb * (1 << DIM), xi, Jn, detn);
          for (int k = 0; k < DIM; ++k)
            {
              double c = 0.;
              for (int i = 0; i < DIM; ++i)
                c += Jn[k][i] * nrm[i];
              fg[(1 + DIM + k) * NF] = c;
              if (k == d)
                cpn = c;
            }
        }
      else
        for (int k = 0; k < DIM; ++k)
          fg[(1 + DIM + k) * NF] = 0.;
      if (j == 0) // the reference evaluates the penalty at face quadrature point 0
        A.sigma[(int64_t)cell * 2 * DIM + f] = nb >= 0 ? A.pc * (fabs(cmn) + fabs(cpn)) : 4. * A.pc * fabs(cmn);
    }

    // host: Lagrange basis on `nodes` at x
    void
    lagrange_on(const std::vector<double> &nodes, const double x, double *L, double *dL)
    {
      const int m = (int)nodes.size();
      for (int a = 0; a < m; ++a)
        {
          double val = 1., der = 0., den = 1.;
          for (int b = 0; b < m; ++b)
            if (b != a)
              {
                const double t = x - nodes[b];
                der            = der * t + val;
                val            = val * t;
                den *= nodes[a] - nodes[b];
              }
          L[a]  = val / den;
          dL[a] = der / den;
        }
    }

    template <int DIM, int DEG, int MINB>
    void
    launch_mapped(pd_handle *h, const double *src, double *dst, const bool add)
    {
      constexpr int N1 = DEG + 1, N = ipow_(N1, DIM), NT = DIM * (N / N1);
      const int     sm = h->sm_count;
      // 1. to the Gauss basis (owned cells; this mode has no ghosts)
      {
        constexpr int   GS = pow2_at_least(N), CPB = 256 / GS;
        ChangeArgs<N1>  c;
        std::memcpy(c.M, h->mp_tab_host.data() + 0, sizeof(c.M)); // V
        c.src     = src;
        c.dst     = h->mp_xg.p;
        c.n_cells = h->np_own;
        c.add     = 0;
        const int64_t want = ((int64_t)h->np_own + CPB - 1) / CPB;
        k_basis_change<DIM, DEG><<<(int)std::min<int64_t>(want, (int64_t)sm * 32), 256, 0, h->stream>>>(c);
      }
      // 2. the operator in the Gauss basis
      {
        constexpr int  GS = pow2_at_least(N > NT ? N : NT), CPB = 256 / GS;
        MappedArgs<N1> a;
        static_assert(sizeof(a.T) == (2 * N1 * N1 + 4 * N1) * sizeof(double), "table layout");
        std::memcpy(&a.T, h->mp_tab_host.data() + 2 * N1 * N1, (N1 * N1 + 4 * N1) * sizeof(double)); // Dt | e | d
        std::memcpy(a.T.Vt, h->mp_tab_host.data() + N1 * N1, N1 * N1 * sizeof(double));
        a.dt_rows   = h->mp_dt.p;
        a.cgeo      = h->mp_cgeo.p;
        a.fgeo      = h->mp_fgeo.p;
        a.sigma     = h->mp_sigma.p;
        a.nbr       = h->mp_nbr.p;
        a.zero      = h->mp_zero.p;
        a.xg        = h->mp_xg.p;
        a.y         = dst;
        a.add       = add ? 1 : 0;
        a.n_cells   = h->np_own;
        a.stiffness = h->op_coef.stiffness;
        a.mass      = h->op_coef.mass;
        a.flags     = h->op_flags;
        const int64_t want = ((int64_t)h->np_own + CPB - 1) / CPB;
        k_mapped_sip<DIM, DEG, MINB><<<(int)std::min<int64_t>(want, (int64_t)sm * 4 * MINB), 256, 0, h->stream>>>(a);
      }
      h->launches += 2;
    }
  } // namespace

  // Recognise "every polytope is one cell, no ghosts, neighbours in standard orientation,
  // QGauss(p+1) on cells and faces" and stage the topology; geometry is computed on first use.
  void
  setup_mapped_operator(pd_handle *h, const pd_mesh_desc &d)
  {
    h->mp_ready     = false;
    h->mp_geo_valid = false;
    const int dim = d.dim, vpc = 1 << dim, nfc = 2 * dim, n1 = h->n1;
    if (h->n_subcells != h->np_own || h->np != h->np_own || h->nq1 != n1 || h->nq1f != n1)
      return;
    if (!((dim == 2 && h->degree >= 1 && h->degree <= 4) || (dim == 3 && h->degree >= 1 && h->degree <= 3)))
      return;
    std::vector<int32_t> cellv((size_t)h->np * vpc), nbr((size_t)h->np * nfc, -2);
    for (int32_t p = 0; p < h->np; ++p)
      {
        const int32_t c = d.poly_subcell_idx[d.poly_subcell_ptr[p]];
        for (int v = 0; v < vpc; ++v)
          cellv[(size_t)d.dof_block[p] * vpc + v] = d.cell_verts[(size_t)c * vpc + v];
      }
    for (int32_t f = 0; f < d.n_ifaces; ++f)
      {
        const int32_t a = d.iface_polyA[f], b = d.iface_polyB[f];
        for (int64_t s = d.iface_sub_ptr[f]; s < d.iface_sub_ptr[f + 1]; ++s)
          {
            const int     lf = d.sub_face[s];
            const int32_t ba = d.dof_block[a];
            nbr[(size_t)ba * nfc + lf] = b >= 0 ? d.dof_block[b] : -1;
            if (b >= 0)
              nbr[(size_t)d.dof_block[b] * nfc + (lf ^ 1)] = ba;
          }
      }
    for (int32_t c = 0; c < h->np; ++c)
      for (int f = 0; f < nfc; ++f)
        {
          const int32_t nb = nbr[(size_t)c * nfc + f];
          if (nb == -2)
            return; // a cell face without an interface entry
          if (nb < 0)
            continue;
          // standard orientation: the neighbour's opposite face carries the same vertices
          const int dd = f / 2, s = f % 2;
          for (int v = 0; v < vpc; ++v)
            if (((v >> dd) & 1) == s && cellv[(size_t)c * vpc + v] != cellv[(size_t)nb * vpc + (v ^ (1 << dd))])
              return;
        }
    // tables: V (nodal -> Gauss), V^T, then Dt | e | d in the Gauss-point basis
    std::vector<double> gx(h->quad.x, h->quad.x + n1), nodes(h->basis.node, h->basis.node + n1);
    h->mp_tab_host.assign(3 * n1 * n1 + 4 * n1, 0.);
    double             *V = h->mp_tab_host.data(), *Vt = V + n1 * n1, *Dt = Vt + n1 * n1, *e = Dt + n1 * n1, *dd = e + 2 * n1;
    std::vector<double> L(n1), dL(n1);
    for (int q = 0; q < n1; ++q)
      {
        lagrange_on(nodes, gx[q], L.data(), dL.data());
        for (int i = 0; i < n1; ++i)
          {
            V[q * n1 + i]  = L[i]; // out_q = sum_i V[q][i] in_i
            Vt[i * n1 + q] = L[i];
          }
        lagrange_on(gx, gx[q], L.data(), dL.data());
        for (int r = 0; r < n1; ++r)
          Dt[q * n1 + r] = dL[r];
      }
    for (int s = 0; s < 2; ++s)
      lagrange_on(gx, (double)s, e + s * n1, dd + s * n1);
    auto put = [](auto &buf, const auto &v) {
      buf.alloc(v.size());
      PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    put(h->mp_cellv, cellv);
    put(h->mp_nbr, nbr);
    std::vector<double> dt(Dt, Dt + n1 * n1);
    put(h->mp_dt, dt);
    h->mp_ready = true;
  }

  void
  launch_mapped_operator(pd_handle *h, const double *src, double *dst, const bool add)
  {
    const int     dim = h->dim, n1 = h->n1, N = h->n, NF = N / n1;
    const int64_t nc = h->np_own;
    if (!h->mp_geo_valid)
      {
        const int ng = dim * (dim + 1) / 2;
        h->mp_cgeo.alloc((size_t)nc * (ng + 1) * N);
        h->mp_fgeo.alloc((size_t)nc * dim * 2 * (1 + 2 * dim) * NF);
        h->mp_sigma.alloc((size_t)nc * 2 * dim);
        h->mp_xg.alloc((size_t)nc * N);
        h->mp_zero.alloc((size_t)N);
        PD_CUDA(cudaMemsetAsync(h->mp_zero.p, 0, (size_t)N * sizeof(double), h->stream));
        GeoArgs g;
        g.verts = h->verts.p;
        g.cellv = h->mp_cellv.p;
        g.nbr   = h->mp_nbr.p;
        for (int q = 0; q < n1; ++q)
          {
            g.gx[q] = h->quad.x[q];
            g.gw[q] = h->quad.w[q];
          }
        g.dim     = dim;
        g.n1      = n1;
        g.n_cells = (int32_t)nc;
        g.pc      = std::max(h->degree, 1) * (h->degree + 1.0);
        g.cgeo    = h->mp_cgeo.p;
        g.fgeo    = h->mp_fgeo.p;
        g.sigma   = h->mp_sigma.p;
        const int64_t total = nc * (N + 2 * dim * NF);
        const unsigned grid = (unsigned)((total + 127) / 128);
        if (dim == 2)
          k_mapped_geometry<2><<<grid, 128, 0, h->stream>>>(g);
        else
          k_mapped_geometry<3><<<grid, 128, 0, h->stream>>>(g);
        PD_CUDA(cudaGetLastError());
        ++h->launches;
        h->mp_geo_valid = true;
      }
    switch (dim * 10 + h->degree)
      {
        case 21: launch_mapped<2, 1, 3>(h, src, dst, add); break;
        case 22: launch_mapped<2, 2, 3>(h, src, dst, add); break;
        case 23: launch_mapped<2, 3, 2>(h, src, dst, add); break;
        case 24: launch_mapped<2, 4, 2>(h, src, dst, add); break;
        case 31: launch_mapped<3, 1, 3>(h, src, dst, add); break;
        case 32: launch_mapped<3, 2, 3>(h, src, dst, add); break;
        case 33: launch_mapped<3, 3, 2>(h, src, dst, add); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no mapped fine-mesh operator kernel for this (dim, degree)", __LINE__};
      }
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
