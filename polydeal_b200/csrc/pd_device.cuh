// -----------------------------------------------------------------------------
// pd_device.cuh -- device helpers shared by the assembly and the matrix-free
// polytopal kernels: compile-time sizes of FE_DGQ<dim>(p) on the bounding box,
// the FP64 tensor-core MMA wrapper and the 1-D Lagrange evaluation.
// -----------------------------------------------------------------------------
#pragma once
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

    constexpr int
    binom(const int n, const int k)
    {
      return k == 0 ? 1 : binom(n - 1, k - 1) * n / k;
    }

    // Compile-time sizes of the element on the bounding box.  The second template argument
    // encodes the element family: DEGX = p is FE_DGQ<dim>(p) (tensor Lagrange basis on the
    // Gauss-Lobatto nodes, (p+1)^dim DoFs); DEGX = DGP_BASE + p is FE_AggloDGP<dim>(p)
    // (source/fe_agglodgp.cc:28-57: products of L2-orthonormal Legendre polynomials of total
    // degree <= p, C(p+dim, dim) DoFs in deal.II's PolynomialSpace order: last coordinate
    // outermost, first coordinate fastest).
    constexpr int DGP_BASE = 10;
    template <int DIM, int DEGX>
    struct Cfg
    {
      static constexpr bool DGP = DEGX >= DGP_BASE;
      static constexpr int  P   = DGP ? DEGX - DGP_BASE : DEGX;
      static constexpr int  N1  = P + 1;
      static constexpr int  N   = DGP ? binom(P + DIM, DIM) : ipow(N1, DIM);
      static constexpr int  NT8 = (N + 7) / 8; // 8x8 MMA tiles per side
      static constexpr int  NP  = NT8 * 8;
      static constexpr int  NU  = ipow(N1, DIM - 1); // generator units per point: the (b[,c]) pairs
      // Row stride (doubles) of operand panels.  A fragment load touches 4 rows x
      // 8 consecutive doubles; with stride = 4, 8 or 12 (mod 16) the four rows
      // cover every bank pair exactly twice => 2 wavefronts, the minimum for 256 B.
      static constexpr int STRIDE = (NP % 16 == 0) ? NP + 8 : NP;
      static constexpr int NTRI   = NT8 * (NT8 + 1) / 2;

      // Generator unit wu = b (2-D) or b + N1 c (3-D) holds the DoFs (a, b[, c]), a = 0..count-1,
      // in consecutive rows starting at base.
      __host__ __device__ static constexpr int
      unit_count(const int wu)
      {
        if (!DGP)
          return N1;
        const int r = P + 1 - (DIM == 3 ? wu % N1 + wu / N1 : wu);
        return r > 0 ? r : 0;
      }
      __host__ __device__ static constexpr int
      unit_base(const int wu)
      {
        if (!DGP)
          return wu * N1;
        const int b = DIM == 3 ? wu % N1 : wu, c = DIM == 3 ? wu / N1 : 0;
        int       base = 0, m = P;
        if (DIM == 3)
          {
            for (int cc = 0; cc < c; ++cc)
              base += (P - cc + 1) * (P - cc + 2) / 2;
            m = P - c;
          }
        return base + b * (m + 1) - b * (b - 1) / 2;
      }
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
      // barrier 0, issued from role-specialised code paths, hence PTX
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

    // L2[0,1]-orthonormal Legendre polynomials sqrt(2k+1) P_k(2x-1) and their derivatives * scale
    template <int N1>
    __device__ __forceinline__ void
    legendre01(const double x, const double scale, double *L, double *dL)
    {
      const double t = 2. * x - 1.;
      double       p0 = 1., p1 = t, d0 = 0., d1 = 1.;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        {
          double pk = k == 0 ? p0 : p1, dk = k == 0 ? d0 : d1;
          if (k >= 2)
            {
              pk = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
              dk = ((2 * k - 1) * (p1 + t * d1) - (k - 1) * d0) / k;
              p0 = p1;
              p1 = pk;
              d0 = d1;
              d1 = dk;
            }
          const double s = sqrt(2. * k + 1.);
          L[k]           = s * pk;
          dL[k]          = s * dk * 2. * scale;
        }
    }

    // the 1-D factors of the element family C at x
    template <class C>
    __device__ __forceinline__ void
    basis_1d(const Basis1D &B, const double x, const double scale, double *L, double *dL)
    {
      if constexpr (C::DGP)
        legendre01<C::N1>(x, scale, L, dL);
      else
        lagrange<C::N1>(B, x, scale, L, dL);
    }

  } // namespace
} // namespace pd
