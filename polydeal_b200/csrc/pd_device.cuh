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

    template <int DIM, int DEG>
    struct Cfg
    {
      static constexpr int N1  = DEG + 1;
      static constexpr int N   = ipow(N1, DIM);
      static constexpr int NT8 = (N + 7) / 8; // 8x8 MMA tiles per side
      static constexpr int NP  = NT8 * 8;
      static constexpr int NU  = ipow(N1, DIM - 1); // generator units per point
      // Row stride (doubles) of operand panels.  A fragment load touches 4 rows x
      // 8 consecutive doubles; with stride = 4, 8 or 12 (mod 16) the four rows
      // cover every bank pair exactly twice => 2 wavefronts, the minimum for 256 B.
      static constexpr int STRIDE = (NP % 16 == 0) ? NP + 8 : NP;
      static constexpr int NTRI   = NT8 * (NT8 + 1) / 2;
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

  } // namespace
} // namespace pd
