// -----------------------------------------------------------------------------
// pd_fine_cell.hpp -- the per-cell arithmetic and the tile plan of the tiled
// fine-mesh SIP kernel (k_fine_tile in pd_finemesh.cu), kept free of CUDA
// runtime types so that tests/csrc/fine_cell_host.cpp can instantiate the very
// same templates with g++ and check them against a dense Kronecker restatement
// (tests/test_fine_tile_host.py).  Nothing here is a CPU execution path of the
// product: the library only ever calls cell_apply from device code.
//
// Reference semantics: LaplaceOperatorDG / MonodomainOperatorDG,
// include/utils.h:819-925, 1565-1659 (see pd_finemesh.cu).
//
// On a Cartesian cell the operator is
//   sum_d  M (x) .. (x) L_d (x) .. (x) M  +  f vol  M (x) M (x) M
//     = (M (x) M (x) M) [ sum_d  I (x) .. (x) M^-1 L_d (x) .. (x) I  +  f vol I ]
// so ONE thread that owns a whole cell applies the premultiplied 1-D stencils
// M^-1 L_d line by line into a register accumulator and finishes with DIM mass
// passes, all indices compile-time after unrolling -- no exchange between
// threads at all.  FE_DGQ(p >= 1) has nodes on both ends of [0,1], so the trace
// functionals l_i(0), l_i(1) are unit vectors and only the derivative
// functionals d_s = l_i'(s) cost arithmetic.
// -----------------------------------------------------------------------------
#ifndef PD_FINE_CELL_HPP
#define PD_FINE_CELL_HPP

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

#if defined(__CUDACC__)
#  define PD_HD __host__ __device__ __forceinline__
#else
#  define PD_HD inline
#endif

namespace pd
{
  namespace fine
  {
    constexpr int
    ipow(const int b, const int e)
    {
      return e == 0 ? 1 : b * ipow(b, e - 1);
    }

    // 1-D tables of the tiled kernel (by value in the kernel parameters: constant-bank operands)
    template <int N1>
    struct TileTables
    {
      double Mh[N1 * N1];  // mass, the final passes
      double Shp[N1 * N1]; // Mh^-1 Sh
      double ep[2][N1];    // Mh^-1 e_s
      double dp[2][N1];    // Mh^-1 d_s
      double d[2][N1];     // l_i'(0), l_i'(1)
    };

    // the coefficients of one (cell, direction): FineRec of pd_finemesh.cu without the neighbour ids
    struct LineCoef
    {
      double cVol, cD[2], P[2], Q[2];
    };

    // u: the cell's coefficients; nbv(d, s, e): coefficient e of the neighbour across face (d, s) (zeros where
    // there is none); coef(d): the folded record; mv = f vol.  out may not alias u.
    template <int DIM, int N1, class Tab, class Nb, class Coef>
    PD_HD void
    cell_apply(const Tab &T, const double *u, Nb &&nbv, Coef &&coef, const double mv, double *out)
    {
      constexpr int N = ipow(N1, DIM), NL = N / N1;
      double        acc[N];
#pragma unroll
      for (int e = 0; e < N; ++e)
        acc[e] = mv * u[e];
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          const LineCoef r      = coef(d);
          const int      stride = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
#pragma unroll
          for (int j = 0; j < NL; ++j)
            {
              const int base = d == 0 ? j * N1 : (d == 1 ? (j % N1) + (j / N1) * N1 * N1 : j);
              double    n0[N1], n1[N1];
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
                  n0[t] = nbv(d, 0, base + t * stride);
                  n1[t] = nbv(d, 1, base + t * stride);
                }
              double du0 = 0., du1 = 0., dn0 = 0., dn1 = 0.;
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
                  du0 += T.d[0][t] * u[base + t * stride];
                  du1 += T.d[1][t] * u[base + t * stride];
                  dn0 += T.d[1][t] * n0[t]; // the neighbour's facing end
                  dn1 += T.d[0][t] * n1[t];
                }
              // av = P [u] - sn (cD du + Q dn),  bv = -sn cD [u]   (sn = -1 / +1)
              const double j0 = u[base] - n0[N1 - 1], j1 = u[base + (N1 - 1) * stride] - n1[0];
              const double a0 = r.P[0] * j0 + (r.cD[0] * du0 + r.Q[0] * dn0), b0 = r.cD[0] * j0;
              const double a1 = r.P[1] * j1 - (r.cD[1] * du1 + r.Q[1] * dn1), b1 = -r.cD[1] * j1;
#pragma unroll
              for (int i = 0; i < N1; ++i)
                {
                  double sv = 0.;
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    sv += T.Shp[i * N1 + t] * u[base + t * stride];
                  acc[base + i * stride] += r.cVol * sv + T.ep[0][i] * a0 + T.ep[1][i] * a1 + T.dp[0][i] * b0 + T.dp[1][i] * b1;
                }
            }
        }
      // M (x) M (x) M, one direction after the other, in registers
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          const int stride = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
#pragma unroll
          for (int j = 0; j < NL; ++j)
            {
              const int base = d == 0 ? j * N1 : (d == 1 ? (j % N1) + (j / N1) * N1 * N1 : j);
              double    v[N1];
#pragma unroll
              for (int t = 0; t < N1; ++t)
                v[t] = acc[base + t * stride];
#pragma unroll
              for (int i = 0; i < N1; ++i)
                {
                  double sm = 0.;
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    sm += T.Mh[i * N1 + t] * v[t];
                  acc[base + i * stride] = sm;
                }
            }
        }
#pragma unroll
      for (int e = 0; e < N; ++e)
        out[e] = acc[e];
    }

    // ---- tile plan -----------------------------------------------------------
    // A CTA of the tiled kernel takes TILE consecutive entries of a cell sequence (all owned cells, or the
    // interior / boundary lists of a sharded apply), stages their coefficients AND those of every
    // neighbour outside the tile ("halo") in shared memory, and each thread then works on slots:
    //   slot i < n_own            own cell i of the tile
    //   slot n_own + k            halo cell halo[tile_ptr[tile] + k]
    //   slot zslot (= max slots)  zeros: what a missing neighbour reads
    // nslot[(seq position) * 2 DIM + face] is the slot of the neighbour across that face.
    struct TilePlan
    {
      int32_t               n_tiles = 0, zslot = 0;
      std::vector<int32_t>  tile_ptr, halo;
      std::vector<uint16_t> nslot;
    };

    // seq: the cells in processing order (nullptr = 0 .. n_seq-1); nbr[cell * nfc + f]: neighbour cell or -1;
    // n_cells_total bounds every id that appears in nbr (owned + ghost cells)
    inline TilePlan
    build_tile_plan(const int32_t n_seq, const int32_t *seq, const int32_t *nbr, const int nfc, const int32_t n_cells_total,
                    const int tile)
    {
      TilePlan p;
      p.n_tiles = (n_seq + tile - 1) / tile;
      p.tile_ptr.assign((size_t)p.n_tiles + 1, 0);
      p.nslot.assign((size_t)n_seq * nfc, 0xFFFF);
      std::vector<int32_t> slot_of((size_t)n_cells_total, -1);
      int                  max_slots = 0;
      for (int32_t k = 0; k < p.n_tiles; ++k)
        {
          const int32_t s0 = k * tile, n_own = std::min<int32_t>(tile, n_seq - s0);
          const size_t  h0 = p.halo.size();
          for (int32_t i = 0; i < n_own; ++i)
            slot_of[(size_t)(seq ? seq[s0 + i] : s0 + i)] = i;
          int32_t n_slots = n_own;
          for (int32_t i = 0; i < n_own; ++i)
            {
              const int32_t c = seq ? seq[s0 + i] : s0 + i;
              for (int f = 0; f < nfc; ++f)
                {
                  const int32_t nb = nbr[(size_t)c * nfc + f];
                  if (nb < 0)
                    continue; // patched to zslot below
                  if (nb >= n_cells_total)
                    throw std::out_of_range("build_tile_plan: neighbour id out of range");
                  if (slot_of[(size_t)nb] < 0)
                    {
                      slot_of[(size_t)nb] = n_slots++;
                      p.halo.push_back(nb);
                    }
                  p.nslot[(size_t)(s0 + i) * nfc + f] = (uint16_t)slot_of[(size_t)nb];
                }
            }
          max_slots = std::max(max_slots, n_slots);
          for (int32_t i = 0; i < n_own; ++i)
            slot_of[(size_t)(seq ? seq[s0 + i] : s0 + i)] = -1;
          for (size_t h = h0; h < p.halo.size(); ++h)
            slot_of[(size_t)p.halo[h]] = -1;
          p.tile_ptr[(size_t)k + 1] = (int32_t)p.halo.size();
        }
      if (max_slots >= 0xFFFF)
        throw std::length_error("build_tile_plan: tile with more than 65534 slots");
      p.zslot = max_slots;
      for (size_t i = 0; i < p.nslot.size(); ++i)
        {
          const int32_t c = seq ? seq[i / nfc] : (int32_t)(i / nfc);
          if (nbr[(size_t)c * nfc + i % nfc] < 0)
            p.nslot[i] = (uint16_t)p.zslot;
        }
      return p;
    }
  } // namespace fine
} // namespace pd

#endif
