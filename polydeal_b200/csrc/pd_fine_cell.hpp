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
// so the thread(s) that own a whole cell apply the premultiplied 1-D stencils
// M^-1 L_d line by line into register accumulators and finish with DIM mass
// passes, all indices compile-time after unrolling -- no exchange between the
// threads of a cell except one partial sum.  FE_DGQ(p >= 1) has nodes on both ends of [0,1], so the trace
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

    // The lines of a cell are split between TWO threads (roles) so that twice as many warps are resident:
    // role 1 takes the first lines_of_role1() lines in (direction, line) order, role 0 the rest plus the mass
    // passes; the split balances the arithmetic (a line costs about N1^2 + 8 N1 + 10 FMAs).
    template <int DIM, int N1>
    constexpr int
    lines_of_role1()
    {
      constexpr int N = ipow(N1, DIM), L = DIM * (N / N1);
      constexpr int line_cost = N1 * N1 + 8 * N1 + 10, mass_cost = DIM * N * N1 + N;
      constexpr int nb = (L * line_cost + mass_cost + line_cost) / (2 * line_cost);
      return nb > L ? L : nb;
    }

    // acc += the premultiplied stencils M^-1 L_d of the lines of `role` (role < 0: all lines) applied to u.
    // u: the cell's coefficients; nbv(d, s, e): coefficient e of the neighbour across face (d, s) (zeros where
    // there is none); coef(d): the folded record of direction d.
    template <int DIM, int N1, class Tab, class Nb, class Coef>
    PD_HD void
    cell_lines(const Tab &T, const int role, const double *u, Nb &&nbv, Coef &&coef, double *acc)
    {
      constexpr int N = ipow(N1, DIM), NL = N / N1, NB = lines_of_role1<DIM, N1>();
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          // (a direction none of whose lines belong to the role costs nothing: everything below is dead code)
          const LineCoef r      = coef(d);
          const int      stride = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
#pragma unroll
          for (int j = 0; j < NL; ++j)
            {
              if (role >= 0 && (d * NL + j < NB) != (role == 1))
                continue;
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
    }

    // acc <- (M (x) M (x) M) acc, one direction after the other, in registers
    template <int DIM, int N1, class Tab>
    PD_HD void
    cell_mass(const Tab &T, double *acc)
    {
      constexpr int N = ipow(N1, DIM), NL = N / N1;
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
    }

    // the whole operator on one cell the way the kernel's two roles compose it: role 0 starts from the mass
    // term mv u (mv = f vol), role 1 from zero, the partial sums meet, then the mass passes.  out may not alias u.
    template <int DIM, int N1, class Tab, class Nb, class Coef>
    PD_HD void
    cell_apply(const Tab &T, const double *u, Nb &&nbv, Coef &&coef, const double mv, double *out)
    {
      constexpr int N = ipow(N1, DIM);
      double        acc0[N], acc1[N];
#pragma unroll
      for (int e = 0; e < N; ++e)
        {
          acc0[e] = mv * u[e];
          acc1[e] = 0.;
        }
      cell_lines<DIM, N1>(T, 0, u, nbv, coef, acc0);
      cell_lines<DIM, N1>(T, 1, u, nbv, coef, acc1);
#pragma unroll
      for (int e = 0; e < N; ++e)
        acc0[e] += acc1[e];
      cell_mass<DIM, N1>(T, acc0);
#pragma unroll
      for (int e = 0; e < N; ++e)
        out[e] = acc0[e];
    }

    // ---- the dense form of the line stencils on a UNIFORM mesh ------------------
    // When all cells are the same box and the penalty depends only on (direction, interior / boundary side), the
    // premultiplied line stencil of direction d is the same three N1 x N1 matrices for every cell,
    //   out_line += A_d u_line + C_d0 n0_line + C_d1 n1_line      (n_s: the neighbour's line across side s)
    // plus a correction dB_ds u_line when side s is a boundary face (a missing neighbour reads zeros, so C needs
    // none).  All entries are kernel constants: a line costs 3 N1^2 multiply-adds with constant-bank operands
    // instead of the N1^2 + 8 N1 + 10 dependent operations of the rank-one form in cell_lines, and the mass
    // term f vol I rides on the diagonal of the last direction's A.
    struct UniformLine
    {
      double cVol, cDi, Pi, Qi, cDb, Pb[2]; // FineRec's fields for an interior side (i) / a boundary side (b)
    };
    template <int DIM, int N1>
    struct DenseTables
    {
      double Mh[N1 * N1];
      double A[DIM][N1 * N1];     // persymmetric (the nodes are symmetric about 1/2): the kernel reads entries k <= N1^2-1-k only
      double C[DIM][N1 * N1];     // across side 0; across side 1 it is the mirror image C[(N1-1-i) N1 + (N1-1-t)]
      double dB[DIM][2][N1 * N1]; // (fewer distinct constants: they stay in uniform registers over the lines of a direction)
    };
    template <int DIM, int N1>
    inline void
    build_dense_tables(const TileTables<N1> &T, const UniformLine *U, const double mv, DenseTables<DIM, N1> &D)
    {
      // the part of the own-line matrix that side s contributes, with penalty P and derivative coefficient cD
      auto side = [&](const int s, const double P, const double cD, const int i, const int t) {
        const int    end = s == 0 ? 0 : N1 - 1;
        const double dl  = t == end ? 1. : 0.;
        return s == 0 ? T.ep[0][i] * (P * dl + cD * T.d[0][t]) + T.dp[0][i] * cD * dl :
                        T.ep[1][i] * (P * dl - cD * T.d[1][t]) - T.dp[1][i] * cD * dl;
      };
      for (int k = 0; k < N1 * N1; ++k)
        D.Mh[k] = T.Mh[k];
      for (int d = 0; d < DIM; ++d)
        for (int i = 0; i < N1; ++i)
          for (int t = 0; t < N1; ++t)
            {
              const UniformLine &u = U[d];
              const int          k = i * N1 + t;
              D.A[d][k] = u.cVol * T.Shp[k] + side(0, u.Pi, u.cDi, i, t) + side(1, u.Pi, u.cDi, i, t) +
                          ((d == DIM - 1 && i == t) ? mv : 0.);
              for (int s = 0; s < 2; ++s)
                D.dB[d][s][k] = side(s, u.Pb[s], u.cDb, i, t) - side(s, u.Pi, u.cDi, i, t);
              const double l1 = t == N1 - 1 ? 1. : 0.;
              D.C[d][k]       = T.ep[0][i] * (-u.Pi * l1 + u.Qi * T.d[1][t]) - T.dp[0][i] * u.cDi * l1;
              // (across side 1: ep1 (-Pi l0 - Qi d0) + dp1 cDi l0 = the mirror image, as ep1, dp1, d1 mirror ep0, -dp0, -d0)
            }
    }

    // role split of the dense form: a line costs about 3 N1^2 multiply-adds + 3 N1 loads, role 0 also does the mass
    // passes (DIM N N1) and both ends of the exchange
    template <int DIM, int N1>
    constexpr int
    dense_lines_of_role1()
    {
      constexpr int N = ipow(N1, DIM), L = DIM * (N / N1);
      constexpr int line_cost = 3 * N1 * N1 + 3 * N1 + 4, mass_cost = DIM * N * N1 + 3 * N;
      constexpr int nb = (L * line_cost + mass_cost + line_cost) / (2 * line_cost);
      return nb > L ? L : nb;
    }

    // acc += the dense line stencils of the lines of `role` (role < 0: all lines).  u: the cell's coefficients (read
    // line by line, not kept); nbv(d, s, e): coefficient e of the neighbour across face (d, s) (zeros where there
    // is none); bnd[2 d + s]: that face is a boundary face.
    template <int DIM, int N1, class Tab, class Nb>
    PD_HD void
    cell_lines_dense(const Tab &T, const int role, const double *u, Nb &&nbv, const bool *bnd, double *acc)
    {
      constexpr int N = ipow(N1, DIM), NL = N / N1, NB = dense_lines_of_role1<DIM, N1>();
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          const int stride = d == 0 ? 1 : (d == 1 ? N1 : N1 * N1);
#pragma unroll
          for (int j = 0; j < NL; ++j)
            {
              if (role >= 0 && (d * NL + j < NB) != (role == 1))
                continue;
              const int base = d == 0 ? j * N1 : (d == 1 ? (j % N1) + (j / N1) * N1 * N1 : j);
              double    v[N1], n0[N1], n1[N1];
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
                  v[t]  = u[base + t * stride];
                  n0[t] = nbv(d, 0, base + t * stride);
                  n1[t] = nbv(d, 1, base + t * stride);
                }
#pragma unroll
              for (int i = 0; i < N1; ++i)
                {
                  double sm = acc[base + i * stride];
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    {
                      const int k = i * N1 + t, km = N1 * N1 - 1 - k;
                      sm += T.A[d][k < km ? k : km] * v[t];
                    }
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    sm += T.C[d][i * N1 + t] * n0[t];
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    sm += T.C[d][N1 * N1 - 1 - (i * N1 + t)] * n1[t];
                  acc[base + i * stride] = sm;
                }
            }
          // boundary sides (rare): one branch per face, not per line -- the lines above stay one straight block that
          // the scheduler can interleave; the own lines are read again
#pragma unroll
          for (int s = 0; s < 2; ++s)
            if (bnd[2 * d + s])
              {
                const double *ub = u;
#if defined(__CUDA_ARCH__)
                asm volatile("" : "+l"(ub)); // (keeps the compiler from holding the lines above in registers for this branch)
#endif
#pragma unroll
                for (int j = 0; j < NL; ++j)
                  {
                    if (role >= 0 && (d * NL + j < NB) != (role == 1))
                      continue;
                    const int base = d == 0 ? j * N1 : (d == 1 ? (j % N1) + (j / N1) * N1 * N1 : j);
                    double    v[N1];
#pragma unroll
                    for (int t = 0; t < N1; ++t)
                      v[t] = ub[base + t * stride];
#pragma unroll
                    for (int i = 0; i < N1; ++i)
                      {
                        double sm = acc[base + i * stride];
#pragma unroll
                        for (int t = 0; t < N1; ++t)
                          sm += T.dB[d][s][i * N1 + t] * v[t];
                        acc[base + i * stride] = sm;
                      }
                  }
              }
        }
    }

    // the whole operator on one cell the way the pipelined kernel's two roles compose it
    template <int DIM, int N1, class Tab, class Nb>
    PD_HD void
    cell_apply_dense(const Tab &T, const double *u, Nb &&nbv, const bool *bnd, double *out)
    {
      constexpr int N = ipow(N1, DIM);
      double        acc0[N], acc1[N];
#pragma unroll
      for (int e = 0; e < N; ++e)
        acc0[e] = acc1[e] = 0.;
      cell_lines_dense<DIM, N1>(T, 0, u, nbv, bnd, acc0);
      cell_lines_dense<DIM, N1>(T, 1, u, nbv, bnd, acc1);
#pragma unroll
      for (int e = 0; e < N; ++e)
        acc0[e] += acc1[e];
      cell_mass<DIM, N1>(T, acc0);
#pragma unroll
      for (int e = 0; e < N; ++e)
        out[e] = acc0[e];
    }

    // ---- tile plan -----------------------------------------------------------
    // A CTA of the tiled kernel takes TILE consecutive entries of a cell sequence (all owned cells, or the
    // interior / boundary lists of a sharded apply), stages their coefficients AND those of every
    // neighbour outside the tile ("halo") in shared memory, and each thread then reads rows of it.
    // Shared-memory layout of the coefficients, in doubles:
    //   i * own_row(n) + [0, n)                     own cell i of the tile (i < TILE)
    //   TILE * own_row(n) + k * halo_row(n) + odd + [0, n)   halo cell halo[tile_ptr[tile] + k]; a halo row is filled by ONE
    //                                               bulk copy (16-byte granularity), so it starts at the 16-byte
    //                                               boundary below the cell's first coefficient: odd = (cell * n) & 1
    //   zoff + [0, n)                               zeros: what a missing neighbour reads
    // noff[(seq position) * 2 DIM + face] is the first double of the neighbour across that face.
    // Halo rows are handed out face by face (all -x neighbours in cell order, then +x, ...), so the lanes of
    // a warp, which read the same face at the same time, mostly read different rows -> few bank conflicts.
    // own rows: n doubles apart when n is odd (the global layout: one bulk copy fills them all), n + 1 when n
    // is even -- consecutive threads read consecutive rows, and an even distance would put them on few banks
    constexpr int
    own_row(const int n)
    {
      return n | 1;
    }
    constexpr int
    halo_row(const int n)
    {
      int r = (n + 2) / 2 * 2; // even, and room for the odd start
      if (r % 4 == 0)
        r += 2; // rows 2 (mod 4) doubles apart spread over all banks
      return r;
    }

    // Tiles are runs of at most TILE consecutive sequence entries that share a block key (cells of one aligned
    // 4x4x4 / 8x8 block of the Morton curve): a list that leaves cells out (the interior / boundary split of a
    // sharded apply) then still gets compact tiles, only smaller ones.
    struct TilePlan
    {
      int32_t               n_tiles = 0, max_halo = 0, zoff = 0;
      std::vector<int32_t>  tile_first, tile_ptr, halo; // [n_tiles + 1]: first sequence entry / first halo cell of a tile
      std::vector<uint16_t> noff;
    };

    // seq: the cells in processing order (nullptr = 0 .. n_seq-1); block_key[i]: block of sequence entry i
    // (nullptr: tiles of `tile` consecutive entries); nbr[cell * nfc + f]: neighbour cell or -1;
    // n_cells_total bounds every id that appears in nbr (owned + ghost cells); n: coefficients per cell
    inline TilePlan
    build_tile_plan(const int32_t n_seq, const int32_t *seq, const uint64_t *block_key, const int32_t *nbr, const int nfc,
                    const int32_t n_cells_total, const int tile, const int n)
    {
      TilePlan  p;
      const int rh = halo_row(n), ro = own_row(n);
      p.tile_first.push_back(0);
      for (int32_t i = 1; i <= n_seq; ++i)
        if (i == n_seq || i - p.tile_first.back() == tile || (block_key && block_key[i] != block_key[i - 1]))
          p.tile_first.push_back(i);
      p.n_tiles = (int32_t)p.tile_first.size() - 1;
      p.tile_ptr.assign((size_t)p.n_tiles + 1, 0);
      p.noff.assign((size_t)n_seq * nfc, 0xFFFF);
      std::vector<int32_t> off_of((size_t)n_cells_total, -1);
      for (int32_t k = 0; k < p.n_tiles; ++k)
        {
          const int32_t s0 = p.tile_first[k], n_own = p.tile_first[k + 1] - s0;
          const size_t  h0 = p.halo.size();
          for (int32_t i = 0; i < n_own; ++i)
            off_of[(size_t)(seq ? seq[s0 + i] : s0 + i)] = i * ro;
          for (int f = 0; f < nfc; ++f)
            for (int32_t i = 0; i < n_own; ++i)
              {
                const int32_t c  = seq ? seq[s0 + i] : s0 + i;
                const int32_t nb = nbr[(size_t)c * nfc + f];
                if (nb < 0)
                  continue; // patched to zoff below
                if (nb >= n_cells_total)
                  throw std::out_of_range("build_tile_plan: neighbour id out of range");
                if (off_of[(size_t)nb] < 0)
                  {
                    const int64_t o = (int64_t)tile * ro + (int64_t)(p.halo.size() - h0) * rh + (((int64_t)nb * n) & 1);
                    if (o + n >= 0xFFFF)
                      throw std::length_error("build_tile_plan: tile too large for 16-bit offsets");
                    off_of[(size_t)nb] = (int32_t)o;
                    p.halo.push_back(nb);
                  }
                p.noff[(size_t)(s0 + i) * nfc + f] = (uint16_t)off_of[(size_t)nb];
              }
          p.max_halo = std::max<int32_t>(p.max_halo, (int32_t)(p.halo.size() - h0));
          for (int32_t i = 0; i < n_own; ++i)
            off_of[(size_t)(seq ? seq[s0 + i] : s0 + i)] = -1;
          for (size_t h = h0; h < p.halo.size(); ++h)
            off_of[(size_t)p.halo[h]] = -1;
          p.tile_ptr[(size_t)k + 1] = (int32_t)p.halo.size();
        }
      const int64_t zoff = (int64_t)tile * ro + (int64_t)p.max_halo * rh;
      if (zoff + n >= 0xFFFF)
        throw std::length_error("build_tile_plan: tile too large for 16-bit offsets");
      p.zoff = (int32_t)zoff;
      for (size_t i = 0; i < p.noff.size(); ++i)
        {
          const int32_t c = seq ? seq[i / nfc] : (int32_t)(i / nfc);
          if (nbr[(size_t)c * nfc + i % nfc] < 0)
            p.noff[i] = (uint16_t)p.zoff;
        }
      return p;
    }
    // ---- tile plan of the pipelined kernel (k_fine_stream) --------------------------------
    // A tile is a run of at most TILE CONSECUTIVE cells (any length: the blocks a METIS partition cuts are partial;
    // any alignment).  Its stage in shared memory, in doubles:
    //   2 - par + i * n + [0, n)    own cell i: the run as it lies in the vector, one double lower when it starts 8
    //                               bytes past a 16-byte boundary there (par = (first cell * n) & 1), so that shared
    //                               memory and vector have the same 16-byte phase (bulk copies)
    //   HB + r * n + [0, n)         halo row r, HB = 2 + TILE * n: the rows after the own rows of an aligned tile
    //   zoff + [0, n)               zeros
    // All rows are n doubles apart, n odd (the 16-byte aligned rows of the other plan are an even number apart: rows r
    // and r + 8 share their banks).  The halo rows are filled by 16-byte cp.async chunks all the same: a halo cell
    // whose first coefficient sits on a 16-byte boundary where it is read gets an EVEN row, the others an odd row --
    // source and destination then have the same phase and the row is (n - 1) / 2 chunks of 16 bytes plus one of 8.
    // rows[tile][r] = cell or -1.
    struct StreamPlan
    {
      int32_t               n_tiles = 0, max_rows = 0, zoff = 0, halo_base = 0;
      std::vector<int32_t>  rows; // [n_tiles][max_rows]
      std::vector<uint16_t> noff; // [n_seq][nfc]
    };
    // tile_first[n_tiles + 1]: the tiles' sequence ranges; the cells of a tile must be consecutive numbers (checked).
    // src_parity (optional, [n_cells_total]): whether a cell's coefficients start 8 bytes past a 16-byte boundary where
    // the kernel reads them as a HALO cell (default: (cell * n) & 1, the vector itself; ghost cells read straight from
    // a peer's export buffer sit wherever that buffer has them)
    inline StreamPlan
    build_stream_plan(const int32_t n_seq, const int32_t *seq, const int32_t *tile_first, const int32_t n_tiles, const int32_t *nbr,
                      const int nfc, const int32_t n_cells_total, const int tile, const int n, const uint8_t *src_parity = nullptr)
    {
      if (n % 2 == 0)
        throw std::invalid_argument("build_stream_plan: n odd only");
      StreamPlan p;
      p.n_tiles   = n_tiles;
      p.halo_base = tile * n + 2;
      std::vector<std::vector<int32_t>> rows((size_t)p.n_tiles);
      std::vector<int32_t>              row_of((size_t)n_cells_total, -1), own_of((size_t)n_cells_total, -1);
      std::vector<int32_t>              noff_row((size_t)n_seq * nfc, -1); // >= 0: halo row, -2 - i: own cell i, -1: none
      std::vector<uint8_t>              tile_par((size_t)p.n_tiles, 0);
      for (int32_t k = 0; k < p.n_tiles; ++k)
        {
          const int32_t s0 = tile_first[k], n_own = tile_first[k + 1] - s0;
          if (n_own < 1 || n_own > tile)
            throw std::invalid_argument("build_stream_plan: tile size");
          const int32_t c0 = seq ? seq[s0] : s0;
          for (int32_t i = 1; i < n_own; ++i)
            if ((seq ? seq[s0 + i] : s0 + i) != c0 + i)
              throw std::invalid_argument("build_stream_plan: the cells of a tile must be consecutive");
          tile_par[(size_t)k] = (uint8_t)(((int64_t)c0 * n) & 1);
          auto   &R           = rows[(size_t)k];
          int32_t next[2]     = {0, 1};
          for (int32_t i = 0; i < n_own; ++i)
            own_of[(size_t)(c0 + i)] = i;
          for (int f = 0; f < nfc; ++f) // face by face: the lanes of a warp, which read one face at a time, read different rows
            for (int32_t i = 0; i < n_own; ++i)
              {
                const int32_t nb = nbr[(size_t)(c0 + i) * nfc + f];
                if (nb < 0)
                  continue;
                if (nb >= n_cells_total)
                  throw std::out_of_range("build_stream_plan: neighbour id out of range");
                if (own_of[(size_t)nb] >= 0)
                  {
                    noff_row[(size_t)(s0 + i) * nfc + f] = -2 - own_of[(size_t)nb];
                    continue;
                  }
                if (row_of[(size_t)nb] < 0)
                  {
                    const int par = src_parity ? (src_parity[(size_t)nb] & 1) : (int)(((int64_t)nb * n) & 1);
                    row_of[(size_t)nb] = next[par];
                    next[par] += 2;
                    if ((int32_t)R.size() <= row_of[(size_t)nb])
                      R.resize((size_t)row_of[(size_t)nb] + 1, -1);
                    R[(size_t)row_of[(size_t)nb]] = nb;
                  }
                noff_row[(size_t)(s0 + i) * nfc + f] = row_of[(size_t)nb];
              }
          for (int32_t i = 0; i < n_own; ++i)
            own_of[(size_t)(c0 + i)] = -1;
          for (const int32_t c : R)
            if (c >= 0)
              row_of[(size_t)c] = -1;
          p.max_rows = std::max<int32_t>(p.max_rows, (int32_t)R.size());
        }
      const int64_t zoff = (int64_t)p.halo_base + (int64_t)p.max_rows * n;
      if (zoff + n >= 0xFFFF)
        throw std::length_error("build_stream_plan: tile too large for 16-bit offsets");
      p.zoff = (int32_t)zoff;
      p.rows.assign((size_t)p.n_tiles * std::max<int32_t>(1, p.max_rows), -1);
      for (int32_t k = 0; k < p.n_tiles; ++k)
        std::copy(rows[(size_t)k].begin(), rows[(size_t)k].end(), p.rows.begin() + (size_t)k * p.max_rows);
      p.noff.resize((size_t)n_seq * nfc);
      for (int32_t k = 0; k < p.n_tiles; ++k)
        for (size_t i = (size_t)tile_first[k] * nfc; i < (size_t)tile_first[k + 1] * nfc; ++i)
          {
            const int32_t r = noff_row[i];
            p.noff[i]       = (uint16_t)(r == -1 ? p.zoff : (r <= -2 ? 2 - tile_par[(size_t)k] + (-2 - r) * n : p.halo_base + r * n));
          }
      return p;
    }
    // Rows build_stream_plan gives the tile seq[s0, s1): its distinct halo cells by 16-byte phase -- even rows 0, 2, ..
    // for the cells on a boundary, odd rows 1, 3, .. for the others -- so twice the larger class.  stamp: scratch of
    // n_cells_total entries, none equal to id.
    inline int32_t
    stream_tile_rows(const int32_t *seq, const int32_t s0, const int32_t s1, const int32_t *nbr, const int nfc, const int n,
                     const uint8_t *src_parity, std::vector<int32_t> &stamp, const int32_t id)
    {
      for (int32_t i = s0; i < s1; ++i)
        stamp[(size_t)(seq ? seq[i] : i)] = id;
      int32_t count[2] = {0, 0};
      for (int32_t i = s0; i < s1; ++i)
        for (int f = 0; f < nfc; ++f)
          {
            const int32_t nb = nbr[(size_t)(seq ? seq[i] : i) * nfc + f];
            if (nb < 0 || nb >= (int32_t)stamp.size() || stamp[(size_t)nb] == id)
              continue;
            stamp[(size_t)nb] = id;
            ++count[src_parity ? (src_parity[(size_t)nb] & 1) : (int)(((int64_t)nb * n) & 1)];
          }
      return std::max(count[0] > 0 ? 2 * count[0] - 1 : 0, 2 * count[1]);
    }
    // Tiles that would need more than max_rows halo rows are halved until they fit (a tile is any run of consecutive
    // cells; one cell needs at most 2 nfc rows).  The ghost cells of the fused sharded apply lie in their owners'
    // export buffers at whatever 16-byte phase those give them: a tile next to a cut most of whose halo cells share
    // one phase needs twice as many rows as it has halo cells, more than the kernel's gather holds.  Returns the new
    // tile_first.
    inline std::vector<int32_t>
    split_stream_tiles(const int32_t *seq, const int32_t *tile_first, const int32_t n_tiles, const int32_t *nbr, const int nfc,
                       const int32_t n_cells_total, const int n, const uint8_t *src_parity, const int32_t max_rows)
    {
      if (max_rows < 2 * nfc)
        throw std::invalid_argument("split_stream_tiles: max_rows below what one cell needs");
      std::vector<int32_t> out, stamp((size_t)n_cells_total, -1), todo;
      int32_t              id = 0;
      for (int32_t k = 0; k < n_tiles; ++k)
        {
          // (pieces of the tile still to check, last first: they come out in sequence order)
          todo.assign({tile_first[k + 1], tile_first[k]});
          while (todo.size() >= 2)
            {
              const int32_t s0 = todo.back(), s1 = todo[todo.size() - 2];
              if (s1 - s0 > 1 && stream_tile_rows(seq, s0, s1, nbr, nfc, n, src_parity, stamp, id++) > max_rows)
                todo.insert(todo.end() - 1, s0 + (s1 - s0) / 2);
              else
                {
                  out.push_back(s0);
                  todo.pop_back();
                }
            }
        }
      out.push_back(tile_first[n_tiles]);
      return out;
    }
    // The plan of the fused sharded apply: ONE tile sequence, the tiles of the interior list followed by those of the
    // boundary list (the only ones that read ghost cells and therefore wait for the owners' epoch flags), tiles that
    // would not fit the gather split.  src_parity as in build_stream_plan (own cells: (cell * n) & 1; ghost cells:
    // where the owner's export buffer has them).
    struct FusedPlan
    {
      std::vector<int32_t> seq, tile_first, tile_base;
      int32_t              first_ghost_tile = 0;
      StreamPlan           sp;
    };
    inline FusedPlan
    build_fused_plan(const std::vector<int32_t> &inner, const std::vector<int32_t> &outer, const std::vector<int32_t> &inner_tile_first,
                     const std::vector<int32_t> &outer_tile_first, const int32_t *nbr, const int nfc, const int32_t n_cells_total,
                     const int tile, const int n, const uint8_t *src_parity, const int32_t max_rows)
    {
      if (inner_tile_first.empty() || outer_tile_first.empty() || inner_tile_first.back() != (int32_t)inner.size() ||
          outer_tile_first.back() != (int32_t)outer.size())
        throw std::invalid_argument("build_fused_plan: tile ranges do not cover the lists");
      FusedPlan p;
      p.seq = inner;
      p.seq.insert(p.seq.end(), outer.begin(), outer.end());
      std::vector<int32_t> tf(inner_tile_first);
      for (size_t k = 1; k < outer_tile_first.size(); ++k)
        tf.push_back(outer_tile_first[k] + (int32_t)inner.size());
      p.tile_first = split_stream_tiles(p.seq.data(), tf.data(), (int32_t)tf.size() - 1, nbr, nfc, n_cells_total, n, src_parity, max_rows);
      const int32_t n_tiles = (int32_t)p.tile_first.size() - 1;
      p.tile_base.resize((size_t)n_tiles);
      p.first_ghost_tile = n_tiles;
      for (int32_t k = 0; k < n_tiles; ++k)
        {
          p.tile_base[(size_t)k] = p.seq[(size_t)p.tile_first[(size_t)k]];
          if (p.tile_first[(size_t)k] >= (int32_t)inner.size() && k < p.first_ghost_tile)
            p.first_ghost_tile = k;
        }
      p.sp = build_stream_plan((int32_t)p.seq.size(), p.seq.data(), p.tile_first.data(), n_tiles, nbr, nfc, n_cells_total, tile, n, src_parity);
      return p;
    }
  } // namespace fine
} // namespace pd

#endif
