// Test-only host instantiation of the tiled fine-mesh kernel's per-cell arithmetic and tile plan
// (polydeal_b200/csrc/pd_fine_cell.hpp), compiled by tests/test_fine_tile_host.py with g++ and
// compared with a dense Kronecker restatement in numpy.  Not part of the product library.
#include "../../polydeal_b200/csrc/pd_fine_cell.hpp"

#include <cstring>

namespace
{
  template <int DIM, int N1>
  void
  run(const double *tab, const double *u, const double *nb, const double *coef, const double mv, double *out)
  {
    constexpr int               N = pd::fine::ipow(N1, DIM);
    pd::fine::TileTables<N1>    T;
    static_assert(sizeof(T) == (2 * N1 * N1 + 6 * N1) * sizeof(double), "table layout");
    std::memcpy(&T, tab, sizeof(T));
    pd::fine::cell_apply<DIM, N1>(
      T, u, [&](const int d, const int s, const int e) { return nb[(2 * d + s) * N + e]; },
      [&](const int d) {
        pd::fine::LineCoef c;
        std::memcpy(&c, coef + 7 * d, sizeof(c));
        return c;
      },
      mv, out);
  }

  // the dense form on a uniform mesh: uni[dim][7] = cVol, cDi, Pi, Qi, cDb, Pb0, Pb1; bnd[2 dim]
  template <int DIM, int N1>
  void
  run_dense(const double *tab, const double *u, const double *nb, const double *uni, const int *bnd, const double mv, double *out)
  {
    constexpr int            N = pd::fine::ipow(N1, DIM);
    pd::fine::TileTables<N1> T;
    std::memcpy(&T, tab, sizeof(T));
    pd::fine::UniformLine U[DIM];
    for (int d = 0; d < DIM; ++d)
      U[d] = pd::fine::UniformLine{uni[7 * d], uni[7 * d + 1], uni[7 * d + 2], uni[7 * d + 3], uni[7 * d + 4], {uni[7 * d + 5], uni[7 * d + 6]}};
    pd::fine::DenseTables<DIM, N1> D;
    pd::fine::build_dense_tables<DIM, N1>(T, U, mv, D);
    bool b[2 * DIM];
    for (int f = 0; f < 2 * DIM; ++f)
      b[f] = bnd[f] != 0;
    pd::fine::cell_apply_dense<DIM, N1>(
      D, u, [&](const int d, const int s, const int e) { return nb[(2 * d + s) * N + e]; }, b, out);
  }
} // namespace

extern "C"
{
  int
  fine_cell_dense_host(const int dim, const int n1, const double *tab, const double *u, const double *nb, const double *uni,
                       const int *bnd, const double mv, double *out)
  {
    switch (dim * 10 + n1)
      {
        case 23: run_dense<2, 3>(tab, u, nb, uni, bnd, mv, out); return 0;
        case 25: run_dense<2, 5>(tab, u, nb, uni, bnd, mv, out); return 0;
        case 33: run_dense<3, 3>(tab, u, nb, uni, bnd, mv, out); return 0;
        case 34: run_dense<3, 4>(tab, u, nb, uni, bnd, mv, out); return 0;
        default: return -1;
      }
  }

  // tab: Mh | Mh^-1 Sh | Mh^-1 e0,e1 | Mh^-1 d0,d1 | d0,d1;  nb[2 dim][N];  coef[dim][7] = cVol, cD0, cD1, P0, P1, Q0, Q1
  int
  fine_cell_host(const int dim, const int n1, const double *tab, const double *u, const double *nb, const double *coef,
                 const double mv, double *out)
  {
    switch (dim * 10 + n1)
      {
        case 22: run<2, 2>(tab, u, nb, coef, mv, out); return 0;
        case 23: run<2, 3>(tab, u, nb, coef, mv, out); return 0;
        case 24: run<2, 4>(tab, u, nb, coef, mv, out); return 0;
        case 25: run<2, 5>(tab, u, nb, coef, mv, out); return 0;
        case 32: run<3, 2>(tab, u, nb, coef, mv, out); return 0;
        case 33: run<3, 3>(tab, u, nb, coef, mv, out); return 0;
        case 34: run<3, 4>(tab, u, nb, coef, mv, out); return 0;
        default: return -1;
      }
  }

  // the pipelined kernel's plan; outputs sized by the caller: rows[n_tiles * rows_cap], noff[n_seq * nfc]
  int
  fine_stream_plan_host(const int32_t n_seq, const int32_t *seq, const int32_t *tile_first, const int32_t n_tiles, const int32_t *nbr,
                        const int nfc, const int32_t n_cells_total, const int tile, const int n, const int32_t rows_cap,
                        int32_t *max_rows, int32_t *zoff, int32_t *halo_base, int32_t *rows, uint16_t *noff)
  {
    try
      {
        const pd::fine::StreamPlan p = pd::fine::build_stream_plan(n_seq, seq, tile_first, n_tiles, nbr, nfc, n_cells_total, tile, n);
        if (p.max_rows > rows_cap)
          return -2;
        *max_rows  = p.max_rows;
        *zoff      = p.zoff;
        *halo_base = p.halo_base;
        for (int32_t k = 0; k < p.n_tiles; ++k)
          for (int32_t r = 0; r < rows_cap; ++r)
            rows[(size_t)k * rows_cap + r] = r < p.max_rows ? p.rows[(size_t)k * p.max_rows + r] : -1;
        std::memcpy(noff, p.noff.data(), p.noff.size() * sizeof(uint16_t));
        return 0;
      }
    catch (const std::exception &)
      {
        return -1;
      }
  }

  // outputs sized by the caller: tile_first[n_seq + 1], tile_ptr[n_seq + 1], noff[n_seq * nfc], halo[halo_cap]
  int
  fine_tile_plan_host(const int32_t n_seq, const int32_t *seq, const uint64_t *block_key, const int32_t *nbr, const int nfc,
                      const int32_t n_cells_total, const int tile, const int n, int32_t *n_tiles, int32_t *max_halo, int32_t *zoff,
                      int32_t *halo_row, int32_t *tile_first, int32_t *tile_ptr, int32_t *halo, const int64_t halo_cap,
                      int64_t *n_halo, uint16_t *noff)
  {
    try
      {
        const pd::fine::TilePlan p = pd::fine::build_tile_plan(n_seq, seq, block_key, nbr, nfc, n_cells_total, tile, n);
        if ((int64_t)p.halo.size() > halo_cap)
          return -2;
        std::memcpy(tile_first, p.tile_first.data(), p.tile_first.size() * sizeof(int32_t));
        *n_tiles                   = p.n_tiles;
        *max_halo                  = p.max_halo;
        *zoff                      = p.zoff;
        *halo_row                  = pd::fine::halo_row(n);
        *n_halo                    = (int64_t)p.halo.size();
        if ((int64_t)p.halo.size() > halo_cap)
          return -2;
        std::memcpy(tile_ptr, p.tile_ptr.data(), p.tile_ptr.size() * sizeof(int32_t));
        std::memcpy(halo, p.halo.data(), p.halo.size() * sizeof(int32_t));
        std::memcpy(noff, p.noff.data(), p.noff.size() * sizeof(uint16_t));
        return 0;
      }
    catch (const std::exception &)
      {
        return -1;
      }
  }
  // The fused sharded plan (interior ++ boundary tiles, tiles split to the gather's row budget).  Outputs sized by the
  // caller: seq[n_inner + n_outer], tile_first[n_inner + n_outer + 1], rows[n_tiles_out * max_rows] (queried through
  // rows_cap: -2 when too small), noff[(n_inner + n_outer) * nfc].  unsplit_max_rows: what the unsplit tiles need.
  int
  fine_fused_plan_host(const int32_t n_inner, const int32_t *inner, const int32_t n_outer, const int32_t *outer, const int32_t n_t1,
                       const int32_t *tf1, const int32_t n_t2, const int32_t *tf2, const int32_t *nbr, const int nfc,
                       const int32_t n_cells_total, const int tile, const int n, const uint8_t *src_parity, const int32_t max_rows,
                       int32_t *n_tiles, int32_t *first_ghost_tile, int32_t *plan_max_rows, int32_t *zoff, int32_t *unsplit_max_rows,
                       int32_t *seq, int32_t *tile_first, int32_t *tile_base, int32_t *rows, const int64_t rows_cap, uint16_t *noff)
  {
    try
      {
        const std::vector<int32_t> in(inner, inner + n_inner), out(outer, outer + n_outer), t1(tf1, tf1 + n_t1 + 1), t2(tf2, tf2 + n_t2 + 1);
        const pd::fine::FusedPlan  big = pd::fine::build_fused_plan(in, out, t1, t2, nbr, nfc, n_cells_total, tile, n, src_parity, 1 << 20);
        *unsplit_max_rows          = big.sp.max_rows;
        const pd::fine::FusedPlan p = pd::fine::build_fused_plan(in, out, t1, t2, nbr, nfc, n_cells_total, tile, n, src_parity, max_rows);
        *n_tiles                    = (int32_t)p.tile_first.size() - 1;
        *first_ghost_tile           = p.first_ghost_tile;
        *plan_max_rows              = p.sp.max_rows;
        *zoff                       = p.sp.zoff;
        if ((int64_t)p.sp.rows.size() > rows_cap)
          return -2;
        std::memcpy(seq, p.seq.data(), p.seq.size() * sizeof(int32_t));
        std::memcpy(tile_first, p.tile_first.data(), p.tile_first.size() * sizeof(int32_t));
        std::memcpy(tile_base, p.tile_base.data(), p.tile_base.size() * sizeof(int32_t));
        std::memcpy(rows, p.sp.rows.data(), p.sp.rows.size() * sizeof(int32_t));
        std::memcpy(noff, p.sp.noff.data(), p.sp.noff.size() * sizeof(uint16_t));
        return 0;
      }
    catch (const std::exception &)
      {
        return -1;
      }
  }
}
