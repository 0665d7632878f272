// -----------------------------------------------------------------------------
// pd_finemesh.cu -- matrix-free, sum-factorised SIP operator apply on a fine
// (non-agglomerated) Cartesian hex/quad mesh.
//
// Reference semantics:
//   Utils::MatrixFreeOperators::LaplaceOperatorDG::{vmult, vmult_add,
//     local_apply, local_apply_face, local_apply_boundary}
//     include/utils.h:423-473, 819-925
//   MonodomainOperatorDG: y = (f M + sigma K_SIP) x, no boundary term
//     include/utils.h:1131-1134, 1565-1659
// i.e. cell term  int grad u . grad v (+ f u v), interior faces
//   sigma_F [u][v] - {dn u}[v] - [u]{dn v},  sigma_F = p(p+1)(1/h_m + 1/h_p)
// and boundary faces  4 p(p+1)/h u v - dn u v - u dn v.  The penalties arrive per
// face through the flattened agglomeration (sub_sigma), so any of the reference's
// penalty rules works.
//
// deal.II evaluates this face-by-face with FEEvaluation / FEFaceEvaluation (sum
// factorisation over SIMD lanes) and scatters both sides of a face.  Here the
// loop is CELL-centric: each cell computes its own rows only, reading the six
// neighbours' coefficients (every interior face is visited from both sides), so
// there are no atomics and the result is deterministic.  On Cartesian cells all
// geometry factors are per-direction scalars and the operator is a sum of
// Kronecker products of the 1-D matrices
//   Mh = int l_i l_j,  Sh = int l_i' l_j'   (the Gauss rule; cell and face rules coincide)
//   e0/e1 = l_i(0)/l_i(1),  d0/d1 = l_i'(0)/l_i'(1)
// applied LINE by line: one thread owns one line of N1 DoFs along one direction,
// keeps it in registers and multiplies by the tables as constant-bank operands
// (the first version, one thread per DoF with every operand fetched from shared
// memory, was shared-memory-bandwidth bound at 2.6x the time).
//
// Algorithmic traffic 16 B per DoF (read src once, write dst once; neighbour reads
// are L1/L2 hits) plus 64 B per (cell, direction) of folded stencil record.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"
#include "pd_fine_cell.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
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

    // 1-D tables, passed BY VALUE as kernel parameters: every index is a compile-time constant
    // after unrolling, so each entry is a constant-bank operand of its DFMA (no load at all).
    template <int N1>
    struct FineTables
    {
      double Mh[N1 * N1], Sh[N1 * N1], e[2][N1], d[2][N1];
    };

    // What one line of DoFs along direction d of one cell needs to know about the cell
    // (a = vol/h_d the face area, sg the face penalty): the geometry record built at set-up
    //   cV = a/h_d,  P[s] = a sg_s,  Q[s] = a / (2 h_d(neighbour_s))   (0 at the boundary)
    struct alignas(16) FineGeo
    {
      int32_t nb[2]; // neighbour block along -d / +d, -1 = boundary
      double  cV;
      double  P[2];
      double  Q[2];
    };
    static_assert(sizeof(FineGeo) == 6 * sizeof(double), "records travel in a double buffer");

    // ... and the record the kernel reads: the operator's coefficient and term flags folded
    // in (k_fine_fold, once per pd_set_operator), so the apply kernel has no per-face logic:
    //   cVol = c cV (volume term on),  cD[s] = c cV/2 (interior) | c cV (boundary),
    //   P[s], Q[s] scaled by c; everything of a switched-off face term is zero.
    struct alignas(16) FineRec
    {
      int32_t nb[2];
      double  cVol;
      double  cD[2];
      double  P[2];
      double  Q[2];
    };
    static_assert(sizeof(FineRec) == 64, "four 16-byte loads");

    __global__ void
    k_fine_fold(const FineGeo *geo, FineRec *rec, const int64_t n, const double c, const uint32_t flags)
    {
      const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n)
        return;
      const FineGeo g = geo[i];
      FineRec       r;
      r.cVol = (flags & PD_ASSEMBLE_VOLUME) ? c * g.cV : 0.;
      for (int s = 0; s < 2; ++s)
        {
          const bool inner = g.nb[s] >= 0;
          const bool on    = (flags & (inner ? PD_ASSEMBLE_INTERIOR : PD_ASSEMBLE_BOUNDARY)) != 0;
          r.nb[s]          = g.nb[s];
          r.cD[s]          = on ? c * (inner ? 0.5 * g.cV : g.cV) : 0.;
          r.P[s]           = on ? c * g.P[s] : 0.;
          r.Q[s]           = on ? c * g.Q[s] : 0.;
        }
      rec[i] = r;
    }

    template <int N1>
    struct FineArgs
    {
      FineTables<N1> T;
      const FineRec *rec;  // [n_cells][DIM], block order
      const double  *vol;  // [n_cells]
      const double  *zero; // N zeros: what a missing neighbour reads
      const double  *x;
      double        *y;
      int32_t        n_cells;
      const int32_t *list; // optional: the cells to process (interior / boundary split of a sharded apply)
      double         mass; // 0 when the volume term is off
      int            add;
    };

    constexpr int
    pow2_at_least(const int v)
    {
      int g = 4;
      while (g < v)
        g *= 2;
      return g;
    }

    // On a Cartesian cell the SIP operator is
    //   sum_d  M (x) ... (x) L_d (x) ... (x) M   +  f vol  M (x) M (x) M
    // where L_d is the 1-D DG stencil along d acting on the cell's own line of DoFs and on
    // the lines of its two neighbours along d (the face parts are rank one / two: outer
    // products of the end values e_s and end derivatives d_s), and M the 1-D mass (cell and
    // face Gauss rules coincide).  The mass term is folded into the last direction's stencil,
    //   M (x) M (x) (L_z + f vol M).
    //
    // One thread per LINE: thread (d, j) of a cell owns line j along direction d.  It reads
    // its own line and the two neighbour lines straight from the vector (L1/L2 hits for the
    // neighbours), applies L_d entirely in registers with the tables as constant operands,
    // and writes the N1 results to the cell's work array W_d in shared memory; the DIM-1
    // remaining mass contractions of W_d are again one line per thread, in place, and the
    // DoF lanes finally sum W_0..W_{DIM-1}.  GS threads per cell (a power of two covering
    // both the DIM*N1^(DIM-1) lines and the N DoFs), 256/GS cells per block, grid-stride.
    template <int DIM, int DEG, int MINB>
    __global__ void __launch_bounds__(256, MINB) k_fine_sip(const __grid_constant__ FineArgs<DEG + 1> A)
    {
      constexpr int N1  = DEG + 1;
      constexpr int N   = ipow_(N1, DIM);
      constexpr int NL  = N / N1;   // lines per direction
      constexpr int NT  = DIM * NL; // line tasks per cell
      constexpr int GS  = pow2_at_least(N > NT ? N : NT);
      constexpr int CPB = 256 / GS;

      constexpr int NP  = N | 1; // odd row length: the DIM work arrays of a cell start on different banks
      __shared__ double sW[CPB][DIM][NP];

      const int  slot = threadIdx.x / GS, l = threadIdx.x % GS;
      // a cell's work arrays are private to its GS threads: warp-level ordering when a cell
      // is (part of) one warp, a named barrier per cell slot otherwise
      auto group_sync = [slot] {
        if constexpr (GS <= 32)
          __syncwarp();
        else
          asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(GS) : "memory");
      };

      const bool task_ok = l < NT;
      const int  d = task_ok ? l / NL : 0, j = l % NL;
      // line j along direction dd: first DoF and stride
      auto line = [](const int dd, const int jj, int &base, int &stride) {
        stride = dd == 0 ? 1 : (dd == 1 ? N1 : N1 * N1);
        base   = dd == 0 ? jj * N1 : (dd == 1 ? (jj % N1) + (jj / N1) * N1 * N1 : jj);
      };
      // loop-invariant offsets: the line in the vector / in W_d, and the lines of the mass passes
      int off[N1], offk[DIM > 1 ? DIM - 1 : 1][N1];
      {
        int base, stride;
        line(d, j, base, stride);
#pragma unroll
        for (int t = 0; t < N1; ++t)
          off[t] = base + t * stride;
#pragma unroll
        for (int k = 1; k < DIM; ++k)
          {
            line((d + k) % DIM, j, base, stride);
#pragma unroll
            for (int t = 0; t < N1; ++t)
              offk[k - 1][t] = base + t * stride;
          }
      }
      double *const  wd     = &sW[slot][d][0];
      const bool     mass_on = A.mass != 0.;
      const double   mass_d = d == DIM - 1 ? A.mass : 0.;
      const FineRec *recp   = A.rec + d;

      // the neighbour indices of the NEXT cell are fetched one iteration ahead, so the two
      // dependent long-latency steps (record -> neighbour lines) overlap across iterations
      // work index -> cell (identity unless a cell list is given)
      auto cell_of = [&](const int idx) { return idx < A.n_cells ? (A.list ? A.list[idx] : idx) : -1; };
      int       idx  = blockIdx.x * CPB + slot;
      const int step = gridDim.x * CPB;
      int       cell = cell_of(idx), cell_next = cell_of(idx + step);
      int2      nbn  = make_int2(-1, -1);
      if (cell >= 0 && task_ok)
        nbn = *reinterpret_cast<const int2 *>(recp + (int64_t)cell * DIM);

      for (int c0 = blockIdx.x * CPB; c0 < A.n_cells;
           c0 += step, idx += step, cell = cell_next, cell_next = cell_of(idx + step))
        {
          const bool cell_ok = cell >= 0;
          double     out[N1];
#pragma unroll
          for (int i = 0; i < N1; ++i)
            out[i] = 0.;
          const int2 nb = nbn;
          if (cell_next >= 0 && task_ok)
            nbn = *reinterpret_cast<const int2 *>(recp + (int64_t)cell_next * DIM);
          if (cell_ok && task_ok)
            {
              // all loads are unconditional (a missing neighbour reads zeros), so they are
              // issued back to back
              const double *xc  = A.x + (int64_t)cell * N;
              const double *xn0 = nb.x >= 0 ? A.x + (int64_t)nb.x * N : A.zero;
              const double *xn1 = nb.y >= 0 ? A.x + (int64_t)nb.y * N : A.zero;
              double        u[N1], nv[2][N1];
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
                  u[t]     = xc[off[t]];
                  nv[0][t] = xn0[off[t]];
                  nv[1][t] = xn1[off[t]];
                }
              const FineRec r = recp[(int64_t)cell * DIM];
              // traces of the own line and of the neighbours' facing ends
              double tu[2] = {0., 0.}, du[2] = {0., 0.}, tn[2] = {0., 0.}, dn[2] = {0., 0.};
#pragma unroll
              for (int t = 0; t < N1; ++t)
                {
#pragma unroll
                  for (int s = 0; s < 2; ++s)
                    {
                      tu[s] += A.T.e[s][t] * u[t];
                      du[s] += A.T.d[s][t] * u[t];
                      tn[s] += A.T.e[1 - s][t] * nv[s][t];
                      dn[s] += A.T.d[1 - s][t] * nv[s][t];
                    }
                }
              // av = P [u] - sn (cD du + Q dn),  bv = -sn cD [u]   (sn = -1 / +1)
              const double j0 = tu[0] - tn[0], j1 = tu[1] - tn[1];
              const double a0 = r.P[0] * j0 + (r.cD[0] * du[0] + r.Q[0] * dn[0]), b0 = r.cD[0] * j0;
              const double a1 = r.P[1] * j1 - (r.cD[1] * du[1] + r.Q[1] * dn[1]), b1 = -r.cD[1] * j1;
#pragma unroll
              for (int i = 0; i < N1; ++i)
                {
                  double sv = 0.;
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    sv += A.T.Sh[i * N1 + t] * u[t];
                  out[i] = r.cVol * sv + A.T.e[0][i] * a0 + A.T.e[1][i] * a1 + A.T.d[0][i] * b0 + A.T.d[1][i] * b1;
                }
              if (mass_on) // kernel-uniform
                {
                  const double mv = mass_d * A.vol[cell];
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                    {
                      double sm = 0.;
#pragma unroll
                      for (int t = 0; t < N1; ++t)
                        sm += A.T.Mh[i * N1 + t] * u[t];
                      out[i] += mv * sm;
                    }
                }
            }
          group_sync(); // the previous cell's final sum has read this slot's work arrays
          if (task_ok)
            {
#pragma unroll
              for (int i = 0; i < N1; ++i)
                wd[off[i]] = out[i];
            }
          group_sync();
          // ---- masses along the other directions, in place, one line per thread
#pragma unroll
          for (int k = 1; k < DIM; ++k)
            {
              if (task_ok)
                {
                  double v[N1];
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    v[t] = wd[offk[k - 1][t]];
#pragma unroll
                  for (int i = 0; i < N1; ++i)
                    {
                      double sm = 0.;
#pragma unroll
                      for (int t = 0; t < N1; ++t)
                        sm += A.T.Mh[i * N1 + t] * v[t];
                      wd[offk[k - 1][i]] = sm;
                    }
                }
              group_sync();
            }
          if (cell_ok && l < N)
            {
              double acc = 0.;
#pragma unroll
              for (int dd = 0; dd < DIM; ++dd)
                acc += sW[slot][dd][l];
              double *yp = A.y + (int64_t)cell * N + l;
              *yp        = A.add ? *yp + acc : acc;
            }
        }
    }

    // ---------------------------------------------------------------------------------------
    // The tiled kernel: ONE CELL PER THREAD PAIR, coefficients staged in shared memory.
    //
    // ncu on k_fine_sip (profiles/ncu_r01_fine_sip_summary.txt) shows the line-per-thread kernel
    // bound by the L1 data pipe (l1tex__data_pipe_lsu_wavefronts 92 %: 61 global + 73 shared
    // wavefronts per DGQ2 cell, every 8-byte access of a line a separate wavefront share), with
    // HBM at 13 % and the FP64 pipe at 29 %.  Here a CTA takes TILE consecutive cells and copies
    // their coefficients and those of the neighbours outside the tile (the "halo", tile plan of
    // pd_fine_cell.hpp) into shared memory with TMA bulk copies (cp.async.bulk: no registers, no
    // LSU wavefronts, everything in flight at once, completion on one mbarrier).  Two threads per
    // cell then apply the whole operator in registers (pd::fine::cell_lines, cell_mass; the lines
    // are split between two roles so that 16 warps are resident per SM, the partial sums meet
    // once through the cell's own row): their reads are row reads of an [cell][row] array with
    // an odd row length -> bank-conflict free, 16 useful doubles per wavefront.  Results go back through the
    // own rows so that the global stores are coalesced.
    // Measured (64^3 DGQ2, Morton order, profiles/ncu_r01_fine_tile_summary.txt): 31 shared + ~6
    // global L1 wavefronts per cell instead of 134, 0.134 -> 0.070 ms.
    // ---------------------------------------------------------------------------------------
    template <int N1>
    struct TileArgs
    {
      fine::TileTables<N1> T;
      const FineRec       *rec;
      const double        *vol;
      const double        *x; // 16-byte aligned (checked at launch): the halo rows are filled by bulk copies
      double              *y;
      const int32_t       *seq;      // cells in processing order (nullptr: 0 .. n_seq-1)
      const int32_t       *tile_first; // [n_tiles + 1]: first sequence entry of a tile (at most FINE_TILE entries)
      const int32_t       *tile_base;  // [n_tiles]: first cell of a tile whose cells are consecutive cell numbers, else -1
      const int32_t       *tile_ptr;   // [n_tiles + 1] into halo
      const int32_t       *halo;
      const uint16_t      *noff; // [n_seq][2 DIM]: first double of the neighbour's coefficients in shared memory
      int32_t              n_seq, max_halo, zoff;
      double               mass;
      int                  add;
    };

    __device__ __forceinline__ void
    cp_async8(void *smem, const void *gmem)
    {
      const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
    }
    __device__ __forceinline__ void
    cp_async16(void *smem, const void *gmem)
    {
      const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
    }
    // arrive on the mbarrier once every cp.async this thread has issued so far has landed (the arrival is part of the
    // barrier's initial count: .noinc)
    __device__ __forceinline__ void
    cp_async_mbar_arrive(uint64_t *bar)
    {
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(b) : "memory");
    }
    __device__ __forceinline__ void
    cp_async_wait_all()
    {
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    // bulk copy global -> shared through the TMA unit (no LSU wavefronts, no registers), completion on an mbarrier
    __device__ __forceinline__ void
    bulk_g2s(void *smem, const void *gmem, const uint32_t bytes, uint64_t *bar)
    {
      const unsigned s = (unsigned)__cvta_generic_to_shared(smem), b = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s), "l"(gmem),
                   "r"(bytes), "r"(b)
                   : "memory");
    }
    __device__ __forceinline__ void
    mbar_init(uint64_t *bar, const uint32_t count)
    {
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void
    mbar_arrive_expect(uint64_t *bar, const uint32_t bytes)
    {
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
      if (bytes)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
      else
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
    }
    __device__ __forceinline__ bool
    mbar_try_wait(uint64_t *bar, const uint32_t parity)
    {
      const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
      uint32_t       ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(b), "r"(parity)
                   : "memory");
      return ok != 0;
    }

    constexpr int FINE_TILE = 64, FINE_TILE_THREADS = 2 * FINE_TILE; // two threads (roles) per cell

    constexpr size_t
    round16(const size_t v)
    {
      return (v + 15) / 16 * 16;
    }
    // shared memory of a tile CTA: coefficients (own | halo rows | zeros, pd_fine_cell.hpp) | records
    // [TILE][DIM*64+16] | mbarrier
    constexpr size_t
    tile_values_bytes(const int n, const int max_halo)
    {
      return round16(((size_t)FINE_TILE * fine::own_row(n) + (size_t)max_halo * fine::halo_row(n) + n + 1) * sizeof(double));
    }
    constexpr size_t
    tile_smem_bytes(const int dim, const int n, const int max_halo)
    {
      return tile_values_bytes(n, max_halo) + (size_t)FINE_TILE * (dim * 64 + 16) + 16;
    }

    template <int DIM, int DEG>
    __global__ void __launch_bounds__(FINE_TILE_THREADS, 4) k_fine_tile(const __grid_constant__ TileArgs<DEG + 1> A)
    {
      constexpr int N1  = DEG + 1;
      constexpr int N   = ipow_(N1, DIM);
      constexpr int NFC = 2 * DIM;
      constexpr int RH  = fine::halo_row(N);
      constexpr int RO  = fine::own_row(N); // distance of the own rows (N odd: the global layout)
      constexpr int RS  = DIM * 64 + 16; // record row of a cell, padded: 16-byte reads of consecutive threads hit distinct banks
      constexpr int SPW = 32 / N;        // cells a warp copies per step
      constexpr int NW  = FINE_TILE_THREADS / 32;
      static_assert(N <= 32, "one warp step copies at least one cell");

      extern __shared__ __align__(16) unsigned char smem[];
      double        *S   = reinterpret_cast<double *>(smem);
      unsigned char *sR  = smem + tile_values_bytes(N, A.max_halo);
      uint64_t      *bar = reinterpret_cast<uint64_t *>(sR + FINE_TILE * RS);

      const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
      const int s0 = A.tile_first[blockIdx.x], n_own = A.tile_first[blockIdx.x + 1] - s0;
      const int h0 = A.tile_ptr[blockIdx.x], nh = A.tile_ptr[blockIdx.x + 1] - h0;
      // cells of a tile: consecutive numbers from tile_base (one load per tile, one bulk copy stages them), else the list
      const int tbase   = A.seq ? A.tile_base[blockIdx.x] : s0; // (no list: the tile is the run s0 .. s0 + n_own - 1)
      auto      cell_of = [&](const int i) { return tbase >= 0 ? tbase + i : A.seq[s0 + i]; };

      if (tid == 0)
        mbar_init(bar, FINE_TILE_THREADS);
      if (tid < N)
        S[A.zoff + tid] = 0.;
      // thread = (role, cell): role 0 (warps 0-1) and role 1 (warps 2-3) split the lines of a cell
      const int  role = tid / FINE_TILE, ci = tid % FINE_TILE;
      const bool mine = ci < n_own;
      uint32_t   no[NFC];
      double     mv = 0.;
      if (mine)
        {
          const uint16_t *np = A.noff + (size_t)(s0 + ci) * NFC;
#pragma unroll
          for (int f = 0; f < NFC; ++f)
            no[f] = np[f];
          if (role == 0 && A.mass != 0.)
            mv = A.mass * A.vol[cell_of(ci)];
        }
      __syncthreads(); // the mbarrier is initialised
      const int  first_cell = cell_of(0);
      const bool contiguous = tbase >= 0;
      // ---- stage the coefficients (own + halo) and the own cells' records; everything is in flight at once
      uint32_t bytes = 0;
      for (int r = tid; r < nh; r += FINE_TILE_THREADS)
        { // halo row r: one bulk copy from the 16-byte boundary below the cell (+ its last double when that leaves 8 bytes)
          const int64_t  first = (int64_t)A.halo[h0 + r] * N, a0 = first & ~(int64_t)1;
          const uint32_t total = (uint32_t)(N + (first - a0)) * 8, sz = total & ~15u;
          double        *dst   = S + FINE_TILE * RO + r * RH;
          bulk_g2s(dst, A.x + a0, sz, bar);
          bytes += sz;
          if (total != sz)
            cp_async8(dst + sz / 8, A.x + a0 + sz / 8);
        }
      if (mine && role == 0)
        {
          bulk_g2s(sR + ci * RS, A.rec + (int64_t)cell_of(ci) * DIM, DIM * 64, bar);
          bytes += DIM * 64;
        }
      // the own cells are one aligned contiguous range (always along the curve's own numbering; with a cell list --
      // the interior / boundary split of a sharded apply -- whenever the tile's cells are consecutive)
      if (RO == N && contiguous && n_own == FINE_TILE && ((int64_t)first_cell * N) % 2 == 0)
        {
          if (tid == 0)
            {
              bulk_g2s(S, A.x + (int64_t)first_cell * N, FINE_TILE * N * 8, bar);
              bytes += FINE_TILE * N * 8;
            }
        }
      else
        {
          const int  sub = lane / N, e = lane % N;
          const bool copier = sub < SPW;
#pragma unroll 2
          for (int i = warp * SPW + sub; i < n_own; i += NW * SPW)
            if (copier)
              cp_async8(S + i * RO + e, A.x + (int64_t)cell_of(i) * N + e);
        }
      mbar_arrive_expect(bar, bytes);
      cp_async_wait_all();
      for (int spin = 0; !mbar_try_wait(bar, 0); ++spin)
        if (spin > (1 << 22))
          __trap(); // a bulk copy that never completes would otherwise hang the device
      __syncthreads();
      // ---- the lines of one cell in registers, split between its two threads
      double        acc[N];
      double *const own = S + ci * RO;
      if (mine)
        {
          const double *nbp[NFC];
#pragma unroll
          for (int f = 0; f < NFC; ++f)
            nbp[f] = S + no[f];
          const unsigned char *myrec = sR + ci * RS;
          auto                 nbv   = [&](const int d, const int s, const int k) { return nbp[2 * d + s][k]; };
          auto                 coef  = [&](const int d) {
            const unsigned char *r  = myrec + d * 64;
            const double2        cd = *reinterpret_cast<const double2 *>(r + 16);
            const double2        pp = *reinterpret_cast<const double2 *>(r + 32);
            const double2        qq = *reinterpret_cast<const double2 *>(r + 48);
            fine::LineCoef       c;
            c.cVol  = *reinterpret_cast<const double *>(r + 8);
            c.cD[0] = cd.x, c.cD[1] = cd.y;
            c.P[0] = pp.x, c.P[1] = pp.y;
            c.Q[0] = qq.x, c.Q[1] = qq.y;
            return c;
          };
          if (role == 0) // warp-uniform
            {
#pragma unroll
              for (int k = 0; k < N; ++k)
                acc[k] = mv * own[k];
              fine::cell_lines<DIM, N1>(A.T, 0, own, nbv, coef, acc);
            }
          else
            {
#pragma unroll
              for (int k = 0; k < N; ++k)
                acc[k] = 0.;
              fine::cell_lines<DIM, N1>(A.T, 1, own, nbv, coef, acc);
            }
        }
      __syncthreads(); // every read of the staged coefficients is done: the own rows become the exchange / output staging
      if (mine && role == 1)
        {
#pragma unroll
          for (int k = 0; k < N; ++k)
            own[k] = acc[k];
        }
      __syncthreads();
      if (mine && role == 0)
        {
#pragma unroll
          for (int k = 0; k < N; ++k)
            acc[k] += own[k];
          fine::cell_mass<DIM, N1>(A.T, acc);
#pragma unroll
          for (int k = 0; k < N; ++k)
            own[k] = acc[k];
        }
      __syncthreads();
      {
        const int  sub = lane / N, e = lane % N;
        const bool copier = sub < SPW;
#pragma unroll 2
        for (int i = warp * SPW + sub; i < n_own; i += NW * SPW)
          if (copier)
            {
              double      *yp = A.y + (int64_t)cell_of(i) * N + e;
              const double v  = S[i * RO + e];
              *yp             = A.add ? *yp + v : v;
            }
      }
    }

    // ---------------------------------------------------------------------------------------
    // The pipelined kernel: the tiles of k_fine_tile STREAMED through a ring of shared-memory stages.
    //
    // ncu on k_fine_tile (profiles/ncu_r01_fine_tile_summary.txt): warps active 24 %, long scoreboard the top
    // stall -- a CTA lives through three dependent steps (tile metadata -> bulk copies -> arithmetic) and only
    // four CTAs per SM overlap them.  Here the CTAs are persistent and warp-specialised:
    //  * one PRODUCER warp walks the CTA's tiles (tile = blockIdx + k gridDim, so the tiles in flight at any time
    //    are one contiguous run of the curve and share their halos in L2), reads the plan, and issues the bulk
    //    copies of tile k + NS - 1 (own cells: one copy; halo cells: one copy each; the 16-bit neighbour offsets:
    //    one copy) while the consumers work on tile k; completion on the stage's `full` mbarrier;
    //  * four CONSUMER warps (two threads per cell, pd::fine::cell_lines / cell_mass as in k_fine_tile) wait on
    //    `full`, apply the operator in registers, leave the result in the own rows, fence towards the async proxy
    //    and arrive on the stage's `done` mbarrier;
    //  * the producer then sends the own rows to y with ONE bulk store (cp.async.bulk shared -> global; the
    //    add variant is cp.reduce.async.bulk .add.f64: every y entry has one writer, so it stays deterministic),
    //    waits until the store has read the stage and refills it.
    // No per-cell records: the kernel is for UNIFORM meshes (all cells the same box, one penalty per direction and
    // face kind; pd_handle::mf_uniform), where the stencil coefficients are kernel constants and a face is a
    // boundary face iff its neighbour offset points at the zero row.  Everything else runs k_fine_tile.
    // Needs: every tile one contiguous run of cells starting on a 16-byte boundary with a multiple of four
    // cells (FineTiles::stream_ok), N odd (own rows as in the vector), x and y 16-byte aligned.
    // ---------------------------------------------------------------------------------------
    constexpr int FINE_MAX_STAGES = 8; // (and at most FINE_TILE_THREADS halo cells per tile: one row per thread)
    template <int DIM, int N1>
    struct StreamArgs
    {
      fine::DenseTables<DIM, N1> T;
      const double              *x;
      double                    *y;
      const int32_t             *tile_first; // [n_tiles + 1] first sequence entry of a tile (at most FINE_TILE entries)
      const int32_t             *tile_base;  // [n_tiles] first cell of a tile: its cells are consecutive numbers
      const int32_t             *halo_pad;   // [n_tiles][max_halo] halo cells of a tile, padded with -1
      const uint16_t            *noff;       // [n_seq][2 DIM] (+ 16 bytes of slack)
      int64_t                    x_len;     // doubles that may be read from x
      int32_t                    n_tiles, max_halo, zoff, n_stages;
      int                        add;
      // fused sharded apply (ghost_src != nullptr): halo cells >= np_own are read from the owner's export buffer over
      // NVLink, tiles >= first_ghost_tile wait (once per warp) until the owners' epoch flags have reached this
      // rank's exchange epoch
      const double *const       *ghost_src;
      int64_t                    parity_stride;
      const unsigned long long  *epochs;
      volatile unsigned long long *flags, *error_word;
      const int32_t             *owners;
      int32_t                    n_owners, np_own, first_ghost_tile;
      int32_t                    refill_late; // PD_FINE_REFILL: 0 own rows refilled at the top of the next tile, 1 (default) after its line phase
    };
    // shared memory of a stage: coefficients (own rows | halo rows | zeros, all n doubles apart: StreamPlan) | neighbour offsets
    // (own rows with two doubles of slack for the tile's 16-byte phase | halo rows | zero row: fine::StreamPlan)
    constexpr size_t
    stream_values_bytes(const int n, const int max_rows)
    {
      return round16(((size_t)FINE_TILE * n + 2 + (size_t)(max_rows + 1) * n + 1) * sizeof(double));
    }
    // ... | neighbour offsets (copied from the 16-byte boundary below the tile's first entry) | n_own, phase, offset shift, first cell
    constexpr size_t
    stream_noff_bytes(const int dim)
    {
      return round16((size_t)FINE_TILE * 2 * dim * 2) + 16;
    }
    constexpr size_t
    stream_stage_bytes(const int dim, const int n, const int max_rows)
    {
      return stream_values_bytes(n, max_rows) + stream_noff_bytes(dim) + 16;
    }
    constexpr size_t
    stream_smem_bytes(const int dim, const int n, const int max_halo, const int stages)
    {
      return stages * stream_stage_bytes(dim, n, max_halo) + FINE_MAX_STAGES * sizeof(uint64_t) + 32;
    }
    __device__ __forceinline__ void
    mbar_wait(uint64_t *bar, const uint32_t parity)
    {
      for (int spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1 << 22))
          __trap(); // a copy that never completes / a lost arrival would otherwise hang the device
    }
    __device__ __forceinline__ void
    group_bar_sync(const int id)
    {
      asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(FINE_TILE_THREADS) : "memory");
    }

    // NG groups of FINE_TILE_THREADS threads; a group works on one tile at a time (thread = (role, cell) as in
    // k_fine_tile), the CTA's tiles are handed out in order to whichever group is free.  Per stage one mbarrier `full`
    // (everything of the tile has landed).
    //  * halo rows: gathered by the group itself -- as soon as it is done reading stage s (tile l), its threads issue the
    //    16-byte cp.async chunks of tile l + NS into the same stage, one halo row per thread, completion counted by
    //    cp.async.mbarrier.arrive on full[s]; the gather overlaps the rest of the group's tile (exchange, mass passes)
    //    and the other groups' arithmetic.  Neither per-cell bulk copies (measured: the TMA unit of an SM retires only
    //    ~40 of these 224-byte copies per microsecond) nor one gathering warp (measured: 6.5 us per tile, too few
    //    loads in flight) keeps up.
    //  * own rows in, neighbour offsets in, results out: one bulk copy each, issued by the group's first thread; the
    //    refill of the own rows waits for the bulk store to have read them (checked at the top of the group's next
    //    tile, when it no longer costs anything).
    // No extra warp: registers are allocated to a CTA in units of four warps, and a 17th warp would cost every thread
    // of four groups 32 registers.
    constexpr int
    stream_max_regs(const int ng)
    {
      return ng == 2 ? 128 : (ng == 3 ? 168 : (ng == 4 ? 128 : 96));
    }
    // GENERAL = false: every tile is FINE_TILE cells starting on a 16-byte boundary of the vector and there are no
    // ghost cells to fetch from peers (one GPU, Morton-numbered mesh): the per-tile bookkeeping is compile-time.
    template <int DIM, int DEG, int NG, bool GENERAL>
    __global__ void __launch_bounds__(NG *FINE_TILE_THREADS) __maxnreg__(stream_max_regs(NG))
      k_fine_stream(const __grid_constant__ StreamArgs<DIM, DEG + 1> A)
    {
      constexpr int N1  = DEG + 1;
      constexpr int N   = ipow_(N1, DIM);
      constexpr int NFC = 2 * DIM;
      constexpr int RO  = N; // own and halo rows as in the vector (N odd), fine::StreamPlan
      constexpr int HB  = FINE_TILE * N + 2; // first halo row
      constexpr uint32_t NOFFB = (uint32_t)stream_noff_bytes(DIM);
      static_assert(N % 2 == 1, "packed rows");

      extern __shared__ __align__(16) unsigned char smem[];
      const int      NS          = A.n_stages;
      const uint32_t vb          = (uint32_t)stream_values_bytes(N, A.max_halo);
      const uint32_t stage_bytes = vb + NOFFB + 16;
      uint64_t      *full        = reinterpret_cast<uint64_t *>(smem + (size_t)NS * stage_bytes);
      int           *next_tile   = reinterpret_cast<int *>(full + FINE_MAX_STAGES); // [0]: next unclaimed tile of the CTA; [1 + g]: group g's claim

      // (the warp index through a broadcast shuffle: the compiler then knows that everything derived from it -- the
      // warp's group, its role -- is warp-uniform and keeps the 1-D matrices in uniform registers)
      const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
      if (tid == 0)
        {
          for (int s = 0; s < NS; ++s)
            mbar_init(full + s, FINE_TILE_THREADS + 2); // the gathering group's threads + the first thread's single doubles + the bulk copies' transaction bytes
          next_tile[0] = NG; // (the first NG tiles are taken by group number)
        }
      if (tid < N)
        for (int s = 0; s < NS; ++s)
          reinterpret_cast<double *>(smem + (size_t)s * stage_bytes)[A.zoff + tid] = 0.;
      __syncthreads();
      const int my_n = (int)blockIdx.x < A.n_tiles ? (A.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      auto      tile_of = [&](const int l) { return (int)blockIdx.x + l * (int)gridDim.x; };

      constexpr int WPG = FINE_TILE_THREADS / 32, WPR = FINE_TILE / 32; // warps per group / per role
      const int     group = warp / WPG, role = (warp % WPG) / WPR, ci = (warp % WPR) * 32 + lane;
      const int     gt      = (warp % WPG) * 32 + lane; // thread of the group: gathers halo row gt
      const bool    elected = gt == 0;                  // issues the group's bulk copies
      // own rows + neighbour offsets of tile l into its stage.  The run of own coefficients [a, a + L) keeps its 16-byte
      // phase (element e at S + 2 - par + e, par = a & 1): the aligned middle is one bulk copy to S + 2, a first / last
      // double that sticks out an 8-byte cp.async; the offsets are copied from the 16-byte boundary below the tile's
      // first entry.  m = {first sequence entry, cells, first cell} of the tile (fetched a tile ahead).
      auto tile_meta = [&](const int l) {
        const int tile = tile_of(l);
        if constexpr (!GENERAL)
          return make_int3(tile * FINE_TILE, FINE_TILE, l < my_n ? A.tile_base[tile] : 0);
        else
          return l < my_n ? make_int3(A.tile_first[tile], A.tile_first[tile + 1] - A.tile_first[tile], A.tile_base[tile]) :
                            make_int3(0, 0, 0);
      };
      auto load_own = [&](const int l, const int3 m) {
        const int      s  = l % NS;
        unsigned char *st = smem + (size_t)s * stage_bytes;
        double        *S  = reinterpret_cast<double *>(st);
        const int      s0 = m.x, n_own = m.y, c0 = m.z;
        const int64_t  a = (int64_t)c0 * N, L = (int64_t)n_own * N;
        const int      par = GENERAL ? (int)(a & 1) : 0, tail = GENERAL ? (int)((a + L) & 1) : 0;
        const uint32_t mid = (uint32_t)(L - par - tail) * 8;
        bulk_g2s(S + 2, A.x + a + par, mid, full + s);
        if (par)
          cp_async8(S + 1, A.x + a);
        if (tail)
          cp_async8(S + 2 - par + L - 1, A.x + a + L - 1);
        cp_async_mbar_arrive(full + s);
        const uint32_t o = (uint32_t)s0 * NFC * 2, sh = o & 15u, len = (uint32_t)round16((size_t)sh + (size_t)n_own * NFC * 2);
        bulk_g2s(st + vb, reinterpret_cast<const unsigned char *>(A.noff) + (o - sh), len, full + s);
        int *meta = reinterpret_cast<int *>(st + vb + NOFFB);
        if constexpr (GENERAL)
          meta[0] = n_own, meta[1] = par, meta[2] = (int)sh;
        meta[3] = c0;
        mbar_arrive_expect(full + s, mid + len);
      };
      // The halo rows of tile l into its stage.  Row and cell have the same 16-byte alignment (fine::StreamPlan), so a
      // row is (N - 1) / 2 chunks of 16 bytes and one of 8 (the last of an even row, the first of an odd row).  N + 1
      // consecutive lanes take the chunks of an (even, odd) pair of rows, PPP pairs per instruction: the shared-memory
      // side of a cp.async costs one wavefront per contiguous 128 bytes, so lanes must write NEIGHBOURING chunks
      // (one row per lane, the first version, took 30 wavefronts per instruction -- 43 % of all the kernel's
      // shared-memory wavefronts).  The warps of the group take the passes round robin; lane q RPP + k of a warp has
      // prefetched the cell of row k of the warp's pass q.
      constexpr int PPP = 32 / (N + 1), RPP = 2 * PPP, HALF = (N + 1) / 2; // pairs / rows per pass, chunks per row
      static_assert(PPP >= 1, "a pair of rows fits a warp");
      const int  gw = warp % WPG; // warp of the group
      const int  pair = lane / (N + 1), jj = lane % (N + 1), odd = jj >= HALF ? 1 : 0, chunk = odd ? jj - HALF : jj;
      const bool gathers = lane < PPP * (N + 1);
      const bool small   = odd ? chunk == 0 : chunk == HALF - 1;                          // the 8-byte chunk of the row
      const int  coff    = odd ? (chunk == 0 ? 0 : 2 * chunk - 1) : 2 * chunk;           // first double of the chunk
      const int  n_pass  = (A.max_halo + RPP - 1) / RPP;                                 // passes of the group
      bool       ghosts_ready = A.ghost_src == nullptr;
      int64_t    ghost_shift  = 0;
      constexpr int MAXQ = 32 / RPP; // passes per warp at most (the lanes of a warp hold 32 rows)
      // Per pass a lane needs one shuffle, one 64-bit multiply-add and the copy: its shared-memory destination (32-bit
      // address, the pass as an immediate offset), its chunk of the row and the kind of chunk are fixed; the copies are
      // predicated, not branched (the lanes of a warp differ in whether and what they copy).  Measured at 128^3 cells:
      // 0.306 ms against 0.320 ms with a rolled loop over generic pointers.
      const uint32_t dst_lane  = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(HB + (gw * RPP + 2 * pair + odd) * RO + coff) * 8u;
      const int      src_lane0 = 2 * pair + odd;
      const unsigned long long x_lane = reinterpret_cast<unsigned long long>(A.x) + (unsigned long long)coff * 8ull;
      constexpr uint32_t PASS_BYTES = (uint32_t)(WPG * RPP * RO) * 8u;
      const int          lane16 = gathers && !small ? 1 : 0, lane8 = gathers && small ? 1 : 0;
      auto               copy_chunk = [&](const uint32_t dst, const unsigned long long src, const int c16, const int c8) {
        asm volatile("{\n\t.reg .pred p16, p8;\n\t"
                     "setp.ne.s32 p16, %2, 0;\n\t"
                     "setp.ne.s32 p8, %3, 0;\n\t"
                     "@p16 cp.async.cg.shared.global [%0], [%1], 16;\n\t"
                     "@p8 cp.async.ca.shared.global [%0], [%1], 8;\n\t}" ::"r"(dst),
                     "l"(src), "r"(c16), "r"(c8)
                     : "memory");
      };
      // hc: what halo_src() returned for tile l -- lane q RPP + k holds the source of row k of this warp's pass q:
      // first coefficient (in doubles from A.x) of a cell of the vector, -2 - g for ghost cell g read from its owner
      // (fused sharded apply only), -1 for no row
      auto gather = [&](const int l, const int32_t hc) {
        const int      s = l % NS;
        const uint32_t d = dst_lane + (uint32_t)s * stage_bytes;
        __syncwarp(); // (the group's first thread may come from its refill: without this its warp runs the passes twice)
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
          if (gw + q * WPG < n_pass) // (warp-uniform)
            {
              const int32_t e = __shfl_sync(0xffffffffu, hc, src_lane0 + q * RPP);
              copy_chunk(d + q * PASS_BYTES, x_lane + (unsigned long long)(uint32_t)e * 8ull, e >= 0 ? lane16 : 0, e >= 0 ? lane8 : 0);
            }
        if constexpr (GENERAL)
          if (A.ghost_src != nullptr && tile_of(l) >= A.first_ghost_tile) // (uniform; fused sharded apply, tiles next to a cut)
            {
              if (!ghosts_ready)
                { // the first such tile of this warp: the owners must have published
                  const unsigned long long e = *A.epochs; // (bumped by this rank's publish, the kernel in front of this one)
                  if (lane < A.n_owners)
                    {
                      const long long t0 = clock64();
                      while (A.flags[A.owners[lane]] < e)
                        {
                          if (clock64() - t0 > 4000000000ll) // the peer is gone: report, do not hang
                            {
                              *A.error_word = 1ull;
                              break;
                            }
                          __nanosleep(64);
                        }
                    }
                  __threadfence_system();
                  __syncwarp();
                  ghost_shift  = (e & 1ull) ? A.parity_stride : 0;
                  ghosts_ready = true;
                }
              for (int q = 0; gw + q * WPG < n_pass; ++q)
                {
                  const int32_t e     = __shfl_sync(0xffffffffu, hc, (src_lane0 + q * RPP) & 31);
                  const bool    ghost = e < -1;
                  const double *src   = ghost ? A.ghost_src[-2 - e] + ghost_shift + coff : A.x;
                  copy_chunk(d + (uint32_t)q * PASS_BYTES, reinterpret_cast<unsigned long long>(src), ghost ? lane16 : 0, ghost ? lane8 : 0);
                }
            }
        cp_async_mbar_arrive(full + s);
      };
      auto halo_src = [&](const int l) {
        const int row = (gw + (lane / RPP) * WPG) * RPP + lane % RPP; // row lane % RPP of this warp's pass lane / RPP
        const int32_t c =
          (l < my_n && lane < MAXQ * RPP && row < A.max_halo) ? A.halo_pad[(size_t)tile_of(l) * A.max_halo + row] : -1;
        return c < 0 ? -1 : ((!GENERAL || c < A.np_own || A.ghost_src == nullptr) ? c * N : -2 - (c - A.np_own));
      };
      // the first NS tiles: tile l by group l % NG
      for (int l = group; l < NS && l < my_n; l += NG)
        {
          gather(l, halo_src(l));
          if (elected)
            load_own(l, tile_meta(l));
        }
      int  pend = -1; // (elected thread) tile whose bulk store may still be reading its stage
      int3 pend_meta = make_int3(0, 0, 0); // ... and the plan entries of the tile that refills it
      auto refill_own = [&]() {
        if (pend >= 0)
          {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (pend + NS < my_n)
              load_own(pend + NS, pend_meta);
            pend = -1;
          }
      };
      for (int it = group; it < my_n;)
        {
          const int      s  = it % NS;
          unsigned char *st = smem + (size_t)s * stage_bytes;
          double        *S  = reinterpret_cast<double *>(st);
          const int32_t  hc = halo_src(it + NS); // (on its way while the lines are computed)
          if (elected && A.refill_late == 0)
            refill_own();
          // (before BLOCKING on a `full` the group's first thread makes sure its own pending refill is out: two groups
          // waiting for each other's refills would otherwise never wake up)
          if (!mbar_try_wait(full + s, (uint32_t)(it / NS) & 1u))
            {
              if (elected && A.refill_late != 2)
                refill_own();
              mbar_wait(full + s, (uint32_t)(it / NS) & 1u);
            }
          // (a tile may hold fewer than FINE_TILE cells; the threads beyond run the same instruction stream on their
          // own rows and the zero row -- in range, never stored --, so nothing below is predicated)
          const int *const meta  = reinterpret_cast<const int *>(st + vb + NOFFB);
          const int        n_own = GENERAL ? meta[0] : FINE_TILE, par = GENERAL ? meta[1] : 0;
          double           acc[N];
          double *const    own = S + 2 - par + ci * RO;
          {
            const uint16_t *np = reinterpret_cast<const uint16_t *>(st + vb + (GENERAL ? meta[2] : 0)) + ci * NFC;
            const double   *nbp[NFC];
            bool            bnd[NFC];
#pragma unroll
            for (int f = 0; f < NFC; ++f)
              {
                const uint32_t o = (!GENERAL || ci < n_own) ? np[f] : (uint32_t)A.zoff;
                nbp[f]           = S + o;
                bnd[f]           = o == (uint32_t)A.zoff;
              }
            auto nbv = [&](const int d, const int sd, const int k) { return nbp[2 * d + sd][k]; };
#pragma unroll
            for (int k = 0; k < N; ++k)
              acc[k] = 0.;
            if (role == 0) // warp-uniform
              fine::cell_lines_dense<DIM, N1>(A.T, 0, own, nbv, bnd, acc);
            else
              fine::cell_lines_dense<DIM, N1>(A.T, 1, own, nbv, bnd, acc);
          }
          group_bar_sync(group + 1); // every read of the staged coefficients is done: the own rows become the exchange / output staging
          // The own rows of the stage this group used for its previous tile can be refilled now: its bulk store was
          // issued a whole line phase ago and has read them (waiting for it at the top of the tile cost the group's
          // first warp 40 % of its time).  No cycle: the refill a group waits for is always issued by a group working
          // on an earlier tile (more stages than groups).
          if (elected)
            {
              refill_own();
              pend_meta = tile_meta(it + NS);
            }
          if (it + NS < my_n)
            gather(it + NS, hc); // ... and the halo rows can take the next tile of this stage
          if (role == 1)
            {
#pragma unroll
              for (int k = 0; k < N; ++k)
                own[k] = acc[k];
            }
          if (elected)
            next_tile[1 + group] = atomicAdd(next_tile, 1);
          group_bar_sync(group + 1);
          const int it_next = next_tile[1 + group];
          if (role == 0)
            {
#pragma unroll
              for (int k = 0; k < N; ++k)
                acc[k] += own[k];
              fine::cell_mass<DIM, N1>(A.T, acc);
#pragma unroll
              for (int k = 0; k < N; ++k)
                own[k] = acc[k];
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the bulk store reads these rows
              asm volatile("bar.sync %0, %1;" ::"r"(NG + 1 + group), "n"(FINE_TILE) : "memory"); // the role's threads
              if (elected)
                { // the run [a, a + L) of y: the aligned middle by one bulk store, a first / last double that sticks out by hand
                  const int64_t  a = (int64_t)meta[3] * N, L = (int64_t)n_own * N;
                  const int      tail = GENERAL ? (int)((a + L) & 1) : 0;
                  double        *dst  = A.y + a + par;
                  const uint32_t sz = (uint32_t)(L - par - tail) * 8, src = (unsigned)__cvta_generic_to_shared(S + 2);
                  if (A.add)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(src),
                                 "r"(sz)
                                 : "memory");
                  else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(sz)
                                 : "memory");
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                  if (par)
                    A.y[a] = A.add ? A.y[a] + S[1] : S[1];
                  if (tail)
                    A.y[a + L - 1] = A.add ? A.y[a + L - 1] + S[2 - par + L - 1] : S[2 - par + L - 1];
                  pend = it;
                }
            }
          it = it_next;
        }
      if (elected)
        {
          refill_own();
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }

    // host: l_a(x), l_a'(x)
    void
    lagrange_host(const Basis1D &B, const int n1, const double x, double *L, double *dL)
    {
      for (int a = 0; a < n1; ++a)
        {
          double val = 1., der = 0.;
          for (int b = 0; b < n1; ++b)
            if (b != a)
              {
                const double t = x - B.node[b];
                der            = der * t + val;
                val            = val * t;
              }
          L[a]  = val * B.wprod[a];
          dL[a] = der * B.wprod[a];
        }
    }
  } // namespace

  // Recognise "every polytope is one axis-aligned cell" and build the per-cell
  // neighbour / penalty tables (block order) from the flattened interface list.
  void
  setup_fine_operator(pd_handle *h, const pd_mesh_desc &d)
  {
    h->mf_ready = false;
    if (h->n_subcells != h->np_own || h->nq1 != h->nq1f)
      return; // (the stencil form needs the cell and face Gauss rules to coincide, as in MatrixFree)
    const int dim = d.dim, vpc = 1 << dim, nfc = 2 * dim;
    for (int32_t p = 0; p < h->np_own; ++p)
      {
        const int32_t  c  = d.poly_subcell_idx[d.poly_subcell_ptr[p]];
        const double  *bb = d.bbox + (size_t)p * 2 * dim;
        for (int v = 0; v < vpc; ++v)
          {
            const double *x = d.verts + (size_t)d.cell_verts[(size_t)c * vpc + v] * dim;
            for (int k = 0; k < dim; ++k)
              if (x[k] != (((v >> k) & 1) ? bb[dim + k] : bb[k]))
                return; // not an axis-aligned box in reference orientation
          }
      }
    // extents for owned AND ghost cells (a neighbour's normal derivative needs its 1/h);
    // neighbour / penalty rows for owned cells only
    std::vector<double>  cell_h((size_t)h->np * dim), sigma((size_t)h->np_own * nfc, 0.);
    std::vector<int32_t> nbr((size_t)h->np_own * nfc, -1);
    std::vector<char>    seen((size_t)h->np_own * nfc, 0);
    for (int32_t p = 0; p < h->np; ++p)
      for (int k = 0; k < dim; ++k)
        cell_h[(size_t)d.dof_block[p] * dim + k] = d.bbox[(size_t)p * 2 * dim + dim + k] - d.bbox[(size_t)p * 2 * dim + k];
    for (int32_t f = 0; f < d.n_ifaces; ++f)
      {
        const int32_t a = d.iface_polyA[f], b = d.iface_polyB[f];
        for (int64_t s = d.iface_sub_ptr[f]; s < d.iface_sub_ptr[f + 1]; ++s)
          {
            const int     lf = d.sub_face[s];
            const int32_t ba = d.dof_block[a];
            nbr[(size_t)ba * nfc + lf]   = b >= 0 ? d.dof_block[b] : -1;
            sigma[(size_t)ba * nfc + lf] = d.sub_sigma[s];
            seen[(size_t)ba * nfc + lf]  = 1;
            if (b >= 0 && b < h->np_own)
              {
                const int32_t bb = d.dof_block[b];
                nbr[(size_t)bb * nfc + (lf ^ 1)]   = ba;
                sigma[(size_t)bb * nfc + (lf ^ 1)] = d.sub_sigma[s];
                seen[(size_t)bb * nfc + (lf ^ 1)]  = 1;
              }
          }
      }
    for (char s : seen)
      if (!s)
        return; // a cell face without an interface entry: not a conforming singleton mesh
    // uniform mesh?  (all cells -- owned and ghost -- the same box, one penalty per direction for the interior faces and
    // one per (direction, side) for the boundary faces, to two units in the last place): the pipelined kernel then
    // takes its stencil coefficients as kernel constants instead of per-cell records
    {
      auto &u    = h->mf_uniform;
      u          = pd_handle::FineUniform{};
      auto close = [](const double a, const double b) { return std::fabs(a - b) <= 4.5e-16 * std::max(std::fabs(a), std::fabs(b)); };
      bool ok    = h->np_own > 0;
      bool have_in[3] = {false, false, false}, have_bd[3][2] = {{false, false}, {false, false}, {false, false}};
      for (int k = 0; k < dim && ok; ++k)
        u.h[k] = cell_h[k];
      for (int32_t c = 0; c < h->np && ok; ++c)
        for (int k = 0; k < dim; ++k)
          ok = ok && close(cell_h[(size_t)c * dim + k], u.h[k]);
      for (int32_t c = 0; c < h->np_own && ok; ++c)
        for (int f = 0; f < nfc; ++f)
          {
            const double sg = sigma[(size_t)c * nfc + f];
            const bool   in = nbr[(size_t)c * nfc + f] >= 0;
            double      &v  = in ? u.sig_in[f / 2] : u.sig_bd[f / 2][f % 2];
            bool        &hv = in ? have_in[f / 2] : have_bd[f / 2][f % 2];
            if (!hv)
              v = sg, hv = true;
            else
              ok = ok && close(sg, v);
          }
      u.ok = ok;
    }
    // the per-(cell, direction) stencil records
    std::vector<double> rec((size_t)h->np_own * dim * 6), vol((size_t)h->np_own);
    for (int32_t c = 0; c < h->np_own; ++c)
      {
        double v = 1.;
        for (int k = 0; k < dim; ++k)
          v *= cell_h[(size_t)c * dim + k];
        vol[c] = v;
        for (int k = 0; k < dim; ++k)
          {
            FineGeo      r;
            const double hk = cell_h[(size_t)c * dim + k], a = v / hk;
            r.cV = a / hk;
            for (int s = 0; s < 2; ++s)
              {
                const int32_t nb = nbr[(size_t)c * nfc + 2 * k + s];
                r.nb[s]          = nb;
                r.P[s]           = a * sigma[(size_t)c * nfc + 2 * k + s];
                r.Q[s]           = nb >= 0 ? 0.5 * a / cell_h[(size_t)nb * dim + k] : 0.;
              }
            std::memcpy(rec.data() + ((size_t)c * dim + k) * 6, &r, sizeof r);
          }
      }
    // 1-D tables: Mh, Sh, e0|e1, d0|d1
    const int n1 = h->n1;
    h->mf_tab_host.assign(2 * n1 * n1 + 4 * n1, 0.);
    std::vector<double> L(n1), dL(n1);
    double             *Mh = h->mf_tab_host.data(), *Sh = Mh + n1 * n1, *e0 = Sh + n1 * n1, *d0 = e0 + 2 * n1;
    for (int q = 0; q < h->nq1; ++q)
      {
        lagrange_host(h->basis, n1, h->quad.x[q], L.data(), dL.data());
        for (int i = 0; i < n1; ++i)
          for (int j = 0; j < n1; ++j)
            {
              Mh[i * n1 + j] += h->quad.w[q] * L[i] * L[j];
              Sh[i * n1 + j] += h->quad.w[q] * dL[i] * dL[j];
            }
      }
    lagrange_host(h->basis, n1, 0., e0, d0);
    lagrange_host(h->basis, n1, 1., e0 + n1, d0 + n1);
    auto put = [](auto &buf, const auto &v) {
      buf.alloc(v.size());
      PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    put(h->mf_geo, rec);
    put(h->mf_vol, vol);
    // Processing order: the cells sorted along the Morton curve through their centres, so that consecutive
    // cells are neighbours in space whatever the caller's numbering (a hyper_cube + refine_global mesh already
    // is in this order); the tiled kernel's tiles then are compact blocks with a small halo.
    std::vector<int32_t>  morton_order((size_t)h->np_own);
    std::vector<uint64_t> key((size_t)h->np_own, 0);
    {
      std::vector<double> ctr((size_t)h->np_own * dim), lo(dim, 1e300), hmin(dim, 1e300);
      for (int32_t p = 0; p < h->np_own; ++p)
        for (int k = 0; k < dim; ++k)
          {
            const double a = d.bbox[(size_t)p * 2 * dim + k], b = d.bbox[(size_t)p * 2 * dim + dim + k];
            ctr[(size_t)d.dof_block[p] * dim + k] = 0.5 * (a + b);
            lo[k]   = std::min(lo[k], a);
            hmin[k] = std::min(hmin[k], b - a);
          }
      // integer cell coordinates q = floor(centre / h) in ABSOLUTE coordinates, shifted by a multiple of 2^20 cells: the
      // curve is then the same whichever part of a mesh a rank holds (the restriction of a Morton-numbered global mesh
      // to a rank is again in the curve's order: no permutation, contiguous tiles), and hyper_cube(-1, 1) splits at 0
      std::vector<int64_t> q0(dim);
      for (int k = 0; k < dim; ++k)
        q0[k] = (int64_t)std::floor(std::floor(lo[k] / hmin[k]) / 1048576.) * 1048576;
      for (int32_t c = 0; c < h->np_own; ++c)
        for (int k = 0; k < dim; ++k)
          {
            const int64_t  qa = (int64_t)std::floor(ctr[(size_t)c * dim + k] / hmin[k]) - q0[k];
            const uint64_t q  = (uint64_t)std::min<int64_t>(2097151, std::max<int64_t>(0, qa));
            for (int b = 0; b < 21; ++b)
              key[c] |= ((q >> b) & 1u) << (b * dim + k);
          }
      for (int32_t c = 0; c < h->np_own; ++c)
        morton_order[c] = c;
      std::stable_sort(morton_order.begin(), morton_order.end(), [&](const int32_t a, const int32_t b) { return key[a] < key[b]; });
    }
    bool identity_order = true;
    for (int32_t c = 0; c < h->np_own; ++c)
      identity_order = identity_order && morton_order[c] == c;
    // sharded handles: cells whose neighbours are all owned can be applied while the ghost blocks travel
    h->mf_list_interior.release();
    h->mf_list_boundary.release();
    h->mf_seq_all.release();
    std::vector<int32_t> inner, outer;
    h->mf_h_nbr = nbr;
    h->mf_h_inner.clear(), h->mf_h_outer.clear();
    h->mf_fused.ok = false;
    if (h->np != h->np_own)
      {
        // the split is made BLOCK by block of the curve (the tiles of the tiled kernel): a block one of whose cells reads
        // ghost data goes to the boundary list as a whole, so that both lists consist of whole tiles -- full
        // contiguous runs of the curve that one bulk copy stages -- instead of blocks with a layer missing
        const int             split_bits = (dim == 3 ? 2 : (dim == 2 ? 3 : 6)) * dim;
        std::vector<uint64_t> bnd_blocks;
        for (const int32_t c : morton_order)
          for (int f = 0; f < nfc; ++f)
            if (nbr[(size_t)c * nfc + f] >= h->np_own)
              {
                if (bnd_blocks.empty() || bnd_blocks.back() != (key[c] >> split_bits))
                  bnd_blocks.push_back(key[c] >> split_bits);
                break;
              }
        std::sort(bnd_blocks.begin(), bnd_blocks.end());
        for (const int32_t c : morton_order)
          (std::binary_search(bnd_blocks.begin(), bnd_blocks.end(), key[c] >> split_bits) ? outer : inner).push_back(c);
        h->mf_h_inner = inner, h->mf_h_outer = outer;
        put(h->mf_list_interior, inner);
        put(h->mf_list_boundary, outer);
      }
    // ---- the tiled kernel: premultiplied tables and one tile plan per cell sequence
    for (auto &t : h->mf_tiles)
      t.ok = false;
    {
      // Which kernel: measured on B200 (profiles/README.md) the tiled kernel wins wherever it was timed -- 3-D
      // DGQ1 / DGQ2 (64^3: 0.049 vs 0.057, 0.070 vs 0.134 ms), 2-D DGQ2 / 3 / 4 (512^2: 0.041 vs 0.048, 0.052 vs
      // 0.100, 0.070 vs 0.200 ms) -- so it is the default there; 2-D DGQ1 (not timed) and 3-D DGQ3 (128
      // accumulators: no tiled instance) run the line kernel.  PD_FINE_KERNEL=tile / line force one of them
      // wherever it exists (the tests run both).
      const char *env    = std::getenv("PD_FINE_KERNEL");
      const bool  faster = (dim == 3 && h->degree <= 2) || (dim == 2 && h->degree >= 2);
      h->mf_kernel = (env && std::strcmp(env, "line") == 0) ? 1 : ((env && std::strcmp(env, "tile") == 0) ? 0 : (faster ? 0 : 1));
      // the pipelined kernel (k_fine_stream) takes over from the tiled one wherever it applies (uniform mesh, N odd,
      // whole aligned tiles); PD_FINE_KERNEL=tile keeps the tiled kernel, =stream asks for the default policy
      h->mf_stream = !(env && std::strcmp(env, "tile") == 0);
    }
    if (h->mf_kernel == 0 && h->n <= 27)
      {
        // X = Mh^-1 B by Gaussian elimination with partial pivoting (Mh is SPD, n1 <= 5)
        auto solve = [n1](const double *M, const double *B, const int ncol, double *X) {
          std::vector<double> a(M, M + n1 * n1), b(B, B + n1 * ncol);
          for (int k = 0; k < n1; ++k)
            {
              int piv = k;
              for (int r = k + 1; r < n1; ++r)
                if (std::fabs(a[r * n1 + k]) > std::fabs(a[piv * n1 + k]))
                  piv = r;
              for (int c = 0; c < n1; ++c)
                std::swap(a[k * n1 + c], a[piv * n1 + c]);
              for (int c = 0; c < ncol; ++c)
                std::swap(b[k * ncol + c], b[piv * ncol + c]);
              for (int r = k + 1; r < n1; ++r)
                {
                  const double f = a[r * n1 + k] / a[k * n1 + k];
                  for (int c = k; c < n1; ++c)
                    a[r * n1 + c] -= f * a[k * n1 + c];
                  for (int c = 0; c < ncol; ++c)
                    b[r * ncol + c] -= f * b[k * ncol + c];
                }
            }
          for (int k = n1 - 1; k >= 0; --k)
            for (int c = 0; c < ncol; ++c)
              {
                double v = b[k * ncol + c];
                for (int r = k + 1; r < n1; ++r)
                  v -= a[k * n1 + r] * X[r * ncol + c];
                X[k * ncol + c] = v / a[k * n1 + k];
              }
        };
        // layout of fine::TileTables: Mh | Shp | ep[2] | dp[2] | d[2]
        h->mf_tile_tab_host.assign(2 * n1 * n1 + 6 * n1, 0.);
        double *tM = h->mf_tile_tab_host.data(), *tS = tM + n1 * n1, *tE = tS + n1 * n1, *tDp = tE + 2 * n1, *tD = tDp + 2 * n1;
        std::copy(Mh, Mh + n1 * n1, tM);
        solve(Mh, Sh, n1, tS);
        std::vector<double> rhs(n1 * 4), sol(n1 * 4); // columns e0, e1, d0, d1
        for (int i = 0; i < n1; ++i)
          {
            rhs[i * 4 + 0] = e0[i], rhs[i * 4 + 1] = e0[n1 + i];
            rhs[i * 4 + 2] = d0[i], rhs[i * 4 + 3] = d0[n1 + i];
          }
        solve(Mh, rhs.data(), 4, sol.data());
        for (int i = 0; i < n1; ++i)
          {
            tE[i] = sol[i * 4 + 0], tE[n1 + i] = sol[i * 4 + 1];
            tDp[i] = sol[i * 4 + 2], tDp[n1 + i] = sol[i * 4 + 3];
            tD[i] = d0[i], tD[n1 + i] = d0[n1 + i];
          }
        // the unit-vector traces cell_apply relies on (FE_DGQ(p >= 1): nodes on both ends)
        bool unit = true;
        for (int i = 0; i < n1; ++i)
          unit = unit && std::fabs(e0[i] - (i == 0 ? 1. : 0.)) < 1e-14 && std::fabs(e0[n1 + i] - (i == n1 - 1 ? 1. : 0.)) < 1e-14;
        h->mf_tiles[3].ok = h->mf_tiles[3].stream_ok = false;
        for (int part = 0; unit && part < 3; ++part)
          {
            const std::vector<int32_t> *seq = part == 0 ? (identity_order ? nullptr : &morton_order) : (part == 1 ? &inner : &outer);
            const int32_t               n_seq = seq ? (int32_t)seq->size() : h->np_own;
            if (n_seq == 0)
              continue;
            // block of a sequence entry: the aligned 4x4x4 (3-D) / 8x8 (2-D) block of the curve, FINE_TILE cells
            const int             block_bits = dim == 3 ? 2 : (dim == 2 ? 3 : 6);
            std::vector<uint64_t> bkey((size_t)n_seq);
            for (int32_t i = 0; i < n_seq; ++i)
              bkey[i] = key[seq ? (*seq)[i] : i] >> (block_bits * dim);
            fine::TilePlan plan;
            try
              {
                plan = fine::build_tile_plan(n_seq, seq ? seq->data() : nullptr, bkey.data(), nbr.data(), nfc, h->np, FINE_TILE, h->n);
              }
            catch (const std::exception &)
              {
                continue;
              }
            if (tile_smem_bytes(dim, h->n, plan.max_halo) > 56 * 1024)
              continue; // tiles with large halos (four CTAs no longer fit an SM): the line-per-thread kernel takes this sequence
            auto &t = h->mf_tiles[part];
            put(t.tile_first, plan.tile_first);
            {
              std::vector<int32_t> base((size_t)plan.n_tiles, -1);
              for (int32_t k = 0; k < plan.n_tiles; ++k)
                {
                  const int32_t s0 = plan.tile_first[k], n_own = plan.tile_first[k + 1] - s0;
                  const int32_t c0 = seq ? (*seq)[s0] : s0;
                  bool          run = true;
                  for (int32_t i = 1; i < n_own && run; ++i)
                    run = (seq ? (*seq)[s0 + i] : s0 + i) == c0 + i;
                  if (run)
                    base[k] = c0;
                }
              put(t.tile_base, base);
            }
            put(t.tile_ptr, plan.tile_ptr);
            put(t.noff, plan.noff);
            if (plan.halo.empty())
              plan.halo.push_back(0);
            put(t.halo, plan.halo);
            t.n_tiles = plan.n_tiles, t.max_halo = plan.max_halo, t.zoff = plan.zoff, t.n_seq = n_seq;
            t.ok = true;
            {
              // the pipelined kernel: tiles whose cells are consecutive numbers (any length, any alignment), own layout
              // of the halo rows (fine::StreamPlan)
              std::vector<int32_t> base_h((size_t)plan.n_tiles);
              bool                 run_ok = h->mf_uniform.ok && h->n % 2 == 1 && (int64_t)h->np * h->n < INT32_MAX; // (32-bit element offsets in the gather)
              for (int32_t k = 0; k < plan.n_tiles && run_ok; ++k)
                {
                  const int32_t s0 = plan.tile_first[k], n_own = plan.tile_first[k + 1] - s0;
                  const int32_t c0 = seq ? (*seq)[s0] : s0;
                  base_h[(size_t)k] = c0;
                  for (int32_t i = 1; i < n_own && run_ok; ++i)
                    run_ok = (seq ? (*seq)[s0 + i] : s0 + i) == c0 + i;
                }
              t.stream_ok = false;
              t.h_tile_first = plan.tile_first, t.h_tile_base = base_h;
              // regular: every tile FINE_TILE cells from a 16-byte boundary of the vector (the kernel's compile-time case)
              t.stream_regular = true;
              for (int32_t k = 0; k < plan.n_tiles; ++k)
                t.stream_regular = t.stream_regular && plan.tile_first[k] == k * FINE_TILE &&
                                   plan.tile_first[k + 1] - plan.tile_first[k] == FINE_TILE && base_h[(size_t)k] % 2 == 0;
              if (run_ok)
                try
                  {
                    const fine::StreamPlan sp = fine::build_stream_plan(n_seq, seq ? seq->data() : nullptr, plan.tile_first.data(),
                                                                        plan.n_tiles, nbr.data(), nfc, h->np, FINE_TILE, h->n);
                    // (the gather: a warp's lanes hold the cells of 32 / rpp passes of rpp rows each, four warps per group)
                    const int rpp = 2 * (32 / (h->n + 1));
                    if (rpp >= 2 && sp.max_rows <= (FINE_TILE_THREADS / 32) * (32 / rpp) * rpp &&
                        stream_smem_bytes(dim, h->n, sp.max_rows, 3) <= 227 * 1024)
                      {
                        std::vector<uint16_t> noff_pad(sp.noff);
                        noff_pad.resize(noff_pad.size() + 8, 0); // (the kernel copies whole 16-byte pieces)
                        put(t.halo_pad, sp.rows);
                        put(t.noff_stream, noff_pad);
                        t.stream_rows = sp.max_rows, t.stream_zoff = sp.zoff;
                        t.stream_ok = true;
                      }
                  }
                catch (const std::exception &)
                  {
                  }
            }
            if (part == 0 && seq)
              put(h->mf_seq_all, morton_order);
          }
      }
    h->mf_rec.alloc((size_t)h->np_own * dim * 8);
    h->mf_zero.alloc((size_t)h->n);
    PD_CUDA(cudaMemset(h->mf_zero.p, 0, (size_t)h->n * sizeof(double)));
    h->mf_rec_valid = false;
    h->mf_ready = true;
  }

  bool
  fine_operator_supported(const int dim, const int degree)
  {
    return (dim == 2 && degree >= 1 && degree <= 4) || (dim == 3 && degree >= 1 && degree <= 3);
  }

  namespace
  {
    // the operator's coefficient and term flags folded into the stencil records (once per pd_set_operator)
    template <int DIM>
    void
    fold_fine_records(pd_handle *h)
    {
      if (!h->mf_rec_valid || h->mf_rec_flags != h->op_flags || h->mf_rec_coef != h->op_coef.stiffness)
        {
          const int64_t n = (int64_t)h->np_own * DIM;
          k_fine_fold<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(reinterpret_cast<const FineGeo *>(h->mf_geo.p),
                                                                        reinterpret_cast<FineRec *>(h->mf_rec.p), n,
                                                                        h->op_coef.stiffness, h->op_flags);
          h->mf_rec_valid = true;
          h->mf_rec_flags = h->op_flags;
          h->mf_rec_coef  = h->op_coef.stiffness;
          ++h->launches;
        }
    }

    template <int DIM, int DEG, int MINB>
    void
    launch_fine(pd_handle *h, const double *src, double *dst, const bool add, const int part)
    {
      constexpr int N1 = DEG + 1, N = ipow_(N1, DIM), NT = DIM * (N / N1);
      constexpr int GS = pow2_at_least(N > NT ? N : NT), CPB = 256 / GS;
      FineArgs<N1>  a;
      static_assert(sizeof(a.T) == (2 * N1 * N1 + 4 * N1) * sizeof(double), "table layout");
      std::memcpy(&a.T, h->mf_tab_host.data(), sizeof(a.T));
      const bool vol_on = (h->op_flags & PD_ASSEMBLE_VOLUME) != 0;
      fold_fine_records<DIM>(h);
      a.rec      = reinterpret_cast<const FineRec *>(h->mf_rec.p);
      a.vol      = h->mf_vol.p;
      a.zero     = h->mf_zero.p;
      a.x        = src;
      a.y        = dst;
      // part 0: all owned cells; 1: cells without ghost neighbours; 2: cells with ghost neighbours
      a.n_cells = part == 0 ? h->np_own : (part == 1 ? (int32_t)h->mf_list_interior.n : (int32_t)h->mf_list_boundary.n);
      a.list    = part == 0 ? nullptr : (part == 1 ? h->mf_list_interior.p : h->mf_list_boundary.p);
      a.mass    = vol_on ? h->op_coef.mass : 0.;
      a.add     = add ? 1 : 0;
      if (a.n_cells == 0)
        return;
      const int64_t want = ((int64_t)a.n_cells + CPB - 1) / CPB;
      const int     grid = (int)std::min<int64_t>(want, (int64_t)h->sm_count * 4 * MINB);
      k_fine_sip<DIM, DEG, MINB><<<grid, 256, 0, h->stream>>>(a);
    }
    template <int DIM, int DEG>
    void
    launch_fine_tiled(pd_handle *h, const double *src, double *dst, const bool add, const int part)
    {
      constexpr int     N1 = DEG + 1;
      const auto       &t  = h->mf_tiles[part];
      TileArgs<N1>      a;
      static_assert(sizeof(a.T) == (2 * N1 * N1 + 6 * N1) * sizeof(double), "table layout");
      std::memcpy(&a.T, h->mf_tile_tab_host.data(), sizeof(a.T));
      fold_fine_records<DIM>(h);
      const bool vol_on = (h->op_flags & PD_ASSEMBLE_VOLUME) != 0;
      a.rec      = reinterpret_cast<const FineRec *>(h->mf_rec.p);
      a.vol      = h->mf_vol.p;
      a.x        = src;
      a.y        = dst;
      a.seq      = part == 0 ? h->mf_seq_all.p /* nullptr: the cells are numbered along the curve already */ :
                               (part == 1 ? h->mf_list_interior.p : h->mf_list_boundary.p);
      a.tile_first = t.tile_first.p;
      a.tile_base  = t.tile_base.p;
      a.tile_ptr = t.tile_ptr.p;
      a.halo     = t.halo.p;
      a.noff     = t.noff.p;
      a.n_seq    = t.n_seq;
      a.max_halo = t.max_halo;
      a.zoff     = t.zoff;
      a.mass     = vol_on ? h->op_coef.mass : 0.;
      a.add      = add ? 1 : 0;
      const size_t smem = tile_smem_bytes(DIM, ipow_(N1, DIM), t.max_halo);
      static size_t smem_set = 0; // per instantiation
      if (smem > smem_set)
        {
          PD_CUDA(cudaFuncSetAttribute(k_fine_tile<DIM, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          smem_set = smem;
        }
      k_fine_tile<DIM, DEG><<<t.n_tiles, FINE_TILE_THREADS, smem, h->stream>>>(a);
    }
    // the uniform mesh's dense line stencils with the operator's coefficient and term flags folded in
    // (what k_fine_fold writes into the records of the other kernels)
    template <int DIM, int N1>
    void
    uniform_dense_tables(const pd_handle *h, fine::DenseTables<DIM, N1> &D)
    {
      const auto          &u = h->mf_uniform;
      fine::TileTables<N1> T;
      static_assert(sizeof(T) == (2 * N1 * N1 + 6 * N1) * sizeof(double), "table layout");
      std::memcpy(&T, h->mf_tile_tab_host.data(), sizeof(T));
      fine::UniformLine U[DIM];
      const double      c = h->op_coef.stiffness;
      const bool        vol_on = (h->op_flags & PD_ASSEMBLE_VOLUME) != 0, in_on = (h->op_flags & PD_ASSEMBLE_INTERIOR) != 0,
                 bd_on = (h->op_flags & PD_ASSEMBLE_BOUNDARY) != 0;
      double vol = 1.;
      for (int k = 0; k < DIM; ++k)
        vol *= u.h[k];
      for (int k = 0; k < DIM; ++k)
        {
          const double a = vol / u.h[k], cV = a / u.h[k];
          U[k].cVol = vol_on ? c * cV : 0.;
          U[k].cDi  = in_on ? c * (0.5 * cV) : 0.;
          U[k].Pi   = in_on ? c * (a * u.sig_in[k]) : 0.;
          U[k].Qi   = in_on ? c * (0.5 * a / u.h[k]) : 0.;
          U[k].cDb  = bd_on ? c * cV : 0.;
          for (int s = 0; s < 2; ++s)
            U[k].Pb[s] = bd_on ? c * (a * u.sig_bd[k][s]) : 0.;
        }
      fine::build_dense_tables<DIM, N1>(T, U, vol_on ? h->op_coef.mass * vol : 0., D);
    }

    template <int DIM, int DEG>
    void
    launch_fine_stream(pd_handle *h, const double *src, double *dst, const bool add, const int part)
    {
      constexpr int       N1 = DEG + 1, N = ipow_(N1, DIM);
      const auto         &t  = h->mf_tiles[part];
      StreamArgs<DIM, N1> a;
      uniform_dense_tables<DIM, N1>(h, a.T);
      a.ghost_src = nullptr, a.parity_stride = 0, a.epochs = nullptr, a.flags = nullptr, a.error_word = nullptr, a.owners = nullptr;
      a.n_owners = 0, a.np_own = h->np_own, a.first_ghost_tile = 0;
      {
        static const char *e = std::getenv("PD_FINE_REFILL");
        a.refill_late        = e ? std::atoi(e) : 0; // (measured at 128^3: 0.306 ms with the refill at the top, 0.330 after the line phase)
      }
      if (part == 3)
        {
          const auto &f      = h->mf_fused;
          a.ghost_src        = f.ghost_src.p;
          a.parity_stride    = f.parity_stride;
          a.epochs           = f.epochs;
          a.flags            = f.flags;
          a.error_word       = f.error_word;
          a.owners           = f.owners;
          a.n_owners         = f.n_owners;
          a.first_ghost_tile = f.first_ghost_tile;
        }
      a.x          = src;
      a.y          = dst;
      a.tile_first = t.tile_first.p;
      a.tile_base  = t.tile_base.p;
      a.halo_pad   = t.halo_pad.p;
      a.noff      = t.noff_stream.p;
      a.x_len     = (int64_t)h->np * N;
      a.n_tiles   = t.n_tiles;
      a.max_halo  = t.stream_rows;
      a.zoff      = t.stream_zoff;
      a.add       = add ? 1 : 0;
      // groups per CTA: PD_FINE_GROUPS (2..5; two groups: two CTAs per SM), stages: PD_FINE_STAGES; default 3 groups
      // and 6 stages.  Measured, 64^3 / 128^3 cells of DGQ2: 3 x 6 0.0467 / 0.278 ms, 2 x 3 0.0469 / 0.287,
      // 4 x 6 0.0478 / 0.284, 5 x 6 (96 registers, spills) 0.058 / 0.338
      static int groups_env = -1, stages_env = -1;
      if (groups_env < 0)
        {
          const char *e = std::getenv("PD_FINE_GROUPS"), *f = std::getenv("PD_FINE_STAGES");
          groups_env    = e ? std::min(5, std::max(2, std::atoi(e))) : 0;
          stages_env    = f ? std::min(FINE_MAX_STAGES, std::max(2, std::atoi(f))) : 0;
        }
      const size_t cap = 227 * 1024;
      int          groups = groups_env ? groups_env : 3;
      while (groups > 2 && stream_smem_bytes(DIM, N, t.stream_rows, groups + 1) > cap)
        --groups;
      int stages = stages_env ? stages_env : (groups <= 2 ? 3 : 6);
      const size_t per_cta_cap = groups <= 2 ? cap / 2 - 1024 : cap; // (two groups: two CTAs per SM)
      while (stages > 2 && stream_smem_bytes(DIM, N, t.stream_rows, stages) > per_cta_cap)
        --stages;
      if (stages <= groups)
        groups = std::max(2, stages - 1); // (more stages than groups: what keeps the refills free of cycles)
      if (stages <= groups)
        stages = groups + 1;
      a.n_stages        = stages;
      const size_t smem = stream_smem_bytes(DIM, N, t.stream_rows, stages);
      auto go = [&](auto kernel, size_t &smem_set, const int threads) {
        if (smem > smem_set)
          {
            PD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set = smem;
          }
        int per_sm = 1;
        PD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
        const int grid = std::max(1, std::min(t.n_tiles, h->sm_count * std::max(1, per_sm)));
        kernel<<<grid, threads, smem, h->stream>>>(a);
      };
      static size_t smem_set[6] = {0, 0, 0, 0, 0, 0}; // per instantiation
      static const bool force_general = std::getenv("PD_FINE_GENERAL") != nullptr; // (tests: the general kernel on regular tiles)
      const bool        general       = part == 3 || !t.stream_regular || force_general;
      static size_t     smem_set_g[6] = {0, 0, 0, 0, 0, 0};
      if (general)
        switch (groups)
          {
            case 2: go(k_fine_stream<DIM, DEG, 2, true>, smem_set_g[2], 2 * FINE_TILE_THREADS); break;
            case 3: go(k_fine_stream<DIM, DEG, 3, true>, smem_set_g[3], 3 * FINE_TILE_THREADS); break;
            case 4: go(k_fine_stream<DIM, DEG, 4, true>, smem_set_g[4], 4 * FINE_TILE_THREADS); break;
            default: go(k_fine_stream<DIM, DEG, 5, true>, smem_set_g[5], 5 * FINE_TILE_THREADS); break;
          }
      else
        switch (groups)
          {
            case 2: go(k_fine_stream<DIM, DEG, 2, false>, smem_set[2], 2 * FINE_TILE_THREADS); break;
            case 3: go(k_fine_stream<DIM, DEG, 3, false>, smem_set[3], 3 * FINE_TILE_THREADS); break;
            case 4: go(k_fine_stream<DIM, DEG, 4, false>, smem_set[4], 4 * FINE_TILE_THREADS); break;
            default: go(k_fine_stream<DIM, DEG, 5, false>, smem_set[5], 5 * FINE_TILE_THREADS); break;
          }
    }
  } // namespace

  void
  launch_fine_operator(pd_handle *h, const double *src, double *dst, const bool add, const int part)
  {
    // the pipelined kernel: uniform mesh, whole aligned tiles, both vectors 16-byte aligned (bulk copies in and out)
    if (h->mf_kernel == 0 && h->mf_stream && h->mf_uniform.ok && h->mf_tiles[part].ok && h->mf_tiles[part].stream_ok &&
        reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(dst) % 16 == 0)
      {
        bool done = true;
        switch (h->dim * 10 + h->degree)
          {
            case 22: launch_fine_stream<2, 2>(h, src, dst, add, part); break;
            case 24: launch_fine_stream<2, 4>(h, src, dst, add, part); break;
            case 32: launch_fine_stream<3, 2>(h, src, dst, add, part); break;
            default: done = false;
          }
        if (done)
          {
            h->mf_kernel_last = PD_FINE_KERNEL_STREAM;
            ++h->launches;
            PD_CUDA(cudaGetLastError());
            return;
          }
      }
    // (the tiled kernel fills its halo rows with 16-byte bulk copies: a source vector that is only 8-byte
    // aligned goes to the line-per-thread kernel)
    const bool tiled = h->mf_kernel == 0 && h->mf_tiles[part].ok && reinterpret_cast<uintptr_t>(src) % 16 == 0;
    switch ((tiled ? 100 : 0) + h->dim * 10 + h->degree)
      {
        case 121: launch_fine_tiled<2, 1>(h, src, dst, add, part); break;
        case 122: launch_fine_tiled<2, 2>(h, src, dst, add, part); break;
        case 123: launch_fine_tiled<2, 3>(h, src, dst, add, part); break;
        case 124: launch_fine_tiled<2, 4>(h, src, dst, add, part); break;
        case 131: launch_fine_tiled<3, 1>(h, src, dst, add, part); break;
        case 132: launch_fine_tiled<3, 2>(h, src, dst, add, part); break;
        case 21: launch_fine<2, 1, 4>(h, src, dst, add, part); break;
        case 22: launch_fine<2, 2, 3>(h, src, dst, add, part); break;
        case 23: launch_fine<2, 3, 3>(h, src, dst, add, part); break;
        case 24: launch_fine<2, 4, 2>(h, src, dst, add, part); break;
        case 31: launch_fine<3, 1, 4>(h, src, dst, add, part); break;
        case 32: launch_fine<3, 2, 3>(h, src, dst, add, part); break;
        case 33: launch_fine<3, 3, 2>(h, src, dst, add, part); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no fine-mesh operator kernel for this (dim, degree)", __LINE__};
      }
    h->mf_kernel_last = tiled ? PD_FINE_KERNEL_TILE : PD_FINE_KERNEL_LINE;
    ++h->launches;
    PD_CUDA(cudaGetLastError());
  }

  // The fused sharded apply: ONE launch of the pipelined kernel over the interior tiles followed by the boundary
  // tiles.  The boundary tiles come last in every CTA's list; before a warp gathers the halo rows of the first of
  // them it waits for the owners' epoch flags, and it reads the ghost cells' coefficients straight from the owners'
  // export buffers over NVLink -- no pull kernel, no ghost section round trip, no second stream, no separate launch
  // for the boundary cells (reference: update_ghost_values() inside MatrixFree::loop, include/utils.h:466-472).
  bool
  setup_fine_fused(pd_handle *h, const double *const *ghost_src_host, const int64_t parity_stride, const unsigned long long *epochs,
                   unsigned long long *flags, unsigned long long *error_word, const int32_t *owners_dev, const int n_owners)
  {
    h->mf_fused.ok = false;
    h->mf_tiles[3].ok = h->mf_tiles[3].stream_ok = false;
    const int32_t n_ghost = h->np - h->np_own;
    // (the interior / boundary lists themselves need not be streamable: their tiles are split here where the gather's
    // row budget asks for it -- a METIS cut of 64^3 cells per rank leaves a few boundary tiles with 129 or 130 rows)
    if (!h->mf_ready || !h->mf_stream || h->mf_kernel != 0 || !h->mf_uniform.ok || n_ghost <= 0 || !h->mf_tiles[1].ok ||
        !h->mf_tiles[2].ok || h->n % 2 == 0 || n_owners > 32 || (int64_t)h->np * h->n >= INT32_MAX ||
        !(h->dim * 10 + h->degree == 22 || h->dim * 10 + h->degree == 24 || h->dim * 10 + h->degree == 32))
      return false;
    const int            nfc = 2 * h->dim;
    std::vector<uint8_t> par((size_t)h->np);
    for (int32_t c = 0; c < h->np_own; ++c)
      par[(size_t)c] = (uint8_t)(((int64_t)c * h->n) & 1);
    for (int32_t g = 0; g < n_ghost; ++g)
      par[(size_t)h->np_own + g] = (uint8_t)((reinterpret_cast<uintptr_t>(ghost_src_host[g]) / sizeof(double)) & 1);
    // (the gather: a warp's lanes hold the cells of 32 / rpp passes of rpp rows each, four warps per group)
    const int     rpp      = 2 * (32 / (h->n + 1));
    int32_t       max_rows = rpp >= 2 ? (FINE_TILE_THREADS / 32) * (32 / rpp) * rpp : 0;
    if (const char *e = std::getenv("PD_FINE_FUSED_MAX_ROWS")) // (tests: a smaller budget, so that small meshes split their tiles too)
      max_rows = std::min(max_rows, std::max(2 * nfc, std::atoi(e)));
    fine::FusedPlan fp;
    try
      {
        fp = fine::build_fused_plan(h->mf_h_inner, h->mf_h_outer, h->mf_tiles[1].h_tile_first, h->mf_tiles[2].h_tile_first,
                                    h->mf_h_nbr.data(), nfc, h->np, FINE_TILE, h->n, par.data(), max_rows);
      }
    catch (const std::exception &)
      {
        return false;
      }
    const fine::StreamPlan     &sp   = fp.sp;
    const std::vector<int32_t> &seq = fp.seq, &tf = fp.tile_first, &base = fp.tile_base;
    if (sp.max_rows > max_rows || stream_smem_bytes(h->dim, h->n, sp.max_rows, 3) > 227 * 1024)
      return false;
    auto put = [](auto &buf, const auto &v) {
      buf.alloc(v.size());
      PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    auto &t = h->mf_tiles[3];
    {
      std::vector<uint16_t> noff_pad(sp.noff);
      noff_pad.resize(noff_pad.size() + 8, 0);
      put(t.noff_stream, noff_pad);
    }
    put(t.tile_first, tf);
    put(t.tile_base, base);
    put(t.halo_pad, sp.rows);
    t.stream_rows = sp.max_rows, t.stream_zoff = sp.zoff, t.n_tiles = sp.n_tiles, t.n_seq = (int32_t)seq.size();
    t.ok = false, t.stream_ok = true; // (the pipelined kernel only)
    auto &f = h->mf_fused;
    {
      std::vector<const double *> gs(ghost_src_host, ghost_src_host + n_ghost);
      put(f.ghost_src, gs);
    }
    f.parity_stride    = parity_stride;
    f.epochs           = epochs;
    f.flags            = flags;
    f.error_word       = error_word;
    f.owners           = owners_dev;
    f.n_owners         = n_owners;
    f.first_ghost_tile = fp.first_ghost_tile;
    f.ok               = true;
    return true;
  }

  bool
  launch_fine_fused(pd_handle *h, const double *src, double *dst, const bool add)
  {
    if (!h->mf_fused.ok || !h->mf_tiles[3].stream_ok || reinterpret_cast<uintptr_t>(src) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(dst) % 16 != 0)
      return false;
    switch (h->dim * 10 + h->degree)
      {
        case 22: launch_fine_stream<2, 2>(h, src, dst, add, 3); break;
        case 24: launch_fine_stream<2, 4>(h, src, dst, add, 3); break;
        case 32: launch_fine_stream<3, 2>(h, src, dst, add, 3); break;
        default: return false;
      }
    h->mf_kernel_last = PD_FINE_KERNEL_STREAM;
    ++h->launches;
    PD_CUDA(cudaGetLastError());
    return true;
  }
} // namespace pd
