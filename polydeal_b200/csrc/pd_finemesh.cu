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
//   Mh = int l_i l_j,  Sh = int l_i' l_j'   (the cell Gauss rule)
//   Mf = int l_i l_j                        (the face Gauss rule)
//   e0/e1 = l_i(0)/l_i(1),  d0/d1 = l_i'(0)/l_i'(1)
// applied by 1-D contractions through shared memory.
//
// HBM-bound by design: 16 B per DoF (read src once, write dst once; neighbour
// reads are L1/L2 hits), O(1) geometry per cell.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

#include <algorithm>
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

    struct FineArgs
    {
      const double  *tables; // Mh[N1*N1], Sh[N1*N1], Mf[N1*N1], e0[N1], e1[N1], d0[N1], d1[N1]
      const double  *cell_h; // [n_cells][DIM]   extents, block order
      const int32_t *nbr;    // [n_cells][2*DIM] neighbour block, -1 = boundary
      const double  *sigma;  // [n_cells][2*DIM] penalty of the face
      const double  *x;
      double        *y;
      int32_t        n_cells;
      double         stiffness, mass;
      uint32_t       flags;
      int            add;
    };

    // One thread per DoF, GS threads per cell (GS = N rounded up to a power of two),
    // CPB cells per block, persistent over cell batches.
    template <int DIM, int DEG>
    __global__ void __launch_bounds__(256)
    k_fine_sip(const FineArgs A)
    {
      constexpr int N1  = DEG + 1;
      constexpr int N   = ipow_(N1, DIM);
      constexpr int NF  = ipow_(N1, DIM - 1); // DoFs of a face trace
      constexpr int GS  = N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 128))));
      constexpr int CPB = 256 / GS;
      constexpr int NFC = 2 * DIM; // faces per cell

      __shared__ double tab[3 * N1 * N1 + 4 * N1];
      __shared__ double sU[CPB][N];        // own coefficients
      __shared__ double sN[CPB][NFC][N];   // neighbours' coefficients
      __shared__ double sA[CPB][NFC][NF];  // coefficient of v(face)
      __shared__ double sB[CPB][NFC][NF];  // coefficient of dn v(face)
      __shared__ double sW[CPB][4][N];     // contraction work arrays
      __shared__ double sT[CPB][NFC][2][NF];

      const double *Mh = tab, *Sh = tab + N1 * N1, *Mf = tab + 2 * N1 * N1;
      const double *e0 = tab + 3 * N1 * N1, *d0 = e0 + 2 * N1; // e0,e1 contiguous; d0,d1 contiguous

      for (int i = threadIdx.x; i < 3 * N1 * N1 + 4 * N1; i += blockDim.x)
        tab[i] = A.tables[i];
      __syncthreads();
      // all shared arrays are private to a cell slot: when a slot is (part of) one warp the
      // phases only need warp-level ordering
      auto group_sync = [] {
        if constexpr (GS <= 32)
          __syncwarp();
        else
          __syncthreads();
      };

      const int  slot = threadIdx.x / GS, l = threadIdx.x % GS;
      const bool lane_ok = l < N;
      int        idx[DIM];
      {
        int r = l;
#pragma unroll
        for (int d = 0; d < DIM; ++d)
          {
            idx[d] = r % N1;
            r /= N1;
          }
      }
      constexpr int stride[3] = {1, N1, N1 * N1};

      for (int c0 = blockIdx.x * CPB; c0 < A.n_cells; c0 += gridDim.x * CPB)
        {
          const int  cell = c0 + slot;
          const bool ok   = lane_ok && cell < A.n_cells;
          group_sync(); // previous batch done with this slot's shared arrays
          double h[DIM];
          int    nb[NFC];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            h[d] = 1.;
#pragma unroll
          for (int f = 0; f < NFC; ++f)
            nb[f] = -1;
          if (cell < A.n_cells)
            {
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                h[d] = A.cell_h[(int64_t)cell * DIM + d];
#pragma unroll
              for (int f = 0; f < NFC; ++f)
                nb[f] = A.nbr[(int64_t)cell * NFC + f];
            }
          if (ok)
            {
              sU[slot][l] = A.x[(int64_t)cell * N + l];
#pragma unroll
              for (int f = 0; f < NFC; ++f)
                sN[slot][f][l] = nb[f] >= 0 ? A.x[(int64_t)nb[f] * N + l] : 0.;
            }
          group_sync();
          double vol = 1.;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            vol *= h[d];

          // ---- face traces: own value/normal derivative and the neighbour's, on every face
          if (cell < A.n_cells)
            for (int m = l; m < NFC * NF; m += GS)
              {
                const int f = m / NF, e = m - f * NF;
                const int d = f >> 1, s = f & 1;
                // base index of face entry e: dims other than d, in increasing order
                int base = 0, r = e;
#pragma unroll
                for (int dd = 0; dd < DIM; ++dd)
                  if (dd != d)
                    {
                      base += (r % N1) * stride[dd];
                      r /= N1;
                    }
                const double *ev_own = e0 + s * N1, *dv_own = d0 + s * N1;             // own side s
                const double *ev_nb = e0 + (1 - s) * N1, *dv_nb = d0 + (1 - s) * N1;   // neighbour: opposite side
                double        u = 0., du = 0., p = 0., dp = 0.;
#pragma unroll
                for (int t = 0; t < N1; ++t)
                  {
                    const double a = sU[slot][base + t * stride[d]], b = sN[slot][f][base + t * stride[d]];
                    u += ev_own[t] * a;
                    du += dv_own[t] * a;
                    p += ev_nb[t] * b;
                    dp += dv_nb[t] * b;
                  }
                const double sn  = s ? 1. : -1.;
                const double ih  = 1. / h[d];
                const double sg  = A.sigma[(int64_t)cell * NFC + f];
                double       av, bv;
                if (nb[f] >= 0)
                  {
                    // neighbour extent along d equals... not necessarily ours: carried in sigma only;
                    // its normal derivative needs 1/h of the NEIGHBOUR
                    const double ihn = 1. / A.cell_h[(int64_t)nb[f] * DIM + d];
                    const double jmp = u - p;
                    av               = (A.flags & PD_ASSEMBLE_INTERIOR) ? sg * jmp - 0.5 * sn * (du * ih + dp * ihn) : 0.;
                    bv               = (A.flags & PD_ASSEMBLE_INTERIOR) ? -0.5 * jmp : 0.;
                  }
                else
                  {
                    av = (A.flags & PD_ASSEMBLE_BOUNDARY) ? sg * u - sn * du * ih : 0.;
                    bv = (A.flags & PD_ASSEMBLE_BOUNDARY) ? -u : 0.;
                  }
                sT[slot][f][0][e] = av;
                sT[slot][f][1][e] = bv;
              }
          group_sync();
          // ---- surface mass (Mf x Mf) on the face arrays, times the face area
          if constexpr (DIM == 3)
            {
              if (cell < A.n_cells)
                for (int m = l; m < NFC * 2 * NF; m += GS)
                  {
                    const int f = m / (2 * NF), w = (m / NF) & 1, e = m % NF;
                    const int a = e % N1, b = e / N1;
                    double    s = 0.;
#pragma unroll
                    for (int t = 0; t < N1; ++t)
                      s += Mf[a * N1 + t] * sT[slot][f][w][t + b * N1];
                    (w ? sB : sA)[slot][f][e] = s;
                  }
              group_sync();
              if (cell < A.n_cells)
                for (int m = l; m < NFC * 2 * NF; m += GS)
                  {
                    const int f = m / (2 * NF), w = (m / NF) & 1, e = m % NF;
                    const int a = e % N1, b = e / N1, d = f >> 1;
                    double    s = 0.;
#pragma unroll
                    for (int t = 0; t < N1; ++t)
                      s += Mf[b * N1 + t] * (w ? sB : sA)[slot][f][a + t * N1];
                    sT[slot][f][w][e] = s * (vol / h[d]);
                  }
              group_sync();
            }
          else
            {
              if (cell < A.n_cells)
                for (int m = l; m < NFC * 2 * NF; m += GS)
                  {
                    const int f = m / (2 * NF), w = (m / NF) & 1, e = m % NF, d = f >> 1;
                    double    s = 0.;
#pragma unroll
                    for (int t = 0; t < N1; ++t)
                      s += Mf[e * N1 + t] * sT[slot][f][w][t];
                    (w ? sB : sA)[slot][f][e] = s * (vol / h[d]);
                  }
              group_sync();
              if (cell < A.n_cells)
                for (int m = l; m < NFC * 2 * NF; m += GS)
                  {
                    const int f = m / (2 * NF), w = (m / NF) & 1, e = m % NF;
                    sT[slot][f][w][e] = (w ? sB : sA)[slot][f][e];
                  }
              group_sync();
            }

          // ---- cell term by 1-D contractions.  3-D:
          //   c0 Sx(My Mz U) + Mx[ c1 Sy(Mz U) + c2 My(Sz U) + f vol My Mz U ],  c_d = vol / h_d^2
          double acc = 0.;
          auto contract = [&](const double *Mat, const double *src, const int d) {
            double s = 0.;
#pragma unroll
            for (int t = 0; t < N1; ++t)
              s += Mat[idx[d] * N1 + t] * src[l + (t - idx[d]) * stride[d]];
            return s;
          };
          const bool vol_on = (A.flags & PD_ASSEMBLE_VOLUME) != 0;
          if constexpr (DIM == 3)
            {
              if (ok)
                {
                  sW[slot][0][l] = contract(Mh, sU[slot], 2); // Mz U
                  sW[slot][1][l] = contract(Sh, sU[slot], 2); // Sz U
                }
              group_sync();
              double yz = 0.;
              if (ok)
                {
                  yz = contract(Mh, sW[slot][0], 1);                                         // My Mz U
                  const double syz = contract(Sh, sW[slot][0], 1), mysz = contract(Mh, sW[slot][1], 1);
                  sW[slot][2][l]   = yz;
                  sW[slot][3][l]   = A.stiffness * ((vol / (h[1] * h[1])) * syz + (vol / (h[2] * h[2])) * mysz) +
                                   A.mass * vol * yz;
                }
              group_sync();
              if (ok && vol_on)
                acc = A.stiffness * (vol / (h[0] * h[0])) * contract(Sh, sW[slot][2], 0) + contract(Mh, sW[slot][3], 0);
            }
          else
            {
              if (ok)
                {
                  sW[slot][0][l] = contract(Mh, sU[slot], 1); // My U
                  sW[slot][1][l] = A.stiffness * (vol / (h[1] * h[1])) * contract(Sh, sU[slot], 1); // Sy U
                }
              group_sync();
              if (ok && vol_on)
                {
                  const double sx = contract(Sh, sW[slot][0], 0), mx_my = contract(Mh, sW[slot][0], 0);
                  acc = A.stiffness * (vol / (h[0] * h[0])) * sx + contract(Mh, sW[slot][1], 0) + A.mass * vol * mx_my;
                }
            }
          // ---- lift the face terms: v(face) = e_s[i_d], dn v(face) = sn d_s[i_d] / h_d
          if (ok)
            {
              double fa = 0.;
#pragma unroll
              for (int f = 0; f < NFC; ++f)
                {
                  const int d = f >> 1, s = f & 1;
                  int       e = 0, mul = 1;
#pragma unroll
                  for (int dd = 0; dd < DIM; ++dd)
                    if (dd != d)
                      {
                        e += idx[dd] * mul;
                        mul *= N1;
                      }
                  const double sn = s ? 1. : -1.;
                  fa += (e0 + s * N1)[idx[d]] * sT[slot][f][0][e] + sn * (d0 + s * N1)[idx[d]] / h[d] * sT[slot][f][1][e];
                }
              acc += A.stiffness * fa;
              double *yp = A.y + (int64_t)cell * N + l;
              *yp        = A.add ? *yp + acc : acc;
            }
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
    if (h->n_subcells != h->np_own)
      return;
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
    // 1-D tables
    const int           n1 = h->n1;
    std::vector<double> tab(3 * n1 * n1 + 4 * n1, 0.), L(n1), dL(n1);
    double             *Mh = tab.data(), *Sh = Mh + n1 * n1, *Mf = Sh + n1 * n1, *e0 = Mf + n1 * n1, *d0 = e0 + 2 * n1;
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
    for (int q = 0; q < h->nq1f; ++q)
      {
        lagrange_host(h->basis, n1, h->quadf.x[q], L.data(), dL.data());
        for (int i = 0; i < n1; ++i)
          for (int j = 0; j < n1; ++j)
            Mf[i * n1 + j] += h->quadf.w[q] * L[i] * L[j];
      }
    lagrange_host(h->basis, n1, 0., e0, d0);
    lagrange_host(h->basis, n1, 1., e0 + n1, d0 + n1);
    auto put = [](auto &buf, const auto &v) {
      buf.alloc(v.size());
      PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    put(h->mf_tables, tab);
    put(h->mf_cell_h, cell_h);
    put(h->mf_nbr, nbr);
    put(h->mf_sigma, sigma);
    h->mf_ready = true;
  }

  bool
  fine_operator_supported(const int dim, const int degree)
  {
    return (dim == 2 && degree >= 1 && degree <= 4) || (dim == 3 && degree >= 1 && degree <= 3);
  }

  void
  launch_fine_operator(pd_handle *h, const double *src, double *dst, const bool add)
  {
    FineArgs a;
    a.tables    = h->mf_tables.p;
    a.cell_h    = h->mf_cell_h.p;
    a.nbr       = h->mf_nbr.p;
    a.sigma     = h->mf_sigma.p;
    a.x         = src;
    a.y         = dst;
    a.n_cells   = h->np_own;
    a.stiffness = h->op_coef.stiffness;
    a.mass      = h->op_coef.mass;
    a.flags     = h->op_flags;
    a.add       = add ? 1 : 0;
    const int key = h->dim * 10 + h->degree;
    auto      go  = [&](auto kern, const int gs) {
      const int     cpb  = 256 / gs;
      const int64_t want = ((int64_t)h->np_own + cpb - 1) / cpb;
      const int     grid = (int)std::min<int64_t>(want, (int64_t)h->sm_count * 16);
      kern<<<grid, 256, 0, h->stream>>>(a);
    };
    switch (key)
      {
        case 21: go(k_fine_sip<2, 1>, 4); break;
        case 22: go(k_fine_sip<2, 2>, 16); break;
        case 23: go(k_fine_sip<2, 3>, 16); break;
        case 24: go(k_fine_sip<2, 4>, 32); break;
        case 31: go(k_fine_sip<3, 1>, 8); break;
        case 32: go(k_fine_sip<3, 2>, 32); break;
        case 33: go(k_fine_sip<3, 3>, 64); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no fine-mesh operator kernel for this (dim, degree)", __LINE__};
      }
    ++h->launches;
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
