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
    //
    // On a Cartesian cell the SIP operator is
    //   sum_d  M (x) ... (x) L_d (x) ... (x) M   +  f vol  M (x) M (x) M
    // where L_d is the 1-D DG stencil along d: three (N1 x N1) matrices acting on the
    // cell's own line of DoFs and on the lines of its two neighbours along d,
    //   B_own   = (a/h) Sh + sum_s a [ sig e_s e_s^T - sn/(2h) (e_s d_s^T + d_s e_s^T) ]
    //   B_nbr,s = a [ -sig e_s e_s'^T - sn/(2 h_N) e_s d_s'^T + sn/(2h) d_s e_s'^T ],  s' = 1-s
    // (boundary face: a [ pen e_s e_s^T - sn/h (e_s d_s^T + d_s e_s^T) ], no neighbour part),
    // a = vol/h the face area, sn = -1/+1 the outward normal sign, and M the 1-D mass
    // (cell and face Gauss rules coincide).  15 one-dimensional contractions per cell in 3-D.
    template <int DIM, int DEG>
    __global__ void __launch_bounds__(256, 4) // latency-bound: occupancy beats spill-free code here (measured)
    k_fine_sip(const FineArgs A)
    {
      constexpr int N1  = DEG + 1;
      constexpr int N   = ipow_(N1, DIM);
      constexpr int GS  = N <= 4 ? 4 : (N <= 8 ? 8 : (N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 128))));
      constexpr int CPB = 256 / GS;
      constexpr int NFC = 2 * DIM; // faces per cell
      constexpr int NM  = N1 * N1;

      __shared__ double tab[3 * NM + 4 * N1];
      __shared__ double sU[CPB][N];           // own coefficients
      __shared__ double sN[CPB][NFC][N];      // neighbours' coefficients
      __shared__ double sW[CPB][4][N];        // contraction work arrays

      const double *Mh = tab, *Sh = tab + NM;
      const double *e0 = tab + 3 * NM, *d0 = e0 + 2 * N1; // e0,e1 contiguous; d0,d1 contiguous

      for (int i = threadIdx.x; i < 3 * NM + 4 * N1; i += blockDim.x)
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
      const bool    vol_on = (A.flags & PD_ASSEMBLE_VOLUME) != 0;

      for (int c0 = blockIdx.x * CPB; c0 < A.n_cells; c0 += gridDim.x * CPB)
        {
          const int  cell  = c0 + slot;
          const bool cell_ok = cell < A.n_cells;
          const bool ok    = lane_ok && cell_ok;
          group_sync(); // previous batch done with this slot's shared arrays
          if (ok)
            {
              sU[slot][l] = A.x[(int64_t)cell * N + l];
#pragma unroll
              for (int f = 0; f < NFC; ++f)
                {
                  const int nb   = A.nbr[(int64_t)cell * NFC + f];
                  sN[slot][f][l] = nb >= 0 ? A.x[(int64_t)nb * N + l] : 0.;
                }
            }
          double vol = 1., h[DIM];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              h[d] = cell_ok ? A.cell_h[(int64_t)cell * DIM + d] : 1.;
              vol *= h[d];
            }
          group_sync();
          auto contract = [&](const double *Mat, const double *src, const int d) {
            double s = 0.;
#pragma unroll
            for (int t = 0; t < N1; ++t)
              s += Mat[idx[d] * N1 + t] * src[l + (t - idx[d]) * stride[d]];
            return s;
          };
          // ---- 1-D stencil along every direction, per lane.  The face parts of the stencil
          //      matrices are rank one / two (outer products of e_s, d_s), so each lane forms
          //      the traces of its own line (and of the neighbours' lines) and lifts them.
          double st[DIM];
          if (ok)
            {
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  const int     i    = idx[d];
                  const double *u    = sU[slot] + l - i * stride[d];
                  const double  ih   = 1. / h[d], a = vol * ih;
                  double        su   = 0.;
                  double        tu[2] = {0., 0.}, du[2] = {0., 0.};
#pragma unroll
                  for (int t = 0; t < N1; ++t)
                    {
                      const double ut = u[t * stride[d]];
                      su += Sh[i * N1 + t] * ut;
                      tu[0] += e0[t] * ut;
                      tu[1] += e0[N1 + t] * ut;
                      du[0] += d0[t] * ut;
                      du[1] += d0[N1 + t] * ut;
                    }
                  double acc_d = vol_on ? a * ih * su : 0.;
#pragma unroll
                  for (int s = 0; s < 2; ++s)
                    {
                      const int     f  = 2 * d + s;
                      const int     nb = A.nbr[(int64_t)cell * NFC + f];
                      const double  sg = A.sigma[(int64_t)cell * NFC + f], sn = s ? 1. : -1.;
                      const double *es = e0 + s * N1, *ds = d0 + s * N1;
                      double        av = 0., bv = 0.;
                      if (nb >= 0)
                        {
                          if (A.flags & PD_ASSEMBLE_INTERIOR)
                            {
                              const double *un  = sN[slot][f] + l - i * stride[d];
                              const double *eo  = e0 + (1 - s) * N1, *dd = d0 + (1 - s) * N1;
                              const double  ihn = 1. / A.cell_h[(int64_t)nb * DIM + d];
                              double        tn = 0., dn = 0.;
#pragma unroll
                              for (int t = 0; t < N1; ++t)
                                {
                                  const double nt = un[t * stride[d]];
                                  tn += eo[t] * nt;
                                  dn += dd[t] * nt;
                                }
                              const double jmp = tu[s] - tn;
                              av               = sg * jmp - 0.5 * sn * (du[s] * ih + dn * ihn);
                              bv               = -0.5 * jmp;
                            }
                        }
                      else if (A.flags & PD_ASSEMBLE_BOUNDARY)
                        {
                          av = sg * tu[s] - sn * du[s] * ih;
                          bv = -tu[s];
                        }
                      acc_d += a * (es[i] * av + sn * ih * ds[i] * bv);
                    }
                  st[d] = A.stiffness * acc_d;
                }
            }
          // ---- masses in the other directions:
          //   3-D: My(Mz W0) + Mx[ Mz W1 + My( W2 + f vol Mz U ) ];  2-D: My W0 + Mx( W1 + f vol My U )
          double acc = 0.;
          if constexpr (DIM == 3)
            {
              if (ok)
                {
                  sW[slot][0][l] = st[0];
                  sW[slot][1][l] = st[1];
                }
              group_sync();
              double t0 = 0., t1 = 0., t2 = 0.;
              if (ok)
                {
                  t0 = contract(Mh, sW[slot][0], 2);
                  t1 = contract(Mh, sW[slot][1], 2);
                  t2 = st[2] + (vol_on && A.mass != 0. ? A.mass * vol * contract(Mh, sU[slot], 2) : 0.);
                }
              group_sync();
              if (ok)
                {
                  sW[slot][2][l] = t0;
                  sW[slot][3][l] = t2;
                }
              group_sync();
              double t4 = 0.;
              if (ok)
                {
                  acc = contract(Mh, sW[slot][2], 1);
                  t4  = t1 + contract(Mh, sW[slot][3], 1);
                }
              group_sync();
              if (ok)
                sW[slot][0][l] = t4;
              group_sync();
              if (ok)
                acc += contract(Mh, sW[slot][0], 0);
            }
          else
            {
              if (ok)
                sW[slot][0][l] = st[0];
              group_sync();
              double t1 = 0.;
              if (ok)
                {
                  acc = contract(Mh, sW[slot][0], 1);
                  t1  = st[1] + (vol_on && A.mass != 0. ? A.mass * vol * contract(Mh, sU[slot], 1) : 0.);
                  sW[slot][1][l] = t1;
                }
              group_sync();
              if (ok)
                acc += contract(Mh, sW[slot][1], 0);
            }
          if (ok)
            {
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
