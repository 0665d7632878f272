// -----------------------------------------------------------------------------
// pd_geometry.cu -- agglomerated quadrature on the device.
//
// Replaces AgglomerationHandler::agglomerated_quadrature
// (source/agglomeration_handler.cc:622-707) and the geometry half of
// reinit_master (:1139-1165): for every sub-cell of every polytope the Gauss
// points and JxW under the Q1 map, for every sub-face of every interface the
// Gauss points, outward normals and surface JxW.  The reference does this with
// one deal.II FEValues::reinit per sub-cell / sub-face on the host.
//
// Here a WARP takes a run of consecutive sub-cells (sub-faces).  Phase 1: one
// lane per (cell, coordinate) gathers the 2^dim vertices and turns them into
// the multilinear coefficients of x_d(xi) = sum_m c_m prod_{k in m} xi_k by
// successive differencing (exactly zero mixed terms on Cartesian cells).
// Phase 2: one thread per quadrature point evaluates x and dx/dxi by nested
// multiplication (11 FMAs per coordinate) and writes the SoA streams fully
// coalesced.  HBM-bound by design: 8(dim+1) B per volume point and 8(2dim+1) B
// per face point written, 2^dim vertices per cell read once.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

#include <algorithm>

namespace pd
{
  namespace
  {
    // c[m], m = bit mask of the unit coordinates in the monomial, for ONE coordinate
    template <int DIM>
    __device__ __forceinline__ void
    multilinear_coefficients(double (&c)[1 << DIM])
    {
      constexpr int VPC = 1 << DIM;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int v = 0; v < VPC; ++v)
          if (v & (1 << d))
            c[v] -= c[v ^ (1 << d)];
    }

    // value and gradient w.r.t. xi of one coordinate from its coefficients (shared memory)
    template <int DIM>
    __device__ __forceinline__ void
    multilinear_eval(const double *c, const double (&xi)[DIM], double &x, double (&dx)[DIM])
    {
      if constexpr (DIM == 2)
        {
          const double a0 = fma(c[1], xi[0], c[0]), a1 = fma(c[3], xi[0], c[2]);
          x     = fma(a1, xi[1], a0);
          dx[0] = fma(c[3], xi[1], c[1]);
          dx[1] = a1;
        }
      else
        {
          const double a0 = fma(c[1], xi[0], c[0]), a1 = fma(c[3], xi[0], c[2]);
          const double a2 = fma(c[5], xi[0], c[4]), a3 = fma(c[7], xi[0], c[6]);
          const double b0 = fma(a1, xi[1], a0), b1 = fma(a3, xi[1], a2);
          x     = fma(b1, xi[2], b0);
          dx[2] = b1;
          dx[1] = fma(a3, xi[2], a1);
          dx[0] = fma(fma(c[7], xi[1], c[5]), xi[2], fma(c[3], xi[1], c[1]));
        }
    }

    // phase 1 shared by both kernels, per WARP (no block barrier: every warp is an
    // independent latency chain, which is what hides the index -> vertex gather):
    // lanes < nslot*DIM compute the coefficients of (slot, coordinate) into sc[slot][d][2^dim]
    template <int DIM>
    __device__ __forceinline__ void
    stage_coefficients(const double *__restrict__ verts,
                       const int32_t *__restrict__ cell_verts,
                       const int32_t *__restrict__ cells, // cell index of every slot of this warp
                       const int nslot,
                       double   *sc)
    {
      constexpr int VPC = 1 << DIM;
      const int     lane = threadIdx.x & 31;
      for (int tv = lane; tv < nslot * DIM; tv += 32)
        {
          const int      sl = tv / DIM, d = tv - sl * DIM;
          const int32_t *cv = cell_verts + (int64_t)cells[sl] * VPC;
          double         c[VPC];
#pragma unroll
          for (int v = 0; v < VPC; ++v)
            c[v] = __ldg(&verts[(int64_t)cv[v] * DIM + d]);
          multilinear_coefficients<DIM>(c);
#pragma unroll
          for (int v = 0; v < VPC; ++v)
            sc[tv * VPC + v] = c[v];
        }
      __syncwarp();
    }

    template <int DIM>
    __global__ void __launch_bounds__(256, 4)
    k_volume_quadrature(const double *__restrict__ verts,
                        const int32_t *__restrict__ cell_verts,
                        const int32_t *__restrict__ subcell_idx,
                        const int64_t n_subcells,
                        const int64_t n_points,
                        const int     nq1,
                        const int     nqc,
                        const int     cpb,
                        const double *__restrict__ rule, // x[8], w[8]
                        double *__restrict__ vq_x,
                        double *__restrict__ vq_w)
    {
      constexpr int VPC = 1 << DIM;
      extern __shared__ double sc_all[]; // [warps][cpb][DIM][VPC], cpb = cells per WARP
      __shared__ double        sq[16];
      if (threadIdx.x < 16)
        sq[threadIdx.x] = __ldg(&rule[threadIdx.x]);
      __syncthreads();
      const int     lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
      double       *sc   = sc_all + warp * cpb * DIM * VPC;
      const int64_t slot0 = ((int64_t)blockIdx.x * nwarp + warp) * cpb;
      if (slot0 >= n_subcells)
        return;
      const int nslot = (int)min((int64_t)cpb, n_subcells - slot0);
      stage_coefficients<DIM>(verts, cell_verts, subcell_idx + slot0, nslot, sc);
      for (int idx = lane; idx < nslot * nqc; idx += 32)
        {
          const int sl = idx / nqc;
          int       q  = idx - sl * nqc;
          double    xi[DIM], w = 1.;
#pragma unroll
          for (int d = 0; d < DIM; ++d) // x fastest (tensor QGauss<dim>)
            {
              const int a = q % nq1;
              q /= nq1;
              xi[d] = sq[a];
              w *= sq[8 + a];
            }
          double x[DIM], J[DIM][DIM];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            multilinear_eval<DIM>(sc + (sl * DIM + d) * VPC, xi, x[d], J[d]);
          double det;
          if constexpr (DIM == 2)
            det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
          else
            det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                  J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
          const int64_t o = slot0 * nqc + idx;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            vq_x[(int64_t)d * n_points + o] = x[d];
          vq_w[o] = w * det;
        }
    }

    // Face rule projected as deal.II does: the face coordinate(s) run along the
    // free axes, 3-D faces 0/1 -> (y,z), 2/3 -> (z,x), 4/5 -> (x,y), first face
    // coordinate fastest.
    template <int DIM>
    __global__ void __launch_bounds__(256, 4)
    k_face_quadrature(const double *__restrict__ verts,
                      const int32_t *__restrict__ cell_verts,
                      const int32_t *__restrict__ sub_cell,
                      const int32_t *__restrict__ sub_face,
                      const int64_t n_subfaces,
                      const int64_t n_points,
                      const int     nq1,
                      const int     nqf,
                      const int     spb,
                      const double *__restrict__ rule,
                      double *__restrict__ fq_x,
                      double *__restrict__ fq_n,
                      double *__restrict__ fq_w)
    {
      constexpr int VPC = 1 << DIM;
      extern __shared__ double sc_all[]; // [warps][spb][DIM][VPC], spb = sub-faces per WARP
      __shared__ double        sq[16];
      if (threadIdx.x < 16)
        sq[threadIdx.x] = __ldg(&rule[threadIdx.x]);
      __syncthreads();
      const int     lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
      double       *sc   = sc_all + warp * spb * DIM * VPC;
      const int64_t slot0 = ((int64_t)blockIdx.x * nwarp + warp) * spb;
      if (slot0 >= n_subfaces)
        return;
      const int nslot = (int)min((int64_t)spb, n_subfaces - slot0);
      stage_coefficients<DIM>(verts, cell_verts, sub_cell + slot0, nslot, sc);
      for (int idx = lane; idx < nslot * nqf; idx += 32)
        {
          const int    sl = idx / nqf, q = idx - sl * nqf;
          const int    f  = sub_face[slot0 + sl];
          const int    nd = f >> 1;
          const double side = (f & 1) ? 1. : 0.;
          double       xi[DIM], w;
          int          t0, t1 = 0;
          if constexpr (DIM == 2)
            {
              t0 = 1 - nd;
              w  = sq[8 + q];
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                xi[d] = d == nd ? side : sq[q];
            }
          else
            {
              t0          = (nd + 1) % 3;
              t1          = (nd + 2) % 3;
              const int a = q % nq1, b = q / nq1;
              w           = sq[8 + a] * sq[8 + b];
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                xi[d] = d == nd ? side : (d == t0 ? sq[a] : sq[b]);
            }
          double x[DIM], J[DIM][DIM];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            multilinear_eval<DIM>(sc + (sl * DIM + d) * VPC, xi, x[d], J[d]);
          // tangents = columns t0 (, t1) of J; outward side fixed by the sign of the
          // component along dx/dxi_nd
          double nrm[DIM], jn[DIM];
          if constexpr (DIM == 2)
            {
              double tx = 0., ty = 0.;
#pragma unroll
              for (int e = 0; e < DIM; ++e)
                {
                  if (e == t0)
                    {
                      tx = J[0][e];
                      ty = J[1][e];
                    }
                  if (e == nd)
                    {
                      jn[0] = J[0][e];
                      jn[1] = J[1][e];
                    }
                }
              nrm[0] = ty;
              nrm[1] = -tx;
            }
          else
            {
              double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
#pragma unroll
              for (int e = 0; e < DIM; ++e)
                {
                  if (e == t0)
                    {
                      a[0] = J[0][e];
                      a[1] = J[1][e];
                      a[2] = J[2][e];
                    }
                  if (e == t1)
                    {
                      b[0] = J[0][e];
                      b[1] = J[1][e];
                      b[2] = J[2][e];
                    }
                  if (e == nd)
                    {
                      jn[0] = J[0][e];
                      jn[1] = J[1][e];
                      jn[2] = J[2][e];
                    }
                }
              nrm[0] = a[1] * b[2] - a[2] * b[1];
              nrm[1] = a[2] * b[0] - a[0] * b[2];
              nrm[2] = a[0] * b[1] - a[1] * b[0];
            }
          double len2 = 0., dotp = 0.;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              len2 += nrm[d] * nrm[d];
              dotp += nrm[d] * jn[d];
            }
          const double  len = sqrt(len2);
          const double  sgn = ((dotp > 0.) == (side > 0.5)) ? 1. : -1.;
          const int64_t o   = slot0 * nqf + idx;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              fq_x[(int64_t)d * n_points + o] = x[d];
              fq_n[(int64_t)d * n_points + o] = sgn * nrm[d] / len;
            }
          fq_w[o] = w * len;
        }
    }
  } // namespace

  void
  launch_quadrature(pd_handle *h)
  {
    const int tb = 256;
    if (h->Q > 0)
      {
        // cells per warp: one lane per (cell, coordinate) gathers vertices, so 32/dim cells share ONE
        // index->vertex latency round trip (the kernel is latency-bound otherwise)
        const int      cpb  = 32 / h->dim;
        const int64_t  per_block = (int64_t)cpb * (tb / 32);
        const unsigned nb   = (unsigned)((h->n_subcells + per_block - 1) / per_block);
        const size_t   smem = sizeof(double) * per_block * (1 << h->dim) * h->dim;
        if (h->dim == 2)
          k_volume_quadrature<2><<<nb, tb, smem, h->stream>>>(h->verts.p, h->cell_verts.p, h->subcell_idx.p,
                                                             h->n_subcells, h->Q, h->nq1, h->nqc, cpb, h->rules.p,
                                                             h->vq_x.p, h->vq_w.p);
        else
          k_volume_quadrature<3><<<nb, tb, smem, h->stream>>>(h->verts.p, h->cell_verts.p, h->subcell_idx.p,
                                                             h->n_subcells, h->Q, h->nq1, h->nqc, cpb, h->rules.p,
                                                             h->vq_x.p, h->vq_w.p);
        ++h->launches;
      }
    if (h->Qf > 0)
      {
        const int      spb  = 32 / h->dim; // sub-faces per warp
        const int64_t  per_block = (int64_t)spb * (tb / 32);
        const unsigned nb   = (unsigned)((h->n_subfaces + per_block - 1) / per_block);
        const size_t   smem = sizeof(double) * per_block * (1 << h->dim) * h->dim;
        if (h->dim == 2)
          k_face_quadrature<2><<<nb, tb, smem, h->stream>>>(h->verts.p, h->cell_verts.p, h->sub_cell.p, h->sub_face.p,
                                                           h->n_subfaces, h->Qf, h->nq1f, h->nqf, spb,
                                                           h->rules.p + 16, h->fq_x.p, h->fq_n.p, h->fq_w.p);
        else
          k_face_quadrature<3><<<nb, tb, smem, h->stream>>>(h->verts.p, h->cell_verts.p, h->sub_cell.p, h->sub_face.p,
                                                           h->n_subfaces, h->Qf, h->nq1f, h->nqf, spb,
                                                           h->rules.p + 16, h->fq_x.p, h->fq_n.p, h->fq_w.p);
        ++h->launches;
      }
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
