// -----------------------------------------------------------------------------
// pd_geometry.cu -- agglomerated quadrature on the device.
//
// Replaces AgglomerationHandler::agglomerated_quadrature
// (source/agglomeration_handler.cc:622-707) and the geometry half of
// reinit_master (:1139-1165): for every sub-cell of every polytope the Gauss
// points and JxW under the Q1 map, for every sub-face of every interface the
// Gauss points, outward normals and surface JxW.  The reference does this with
// one deal.II FEValues::reinit per sub-cell / sub-face on the host; here one
// thread computes one quadrature point and the results are written as SoA
// streams that the assembly kernels read fully coalesced.
//
// HBM-bound: reads 2^dim vertices per sub-cell (L1/L2 hits after the first
// point of the cell), writes 8(dim+1) B per volume point and 8(2 dim+1) B per
// face point.
// -----------------------------------------------------------------------------
#include "pd_internal.hpp"

#include <algorithm>

namespace pd
{
  namespace
  {
    template <int DIM>
    struct CellVerts
    {
      double x[1 << DIM][DIM];
    };

    template <int DIM>
    __device__ __forceinline__ void
    load_cell(const double *__restrict__ verts, const int32_t *__restrict__ cell_verts, const int32_t cell, CellVerts<DIM> &cv)
    {
      constexpr int VPC = 1 << DIM;
#pragma unroll
      for (int v = 0; v < VPC; ++v)
        {
          const int64_t vi = cell_verts[(int64_t)cell * VPC + v];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            cv.x[v][d] = __ldg(&verts[vi * DIM + d]);
        }
    }

    // x(xi) and J[a][b] = dx_a/dxi_b of the multilinear map
    template <int DIM>
    __device__ __forceinline__ void
    q1_map(const CellVerts<DIM> &cv, const double (&xi)[DIM], double (&x)[DIM], double (&J)[DIM][DIM])
    {
      constexpr int VPC = 1 << DIM;
#pragma unroll
      for (int a = 0; a < DIM; ++a)
        {
          x[a] = 0.;
#pragma unroll
          for (int b = 0; b < DIM; ++b)
            J[a][b] = 0.;
        }
#pragma unroll
      for (int v = 0; v < VPC; ++v)
        {
          double f[DIM], s[DIM];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              const bool up = (v >> d) & 1;
              f[d]          = up ? xi[d] : 1. - xi[d];
              s[d]          = up ? 1. : -1.;
            }
          double N = 1.;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            N *= f[d];
          double dN[DIM];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              double g = s[d];
#pragma unroll
              for (int e = 0; e < DIM; ++e)
                if (e != d)
                  g *= f[e];
              dN[d] = g;
            }
#pragma unroll
          for (int a = 0; a < DIM; ++a)
            {
              x[a] += N * cv.x[v][a];
#pragma unroll
              for (int b = 0; b < DIM; ++b)
                J[a][b] += dN[b] * cv.x[v][a];
            }
        }
    }

    // The 1-D rules live in a small device array (x[8], w[8]): indexing a kernel
    // PARAMETER array with a runtime index makes the compiler copy the whole struct to
    // local memory in every thread (16 STL + LDLs per point, seen in the SASS).
    // One block handles CPB consecutive sub-cells: their vertices are fetched ONCE into
    // shared memory (one thread per vertex), then one thread computes one point.
    template <int DIM>
    __global__ void __launch_bounds__(256)
    k_volume_quadrature(const double *__restrict__ verts,
                        const int32_t *__restrict__ cell_verts,
                        const int32_t *__restrict__ subcell_idx,
                        const int64_t n_subcells,
                        const int64_t n_points,
                        const int     nq1,
                        const int     nqc,
                        const int     cpb,
                        const double *__restrict__ rule,
                        double *__restrict__ vq_x,
                        double *__restrict__ vq_w)
    {
      constexpr int VPC = 1 << DIM;
      extern __shared__ double sv[]; // [cpb][VPC][DIM]
      __shared__ double        sq[16];
      if (threadIdx.x < 16)
        sq[threadIdx.x] = __ldg(&rule[threadIdx.x]);
      const int64_t slot0 = (int64_t)blockIdx.x * cpb;
      const int     nslot = (int)min((int64_t)cpb, n_subcells - slot0);
      for (int tv = threadIdx.x; tv < nslot * VPC; tv += blockDim.x)
        {
          const int     sl = tv / VPC, v = tv - sl * VPC;
          const int64_t vi = cell_verts[(int64_t)subcell_idx[slot0 + sl] * VPC + v];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            sv[tv * DIM + d] = __ldg(&verts[vi * DIM + d]);
        }
      __syncthreads();
      for (int idx = threadIdx.x; idx < nslot * nqc; idx += blockDim.x)
        {
          const int      sl = idx / nqc;
          int            q  = idx - sl * nqc;
          CellVerts<DIM> cv;
#pragma unroll
          for (int v = 0; v < VPC; ++v)
#pragma unroll
            for (int d = 0; d < DIM; ++d)
              cv.x[v][d] = sv[(sl * VPC + v) * DIM + d];
          double xi[DIM], w = 1.;
#pragma unroll
          for (int d = 0; d < DIM; ++d) // x fastest (tensor QGauss<dim>)
            {
              const int a = q % nq1;
              q /= nq1;
              xi[d] = sq[a];
              w *= sq[8 + a];
            }
          double x[DIM], J[DIM][DIM];
          q1_map<DIM>(cv, xi, x, J);
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
    __global__ void __launch_bounds__(256)
    k_face_quadrature(const double *__restrict__ verts,
                      const int32_t *__restrict__ cell_verts,
                      const int32_t *__restrict__ sub_cell,
                      const int32_t *__restrict__ sub_face,
                      const int64_t n_points,
                      const int     nq1,
                      const int     nqf,
                      const double *__restrict__ rule,
                      double *__restrict__ fq_x,
                      double *__restrict__ fq_n,
                      double *__restrict__ fq_w)
    {
      __shared__ double sq[16];
      if (threadIdx.x < 16)
        sq[threadIdx.x] = __ldg(&rule[threadIdx.x]);
      __syncthreads();
      const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (idx >= n_points)
        return;
      const int64_t s  = idx / nqf;
      const int     q  = (int)(idx - s * nqf);
      const int     f  = sub_face[s];
      const int     nd = f >> 1;
      const double  side = (f & 1) ? 1. : 0.;
      CellVerts<DIM> cv;
      load_cell<DIM>(verts, cell_verts, sub_cell[s], cv);
      double xi[DIM], w;
      int    t0, t1 = 0;
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
      q1_map<DIM>(cv, xi, x, J);
      double nrm[DIM];
      if constexpr (DIM == 2)
        {
          double tx = 0., ty = 0.;
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            if (d == t0)
              {
                tx = J[0][d];
                ty = J[1][d];
              }
          nrm[0] = ty;
          nrm[1] = -tx;
        }
      else
        {
          double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              if (d == t0)
                {
                  a[0] = J[0][d];
                  a[1] = J[1][d];
                  a[2] = J[2][d];
                }
              if (d == t1)
                {
                  b[0] = J[0][d];
                  b[1] = J[1][d];
                  b[2] = J[2][d];
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
          double jn = 0.;
#pragma unroll
          for (int e = 0; e < DIM; ++e)
            if (e == nd)
              jn = J[d][e];
          dotp += nrm[d] * jn;
        }
      const double len = sqrt(len2);
      const double sgn = ((dotp > 0.) == (side > 0.5)) ? 1. : -1.;
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        {
          fq_x[(int64_t)d * n_points + idx] = x[d];
          fq_n[(int64_t)d * n_points + idx] = sgn * nrm[d] / len;
        }
      fq_w[idx] = w * len;
    }
  } // namespace

  void
  launch_quadrature(pd_handle *h)
  {
    const int tb = 256;
    if (h->Q > 0)
      {
        const int      cpb  = std::max(1, tb / h->nqc);
        const unsigned nb   = (unsigned)((h->n_subcells + cpb - 1) / cpb);
        const size_t   smem = sizeof(double) * cpb * (1 << h->dim) * h->dim;
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
        const unsigned nb = (unsigned)((h->Qf + tb - 1) / tb);
        if (h->dim == 2)
          k_face_quadrature<2><<<nb, tb, 0, h->stream>>>(h->verts.p, h->cell_verts.p, h->sub_cell.p, h->sub_face.p,
                                                        h->Qf, h->nq1f, h->nqf, h->rules.p + 16, h->fq_x.p, h->fq_n.p,
                                                        h->fq_w.p);
        else
          k_face_quadrature<3><<<nb, tb, 0, h->stream>>>(h->verts.p, h->cell_verts.p, h->sub_cell.p, h->sub_face.p,
                                                        h->Qf, h->nq1f, h->nqf, h->rules.p + 16, h->fq_x.p, h->fq_n.p,
                                                        h->fq_w.p);
        ++h->launches;
      }
    PD_CUDA(cudaGetLastError());
  }
} // namespace pd
