// -----------------------------------------------------------------------------
// pd_reinit.cu -- the reinit() family through the C ABI: the tables the reference's
// hand-written loops read, materialised for ONE polytope / polytope face at a time.
//
// Reference: AgglomerationHandler::reinit(polytope) (source/agglomeration_handler.cc:
// 729-767: FEValues of the element on the bounding box at the agglomerated quadrature),
// reinit(polytope, f) and reinit_interface (:785-906 -> reinit_master :1103-1243:
// FEImmersedSurfaceValues with the face points, normals and JxW of the sub-faces),
// MappingBox (source/mapping_box.cc:393-439: value = phihat(xhat), gradient =
// grad-hat / h_bbox per coordinate; JxW and normals as given) and the element families
// FE_DGQ / FE_AggloDGP (source/fe_agglodgp.cc:28-57).
//
// The assembly kernels never materialise these tables (they are generated per 32-point
// stage in shared memory); this file exists so that a reference-style loop
// (examples/poisson.cc:745-905) can be pointed at the library unchanged.  One thread per
// quadrature point evaluates the 1-D factors in registers and writes every DoF's value
// and gradient (coalesced over the points); not a hot path.
// Layout = FEValues accessors: values[i * Q + q] = shape_value(i, q),
// grads[(i * Q + q) * dim + d] = shape_grad(i, q)[d], jxw[q], points[q * dim + d],
// normals[q * dim + d].
// -----------------------------------------------------------------------------
#include "pd_host.hpp"
#include "pd_internal.hpp"

#include <algorithm>
#include <cstring>

namespace pd
{
  namespace
  {
    struct ReinitArgs
    {
      int           dim, n1, dgp, n;
      Basis1D       B;
      const double *x[3], *nrm[3], *w; // SoA streams, already offset to the first point of the item
      double        lo[3], inv_h[3];
      int64_t       Q;
      double        normal_sign;
      double       *values, *grads, *jxw, *points, *normals, *unit_points;
    };

    // 1-D factors at xhat: Lagrange on the Gauss-Lobatto nodes (product form) or L2[0,1]-orthonormal Legendre
    __device__ void
    factors_1d(const ReinitArgs &A, const double xh, const double scale, double *L, double *dL)
    {
      if (!A.dgp)
        {
          for (int a = 0; a < A.n1; ++a)
            {
              double val = 1., der = 0.;
              for (int b = 0; b < A.n1; ++b)
                if (b != a)
                  {
                    const double t = xh - A.B.node[b];
                    der            = der * t + val;
                    val            = val * t;
                  }
              L[a]  = val * A.B.wprod[a];
              dL[a] = der * A.B.wprod[a] * scale;
            }
          return;
        }
      const double t = 2. * xh - 1.;
      double       p0 = 1., p1 = t, d0 = 0., d1 = 1.;
      for (int k = 0; k < A.n1; ++k)
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

    __global__ void __launch_bounds__(128)
    k_reinit_tables(const ReinitArgs A)
    {
      const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (q >= A.Q)
        return;
      double L[3][6], dL[3][6], xr[3];
      for (int d = 0; d < 3; ++d)
        {
          L[d][0]  = 1.;
          dL[d][0] = 0.;
        }
      for (int d = 0; d < A.dim; ++d)
        {
          xr[d]           = A.x[d][q];
          const double xh = (xr[d] - A.lo[d]) * A.inv_h[d]; // BoundingBox::real_to_unit
          factors_1d(A, xh, A.inv_h[d], L[d], dL[d]);
          if (A.points)
            A.points[q * A.dim + d] = xr[d];
          if (A.unit_points)
            A.unit_points[q * A.dim + d] = xh;
          if (A.normals)
            A.normals[q * A.dim + d] = A.normal_sign * A.nrm[d][q];
        }
      if (A.jxw)
        A.jxw[q] = A.w[q];
      if (!A.values && !A.grads)
        return;
      const int p  = A.n1 - 1;
      const int nc = A.dim == 3 ? A.n1 : 1;
      int       i  = 0;
      for (int c = 0; c < nc; ++c)
        for (int b = 0; b < A.n1; ++b)
          for (int a = 0; a < A.n1; ++a)
            {
              if (A.dgp && a + b + c > p)
                continue; // PolynomialSpace: total degree <= p, first coordinate fastest
              const double vyz = L[1][b] * L[2][c];
              if (A.values)
                A.values[(int64_t)i * A.Q + q] = L[0][a] * vyz;
              if (A.grads)
                {
                  double *g = A.grads + ((int64_t)i * A.Q + q) * A.dim;
                  g[0]      = dL[0][a] * vyz;
                  g[1]      = L[0][a] * dL[1][b] * L[2][c];
                  if (A.dim == 3)
                    g[2] = L[0][a] * L[1][b] * dL[2][c];
                }
              ++i;
            }
    }

    // tables of `Q` points into scratch, then to the caller's (host or device) pointers
    void
    run(pd_handle *h, ReinitArgs A, const int32_t poly, double *values, double *grads, double *jxw, double *points,
        double *normals, double *unit_points)
    {
      const int     dim = h->dim, n = h->n;
      const int64_t Q   = A.Q;
      if (Q == 0)
        return;
      A.dim = dim;
      A.n1  = h->n1;
      A.dgp = h->fe_kind == PD_FE_AGGLODGP;
      A.n   = n;
      A.B   = h->basis;
      const double *bb = &h->h_bbox[(size_t)poly * 2 * dim];
      for (int d = 0; d < dim; ++d)
        {
          A.lo[d]    = bb[d];
          A.inv_h[d] = 1. / (bb[dim + d] - bb[d]);
        }
      // scratch: values n Q | grads n Q dim | jxw Q | points Q dim | normals Q dim | unit points Q dim
      const size_t need = (size_t)Q * ((size_t)n * (1 + dim) + 1 + 3 * dim);
      if (h->reinit_scratch.n < need)
        {
          PD_CUDA(cudaStreamSynchronize(h->stream));
          h->reinit_scratch.alloc(need + need / 4);
        }
      double *s = h->reinit_scratch.p;
      A.values  = values ? s : nullptr;
      s += (size_t)Q * n;
      A.grads = grads ? s : nullptr;
      s += (size_t)Q * n * dim;
      A.jxw = jxw ? s : nullptr;
      s += Q;
      A.points = points ? s : nullptr;
      s += (size_t)Q * dim;
      A.normals = normals ? s : nullptr;
      s += (size_t)Q * dim;
      A.unit_points = unit_points ? s : nullptr;
      k_reinit_tables<<<(unsigned)((Q + 127) / 128), 128, 0, h->stream>>>(A);
      ++h->launches;
      PD_CUDA(cudaGetLastError());
      auto out = [&](double *dst, const double *src, const size_t count) {
        if (dst)
          PD_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDefault, h->stream));
      };
      out(values, A.values, (size_t)Q * n);
      out(grads, A.grads, (size_t)Q * n * dim);
      out(jxw, A.jxw, (size_t)Q);
      out(points, A.points, (size_t)Q * dim);
      out(normals, A.normals, (size_t)Q * dim);
      out(unit_points, A.unit_points, (size_t)Q * dim);
      PD_CUDA(cudaStreamSynchronize(h->stream)); // like the reference: the tables are valid when reinit returns
    }
  } // namespace

  int64_t
  reinit_n_points(const pd_handle *h, const int32_t poly)
  {
    if (poly < 0 || poly >= h->np_own)
      throw Error(PD_ERR_INVALID, "polytope index out of range (owned polytopes only)");
    return (h->h_subcell_ptr[poly + 1] - h->h_subcell_ptr[poly]) * h->nqc;
  }

  int64_t
  reinit_iface_n_points(const pd_handle *h, const int32_t iface)
  {
    if (iface < 0 || iface >= h->n_ifaces)
      throw Error(PD_ERR_INVALID, "interface index out of range");
    return (h->h_if_sub_ptr[iface + 1] - h->h_if_sub_ptr[iface]) * h->nqf;
  }

  // reinit(polytope): FEValues on the agglomerated quadrature of the polytope (quadrature must be valid)
  void
  reinit_polytope(pd_handle *h, const int32_t poly, double *values, double *grads, double *jxw, double *points,
                  double *unit_points)
  {
    ReinitArgs    A{};
    const int64_t q0 = h->h_subcell_ptr[poly] * h->nqc;
    A.Q              = reinit_n_points(h, poly);
    for (int d = 0; d < h->dim; ++d)
      {
        A.x[d]   = h->vq_x.p + (size_t)d * h->Q + q0;
        A.nrm[d] = nullptr;
      }
    A.w           = h->vq_w.p + q0;
    A.normal_sign = 0.;
    run(h, A, poly, values, grads, jxw, points, nullptr, unit_points);
  }

  // reinit(polytope, f) / one side of reinit_interface: side 0 = the listing polytope A (its outward normals),
  // side 1 = the neighbour B at the SAME points (aligned, source/agglomeration_handler.cc:1375-1397) with its own
  // outward normals = -n_A and its own bounding box
  void
  reinit_iface(pd_handle *h, const int32_t iface, const int side, double *values, double *grads, double *jxw,
               double *points, double *normals)
  {
    ReinitArgs    A{};
    const int64_t q0 = h->h_if_sub_ptr[iface] * h->nqf;
    A.Q              = reinit_iface_n_points(h, iface);
    const int32_t poly = side == 0 ? h->h_ifA[iface] : h->h_ifB[iface];
    if (side != 0 && side != 1)
      throw Error(PD_ERR_INVALID, "side must be 0 or 1");
    if (poly < 0)
      throw Error(PD_ERR_INVALID, "a boundary face has no second side");
    for (int d = 0; d < h->dim; ++d)
      {
        A.x[d]   = h->fq_x.p + (size_t)d * h->Qf + q0;
        A.nrm[d] = h->fq_n.p + (size_t)d * h->Qf + q0;
      }
    A.w           = h->fq_w.p + q0;
    A.normal_sign = side == 0 ? 1. : -1.;
    run(h, A, poly, values, grads, jxw, points, normals, nullptr);
  }

  // FE_DGQ / FE_AggloDGP on the unit cell at arbitrary unit points (AoS [n_points][dim], host or device)
  void
  fe_evaluate(const int fe_kind, const int dim, const int degree, const int64_t n_points, const double *unit_points,
              double *values, double *grads)
  {
    if ((dim != 2 && dim != 3) || degree < 0 || degree > 5 || n_points < 0 || (fe_kind != PD_FE_DGQ && fe_kind != PD_FE_AGGLODGP))
      throw Error(PD_ERR_INVALID, "pd_fe_evaluate: bad argument");
    if (n_points == 0)
      return;
    pd_handle tmp; // only the fields run() reads; no device state of its own besides the scratch
    tmp.dim     = dim;
    tmp.degree  = degree;
    tmp.n1      = degree + 1;
    tmp.fe_kind = fe_kind;
    tmp.n       = 1;
    if (fe_kind == PD_FE_DGQ)
      for (int k = 0; k < dim; ++k)
        tmp.n *= degree + 1;
    else
      for (int k = 1; k <= dim; ++k)
        tmp.n = tmp.n * (degree + k) / k;
    make_basis_1d(degree, tmp.basis);
    tmp.h_bbox.assign(2 * dim, 0.);
    for (int d = 0; d < dim; ++d)
      tmp.h_bbox[dim + d] = 1.;
    tmp.stream = nullptr;
    // transpose the AoS unit points into SoA streams on the device
    DevBuf<double>      soa;
    std::vector<double> host((size_t)n_points * dim), t((size_t)n_points * dim);
    PD_CUDA(cudaMemcpy(host.data(), unit_points, host.size() * sizeof(double), cudaMemcpyDefault));
    for (int64_t q = 0; q < n_points; ++q)
      for (int d = 0; d < dim; ++d)
        t[(size_t)d * n_points + q] = host[(size_t)q * dim + d];
    soa.alloc(t.size() + (size_t)n_points);
    PD_CUDA(cudaMemcpy(soa.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    ReinitArgs A{};
    A.Q = n_points;
    for (int d = 0; d < dim; ++d)
      A.x[d] = soa.p + (size_t)d * n_points;
    A.w           = soa.p + t.size();
    A.normal_sign = 0.;
    run(&tmp, A, 0, values, grads, nullptr, nullptr, nullptr, nullptr);
  }
} // namespace pd
