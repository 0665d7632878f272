// =============================================================================
//  oracle/polyoracle.hpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
//
//  CPU restatement of polyDEAL's hot path (SIP-DG assembly over agglomerated
//  polytopes + operator apply).  Only tests/, __graft_entry__.smoke() and
//  bench.py's cpu_baseline / --impl reference legs may use anything in oracle/.
//  The product (polydeal_b200/) never includes, links or calls this.
//
//  The reference (/root/reference, polyDEAL on deal.II >= 9.7) cannot be built
//  here (no deal.II / Trilinos / MPI / Boost), so this file restates
//    source/agglomeration_handler.cc   45-104   define_agglomerate
//                                      476-491  create_bounding_box
//                                      622-707  agglomerated_quadrature
//                                      729-801  reinit(polytope[,f])
//                                      805-834  reinit_interface (owned/owned)
//                                      910-1022 create_agglomeration_sparsity_pattern
//                                      1103-1243 reinit_master
//                                      1253-1645 setup_master_neighbor_connectivity
//    include/agglomeration_accessor.h  335-481  neighbor / neighbor_of_agglomerated_neighbor
//                                      562-601  get_agglomerate / diameter
//    source/mapping_box.cc             393-439, 465-503, 522-532
//    source/fe_agglodgp.cc             28-57, 91-101
//    include/poly_utils.h              1870-1926 jumps/averages, 2000-2195 assemble_dg_matrix
//    include/utils.h                   819-925  LaplaceOperatorDG (as a matrix, see
//                                               examples/monodomain_DG3D.cc:1470-1498)
//  and the deal.II behaviours it relies on (hyper_cube/refine_global cell order,
//  reference-cell face numbering, QGauss, FE_DGQ on Gauss-Lobatto nodes, DG
//  DoF numbering in active-cell order), which are not citable in-tree and are
//  pinned through the reference's own test goldens (tests/golden/).
//
//  Parity status: structure (numbering, sparsity, face lists) PINNED against
//  reference goldens; SIP arithmetic pinned through the reference tests'
//  invariants (minimal_SIP_Poisson equality, poisson_sanity_check energies,
//  exactness); assemble_dg_matrix / LaplaceOperatorDG numbers themselves have
//  no golden in the reference => "parity unpinned" for those two.
// =============================================================================
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace po
{
  constexpr unsigned int invalid_uint = std::numeric_limits<unsigned int>::max();

  // ---------------------------------------------------------------------------
  // 1-D rules (deal.II QGauss<1>, QGaussLobatto<1> on [0,1])
  // ---------------------------------------------------------------------------
  inline void
  legendre(const int n, const long double x, long double &p, long double &dp)
  {
    // P_n(x) and P_n'(x) on [-1,1] by the three-term recurrence
    long double p0 = 1.0L, p1 = x;
    if (n == 0)
      {
        p  = 1.0L;
        dp = 0.0L;
        return;
      }
    for (int k = 2; k <= n; ++k)
      {
        const long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0                   = p1;
        p1                   = pk;
      }
    p  = p1;
    dp = n * (x * p1 - p0) / (x * x - 1.0L);
  }

  inline void
  gauss_1d(const int n, std::vector<double> &pts, std::vector<double> &wts)
  {
    pts.assign(n, 0.);
    wts.assign(n, 0.);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < n; ++i)
      {
        long double x = -std::cos(pi * (i + 0.75L) / (n + 0.5L));
        for (int it = 0; it < 100; ++it)
          {
            long double p, dp;
            legendre(n, x, p, dp);
            const long double dx = p / dp;
            x -= dx;
            if (std::fabs((double)dx) < 1e-20)
              break;
          }
        long double p, dp;
        legendre(n, x, p, dp);
        pts[i] = (double)((x + 1.0L) / 2.0L);
        wts[i] = (double)(1.0L / ((1.0L - x * x) * dp * dp));
      }
    // enforce exact symmetry about 1/2 like a careful table would
    for (int i = 0; i < n / 2; ++i)
      {
        const double w  = 0.5 * (wts[i] + wts[n - 1 - i]);
        wts[i]          = w;
        wts[n - 1 - i]  = w;
        const double d  = 0.5 * ((0.5 - pts[i]) + (pts[n - 1 - i] - 0.5));
        pts[i]          = 0.5 - d;
        pts[n - 1 - i]  = 0.5 + d;
      }
    if (n % 2 == 1)
      pts[n / 2] = 0.5;
  }

  inline std::vector<double>
  gauss_lobatto_nodes(const int n) // n points on [0,1], n >= 2
  {
    std::vector<double> x(n);
    x[0]     = 0.;
    x[n - 1] = 1.;
    const int         m  = n - 1; // interior nodes = roots of P_m'
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 1; i < n - 1; ++i)
      {
        long double t = -std::cos(pi * i / m);
        for (int it = 0; it < 100; ++it)
          {
            // Newton on P_m'(t); P_m'' from the Legendre ODE
            long double p, dp;
            legendre(m, t, p, dp);
            const long double ddp =
              (2 * t * dp - m * (m + 1) * p) / (1.0L - t * t);
            const long double dt = dp / ddp;
            t -= dt;
            if (std::fabs((double)dt) < 1e-20)
              break;
          }
        x[i] = (double)((t + 1.0L) / 2.0L);
      }
    for (int i = 1; i < n / 2; ++i)
      {
        const double d = 0.5 * ((0.5 - x[i]) + (x[n - 1 - i] - 0.5));
        x[i]           = 0.5 - d;
        x[n - 1 - i]   = 0.5 + d;
      }
    if (n % 2 == 1)
      x[n / 2] = 0.5;
    return x;
  }

  // ---------------------------------------------------------------------------
  // Finite elements on the unit box [0,1]^d
  // ---------------------------------------------------------------------------
  enum FEKind
  {
    FE_DGQ      = 0, // deal.II FE_DGQ<dim>(p): tensor Lagrange on GLL nodes
    FE_AGGLODGP = 1  // source/fe_agglodgp.cc:28-57: orthonormal Legendre, |alpha|<=p
  };

  struct FiniteElement
  {
    int                 kind   = FE_DGQ;
    int                 dim    = 2;
    int                 degree = 1;
    int                 n_dofs = 0;
    std::vector<double> nodes;               // DGQ 1-D nodes
    std::vector<double> lag_w;               // DGQ barycentric-free product weights
    std::vector<std::array<int, 3>> multi;   // DGP multi-indices

    void
    init(const int kind_, const int dim_, const int p)
    {
      kind   = kind_;
      dim    = dim_;
      degree = p;
      if (kind == FE_DGQ)
        {
          if (p == 0)
            nodes = {0.5};
          else
            nodes = gauss_lobatto_nodes(p + 1);
          const int m = p + 1;
          lag_w.assign(m, 1.);
          for (int a = 0; a < m; ++a)
            for (int b = 0; b < m; ++b)
              if (a != b)
                lag_w[a] /= (nodes[a] - nodes[b]);
          n_dofs = 1;
          for (int d = 0; d < dim; ++d)
            n_dofs *= m;
        }
      else
        {
          // deal.II PolynomialSpace ordering: last coordinate outermost,
          // first coordinate fastest, total degree <= p.
          multi.clear();
          if (dim == 2)
            {
              for (int iy = 0; iy <= p; ++iy)
                for (int ix = 0; ix + iy <= p; ++ix)
                  multi.push_back({ix, iy, 0});
            }
          else
            {
              for (int iz = 0; iz <= p; ++iz)
                for (int iy = 0; iy + iz <= p; ++iy)
                  for (int ix = 0; ix + iy + iz <= p; ++ix)
                    multi.push_back({ix, iy, iz});
            }
          n_dofs = (int)multi.size();
        }
    }

    // 1-D Lagrange value/derivative in product form
    void
    lagrange_1d(const double x, double *v, double *dv) const
    {
      const int m = degree + 1;
      if (m == 1)
        {
          v[0]  = 1.;
          dv[0] = 0.;
          return;
        }
      for (int a = 0; a < m; ++a)
        {
          double val = 1., der = 0.;
          for (int b = 0; b < m; ++b)
            if (b != a)
              {
                // d/dx of running product
                der = der * (x - nodes[b]) + val;
                val = val * (x - nodes[b]);
              }
          v[a]  = val * lag_w[a];
          dv[a] = der * lag_w[a];
        }
    }

    // orthonormal Legendre on [0,1]: sqrt(2k+1) P_k(2x-1)
    static void
    legendre01(const int kmax, const double x, double *v, double *dv)
    {
      const double t = 2. * x - 1.;
      double       p0 = 1., p1 = t, d0 = 0., d1 = 1.;
      for (int k = 0; k <= kmax; ++k)
        {
          double pk, dk;
          if (k == 0)
            {
              pk = p0;
              dk = d0;
            }
          else if (k == 1)
            {
              pk = p1;
              dk = d1;
            }
          else
            {
              pk = ((2 * k - 1) * t * p1 - (k - 1) * p0) / k;
              dk = ((2 * k - 1) * (p1 + t * d1) - (k - 1) * d0) / k;
              p0 = p1;
              p1 = pk;
              d0 = d1;
              d1 = dk;
            }
          const double s = std::sqrt(2. * k + 1.);
          v[k]           = s * pk;
          dv[k]          = s * dk * 2.; // chain rule d(2x-1)/dx
        }
    }

    // values[n], grads[n*dim] w.r.t. unit coordinates
    void
    evaluate(const double *xhat, double *values, double *grads) const
    {
      double v[3][16], dv[3][16];
      if (kind == FE_DGQ)
        {
          const int m = degree + 1;
          for (int d = 0; d < dim; ++d)
            lagrange_1d(xhat[d], v[d], dv[d]);
          int i = 0;
          if (dim == 2)
            {
              for (int b = 0; b < m; ++b)
                for (int a = 0; a < m; ++a, ++i)
                  {
                    values[i]        = v[0][a] * v[1][b];
                    grads[i * 2 + 0] = dv[0][a] * v[1][b];
                    grads[i * 2 + 1] = v[0][a] * dv[1][b];
                  }
            }
          else
            {
              for (int c = 0; c < m; ++c)
                for (int b = 0; b < m; ++b)
                  for (int a = 0; a < m; ++a, ++i)
                    {
                      values[i]        = v[0][a] * v[1][b] * v[2][c];
                      grads[i * 3 + 0] = dv[0][a] * v[1][b] * v[2][c];
                      grads[i * 3 + 1] = v[0][a] * dv[1][b] * v[2][c];
                      grads[i * 3 + 2] = v[0][a] * v[1][b] * dv[2][c];
                    }
            }
        }
      else
        {
          for (int d = 0; d < dim; ++d)
            legendre01(degree, xhat[d], v[d], dv[d]);
          for (int i = 0; i < n_dofs; ++i)
            {
              const auto &mi = multi[i];
              if (dim == 2)
                {
                  values[i]        = v[0][mi[0]] * v[1][mi[1]];
                  grads[i * 2 + 0] = dv[0][mi[0]] * v[1][mi[1]];
                  grads[i * 2 + 1] = v[0][mi[0]] * dv[1][mi[1]];
                }
              else
                {
                  values[i]        = v[0][mi[0]] * v[1][mi[1]] * v[2][mi[2]];
                  grads[i * 3 + 0] = dv[0][mi[0]] * v[1][mi[1]] * v[2][mi[2]];
                  grads[i * 3 + 1] = v[0][mi[0]] * dv[1][mi[1]] * v[2][mi[2]];
                  grads[i * 3 + 2] = v[0][mi[0]] * v[1][mi[1]] * dv[2][mi[2]];
                }
            }
        }
    }

    // DGQ support points on the unit box (for interpolation of test functions)
    void
    unit_support_point(const int i, double *xhat) const
    {
      const int m = degree + 1;
      int       r = i;
      for (int d = 0; d < dim; ++d)
        {
          xhat[d] = nodes[r % m];
          r /= m;
        }
    }
  };

  // ---------------------------------------------------------------------------
  // Grid: hypercube cells, deal.II conventions
  //   vertices lexicographic in a cell; faces 0:x-,1:x+,2:y-,3:y+,4:z-,5:z+
  // ---------------------------------------------------------------------------
  struct Grid
  {
    int                 dim = 2;
    std::vector<double> verts;      // n_verts * dim
    std::vector<int>    cell_verts; // n_cells * 2^dim
    std::vector<int>    nbr;        // n_cells * 2*dim, -1 on the boundary
    std::vector<char>   vert_on_boundary;

    int
    n_cells() const
    {
      return (int)(cell_verts.size() >> dim);
    }
    int
    n_verts() const
    {
      return (int)(verts.size() / dim);
    }
    const double *
    vertex(const int c, const int v) const
    {
      return &verts[(size_t)cell_verts[((size_t)c << dim) + v] * dim];
    }
    int
    neighbor(const int c, const int f) const
    {
      return nbr[(size_t)c * 2 * dim + f];
    }
    // cell->neighbor_of_neighbor(f) of deal.II: the face through which the neighbour across face f sees
    // this cell.  The opposite face on structured grids; on unstructured meshes (GridIn) the neighbour may
    // be rotated, so look it up.
    int
    neighbor_of_neighbor(const int c, const int f) const
    {
      const int q = neighbor(c, f);
      if (q >= 0 && neighbor(q, f ^ 1) == c)
        return f ^ 1;
      for (int g = 0; q >= 0 && g < 2 * dim; ++g)
        if (neighbor(q, g) == c)
          return g;
      throw std::runtime_error("Grid::neighbor_of_neighbor: the neighbour table is not symmetric");
    }

    // any conforming quadrilateral / hexahedral mesh from its arrays (vertices of a cell in deal.II's
    // lexicographic order, faces 2 * direction + side); the boundary vertices are found from the faces
    void
    build_from_arrays(const int dim_, const int n_verts_, const double *v, const int n_cells_, const int *cv, const int *nb)
    {
      dim = dim_;
      verts.assign(v, v + (size_t)n_verts_ * dim);
      cell_verts.assign(cv, cv + ((size_t)n_cells_ << dim));
      nbr.assign(nb, nb + (size_t)n_cells_ * 2 * dim);
      vert_on_boundary.assign((size_t)n_verts_, 0);
      for (int c = 0; c < n_cells_; ++c)
        for (int f = 0; f < 2 * dim; ++f)
          if (nbr[(size_t)c * 2 * dim + f] < 0)
            for (int k = 0; k < (1 << dim); ++k)
              if (((k >> (f / 2)) & 1) == (f % 2))
                vert_on_boundary[(size_t)cell_verts[((size_t)c << dim) + k]] = 1;
    }

    // Build an nx*ny*nz structured grid on [lo,hi].  order==0: deal.II
    // hyper_cube + refine_global (hierarchical child order == Morton, child c
    // <-> bits x|y<<1|z<<2; requires nx==ny==nz==2^k); order==1:
    // subdivided_hyper_rectangle (lexicographic, x fastest).
    void
    build_structured(const int    dim_,
                     const int   *n,
                     const double *lo,
                     const double *hi,
                     const int     order)
    {
      dim        = dim_;
      int nn[3]  = {n[0], n[1], dim == 3 ? n[2] : 1};
      int nv[3]  = {nn[0] + 1, nn[1] + 1, dim == 3 ? nn[2] + 1 : 1};
      size_t ncell = (size_t)nn[0] * nn[1] * nn[2];
      size_t nvert = (size_t)nv[0] * nv[1] * nv[2];
      verts.resize(nvert * dim);
      vert_on_boundary.assign(nvert, 0);
      for (int k = 0; k < nv[2]; ++k)
        for (int j = 0; j < nv[1]; ++j)
          for (int i = 0; i < nv[0]; ++i)
            {
              const size_t v  = ((size_t)k * nv[1] + j) * nv[0] + i;
              const int    ijk[3] = {i, j, k};
              bool         bd = false;
              for (int d = 0; d < dim; ++d)
                {
                  // same arithmetic as repeated midpoint bisection on dyadic
                  // grids: lo + i*(hi-lo)/n
                  verts[v * dim + d] =
                    lo[d] + (hi[d] - lo[d]) * ((double)ijk[d] / (double)nn[d]);
                  if (ijk[d] == 0 || ijk[d] == nn[d])
                    bd = true;
                }
              vert_on_boundary[v] = bd;
            }
      int levels = 0;
      if (order == 0)
        {
          while ((1 << levels) < nn[0])
            ++levels;
          if ((1 << levels) != nn[0] || nn[1] != nn[0] ||
              (dim == 3 && nn[2] != nn[0]))
            throw std::runtime_error("Morton order needs n = 2^k in every direction");
        }
      auto cell_index = [&](int i, int j, int k) -> size_t {
        if (order == 1)
          return ((size_t)k * nn[1] + j) * nn[0] + i;
        size_t m = 0;
        for (int l = 0; l < levels; ++l)
          {
            m |= (size_t)((i >> l) & 1) << (dim * l + 0);
            m |= (size_t)((j >> l) & 1) << (dim * l + 1);
            if (dim == 3)
              m |= (size_t)((k >> l) & 1) << (dim * l + 2);
          }
        return m;
      };
      const int vpc = 1 << dim;
      cell_verts.resize(ncell * vpc);
      nbr.resize(ncell * 2 * dim);
      for (int k = 0; k < nn[2]; ++k)
        for (int j = 0; j < nn[1]; ++j)
          for (int i = 0; i < nn[0]; ++i)
            {
              const size_t c = cell_index(i, j, k);
              for (int v = 0; v < vpc; ++v)
                {
                  const int di = v & 1, dj = (v >> 1) & 1, dk = (v >> 2) & 1;
                  cell_verts[c * vpc + v] =
                    (int)(((size_t)(k + dk) * nv[1] + (j + dj)) * nv[0] + (i + di));
                }
              int *nb = &nbr[c * 2 * dim];
              nb[0]   = i > 0 ? (int)cell_index(i - 1, j, k) : -1;
              nb[1]   = i < nn[0] - 1 ? (int)cell_index(i + 1, j, k) : -1;
              nb[2]   = j > 0 ? (int)cell_index(i, j - 1, k) : -1;
              nb[3]   = j < nn[1] - 1 ? (int)cell_index(i, j + 1, k) : -1;
              if (dim == 3)
                {
                  nb[4] = k > 0 ? (int)cell_index(i, j, k - 1) : -1;
                  nb[5] = k < nn[2] - 1 ? (int)cell_index(i, j, k + 1) : -1;
                }
            }
    }

    // GridTools::distort_random(factor, tria, keep_boundary=true) analogue:
    // interior vertices move by U(-1,1)*factor*(shortest adjacent edge) per
    // coordinate.  (deal.II's own RNG stream is not reproducible here; tests
    // that use distorted grids are invariant-only)
    void
    distort_random(const double factor, const uint64_t seed)
    {
      const int           vpc = 1 << dim;
      std::vector<double> minlen(n_verts(), 1e300);
      for (int c = 0; c < n_cells(); ++c)
        for (int a = 0; a < vpc; ++a)
          for (int d = 0; d < dim; ++d)
            {
              const int b = a ^ (1 << d);
              if (b < a)
                continue;
              const int    va = cell_verts[(size_t)c * vpc + a];
              const int    vb = cell_verts[(size_t)c * vpc + b];
              double       l2 = 0;
              for (int e = 0; e < dim; ++e)
                {
                  const double t = verts[(size_t)va * dim + e] - verts[(size_t)vb * dim + e];
                  l2 += t * t;
                }
              const double l = std::sqrt(l2);
              minlen[va]     = std::min(minlen[va], l);
              minlen[vb]     = std::min(minlen[vb], l);
            }
      std::mt19937_64 rng(seed);
      for (int v = 0; v < n_verts(); ++v)
        for (int d = 0; d < dim; ++d)
          {
            // draw for every vertex so the stream does not depend on the mask
            const double u =
              2. * ((double)(rng() >> 11) * (1.0 / 9007199254740992.0)) - 1.;
            if (!vert_on_boundary[v])
              verts[(size_t)v * dim + d] += u * factor * minlen[v];
          }
    }
  };

  // ---------------------------------------------------------------------------
  // Q1 geometry of one cell / one cell face (deal.II FEValues / FEFaceValues
  // with MappingQ<dim>(1) and FE_Nothing, as set up in
  // source/agglomeration_handler.cc:224-235)
  // ---------------------------------------------------------------------------
  struct Quad1D
  {
    std::vector<double> x, w;
  };

  inline void
  q1_shape(const int dim, const double *xi, double *N, double *dN /*[v][dim]*/)
  {
    const int vpc = 1 << dim;
    for (int v = 0; v < vpc; ++v)
      {
        double f[3], df[3];
        for (int d = 0; d < dim; ++d)
          {
            const int bit = (v >> d) & 1;
            f[d]          = bit ? xi[d] : 1. - xi[d];
            df[d]         = bit ? 1. : -1.;
          }
        double val = 1.;
        for (int d = 0; d < dim; ++d)
          val *= f[d];
        N[v] = val;
        for (int d = 0; d < dim; ++d)
          {
            double g = df[d];
            for (int e = 0; e < dim; ++e)
              if (e != d)
                g *= f[e];
            dN[v * dim + d] = g;
          }
      }
  }

  // x = sum_v N_v x_v ; J[a][b] = d x_a / d xi_b
  inline void
  q1_map(const Grid &g, const int cell, const double *xi, double *x, double *J)
  {
    const int dim = g.dim, vpc = 1 << dim;
    double    N[8], dN[24];
    q1_shape(dim, xi, N, dN);
    for (int a = 0; a < dim; ++a)
      {
        x[a] = 0.;
        for (int b = 0; b < dim; ++b)
          J[a * dim + b] = 0.;
      }
    for (int v = 0; v < vpc; ++v)
      {
        const double *xv = g.vertex(cell, v);
        for (int a = 0; a < dim; ++a)
          {
            x[a] += N[v] * xv[a];
            for (int b = 0; b < dim; ++b)
              J[a * dim + b] += dN[v * dim + b] * xv[a];
          }
      }
  }

  inline double
  det(const int dim, const double *J)
  {
    if (dim == 2)
      return J[0] * J[3] - J[1] * J[2];
    return J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) +
           J[2] * (J[3] * J[7] - J[4] * J[6]);
  }

  // volume quadrature of one cell: QGauss<dim>(nq), x fastest
  inline void
  cell_quadrature(const Grid          &g,
                  const int            cell,
                  const Quad1D        &q,
                  std::vector<double> &pts,
                  std::vector<double> &jxw)
  {
    const int dim = g.dim, nq = (int)q.x.size();
    const int nz = dim == 3 ? nq : 1;
    for (int c = 0; c < nz; ++c)
      for (int b = 0; b < nq; ++b)
        for (int a = 0; a < nq; ++a)
          {
            double xi[3] = {q.x[a], q.x[b], dim == 3 ? q.x[c] : 0.};
            double w     = q.w[a] * q.w[b] * (dim == 3 ? q.w[c] : 1.);
            double x[3], J[9];
            q1_map(g, cell, xi, x, J);
            for (int d = 0; d < dim; ++d)
              pts.push_back(x[d]);
            jxw.push_back(w * det(dim, J));
          }
  }

  // face quadrature of (cell, face): QGauss<dim-1>(nq) projected to the face as
  // deal.II's QProjector does: 2-D: the single face coordinate runs along the
  // free axis; 3-D: faces 0/1 -> (y,z), faces 2/3 -> (z,x), faces 4/5 -> (x,y)
  // with the first face coordinate fastest.
  inline void
  face_quadrature(const Grid          &g,
                  const int            cell,
                  const int            face,
                  const Quad1D        &q,
                  std::vector<double> &pts,
                  std::vector<double> &jxw,
                  std::vector<double> &normals)
  {
    const int    dim = g.dim, nq = (int)q.x.size();
    const int    nd   = face / 2;               // normal direction
    const double side = (face % 2) ? 1. : 0.;
    int          t0, t1 = -1;                   // face coordinate axes
    if (dim == 2)
      t0 = 1 - nd;
    else
      {
        t0 = (nd + 1) % 3;
        t1 = (nd + 2) % 3;
      }
    const int n1 = dim == 3 ? nq : 1;
    for (int b = 0; b < n1; ++b)
      for (int a = 0; a < nq; ++a)
        {
          double xi[3] = {0., 0., 0.};
          xi[nd]       = side;
          xi[t0]       = q.x[a];
          double w     = q.w[a];
          if (dim == 3)
            {
              xi[t1] = q.x[b];
              w *= q.w[b];
            }
          double x[3], J[9];
          q1_map(g, cell, xi, x, J);
          // surface element and outward normal from the tangents
          double nrm[3] = {0, 0, 0};
          if (dim == 2)
            {
              const double tx = J[0 * 2 + t0], ty = J[1 * 2 + t0];
              // rotate tangent; orientation fixed below via the sign test
              nrm[0] = ty;
              nrm[1] = -tx;
            }
          else
            {
              const double a0 = J[0 * 3 + t0], a1 = J[1 * 3 + t0], a2 = J[2 * 3 + t0];
              const double b0 = J[0 * 3 + t1], b1 = J[1 * 3 + t1], b2 = J[2 * 3 + t1];
              nrm[0]          = a1 * b2 - a2 * b1;
              nrm[1]          = a2 * b0 - a0 * b2;
              nrm[2]          = a0 * b1 - a1 * b0;
            }
          double len = 0;
          for (int d = 0; d < dim; ++d)
            len += nrm[d] * nrm[d];
          len = std::sqrt(len);
          // outward: must have positive component along +-(dx/dxi_nd)
          double dotp = 0;
          for (int d = 0; d < dim; ++d)
            dotp += nrm[d] * J[d * dim + nd];
          const double sgn = ((dotp > 0) == (side > 0.5)) ? 1. : -1.;
          for (int d = 0; d < dim; ++d)
            {
              pts.push_back(x[d]);
              normals.push_back(sgn * nrm[d] / len);
            }
          jxw.push_back(w * len);
        }
  }

  // ---------------------------------------------------------------------------
  // AgglomerationHandler restatement
  // ---------------------------------------------------------------------------
  struct BBox
  {
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  };

  struct FEValuesTable
  {
    int                 n_dofs = 0, n_q = 0, dim = 2;
    std::vector<double> points;  // real space, n_q*dim
    std::vector<double> jxw;     // n_q
    std::vector<double> normals; // n_q*dim (faces only)
    std::vector<double> values;  // [i][q]
    std::vector<double> grads;   // [i][q][d]   (real-space gradients)
    double
    shape_value(int i, int q) const
    {
      return values[(size_t)i * n_q + q];
    }
    const double *
    shape_grad(int i, int q) const
    {
      return &grads[((size_t)i * n_q + q) * dim];
    }
  };

  struct Handler
  {
    const Grid *grid = nullptr;
    int         dim  = 2;

    // include/agglomeration_handler.h:688  (Vector<float> in the reference;
    // zero-initialised, -1 marks masters)
    std::vector<float> master_slave_relationships;
    std::vector<int>   master_of_cell; // master_slave_relationships_iterators
    std::vector<int>   master_cells_container;
    std::map<int, std::vector<int>> master2slaves;
    std::map<int, int>              master2polygon;
    std::vector<BBox>               bboxes;
    unsigned int                    n_agglomerations = 0;

    // polytope_cache
    std::map<std::pair<int, unsigned>, std::pair<bool, int>> cell_face_at_boundary;
    std::map<std::pair<int, int>, std::vector<std::pair<int, int>>> interface;
    std::set<std::pair<int, int>> visited_cell_and_faces;
    std::vector<unsigned int>     number_of_agglomerated_faces;

    FiniteElement    fe;
    Quad1D           cell_q, face_q;
    std::vector<int> dof_offset_of_cell; // first DoF on each master cell, -1 on slaves
    int              n_dofs = 0;

    explicit Handler(const Grid *g)
      : grid(g)
      , dim(g->dim)
    {
      master_slave_relationships.assign(g->n_cells(), 0.f);
      master_of_cell.assign(g->n_cells(), -1);
    }

    // source/agglomeration_handler.cc:45-104
    int
    define_agglomerate(const std::vector<int> &cells)
    {
      if (cells.empty())
        throw std::runtime_error("No cells to be agglomerated.");
      const int master = cells[0];
      master_cells_container.push_back(master);
      master_slave_relationships[master] = -1.f;
      std::vector<int> slaves;
      for (size_t i = 1; i < cells.size(); ++i)
        {
          slaves.push_back(cells[i]);
          master_slave_relationships[cells[i]] = (float)master;
          master_of_cell[cells[i]]             = master;
        }
      master_of_cell[master]  = master;
      master2slaves[master]   = slaves;
      master2polygon[master]  = (int)n_agglomerations;
      ++n_agglomerations;
      // create_bounding_box, :476-491 : min/max over all vertices of all cells
      BBox bb;
      for (int d = 0; d < dim; ++d)
        {
          bb.lo[d] = 1e300;
          bb.hi[d] = -1e300;
        }
      for (const int c : cells)
        for (int v = 0; v < (1 << dim); ++v)
          {
            const double *x = grid->vertex(c, v);
            for (int d = 0; d < dim; ++d)
              {
                bb.lo[d] = std::min(bb.lo[d], x[d]);
                bb.hi[d] = std::max(bb.hi[d], x[d]);
              }
          }
      bboxes.push_back(bb);
      return (int)n_agglomerations - 1;
    }

    bool
    is_master_cell(const int c) const
    {
      return master_slave_relationships[c] == -1.f;
    }
    bool
    is_slave_cell(const int c) const
    {
      return master_slave_relationships[c] >= 0.f && master_of_cell[c] != c;
    }
    int
    get_master_idx_of_cell(const int c) const
    {
      return master_of_cell[c];
    }
    bool
    are_cells_agglomerated(const int a, const int b) const
    {
      return get_master_idx_of_cell(a) == get_master_idx_of_cell(b);
    }
    // include/agglomeration_accessor.h:562-569 : slaves first, master last
    std::vector<int>
    get_agglomerate(const int master) const
    {
      std::vector<int> a = master2slaves.at(master);
      a.push_back(master);
      return a;
    }

    void
    initialize_fe_values(const int nq_cell, const int nq_face)
    {
      gauss_1d(nq_cell, cell_q.x, cell_q.w);
      gauss_1d(nq_face, face_q.x, face_q.w);
    }

    // source/agglomeration_handler.cc:326-379 (serial path)
    void
    distribute_agglomerated_dofs(const int fe_kind, const int degree)
    {
      fe.init(fe_kind, dim, degree);
      // initialize_hp_structure -> DoFHandler::distribute_dofs: consecutive
      // blocks on master cells in active-cell order, nothing on slaves
      dof_offset_of_cell.assign(grid->n_cells(), -1);
      n_dofs = 0;
      for (int c = 0; c < grid->n_cells(); ++c)
        if (master_of_cell[c] == c && is_master_cell(c))
          {
            dof_offset_of_cell[c] = n_dofs;
            n_dofs += fe.n_dofs;
          }
      setup_connectivity_of_agglomeration();
    }

    // :495-527
    void
    setup_connectivity_of_agglomeration()
    {
      number_of_agglomerated_faces.assign(master2polygon.size(), 0);
      cell_face_at_boundary.clear();
      interface.clear();
      visited_cell_and_faces.clear();
      for (const int m : master_cells_container)
        setup_master_neighbor_connectivity(m);
    }

    // :1253-1645, serial branches only (locally owned neighbours + boundary)
    void
    setup_master_neighbor_connectivity(const int master_cell)
    {
      const std::vector<int> agglomeration = get_agglomerate(master_cell);
      const int current_polytope_index     = master2polygon.at(master_cell);
      const int current_polytope_id        = master_cell; // CellId order == active index order
      std::set<unsigned int> visited_polygonal_neighbors;
      for (const int cell : agglomeration)
        {
          for (int f = 0; f < 2 * dim; ++f)
            {
              const int neighboring_cell = grid->neighbor(cell, f);
              if (neighboring_cell >= 0)
                {
                  if (!are_cells_agglomerated(cell, neighboring_cell))
                    {
                      const int master_of_neighbor = master_of_cell[neighboring_cell];
                      const int nof = grid->neighbor_of_neighbor(cell, f);
                      const int neighbor_polytope_index =
                        master2polygon.at(master_of_neighbor);
                      const int neighbor_polytope_id = master_of_neighbor;
                      if (visited_polygonal_neighbors.find(neighbor_polytope_index) ==
                          visited_polygonal_neighbors.end())
                        {
                          const unsigned int n_face =
                            number_of_agglomerated_faces[current_polytope_index];
                          cell_face_at_boundary[{current_polytope_index, n_face}] = {
                            false, master_of_neighbor};
                          ++number_of_agglomerated_faces[current_polytope_index];
                          visited_polygonal_neighbors.insert(neighbor_polytope_index);
                        }
                      if (visited_cell_and_faces.find({cell, f}) ==
                          visited_cell_and_faces.end())
                        {
                          interface[{current_polytope_id, neighbor_polytope_id}]
                            .emplace_back(cell, f);
                          visited_cell_and_faces.insert({cell, f});
                        }
                      if (visited_cell_and_faces.find({neighboring_cell, nof}) ==
                          visited_cell_and_faces.end())
                        {
                          interface[{neighbor_polytope_id, current_polytope_id}]
                            .emplace_back(neighboring_cell, nof);
                          visited_cell_and_faces.insert({neighboring_cell, nof});
                        }
                    }
                }
              else
                {
                  // boundary face of a boundary cell: all of them share ONE
                  // polytope face (sentinel UINT_MAX), :1575-1598
                  if (visited_polygonal_neighbors.find(invalid_uint) ==
                      visited_polygonal_neighbors.end())
                    {
                      const unsigned int n_face =
                        number_of_agglomerated_faces[current_polytope_index];
                      cell_face_at_boundary[{current_polytope_index, n_face}] = {true, -1};
                      ++number_of_agglomerated_faces[current_polytope_index];
                      visited_polygonal_neighbors.insert(invalid_uint);
                    }
                  if (visited_cell_and_faces.find({cell, f}) ==
                      visited_cell_and_faces.end())
                    {
                      interface[{current_polytope_id, current_polytope_id}].emplace_back(
                        cell, f);
                      visited_cell_and_faces.insert({cell, f});
                    }
                }
            }
        }
    }

    // ---- accessor-like queries; `poly` is the polytope index (define order)
    int
    n_polytopes() const
    {
      return (int)master_cells_container.size();
    }
    int
    master_cell(const int poly) const
    {
      return master_cells_container[poly];
    }
    unsigned int
    n_faces(const int poly) const
    {
      return number_of_agglomerated_faces[poly];
    }
    bool
    at_boundary(const int poly, const unsigned f) const
    {
      return cell_face_at_boundary.at({poly, f}).first;
    }
    // neighbouring polytope index, -1 on the boundary (accessor.h:335-422)
    int
    neighbor(const int poly, const unsigned f) const
    {
      const auto &e = cell_face_at_boundary.at({poly, f});
      if (e.first)
        return -1;
      return master2polygon.at(e.second);
    }
    // accessor.h:426-481 : linear scan over the neighbour's faces
    unsigned int
    neighbor_of_agglomerated_neighbor(const int poly, const unsigned f) const
    {
      if (at_boundary(poly, f))
        return invalid_uint;
      const int nb = neighbor(poly, f);
      for (unsigned int f_out = 0; f_out < n_faces(nb); ++f_out)
        if (!at_boundary(nb, f_out) && neighbor(nb, f_out) == poly)
          return f_out;
      return invalid_uint;
    }
    const std::vector<std::pair<int, int>> &
    common_face(const int poly, const unsigned f) const
    {
      const int id_in = master_cell(poly);
      const int nb    = neighbor(poly, f);
      const int id_out = nb < 0 ? id_in : master_cell(nb);
      return interface.at({id_in, id_out});
    }
    void
    get_dof_indices(const int poly, unsigned int *out) const
    {
      const int off = dof_offset_of_cell[master_cell(poly)];
      for (int i = 0; i < fe.n_dofs; ++i)
        out[i] = (unsigned int)(off + i);
    }
    // accessor.h:582-601 : diagonal of the bounding box
    double
    diameter(const int poly) const
    {
      const BBox &b = bboxes[poly];
      double      s = 0;
      for (int d = 0; d < dim; ++d)
        s += (b.hi[d] - b.lo[d]) * (b.hi[d] - b.lo[d]);
      return std::sqrt(s);
    }
    double
    volume(const int poly) const
    {
      const BBox &b = bboxes[poly];
      double      v = 1;
      for (int d = 0; d < dim; ++d)
        v *= (b.hi[d] - b.lo[d]);
      return v;
    }

    // evaluate the FE on polytope `poly`'s bounding box at real points
    // (MappingBox: xhat = (x-lo)/(hi-lo); grad = grad_hat * (1/h),
    //  source/mapping_box.cc:522-532)
    void
    fill_tables(const int poly, FEValuesTable &t) const
    {
      const BBox &bb = bboxes[poly];
      t.n_dofs       = fe.n_dofs;
      t.dim          = dim;
      t.n_q          = (int)t.jxw.size();
      t.values.resize((size_t)t.n_dofs * t.n_q);
      t.grads.resize((size_t)t.n_dofs * t.n_q * dim);
      double inv_h[3];
      for (int d = 0; d < dim; ++d)
        inv_h[d] = 1. / (bb.hi[d] - bb.lo[d]);
      std::vector<double> v(fe.n_dofs), g((size_t)fe.n_dofs * dim);
      for (int q = 0; q < t.n_q; ++q)
        {
          double xhat[3];
          for (int d = 0; d < dim; ++d) // BoundingBox::real_to_unit
            xhat[d] = (t.points[(size_t)q * dim + d] - bb.lo[d]) / (bb.hi[d] - bb.lo[d]);
          fe.evaluate(xhat, v.data(), g.data());
          for (int i = 0; i < fe.n_dofs; ++i)
            {
              t.values[(size_t)i * t.n_q + q] = v[i];
              for (int d = 0; d < dim; ++d)
                t.grads[((size_t)i * t.n_q + q) * dim + d] = g[(size_t)i * dim + d] * inv_h[d];
            }
        }
    }

    // source/agglomeration_handler.cc:729-767 + 622-707
    void
    reinit(const int poly, FEValuesTable &t) const
    {
      t.points.clear();
      t.jxw.clear();
      t.normals.clear();
      for (const int c : get_agglomerate(master_cell(poly)))
        cell_quadrature(*grid, c, cell_q, t.points, t.jxw);
      fill_tables(poly, t);
    }

    // reinit(polytope, f) / reinit_master, :785-801, :1103-1243
    void
    reinit_face(const int poly, const unsigned f, FEValuesTable &t) const
    {
      t.points.clear();
      t.jxw.clear();
      t.normals.clear();
      for (const auto &cf : common_face(poly, f))
        face_quadrature(*grid, cf.first, cf.second, face_q, t.points, t.jxw, t.normals);
      fill_tables(poly, t);
    }

    // reinit_interface, :805-834: each side integrates over ITS OWN list of
    // (cell, local face) pairs; the two lists are index-aligned by construction
    void
    reinit_interface(const int      poly_in,
                     const int      poly_out,
                     const unsigned local_in,
                     const unsigned local_out,
                     FEValuesTable &t_in,
                     FEValuesTable &t_out) const
    {
      reinit_face(poly_in, local_in, t_in);
      reinit_face(poly_out, local_out, t_out);
    }

    // :910-1022  scalar CSR pattern, ascending columns
    void
    sparsity_pattern(std::vector<int64_t> &rowptr, std::vector<int> &cols) const
    {
      const int                     n = fe.n_dofs;
      std::vector<std::set<int>>    brow(n_dofs / n); // block columns per block row
      std::vector<unsigned int>     cur(n), nb(n);
      for (int p = 0; p < n_polytopes(); ++p)
        {
          get_dof_indices(p, cur.data());
          brow[cur[0] / n].insert(cur[0] / n); // make_sparsity_pattern (volumetric)
          for (unsigned f = 0; f < n_faces(p); ++f)
            {
              const int q = neighbor(p, f);
              if (q >= 0)
                {
                  get_dof_indices(q, nb.data());
                  brow[cur[0] / n].insert(nb[0] / n);
                }
            }
        }
      rowptr.assign(n_dofs + 1, 0);
      cols.clear();
      for (int r = 0; r < n_dofs; ++r)
        {
          for (const int bc : brow[r / n])
            for (int j = 0; j < n; ++j)
              cols.push_back(bc * n + j);
          rowptr[r + 1] = (int64_t)cols.size();
        }
    }
  };

  // ---------------------------------------------------------------------------
  // SIP assembly (include/poly_utils.h:2000-2195) with the reference's
  // penalty-rule zoo as parameters
  // ---------------------------------------------------------------------------
  enum HRule
  {
    H_DIAMETER_OF_VISITOR = 0, // C / polytope->diameter()              poly_utils.h:2057
    H_MAX_INVERSE_DIAMETER = 1, // C * max(1/h_A, 1/h_B)                 poisson_sanity_check_01.cc:261-266
    H_CONSTANT            = 2, // C / h_const                            minimal_SIP_Poisson.cc:308
    H_NORMAL_EXTENT       = 3  // C*(1/h_n^A + 1/h_n^B), bdary 4*C/h_n   include/utils.h:861-866,906-909
  };
  enum VisitRule
  {
    VISIT_BY_ID    = 0, // polytope->id() < neigh->id()       poly_utils.h:2089
    VISIT_BY_INDEX = 1  // polytope->index() < neigh->index() examples/poisson.cc:841
  };

  struct AssembleParams
  {
    double penalty_constant = 0.;
    int    h_rule           = H_DIAMETER_OF_VISITOR;
    double h_const          = 1.;
    int    visit_rule       = VISIT_BY_ID;
    int    with_boundary    = 1;
    double stiffness_coeff  = 1.; // sigma: scales volume gradient AND all face terms
    double mass_coeff       = 0.; // f:     + f * phi_i phi_j
    int    n_threads        = 1;
    // bounded sampling for CPU-baseline timing: only polytopes p with
    // p % poly_stride == poly_offset do their (visiting) work
    int    poly_stride      = 1;
    int    poly_offset      = 0;
    // timing of a sample at sizes where the whole pattern would not fit the host: the local matrices are
    // computed exactly as always and summed into a per-thread n x n sink instead of the global matrix
    // (distribute_local_to_global left out: slightly in the CPU's favour)
    int    discard_scatter  = 0;
  };

  struct CSRMatrix
  {
    std::vector<int64_t> rowptr;
    std::vector<int>     cols;
    std::vector<double>  vals;
    inline void
    add(const unsigned r, const unsigned c, const double v)
    {
      const int *b = &cols[rowptr[r]], *e = &cols[rowptr[r + 1]];
      const int *p = std::lower_bound(b, e, (int)c);
      if (p == e || *p != (int)c)
        throw std::runtime_error("entry not in sparsity pattern");
      vals[rowptr[r] + (p - b)] += v;
    }
  };

  inline double
  dotn(const int dim, const double *g, const double *n)
  {
    double s = g[0] * n[0] + g[1] * n[1];
    if (dim == 3)
      s += g[2] * n[2];
    return s;
  }

  inline double
  normal_extent(const Handler &ah, const int poly, const double *n)
  {
    // |n . J^-1| for an axis-aligned box: 1/h along the dominant normal axis
    const BBox &b = ah.bboxes[poly];
    double      s = 0;
    for (int d = 0; d < ah.dim; ++d)
      s += std::fabs(n[d]) / (b.hi[d] - b.lo[d]);
    return s; // this is already the inverse extent
  }

  // penalty of a boundary point / an interior point seen from the visitor p (SURVEY 8c zoo)
  inline double
  boundary_penalty(const Handler &ah, const AssembleParams &prm, const int p, const double *nrm)
  {
    switch (prm.h_rule)
      {
        case H_CONSTANT:
          return prm.penalty_constant / prm.h_const;
        case H_NORMAL_EXTENT:
          return 4. * prm.penalty_constant * normal_extent(ah, p, nrm);
        default:
          return prm.penalty_constant / ah.diameter(p);
      }
  }
  inline double
  interior_penalty(const Handler &ah, const AssembleParams &prm, const int p, const int nbp, const double *nrm)
  {
    switch (prm.h_rule)
      {
        case H_MAX_INVERSE_DIAMETER:
          return prm.penalty_constant * std::max(1. / ah.diameter(p), 1. / ah.diameter(nbp));
        case H_CONSTANT:
          return prm.penalty_constant / prm.h_const;
        case H_NORMAL_EXTENT:
          return prm.penalty_constant * (normal_extent(ah, p, nrm) + normal_extent(ah, nbp, nrm));
        default:
          return prm.penalty_constant / ah.diameter(p);
      }
  }

  // volume term (include/poly_utils.h:2038-2052) plus the boundary faces of polytope p
  // (:2060-2085) into cell_matrix (n x n, zeroed here)
  inline void
  cell_and_boundary_matrix(const Handler        &ah,
                           const AssembleParams &prm,
                           const int             p,
                           FEValuesTable        &av,
                           FEValuesTable        &f0,
                           std::vector<double>  &cell_matrix)
  {
    const int    n   = ah.fe.n_dofs;
    const int    dim = ah.dim;
    const double sc  = prm.stiffness_coeff;
    std::fill(cell_matrix.begin(), cell_matrix.end(), 0.);
    ah.reinit(p, av);
    for (int q = 0; q < av.n_q; ++q)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          {
            const double *gi = av.shape_grad(i, q), *gj = av.shape_grad(j, q);
            double        gg = gi[0] * gj[0] + gi[1] * gj[1];
            if (dim == 3)
              gg += gi[2] * gj[2];
            double v = sc * gg;
            if (prm.mass_coeff != 0.)
              v += prm.mass_coeff * av.shape_value(i, q) * av.shape_value(j, q);
            cell_matrix[(size_t)i * n + j] += v * av.jxw[q];
          }
    if (!prm.with_boundary)
      return;
    const unsigned nf = ah.n_faces(p);
    for (unsigned f = 0; f < nf; ++f)
      {
        if (!ah.at_boundary(p, f))
          continue;
        ah.reinit_face(p, f, f0);
        for (int q = 0; q < f0.n_q; ++q)
          {
            const double *nrm = &f0.normals[(size_t)q * dim];
            const double  pen = boundary_penalty(ah, prm, p, nrm);
            for (int i = 0; i < n; ++i)
              for (int j = 0; j < n; ++j)
                cell_matrix[(size_t)i * n + j] +=
                  sc *
                  (-f0.shape_value(i, q) * dotn(dim, f0.shape_grad(j, q), nrm) -
                   dotn(dim, f0.shape_grad(i, q), nrm) * f0.shape_value(j, q) +
                   pen * f0.shape_value(i, q) * f0.shape_value(j, q)) *
                  f0.jxw[q];
          }
      }
  }

  // the four blocks of interface (p, neighbor(p, f)) seen from the visitor p:
  // reinit_interface + assemble_local_jumps_and_averages (include/poly_utils.h:1870-1926)
  inline void
  interface_matrices(const Handler        &ah,
                     const AssembleParams &prm,
                     const int             p,
                     const unsigned        f,
                     FEValuesTable        &f0,
                     FEValuesTable        &f1,
                     std::vector<double>  &M11,
                     std::vector<double>  &M12,
                     std::vector<double>  &M21,
                     std::vector<double>  &M22)
  {
    const int      n    = ah.fe.n_dofs;
    const int      dim  = ah.dim;
    const double   sc   = prm.stiffness_coeff;
    const int      nbp  = ah.neighbor(p, f);
    const unsigned nofn = ah.neighbor_of_agglomerated_neighbor(p, f);
    ah.reinit_interface(p, nbp, f, nofn, f0, f1);
    std::fill(M11.begin(), M11.end(), 0.);
    std::fill(M12.begin(), M12.end(), 0.);
    std::fill(M21.begin(), M21.end(), 0.);
    std::fill(M22.begin(), M22.end(), 0.);
    for (int q = 0; q < f0.n_q; ++q)
      {
        const double *nrm = &f0.normals[(size_t)q * dim];
        const double  pen = interior_penalty(ah, prm, p, nbp, nrm);
        // include/poly_utils.h:1884-1925
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j)
            {
              const double g0i = dotn(dim, f0.shape_grad(i, q), nrm);
              const double g0j = dotn(dim, f0.shape_grad(j, q), nrm);
              const double g1i = dotn(dim, f1.shape_grad(i, q), nrm);
              const double g1j = dotn(dim, f1.shape_grad(j, q), nrm);
              const double v0i = f0.shape_value(i, q), v0j = f0.shape_value(j, q);
              const double v1i = f1.shape_value(i, q), v1j = f1.shape_value(j, q);
              const size_t ij  = (size_t)i * n + j;
              M11[ij] += sc * (-0.5 * g0i * v0j - 0.5 * g0j * v0i + pen * v0i * v0j) * f0.jxw[q];
              M12[ij] += sc * (0.5 * g0i * v1j - 0.5 * g1j * v0i - pen * v0i * v1j) * f1.jxw[q];
              M21[ij] += sc * (-0.5 * g1i * v0j + 0.5 * g0j * v1i - pen * v1i * v0j) * f1.jxw[q];
              M22[ij] += sc * (0.5 * g1i * v1j + 0.5 * g1j * v1i + pen * v1i * v1j) * f1.jxw[q];
            }
      }
  }

  inline bool
  visits(const Handler &ah, const AssembleParams &prm, const int p, const int nbp)
  {
    return prm.visit_rule == VISIT_BY_ID ? ah.master_cell(p) < ah.master_cell(nbp) : p < nbp;
  }

  inline void
  assemble_dg_matrix(const Handler &ah, const AssembleParams &prm, CSRMatrix &A)
  {
    const int n  = ah.fe.n_dofs;
    const int np = ah.n_polytopes();
    const int nb = ah.n_dofs / n;
    if (prm.discard_scatter)
      {
        A.rowptr.assign(2, 0); // one row holding the checksum of everything computed
        A.rowptr[1] = 1;
        A.cols.assign(1, 0);
        A.vals.assign(1, 0.);
      }
    else
      {
        ah.sparsity_pattern(A.rowptr, A.cols);
        A.vals.assign(A.cols.size(), 0.);
      }
    std::atomic<int> sink_lock{0};
    // one spin lock per block row so the threaded variant stays race free
    std::unique_ptr<std::atomic_flag[]> locks(new std::atomic_flag[nb]);
    for (int i = 0; i < nb; ++i)
      locks[i].clear();
    auto scatter = [&](const std::vector<double> &M,
                       const unsigned int        *rows,
                       const unsigned int        *cols) {
      if (prm.discard_scatter)
        {
          double t = 0.;
          for (const double v : M)
            t += v;
          int expected = 0;
          while (!sink_lock.compare_exchange_weak(expected, 1, std::memory_order_acquire))
            expected = 0;
          A.vals[0] += t;
          sink_lock.store(0, std::memory_order_release);
          return;
        }
      const int br = rows[0] / n;
      while (locks[br].test_and_set(std::memory_order_acquire))
        ;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          A.add(rows[i], cols[j], M[(size_t)i * n + j]);
      locks[br].clear(std::memory_order_release);
    };

    std::atomic<int> next_poly{0};
    auto work = [&]() {
      std::vector<double>       cell_matrix((size_t)n * n), M11((size_t)n * n),
        M12((size_t)n * n), M21((size_t)n * n), M22((size_t)n * n);
      std::vector<unsigned int> ldi(n), ldin(n);
      FEValuesTable             av, f0, f1;
      // dynamic schedule: the visiting polytope does all the work of an interface,
      // so static ranges would be badly balanced
      for (int p = next_poly.fetch_add(1); p < np; p = next_poly.fetch_add(1))
        {
          if (p % prm.poly_stride != prm.poly_offset)
            continue;
          cell_and_boundary_matrix(ah, prm, p, av, f0, cell_matrix);
          ah.get_dof_indices(p, ldi.data());
          const unsigned nf = ah.n_faces(p);
          for (unsigned f = 0; f < nf; ++f)
            {
              if (ah.at_boundary(p, f))
                continue;
              const int nbp = ah.neighbor(p, f);
              if (!visits(ah, prm, p, nbp))
                continue;
              interface_matrices(ah, prm, p, f, f0, f1, M11, M12, M21, M22);
              ah.get_dof_indices(nbp, ldin.data());
              scatter(M11, ldi.data(), ldi.data());
              scatter(M12, ldi.data(), ldin.data());
              scatter(M21, ldin.data(), ldi.data());
              scatter(M22, ldin.data(), ldin.data());
            }
          scatter(cell_matrix, ldi.data(), ldi.data());
        }
    };

    const int nt = std::max(1, std::min(prm.n_threads, np));
    if (nt == 1)
      work();
    else
      {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t)
          th.emplace_back(work);
        for (auto &t : th)
          t.join();
      }
  }

  // The COMPLETE block rows of a few polytopes (parity checks at sizes where the whole matrix is
  // out of reach for a scalar CPU code): for polytope p its volume and boundary terms, and for
  // every interior face the interface evaluated from its VISITING side exactly as assemble_dg_matrix
  // does -- p keeps M11, M12 if it visits, M22, M21 if its neighbour does.  Row layout = the scalar
  // CSR rows of the reference pattern: block columns ascending; entry (i, k, j) of polytope s at
  // ptr[s] n^2 + i (nb_s n) + k n + j, nb_s = ptr[s+1] - ptr[s].
  struct BlockRows
  {
    std::vector<int64_t> ptr;
    std::vector<int>     bcol;
    std::vector<double>  vals;
  };
  inline void
  assemble_block_rows(const Handler &ah, const AssembleParams &prm, const int *polys, const int n_polys, BlockRows &R)
  {
    const int n = ah.fe.n_dofs;
    R.ptr.assign(n_polys + 1, 0);
    std::vector<std::vector<std::pair<int, unsigned>>> cols(n_polys); // (block column, face or ~0u for the diagonal)
    std::vector<unsigned int>                          ldi(n);
    for (int s = 0; s < n_polys; ++s)
      {
        const int p = polys[s];
        ah.get_dof_indices(p, ldi.data());
        cols[s].push_back({(int)(ldi[0] / n), ~0u});
        for (unsigned f = 0; f < ah.n_faces(p); ++f)
          if (!ah.at_boundary(p, f))
            {
              ah.get_dof_indices(ah.neighbor(p, f), ldi.data());
              cols[s].push_back({(int)(ldi[0] / n), f});
            }
        std::sort(cols[s].begin(), cols[s].end());
        R.ptr[s + 1] = R.ptr[s] + (int64_t)cols[s].size();
      }
    R.bcol.resize(R.ptr[n_polys]);
    R.vals.assign((size_t)R.ptr[n_polys] * n * n, 0.);
    for (int s = 0; s < n_polys; ++s)
      for (size_t k = 0; k < cols[s].size(); ++k)
        R.bcol[R.ptr[s] + k] = cols[s][k].first;
    std::atomic<int> next{0};
    auto             work = [&]() {
      std::vector<double> cell_matrix((size_t)n * n), M11((size_t)n * n), M12((size_t)n * n), M21((size_t)n * n),
        M22((size_t)n * n);
      FEValuesTable av, f0, f1;
      for (int s = next.fetch_add(1); s < n_polys; s = next.fetch_add(1))
        {
          const int     p   = polys[s];
          const int64_t nbk = R.ptr[s + 1] - R.ptr[s];
          double       *row = &R.vals[(size_t)R.ptr[s] * n * n];
          auto          add = [&](const size_t k, const std::vector<double> &M) {
            for (int i = 0; i < n; ++i)
              for (int j = 0; j < n; ++j)
                row[(size_t)i * nbk * n + k * n + j] += M[(size_t)i * n + j];
          };
          size_t kd = 0;
          while (cols[s][kd].second != ~0u)
            ++kd;
          cell_and_boundary_matrix(ah, prm, p, av, f0, cell_matrix);
          add(kd, cell_matrix);
          for (size_t k = 0; k < cols[s].size(); ++k)
            {
              const unsigned f = cols[s][k].second;
              if (f == ~0u)
                continue;
              const int nbp = ah.neighbor(p, f);
              if (visits(ah, prm, p, nbp))
                {
                  interface_matrices(ah, prm, p, f, f0, f1, M11, M12, M21, M22);
                  add(kd, M11);
                  add(k, M12);
                }
              else
                {
                  interface_matrices(ah, prm, nbp, ah.neighbor_of_agglomerated_neighbor(p, f), f0, f1, M11, M12, M21, M22);
                  add(kd, M22);
                  add(k, M21);
                }
            }
        }
    };
    const int nt = std::max(1, std::min(prm.n_threads, n_polys));
    if (nt == 1)
      work();
    else
      {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t)
          th.emplace_back(work);
        for (auto &t : th)
          t.join();
      }
  }

  // y = A x  (row 12: the matrix-based vmult on agglomerated levels)
  inline void
  spmv(const CSRMatrix &A, const double *x, double *y, const int n_threads = 1)
  {
    const int64_t nr   = (int64_t)A.rowptr.size() - 1;
    auto          work = [&](int64_t r0, int64_t r1) {
      for (int64_t r = r0; r < r1; ++r)
        {
          double s = 0;
          for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k)
            s += A.vals[k] * x[A.cols[k]];
          y[r] = s;
        }
    };
    const int nt = std::max(1, n_threads);
    if (nt == 1)
      work(0, nr);
    else
      {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t)
          th.emplace_back(work, nr * t / nt, nr * (t + 1) / nt);
        for (auto &t : th)
          t.join();
      }
  }
  // ---------------------------------------------------------------------------
  // Fine-mesh (non-agglomerated) SIP operator with the standard MAPPED FE_DGQ(p)
  // basis, phi_i(x) = phihat_i(F_K^{-1}(x)), F_K the Q1 map of the cell:
  //   y = (mass M + stiffness K_SIP) x
  // Restates the matrix-based assembly of /root/reference/examples/monodomain_DG3D.cc:1374-1622
  // (cell term :1408-1449; interior faces visited from the smaller active-cell index
  // :1459-1460, M11/M12/M21/M22 :1514-1575; penalty = pc (|J0^{-1}[nd0] . n| + |J1^{-1}[nd1] . n|)
  // at face quadrature point 0, :1470-1498, which the reference documents as the emulation of
  // the matrix-free sigmaF of include/utils.h:861-866) applied to a vector, plus the boundary
  // term of LaplaceOperatorDG::local_apply_boundary (include/utils.h:895-925):
  //   (4 pc |J^{-1}[nd] . n| u - dn u) v - u dn v,   pc = max(p,1)(p+1).
  // DoFs: n per cell in active-cell order.  Neighbouring cells are assumed to be in standard
  // orientation (opposite local face, aligned tangential axes); checked point by point.
  // PARITY UNPINNED: no reference test fixes numbers for this operator.
  // ---------------------------------------------------------------------------
  inline void
  mapped_fine_vmult(const Grid   &g,
                    const int     degree,
                    const int     nq,
                    const double  stiffness,
                    const double  mass,
                    const bool    with_volume,
                    const bool    with_boundary,
                    const bool    with_interior,
                    const double *x,
                    double       *y)
  {
    const int     dim = g.dim;
    FiniteElement fe;
    fe.init(FE_DGQ, dim, degree);
    const int n = fe.n_dofs;
    Quad1D    q;
    gauss_1d(nq, q.x, q.w);
    const double pc = std::max(degree, 1) * (degree + 1.0);
    std::fill(y, y + (size_t)g.n_cells() * n, 0.);

    // basis of `cell` at reference point xi: values, physical gradients, J^{-1}
    std::vector<double> gh(n * dim);
    auto eval = [&](const int cell, const double *xi, double *val, double *grad, double *Jinv, double *xr) {
      double J[9];
      q1_map(g, cell, xi, xr, J);
      const double dt = det(dim, J);
      if (dim == 2)
        {
          Jinv[0] = J[3] / dt;
          Jinv[1] = -J[1] / dt;
          Jinv[2] = -J[2] / dt;
          Jinv[3] = J[0] / dt;
        }
      else
        {
          Jinv[0] = (J[4] * J[8] - J[5] * J[7]) / dt;
          Jinv[1] = (J[2] * J[7] - J[1] * J[8]) / dt;
          Jinv[2] = (J[1] * J[5] - J[2] * J[4]) / dt;
          Jinv[3] = (J[5] * J[6] - J[3] * J[8]) / dt;
          Jinv[4] = (J[0] * J[8] - J[2] * J[6]) / dt;
          Jinv[5] = (J[2] * J[3] - J[0] * J[5]) / dt;
          Jinv[6] = (J[3] * J[7] - J[4] * J[6]) / dt;
          Jinv[7] = (J[1] * J[6] - J[0] * J[7]) / dt;
          Jinv[8] = (J[0] * J[4] - J[1] * J[3]) / dt;
        }
      fe.evaluate(xi, val, gh.data());
      // grad phi = J^{-T} gradhat phi :  (d/dx_a) = sum_b (dxi_b/dx_a) d/dxi_b = sum_b Jinv[b][a] ...
      for (int i = 0; i < n; ++i)
        for (int a = 0; a < dim; ++a)
          {
            double s = 0.;
            for (int b = 0; b < dim; ++b)
              s += Jinv[b * dim + a] * gh[i * dim + b];
            grad[i * dim + a] = s;
          }
      return dt;
    };

    std::vector<double> v0(n), g0(n * dim), v1(n), g1(n * dim);
    const int           nz = dim == 3 ? nq : 1;
    for (int c = 0; c < g.n_cells(); ++c)
      {
        const double *xc = x + (size_t)c * n;
        double       *yc = y + (size_t)c * n;
        if (with_volume)
          for (int kz = 0; kz < nz; ++kz)
            for (int ky = 0; ky < nq; ++ky)
              for (int kx = 0; kx < nq; ++kx)
                {
                  const double xi[3] = {q.x[kx], q.x[ky], dim == 3 ? q.x[kz] : 0.};
                  double       Ji[9], xr[3];
                  const double dt  = eval(c, xi, v0.data(), g0.data(), Ji, xr);
                  const double jxw = q.w[kx] * q.w[ky] * (dim == 3 ? q.w[kz] : 1.) * dt;
                  double       u = 0., gu[3] = {0, 0, 0};
                  for (int j = 0; j < n; ++j)
                    {
                      u += v0[j] * xc[j];
                      for (int a = 0; a < dim; ++a)
                        gu[a] += g0[j * dim + a] * xc[j];
                    }
                  for (int i = 0; i < n; ++i)
                    {
                      double s = 0.;
                      for (int a = 0; a < dim; ++a)
                        s += g0[i * dim + a] * gu[a];
                      yc[i] += (stiffness * s + mass * v0[i] * u) * jxw;
                    }
                }
        for (int f = 0; f < 2 * dim; ++f)
          {
            const int nbc = g.neighbor(c, f);
            if (nbc >= 0 && !(with_interior && c < nbc))
              continue;
            if (nbc < 0 && !with_boundary)
              continue;
            std::vector<double> pts, jxw, nrm;
            face_quadrature(g, c, f, q, pts, jxw, nrm);
            const int    nd = f / 2, nqf = (int)jxw.size();
            const double side = (f % 2) ? 1. : 0.;
            int          t0, t1 = -1;
            if (dim == 2)
              t0 = 1 - nd;
            else
              {
                t0 = (nd + 1) % 3;
                t1 = (nd + 2) % 3;
              }
            double        pen = 0.;
            const double *xn  = nbc >= 0 ? x + (size_t)nbc * n : nullptr;
            double       *yn  = nbc >= 0 ? y + (size_t)nbc * n : nullptr;
            for (int k = 0; k < nqf; ++k)
              {
                double xi[3] = {0, 0, 0}, xo[3] = {0, 0, 0};
                xi[nd]       = side;
                xi[t0]       = q.x[k % nq];
                if (dim == 3)
                  xi[t1] = q.x[k / nq];
                for (int a = 0; a < 3; ++a)
                  xo[a] = xi[a];
                xo[nd] = 1. - side;
                double        J0i[9], J1i[9], xr0[3], xr1[3];
                const double *nv = nrm.data() + (size_t)k * dim;
                eval(c, xi, v0.data(), g0.data(), J0i, xr0);
                if (nbc >= 0)
                  {
                    eval(nbc, xo, v1.data(), g1.data(), J1i, xr1);
                    for (int a = 0; a < dim; ++a)
                      if (std::fabs(xr0[a] - xr1[a]) > 1e-12 * (1. + std::fabs(xr0[a])))
                        throw std::runtime_error("mapped_fine_vmult: neighbouring cells are not in standard orientation");
                  }
                if (k == 0)
                  {
                    double h0 = 0., h1 = 0.;
                    for (int a = 0; a < dim; ++a)
                      {
                        h0 += J0i[nd * dim + a] * nv[a];
                        if (nbc >= 0)
                          h1 += J1i[nd * dim + a] * nv[a];
                      }
                    pen = nbc >= 0 ? pc * (std::fabs(h0) + std::fabs(h1)) : 4. * pc * std::fabs(h0);
                  }
                double u0 = 0., u1 = 0., dn0 = 0., dn1 = 0.;
                for (int j = 0; j < n; ++j)
                  {
                    u0 += v0[j] * xc[j];
                    for (int a = 0; a < dim; ++a)
                      dn0 += g0[j * dim + a] * nv[a] * xc[j];
                    if (nbc >= 0)
                      {
                        u1 += v1[j] * xn[j];
                        for (int a = 0; a < dim; ++a)
                          dn1 += g1[j * dim + a] * nv[a] * xn[j];
                      }
                  }
                const double w = stiffness * jxw[k];
                for (int i = 0; i < n; ++i)
                  {
                    double gn0 = 0., gn1 = 0.;
                    for (int a = 0; a < dim; ++a)
                      {
                        gn0 += g0[i * dim + a] * nv[a];
                        if (nbc >= 0)
                          gn1 += g1[i * dim + a] * nv[a];
                      }
                    if (nbc >= 0)
                      {
                        // rows of M11 x0 + M12 x1 and of M21 x0 + M22 x1
                        yc[i] += w * (-0.5 * gn0 * (u0 - u1) - 0.5 * (dn0 + dn1) * v0[i] + pen * v0[i] * (u0 - u1));
                        yn[i] += w * (-0.5 * gn1 * (u0 - u1) + 0.5 * (dn0 + dn1) * v1[i] - pen * v1[i] * (u0 - u1));
                      }
                    else
                      yc[i] += w * ((pen * u0 - dn0) * v0[i] - u0 * gn0);
                  }
              }
          }
      }
  }
} // namespace po
