// =============================================================================
// polydeal_b200_shim.hpp -- header-only C++ surface over the C ABI (polydeal_b200.h)
// with the reference's names, so that loops written against polyDEAL's
// AgglomerationHandler / FEValues / LinearOperatorMG compile against this library.
//
// What it mirrors (paths relative to /root/reference):
//   AgglomerationHandler<dim>   include/agglomeration_handler.h:171-575
//     define_agglomerate, initialize_fe_values, distribute_agglomerated_dofs, n_agglomerates, n_dofs,
//     create_agglomeration_sparsity_pattern, reinit(polytope), reinit(polytope, f), reinit_interface,
//     agglomerated_quadrature, polytope_iterators()
//   AgglomerationAccessor       include/agglomeration_accessor.h:41-299
//     index, n_faces, neighbor, at_boundary, neighbor_of_agglomerated_neighbor, diameter, volume,
//     get_dof_indices, n_background_cells, get_bounding_box
//   FEValues / FEValuesBase     the accessors the reference's loops use: shape_value, shape_grad, JxW,
//     quadrature_point, normal_vector, get_JxW_values, get_quadrature_points, get_normal_vectors,
//     n_quadrature_points, dofs_per_cell
//   BoundingBox / MappingBox    real_to_unit, unit_to_real (source/mapping_box.cc:923-972)
//   LinearOperatorMG            include/linear_operator_for_mg.h:295-322 (vmult / vmult_add / Tvmult /
//     Tvmult_add std::function members, n_rows / n_cols)
//   PolyUtils::assemble_dg_matrix  include/poly_utils.h:2000-2195
//
// No deal.II types: points / tensors are a small Tensor1<dim> struct, vectors are std::vector<double> or raw
// pointers (host).  Like the reference, the FEValues reference returned by reinit* is invalidated by the next
// reinit* call on the same handler (include/agglomeration_handler.h:841-851) -- except that reinit_interface
// returns a pair that stays valid together.  Errors: polydeal_b200::Error carrying the C status code
// (the reference throws deal.II exceptions).  Compiles with any C++17 compiler; links -lpolydeal_b200.
// =============================================================================
#ifndef POLYDEAL_B200_SHIM_HPP
#define POLYDEAL_B200_SHIM_HPP

#include "polydeal_b200.h"

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace polydeal_b200
{
  struct Error : std::runtime_error
  {
    int code;
    Error(const int c, const std::string &m)
      : std::runtime_error(m)
      , code(c)
    {}
  };
  inline int
  check(const int rc)
  {
    if (rc < 0)
      throw Error(rc, pd_last_error());
    return rc;
  }

  // Tensor<1, dim> / Point<dim> as far as the loops need them: components and the scalar product a * b
  template <int dim>
  struct Tensor1
  {
    double v[dim];
    double &
    operator[](const unsigned int d)
    {
      return v[d];
    }
    const double &
    operator[](const unsigned int d) const
    {
      return v[d];
    }
    double *
    data()
    {
      return v;
    }
    const double *
    data() const
    {
      return v;
    }
    friend double
    operator*(const Tensor1 &a, const Tensor1 &b)
    {
      double s = 0;
      for (int d = 0; d < dim; ++d)
        s += a.v[d] * b.v[d];
      return s;
    }
  };
  template <int dim>
  using Point = Tensor1<dim>;

  // The tables of one reinit*: FEValues<dim> / FEValuesBase<dim> as far as the reference's loops read them.
  template <int dim>
  class FEValues
  {
  public:
    unsigned int n_quadrature_points = 0, dofs_per_cell = 0;
    double
    shape_value(const unsigned int i, const unsigned int q) const
    {
      return values[(std::size_t)i * n_quadrature_points + q];
    }
    Tensor1<dim>
    shape_grad(const unsigned int i, const unsigned int q) const
    {
      Tensor1<dim>  g;
      const double *p = &grads[((std::size_t)i * n_quadrature_points + q) * dim];
      for (int d = 0; d < dim; ++d)
        g[d] = p[d];
      return g;
    }
    double
    JxW(const unsigned int q) const
    {
      return jxw[q];
    }
    Point<dim>
    quadrature_point(const unsigned int q) const
    {
      Point<dim> x;
      for (int d = 0; d < dim; ++d)
        x[d] = points[(std::size_t)q * dim + d];
      return x;
    }
    Tensor1<dim>
    normal_vector(const unsigned int q) const
    {
      Tensor1<dim> x;
      for (int d = 0; d < dim; ++d)
        x[d] = normals[(std::size_t)q * dim + d];
      return x;
    }
    const std::vector<double> &
    get_JxW_values() const
    {
      return jxw;
    }
    std::vector<Point<dim>>
    get_quadrature_points() const
    {
      std::vector<Point<dim>> r(n_quadrature_points);
      for (unsigned int q = 0; q < n_quadrature_points; ++q)
        r[q] = quadrature_point(q);
      return r;
    }
    std::vector<Tensor1<dim>>
    get_normal_vectors() const
    {
      std::vector<Tensor1<dim>> r(n_quadrature_points);
      for (unsigned int q = 0; q < n_quadrature_points; ++q)
        r[q] = normal_vector(q);
      return r;
    }
    std::vector<unsigned int>
    quadrature_point_indices() const
    {
      std::vector<unsigned int> r(n_quadrature_points);
      for (unsigned int q = 0; q < n_quadrature_points; ++q)
        r[q] = q;
      return r;
    }
    // raw tables (layout documented in polydeal_b200.h)
    std::vector<double> values, grads, jxw, points, normals;

    void
    resize(const unsigned int n, const unsigned int Q, const bool face)
    {
      dofs_per_cell       = n;
      n_quadrature_points = Q;
      values.resize((std::size_t)n * Q);
      grads.resize((std::size_t)n * Q * dim);
      jxw.resize(Q);
      points.resize((std::size_t)Q * dim);
      normals.resize(face ? (std::size_t)Q * dim : 0);
    }
  };
  template <int dim>
  using FEValuesBase = FEValues<dim>;

  template <int dim>
  class BoundingBox
  {
  public:
    Point<dim> lo, hi;
    Point<dim>
    real_to_unit(const Point<dim> &x) const
    {
      Point<dim> u;
      for (int d = 0; d < dim; ++d)
        u[d] = (x[d] - lo[d]) / (hi[d] - lo[d]);
      return u;
    }
    Point<dim>
    unit_to_real(const Point<dim> &u) const
    {
      Point<dim> x;
      for (int d = 0; d < dim; ++d)
        x[d] = lo[d] + u[d] * (hi[d] - lo[d]);
      return x;
    }
    double
    side_length(const unsigned int d) const
    {
      return hi[d] - lo[d];
    }
    double
    volume() const
    {
      double v = 1;
      for (int d = 0; d < dim; ++d)
        v *= hi[d] - lo[d];
      return v;
    }
    std::pair<Point<dim>, Point<dim>>
    get_boundary_points() const
    {
      return {lo, hi};
    }
  };

  // Quadrature<dim> as agglomerated_quadrature returns it: unit points of the bounding box, physical JxW
  template <int dim>
  struct Quadrature
  {
    std::vector<Point<dim>> points;
    std::vector<double>     weights;
    std::size_t
    size() const
    {
      return weights.size();
    }
    const Point<dim> &
    point(const std::size_t q) const
    {
      return points[q];
    }
    double
    weight(const std::size_t q) const
    {
      return weights[q];
    }
  };

  template <int dim>
  class AgglomerationHandler;

  // AgglomerationAccessor / AgglomerationIterator: a polytope by its index()
  template <int dim>
  class Polytope
  {
  public:
    Polytope(const AgglomerationHandler<dim> *ah, const int32_t p)
      : ah(ah)
      , p(p)
    {}
    const Polytope *
    operator->() const
    {
      return this;
    }
    int32_t
    index() const
    {
      return p;
    }
    // CellId of the master cell orders like its active-cell index on the meshes this library sees
    int32_t
    id() const
    {
      return check(pdh_master_cell(ah->host(), p));
    }
    unsigned int
    n_faces() const
    {
      const uint32_t r = pdh_n_faces(ah->host(), p);
      if (r == PD_INVALID_UINT)
        throw Error(PD_ERR_INVALID, pd_last_error());
      return r;
    }
    bool
    at_boundary(const unsigned int f) const
    {
      return check(pdh_at_boundary(ah->host(), p, f)) != 0;
    }
    Polytope
    neighbor(const unsigned int f) const
    {
      const int32_t q = pdh_neighbor(ah->host(), p, f);
      if (q < 0)
        throw Error(PD_ERR_INVALID, "neighbor(): boundary face or face index out of range");
      return Polytope(ah, q);
    }
    unsigned int
    neighbor_of_agglomerated_neighbor(const unsigned int f) const
    {
      return pdh_neighbor_of_agglomerated_neighbor(ah->host(), p, f);
    }
    unsigned int
    n_background_cells() const
    {
      return (unsigned int)check(pdh_n_background_cells(ah->host(), p));
    }
    std::vector<int32_t>
    get_agglomerate() const
    {
      std::vector<int32_t> c(n_background_cells());
      check(pdh_get_agglomerate(ah->host(), p, c.data()));
      return c;
    }
    double
    diameter() const
    {
      return pdh_diameter(ah->host(), p);
    }
    double
    volume() const
    {
      return pdh_volume(ah->host(), p);
    }
    BoundingBox<dim>
    get_bounding_box() const
    {
      BoundingBox<dim> b;
      check(pdh_bounding_box(ah->host(), p, b.lo.data(), b.hi.data()));
      return b;
    }
    void
    get_dof_indices(std::vector<unsigned int> &dofs) const
    {
      dofs.resize(ah->n_dofs_per_cell());
      check(pdh_get_dof_indices(ah->host(), p, dofs.data()));
    }
    bool
    is_locally_owned() const
    {
      return true;
    }

  private:
    const AgglomerationHandler<dim> *ah;
    int32_t                          p;
  };

  template <int dim>
  class AgglomerationHandler
  {
  public:
    // GridGenerator::hyper_cube(tria, a, b) + tria.refine_global(n_refine)
    AgglomerationHandler(const double a, const double b, const unsigned int n_refine)
    {
      int32_t n[3];
      double  lo[3], hi[3];
      for (int d = 0; d < 3; ++d)
        {
          n[d]  = 1 << n_refine;
          lo[d] = a;
          hi[d] = b;
        }
      check(pdh_grid_create_structured(dim, n, lo, hi, 0, &grid));
      check(pdh_handler_create(grid, &ah));
    }
    // any quad / hex mesh in deal.II conventions (what GridIn would deliver)
    AgglomerationHandler(const std::vector<double> &verts, const std::vector<int32_t> &cell_verts,
                         const std::vector<int32_t> &neighbours)
    {
      check(pdh_grid_create(dim, (int64_t)verts.size() / dim, verts.data(), (int64_t)cell_verts.size() >> dim,
                            cell_verts.data(), neighbours.data(), &grid));
      check(pdh_handler_create(grid, &ah));
    }
    AgglomerationHandler(const AgglomerationHandler &)            = delete;
    AgglomerationHandler &operator=(const AgglomerationHandler &) = delete;
    ~AgglomerationHandler()
    {
      if (dev)
        pd_destroy(dev);
      if (ah)
        pdh_handler_destroy(ah);
      if (grid)
        pdh_grid_destroy(grid);
    }

    Polytope<dim>
    define_agglomerate(const std::vector<int32_t> &cells)
    {
      return Polytope<dim>(this, check(pdh_define_agglomerate(ah, cells.data(), (int32_t)cells.size())));
    }
    void
    initialize_fe_values(const unsigned int n_q_points_1d_cell, const unsigned int n_q_points_1d_face)
    {
      check(pdh_initialize_fe_values(ah, (int32_t)n_q_points_1d_cell, (int32_t)n_q_points_1d_face));
    }
    // fe_kind: PD_FE_DGQ or PD_FE_AGGLODGP
    void
    distribute_agglomerated_dofs(const int fe_kind, const unsigned int degree)
    {
      check(pdh_distribute_agglomerated_dofs(ah, fe_kind, (int32_t)degree));
    }
    unsigned int
    n_agglomerates() const
    {
      return (unsigned int)pdh_n_polytopes(ah);
    }
    unsigned int
    n_dofs() const
    {
      return (unsigned int)pdh_n_dofs(ah);
    }
    unsigned int
    n_dofs_per_cell() const
    {
      return (unsigned int)pdh_n_dofs_per_cell(ah);
    }
    std::vector<Polytope<dim>>
    polytope_iterators() const
    {
      std::vector<Polytope<dim>> r;
      for (unsigned int p = 0; p < n_agglomerates(); ++p)
        r.emplace_back(this, (int32_t)p);
      return r;
    }
    // scalar CSR pattern, ascending columns (DynamicSparsityPattern order)
    void
    create_agglomeration_sparsity_pattern(std::vector<int64_t> &rowptr, std::vector<int32_t> &cols) const
    {
      const int64_t nnz = pdh_sparsity_nnz(ah);
      if (nnz < 0)
        throw Error(PD_ERR_STATE, pd_last_error());
      rowptr.resize((std::size_t)n_dofs() + 1);
      cols.resize((std::size_t)nnz);
      check(pdh_create_agglomeration_sparsity_pattern(ah, rowptr.data(), cols.data()));
    }

    // ---- device side: flatten + pd_create (lazily, on the first call that needs the GPU) ----
    void
    set_penalty(const pdh_flatten_params &prm)
    {
      params = prm;
      if (dev)
        {
          pd_destroy(dev);
          dev = nullptr;
        }
    }
    pd_handle *
    device() const
    {
      if (!dev)
        check(pdh_create_device(ah, &params, &dev)); // pdh_flatten + pd_create
      return dev;
    }
    pdh_handler *
    host() const
    {
      return ah;
    }

    // reinit(polytope): FEValues of the element on the bounding box at the agglomerated quadrature
    const FEValues<dim> &
    reinit(const Polytope<dim> &polytope) const
    {
      pd_handle    *h = device();
      const int64_t Q = pd_reinit_n_points(h, polytope.index());
      check((int)(Q < 0 ? Q : 0));
      scratch.resize(n_dofs_per_cell(), (unsigned int)Q, false);
      check(pd_reinit_polytope(h, polytope.index(), scratch.values.data(), scratch.grads.data(), scratch.jxw.data(),
                               scratch.points.data()));
      return scratch;
    }
    // reinit(polytope, f): FEValuesBase on face f of the polytope (its sub-faces, outward normals)
    const FEValuesBase<dim> &
    reinit(const Polytope<dim> &polytope, const unsigned int face_index) const
    {
      fill_face(polytope.index(), face_index, scratch_face);
      return scratch_face;
    }
    // reinit_interface(polytope_in, neigh_polytope, local_in, local_neigh): both sides at aligned points
    std::pair<const FEValuesBase<dim> &, const FEValuesBase<dim> &>
    reinit_interface(const Polytope<dim> &polytope_in, const Polytope<dim> &neigh_polytope, const unsigned int local_in,
                     const unsigned int local_neigh) const
    {
      fill_face(polytope_in.index(), local_in, scratch_face);
      fill_face(neigh_polytope.index(), local_neigh, scratch_neigh);
      return {scratch_face, scratch_neigh};
    }
    Quadrature<dim>
    agglomerated_quadrature(const Polytope<dim> &polytope) const
    {
      pd_handle    *h = device();
      const int64_t Q = pd_reinit_n_points(h, polytope.index());
      check((int)(Q < 0 ? Q : 0));
      Quadrature<dim>     quad;
      std::vector<double> u((std::size_t)Q * dim);
      quad.weights.resize((std::size_t)Q);
      check(pd_agglomerated_quadrature(h, polytope.index(), u.data(), quad.weights.data(), nullptr));
      quad.points.resize((std::size_t)Q);
      for (int64_t q = 0; q < Q; ++q)
        for (int d = 0; d < dim; ++d)
          quad.points[(std::size_t)q][d] = u[(std::size_t)q * dim + d];
      return quad;
    }

  private:
    void
    fill_face(const int32_t p, const unsigned int f, FEValues<dim> &out) const
    {
      pd_handle *h     = device();
      int32_t    iface = -1, side = -1;
      check(pdh_face_work_item(ah, p, f, &iface, &side));
      const int64_t Q = pd_reinit_iface_n_points(h, iface);
      check((int)(Q < 0 ? Q : 0));
      out.resize(n_dofs_per_cell(), (unsigned int)Q, true);
      check(pd_reinit_face(h, iface, side, out.values.data(), out.grads.data(), out.jxw.data(), out.points.data(),
                           out.normals.data()));
    }

    pdh_grid            *grid = nullptr;
    pdh_handler         *ah   = nullptr;
    mutable pd_handle   *dev  = nullptr;
    pdh_flatten_params   params{-1.0, PD_H_DIAMETER_OF_VISITOR, 1.0, PD_VISIT_BY_ID};
    mutable FEValues<dim> scratch, scratch_face, scratch_neigh; // agglomerated_scratch / _isv / _isv_neigh
  };

  // PolyUtils::assemble_dg_matrix(system_matrix, fe_dg, ah): assembles on the device; the values of the scalar CSR
  // matrix (pattern of create_agglomeration_sparsity_pattern) come back in `values`.
  template <int dim>
  inline void
  assemble_dg_matrix(std::vector<double> &values, const AgglomerationHandler<dim> &ah, const uint32_t flags = PD_ASSEMBLE_ALL,
                     const pd_coefficients coefficients = {1.0, 0.0})
  {
    pd_handle *h = ah.device();
    check(pd_assemble(h, flags, &coefficients));
    values.resize((std::size_t)pd_nnz(h));
    check(pd_matrix_values_to_host(h, values.data()));
  }

  // LinearOperatorMG<Range, Domain> (include/linear_operator_for_mg.h:295-322): std::function hooks a solver calls.
  // Vectors are host std::vector<double> here (pd_vmult_host); device pointers go through pd_vmult directly.
  struct LinearOperatorMG
  {
    std::function<void(std::vector<double> &, const std::vector<double> &)> vmult, vmult_add, Tvmult, Tvmult_add;
    std::function<void(std::vector<double> &, bool)>                          reinit_range_vector, reinit_domain_vector;
    std::size_t                                                               n_rows = 0, n_cols = 0;
    std::size_t
    m() const
    {
      return n_rows;
    }
    std::size_t
    n() const
    {
      return n_cols;
    }
  };
  // linear_operator_mg(matrix): the operator of an assembled (BLOCK_CSR) or matrix-free handle
  inline LinearOperatorMG
  linear_operator_mg(pd_handle *h, const int mode = PD_VMULT_BLOCK_CSR)
  {
    LinearOperatorMG op;
    op.n_rows = (std::size_t)pd_n_dofs(h);
    op.n_cols = (std::size_t)pd_n_source_dofs(h);
    op.vmult  = [h, mode](std::vector<double> &dst, const std::vector<double> &src) {
      dst.resize((std::size_t)pd_n_dofs(h));
      check(pd_vmult_host(h, mode, src.data(), dst.data()));
    };
    op.vmult_add = [h, mode](std::vector<double> &dst, const std::vector<double> &src) {
      std::vector<double> t((std::size_t)pd_n_dofs(h));
      check(pd_vmult_host(h, mode, src.data(), t.data()));
      for (std::size_t i = 0; i < t.size(); ++i)
        dst[i] += t[i];
    };
    op.Tvmult              = op.vmult; // the SIP operator is symmetric (include/utils.h:431-445 does the same)
    op.Tvmult_add          = op.vmult_add;
    op.reinit_range_vector = [h](std::vector<double> &v, bool) { v.assign((std::size_t)pd_n_dofs(h), 0.); };
    op.reinit_domain_vector = [h](std::vector<double> &v, bool) { v.assign((std::size_t)pd_n_source_dofs(h), 0.); };
    return op;
  }
} // namespace polydeal_b200

#endif // POLYDEAL_B200_SHIM_HPP
