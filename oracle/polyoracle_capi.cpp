// =============================================================================
//  oracle/polyoracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE
//  extern "C" surface of the CPU oracle (see polyoracle.hpp) for ctypes.
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
//  --impl reference legs may load this library.
// =============================================================================
#include "polyoracle.hpp"

#include <chrono>
#include <cstdio>

using namespace po;

namespace
{
  thread_local std::string g_err;
  template <class F>
  int
  guard(F &&f)
  {
    try
      {
        f();
        return 0;
      }
    catch (const std::exception &e)
      {
        g_err = e.what();
        return -1;
      }
  }
  struct Matrix
  {
    CSRMatrix A;
  };
} // namespace

extern "C"
{
  const char *
  po_last_error()
  {
    return g_err.c_str();
  }

  // ---- 1-D rules / FE -------------------------------------------------------
  void
  po_gauss_1d(int n, double *x, double *w)
  {
    std::vector<double> p, q;
    gauss_1d(n, p, q);
    std::copy(p.begin(), p.end(), x);
    std::copy(q.begin(), q.end(), w);
  }
  void
  po_gauss_lobatto_nodes(int n, double *x)
  {
    auto p = gauss_lobatto_nodes(n);
    std::copy(p.begin(), p.end(), x);
  }
  int
  po_fe_n_dofs(int kind, int dim, int degree)
  {
    FiniteElement fe;
    fe.init(kind, dim, degree);
    return fe.n_dofs;
  }
  // values[n], grads[n*dim] on the unit box
  void
  po_fe_evaluate(int kind, int dim, int degree, const double *xhat, double *values, double *grads)
  {
    FiniteElement fe;
    fe.init(kind, dim, degree);
    fe.evaluate(xhat, values, grads);
  }

  // ---- grid -----------------------------------------------------------------
  void *
  po_grid_structured(int dim, const int *n, const double *lo, const double *hi, int order)
  {
    Grid *g = new Grid;
    if (guard([&] { g->build_structured(dim, n, lo, hi, order); }))
      {
        delete g;
        return nullptr;
      }
    return g;
  }
  void *
  po_grid_from_arrays(int dim, int n_verts, const double *verts, int n_cells, const int *cell_verts, const int *nbr)
  {
    Grid *g = new Grid;
    if (guard([&] { g->build_from_arrays(dim, n_verts, verts, n_cells, cell_verts, nbr); }))
      {
        delete g;
        return nullptr;
      }
    return g;
  }
  void
  po_grid_free(void *g)
  {
    delete static_cast<Grid *>(g);
  }
  void
  po_grid_distort_random(void *g, double factor, uint64_t seed)
  {
    static_cast<Grid *>(g)->distort_random(factor, seed);
  }
  int
  po_grid_dim(void *g)
  {
    return static_cast<Grid *>(g)->dim;
  }
  int
  po_grid_n_cells(void *g)
  {
    return static_cast<Grid *>(g)->n_cells();
  }
  int
  po_grid_n_verts(void *g)
  {
    return static_cast<Grid *>(g)->n_verts();
  }
  void
  po_grid_cell_vertices(void *g_, int cell, double *out)
  {
    Grid *g = static_cast<Grid *>(g_);
    for (int v = 0; v < (1 << g->dim); ++v)
      for (int d = 0; d < g->dim; ++d)
        out[v * g->dim + d] = g->vertex(cell, v)[d];
  }
  int
  po_grid_neighbor(void *g, int cell, int f)
  {
    return static_cast<Grid *>(g)->neighbor(cell, f);
  }
  // raw arrays (to feed the SAME mesh to the product under test)
  void
  po_grid_copy_arrays(void *g_, double *verts, int *cell_verts, int *nbr)
  {
    Grid *g = static_cast<Grid *>(g_);
    std::copy(g->verts.begin(), g->verts.end(), verts);
    std::copy(g->cell_verts.begin(), g->cell_verts.end(), cell_verts);
    std::copy(g->nbr.begin(), g->nbr.end(), nbr);
  }

  // ---- handler --------------------------------------------------------------
  void *
  po_ah_create(void *g)
  {
    return new Handler(static_cast<Grid *>(g));
  }
  void
  po_ah_free(void *ah)
  {
    delete static_cast<Handler *>(ah);
  }
  int
  po_ah_define_agglomerate(void *ah, const int *cells, int n)
  {
    int r = -1;
    if (guard([&] {
          r = static_cast<Handler *>(ah)->define_agglomerate(std::vector<int>(cells, cells + n));
        }))
      return -1;
    return r;
  }
  void
  po_ah_initialize_fe_values(void *ah, int nq_cell, int nq_face)
  {
    static_cast<Handler *>(ah)->initialize_fe_values(nq_cell, nq_face);
  }
  int
  po_ah_distribute_agglomerated_dofs(void *ah, int fe_kind, int degree)
  {
    return guard([&] { static_cast<Handler *>(ah)->distribute_agglomerated_dofs(fe_kind, degree); });
  }
  int
  po_ah_n_polytopes(void *ah)
  {
    return static_cast<Handler *>(ah)->n_polytopes();
  }
  int
  po_ah_n_dofs(void *ah)
  {
    return static_cast<Handler *>(ah)->n_dofs;
  }
  int
  po_ah_n_dofs_per_cell(void *ah)
  {
    return static_cast<Handler *>(ah)->fe.n_dofs;
  }
  int
  po_ah_master_cell(void *ah, int poly)
  {
    return static_cast<Handler *>(ah)->master_cell(poly);
  }
  double
  po_ah_master_slave_value(void *ah, int cell)
  {
    return static_cast<Handler *>(ah)->master_slave_relationships[cell];
  }
  int
  po_ah_n_subcells(void *ah, int poly)
  {
    Handler *h = static_cast<Handler *>(ah);
    return (int)h->master2slaves.at(h->master_cell(poly)).size() + 1;
  }
  void
  po_ah_get_agglomerate(void *ah, int poly, int *cells)
  {
    Handler *h = static_cast<Handler *>(ah);
    auto     a = h->get_agglomerate(h->master_cell(poly));
    std::copy(a.begin(), a.end(), cells);
  }
  void
  po_ah_bbox(void *ah, int poly, double *lo, double *hi)
  {
    Handler *h = static_cast<Handler *>(ah);
    for (int d = 0; d < h->dim; ++d)
      {
        lo[d] = h->bboxes[poly].lo[d];
        hi[d] = h->bboxes[poly].hi[d];
      }
  }
  unsigned int
  po_ah_n_faces(void *ah, int poly)
  {
    return static_cast<Handler *>(ah)->n_faces(poly);
  }
  int
  po_ah_at_boundary(void *ah, int poly, unsigned f)
  {
    return static_cast<Handler *>(ah)->at_boundary(poly, f) ? 1 : 0;
  }
  int
  po_ah_neighbor(void *ah, int poly, unsigned f)
  {
    return static_cast<Handler *>(ah)->neighbor(poly, f);
  }
  unsigned int
  po_ah_neighbor_of_agglomerated_neighbor(void *ah, int poly, unsigned f)
  {
    return static_cast<Handler *>(ah)->neighbor_of_agglomerated_neighbor(poly, f);
  }
  // interface.at({id(poly), id(neighbor(f))}) (boundary: {id,id}); returns count
  int
  po_ah_interface(void *ah, int poly, unsigned f, int *cells, int *faces, int cap)
  {
    const auto &cf = static_cast<Handler *>(ah)->common_face(poly, f);
    for (int i = 0; i < (int)cf.size() && i < cap; ++i)
      {
        cells[i] = cf[i].first;
        faces[i] = cf[i].second;
      }
    return (int)cf.size();
  }
  void
  po_ah_get_dof_indices(void *ah, int poly, unsigned int *out)
  {
    static_cast<Handler *>(ah)->get_dof_indices(poly, out);
  }
  double
  po_ah_diameter(void *ah, int poly)
  {
    return static_cast<Handler *>(ah)->diameter(poly);
  }
  double
  po_ah_volume(void *ah, int poly)
  {
    return static_cast<Handler *>(ah)->volume(poly);
  }

  // ---- reinit tables --------------------------------------------------------
  // kind: 0 = reinit(polytope), 1 = reinit(polytope, f).  Call with null
  // outputs to query n_q.  values[i][q], grads[i][q][d].
  int
  po_ah_reinit(void        *ah,
               int          kind,
               int          poly,
               unsigned int f,
               double      *points,
               double      *jxw,
               double      *normals,
               double      *values,
               double      *grads)
  {
    Handler      *h = static_cast<Handler *>(ah);
    FEValuesTable t;
    if (guard([&] {
          if (kind == 0)
            h->reinit(poly, t);
          else
            h->reinit_face(poly, f, t);
        }))
      return -1;
    if (points)
      std::copy(t.points.begin(), t.points.end(), points);
    if (jxw)
      std::copy(t.jxw.begin(), t.jxw.end(), jxw);
    if (normals && kind == 1)
      std::copy(t.normals.begin(), t.normals.end(), normals);
    if (values)
      std::copy(t.values.begin(), t.values.end(), values);
    if (grads)
      std::copy(t.grads.begin(), t.grads.end(), grads);
    return t.n_q;
  }

  // ---- sparsity / assembly / vmult -----------------------------------------
  int64_t
  po_ah_sparsity_nnz(void *ah)
  {
    std::vector<int64_t> rp;
    std::vector<int>     c;
    static_cast<Handler *>(ah)->sparsity_pattern(rp, c);
    return (int64_t)c.size();
  }
  void
  po_ah_sparsity(void *ah, int64_t *rowptr, int *cols)
  {
    std::vector<int64_t> rp;
    std::vector<int>     c;
    static_cast<Handler *>(ah)->sparsity_pattern(rp, c);
    std::copy(rp.begin(), rp.end(), rowptr);
    std::copy(c.begin(), c.end(), cols);
  }

  struct po_assemble_params
  {
    double penalty_constant;
    int    h_rule;
    double h_const;
    int    visit_rule;
    int    with_boundary;
    double stiffness_coeff;
    double mass_coeff;
    int    n_threads;
    int    poly_stride;
    int    poly_offset;
    int    discard_scatter;
  };

  void *
  po_assemble_dg_matrix(void *ah, const po_assemble_params *p, double *seconds)
  {
    Matrix        *M = new Matrix;
    AssembleParams prm;
    prm.penalty_constant = p->penalty_constant;
    prm.h_rule           = p->h_rule;
    prm.h_const          = p->h_const;
    prm.visit_rule       = p->visit_rule;
    prm.with_boundary    = p->with_boundary;
    prm.stiffness_coeff  = p->stiffness_coeff;
    prm.mass_coeff       = p->mass_coeff;
    prm.n_threads        = p->n_threads;
    prm.poly_stride      = p->poly_stride > 0 ? p->poly_stride : 1;
    prm.poly_offset      = p->poly_offset;
    prm.discard_scatter  = p->discard_scatter;
    const auto t0        = std::chrono::steady_clock::now();
    if (guard([&] { assemble_dg_matrix(*static_cast<Handler *>(ah), prm, M->A); }))
      {
        delete M;
        return nullptr;
      }
    if (seconds)
      *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return M;
  }
  // complete block rows of the listed polytopes (assemble_block_rows): returns the number of blocks;
  // bcol / vals may be null to query sizes (vals: n_blocks * n^2 doubles)
  int64_t
  po_assemble_block_rows(void *ah, const po_assemble_params *p, const int *polys, int n_polys, int64_t *ptr, int *bcol,
                         double *vals, double *seconds)
  {
    AssembleParams prm;
    prm.penalty_constant = p->penalty_constant;
    prm.h_rule           = p->h_rule;
    prm.h_const          = p->h_const;
    prm.visit_rule       = p->visit_rule;
    prm.with_boundary    = p->with_boundary;
    prm.stiffness_coeff  = p->stiffness_coeff;
    prm.mass_coeff       = p->mass_coeff;
    prm.n_threads        = p->n_threads;
    BlockRows  R;
    const auto t0 = std::chrono::steady_clock::now();
    if (guard([&] { assemble_block_rows(*static_cast<Handler *>(ah), prm, polys, n_polys, R); }))
      return -1;
    if (seconds)
      *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::copy(R.ptr.begin(), R.ptr.end(), ptr);
    if (bcol)
      std::copy(R.bcol.begin(), R.bcol.end(), bcol);
    if (vals)
      std::copy(R.vals.begin(), R.vals.end(), vals);
    return R.ptr.back();
  }
  void
  po_matrix_free(void *M)
  {
    delete static_cast<Matrix *>(M);
  }
  int64_t
  po_matrix_n_rows(void *M)
  {
    return (int64_t) static_cast<Matrix *>(M)->A.rowptr.size() - 1;
  }
  int64_t
  po_matrix_nnz(void *M)
  {
    return (int64_t) static_cast<Matrix *>(M)->A.cols.size();
  }
  void
  po_matrix_copy(void *M_, int64_t *rowptr, int *cols, double *vals)
  {
    Matrix *M = static_cast<Matrix *>(M_);
    if (rowptr)
      std::copy(M->A.rowptr.begin(), M->A.rowptr.end(), rowptr);
    if (cols)
      std::copy(M->A.cols.begin(), M->A.cols.end(), cols);
    if (vals)
      std::copy(M->A.vals.begin(), M->A.vals.end(), vals);
  }
  void
  po_matrix_vmult(void *M, const double *x, double *y, int n_threads)
  {
    spmv(static_cast<Matrix *>(M)->A, x, y, n_threads);
  }
  int
  po_mapped_fine_vmult(void *g, int degree, int nq, double stiffness, double mass, int with_volume, int with_boundary,
                       int with_interior, const double *x, double *y)
  {
    return guard([&] {
      mapped_fine_vmult(*static_cast<Grid *>(g), degree, nq, stiffness, mass, with_volume != 0, with_boundary != 0,
                        with_interior != 0, x, y);
    });
  }
}
