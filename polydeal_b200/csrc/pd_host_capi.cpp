// -----------------------------------------------------------------------------
// pd_host_capi.cpp -- C ABI of the host mirror (include/polydeal_b200.h, pdh_*).
// -----------------------------------------------------------------------------
#include "pd_host.hpp"

#include <cstring>
#include <memory>

namespace pd
{
  void set_last_error(const std::string &m);
}

struct pdh_grid
{
  pd::Grid g;
};
struct pdh_handler
{
  std::unique_ptr<pd::AgglomerationHandler> ah;
};

namespace
{
  template <class F>
  int
  guarded(F &&f)
  {
    try
      {
        f();
        return PD_OK;
      }
    catch (const pd::Error &e)
      {
        pd::set_last_error(e.what());
        return e.code;
      }
    catch (const std::exception &e)
      {
        pd::set_last_error(e.what());
        return PD_ERR_INVALID;
      }
  }
  pd::AgglomerationHandler &
  H(const pdh_handler *ah)
  {
    if (!ah || !ah->ah)
      throw pd::Error(PD_ERR_INVALID, "null handler");
    return *ah->ah;
  }
} // namespace

extern "C"
{
  // METIS 5 API of the CUDA toolkit's libmetis_static.a: idx_t = int64_t, real_t = float
  int METIS_SetDefaultOptions(int64_t *options);
  int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                          int64_t *adjwgt, int64_t *nparts, float *tpwgts, float *ubvec, int64_t *options, int64_t *objval,
                          int64_t *part);
  int METIS_PartGraphRecursive(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                               int64_t *adjwgt, int64_t *nparts, float *tpwgts, float *ubvec, int64_t *options,
                               int64_t *objval, int64_t *part);
}

extern "C"
{
  int
  pdh_grid_create_structured(int32_t dim, const int32_t *n, const double *lo, const double *hi, int32_t order, pdh_grid **out)
  {
    return guarded([&] {
      if (!n || !lo || !hi || !out)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      std::unique_ptr<pdh_grid> g(new pdh_grid);
      g->g.make_structured(dim, n, lo, hi, order);
      *out = g.release();
    });
  }

  int
  pdh_grid_create(int32_t dim, int64_t n_verts, const double *verts, int64_t n_cells, const int32_t *cell_verts,
                  const int32_t *nbr, pdh_grid **out)
  {
    return guarded([&] {
      if (!verts || !cell_verts || !nbr || !out || (dim != 2 && dim != 3) || n_verts <= 0 || n_cells <= 0)
        throw pd::Error(PD_ERR_INVALID, "pdh_grid_create: bad argument");
      std::unique_ptr<pdh_grid> g(new pdh_grid);
      g->g.dim = dim;
      g->g.verts.assign(verts, verts + n_verts * dim);
      g->g.cell_verts.assign(cell_verts, cell_verts + (n_cells << dim));
      g->g.nbr.assign(nbr, nbr + n_cells * 2 * dim);
      for (int32_t v : g->g.cell_verts)
        if (v < 0 || v >= n_verts)
          throw pd::Error(PD_ERR_INVALID, "pdh_grid_create: vertex index out of range");
      for (int32_t c : g->g.nbr)
        if (c < -1 || c >= n_cells)
          throw pd::Error(PD_ERR_INVALID, "pdh_grid_create: neighbour index out of range");
      *out = g.release();
    });
  }

  int
  pdh_grid_destroy(pdh_grid *g)
  {
    delete g;
    return PD_OK;
  }
  int64_t
  pdh_grid_n_cells(const pdh_grid *g)
  {
    return g ? g->g.n_cells() : 0;
  }
  int64_t
  pdh_grid_n_verts(const pdh_grid *g)
  {
    return g ? g->g.n_verts() : 0;
  }
  int
  pdh_grid_set_vertices(pdh_grid *g, const double *verts)
  {
    return guarded([&] {
      if (!g || !verts)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      std::memcpy(g->g.verts.data(), verts, g->g.verts.size() * sizeof(double));
    });
  }
  int
  pdh_grid_get_arrays(const pdh_grid *g, double *verts, int32_t *cell_verts, int32_t *nbr)
  {
    return guarded([&] {
      if (!g)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      if (verts)
        std::memcpy(verts, g->g.verts.data(), g->g.verts.size() * sizeof(double));
      if (cell_verts)
        std::memcpy(cell_verts, g->g.cell_verts.data(), g->g.cell_verts.size() * sizeof(int32_t));
      if (nbr)
        std::memcpy(nbr, g->g.nbr.data(), g->g.nbr.size() * sizeof(int32_t));
    });
  }

  int
  pdh_handler_create(pdh_grid *g, pdh_handler **out)
  {
    return guarded([&] {
      if (!g || !out)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      std::unique_ptr<pdh_handler> h(new pdh_handler);
      h->ah.reset(new pd::AgglomerationHandler(&g->g));
      *out = h.release();
    });
  }
  int
  pdh_handler_destroy(pdh_handler *ah)
  {
    delete ah;
    return PD_OK;
  }
  int32_t
  pdh_define_agglomerate(pdh_handler *ah, const int32_t *cells, int32_t n)
  {
    int32_t r  = -1;
    const int e = guarded([&] { r = H(ah).define_agglomerate(cells, n); });
    return e == PD_OK ? r : e;
  }
  int
  pdh_define_agglomerates(pdh_handler *ah, int32_t n_groups, const int64_t *ptr, const int32_t *cells)
  {
    return guarded([&] {
      if (!ptr || !cells || n_groups < 0)
        throw pd::Error(PD_ERR_INVALID, "pdh_define_agglomerates: bad argument");
      for (int32_t g = 0; g < n_groups; ++g)
        H(ah).define_agglomerate(cells + ptr[g], (int32_t)(ptr[g + 1] - ptr[g]));
    });
  }
  int64_t
  pdh_polytope_graph(const pdh_handler *ah, int64_t *xadj, int64_t *adjncy, int64_t *vertex_weights, int64_t *edge_weights)
  {
    int64_t   n_edges = -1;
    const int e       = guarded([&] {
      const pd::AgglomerationHandler &a = H(ah);
      a.require_connectivity();
      const int32_t np = a.n_polytopes();
      int64_t       k  = 0;
      for (int32_t p = 0; p < np; ++p)
        {
          if (xadj)
            xadj[p] = k;
          if (vertex_weights)
            vertex_weights[p] = a.subcell_ptr[p + 1] - a.subcell_ptr[p];
          for (uint32_t f = 0; f < a.n_faces(p); ++f)
            if (!a.at_boundary(p, f))
              {
                if (adjncy)
                  adjncy[k] = a.neighbor(p, f);
                if (edge_weights)
                  edge_weights[k] = a.face_sub_ptr[a.face_ptr[p] + f + 1] - a.face_sub_ptr[a.face_ptr[p] + f];
                ++k;
              }
        }
      if (xadj)
        xadj[np] = k;
      n_edges = k;
    });
    return e == PD_OK ? n_edges : e;
  }
  int
  pdh_face_work_item(const pdh_handler *ah, int32_t poly, uint32_t f, int32_t *iface, int32_t *side)
  {
    return guarded([&] {
      if (!iface || !side)
        throw pd::Error(PD_ERR_INVALID, "pdh_face_work_item: null argument");
      H(ah).face_work_item(poly, f, *iface, *side);
    });
  }
  static int
  box_map(const pdh_handler *ah, int32_t poly, int64_t n, const double *in, double *out, bool to_unit)
  {
    return guarded([&] {
      const pd::AgglomerationHandler &a = H(ah);
      a.check_poly(poly);
      if ((!in || !out) && n > 0)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      const int     dim = a.dim;
      const double *b   = &a.bbox[(size_t)poly * 2 * dim];
      for (int64_t q = 0; q < n; ++q)
        for (int d = 0; d < dim; ++d)
          out[q * dim + d] = to_unit ? (in[q * dim + d] - b[d]) / (b[dim + d] - b[d]) : b[d] + in[q * dim + d] * (b[dim + d] - b[d]);
    });
  }
  int
  pdh_real_to_unit(const pdh_handler *ah, int32_t poly, int64_t n, const double *real_points, double *unit_points)
  {
    return box_map(ah, poly, n, real_points, unit_points, true);
  }
  int
  pdh_unit_to_real(const pdh_handler *ah, int32_t poly, int64_t n, const double *unit_points, double *real_points)
  {
    return box_map(ah, poly, n, unit_points, real_points, false);
  }
  int
  pdh_initialize_fe_values(pdh_handler *ah, int32_t nq_cell, int32_t nq_face)
  {
    return guarded([&] { H(ah).initialize_fe_values(nq_cell, nq_face); });
  }
  int
  pdh_distribute_agglomerated_dofs(pdh_handler *ah, int32_t fe_kind, int32_t degree)
  {
    return guarded([&] { H(ah).distribute_agglomerated_dofs(fe_kind, degree); });
  }
  int32_t
  pdh_n_polytopes(const pdh_handler *ah)
  {
    return ah && ah->ah ? ah->ah->n_polytopes() : 0;
  }
  int64_t
  pdh_n_dofs(const pdh_handler *ah)
  {
    return ah && ah->ah ? ah->ah->n_dofs() : 0;
  }
  int32_t
  pdh_n_dofs_per_cell(const pdh_handler *ah)
  {
    return ah && ah->ah ? ah->ah->dofs_per_cell : 0;
  }
  int32_t
  pdh_master_cell(const pdh_handler *ah, int32_t poly)
  {
    int32_t r  = -1;
    const int e = guarded([&] {
      H(ah).check_poly(poly);
      r = H(ah).masters[poly];
    });
    return e == PD_OK ? r : e;
  }
  int32_t
  pdh_n_background_cells(const pdh_handler *ah, int32_t poly)
  {
    int32_t r  = -1;
    const int e = guarded([&] {
      H(ah).check_poly(poly);
      r = (int32_t)(H(ah).subcell_ptr[poly + 1] - H(ah).subcell_ptr[poly]);
    });
    return e == PD_OK ? r : e;
  }
  int
  pdh_get_agglomerate(const pdh_handler *ah, int32_t poly, int32_t *cells)
  {
    return guarded([&] {
      auto &h = H(ah);
      h.check_poly(poly);
      std::copy(h.subcell_idx.begin() + h.subcell_ptr[poly], h.subcell_idx.begin() + h.subcell_ptr[poly + 1], cells);
    });
  }
  uint32_t
  pdh_n_faces(const pdh_handler *ah, int32_t poly)
  {
    uint32_t r = PD_INVALID_UINT;
    guarded([&] {
      auto &h = H(ah);
      h.check_poly(poly);
      h.require_connectivity();
      r = h.n_faces(poly);
    });
    return r;
  }
  int32_t
  pdh_at_boundary(const pdh_handler *ah, int32_t poly, uint32_t f)
  {
    int32_t r  = -1;
    const int e = guarded([&] {
      H(ah).check_face(poly, f);
      r = H(ah).at_boundary(poly, f) ? 1 : 0;
    });
    return e == PD_OK ? r : e;
  }
  int32_t
  pdh_neighbor(const pdh_handler *ah, int32_t poly, uint32_t f)
  {
    int32_t r = -1;
    guarded([&] {
      H(ah).check_face(poly, f);
      r = H(ah).neighbor(poly, f);
    });
    return r;
  }
  uint32_t
  pdh_neighbor_of_agglomerated_neighbor(const pdh_handler *ah, int32_t poly, uint32_t f)
  {
    uint32_t r = PD_INVALID_UINT;
    guarded([&] {
      H(ah).check_face(poly, f);
      r = H(ah).neighbor_of_agglomerated_neighbor(poly, f);
    });
    return r;
  }
  int32_t
  pdh_interface(const pdh_handler *ah, int32_t poly, uint32_t f, int32_t *cells, int32_t *faces, int32_t cap)
  {
    int32_t r  = -1;
    const int e = guarded([&] {
      auto &h = H(ah);
      h.check_face(poly, f);
      const int64_t fi = h.face_ptr[poly] + f;
      const int64_t s0 = h.face_sub_ptr[fi], s1 = h.face_sub_ptr[fi + 1];
      for (int64_t s = s0; s < s1 && s - s0 < cap; ++s)
        {
          cells[s - s0] = h.sub_cell[s];
          faces[s - s0] = h.sub_face[s];
        }
      r = (int32_t)(s1 - s0);
    });
    return e == PD_OK ? r : e;
  }
  int
  pdh_get_dof_indices(const pdh_handler *ah, int32_t poly, uint32_t *dofs)
  {
    return guarded([&] {
      auto &h = H(ah);
      h.check_poly(poly);
      if (h.dofs_per_cell <= 0)
        throw pd::Error(PD_ERR_STATE, "DoFs have not been distributed");
      for (int32_t i = 0; i < h.dofs_per_cell; ++i)
        dofs[i] = (uint32_t)(h.dof_block[poly] * h.dofs_per_cell + i);
    });
  }
  int
  pdh_bounding_box(const pdh_handler *ah, int32_t poly, double *lo, double *hi)
  {
    return guarded([&] {
      auto &h = H(ah);
      h.check_poly(poly);
      for (int d = 0; d < h.dim; ++d)
        {
          lo[d] = h.bbox[(size_t)poly * 2 * h.dim + d];
          hi[d] = h.bbox[(size_t)poly * 2 * h.dim + h.dim + d];
        }
    });
  }
  double
  pdh_diameter(const pdh_handler *ah, int32_t poly)
  {
    double r = -1;
    guarded([&] {
      H(ah).check_poly(poly);
      r = H(ah).diameter(poly);
    });
    return r;
  }
  double
  pdh_volume(const pdh_handler *ah, int32_t poly)
  {
    double r = -1;
    guarded([&] {
      H(ah).check_poly(poly);
      r = H(ah).volume(poly);
    });
    return r;
  }
  int64_t
  pdh_sparsity_nnz(const pdh_handler *ah)
  {
    int64_t r = -1;
    guarded([&] {
      std::vector<int64_t> bp;
      std::vector<int32_t> bc;
      H(ah).block_pattern(bp, bc);
      r = (int64_t)bc.size() * H(ah).dofs_per_cell * H(ah).dofs_per_cell;
    });
    return r;
  }
  int
  pdh_create_agglomeration_sparsity_pattern(const pdh_handler *ah, int64_t *rowptr, int32_t *cols)
  {
    return guarded([&] {
      auto                &h = H(ah);
      std::vector<int64_t> bp;
      std::vector<int32_t> bc;
      h.block_pattern(bp, bc);
      const int n = h.dofs_per_cell;
      int64_t   k = 0;
      rowptr[0]   = 0;
      for (int32_t b = 0; b < h.n_polytopes(); ++b)
        for (int i = 0; i < n; ++i)
          {
            for (int64_t e = bp[b]; e < bp[b + 1]; ++e)
              for (int j = 0; j < n; ++j)
                cols[k++] = bc[e] * n + j;
            rowptr[(int64_t)b * n + i + 1] = k;
          }
    });
  }
  int
  pdh_flatten(pdh_handler *ah, const pdh_flatten_params *prm, pd_mesh_desc *out)
  {
    return guarded([&] {
      if (!prm || !out)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      H(ah).flatten(*prm, *out);
    });
  }
  int
  pdh_flatten_local(pdh_handler *ah, const pdh_flatten_params *prm, const int32_t *owner, int32_t rank, pd_mesh_desc *out,
                    pdh_local_info *info)
  {
    return guarded([&] {
      if (!prm || !out || !owner || !info)
        throw pd::Error(PD_ERR_INVALID, "null argument");
      H(ah).flatten_local(*prm, owner, rank, *out, *info);
    });
  }
  int
  pdh_create_device(pdh_handler *ah, const pdh_flatten_params *prm, pd_handle **out)
  {
    pd_mesh_desc d;
    const int    e = pdh_flatten(ah, prm, &d);
    if (e != PD_OK)
      return e;
    return pd_create(&d, out);
  }

  // Graph partitioning with METIS (the 64-bit-index build shipped with the CUDA toolkit as
  // libmetis_static.a), called the way deal.II's SparsityTools::partition does
  // (include/poly_utils.h:603-606, GridTools::partition_triangulation in the examples and tests):
  // default options, METIS_PartGraphRecursive for nparts <= 8, METIS_PartGraphKway above.
  int
  pdh_partition_graph(int64_t n_vertices, const int64_t *xadj, const int64_t *adjncy, const int64_t *vertex_weights,
                      const int64_t *edge_weights, int32_t n_parts, int32_t *part_out)
  {
    return guarded([&] {
      if (n_vertices <= 0 || !xadj || !adjncy || !part_out || n_parts < 1)
        throw pd::Error(PD_ERR_INVALID, "pdh_partition_graph: bad argument");
      if (n_parts == 1)
        {
          std::fill(part_out, part_out + n_vertices, 0);
          return;
        }
      int64_t              nv = n_vertices, ncon = 1, np = n_parts, objval = 0;
      std::vector<int64_t> part((size_t)n_vertices), options(40);
      METIS_SetDefaultOptions(options.data());
      auto *xa = const_cast<int64_t *>(xadj), *ad = const_cast<int64_t *>(adjncy);
      auto *vw = const_cast<int64_t *>(vertex_weights), *ew = const_cast<int64_t *>(edge_weights);
      const int rc = n_parts <= 8 ? METIS_PartGraphRecursive(&nv, &ncon, xa, ad, vw, nullptr, ew, &np, nullptr, nullptr,
                                                             options.data(), &objval, part.data()) :
                                    METIS_PartGraphKway(&nv, &ncon, xa, ad, vw, nullptr, ew, &np, nullptr, nullptr,
                                                        options.data(), &objval, part.data());
      if (rc != 1)
        throw pd::Error(PD_ERR_INVALID, "pdh_partition_graph: METIS returned error " + std::to_string(rc));
      for (int64_t v = 0; v < n_vertices; ++v)
        part_out[v] = (int32_t)part[(size_t)v];
    });
  }
}
