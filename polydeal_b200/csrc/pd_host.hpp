// -----------------------------------------------------------------------------
// pd_host.hpp -- host mirror of the reference's polytopal-mesh layer.
//
// Produces, with the reference's numbering, the flattened agglomeration the
// device kernels consume.  Names follow the reference:
//   AgglomerationHandler  include/agglomeration_handler.h:171-575,
//                         source/agglomeration_handler.cc
//   accessor queries      include/agglomeration_accessor.h:41-299
//   MappingBox            source/mapping_box.cc:194-224, 923-972
// Unlike the reference (std::map / std::set keyed on CellId pairs, one heap
// object per reinit) everything is flat CSR built in two linear sweeps.
// No arithmetic of the hot path runs here: quadrature, basis evaluation and
// the SIP terms are computed by the CUDA kernels only.
// -----------------------------------------------------------------------------
#pragma once
#include "../../include/polydeal_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace pd
{
  struct Error : std::runtime_error
  {
    int code;
    Error(int c, const std::string &m)
      : std::runtime_error(m)
      , code(c)
    {}
  };

  // ---------------------------------------------------------------------------
  // Background mesh (quads / hexes, deal.II vertex and face numbering)
  // ---------------------------------------------------------------------------
  struct Grid
  {
    int                  dim = 0;
    std::vector<double>  verts;      // [n_verts][dim]
    std::vector<int32_t> cell_verts; // [n_cells][2^dim]
    std::vector<int32_t> nbr;        // [n_cells][2*dim]

    int64_t
    n_cells() const
    {
      return (int64_t)(cell_verts.size() >> dim);
    }
    int64_t
    n_verts() const
    {
      return (int64_t)(verts.size() / dim);
    }

    // cell->neighbor_of_neighbor(f): the face through which the neighbour across face f sees cell c -- the opposite
    // face on structured grids; looked up on unstructured meshes, where neighbours may be rotated against each other
    int32_t
    neighbor_of_neighbor(const int32_t c, const int f) const
    {
      const int     fpc = 2 * dim;
      const int32_t q   = nbr[(size_t)c * fpc + f];
      if (q >= 0 && nbr[(size_t)q * fpc + (f ^ 1)] == c)
        return f ^ 1;
      for (int g = 0; q >= 0 && g < fpc; ++g)
        if (nbr[(size_t)q * fpc + g] == c)
          return g;
      throw Error(PD_ERR_INVALID, "Grid: the neighbour table is not symmetric");
    }

    // interleave the bits of (i,j,k): position of a cell in the hierarchical
    // (refine_global) ordering of a 2^L grid
    static inline uint64_t
    morton(const int dim, const int levels, const uint32_t i, const uint32_t j, const uint32_t k)
    {
      uint64_t m = 0;
      for (int l = 0; l < levels; ++l)
        {
          m |= (uint64_t)((i >> l) & 1u) << (dim * l);
          m |= (uint64_t)((j >> l) & 1u) << (dim * l + 1);
          if (dim == 3)
            m |= (uint64_t)((k >> l) & 1u) << (dim * l + 2);
        }
      return m;
    }

    void
    make_structured(const int dim_, const int32_t *n, const double *lo, const double *hi, const int order)
    {
      if (dim_ != 2 && dim_ != 3)
        throw Error(PD_ERR_INVALID, "dim must be 2 or 3");
      dim             = dim_;
      const int64_t nx = n[0], ny = n[1], nz = dim == 3 ? n[2] : 1;
      if (nx < 1 || ny < 1 || nz < 1)
        throw Error(PD_ERR_INVALID, "grid sizes must be positive");
      int levels = 0;
      if (order == 0)
        {
          while ((int64_t(1) << levels) < nx)
            ++levels;
          if ((int64_t(1) << levels) != nx || ny != nx || (dim == 3 && nz != nx))
            throw Error(PD_ERR_INVALID, "hierarchical (Morton) order needs n = 2^k in every direction");
        }
      const int64_t vx = nx + 1, vy = ny + 1, vz = dim == 3 ? nz + 1 : 1;
      verts.resize((size_t)(vx * vy * vz) * dim);
      const double inv[3] = {1.0 / nx, 1.0 / ny, 1.0 / nz};
      for (int64_t k = 0; k < vz; ++k)
        for (int64_t j = 0; j < vy; ++j)
          for (int64_t i = 0; i < vx; ++i)
            {
              double      *x      = &verts[(size_t)((k * vy + j) * vx + i) * dim];
              const int64_t ijk[3] = {i, j, k};
              for (int d = 0; d < dim; ++d)
                x[d] = lo[d] + (hi[d] - lo[d]) * ((double)ijk[d] * inv[d]);
            }
      const int     vpc = 1 << dim, fpc = 2 * dim;
      const int64_t nc  = nx * ny * nz;
      cell_verts.resize((size_t)nc * vpc);
      nbr.resize((size_t)nc * fpc);
      auto id = [&](int64_t i, int64_t j, int64_t k) -> int64_t {
        return order == 0 ? (int64_t)morton(dim, levels, (uint32_t)i, (uint32_t)j, (uint32_t)k) :
                            (k * ny + j) * nx + i;
      };
      for (int64_t k = 0; k < nz; ++k)
        for (int64_t j = 0; j < ny; ++j)
          for (int64_t i = 0; i < nx; ++i)
            {
              const int64_t c = id(i, j, k);
              for (int v = 0; v < vpc; ++v)
                cell_verts[(size_t)c * vpc + v] =
                  (int32_t)(((k + ((v >> 2) & 1)) * vy + (j + ((v >> 1) & 1))) * vx + (i + (v & 1)));
              int32_t *b = &nbr[(size_t)c * fpc];
              b[0]       = i ? (int32_t)id(i - 1, j, k) : -1;
              b[1]       = i + 1 < nx ? (int32_t)id(i + 1, j, k) : -1;
              b[2]       = j ? (int32_t)id(i, j - 1, k) : -1;
              b[3]       = j + 1 < ny ? (int32_t)id(i, j + 1, k) : -1;
              if (dim == 3)
                {
                  b[4] = k ? (int32_t)id(i, j, k - 1) : -1;
                  b[5] = k + 1 < nz ? (int32_t)id(i, j, k + 1) : -1;
                }
            }
    }
  };

  // ---------------------------------------------------------------------------
  // AgglomerationHandler mirror
  // ---------------------------------------------------------------------------
  class AgglomerationHandler
  {
  public:
    explicit AgglomerationHandler(Grid *g)
      : grid(g)
      , dim(g->dim)
      , poly_of_cell((size_t)g->n_cells(), -1)
    {
      if (g->n_cells() == 0)
        throw Error(PD_ERR_INVALID, "The triangulation must not be empty upon calling this function.");
      subcell_ptr.push_back(0);
    }

    // source/agglomeration_handler.cc:45-104.  cells[0] becomes the master.
    int32_t
    define_agglomerate(const int32_t *cells, const int32_t n)
    {
      if (n <= 0)
        throw Error(PD_ERR_INVALID, "No cells to be agglomerated.");
      const int32_t p = (int32_t)masters.size();
      for (int32_t i = 0; i < n; ++i)
        {
          if (cells[i] < 0 || cells[i] >= grid->n_cells())
            throw Error(PD_ERR_INVALID, "cell index out of range in define_agglomerate");
          if (poly_of_cell[cells[i]] >= 0)
            throw Error(PD_ERR_INVALID, "cell " + std::to_string(cells[i]) + " already belongs to an agglomerate");
        }
      masters.push_back(cells[0]);
      // stored order = the accessor's get_agglomerate(): slaves first, master last
      for (int32_t i = 1; i < n; ++i)
        subcell_idx.push_back(cells[i]);
      subcell_idx.push_back(cells[0]);
      subcell_ptr.push_back((int64_t)subcell_idx.size());
      for (int32_t i = 0; i < n; ++i)
        poly_of_cell[cells[i]] = p;
      // bounding box of all vertices of all cells (:476-491)
      double b[6];
      for (int d = 0; d < dim; ++d)
        {
          b[d]       = std::numeric_limits<double>::infinity();
          b[dim + d] = -std::numeric_limits<double>::infinity();
        }
      const int vpc = 1 << dim;
      for (int32_t i = 0; i < n; ++i)
        for (int v = 0; v < vpc; ++v)
          {
            const double *x = &grid->verts[(size_t)grid->cell_verts[(size_t)cells[i] * vpc + v] * dim];
            for (int d = 0; d < dim; ++d)
              {
                b[d]       = std::min(b[d], x[d]);
                b[dim + d] = std::max(b[dim + d], x[d]);
              }
          }
      bbox.insert(bbox.end(), b, b + 2 * dim);
      connectivity_ready = false;
      return p;
    }

    void
    initialize_fe_values(const int32_t nq_cell, const int32_t nq_face)
    {
      if (nq_cell < 1 || nq_face < 1 || nq_cell > 8 || nq_face > 8)
        throw Error(PD_ERR_INVALID, "number of Gauss points per direction must be in [1,8]");
      n_q1d      = nq_cell;
      n_q1d_face = nq_face;
    }

    // :326-379 -- DoFs (hp: FE on masters, FE_Nothing on slaves => consecutive
    // blocks in active-cell order of the masters) + connectivity
    void
    distribute_agglomerated_dofs(const int32_t fe_kind, const int32_t degree)
    {
      // the reference accepts FE_DGQ, FE_AggloDGP (and FE_SimplexDGP, not on this path),
      // source/agglomeration_handler.cc:331-337
      if (fe_kind != PD_FE_DGQ && fe_kind != PD_FE_AGGLODGP)
        throw Error(PD_ERR_UNSUPPORTED, "Currently, this interface supports only DGQ and DGP bases.");
      if (degree < 0 || degree > 5)
        throw Error(PD_ERR_UNSUPPORTED, "the polynomial degree must be in [0,5]");
      if (masters.empty())
        throw Error(PD_ERR_STATE, "No agglomeration has been performed.");
      fe_degree     = degree;
      this->fe_kind = fe_kind;
      dofs_per_cell = 1;
      if (fe_kind == PD_FE_DGQ)
        for (int d = 0; d < dim; ++d)
          dofs_per_cell *= degree + 1;
      else // C(p + dim, dim)
        for (int d = 1; d <= dim; ++d)
          dofs_per_cell = dofs_per_cell * (degree + d) / d;
      // rank of each master among all masters by active cell index
      std::vector<int32_t> order(masters.size());
      for (size_t i = 0; i < order.size(); ++i)
        order[i] = (int32_t)i;
      std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return masters[a] < masters[b]; });
      dof_block.resize(masters.size());
      block_to_poly.resize(masters.size());
      for (size_t r = 0; r < order.size(); ++r)
        {
          dof_block[order[r]] = (int32_t)r;
          block_to_poly[r]    = order[r];
        }
      setup_connectivity_of_agglomeration();
    }

    // --- accessor-like queries ------------------------------------------------
    int32_t
    n_polytopes() const
    {
      return (int32_t)masters.size();
    }
    int64_t
    n_dofs() const
    {
      return (int64_t)masters.size() * dofs_per_cell;
    }
    void
    check_poly(const int32_t p) const
    {
      if (p < 0 || p >= n_polytopes())
        throw Error(PD_ERR_INVALID, "polytope index out of range");
    }
    void
    check_face(const int32_t p, const uint32_t f) const
    {
      check_poly(p);
      require_connectivity();
      if (f >= n_faces(p))
        throw Error(PD_ERR_INVALID, "face index out of range"); // reference: std::map::at -> std::out_of_range
    }
    void
    require_connectivity() const
    {
      if (!connectivity_ready)
        throw Error(PD_ERR_STATE,
                    "The DoFHandler associated to the agglomeration has not been initialized. "
                    "It's likely that you forgot to distribute the DoFs.");
    }
    uint32_t
    n_faces(const int32_t p) const
    {
      return (uint32_t)(face_ptr[p + 1] - face_ptr[p]);
    }
    bool
    at_boundary(const int32_t p, const uint32_t f) const
    {
      return face_nbr[face_ptr[p] + f] < 0;
    }
    int32_t
    neighbor(const int32_t p, const uint32_t f) const
    {
      return face_nbr[face_ptr[p] + f];
    }
    uint32_t
    neighbor_of_agglomerated_neighbor(const int32_t p, const uint32_t f) const
    {
      return face_nofn[face_ptr[p] + f];
    }
    double
    diameter(const int32_t p) const
    {
      const double *b = &bbox[(size_t)p * 2 * dim];
      double        s = 0;
      for (int d = 0; d < dim; ++d)
        s += (b[dim + d] - b[d]) * (b[dim + d] - b[d]);
      return std::sqrt(s);
    }
    double
    volume(const int32_t p) const
    {
      const double *b = &bbox[(size_t)p * 2 * dim];
      double        v = 1;
      for (int d = 0; d < dim; ++d)
        v *= (b[dim + d] - b[d]);
      return v;
    }

    // block-CSR pattern (:910-1022): diagonal + one block per neighbouring polytope
    void
    block_pattern(std::vector<int64_t> &brow_ptr, std::vector<int32_t> &bcol) const
    {
      require_connectivity();
      const int32_t np = n_polytopes();
      brow_ptr.assign(np + 1, 0);
      bcol.clear();
      std::vector<int32_t> row;
      for (int32_t b = 0; b < np; ++b)
        {
          const int32_t p = block_to_poly[b];
          row.clear();
          row.push_back(b);
          for (uint32_t f = 0; f < n_faces(p); ++f)
            if (!at_boundary(p, f))
              row.push_back(dof_block[neighbor(p, f)]);
          std::sort(row.begin(), row.end());
          bcol.insert(bcol.end(), row.begin(), row.end());
          brow_ptr[b + 1] = (int64_t)bcol.size();
        }
    }

    // Fill a descriptor; the arrays it points to live in this object.
    void
    flatten(const pdh_flatten_params &prm, pd_mesh_desc &d)
    {
      require_connectivity();
      if (n_q1d <= 0)
        throw Error(PD_ERR_STATE, "initialize_fe_values() must be called before flattening");
      const double C =
        prm.penalty_constant >= 0 ? prm.penalty_constant : 10.0 * (fe_degree + dim) * (fe_degree + 1);
      fl_polyA.clear();
      fl_polyB.clear();
      fl_sub_ptr.assign(1, 0);
      fl_sub_cell.clear();
      fl_sub_face.clear();
      fl_sub_sigma.clear();
      const int32_t np = n_polytopes();
      fl_face_item.assign((size_t)face_ptr[np], -1);
      for (int32_t p = 0; p < np; ++p)
        {
          const double hp = diameter(p);
          for (uint32_t f = 0; f < n_faces(p); ++f)
            {
              const int32_t q = neighbor(p, f);
              if (q >= 0)
                {
                  const bool visit = prm.visit_rule == PD_VISIT_BY_ID ? masters[p] < masters[q] : p < q;
                  if (!visit)
                    continue;
                  fl_face_item[(size_t)(face_ptr[q] + neighbor_of_agglomerated_neighbor(p, f))] = 2 * (int64_t)fl_polyA.size() + 1;
                }
              fl_face_item[(size_t)(face_ptr[p] + f)] = 2 * (int64_t)fl_polyA.size();
              fl_polyA.push_back(p);
              fl_polyB.push_back(q);
              const int64_t fi = face_ptr[p] + f;
              for (int64_t s = face_sub_ptr[fi]; s < face_sub_ptr[fi + 1]; ++s)
                {
                  fl_sub_cell.push_back(sub_cell[s]);
                  fl_sub_face.push_back(sub_face[s]);
                  double sigma;
                  switch (prm.h_rule)
                    {
                      case PD_H_MAX_INVERSE_DIAMETER:
                        sigma = q >= 0 ? C * std::max(1.0 / hp, 1.0 / diameter(q)) : C / hp;
                        break;
                      case PD_H_CONSTANT:
                        sigma = C / prm.h_const;
                        break;
                      case PD_H_NORMAL_EXTENT:
                        {
                          const int     nd = sub_face[s] / 2;
                          const double *ba = &bbox[(size_t)p * 2 * dim];
                          const double  ia = 1.0 / (ba[dim + nd] - ba[nd]);
                          if (q >= 0)
                            {
                              const double *bb = &bbox[(size_t)q * 2 * dim];
                              sigma            = C * (ia + 1.0 / (bb[dim + nd] - bb[nd]));
                            }
                          else
                            sigma = 4.0 * C * ia;
                          break;
                        }
                      default:
                        sigma = C / hp;
                    }
                  fl_sub_sigma.push_back(sigma);
                }
              fl_sub_ptr.push_back((int64_t)fl_sub_cell.size());
            }
        }
      block_pattern(fl_brow_ptr, fl_bcol);

      d                  = pd_mesh_desc{};
      d.dim              = dim;
      d.fe_degree        = fe_degree;
      d.fe_kind          = fe_kind;
      d.n_q1d            = n_q1d;
      d.n_q1d_face       = n_q1d_face;
      d.n_verts          = grid->n_verts();
      d.verts            = grid->verts.data();
      d.n_cells          = grid->n_cells();
      d.cell_verts       = grid->cell_verts.data();
      d.n_polytopes      = np;
      d.poly_subcell_ptr = subcell_ptr.data();
      d.poly_subcell_idx = subcell_idx.data();
      d.bbox             = bbox.data();
      d.dof_block        = dof_block.data();
      d.n_ifaces         = (int32_t)fl_polyA.size();
      d.iface_polyA      = fl_polyA.data();
      d.iface_polyB      = fl_polyB.data();
      d.iface_sub_ptr    = fl_sub_ptr.data();
      d.sub_cell         = fl_sub_cell.data();
      d.sub_face         = fl_sub_face.data();
      d.sub_sigma        = fl_sub_sigma.data();
      d.n_block_rows     = np;
      d.brow_ptr         = fl_brow_ptr.data();
      d.bcol_idx         = fl_bcol.data();
    }

    // Flatten the piece of the agglomeration one rank needs (owner-computes-rows):
    // its owned polytopes (define order; local block order = global block order
    // restricted to the owned ones), then the ghost polytopes adjacent to them, grouped
    // by owning rank and ordered by global block inside a group -- so that the ghost
    // section of a source vector is the concatenation of what each peer sends.
    // Replaces setup_ghost_polytopes / exchange_interface_values
    // (source/agglomeration_handler.cc:531-618, 1026-1091): instead of shipping basis
    // values and gradients at the face points, the ghost's bounding box and DoF block are
    // enough, because the owning side re-evaluates the ghost basis itself.
    void
    flatten_local(const pdh_flatten_params &prm, const int32_t *owner, const int32_t rank, pd_mesh_desc &d,
                  pdh_local_info &info)
    {
      require_connectivity();
      if (n_q1d <= 0)
        throw Error(PD_ERR_STATE, "initialize_fe_values() must be called before flattening");
      const double C =
        prm.penalty_constant >= 0 ? prm.penalty_constant : 10.0 * (fe_degree + dim) * (fe_degree + 1);
      const int32_t np = n_polytopes();
      lc_poly_global.clear();
      std::vector<int32_t> g2l(np, -1);
      for (int32_t p = 0; p < np; ++p)
        if (owner[p] == rank)
          {
            g2l[p] = (int32_t)lc_poly_global.size();
            lc_poly_global.push_back(p);
          }
      const int32_t n_own = (int32_t)lc_poly_global.size();
      if (n_own == 0)
        throw Error(PD_ERR_INVALID, "rank " + std::to_string(rank) + " owns no polytope");
      // ghosts: neighbours owned elsewhere, sorted by (owner, global block)
      std::vector<int32_t> ghosts;
      for (int32_t lp = 0; lp < n_own; ++lp)
        {
          const int32_t p = lc_poly_global[lp];
          for (uint32_t f = 0; f < n_faces(p); ++f)
            {
              const int32_t q = neighbor(p, f);
              if (q >= 0 && owner[q] != rank && g2l[q] == -1)
                {
                  g2l[q] = -2; // mark
                  ghosts.push_back(q);
                }
            }
        }
      std::sort(ghosts.begin(), ghosts.end(), [&](int32_t a, int32_t b) {
        return owner[a] != owner[b] ? owner[a] < owner[b] : dof_block[a] < dof_block[b];
      });
      for (const int32_t q : ghosts)
        {
          g2l[q] = (int32_t)lc_poly_global.size();
          lc_poly_global.push_back(q);
        }
      const int32_t n_loc = (int32_t)lc_poly_global.size(), n_ghost = n_loc - n_own;
      // local block numbering
      std::vector<int32_t> own_sorted(lc_poly_global.begin(), lc_poly_global.begin() + n_own);
      std::sort(own_sorted.begin(), own_sorted.end(), [&](int32_t a, int32_t b) { return dof_block[a] < dof_block[b]; });
      lc_dof_block.assign(n_loc, 0);
      lc_owned_global_block.resize(n_own);
      for (int32_t r = 0; r < n_own; ++r)
        {
          lc_dof_block[g2l[own_sorted[r]]] = r;
          lc_owned_global_block[r]          = dof_block[own_sorted[r]];
        }
      lc_ghost_global_block.resize(n_ghost);
      lc_ghost_owner.resize(n_ghost);
      for (int32_t k = 0; k < n_ghost; ++k)
        {
          lc_dof_block[n_own + k]  = n_own + k;
          lc_ghost_global_block[k] = dof_block[ghosts[k]];
          lc_ghost_owner[k]        = owner[ghosts[k]];
        }
      // sub-cells, bounding boxes
      lc_subcell_ptr.assign(1, 0);
      lc_subcell_idx.clear();
      lc_bbox.clear();
      for (int32_t lp = 0; lp < n_loc; ++lp)
        {
          const int32_t p = lc_poly_global[lp];
          if (lp < n_own)
            lc_subcell_idx.insert(lc_subcell_idx.end(), subcell_idx.begin() + subcell_ptr[p],
                                  subcell_idx.begin() + subcell_ptr[p + 1]);
          lc_subcell_ptr.push_back((int64_t)lc_subcell_idx.size());
          lc_bbox.insert(lc_bbox.end(), bbox.begin() + (size_t)p * 2 * dim, bbox.begin() + (size_t)(p + 1) * 2 * dim);
        }
      // faces: always listed from the owned side; sigma by the reference's rule
      lc_polyA.clear();
      lc_polyB.clear();
      lc_sub_ptr.assign(1, 0);
      lc_sub_cell.clear();
      lc_sub_face.clear();
      lc_sub_sigma.clear();
      std::vector<std::vector<int32_t>> rows(n_own);
      for (int32_t lp = 0; lp < n_own; ++lp)
        {
          const int32_t p = lc_poly_global[lp];
          rows[lc_dof_block[lp]].push_back(lc_dof_block[lp]);
          for (uint32_t f = 0; f < n_faces(p); ++f)
            {
              const int32_t q = neighbor(p, f);
              if (q >= 0)
                rows[lc_dof_block[lp]].push_back(lc_dof_block[g2l[q]]);
              const bool p_visits = q < 0 || (prm.visit_rule == PD_VISIT_BY_ID ? masters[p] < masters[q] : p < q);
              if (q >= 0 && owner[q] == rank && !p_visits)
                continue; // both owned: listed once, from the visiting side
              const double h_visitor = p_visits ? diameter(p) : diameter(q);
              lc_polyA.push_back(lp);
              lc_polyB.push_back(q >= 0 ? g2l[q] : -1);
              const int64_t fi = face_ptr[p] + f;
              for (int64_t s = face_sub_ptr[fi]; s < face_sub_ptr[fi + 1]; ++s)
                {
                  lc_sub_cell.push_back(sub_cell[s]);
                  lc_sub_face.push_back(sub_face[s]);
                  double sigma;
                  switch (prm.h_rule)
                    {
                      case PD_H_MAX_INVERSE_DIAMETER:
                        sigma = q >= 0 ? C * std::max(1.0 / diameter(p), 1.0 / diameter(q)) : C / diameter(p);
                        break;
                      case PD_H_CONSTANT:
                        sigma = C / prm.h_const;
                        break;
                      case PD_H_NORMAL_EXTENT:
                        {
                          const int     nd = sub_face[s] / 2;
                          const double *ba = &bbox[(size_t)p * 2 * dim];
                          const double  ia = 1.0 / (ba[dim + nd] - ba[nd]);
                          if (q >= 0)
                            {
                              const double *bb = &bbox[(size_t)q * 2 * dim];
                              sigma            = C * (ia + 1.0 / (bb[dim + nd] - bb[nd]));
                            }
                          else
                            sigma = 4.0 * C * ia;
                          break;
                        }
                      default:
                        sigma = C / h_visitor;
                    }
                  lc_sub_sigma.push_back(sigma);
                }
              lc_sub_ptr.push_back((int64_t)lc_sub_cell.size());
            }
        }
      lc_brow_ptr.assign(1, 0);
      lc_bcol.clear();
      for (int32_t b = 0; b < n_own; ++b)
        {
          std::sort(rows[b].begin(), rows[b].end());
          lc_bcol.insert(lc_bcol.end(), rows[b].begin(), rows[b].end());
          lc_brow_ptr.push_back((int64_t)lc_bcol.size());
        }
      d                   = pd_mesh_desc{};
      d.dim               = dim;
      d.fe_degree         = fe_degree;
      d.fe_kind           = fe_kind;
      d.n_q1d             = n_q1d;
      d.n_q1d_face        = n_q1d_face;
      // the rank only needs the cells of its own polytopes (interfaces are listed from the
      // owned side): compress the mesh to them and renumber cells and vertices locally
      {
        const int            vpc = 1 << dim;
        std::vector<int32_t> cell_l(grid->n_cells(), -1), vert_l(grid->n_verts(), -1);
        lc_cell_verts.clear();
        lc_verts.clear();
        int32_t n_lc = 0, n_lv = 0;
        for (int32_t &c : lc_subcell_idx)
          {
            if (cell_l[c] < 0)
              {
                cell_l[c] = n_lc++;
                for (int v = 0; v < vpc; ++v)
                  {
                    const int32_t gv = grid->cell_verts[(size_t)c * vpc + v];
                    if (vert_l[gv] < 0)
                      {
                        vert_l[gv] = n_lv++;
                        lc_verts.insert(lc_verts.end(), grid->verts.begin() + (size_t)gv * dim,
                                        grid->verts.begin() + (size_t)(gv + 1) * dim);
                      }
                    lc_cell_verts.push_back(vert_l[gv]);
                  }
              }
            c = cell_l[c];
          }
        for (int32_t &c : lc_sub_cell)
          c = cell_l[c]; // always an owned cell
        d.n_verts    = n_lv;
        d.verts      = lc_verts.data();
        d.n_cells    = n_lc;
        d.cell_verts = lc_cell_verts.data();
      }
      d.n_polytopes       = n_loc;
      d.n_owned_polytopes = n_own;
      d.poly_subcell_ptr  = lc_subcell_ptr.data();
      d.poly_subcell_idx  = lc_subcell_idx.data();
      d.bbox              = lc_bbox.data();
      d.dof_block         = lc_dof_block.data();
      d.n_ifaces          = (int32_t)lc_polyA.size();
      d.iface_polyA       = lc_polyA.data();
      d.iface_polyB       = lc_polyB.data();
      d.iface_sub_ptr     = lc_sub_ptr.data();
      d.sub_cell          = lc_sub_cell.data();
      d.sub_face          = lc_sub_face.data();
      d.sub_sigma         = lc_sub_sigma.data();
      d.n_block_rows      = n_own;
      d.brow_ptr          = lc_brow_ptr.data();
      d.bcol_idx          = lc_bcol.data();
      info.n_owned            = n_own;
      info.n_ghost            = n_ghost;
      info.owned_global_block = lc_owned_global_block.data();
      info.ghost_global_block = lc_ghost_global_block.data();
      info.ghost_owner        = lc_ghost_owner.data();
      info.local_poly_global  = lc_poly_global.data();
    }

    // polytope face -> (entry of the last flatten's work list, side)
    void
    face_work_item(const int32_t p, const uint32_t f, int32_t &iface, int32_t &side) const
    {
      check_face(p, f);
      if (fl_face_item.size() != (size_t)face_ptr[n_polytopes()])
        throw Error(PD_ERR_STATE, "pdh_face_work_item: pdh_flatten has not been called");
      const int64_t v = fl_face_item[(size_t)(face_ptr[p] + f)];
      iface           = (int32_t)(v >> 1);
      side            = (int32_t)(v & 1);
    }
    std::vector<int64_t> fl_face_item;

    Grid     *grid;
    int       dim;
    int32_t   fe_degree = -1, fe_kind = 0, dofs_per_cell = 0, n_q1d = 0, n_q1d_face = 0;
    bool      connectivity_ready = false;

    std::vector<int32_t> poly_of_cell; // -1: cell not agglomerated
    std::vector<int32_t> masters;      // master cell of each polytope (define order)
    std::vector<int64_t> subcell_ptr;
    std::vector<int32_t> subcell_idx;
    std::vector<double>  bbox;
    std::vector<int32_t> dof_block, block_to_poly;

    // polytope faces in the reference's discovery order
    std::vector<int64_t>  face_ptr;     // [np+1]
    std::vector<int32_t>  face_nbr;     // neighbour polytope or -1 (boundary)
    std::vector<uint32_t> face_nofn;    // neighbor_of_agglomerated_neighbor
    std::vector<int64_t>  face_sub_ptr; // [n_faces_total+1]
    std::vector<int32_t>  sub_cell, sub_face;

  private:
    // Face enumeration with the ordering of
    // source/agglomeration_handler.cc:1253-1645: walk the sub-cells (slaves...,
    // master) and their local faces; the first sight of a neighbouring polytope
    // opens the next face index; all physical-boundary sub-faces share one face.
    // The (cell, face) list of an interface is recorded in the traversal order of
    // whichever of the two polytopes is set up FIRST (define order) and mirrored
    // for the other one -- which is what the reference's global visited set
    // yields (:1370-1397).
    void
    setup_connectivity_of_agglomeration()
    {
      for (int64_t c = 0; c < grid->n_cells(); ++c)
        if (poly_of_cell[c] < 0)
          throw Error(PD_ERR_INVALID,
                      "cell " + std::to_string(c) +
                        " belongs to no agglomerate (define singletons with define_agglomerate({cell}))");
      const int32_t np  = n_polytopes();
      const int     fpc = 2 * dim;
      face_ptr.assign(np + 1, 0);
      face_nbr.clear();
      std::vector<int32_t>              stamp(np, -1), slot(np, -1);
      std::vector<std::vector<int32_t>> own_cells, own_faces; // per global face, own traversal order
      for (int32_t p = 0; p < np; ++p)
        {
          const int64_t f0       = (int64_t)face_nbr.size();
          int32_t       bnd_slot = -1;
          for (int64_t s = subcell_ptr[p]; s < subcell_ptr[p + 1]; ++s)
            {
              const int32_t c = subcell_idx[s];
              for (int f = 0; f < fpc; ++f)
                {
                  const int32_t nc = grid->nbr[(size_t)c * fpc + f];
                  int32_t       fs;
                  if (nc < 0)
                    {
                      if (bnd_slot < 0)
                        {
                          bnd_slot = (int32_t)(face_nbr.size() - f0);
                          face_nbr.push_back(-1);
                          own_cells.emplace_back();
                          own_faces.emplace_back();
                        }
                      fs = bnd_slot;
                    }
                  else
                    {
                      const int32_t q = poly_of_cell[nc];
                      if (q == p)
                        continue;
                      if (stamp[q] != p)
                        {
                          stamp[q] = p;
                          slot[q]  = (int32_t)(face_nbr.size() - f0);
                          face_nbr.push_back(q);
                          own_cells.emplace_back();
                          own_faces.emplace_back();
                        }
                      fs = slot[q];
                    }
                  own_cells[f0 + fs].push_back(c);
                  own_faces[f0 + fs].push_back(f);
                }
            }
          face_ptr[p + 1] = (int64_t)face_nbr.size();
        }
      // neighbor_of_agglomerated_neighbor
      const int64_t nf = (int64_t)face_nbr.size();
      face_nofn.assign(nf, PD_INVALID_UINT);
      for (int32_t p = 0; p < np; ++p)
        for (int64_t fi = face_ptr[p]; fi < face_ptr[p + 1]; ++fi)
          {
            const int32_t q = face_nbr[fi];
            if (q < 0)
              continue;
            for (int64_t gi = face_ptr[q]; gi < face_ptr[q + 1]; ++gi)
              if (face_nbr[gi] == p)
                {
                  face_nofn[fi] = (uint32_t)(gi - face_ptr[q]);
                  break;
                }
          }
      // sub-face lists: own order if this side is set up first, else mirrored
      face_sub_ptr.assign(nf + 1, 0);
      sub_cell.clear();
      sub_face.clear();
      for (int32_t p = 0; p < np; ++p)
        for (int64_t fi = face_ptr[p]; fi < face_ptr[p + 1]; ++fi)
          {
            const int32_t q = face_nbr[fi];
            if (q < 0 || p < q)
              {
                sub_cell.insert(sub_cell.end(), own_cells[fi].begin(), own_cells[fi].end());
                sub_face.insert(sub_face.end(), own_faces[fi].begin(), own_faces[fi].end());
              }
            else
              {
                const int64_t gi = face_ptr[q] + face_nofn[fi];
                for (size_t k = 0; k < own_cells[gi].size(); ++k)
                  {
                    const int32_t c = own_cells[gi][k], f = own_faces[gi][k];
                    const int32_t nc = grid->nbr[(size_t)c * fpc + f];
                    sub_cell.push_back(nc);
                    sub_face.push_back(grid->neighbor_of_neighbor(c, f));
                  }
              }
            face_sub_ptr[fi + 1] = (int64_t)sub_cell.size();
          }
      connectivity_ready = true;
    }

    // storage behind flatten_local()
    std::vector<int32_t> lc_poly_global, lc_dof_block, lc_owned_global_block, lc_ghost_global_block, lc_ghost_owner,
      lc_subcell_idx, lc_polyA, lc_polyB, lc_sub_cell, lc_sub_face, lc_bcol;
    std::vector<int64_t> lc_subcell_ptr, lc_sub_ptr, lc_brow_ptr;
    std::vector<double>  lc_bbox, lc_sub_sigma, lc_verts;
    std::vector<int32_t> lc_cell_verts;
    // storage behind flatten()
    std::vector<int32_t> fl_polyA, fl_polyB, fl_sub_cell, fl_sub_face, fl_bcol;
    std::vector<int64_t> fl_sub_ptr, fl_brow_ptr;
    std::vector<double>  fl_sub_sigma;
  };
} // namespace pd
