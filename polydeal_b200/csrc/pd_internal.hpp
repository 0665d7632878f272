// -----------------------------------------------------------------------------
// pd_internal.hpp -- device-side state behind a pd_handle and launch prototypes.
// -----------------------------------------------------------------------------
#pragma once
#include "../../include/polydeal_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace pd
{
  void set_last_error(const std::string &m);

  struct CudaError
  {
    cudaError_t e;
    const char *what;
    int         line;
  };

#define PD_CUDA(call)                                             \
  do                                                              \
    {                                                             \
      cudaError_t e__ = (call);                                   \
      if (e__ != cudaSuccess)                                     \
        throw ::pd::CudaError{e__, #call, __LINE__};              \
    }                                                             \
  while (0)

  // 1-D Lagrange basis of FE_DGQ(p) on the Gauss-Lobatto nodes of [0,1]
  // (deal.II FE_DGQ; accepted by source/agglomeration_handler.cc:331-332),
  // in product form  l_a(x) = w_a prod_{b != a} (x - x_b)
  struct Basis1D
  {
    double node[6];
    double wprod[6];
  };
  // Gauss-Legendre rule on [0,1] (deal.II QGauss<1>)
  struct Quad1D
  {
    double x[8];
    double w[8];
  };
  void make_basis_1d(int degree, Basis1D &b);
  void make_gauss_1d(int n, Quad1D &q);

  template <class T>
  struct DevBuf
  {
    T     *p = nullptr;
    size_t n = 0;
    void
    alloc(size_t count)
    {
      release();
      n = count;
      if (count)
        PD_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void
    release()
    {
      if (p)
        cudaFree(p);
      p = nullptr;
      n = 0;
    }
    ~DevBuf()
    {
      release();
    }
    DevBuf()                          = default;
    DevBuf(const DevBuf &)            = delete;
    DevBuf &operator=(const DevBuf &) = delete;
  };
} // namespace pd

struct pd_peer;
struct pd_handle
{
  int     dim = 0, degree = 0, n1 = 0, n = 0; // n = dofs per polytope
  int     fe_kind = PD_FE_DGQ;
  int     nq1 = 0, nq1f = 0, nqc = 0, nqf = 0;
  int64_t n_verts = 0, n_cells = 0, n_subcells = 0, n_subfaces = 0;
  int32_t np = 0, np_own = 0, n_ifaces = 0; // np_own owned polytopes (rows) + ghosts = np
  int64_t Q = 0, Qf = 0; // total volume / face quadrature points
  int64_t n_blocks = 0, nnz = 0, n_dofs = 0;
  int32_t n_vitems = 0;

  pd::Basis1D basis{};
  pd::Quad1D  quad{}, quadf{};

  // descriptor arrays
  pd::DevBuf<double>  verts, bbox, sub_sigma;
  pd::DevBuf<double>  rules; // [32]: cell rule x[8], w[8], face rule x[8], w[8]
  pd::DevBuf<int32_t> cell_verts, subcell_idx, dof_block, ifA, ifB, sub_cell, sub_face, bcol;
  pd::DevBuf<int64_t> subcell_ptr, if_sub_ptr, brow_ptr;
  // derived index data
  pd::DevBuf<int64_t> diag_base, if_baseAB, if_baseBA, vitem_q0, vitem_q1, poly_vitem_ptr, padj_ptr, padj;
  pd::DevBuf<int32_t> row_stride, vitem_poly; // row_stride[b] = (#blocks in block row b) * n
  pd::DevBuf<int32_t> cta_item_ptr;           // volume items of every persistent CTA
  int                 vol_plan_tq = 0, vol_plan_grid = 0; // schedule the item lists were built for
  // quadrature (SoA by coordinate)
  pd::DevBuf<double> vq_x, vq_w, fq_x, fq_n, fq_w;
  // work buffers and result
  pd::DevBuf<double> vol_partial, face_diag, values;
  pd::DevBuf<double> vec_a, vec_b; // staging for pd_vmult_host
  // matrix-free fine-mesh operator (every polytope = one Cartesian cell), pd_finemesh.cu
  bool                mf_ready = false, force_generic_mf = false;
  pd::DevBuf<int32_t> spmv_list_interior, spmv_list_boundary; // sharded: block rows without / with ghost columns
  pd::DevBuf<int32_t> mf_list_interior, mf_list_boundary; // sharded: cells without / with ghost neighbours (Morton order)
  pd::DevBuf<int32_t> mf_seq_all;                         // all owned cells in Morton order (empty: already numbered that way)
  pd::DevBuf<double>  mf_geo, mf_rec, mf_vol, mf_zero; // per (cell, direction) geometry / folded stencil records, cell volumes (pd_finemesh.cu)
  bool                mf_rec_valid = false;
  uint32_t            mf_rec_flags = 0;
  double              mf_rec_coef  = 0.;
  std::vector<double> mf_tab_host;    // 1-D tables Mh, Sh, e0|e1, d0|d1 (passed as kernel parameters)
  // tiled variant (one thread per cell, coefficients staged in shared memory): the tile plans of the three
  // cell sequences (all owned cells, interior list, boundary list) and the premultiplied 1-D tables
  struct FineTiles
  {
    pd::DevBuf<int32_t>  halo_pad; // pipelined kernel (fine::StreamPlan): [n_tiles][stream_rows] halo cell of a row or -1
    pd::DevBuf<uint16_t> noff_stream;
    int32_t              stream_rows = 0, stream_zoff = 0;
    bool                 stream_regular = false; // every tile FINE_TILE cells from a 16-byte boundary
    std::vector<int32_t> h_tile_first, h_tile_base; // host copies (the fused sharded plan concatenates two sequences)
    pd::DevBuf<int32_t>  tile_first, tile_ptr, halo, tile_base; // tile_base: first cell of a tile whose cells are consecutive, else -1
    pd::DevBuf<uint16_t> noff;
    int32_t              n_tiles = 0, max_halo = 0, zoff = 0, n_seq = 0;
    bool                 ok = false;
    bool                 stream_ok = false; // every tile one aligned contiguous run: the pipelined kernel (k_fine_stream) can take it
  };
  // a uniform mesh (all cells the same box, one penalty per direction and face kind): the stencil coefficients are
  // kernel constants and the pipelined kernel needs no per-cell records
  struct FineUniform
  {
    bool   ok = false;
    double h[3] = {1., 1., 1.}, sig_in[3] = {0., 0., 0.}, sig_bd[3][2] = {{0., 0.}, {0., 0.}, {0., 0.}};
  };
  FineUniform         mf_uniform;
  FineTiles           mf_tiles[4]; // all owned cells | interior list | boundary list | interior ++ boundary (fused sharded apply)
  // host copies kept for the fused sharded apply's plan (built when the peer memory is connected, pd_peer.cu)
  std::vector<int32_t> mf_h_nbr, mf_h_inner, mf_h_outer;
  // the fused sharded apply (k_fine_stream over interior ++ boundary tiles, ghost rows read from the peers' export
  // buffers once their epoch flags are up): set by setup_fine_fused
  struct FineFused
  {
    bool                      ok = false;
    int32_t                   first_ghost_tile = 0;
    pd::DevBuf<const double *> ghost_src; // [n_ghost] coefficients of a ghost cell in its owner's export buffer, epoch parity 0
    int64_t                   parity_stride = 0; // doubles between the two epoch copies
    const unsigned long long *epochs = nullptr;
    unsigned long long       *flags = nullptr, *error_word = nullptr;
    const int32_t            *owners = nullptr;
    int                       n_owners = 0;
  };
  FineFused mf_fused;
  std::vector<double> mf_tile_tab_host; // Mh | Mh^-1 Sh | Mh^-1 e0,e1 | Mh^-1 d0,d1 | d0,d1
  bool                mf_stream = true;  // the pipelined kernel wherever it applies (PD_FINE_KERNEL=tile switches it off)
  int                 mf_kernel_last = 0; // PD_FINE_KERNEL_* of the last launch
  int                 mf_kernel = 0;    // 0: tiled where available, 1: line-per-thread (default policy and PD_FINE_KERNEL in pd_finemesh.cu)
  // matrix-free fine-mesh operator on general (Q1-mapped) cells with the mapped basis, pd_mappedfine.cu
  bool                mp_ready = false, mp_geo_valid = false;
  std::vector<double> mp_tab_host; // V | V^T | Dt | e0 e1 | d0 d1
  pd::DevBuf<int32_t> mp_cellv, mp_nbr;
  pd::DevBuf<double>  mp_dt, mp_cgeo, mp_fgeo, mp_sigma, mp_xg, mp_zero;
  pd::DevBuf<double>  mf_vol_partial, mf_face_partial;
  bool                pw_plan_valid = false; // work items of the point-wise kernels (pd_polyapply.cu)
  int32_t             pw_n_items    = 0;
  pd::DevBuf<int32_t> pw_item_poly;
  pd::DevBuf<int64_t> pw_item_q0, pw_item_q1, pw_poly_item_ptr;
  int32_t             pw_n_fitems = 0;
  pd::DevBuf<int32_t> pw_fitem_iface;
  pd::DevBuf<int64_t> pw_fitem_q0, pw_fitem_q1, pw_iface_item_ptr;
  std::vector<int64_t> h_if_sub_ptr; // polytopal matrix-free apply (pd_polyapply.cu)
  pd_coefficients     op_coef{1.0, 0.0}; // operator of the matrix-free apply
  uint32_t            op_flags = PD_ASSEMBLE_ALL;
  // device-resident solvers around vmult (pd_solver.cu)
  pd::DevBuf<double> sv_r, sv_z, sv_p, sv_Ap, sv_dinv, sv_partial, sv_scal;
  pd::DevBuf<unsigned int> sv_ticket; // last-CTA ticket of the fused dot-product kernels
  cudaGraphExec_t    cg_graph_exec   = nullptr;
  int                cg_graph_mode   = -1, cg_graph_jacobi = -1;
  pd_peer           *cg_graph_peer   = nullptr;
  // inverse diagonal of the matrix-free operators (pd_solver.cu)
  pd::DevBuf<int32_t> mfd_colour;
  pd::DevBuf<double>  mfd_dinv;
  int                 mfd_n_colours = 0, mfd_mode = -1;
  bool                mfd_valid = false;
  uint32_t            mfd_flags = 0;
  double              mfd_coef[2] = {0., 0.};
  const double      *cg_graph_x      = nullptr, *cg_graph_b = nullptr;
  int64_t            cg_launches_per_chunk = 0;

  std::vector<int64_t> h_brow_ptr, h_subcell_ptr;
  std::vector<int32_t> h_bcol, h_dof_block, h_ifA, h_ifB;
  bool                 cartesian = false; // every owned sub-cell is an axis-aligned box: tensor assembly path (pd_cartesian.cu)
  pd::DevBuf<int>      cart_flag;
  // bricks of the tensor path (pd_cartesian.cu: build_cartesian_bricks)
  pd::DevBuf<int64_t>  cbk_ptr, fbk_ptr;
  pd::DevBuf<int32_t>  cbk_iv, civ, cpos, fbk_s, fbk_iv, fiv, fpos, fbk_lf, pit_brick, pit_meta, pit_q;
  pd::DevBuf<int64_t>  pit_ptr, pit_diag_end;
  pd::DevBuf<double2>  civ_box, fiv_box;
  pd::DevBuf<double>   fbk_plane, fbk_sigma, cmat, fmat;
  pd::DevBuf<int32_t>  cbk_poly, fbk_iface;
  bool                 brick_mat_valid = false;
  int64_t              n_cell_bricks = 0, n_face_bricks = 0, n_diag_items = 0, n_apply_items = 0;
  bool                 bricks_ready = false;
  std::vector<int32_t> h_subcell_idx, h_sub_cell, h_sub_face;
  int                  last_assembly_path = -1; // 0: DMMA kernels on the agglomerated quadrature, 1: tensor path
  std::vector<double>  h_bbox;          // bounding boxes (reinit tables, pd_reinit.cu)
  pd::DevBuf<double>   reinit_scratch;  // tables of one polytope / face before they go to the caller
  uint64_t             fine_geo_hash = 0; // fine meshes: hash of the arrays the stencil / mapped operators derive from
  // bumped by everything a captured solver graph bakes in (operator terms and coefficients, kernel choice,
  // stream, work buffers, uploads); part of the graph cache key (pd_solver.cu)
  uint64_t             op_generation = 0;
  uint64_t             cg_graph_generation = ~0ull;
  cudaGraphExec_t      cg_graph1_exec = nullptr; // one iteration: the tail below a chunk

  cudaStream_t stream     = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t  ev[5]      = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t aux_stream = nullptr; // the tensor path's off-diagonal kernel runs beside the diagonal one (pd_cartesian.cu)
  cudaEvent_t  ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t  ev_order   = nullptr; // orders the own stream behind a caller stream that cannot capture (pd_solver.cu)
  float        last_ms[4] = {0, 0, 0, 0};
  bool         quad_valid = false, assembled = false;
  int64_t      launches   = 0;
  int          sm_count   = 148;
  int64_t      max_row_len = -1; // longest scalar row (doubles), computed lazily for the SpMV dispatch
};

// level transfer between a handle's polytopal space and a finer space (pd_polyapply.cu)
struct pd_transfer
{
  pd_handle          *coarse     = nullptr;
  int                 kind       = 0; // 0: finer agglomeration level, 1: the mesh cells
  int32_t             n_children = 0;
  int64_t             n_fine_dofs = 0;
  pd::DevBuf<int32_t> parent, child_blk, pc_idx;
  pd::DevBuf<int64_t> pc_ptr;
  pd::DevBuf<double>  partial;
  const double       *child_bbox = nullptr; // kind 0: the finer handle's bounding boxes (device)
};

namespace pd
{
  // pd_geometry.cu
  void launch_quadrature(pd_handle *h);
  // pd_assemble.cu
  void launch_assemble(pd_handle *h, uint32_t flags, const pd_coefficients &coef);
  bool assemble_supported(int dim, int degree, int fe_kind);
  // pd_finemesh.cu
  void setup_fine_operator(pd_handle *h, const pd_mesh_desc &d);
  void launch_fine_operator(pd_handle *h, const double *src, double *dst, bool add, int part = 0);
  // fused sharded apply: plan over interior ++ boundary tiles; ghost_src_host[g] = device address of ghost cell g's
  // coefficients (epoch parity 0) as mapped on this rank; returns whether the fused kernel can be used
  bool setup_fine_fused(pd_handle *h, const double *const *ghost_src_host, int64_t parity_stride, const unsigned long long *epochs,
                        unsigned long long *flags, unsigned long long *error_word, const int32_t *owners_dev, int n_owners);
  bool launch_fine_fused(pd_handle *h, const double *src, double *dst, bool add);
  // pd_mappedfine.cu
  void setup_mapped_operator(pd_handle *h, const pd_mesh_desc &d);
  void launch_mapped_operator(pd_handle *h, const double *src, double *dst, bool add);
  // pd_polyapply.cu
  void launch_poly_apply(pd_handle *h, const double *src, double *dst, bool add);
  void launch_poly_rhs(pd_handle *h, const double *f_vol, const double *g_face, double stiffness, double *rhs);
  void launch_transfer(pd_handle *h, const pd_transfer &t, bool transpose, const double *src, double *dst, bool add);
  void launch_poly_error(pd_handle *h, const double *u, const double *exact, const double *exact_grad, double *out2_dev);
  // pd_solver.cu
  bool   solver_cg(pd_handle *h, int mode, const double *b, double *x, int max_iter, double rel_tol, int jacobi,
                   int *iters_out, double *relres_out, pd_peer *peer = nullptr);
  void   solver_diagonal_inverse(pd_handle *h, int mode, double *dst);
  double solver_lambda_max(pd_handle *h, int mode, int n_iter, pd_peer *peer = nullptr);
  void   solver_chebyshev(pd_handle *h, int mode, int degree, double lambda_max, double smoothing_range, const double *b,
                          double *x, int zero_initial_guess, pd_peer *peer = nullptr);
  // pd_peer.cu (all throw; the extern "C" wrappers live in pd_api.cu)
  pd_peer *peer_create(pd_handle *h, int rank, int world, const int64_t *send_ptr, const int32_t *send_blocks,
                       const int64_t *recv_ptr, const int64_t *remote_offset);
  int      peer_handle_bytes();
  void     peer_export(pd_peer *p, void *handles_out);
  void     peer_connect(pd_peer *p, const void *all_handles);
  void     peer_exchange(pd_peer *p, double *x_full_dev);
  void     peer_vmult(pd_peer *p, int mode, double *x_full_dev, double *dst, bool add);
  void     peer_allreduce(pd_peer *p, double *scal_dev, int dst0, int nk);
  pd_handle *peer_handle(pd_peer *p);
  int      peer_status(pd_peer *p);
  int      peer_fused(pd_peer *p); // tiles of the fused fine-mesh plan, 0: not fused
  void     peer_destroy(pd_peer *p);
  // pd_cartesian.cu
  bool check_axis_aligned(pd_handle *h, bool *bricks_stale = nullptr);
  void setup_cartesian(pd_handle *h, const pd_mesh_desc &d);
  bool cartesian_assembly_selected(const pd_handle *h);
  void launch_assemble_cartesian(pd_handle *h, uint32_t flags, const pd_coefficients &coef);
  bool cartesian_apply_available(const pd_handle *h);
  void launch_cart_apply(pd_handle *h, const double *src, double *dst, bool add);
  // pd_reinit.cu
  int64_t reinit_n_points(const pd_handle *h, int32_t poly);
  int64_t reinit_iface_n_points(const pd_handle *h, int32_t iface);
  void    reinit_polytope(pd_handle *h, int32_t poly, double *values, double *grads, double *jxw, double *points,
                          double *unit_points);
  void    reinit_iface(pd_handle *h, int32_t iface, int side, double *values, double *grads, double *jxw, double *points,
                       double *normals);
  void    fe_evaluate(int fe_kind, int dim, int degree, int64_t n_points, const double *unit_points, double *values,
                      double *grads);
  // pd_vmult.cu
  void launch_spmv(pd_handle *h, const double *src, double *dst, bool add, int part = 0);
  bool spmv_can_split(pd_handle *h);
  void launch_diagonal_inverse(pd_handle *h, double *dst);
} // namespace pd
