/* =============================================================================
 * polydeal_b200.h -- C ABI of the B200-native SIP-DG hot path for agglomerated
 * polytopes (assembly + operator apply).
 *
 * The reference (polyDEAL) has no FFI layer: its seam is C++.
 * The entry points below are what a binding on the reference side would call
 * (INTEGRATION.md shows the C++ shim).  Each one cites the reference interface
 * it replaces (paths relative to /root/reference).
 *
 * Conventions: every function returns an int status (PD_OK = 0, negative =
 * error; pd_last_error() gives the message of the last failure on the calling
 * thread).  No exceptions cross the ABI.  All pointers are caller owned unless
 * stated.  One handle is single-threaded, like the reference's handler
 * (include/agglomeration_handler.h:841-851: reinit* invalidates the previous
 * result); separate handles may be used concurrently.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with
 * PD_ERR_NO_DEVICE.
 *
 * Two layers:
 *   pd_*   device core: takes the flattened agglomeration (SoA arrays) and
 *          runs the sm_100a kernels.
 *   pdh_*  host mirror of AgglomerationHandler / AgglomerationAccessor /
 *          MappingBox / PolyUtils::assemble_dg_matrix that produces that
 *          flattened form with the reference's numbering.
 * ========================================================================== */
#ifndef POLYDEAL_B200_H
#define POLYDEAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_OK 0
#define PD_ERR_INVALID (-1)     /* bad argument / inconsistent descriptor        */
#define PD_ERR_CUDA (-2)        /* a CUDA runtime call or kernel failed          */
#define PD_ERR_UNSUPPORTED (-3) /* (dim, degree, FE) combination has no kernel   */
#define PD_ERR_NO_DEVICE (-4)   /* no CUDA device: there is no CPU fallback      */
#define PD_ERR_STATE (-5)       /* call order violated (e.g. vmult before assemble) */
#define PD_NOT_CONVERGED 1      /* pd_cg_solve*: tolerance not reached within max_iter; the iterate, the iteration
                                   count and the residual are valid (SolverCG throws NoConvergence there)          */

#define PD_INVALID_UINT 0xFFFFFFFFu /* numbers::invalid_unsigned_int */

typedef struct pd_handle pd_handle;     /* device-resident flattened agglomeration */
typedef struct pdh_grid pdh_grid;       /* background hypercube mesh               */
typedef struct pdh_handler pdh_handler; /* host mirror of AgglomerationHandler<dim> */

const char *pd_last_error(void);
/* number of CUDA devices visible (0 => every compute call fails loudly) */
int pd_device_count(void);

/* -----------------------------------------------------------------------------
 * Flattened agglomeration (what `AgglomerationHandler` + `PolytopeCache` hold,
 * include/agglomeration_handler.h:600-851, as SoA arrays).  Host pointers.
 * -------------------------------------------------------------------------- */
typedef struct pd_mesh_desc
{
  int32_t dim;       /* 2 or 3 */
  int32_t fe_degree; /* p of the element on the bounding box, see fe_kind at the end */
  int32_t n_q1d;      /* QGauss<dim>(n_q1d) on every sub-cell    (initialize_fe_values,
                         source/agglomeration_handler.cc:212-236) */
  int32_t n_q1d_face; /* QGauss<dim-1>(n_q1d_face) on every sub-face */

  /* background mesh, deal.II conventions (vertices lexicographic in a cell) */
  int64_t        n_verts;
  const double  *verts;      /* [n_verts][dim] */
  int64_t        n_cells;
  const int32_t *cell_verts; /* [n_cells][2^dim] */

  /* polytopes, in define_agglomerate order (= polytope->index()) */
  int32_t        n_polytopes;
  const int64_t *poly_subcell_ptr; /* [n_polytopes+1] CSR                              */
  const int32_t *poly_subcell_idx; /* sub-cells, slaves first, master last
                                      (include/agglomeration_accessor.h:562-569)      */
  const double  *bbox;             /* [n_polytopes][2*dim]: lo[dim] then hi[dim]
                                      (source/agglomeration_handler.cc:476-491)       */
  const int32_t *dof_block;        /* [n_polytopes] first global DoF / n_dofs_per_cell:
                                      DoF blocks follow the master's active-cell index */

  /* polytope faces as a work list.  One entry per boundary face of a polytope
   * (polyB = -1) and one per interior interface, listed from the VISITING side A
   * (include/poly_utils.h:2089).  Sub-faces are (cell, local face) pairs of A's
   * sub-cells in the reference's interface order
   * (source/agglomeration_handler.cc:1375-1397,1442-1463,1606-1609). */
  int32_t        n_ifaces;
  const int32_t *iface_polyA;   /* [n_ifaces] */
  const int32_t *iface_polyB;   /* [n_ifaces], -1 on the boundary */
  const int64_t *iface_sub_ptr; /* [n_ifaces+1] */
  const int32_t *sub_cell;      /* [n_subfaces] */
  const int32_t *sub_face;      /* [n_subfaces] local face 0..2*dim-1 */
  const double  *sub_sigma;     /* [n_subfaces] penalty (C/h as the caller's rule says) */

  /* block-CSR pattern = create_agglomeration_sparsity_pattern
   * (source/agglomeration_handler.cc:910-1022) in units of n x n blocks, block
   * columns ascending.  Scalar CSR follows: row b*n+i holds, for every block k of
   * block row b in order, columns bcol[k]*n + (0..n-1). */
  int32_t        n_block_rows;
  const int64_t *brow_ptr; /* [n_block_rows+1] */
  const int32_t *bcol_idx; /* [n_blocks] */

  /* Sharding (one rank of a partition by polytopes; an agglomerate never straddles ranks,
   * source/agglomeration_handler.cc:83-87).  0 = everything is owned.  Otherwise polytopes
   * [0, n_owned) are owned and [n_owned, n_polytopes) are GHOSTS: no sub-cells, only bbox and
   * dof_block (>= n_owned), the role of recv_ghosted_bbox / recv_ghost_dofs
   * (include/agglomeration_handler.h:525-548).  Rows exist for owned polytopes only
   * (n_block_rows = n_owned); column block indices run over owned + ghost.  An interface
   * (A owned, B ghost) is listed from the OWNED side whatever the visiting rule says (the SIP
   * form is symmetric under swapping sides with the normal flipped) with the sigma of the rule. */
  int32_t n_owned_polytopes;

  /* element family on the bounding box (source/agglomeration_handler.cc:331-337):
   *  PD_FE_DGQ      FE_DGQ<dim>(p): tensor Lagrange basis on the Gauss-Lobatto nodes, (p+1)^dim DoFs
   *  PD_FE_AGGLODGP FE_AggloDGP<dim>(p) (source/fe_agglodgp.cc:28-57): products of L2[0,1]-
   *                 orthonormal Legendre polynomials of total degree <= p, C(p+dim, dim) DoFs,
   *                 last coordinate outermost / first fastest.  Assembly, block-CSR apply, the
   *                 polytopal matrix-free apply, right-hand side and error norms; the fine-mesh
   *                 operators and the level transfers are FE_DGQ notions (nodal support points).
   * Last member, so that zero-initialised descriptors of older callers mean FE_DGQ. */
  int32_t fe_kind;
} pd_mesh_desc;

typedef struct pd_coefficients
{
  double stiffness; /* sigma: scales grad-grad AND all face terms
                       (include/utils.h:1628-1636)                     */
  double mass;      /* f: adds f * phi_i phi_j  (reaction c / chi*C_m/dt,
                       examples/diffusion_reaction.cc:489-506)          */
} pd_coefficients;

/* assemble flags */
#define PD_ASSEMBLE_VOLUME 1u   /* include/poly_utils.h:2038-2052 */
#define PD_ASSEMBLE_BOUNDARY 2u /* include/poly_utils.h:2060-2085 */
#define PD_ASSEMBLE_INTERIOR 4u /* include/poly_utils.h:1870-1926, 2086-2132 */
#define PD_ASSEMBLE_ALL 7u

/* vmult modes */
#define PD_FE_DGQ 0
#define PD_FE_AGGLODGP 1

#define PD_VMULT_BLOCK_CSR 0   /* y = A x with the assembled matrix */
#define PD_VMULT_MATRIX_FREE 1 /* y = A x recomputed from the quadrature data           */
#define PD_VMULT_MAPPED_FINE 2 /* fine mesh of general hexes, mapped FE_DGQ basis (below)  */

/* Create the device-resident copy.  Replaces the state built by
 * AgglomerationHandler::define_agglomerate / distribute_agglomerated_dofs /
 * setup_connectivity_of_agglomeration (source/agglomeration_handler.cc:45-104,
 * 326-379, 495-527).  The descriptor is copied; it may be freed afterwards. */
int pd_create(const pd_mesh_desc *desc, pd_handle **out);
int pd_destroy(pd_handle *h);

/* Re-send every descriptor array host->device into the existing buffers.  This is the
 * per-step host->device leg of the end-to-end path (a mesh whose vertices move, new
 * penalties).  The TOPOLOGY must be the one of pd_create: sizes, poly_subcell_ptr,
 * dof_block, the interface list (iface_polyA/B, iface_sub_ptr) and the block pattern are
 * compared with the handle's copies and a difference is PD_ERR_INVALID (create a new
 * handle).  verts, bbox, sub_sigma and the range-checked index arrays cell_verts,
 * poly_subcell_idx, sub_cell, sub_face may change: the quadrature, the assembled matrix
 * and the cached inverse diagonals are invalidated, and on fine meshes the stencil /
 * mapped-geometry tables of the matrix-free operators are re-derived when they differ. */
int pd_upload(pd_handle *h, const pd_mesh_desc *desc);

/* Stream contract: every device call of a handle is enqueued on ONE stream and nothing else
 * orders it against the caller's work.  By default that is a stream the handle creates
 * (cudaStreamNonBlocking: NOT ordered with the legacy default stream); a caller that
 * produces src / consumes dst on its own stream must either hand that stream over here or
 * synchronise itself.  cuda_stream: a cudaStream_t passed as void*; NULL = back to the
 * handle's own stream; the legacy default stream is named by cudaStreamLegacy
 * ((void*)0x1), the per-thread default stream by cudaStreamPerThread ((void*)0x2).
 * Synchronises the previous stream; invalidates captured solver graphs. */
int pd_set_stream(pd_handle *h, void *cuda_stream);
int pd_synchronize(pd_handle *h);

/* agglomerated_quadrature + the geometry half of reinit_master on the device
 * (source/agglomeration_handler.cc:622-707, 1139-1165): fills the volume
 * q-points/JxW of every polytope and the face q-points/normals/JxW of every
 * interface.  Called implicitly by pd_assemble when stale. */
int pd_build_quadrature(pd_handle *h);
/* the agglomerated quadrature, structure of arrays: vol_x [dim][Q], vol_jxw [Q] with the points
 * of polytope p at poly_subcell_ptr[p] * n_q^dim ...; face_x / face_n [dim][Qf], face_jxw [Qf]
 * with the points of sub-face s at s * n_q^(dim-1) ...  (built on demand).  Callers evaluate
 * their data (right-hand side f, Dirichlet values g, exact solution) at these points. */
int64_t pd_n_quadrature_points(const pd_handle *h, int faces);
int pd_quadrature_device(pd_handle *h, const double **vol_x, const double **vol_jxw, const double **face_x,
                         const double **face_n, const double **face_jxw);
int pd_quadrature_to_host(pd_handle *h, double *vol_x, double *vol_jxw, double *face_x, double *face_n,
                          double *face_jxw);
/* --- the reinit() family: the FEValues tables of ONE polytope / polytope face -------------------
 * What AgglomerationHandler::reinit(polytope), reinit(polytope, f) and reinit_interface return
 * (include/agglomeration_handler.h:431-452, source/agglomeration_handler.cc:729-906): the element on
 * the bounding box (MappingBox, source/mapping_box.cc:393-439: value phihat(xhat), gradient
 * grad-hat / h_bbox) at the agglomerated quadrature.  The assembly kernels generate these tables on
 * the fly; these entry points materialise them so that the reference's hand-written loops
 * (examples/poisson.cc:745-905) can be pointed at the library.  Output pointers are HOST or DEVICE
 * memory (any may be NULL); the tables are complete when the call returns, and -- as in the
 * reference, where reinit* invalidates the previous FEValues -- one call at a time per handle.
 * Layout = the FEValues accessors, Q = number of points of the item, n = pd_n_dofs_per_cell:
 *   values[i * Q + q] = shape_value(i, q)       grads[(i * Q + q) * dim + d] = shape_grad(i, q)[d]
 *   jxw[q] = JxW(q)    points[q * dim + d] = quadrature_point(q)[d]    normals[q * dim + d] = normal_vector(q)[d]
 * Faces are addressed through the flattened work list: interface `iface` (pdh_face_work_item maps a
 * polytope's face number to it), side 0 = the listing polytope iface_polyA with its outward normal,
 * side 1 = iface_polyB at the SAME (aligned) points with its own outward normal = -n_A and its own
 * bounding box -- the pair reinit_interface returns. */
int64_t pd_reinit_n_points(const pd_handle *h, int32_t poly);        /* Q of an owned polytope, < 0: error */
int64_t pd_reinit_iface_n_points(const pd_handle *h, int32_t iface); /* Q of an interface / boundary face */
int pd_reinit_polytope(pd_handle *h, int32_t poly, double *values, double *grads, double *jxw, double *points);
int pd_reinit_face(pd_handle *h, int32_t iface, int32_t side, double *values, double *grads, double *jxw,
                   double *points, double *normals);
int pd_reinit_interface(pd_handle *h, int32_t iface, double *values0, double *grads0, double *values1, double *grads1,
                        double *jxw, double *points, double *normals);
/* AgglomerationHandler::agglomerated_quadrature (source/agglomeration_handler.cc:622-707) of one polytope:
 * the Quadrature<dim> the reference builds holds the points in bounding-box UNIT coordinates and the
 * physical JxW as weights; real_points are the same points before BoundingBox::real_to_unit. */
int pd_agglomerated_quadrature(pd_handle *h, int32_t poly, double *unit_points, double *jxw, double *real_points);
/* The element families on the unit cell at arbitrary points (unit_points [n_points][dim], host or device):
 * FE_DGQ<dim>(degree) / FE_AggloDGP<dim>(degree) shape values [n][n_points] and gradients [n][n_points][dim]
 * (source/fe_agglodgp.cc:28-57; deal.II FE_DGQ).  Needs a CUDA device like every compute entry point. */
int pd_fe_evaluate(int32_t fe_kind, int32_t dim, int32_t degree, int64_t n_points, const double *unit_points,
                   double *values, double *grads);

/* Right-hand side over polytopes (examples/poisson.cc:745-761, examples/diffusion_reaction.cc:
 * 550-556): rhs_i = sum_q f_q phi_i w_q  +  stiffness * sum_{boundary q} (sigma g_q phi_i -
 * (grad phi_i . n) g_q) w_q, sigma the boundary sub-face penalty of the descriptor.
 * f_vol_dev [Q] and g_face_dev [Qf] (only boundary points are read) are device arrays of
 * values at the quadrature points; either may be NULL.  rhs_dev: pd_n_dofs doubles. */
int pd_assemble_rhs(pd_handle *h, const double *f_vol_dev, const double *g_face_dev, double stiffness,
                    double *rhs_dev);
/* PolyUtils::compute_global_error (include/poly_utils.h:1647-1750): L2 norm and H1 seminorm of
 * u_h - u over the agglomerated quadrature; exact_dev [Q], exact_grad_dev [dim][Q] (may be NULL
 * together with h1_seminorm).  Owned polytopes only: a sharded caller sums the squares. */
int pd_error_norms(pd_handle *h, const double *u_dev, const double *exact_dev, const double *exact_grad_dev,
                   double *l2, double *h1_seminorm);
/* --- level transfers (the step either side of vmult in the V-cycle) ---------------------------
 * P evaluates a polytope's function at the FE_DGQ support points of a child element,
 * local_matrix(i, j) = phi^parent_j(p_i), applied on the fly (no transfer matrix is stored):
 *  pd_transfer_create          child = polytope of a FINER agglomeration level (another handle on
 *                              the same mesh and FE), points = the nodes of the child's bounding
 *                              box: Utils::fill_injection_matrix (include/utils.h:95-270).
 *                              parent_of_fine[q] = the coarse polytope that contains fine
 *                              polytope q (polytope->children(), inverted).
 *  pd_transfer_create_to_cells child = every mesh cell of a polytope, points = its Q1-mapped nodes:
 *                              PolyUtils::fill_interpolation_matrix (include/poly_utils.h:1469-1634);
 *                              fine DoFs = n per cell in active-cell order.
 *  prolongate  y_fine (+)= P x_coarse      (MGTransferAgglomeration::prolongate[_and_add])
 *  restrict    x_coarse (+)= P^T y_fine    (restrict_and_add, source/multigrid_amg.cc:66-110)
 * Both handles must outlive the transfer.  Vectors are device pointers. */
typedef struct pd_transfer pd_transfer;
int pd_transfer_create(pd_handle *coarse, pd_handle *fine, const int32_t *parent_of_fine, pd_transfer **out);
int pd_transfer_create_to_cells(pd_handle *h, pd_transfer **out);
void pd_transfer_destroy(pd_transfer *t);
int64_t pd_transfer_m(const pd_transfer *t); /* fine DoFs (rows of P) */
int64_t pd_transfer_n(const pd_transfer *t); /* coarse DoFs */
int pd_transfer_prolongate(pd_transfer *t, const double *src_coarse_dev, double *dst_fine_dev, int add);
int pd_transfer_restrict(pd_transfer *t, const double *src_fine_dev, double *dst_coarse_dev, int add);

/* --- ghost exchange of the sharded vmult over NVLink peer memory ------------------------------
 * One process per GPU on one node.  Instead of a message per apply (update_ghost_values() inside
 * MatrixFree::loop, include/utils.h:466-472), every rank publishes the blocks its neighbours need
 * into a buffer they have mapped through CUDA IPC and pulls its ghost blocks with loads over
 * NVLink; ordering is an epoch-flag handshake in the same memory, double buffered (pd_peer.cu).
 * Plan (all in polytope blocks): send_ptr[world+1] / send_blocks = my owned blocks needed by
 * each peer; recv_ptr[world+1] = my ghost section grouped by owner rank; remote_offset[s] = where
 * my segment starts in rank s's send list (= s's send_ptr[my rank]).
 *   pd_peer_create -> pd_peer_export (pd_peer_handle_bytes() bytes) -> all-gather the handles by
 *   any means (once) -> pd_peer_connect([world][bytes]) -> pd_peer_exchange(x_full) before every
 *   pd_vmult: fills the ghost section of x_full (length pd_n_source_dofs) on the handle's stream.
 * pd_peer_status != PD_OK after a synchronize means a neighbour did not arrive within ~2 s. */
typedef struct pd_peer pd_peer;
int pd_peer_create(pd_handle *h, int rank, int world, const int64_t *send_ptr, const int32_t *send_blocks,
                   const int64_t *recv_ptr, const int64_t *remote_offset, pd_peer **out);
int pd_peer_handle_bytes(void);
int pd_peer_export(pd_peer *p, void *handles_out);
int pd_peer_connect(pd_peer *p, const void *all_handles);
int pd_peer_exchange(pd_peer *p, double *x_full_dev);
/* pd_peer_exchange + pd_vmult in one call; with the fine-mesh stencil kernel the cells that need no
 * ghost data are applied while the ghost blocks travel on a second stream */
int pd_peer_vmult(pd_peer *p, int mode, double *x_full_dev, double *dst_dev, int add);
int pd_peer_status(pd_peer *p);
/* > 0 when pd_peer_vmult(PD_VMULT_MATRIX_FREE) runs the FUSED fine-mesh apply (uniform fine mesh whose cells are numbered
 * along the Morton curve on every rank): after the publish ONE kernel applies all owned cells; the tiles next to a cut
 * come last, wait for the owners' epoch flags and read the ghost cells straight from the owners' export buffers over
 * NVLink.  The ghost section of x_full_dev is not written on that path.  The value is the number of tiles of that
 * kernel's plan (blocks of the curve; a block whose halo would not fit the kernel's gather is split); 0: not fused. */
int pd_peer_fused(pd_peer *p);
/* sum of `count` <= 4 doubles (device memory, in place) over all ranks: one warp, stores into every
 * rank's mapped buffer + flag handshake; the result is bitwise identical on all ranks */
int pd_peer_allreduce(pd_peer *p, double *scalars_dev, int count);
/* pd_cg_solve on a sharded handle (collective call): b_dev / x_dev hold the OWNED DoFs; the
 * ghost exchange of every apply and the all-reduce of every dot product run over peer memory
 * inside the replayed CUDA graph (MPI_Allreduce + update_ghost_values in the reference's
 * SolverCG on LinearAlgebra::distributed::Vector).  Residual norms are global. */
int pd_cg_solve_sharded(pd_peer *p, int mode, const double *b_dev, double *x_dev, int max_iter, double rel_tol,
                        int jacobi, int *iterations, double *relative_residual);
/* pd_estimate_lambda_max / pd_chebyshev_smooth on a sharded handle (collective calls).  The
 * smoother's x_full_dev is a vmult source: pd_n_source_dofs long, owned DoFs first; b_dev owned. */
int pd_estimate_lambda_max_sharded(pd_peer *p, int mode, int n_iterations, double *lambda_max);
int pd_chebyshev_smooth_sharded(pd_peer *p, int mode, int degree, double lambda_max, double smoothing_range,
                                const double *b_dev, double *x_full_dev, int zero_initial_guess);
void pd_peer_destroy(pd_peer *p);

/* mark the device quadrature stale (vertices changed through pd_upload do this
 * implicitly): the next pd_assemble rebuilds it */
int pd_invalidate_quadrature(pd_handle *h);

/* PolyUtils::assemble_dg_matrix (include/poly_utils.h:2000-2195).  Result: the
 * scalar-CSR value array of the reference pattern (ascending columns), kept on
 * the device inside the handle.
 * Two kernel families compute the same matrix (the parity tests hold both to the CPU restatement of the reference):
 *   PD_PATH_TENSOR  every owned sub-cell is an axis-aligned box (checked on the device at
 *                   pd_create / pd_upload): the quadrature sums factorise per sub-cell / sub-face
 *                   into 1-D matrices and a block is a sum of Kronecker products (pd_cartesian.cu)
 *   PD_PATH_DMMA    any mesh: rank-k updates over the agglomerated quadrature points on the FP64
 *                   tensor cores (pd_assemble.cu)
 * The environment variable PD_ASSEMBLE_KERNELS=generic forces PD_PATH_DMMA (measurements, tests). */
int pd_assemble(pd_handle *h, uint32_t flags, const pd_coefficients *coef);
#define PD_PATH_DMMA 0
#define PD_PATH_TENSOR 1
/* which family the last pd_assemble ran (-1: none yet) */
int pd_assembly_path(const pd_handle *h);
/* the tensor path's work: stats5 = {axis-aligned (0/1), cell bricks, face bricks, items of all diagonal blocks,
 * items of the matrix-free apply} (a brick = a tensor-product set of sub-cells / coplanar sub-faces whose quadrature
 * sums factorise, pd_cartesian.cu) */
int pd_tensor_path_stats(const pd_handle *h, int64_t *stats5);

int64_t pd_n_dofs(const pd_handle *h);        /* rows = owned DoFs */
int64_t pd_n_source_dofs(const pd_handle *h); /* length of vmult source vectors = owned + ghost DoFs */
int64_t pd_nnz(const pd_handle *h);
int32_t pd_n_dofs_per_cell(const pd_handle *h);
/* device pointer to the matrix values (length pd_nnz) */
int pd_matrix_values_device(pd_handle *h, double **dev_values);
/* device->host copy of the matrix values (the per-step device->host leg) */
int pd_matrix_values_to_host(pd_handle *h, double *host_values);
/* the same without the final stream synchronisation (pinned host memory; complete after
 * pd_synchronize): lets a caller that assembles many matrices double-buffer two handles on two
 * streams, so that the download of one overlaps the upload and the kernels of the next */
int pd_matrix_values_to_host_async(pd_handle *h, double *host_values_pinned);
/* scalar CSR pattern of the result: rowptr[n_dofs+1], cols[nnz] (host) */
int pd_matrix_pattern_to_host(pd_handle *h, int64_t *rowptr, int32_t *cols);

/* vmult: what LinearOperatorMG::vmult / TrilinosWrappers::SparseMatrix::vmult do
 * on agglomerated levels (include/multigrid_amg.h:345-355,
 * include/linear_operator_for_mg.h:295).  src/dst are DEVICE pointers of
 * pd_n_dofs doubles.  vmult_add accumulates into dst. */
int pd_vmult(pd_handle *h, int mode, const double *src_dev, double *dst_dev);
/* Operator of the MATRIX_FREE apply: which terms (PD_ASSEMBLE_* flags) and the
 * coefficients, e.g. flags = VOLUME|INTERIOR, {sigma, chi*Cm/dt} for
 * MonodomainOperatorDG (include/utils.h:1131-1134, 1565-1659: no boundary term).
 * Default: all terms, {1, 0} = LaplaceOperatorDG (include/utils.h:819-925). */
int pd_set_operator(pd_handle *h, uint32_t flags, const pd_coefficients *coef);
/* PD_VMULT_MATRIX_FREE picks one of two kernels:
 *  - every polytope is one axis-aligned cell (the reference's fine-mesh MatrixFree case,
 *    pd_matrix_free_available() == 1): the sum-factorised 1-D stencil kernel;
 *  - genuine agglomerates: the basis is regenerated at the agglomerated quadrature points
 *    (the same operand rows the assembly contracts) and applied to the polytope's coefficients --
 *    no matrix memory, ~12 n flops per volume point; the block-CSR apply is the faster one
 *    whenever the matrix fits.
 * pd_force_generic_matrix_free(h, 1) selects the second kernel on fine meshes too (testing). */
int pd_matrix_free_available(const pd_handle *h);
/* which fine-mesh kernel the last PD_VMULT_MATRIX_FREE apply on a fine mesh launched (diagnostic; the tests use it to
 * make sure the kernel they mean to check is the one that ran): 0 none yet, 1 line per thread (k_fine_sip), 2 tiled
 * (k_fine_tile), 3 pipelined tiles on a uniform mesh (k_fine_stream) */
#define PD_FINE_KERNEL_LINE 1
#define PD_FINE_KERNEL_TILE 2
#define PD_FINE_KERNEL_STREAM 3
int pd_fine_kernel_last(const pd_handle *h);
int pd_force_generic_matrix_free(pd_handle *h, int on);
/* PD_VMULT_MAPPED_FINE: the reference's fine-mesh MatrixFree operators on GENERAL (Q1-mapped,
 * distorted) cells -- LaplaceOperatorDG / MonodomainOperatorDG with the standard mapped
 * FE_DGQ(p) basis, n_q_points_1d = p+1 and the penalty of include/utils.h:861-866, 906-909
 * (max(p,1)(p+1) (|n J_m^-1| + |n J_p^-1|) at face point 0; 4 max(p,1)(p+1) |n J^-1| on the
 * boundary), computed here from the cell vertices: the descriptor's sub_sigma / bbox are NOT
 * used, and on distorted cells this is a different operator from the bounding-box-basis one
 * that pd_assemble / PD_VMULT_MATRIX_FREE apply (they coincide on Cartesian cells with the
 * normal-extent penalty rule).  Terms and coefficients come from pd_set_operator.
 * Available (== 1) when every polytope is one cell, the handle has no ghost polytopes, the
 * cell and face rules are QGauss(p+1) and neighbouring cells are in standard orientation. */
int pd_mapped_fine_available(const pd_handle *h);
int pd_vmult_add(pd_handle *h, int mode, const double *src_dev, double *dst_dev);
/* same with HOST buffers (pinned or pageable): H2D, apply, D2H */
int pd_vmult_host(pd_handle *h, int mode, const double *src_host, double *dst_host);
/* inverse of the matrix diagonal, entries below 1e-10 kept as they are
 * (include/utils.h:797-814); device pointer of pd_n_dofs doubles */
int pd_diagonal_inverse(pd_handle *h, double *dst_dev);
/* the same for the operator a vmult mode applies.  Matrix-free modes need no assembled matrix: the
 * operator is applied to sums of unit vectors over independent sets of polytopes (greedy colouring of
 * the block pattern), n_dofs_per_cell * n_colours applies, cached until pd_set_operator changes
 * (MatrixFreeTools::compute_diagonal in the reference, include/utils.h:929-1100).  pd_cg_solve with
 * jacobi != 0, pd_chebyshev_smooth and pd_estimate_lambda_max use it for their mode. */
int pd_diagonal_inverse_of(pd_handle *h, int mode, double *dst_dev);

/* --- the immediate callers of vmult, device resident (single-rank handles) ------------------
 * Preconditioned conjugate gradients around pd_vmult: what SolverCG does in
 * examples/diffusion_reaction.cc:709-724 / examples/matrix_free_agglo.cc:377-384.
 * jacobi != 0 preconditions with the inverse diagonal (needs pd_assemble).  x_dev holds the
 * initial guess on entry.  Stops when |r| <= rel_tol |b| (checked every 8 iterations: the
 * iteration body is replayed from a CUDA graph; *iterations is the number actually run,
 * never more than max_iter) and returns PD_OK, or returns PD_NOT_CONVERGED after max_iter.
 * rel_tol <= 0: exactly max_iter iterations without a convergence test (PD_OK).
 * The captured graphs are keyed on everything they bake in (mode, jacobi, x_dev, b_dev,
 * operator terms / coefficients, stream, uploads), so changing any of it between two
 * solves is safe. */
int pd_cg_solve(pd_handle *h, int mode, const double *b_dev, double *x_dev, int max_iter, double rel_tol,
                int jacobi, int *iterations, double *relative_residual);
/* largest eigenvalue of D^-1 A by n_iterations of the power method (the bound
 * PreconditionChebyshev needs; deal.II estimates it with eig_cg_n_iterations CG steps) */
int pd_estimate_lambda_max(pd_handle *h, int mode, int n_iterations, double *lambda_max);
/* PreconditionChebyshev with a Jacobi inner preconditioner as used for the multigrid smoothers
 * (examples/matrix_free_agglo.cc:264-319: degree 3, smoothing_range 20): `degree` steps of the
 * Chebyshev recurrence on [lambda_max/smoothing_range, lambda_max]; degree - 1 operator applies
 * when zero_initial_guess != 0. */
int pd_chebyshev_smooth(pd_handle *h, int mode, int degree, double lambda_max, double smoothing_range,
                        const double *b_dev, double *x_dev, int zero_initial_guess);

/* Debug/parity access to device-resident arrays.  name in: "vol_qpt" [dim][Q],
 * "vol_jxw" [Q], "face_qpt" [dim][Qf], "face_normal" [dim][Qf], "face_jxw" [Qf].
 * Returns the element count through *count when host_out is NULL. */
int pd_copy_array(pd_handle *h, const char *name, double *host_out, int64_t *count);

/* number of kernel launches issued through this handle so far */
int64_t pd_launch_count(const pd_handle *h);
/* device time (ms) of the dominant assembly kernels of the LAST pd_assemble,
 * measured with CUDA events on the handle's stream: [0] volume, [1] faces,
 * [2] reduce, [3] quadrature */
int pd_last_kernel_ms(pd_handle *h, float *ms4);

/* The 1-D tables the kernels use: QGauss<1>(n) on [0,1] and the Gauss-Lobatto
 * support points of FE_DGQ(degree).  Host-only; lets a CPU test-suite check them. */
int pd_quadrature_rule_1d(int n, double *x, double *w);
int pd_dgq_nodes_1d(int degree, double *nodes);

/* -----------------------------------------------------------------------------
 * Host mirror of the reference classes
 * -------------------------------------------------------------------------- */
/* GridGenerator::hyper_cube + refine_global (order 0: hierarchical = Morton cell
 * order, needs n = 2^k) or subdivided_hyper_rectangle (order 1: lexicographic) */
int pdh_grid_create_structured(int32_t dim, const int32_t *n, const double *lo, const double *hi,
                               int32_t order, pdh_grid **out);
/* any hypercube mesh in deal.II conventions; nbr[n_cells][2*dim] = neighbouring
 * cell behind each local face (-1 on the boundary), standard orientation */
int pdh_grid_create(int32_t dim, int64_t n_verts, const double *verts, int64_t n_cells,
                    const int32_t *cell_verts, const int32_t *nbr, pdh_grid **out);
int pdh_grid_destroy(pdh_grid *g);
int64_t pdh_grid_n_cells(const pdh_grid *g);
int64_t pdh_grid_n_verts(const pdh_grid *g);
/* overwrite vertex coordinates (e.g. GridTools::distort_random done by the caller) */
int pdh_grid_set_vertices(pdh_grid *g, const double *verts);
int pdh_grid_get_arrays(const pdh_grid *g, double *verts, int32_t *cell_verts, int32_t *nbr);

/* AgglomerationHandler<dim>(cached_tria)  (source/agglomeration_handler.cc:20-40) */
int pdh_handler_create(pdh_grid *g, pdh_handler **out);
int pdh_handler_destroy(pdh_handler *ah);
/* define_agglomerate(cells): cells[0] is the master; returns polytope index or <0
 * (source/agglomeration_handler.cc:45-104) */
int32_t pdh_define_agglomerate(pdh_handler *ah, const int32_t *cells, int32_t n);
/* the same for many agglomerates in one call: group g = cells[ptr[g] .. ptr[g+1]) (define order = group order) */
int pdh_define_agglomerates(pdh_handler *ah, int32_t n_groups, const int64_t *ptr, const int32_t *cells);
/* initialize_fe_values(QGauss<dim>(nq_cell), ..., QGauss<dim-1>(nq_face)) (:210-236) */
int pdh_initialize_fe_values(pdh_handler *ah, int32_t nq_cell, int32_t nq_face);
/* distribute_agglomerated_dofs(fe)  (:326-379); fe_kind PD_FE_DGQ or PD_FE_AGGLODGP */
int pdh_distribute_agglomerated_dofs(pdh_handler *ah, int32_t fe_kind, int32_t degree);

int32_t  pdh_n_polytopes(const pdh_handler *ah);
int64_t  pdh_n_dofs(const pdh_handler *ah);
int32_t  pdh_n_dofs_per_cell(const pdh_handler *ah);
/* AgglomerationAccessor (include/agglomeration_accessor.h:41-299) by polytope index */
int32_t  pdh_master_cell(const pdh_handler *ah, int32_t poly);
int32_t  pdh_n_background_cells(const pdh_handler *ah, int32_t poly);
int      pdh_get_agglomerate(const pdh_handler *ah, int32_t poly, int32_t *cells);
uint32_t pdh_n_faces(const pdh_handler *ah, int32_t poly);
int32_t  pdh_at_boundary(const pdh_handler *ah, int32_t poly, uint32_t f);
int32_t  pdh_neighbor(const pdh_handler *ah, int32_t poly, uint32_t f); /* -1: boundary */
uint32_t pdh_neighbor_of_agglomerated_neighbor(const pdh_handler *ah, int32_t poly, uint32_t f);
/* get_interface().at({id, neighbour id}): fills (cell, local face) pairs, returns count */
int32_t  pdh_interface(const pdh_handler *ah, int32_t poly, uint32_t f, int32_t *cells,
                       int32_t *faces, int32_t cap);
int      pdh_get_dof_indices(const pdh_handler *ah, int32_t poly, uint32_t *dofs);
int      pdh_bounding_box(const pdh_handler *ah, int32_t poly, double *lo, double *hi);
double   pdh_diameter(const pdh_handler *ah, int32_t poly);
double   pdh_volume(const pdh_handler *ah, int32_t poly);
/* create_agglomeration_sparsity_pattern (:910-1022): scalar CSR, ascending columns */
int64_t  pdh_sparsity_nnz(const pdh_handler *ah);
int      pdh_create_agglomeration_sparsity_pattern(const pdh_handler *ah, int64_t *rowptr, int32_t *cols);

/* penalty rules found in the reference */
#define PD_H_DIAMETER_OF_VISITOR 0  /* C / polytope->diameter()       include/poly_utils.h:2057   */
#define PD_H_MAX_INVERSE_DIAMETER 1 /* C max(1/hA,1/hB)  test/polydeal/poisson_sanity_check_01.cc:261 */
#define PD_H_CONSTANT 2             /* C / h_const       test/polydeal/minimal_SIP_Poisson.cc:308 */
#define PD_H_NORMAL_EXTENT 3        /* C (1/hn_A + 1/hn_B), boundary 4C/hn  include/utils.h:861-866,906-909 */
#define PD_VISIT_BY_ID 0            /* polytope->id() < neighbour->id()        include/poly_utils.h:2089 */
#define PD_VISIT_BY_INDEX 1         /* polytope->index() < neighbour->index()  examples/poisson.cc:841   */

typedef struct pdh_flatten_params
{
  double  penalty_constant; /* C; <0 => library default 10 (p+dim)(p+1), include/poly_utils.h:2018 */
  int32_t h_rule;
  double  h_const;
  int32_t visit_rule;
} pdh_flatten_params;

/* Flatten the agglomeration into a descriptor whose arrays stay owned by the handler: they are valid until the next
 * pdh_flatten / pdh_flatten_local on the same handler (which reuses the storage) or its destruction.  pd_create copies
 * everything it needs to the device, so a descriptor need not outlive that call. */
int pdh_flatten(pdh_handler *ah, const pdh_flatten_params *prm, pd_mesh_desc *out);
/* Where face f of polytope `poly` (the reference's face numbering, n_faces / neighbor / at_boundary) sits in
 * the work list of the LAST pdh_flatten: *iface = its entry, *side = 0 if `poly` is the listing (visiting)
 * polytope iface_polyA, 1 if it is iface_polyB.  Feeds pd_reinit_face / pd_reinit_interface. */
int pdh_face_work_item(const pdh_handler *ah, int32_t poly, uint32_t f, int32_t *iface, int32_t *side);
/* MappingBox / BoundingBox of a polytope (source/mapping_box.cc:923-972, agglomeration_handler.cc:698-704):
 * xhat = (x - lo) / (hi - lo) and back, n points [n][dim]. */
int pdh_real_to_unit(const pdh_handler *ah, int32_t poly, int64_t n, const double *real_points, double *unit_points);
int pdh_unit_to_real(const pdh_handler *ah, int32_t poly, int64_t n, const double *unit_points, double *real_points);
/* One rank's share of a partition of the polytopes (owner[p] = rank of polytope p, all
 * polytopes of the handler): owned polytopes + the ghost polytopes adjacent to them, as a
 * descriptor with n_owned_polytopes set, plus the global block numbers needed to build the
 * halo exchange.  Local block r of an owned polytope <-> owned_global_block[r] (ascending);
 * ghost k (local block n_owned + k) <-> ghost_global_block[k], owned by ghost_owner[k];
 * ghosts are grouped by owner rank (ascending) and ordered by global block inside a group.
 * Arrays stay owned by the handler and are valid until its next pdh_flatten / pdh_flatten_local (one rank's share at a
 * time: emulating several ranks in one process means pd_create, or a copy, before flattening the next rank). */
typedef struct pdh_local_info
{
  int32_t        n_owned, n_ghost;
  const int32_t *owned_global_block; /* [n_owned] */
  const int32_t *ghost_global_block; /* [n_ghost] */
  const int32_t *ghost_owner;        /* [n_ghost] */
  const int32_t *local_poly_global;  /* [n_owned + n_ghost] global polytope index of local polytope */
} pdh_local_info;
int pdh_flatten_local(pdh_handler *ah, const pdh_flatten_params *prm, const int32_t *owner, int32_t rank,
                      pd_mesh_desc *out, pdh_local_info *info);
/* Graph partitioning with METIS (libmetis_static.a of the CUDA toolkit, 64-bit indices), called like
 * deal.II's SparsityTools::partition: default options, PartGraphRecursive for n_parts <= 8, PartGraphKway
 * above.  Two uses on this path, both INPUTS of it: the METIS agglomeration shape
 * (GridTools::partition_triangulation on the cell face-adjacency graph, examples/poisson.cc) and the
 * distribution of polytopes over GPUs (vertex weight = sub-cells of the polytope, edge weight = shared
 * sub-faces).  CSR graph without self loops, both directions listed; weights may be NULL. */
int pdh_partition_graph(int64_t n_vertices, const int64_t *xadj, const int64_t *adjncy, const int64_t *vertex_weights,
                        const int64_t *edge_weights, int32_t n_parts, int32_t *part_out);
/* The polytope adjacency graph in the CSR form pdh_partition_graph takes (vertex weight = sub-cells of
 * the polytope, edge weight = shared sub-faces; neighbours in face order).  Returns the number of directed edges
 * (or < 0); any output may be NULL (call once with NULLs for the size). */
int64_t pdh_polytope_graph(const pdh_handler *ah, int64_t *xadj, int64_t *adjncy, int64_t *vertex_weights,
                           int64_t *edge_weights);
/* pdh_flatten + pd_create in one call */
int pdh_create_device(pdh_handler *ah, const pdh_flatten_params *prm, pd_handle **out);

#ifdef __cplusplus
}
#endif
#endif /* POLYDEAL_B200_H */
