// -----------------------------------------------------------------------------
// pd_api.cu -- C ABI of the device core (include/polydeal_b200.h, pd_* part).
// -----------------------------------------------------------------------------
#include "pd_host.hpp"
#include "pd_internal.hpp"

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <thread>

namespace pd
{
  static thread_local std::string g_last_error;
  void
  set_last_error(const std::string &m)
  {
    g_last_error = m;
  }

  // Gauss-Legendre on [0,1]: Newton on P_n from the Chebyshev guess, symmetric
  // pairs computed once and mirrored
  void
  make_gauss_1d(const int n, Quad1D &q)
  {
    if (n < 1 || n > 8)
      throw Error(PD_ERR_INVALID, "QGauss: 1 <= n <= 8 points per direction supported");
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < (n + 1) / 2; ++i)
      {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L)), pp = 1;
        for (int it = 0; it < 64; ++it)
          {
            long double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j)
              {
                const long double p3 = p2;
                p2                   = p1;
                p1                   = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
              }
            pp                   = n * (z * p1 - p2) / (z * z - 1);
            const long double dz = p1 / pp;
            z -= dz;
            if (fabsl(dz) < 1e-19L)
              break;
          }
        // z > 0 is the upper root of the pair
        const long double w = 1 / ((1 - z * z) * pp * pp); // weight on [0,1]
        q.x[n - 1 - i]      = (double)((1 + z) / 2);
        q.x[i]              = (double)((1 - z) / 2);
        q.w[i] = q.w[n - 1 - i] = (double)w;
      }
    if (n % 2)
      q.x[n / 2] = 0.5;
  }

  // Gauss-Lobatto nodes of FE_DGQ(p) on [0,1] in closed form
  void
  make_basis_1d(const int p, Basis1D &b)
  {
    std::memset(&b, 0, sizeof(b));
    switch (p)
      {
        case 0:
          b.node[0] = 0.5;
          break;
        case 1:
          b.node[0] = 0;
          b.node[1] = 1;
          break;
        case 2:
          b.node[0] = 0;
          b.node[1] = 0.5;
          b.node[2] = 1;
          break;
        case 3:
          {
            const long double s = sqrtl(5.0L) / 10;
            b.node[0]           = 0;
            b.node[1]           = (double)(0.5L - s);
            b.node[2]           = (double)(0.5L + s);
            b.node[3]           = 1;
            break;
          }
        case 4:
          {
            const long double s = sqrtl(3.0L / 7.0L) / 2;
            b.node[0]           = 0;
            b.node[1]           = (double)(0.5L - s);
            b.node[2]           = 0.5;
            b.node[3]           = (double)(0.5L + s);
            b.node[4]           = 1;
            break;
          }
        case 5:
          {
            const long double r = 2 * sqrtl(7.0L) / 21;
            const long double s1 = sqrtl(1.0L / 3 - r) / 2, s2 = sqrtl(1.0L / 3 + r) / 2;
            b.node[0] = 0;
            b.node[1] = (double)(0.5L - s2);
            b.node[2] = (double)(0.5L - s1);
            b.node[3] = (double)(0.5L + s1);
            b.node[4] = (double)(0.5L + s2);
            b.node[5] = 1;
            break;
          }
        default:
          throw Error(PD_ERR_UNSUPPORTED, "FE_DGQ degree must be in [0,5]");
      }
    for (int a = 0; a <= p; ++a)
      {
        double w = 1;
        for (int c = 0; c <= p; ++c)
          if (c != a)
            w /= (b.node[a] - b.node[c]);
        b.wprod[a] = w;
      }
  }

  template <class F>
  static int
  guarded(F &&f)
  {
    try
      {
        f();
        return PD_OK;
      }
    catch (const Error &e)
      {
        set_last_error(e.what());
        return e.code;
      }
    catch (const CudaError &e)
      {
        set_last_error(std::string("CUDA: ") + cudaGetErrorString(e.e) + " in `" + e.what + "` (line " +
                       std::to_string(e.line) + ")");
        return e.e == cudaErrorNotSupported ? PD_ERR_UNSUPPORTED : PD_ERR_CUDA;
      }
    catch (const std::exception &e)
      {
        set_last_error(e.what());
        return PD_ERR_INVALID;
      }
    catch (...)
      {
        set_last_error("unknown exception");
        return PD_ERR_INVALID;
      }
  }

  static void
  require_device()
  {
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      {
        cudaGetLastError();
        throw Error(PD_ERR_NO_DEVICE,
                    "polydeal_b200: no CUDA device visible; this path has no CPU fallback (sm_100a kernels only)");
      }
  }

  template <class T>
  static void
  h2d(DevBuf<T> &b, const T *src, size_t n, cudaStream_t s)
  {
    if (n)
      PD_CUDA(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }

  // every entry of a[0..n) inside [lo, hi)?  (the large index arrays: a few threads, one pass)
  static bool
  all_in_range(const int32_t *a, const int64_t n, const int64_t lo, const int64_t hi)
  {
    auto scan = [&](int64_t b, int64_t e) {
      int32_t mn = INT32_MAX, mx = INT32_MIN;
      for (int64_t i = b; i < e; ++i)
        {
          mn = std::min(mn, a[i]);
          mx = std::max(mx, a[i]);
        }
      return e <= b || ((int64_t)mn >= lo && (int64_t)mx < hi);
    };
    if (n < (int64_t)1 << 20)
      return scan(0, n);
    constexpr int            NT = 4;
    std::array<char, NT>     ok{};
    std::vector<std::thread> th;
    for (int t = 0; t < NT; ++t)
      th.emplace_back([&, t] { ok[t] = scan(n * t / NT, n * (t + 1) / NT); });
    for (auto &t : th)
      t.join();
    return std::all_of(ok.begin(), ok.end(), [](char c) { return c != 0; });
  }

  // the arrays pd_upload may change without a new handle: coordinates and penalties only
  static void
  validate_large_arrays(const pd_mesh_desc &d)
  {
    auto need = [](bool ok, const char *m) {
      if (!ok)
        throw Error(PD_ERR_INVALID, std::string("pd_mesh_desc: ") + m);
    };
    need(all_in_range(d.cell_verts, d.n_cells << d.dim, 0, d.n_verts), "cell_verts out of range");
    need(all_in_range(d.poly_subcell_idx, d.poly_subcell_ptr[d.n_polytopes], 0, d.n_cells), "sub-cell index out of range");
    const int64_t nsf = d.n_ifaces ? d.iface_sub_ptr[d.n_ifaces] : 0;
    need(nsf == 0 || (d.sub_cell && d.sub_face && d.sub_sigma), "missing sub-face arrays");
    need(all_in_range(d.sub_cell, nsf, 0, d.n_cells), "sub_cell out of range");
    need(all_in_range(d.sub_face, nsf, 0, 2 * d.dim), "sub_face out of range");
  }

  static void
  validate(const pd_mesh_desc &d)
  {
    auto need = [](bool ok, const char *m) {
      if (!ok)
        throw Error(PD_ERR_INVALID, std::string("pd_mesh_desc: ") + m);
    };
    need(d.dim == 2 || d.dim == 3, "dim must be 2 or 3");
    need(d.fe_degree >= 0 && d.fe_degree <= 5, "fe_degree out of range");
    need(d.n_q1d >= 1 && d.n_q1d <= 8 && d.n_q1d_face >= 1 && d.n_q1d_face <= 8, "n_q1d out of range");
    need(d.n_verts > 0 && d.verts, "no vertices");
    need(d.n_cells > 0 && d.cell_verts, "no cells");
    need(d.n_polytopes > 0 && d.poly_subcell_ptr && d.poly_subcell_idx && d.bbox && d.dof_block, "no polytopes");
    need(d.n_ifaces >= 0 && (d.n_ifaces == 0 || (d.iface_polyA && d.iface_polyB && d.iface_sub_ptr)), "bad interface list");
    const int32_t n_own = d.n_owned_polytopes > 0 ? d.n_owned_polytopes : d.n_polytopes;
    need(n_own <= d.n_polytopes, "n_owned_polytopes exceeds n_polytopes");
    need(d.n_block_rows == n_own && d.brow_ptr && d.bcol_idx, "block pattern must have one row per owned polytope");
    need(d.poly_subcell_ptr[0] == 0, "poly_subcell_ptr[0] != 0");
    for (int32_t p = 0; p < d.n_polytopes; ++p)
      {
        if (p < n_own)
          {
            need(d.poly_subcell_ptr[p + 1] > d.poly_subcell_ptr[p], "empty polytope");
            need(d.dof_block[p] < n_own, "owned polytopes must carry the first n_owned block indices");
          }
        else
          {
            need(d.poly_subcell_ptr[p + 1] == d.poly_subcell_ptr[p], "ghost polytopes carry no sub-cells");
            need(d.dof_block[p] >= n_own, "ghost polytopes must carry block indices >= n_owned");
          }
        need(d.dof_block[p] >= 0 && d.dof_block[p] < d.n_polytopes, "dof_block out of range");
        for (int k = 0; k < d.dim; ++k)
          need(d.bbox[(size_t)p * 2 * d.dim + d.dim + k] > d.bbox[(size_t)p * 2 * d.dim + k], "degenerate bounding box");
      }
    for (int32_t f = 0; f < d.n_ifaces; ++f)
      {
        need(d.iface_polyA[f] >= 0 && d.iface_polyA[f] < n_own, "iface_polyA must be an owned polytope");
        need(d.iface_polyB[f] >= -1 && d.iface_polyB[f] < d.n_polytopes, "iface_polyB out of range");
        need(d.iface_sub_ptr[f + 1] >= d.iface_sub_ptr[f], "iface_sub_ptr not monotone");
      }
    validate_large_arrays(d);
  }

  static int64_t
  find_block(const pd_mesh_desc &d, const int32_t brow, const int32_t bcol)
  {
    const int32_t *b = d.bcol_idx + d.brow_ptr[brow], *e = d.bcol_idx + d.brow_ptr[brow + 1];
    const int32_t *p = std::lower_bound(b, e, bcol);
    if (p == e || *p != bcol)
      throw Error(PD_ERR_INVALID, "pd_mesh_desc: block (" + std::to_string(brow) + "," + std::to_string(bcol) +
                                    ") needed by an interface is not in the block pattern");
    return p - b;
  }

  static void
  upload_descriptor(pd_handle *h, const pd_mesh_desc &d)
  {
    cudaStream_t s = h->stream;
    h2d(h->verts, d.verts, (size_t)d.n_verts * d.dim, s);
    h2d(h->cell_verts, d.cell_verts, (size_t)d.n_cells << d.dim, s);
    h2d(h->subcell_ptr, d.poly_subcell_ptr, (size_t)d.n_polytopes + 1, s);
    h2d(h->subcell_idx, d.poly_subcell_idx, (size_t)h->n_subcells, s);
    h2d(h->bbox, d.bbox, (size_t)d.n_polytopes * 2 * d.dim, s);
    h2d(h->dof_block, d.dof_block, (size_t)d.n_polytopes, s);
    h2d(h->ifA, d.iface_polyA, (size_t)d.n_ifaces, s);
    h2d(h->ifB, d.iface_polyB, (size_t)d.n_ifaces, s);
    h2d(h->if_sub_ptr, d.iface_sub_ptr, (size_t)d.n_ifaces + 1, s);
    h2d(h->sub_cell, d.sub_cell, (size_t)h->n_subfaces, s);
    h2d(h->sub_face, d.sub_face, (size_t)h->n_subfaces, s);
    h2d(h->sub_sigma, d.sub_sigma, (size_t)h->n_subfaces, s);
    h2d(h->brow_ptr, d.brow_ptr, (size_t)d.n_block_rows + 1, s);
    h2d(h->bcol, d.bcol_idx, (size_t)h->n_blocks, s);
    h->h_bbox.assign(d.bbox, d.bbox + (size_t)d.n_polytopes * 2 * d.dim);
    h->quad_valid = false;
    h->assembled  = false;
    h->mfd_valid  = false; // cached inverse diagonal of the matrix-free operators
    ++h->op_generation;    // graphs captured around the old state are stale
  }

  // 64-bit mix of the arrays the fine-mesh operators are derived from (pd_upload re-derives them on change)
  static uint64_t
  hash_bytes(uint64_t hsh, const void *p, const size_t bytes)
  {
    const uint64_t *w = static_cast<const uint64_t *>(p);
    const size_t    n = bytes / 8;
    uint64_t        a = hsh, b = 0x9e3779b97f4a7c15ull, c = 0xc2b2ae3d27d4eb4full, d = 0x165667b19e3779f9ull;
    size_t          i = 0;
    for (; i + 4 <= n; i += 4)
      {
        a = (a ^ w[i]) * 0x100000001b3ull;
        b = (b ^ w[i + 1]) * 0x100000001b3ull;
        c = (c ^ w[i + 2]) * 0x100000001b3ull;
        d = (d ^ w[i + 3]) * 0x100000001b3ull;
      }
    for (; i < n; ++i)
      a = (a ^ w[i]) * 0x100000001b3ull;
    const unsigned char *t = static_cast<const unsigned char *>(p) + n * 8;
    for (size_t k = 0; k < bytes % 8; ++k)
      a = (a ^ t[k]) * 0x100000001b3ull;
    return a ^ (b << 1) ^ (c << 2) ^ (d << 3);
  }
  static uint64_t
  fine_geometry_hash(const pd_handle *h, const pd_mesh_desc &d)
  {
    uint64_t x = 0xcbf29ce484222325ull;
    x = hash_bytes(x, d.verts, sizeof(double) * (size_t)d.n_verts * d.dim);
    x = hash_bytes(x, d.cell_verts, sizeof(int32_t) * ((size_t)d.n_cells << d.dim));
    x = hash_bytes(x, d.poly_subcell_idx, sizeof(int32_t) * (size_t)h->n_subcells);
    x = hash_bytes(x, d.bbox, sizeof(double) * (size_t)d.n_polytopes * 2 * d.dim);
    x = hash_bytes(x, d.sub_cell, sizeof(int32_t) * (size_t)h->n_subfaces);
    x = hash_bytes(x, d.sub_face, sizeof(int32_t) * (size_t)h->n_subfaces);
    x = hash_bytes(x, d.sub_sigma, sizeof(double) * (size_t)h->n_subfaces);
    return x;
  }

  static void
  ensure_quadrature_buffers(pd_handle *h)
  {
    if (h->vq_w.p || h->Q == 0)
      return;
    h->vq_x.alloc((size_t)h->Q * h->dim);
    h->vq_w.alloc((size_t)h->Q);
    h->fq_x.alloc((size_t)h->Qf * h->dim);
    h->fq_n.alloc((size_t)h->Qf * h->dim);
    h->fq_w.alloc((size_t)h->Qf);
  }

  // generic = the DMMA kernels (agglomerated quadrature streams + per-interface diagonal parts); the tensor path
  // of axis-aligned meshes needs neither (config C: 4.3 GB of points + 6.5 GB of face parts less)
  static void
  ensure_assembly_buffers(pd_handle *h, const bool generic = false)
  {
    if (generic)
      {
        ensure_quadrature_buffers(h);
        if (!h->face_diag.p && h->n_ifaces > 0)
          h->face_diag.alloc((size_t)h->n_ifaces * 2 * h->n * h->n);
      }
    if (h->values.p || h->nnz == 0)
      return;
    h->values.alloc((size_t)h->nnz);
    // pattern blocks no interface covers (a caller's wider pattern) stay zero
    PD_CUDA(cudaMemsetAsync(h->values.p, 0, sizeof(double) * (size_t)h->nnz, h->stream));
  }

  static void
  create(const pd_mesh_desc &d, pd_handle **out)
  {
    require_device();
    validate(d);
    if (!assemble_supported(d.dim, d.fe_degree, d.fe_kind))
      throw Error(PD_ERR_UNSUPPORTED, std::string("no sm_100a kernel for ") +
                                        (d.fe_kind == PD_FE_AGGLODGP ? "FE_AggloDGP<" : "FE_DGQ<") + std::to_string(d.dim) +
                                        ">(" + std::to_string(d.fe_degree) + "); supported: 2-D p=1..4, 3-D p=1..3");
    pd_handle *h = new pd_handle;
    try
      {
        h->dim    = d.dim;
        h->degree = d.fe_degree;
        h->n1     = d.fe_degree + 1;
        h->fe_kind = d.fe_kind;
        h->n       = 1;
        if (d.fe_kind == PD_FE_DGQ)
          for (int k = 0; k < d.dim; ++k)
            h->n *= h->n1;
        else
          for (int k = 1; k <= d.dim; ++k)
            h->n = h->n * (d.fe_degree + k) / k;
        h->nq1  = d.n_q1d;
        h->nq1f = d.n_q1d_face;
        h->nqc  = 1;
        h->nqf  = 1;
        for (int k = 0; k < d.dim; ++k)
          h->nqc *= h->nq1;
        for (int k = 0; k + 1 < d.dim; ++k)
          h->nqf *= h->nq1f;
        make_basis_1d(h->degree, h->basis);
        make_gauss_1d(h->nq1, h->quad);
        make_gauss_1d(h->nq1f, h->quadf);
        {
          double r[32];
          for (int i = 0; i < 8; ++i)
            {
              r[i]      = h->quad.x[i];
              r[8 + i]  = h->quad.w[i];
              r[16 + i] = h->quadf.x[i];
              r[24 + i] = h->quadf.w[i];
            }
          h->rules.alloc(32);
          PD_CUDA(cudaMemcpy(h->rules.p, r, sizeof(r), cudaMemcpyHostToDevice));
        }
        h->n_verts    = d.n_verts;
        h->n_cells    = d.n_cells;
        h->np         = d.n_polytopes;
        h->np_own     = d.n_owned_polytopes > 0 ? d.n_owned_polytopes : d.n_polytopes;
        h->n_ifaces   = d.n_ifaces;
        h->n_subcells = d.poly_subcell_ptr[d.n_polytopes];
        h->n_subfaces = d.n_ifaces ? d.iface_sub_ptr[d.n_ifaces] : 0;
        h->Q          = h->n_subcells * h->nqc;
        h->Qf         = h->n_subfaces * h->nqf;
        h->n_blocks   = d.brow_ptr[d.n_block_rows];
        h->n_dofs     = (int64_t)h->np_own * h->n; // rows; source vectors have np * n entries (owned + ghost)
        h->nnz        = h->n_blocks * h->n * h->n;
        int dev       = 0;
        PD_CUDA(cudaGetDevice(&dev));
        PD_CUDA(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
        PD_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        for (auto &e : h->ev)
          PD_CUDA(cudaEventCreate(&e));

        h->verts.alloc((size_t)d.n_verts * d.dim);
        h->cell_verts.alloc((size_t)d.n_cells << d.dim);
        h->subcell_ptr.alloc((size_t)h->np + 1);
        h->subcell_idx.alloc((size_t)h->n_subcells);
        h->bbox.alloc((size_t)h->np * 2 * d.dim);
        h->dof_block.alloc((size_t)h->np);
        h->ifA.alloc((size_t)h->n_ifaces);
        h->ifB.alloc((size_t)h->n_ifaces);
        h->if_sub_ptr.alloc((size_t)h->n_ifaces + 1);
        h->sub_cell.alloc((size_t)h->n_subfaces);
        h->sub_face.alloc((size_t)h->n_subfaces);
        h->sub_sigma.alloc((size_t)h->n_subfaces);
        h->brow_ptr.alloc((size_t)h->np_own + 1);
        h->bcol.alloc((size_t)h->n_blocks);
        if (h->n_ifaces == 0)
          {
            // keep a valid pointer for the one-element ptr array
            const int64_t zero = 0;
            PD_CUDA(cudaMemcpy(h->if_sub_ptr.p, &zero, sizeof(zero), cudaMemcpyHostToDevice));
          }
        upload_descriptor(h, d);
        setup_cartesian(h, d);

        h->h_brow_ptr.assign(d.brow_ptr, d.brow_ptr + h->np_own + 1);
        h->h_bcol.assign(d.bcol_idx, d.bcol_idx + h->n_blocks);
        h->h_dof_block.assign(d.dof_block, d.dof_block + h->np);
        h->h_subcell_ptr.assign(d.poly_subcell_ptr, d.poly_subcell_ptr + d.n_polytopes + 1);
        if (d.n_ifaces > 0)
          {
            h->h_if_sub_ptr.assign(d.iface_sub_ptr, d.iface_sub_ptr + d.n_ifaces + 1);
            h->h_ifA.assign(d.iface_polyA, d.iface_polyA + d.n_ifaces);
            h->h_ifB.assign(d.iface_polyB, d.iface_polyB + d.n_ifaces);
          }
        else
          h->h_if_sub_ptr.assign(1, 0);

        // ---- derived index data ------------------------------------------------
        const int64_t        nn = (int64_t)h->n * h->n;
        // Rows exist for OWNED polytopes only.  An interface whose B side is a ghost
        // polytope (owned by another rank) contributes M11 and M12 to A's row; its M21 / M22
        // belong to the other rank, which evaluates the same interface itself.
        const int32_t        n_own = h->np_own;
        std::vector<int32_t> row_stride(n_own);
        for (int32_t b = 0; b < n_own; ++b)
          row_stride[b] = (int32_t)(d.brow_ptr[b + 1] - d.brow_ptr[b]) * h->n;
        std::vector<int64_t> diag_base(n_own);
        std::vector<char>    seen_block(h->np, 0);
        for (int32_t p = 0; p < h->np; ++p)
          {
            const int32_t b = d.dof_block[p];
            if (seen_block[b])
              throw Error(PD_ERR_INVALID, "pd_mesh_desc: dof_block is not a permutation");
            seen_block[b] = 1;
            if (p < n_own)
              diag_base[p] = d.brow_ptr[b] * nn + find_block(d, b, b) * h->n;
          }
        std::vector<int64_t> baseAB(h->n_ifaces, -1), baseBA(h->n_ifaces, -1);
        std::vector<int64_t> padj_ptr(n_own + 1, 0);
        std::vector<char>    slot_taken((size_t)h->n_blocks, 0); // an off-diagonal block has exactly one writer
        auto                 take = [&](const int64_t slot) {
          if (slot_taken[(size_t)slot])
            throw Error(PD_ERR_INVALID, "pd_mesh_desc: two interfaces join the same pair of polytopes");
          slot_taken[(size_t)slot] = 1;
        };
        for (int32_t f = 0; f < h->n_ifaces; ++f)
          {
            const int32_t pa = d.iface_polyA[f], pb = d.iface_polyB[f];
            ++padj_ptr[pa + 1];
            if (pb >= 0)
              {
                if (pb == pa)
                  throw Error(PD_ERR_INVALID, "pd_mesh_desc: interface joins a polytope with itself");
                const int32_t ba = d.dof_block[pa], bb = d.dof_block[pb];
                const int64_t kab = find_block(d, ba, bb);
                take(d.brow_ptr[ba] + kab);
                baseAB[f] = d.brow_ptr[ba] * nn + kab * h->n;
                if (pb < n_own)
                  {
                    ++padj_ptr[pb + 1];
                    const int64_t kba = find_block(d, bb, ba);
                    take(d.brow_ptr[bb] + kba);
                    baseBA[f] = d.brow_ptr[bb] * nn + kba * h->n;
                  }
              }
          }
        for (int32_t p = 0; p < n_own; ++p)
          padj_ptr[p + 1] += padj_ptr[p];
        std::vector<int64_t> padj(padj_ptr[n_own]), cursor(padj_ptr.begin(), padj_ptr.end() - 1);
        for (int32_t f = 0; f < h->n_ifaces; ++f)
          {
            padj[cursor[d.iface_polyA[f]]++] = (int64_t)f * 2;
            if (d.iface_polyB[f] >= 0 && d.iface_polyB[f] < n_own)
              padj[cursor[d.iface_polyB[f]]++] = (int64_t)f * 2 + 1;
          }
        auto put32 = [&](DevBuf<int32_t> &b, const std::vector<int32_t> &v) {
          b.alloc(v.size());
          if (!v.empty())
            PD_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        };
        auto put64 = [&](DevBuf<int64_t> &b, const std::vector<int64_t> &v) {
          b.alloc(v.size());
          if (!v.empty())
            PD_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        };
        put32(h->row_stride, row_stride);
        put64(h->diag_base, diag_base);
        put64(h->if_baseAB, baseAB);
        put64(h->if_baseBA, baseBA);
        put64(h->padj_ptr, padj_ptr);
        put64(h->padj, padj);

        // quadrature streams, work buffers and the matrix are allocated on first use
        // (ensure_assembly_buffers): a handle used only for the matrix-free apply never
        // pays for them
        if (h->fe_kind == PD_FE_DGQ)
          {
            setup_fine_operator(h, d);
            setup_mapped_operator(h, d);
            if (h->n_subcells == h->np_own)
              h->fine_geo_hash = fine_geometry_hash(h, d);
          }
        PD_CUDA(cudaStreamSynchronize(h->stream));
      }
    catch (...)
      {
        for (auto &e : h->ev)
          if (e)
            cudaEventDestroy(e);
        if (h->own_stream)
          cudaStreamDestroy(h->own_stream);
        delete h;
        throw;
      }
    *out = h;
  }
} // namespace pd

using namespace pd;

extern "C"
{
  const char *
  pd_last_error(void)
  {
    return g_last_error.c_str();
  }

  int
  pd_device_count(void)
  {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
      {
        cudaGetLastError();
        return 0;
      }
    return n;
  }

  int
  pd_create(const pd_mesh_desc *desc, pd_handle **out)
  {
    return guarded([&] {
      if (!desc || !out)
        throw Error(PD_ERR_INVALID, "pd_create: null argument");
      *out = nullptr;
      create(*desc, out);
    });
  }

  int
  pd_destroy(pd_handle *h)
  {
    return guarded([&] {
      if (!h)
        return;
      cudaStreamSynchronize(h->stream);
      if (h->cg_graph_exec)
        cudaGraphExecDestroy(h->cg_graph_exec);
      if (h->cg_graph1_exec)
        cudaGraphExecDestroy(h->cg_graph1_exec);
      if (h->ev_order)
        cudaEventDestroy(h->ev_order);
      if (h->ev_fork)
        cudaEventDestroy(h->ev_fork);
      if (h->ev_join)
        cudaEventDestroy(h->ev_join);
      if (h->aux_stream)
        cudaStreamDestroy(h->aux_stream);
      for (auto &e : h->ev)
        if (e)
          cudaEventDestroy(e);
      if (h->own_stream)
        cudaStreamDestroy(h->own_stream);
      delete h;
    });
  }

  int
  pd_upload(pd_handle *h, const pd_mesh_desc *d)
  {
    return guarded([&] {
      if (!h || !d)
        throw Error(PD_ERR_INVALID, "pd_upload: null argument");
      if (!d->verts || !d->cell_verts || !d->poly_subcell_ptr || !d->poly_subcell_idx || !d->bbox || !d->dof_block ||
          !d->brow_ptr || !d->bcol_idx || (d->n_ifaces > 0 && (!d->iface_polyA || !d->iface_polyB || !d->iface_sub_ptr)))
        throw Error(PD_ERR_INVALID, "pd_upload: null array in the descriptor");
      if (d->dim != h->dim || d->fe_degree != h->degree || d->fe_kind != h->fe_kind || d->n_q1d != h->nq1 || d->n_q1d_face != h->nq1f ||
          d->n_verts != h->n_verts || d->n_cells != h->n_cells || d->n_polytopes != h->np ||
          (d->n_owned_polytopes > 0 ? d->n_owned_polytopes : d->n_polytopes) != h->np_own ||
          d->n_ifaces != h->n_ifaces || d->n_block_rows != h->np_own)
        throw Error(PD_ERR_INVALID, "pd_upload: descriptor sizes differ from the ones the handle was created with");
      // the TOPOLOGY (polytope -> sub-cell ranges, DoF blocks, interface list, block pattern) is what every derived
      // table of the handle was built from: it must be the one of pd_create.  Coordinates, bounding boxes,
      // penalties and the (range-checked) cell / sub-cell / sub-face index arrays may change.
      auto same = [](const auto &v, const auto *p) { return v.empty() || std::memcmp(v.data(), p, v.size() * sizeof(v[0])) == 0; };
      if (!same(h->h_subcell_ptr, d->poly_subcell_ptr) || !same(h->h_dof_block, d->dof_block) ||
          !same(h->h_brow_ptr, d->brow_ptr) || !same(h->h_bcol, d->bcol_idx) ||
          (d->n_ifaces > 0 && (!same(h->h_if_sub_ptr, d->iface_sub_ptr) || !same(h->h_ifA, d->iface_polyA) ||
                               !same(h->h_ifB, d->iface_polyB))))
        throw Error(PD_ERR_INVALID,
                    "pd_upload: the topology (sub-cell ranges, dof_block, interface list, block pattern) differs from "
                    "pd_create's; only coordinates, bounding boxes, penalties and cell indices may change -- create a new handle");
      validate_large_arrays(*d);
      for (int32_t p = 0; p < d->n_polytopes; ++p)
        for (int k = 0; k < d->dim; ++k)
          if (!(d->bbox[(size_t)p * 2 * d->dim + d->dim + k] > d->bbox[(size_t)p * 2 * d->dim + k]))
            throw Error(PD_ERR_INVALID, "pd_upload: degenerate bounding box");
      upload_descriptor(h, *d);
      setup_cartesian(h, *d); // waits for the upload (an 8-byte read-back): which assembly path applies, bricks still valid?
      // fine-mesh operators (every polytope one cell): geometry tables, stencil records and tile plans are derived
      // from the coordinates and penalties -- re-derive them when those changed
      if (h->fe_kind == PD_FE_DGQ && h->n_subcells == h->np_own)
        {
          const uint64_t hsh = fine_geometry_hash(h, *d);
          if (hsh != h->fine_geo_hash)
            {
              h->mf_rec_valid = false;
              h->mp_geo_valid = false;
              setup_fine_operator(h, *d);
              setup_mapped_operator(h, *d);
              h->fine_geo_hash = hsh;
            }
        }
    });
  }

  int
  pd_set_stream(pd_handle *h, void *cuda_stream)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      PD_CUDA(cudaStreamSynchronize(h->stream));
      h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
      ++h->op_generation;
    });
  }

  int
  pd_synchronize(pd_handle *h)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      PD_CUDA(cudaStreamSynchronize(h->stream));
    });
  }

  int
  pd_build_quadrature(pd_handle *h)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      ensure_quadrature_buffers(h);
      launch_quadrature(h);
      h->quad_valid = true;
    });
  }

  static void
  need_quadrature(pd_handle *h)
  {
    ensure_quadrature_buffers(h);
    if (!h->quad_valid)
      {
        launch_quadrature(h);
        h->quad_valid = true;
      }
  }

  int64_t
  pd_n_quadrature_points(const pd_handle *h, int faces)
  {
    return h ? (faces ? h->Qf : h->Q) : 0;
  }

  int
  pd_quadrature_device(pd_handle *h, const double **vol_x, const double **vol_jxw, const double **face_x,
                       const double **face_n, const double **face_jxw)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      need_quadrature(h);
      if (vol_x)
        *vol_x = h->vq_x.p;
      if (vol_jxw)
        *vol_jxw = h->vq_w.p;
      if (face_x)
        *face_x = h->fq_x.p;
      if (face_n)
        *face_n = h->fq_n.p;
      if (face_jxw)
        *face_jxw = h->fq_w.p;
    });
  }

  int
  pd_quadrature_to_host(pd_handle *h, double *vol_x, double *vol_jxw, double *face_x, double *face_n, double *face_jxw)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      need_quadrature(h);
      auto get = [&](double *dst, const double *src, const size_t count) {
        if (dst && count)
          PD_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      };
      get(vol_x, h->vq_x.p, (size_t)h->Q * h->dim);
      get(vol_jxw, h->vq_w.p, (size_t)h->Q);
      get(face_x, h->fq_x.p, (size_t)h->Qf * h->dim);
      get(face_n, h->fq_n.p, (size_t)h->Qf * h->dim);
      get(face_jxw, h->fq_w.p, (size_t)h->Qf);
      PD_CUDA(cudaStreamSynchronize(h->stream));
    });
  }

  int
  pd_assemble_rhs(pd_handle *h, const double *f_vol_dev, const double *g_face_dev, double stiffness, double *rhs_dev)
  {
    return guarded([&] {
      if (!h || !rhs_dev)
        throw Error(PD_ERR_INVALID, "pd_assemble_rhs: null argument");
      need_quadrature(h);
      launch_poly_rhs(h, f_vol_dev, g_face_dev, stiffness, rhs_dev);
    });
  }

  int
  pd_error_norms(pd_handle *h, const double *u_dev, const double *exact_dev, const double *exact_grad_dev, double *l2,
                 double *h1_seminorm)
  {
    return guarded([&] {
      if (!h || !u_dev || !exact_dev || !l2)
        throw Error(PD_ERR_INVALID, "pd_error_norms: null argument");
      if (h1_seminorm && !exact_grad_dev)
        throw Error(PD_ERR_INVALID, "pd_error_norms: the H1 seminorm needs the exact gradient");
      need_quadrature(h);
      if (h->sv_scal.n < 8)
        h->sv_scal.alloc(8);
      launch_poly_error(h, u_dev, exact_dev, h1_seminorm ? exact_grad_dev : nullptr, h->sv_scal.p);
      double e[2];
      PD_CUDA(cudaMemcpyAsync(e, h->sv_scal.p, sizeof e, cudaMemcpyDeviceToHost, h->stream));
      PD_CUDA(cudaStreamSynchronize(h->stream));
      *l2 = std::sqrt(e[0]);
      if (h1_seminorm)
        *h1_seminorm = std::sqrt(e[1]);
    });
  }

  static void
  finish_transfer(pd_transfer *t, const std::vector<int32_t> &parent, const std::vector<int32_t> &blk)
  {
    pd_handle *h  = t->coarse;
    t->n_children = (int32_t)parent.size();
    std::vector<int64_t> pc_ptr((size_t)h->np_own + 1, 0);
    for (const int32_t p : parent)
      {
        if (p < 0 || p >= h->np_own)
          throw Error(PD_ERR_INVALID, "pd_transfer: parent polytope out of range");
        ++pc_ptr[p + 1];
      }
    for (int32_t p = 0; p < h->np_own; ++p)
      pc_ptr[p + 1] += pc_ptr[p];
    std::vector<int32_t> pc_idx(parent.size());
    std::vector<int64_t> fill(pc_ptr.begin(), pc_ptr.end() - 1);
    for (size_t c = 0; c < parent.size(); ++c)
      pc_idx[fill[parent[c]]++] = (int32_t)c;
    auto put = [](auto &buf, const auto &v) {
      buf.alloc(v.size());
      if (!v.empty())
        PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    put(t->parent, parent);
    put(t->child_blk, blk);
    put(t->pc_ptr, pc_ptr);
    put(t->pc_idx, pc_idx);
    t->partial.alloc(parent.size() * (size_t)h->n);
  }

  int
  pd_transfer_create(pd_handle *coarse, pd_handle *fine, const int32_t *parent_of_fine, pd_transfer **out)
  {
    return guarded([&] {
      if (!coarse || !fine || !parent_of_fine || !out)
        throw Error(PD_ERR_INVALID, "pd_transfer_create: null argument");
      if (coarse->fe_kind != PD_FE_DGQ || fine->fe_kind != PD_FE_DGQ)
        throw Error(PD_ERR_UNSUPPORTED, "pd_transfer_create: FE_DGQ only");
      if (coarse->dim != fine->dim || coarse->degree != fine->degree)
        throw Error(PD_ERR_INVALID, "pd_transfer_create: the two levels must use the same FE_DGQ space");
      std::unique_ptr<pd_transfer> t(new pd_transfer);
      t->coarse      = coarse;
      t->kind        = 0;
      t->child_bbox  = fine->bbox.p;
      t->n_fine_dofs = (int64_t)fine->np_own * fine->n;
      std::vector<int32_t> parent(parent_of_fine, parent_of_fine + fine->np_own);
      std::vector<int32_t> blk(fine->h_dof_block.begin(), fine->h_dof_block.begin() + fine->np_own);
      finish_transfer(t.get(), parent, blk);
      *out = t.release();
    });
  }

  int
  pd_transfer_create_to_cells(pd_handle *h, pd_transfer **out)
  {
    return guarded([&] {
      if (!h || !out)
        throw Error(PD_ERR_INVALID, "pd_transfer_create_to_cells: null argument");
      if (h->fe_kind != PD_FE_DGQ)
        throw Error(PD_ERR_UNSUPPORTED, "pd_transfer_create_to_cells: FE_DGQ only");
      std::unique_ptr<pd_transfer> t(new pd_transfer);
      t->coarse      = h;
      t->kind        = 1;
      t->n_fine_dofs = h->n_cells * h->n;
      const int64_t        ns = h->h_subcell_ptr[h->np_own];
      std::vector<int32_t> cells((size_t)ns), parent((size_t)ns);
      if (ns)
        PD_CUDA(cudaMemcpy(cells.data(), h->subcell_idx.p, sizeof(int32_t) * (size_t)ns, cudaMemcpyDeviceToHost));
      for (int32_t p = 0; p < h->np_own; ++p)
        for (int64_t k = h->h_subcell_ptr[p]; k < h->h_subcell_ptr[p + 1]; ++k)
          parent[(size_t)k] = p;
      finish_transfer(t.get(), parent, cells);
      *out = t.release();
    });
  }

  void
  pd_transfer_destroy(pd_transfer *t)
  {
    delete t;
  }

  int64_t
  pd_transfer_m(const pd_transfer *t)
  {
    return t ? t->n_fine_dofs : 0;
  }

  int64_t
  pd_transfer_n(const pd_transfer *t)
  {
    return t ? (int64_t)t->coarse->np_own * t->coarse->n : 0;
  }

  int
  pd_transfer_prolongate(pd_transfer *t, const double *src_coarse_dev, double *dst_fine_dev, int add)
  {
    return guarded([&] {
      if (!t || !src_coarse_dev || !dst_fine_dev)
        throw Error(PD_ERR_INVALID, "pd_transfer_prolongate: null argument");
      launch_transfer(t->coarse, *t, false, src_coarse_dev, dst_fine_dev, add != 0);
    });
  }

  int
  pd_transfer_restrict(pd_transfer *t, const double *src_fine_dev, double *dst_coarse_dev, int add)
  {
    return guarded([&] {
      if (!t || !src_fine_dev || !dst_coarse_dev)
        throw Error(PD_ERR_INVALID, "pd_transfer_restrict: null argument");
      launch_transfer(t->coarse, *t, true, src_fine_dev, dst_coarse_dev, add != 0);
    });
  }

  int
  pd_peer_create(pd_handle *h, int rank, int world, const int64_t *send_ptr, const int32_t *send_blocks,
                 const int64_t *recv_ptr, const int64_t *remote_offset, pd_peer **out)
  {
    return guarded([&] {
      if (!out)
        throw Error(PD_ERR_INVALID, "pd_peer_create: null argument");
      *out = peer_create(h, rank, world, send_ptr, send_blocks, recv_ptr, remote_offset);
    });
  }
  int
  pd_peer_handle_bytes(void)
  {
    return peer_handle_bytes();
  }
  int
  pd_peer_export(pd_peer *p, void *handles_out)
  {
    return guarded([&] { peer_export(p, handles_out); });
  }
  int
  pd_peer_connect(pd_peer *p, const void *all_handles)
  {
    return guarded([&] { peer_connect(p, all_handles); });
  }
  int
  pd_peer_exchange(pd_peer *p, double *x_full_dev)
  {
    return guarded([&] { peer_exchange(p, x_full_dev); });
  }
  int
  pd_peer_vmult(pd_peer *p, int mode, double *x_full_dev, double *dst_dev, int add)
  {
    return guarded([&] { peer_vmult(p, mode, x_full_dev, dst_dev, add != 0); });
  }
  int
  pd_peer_fused(pd_peer *p)
  {
    return peer_fused(p);
  }

  int
  pd_peer_status(pd_peer *p)
  {
    return peer_status(p);
  }
  void
  pd_peer_destroy(pd_peer *p)
  {
    peer_destroy(p);
  }

  // ---- the reinit() family (pd_reinit.cu) ----
  int64_t
  pd_reinit_n_points(const pd_handle *h, int32_t poly)
  {
    int64_t   n = -1;
    const int e = guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      n = reinit_n_points(h, poly);
    });
    return e == PD_OK ? n : e;
  }
  int64_t
  pd_reinit_iface_n_points(const pd_handle *h, int32_t iface)
  {
    int64_t   n = -1;
    const int e = guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      n = reinit_iface_n_points(h, iface);
    });
    return e == PD_OK ? n : e;
  }
  int
  pd_reinit_polytope(pd_handle *h, int32_t poly, double *values, double *grads, double *jxw, double *points)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      reinit_n_points(h, poly); // range check
      need_quadrature(h);
      reinit_polytope(h, poly, values, grads, jxw, points, nullptr);
    });
  }
  int
  pd_agglomerated_quadrature(pd_handle *h, int32_t poly, double *unit_points, double *jxw, double *real_points)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      reinit_n_points(h, poly);
      need_quadrature(h);
      reinit_polytope(h, poly, nullptr, nullptr, jxw, real_points, unit_points);
    });
  }
  int
  pd_reinit_face(pd_handle *h, int32_t iface, int32_t side, double *values, double *grads, double *jxw, double *points,
                 double *normals)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      reinit_iface_n_points(h, iface);
      need_quadrature(h);
      reinit_iface(h, iface, side, values, grads, jxw, points, normals);
    });
  }
  int
  pd_reinit_interface(pd_handle *h, int32_t iface, double *values0, double *grads0, double *values1, double *grads1,
                      double *jxw, double *points, double *normals)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      reinit_iface_n_points(h, iface);
      if (h->h_ifB[iface] < 0)
        throw Error(PD_ERR_INVALID, "pd_reinit_interface: a boundary face has no second side (use pd_reinit_face)");
      need_quadrature(h);
      reinit_iface(h, iface, 0, values0, grads0, jxw, points, normals);
      reinit_iface(h, iface, 1, values1, grads1, nullptr, nullptr, nullptr);
    });
  }
  int
  pd_fe_evaluate(int32_t fe_kind, int32_t dim, int32_t degree, int64_t n_points, const double *unit_points, double *values,
                 double *grads)
  {
    return guarded([&] {
      require_device();
      if (!unit_points && n_points > 0)
        throw Error(PD_ERR_INVALID, "pd_fe_evaluate: null argument");
      fe_evaluate(fe_kind, dim, degree, n_points, unit_points, values, grads);
    });
  }

  int
  pd_invalidate_quadrature(pd_handle *h)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      h->quad_valid      = false;
      h->brick_mat_valid = false; // the tensor path's analogue: cached brick geometry and 1-D matrices
    });
  }

  int
  pd_assemble(pd_handle *h, uint32_t flags, const pd_coefficients *coef)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      pd_coefficients c{1.0, 0.0};
      if (coef)
        c = *coef;
      // axis-aligned sub-cells: per-sub-cell sum factorisation (pd_cartesian.cu), no quadrature streams;
      // anything else: the DMMA kernels on the agglomerated quadrature (pd_assemble.cu)
      const bool cart = cartesian_assembly_selected(h);
      ensure_assembly_buffers(h, !cart);
      PD_CUDA(cudaEventRecord(h->ev[4], h->stream));
      const bool built = !cart && !h->quad_valid;
      if (built)
        {
          launch_quadrature(h);
          h->quad_valid = true;
        }
      if (cart)
        {
          if (!(flags & PD_ASSEMBLE_INTERIOR))
            PD_CUDA(cudaMemsetAsync(h->values.p, 0, sizeof(double) * h->nnz, h->stream));
          launch_assemble_cartesian(h, flags, c);
        }
      else
        launch_assemble(h, flags, c);
      h->last_assembly_path = cart ? 1 : 0;
      h->assembled  = true;
      h->last_ms[3] = built ? -1.f : 0.f; // resolved lazily in pd_last_kernel_ms
    });
  }

  int64_t
  pd_n_dofs(const pd_handle *h)
  {
    return h ? h->n_dofs : 0;
  }
  int64_t
  pd_nnz(const pd_handle *h)
  {
    return h ? h->nnz : 0;
  }
  int64_t
  pd_n_source_dofs(const pd_handle *h)
  {
    return h ? (int64_t)h->np * h->n : 0;
  }
  int32_t
  pd_n_dofs_per_cell(const pd_handle *h)
  {
    return h ? h->n : 0;
  }

  int
  pd_matrix_values_device(pd_handle *h, double **dev_values)
  {
    return guarded([&] {
      if (!h || !dev_values)
        throw Error(PD_ERR_INVALID, "null argument");
      ensure_assembly_buffers(h);
      *dev_values = h->values.p;
    });
  }

  int
  pd_matrix_values_to_host(pd_handle *h, double *host_values)
  {
    return guarded([&] {
      if (!h || !host_values)
        throw Error(PD_ERR_INVALID, "null argument");
      if (!h->assembled)
        throw Error(PD_ERR_STATE, "pd_matrix_values_to_host: pd_assemble has not been called");
      PD_CUDA(cudaMemcpyAsync(host_values, h->values.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
      PD_CUDA(cudaStreamSynchronize(h->stream));
    });
  }

  int
  pd_matrix_values_to_host_async(pd_handle *h, double *host_values_pinned)
  {
    return guarded([&] {
      if (!h || !host_values_pinned)
        throw Error(PD_ERR_INVALID, "null argument");
      if (!h->assembled)
        throw Error(PD_ERR_STATE, "pd_matrix_values_to_host_async: pd_assemble has not been called");
      PD_CUDA(cudaMemcpyAsync(host_values_pinned, h->values.p, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
    });
  }

  int
  pd_matrix_pattern_to_host(pd_handle *h, int64_t *rowptr, int32_t *cols)
  {
    return guarded([&] {
      if (!h || !rowptr || !cols)
        throw Error(PD_ERR_INVALID, "null argument");
      const int n = h->n;
      int64_t   k = 0;
      rowptr[0]   = 0;
      for (int32_t b = 0; b < h->np_own; ++b)
        for (int i = 0; i < n; ++i)
          {
            for (int64_t e = h->h_brow_ptr[b]; e < h->h_brow_ptr[b + 1]; ++e)
              for (int j = 0; j < n; ++j)
                cols[k++] = h->h_bcol[e] * n + j;
            rowptr[(int64_t)b * n + i + 1] = k;
          }
    });
  }

  static void vmult_impl(pd_handle *h, int mode, const double *src, double *dst, bool add);
}
namespace pd
{
  void
  vmult_dispatch(pd_handle *h, int mode, const double *src, double *dst, bool add)
  {
    vmult_impl(h, mode, src, dst, add);
  }
} // namespace pd
extern "C"
{
  static void
  vmult_impl(pd_handle *h, int mode, const double *src, double *dst, bool add)
  {
    if (!h || !src || !dst)
      throw Error(PD_ERR_INVALID, "pd_vmult: null argument");
    if (mode == PD_VMULT_BLOCK_CSR)
      {
        if (!h->assembled)
          throw Error(PD_ERR_STATE, "pd_vmult(BLOCK_CSR): pd_assemble has not been called");
        launch_spmv(h, src, dst, add);
      }
    else if (h->fe_kind != PD_FE_DGQ && mode != PD_VMULT_MATRIX_FREE)
      throw Error(PD_ERR_UNSUPPORTED, "the fine-mesh matrix-free operators are defined for FE_DGQ only");
    else if (mode == PD_VMULT_MATRIX_FREE)
      {
        if (h->mf_ready && !h->force_generic_mf)
          launch_fine_operator(h, src, dst, add); // fine Cartesian mesh: sum-factorised stencil kernel
        else if (cartesian_apply_available(h))
          launch_cart_apply(h, src, dst, add); // axis-aligned sub-cells: sum factorisation per sub-cell / sub-face
        else
          {
            // agglomerated polytopes: regenerate the basis at the agglomerated quadrature points
            need_quadrature(h);
            launch_poly_apply(h, src, dst, add);
          }
      }
    else if (mode == PD_VMULT_MAPPED_FINE)
      {
        if (!h->mp_ready)
          throw Error(PD_ERR_STATE,
                      "pd_vmult(MAPPED_FINE): needs one cell per polytope, no ghosts, QGauss(p+1) on cells and faces "
                      "and neighbouring cells in standard orientation");
        launch_mapped_operator(h, src, dst, add);
      }
    else
      throw Error(PD_ERR_INVALID, "pd_vmult: unknown mode");
  }

  int
  pd_set_operator(pd_handle *h, uint32_t flags, const pd_coefficients *coef)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      h->op_flags = flags;
      h->op_coef  = coef ? *coef : pd_coefficients{1.0, 0.0};
      ++h->op_generation;
    });
  }

  int
  pd_matrix_free_available(const pd_handle *h)
  {
    return h && h->mf_ready ? 1 : 0;
  }

  int
  pd_fine_kernel_last(const pd_handle *h)
  {
    return h ? h->mf_kernel_last : 0;
  }

  int
  pd_mapped_fine_available(const pd_handle *h)
  {
    return h && h->mp_ready ? 1 : 0;
  }

  int
  pd_force_generic_matrix_free(pd_handle *h, int on)
  {
    return guarded([&] {
      if (!h)
        throw Error(PD_ERR_INVALID, "null handle");
      h->force_generic_mf = on != 0;
      ++h->op_generation;
    });
  }

  int
  pd_vmult(pd_handle *h, int mode, const double *src_dev, double *dst_dev)
  {
    return guarded([&] { vmult_impl(h, mode, src_dev, dst_dev, false); });
  }

  int
  pd_vmult_add(pd_handle *h, int mode, const double *src_dev, double *dst_dev)
  {
    return guarded([&] { vmult_impl(h, mode, src_dev, dst_dev, true); });
  }

  int
  pd_vmult_host(pd_handle *h, int mode, const double *src_host, double *dst_host)
  {
    return guarded([&] {
      if (!h || !src_host || !dst_host)
        throw Error(PD_ERR_INVALID, "pd_vmult_host: null argument");
      const size_t n_src = (size_t)h->np * h->n;
      if (h->vec_a.n != n_src)
        {
          h->vec_a.alloc(n_src);
          h->vec_b.alloc((size_t)h->n_dofs);
        }
      PD_CUDA(cudaMemcpyAsync(h->vec_a.p, src_host, sizeof(double) * n_src, cudaMemcpyHostToDevice, h->stream));
      vmult_impl(h, mode, h->vec_a.p, h->vec_b.p, false);
      PD_CUDA(cudaMemcpyAsync(dst_host, h->vec_b.p, sizeof(double) * h->n_dofs, cudaMemcpyDeviceToHost, h->stream));
      PD_CUDA(cudaStreamSynchronize(h->stream));
    });
  }

  int
  pd_cg_solve(pd_handle *h, int mode, const double *b_dev, double *x_dev, int max_iter, double rel_tol, int jacobi,
              int *iterations, double *relative_residual)
  {
    bool      converged = true;
    const int rc        = guarded([&] {
      if (!h || !b_dev || !x_dev)
        throw Error(PD_ERR_INVALID, "pd_cg_solve: null argument");
      converged = solver_cg(h, mode, b_dev, x_dev, max_iter, rel_tol, jacobi, iterations, relative_residual);
    });
    return rc != PD_OK ? rc : (converged ? PD_OK : PD_NOT_CONVERGED);
  }

  int
  pd_cg_solve_sharded(pd_peer *peer, int mode, const double *b_dev, double *x_dev, int max_iter, double rel_tol,
                      int jacobi, int *iterations, double *relative_residual)
  {
    bool      converged = true;
    const int rc        = guarded([&] {
      if (!peer || !b_dev || !x_dev)
        throw Error(PD_ERR_INVALID, "pd_cg_solve_sharded: null argument");
      converged =
        solver_cg(peer_handle(peer), mode, b_dev, x_dev, max_iter, rel_tol, jacobi, iterations, relative_residual, peer);
      if (peer_status(peer) != PD_OK)
        throw Error(PD_ERR_STATE, "pd_cg_solve_sharded: a rank did not arrive at a collective within the time-out");
    });
    return rc != PD_OK ? rc : (converged ? PD_OK : PD_NOT_CONVERGED);
  }

  int
  pd_peer_allreduce(pd_peer *p, double *scalars_dev, int count)
  {
    return guarded([&] { peer_allreduce(p, scalars_dev, 0, count); });
  }

  int
  pd_estimate_lambda_max(pd_handle *h, int mode, int n_iterations, double *lambda_max)
  {
    return guarded([&] {
      if (!h || !lambda_max || n_iterations < 1)
        throw Error(PD_ERR_INVALID, "pd_estimate_lambda_max: bad argument");
      *lambda_max = solver_lambda_max(h, mode, n_iterations);
    });
  }

  int
  pd_estimate_lambda_max_sharded(pd_peer *peer, int mode, int n_iterations, double *lambda_max)
  {
    return guarded([&] {
      if (!peer || !lambda_max || n_iterations < 1)
        throw Error(PD_ERR_INVALID, "pd_estimate_lambda_max_sharded: bad argument");
      *lambda_max = solver_lambda_max(peer_handle(peer), mode, n_iterations, peer);
      if (peer_status(peer) != PD_OK)
        throw Error(PD_ERR_STATE, "pd_estimate_lambda_max_sharded: a rank did not arrive at a collective within the time-out");
    });
  }

  int
  pd_chebyshev_smooth_sharded(pd_peer *peer, int mode, int degree, double lambda_max, double smoothing_range,
                              const double *b_dev, double *x_full_dev, int zero_initial_guess)
  {
    return guarded([&] {
      if (!peer || !b_dev || !x_full_dev)
        throw Error(PD_ERR_INVALID, "pd_chebyshev_smooth_sharded: null argument");
      solver_chebyshev(peer_handle(peer), mode, degree, lambda_max, smoothing_range, b_dev, x_full_dev, zero_initial_guess,
                       peer);
    });
  }

  int
  pd_chebyshev_smooth(pd_handle *h, int mode, int degree, double lambda_max, double smoothing_range, const double *b_dev,
                      double *x_dev, int zero_initial_guess)
  {
    return guarded([&] {
      if (!h || !b_dev || !x_dev)
        throw Error(PD_ERR_INVALID, "pd_chebyshev_smooth: null argument");
      solver_chebyshev(h, mode, degree, lambda_max, smoothing_range, b_dev, x_dev, zero_initial_guess);
    });
  }

  int
  pd_diagonal_inverse(pd_handle *h, double *dst_dev)
  {
    return guarded([&] {
      if (!h || !dst_dev)
        throw Error(PD_ERR_INVALID, "null argument");
      if (!h->assembled)
        throw Error(PD_ERR_STATE, "pd_diagonal_inverse: pd_assemble has not been called");
      launch_diagonal_inverse(h, dst_dev);
    });
  }

  int
  pd_diagonal_inverse_of(pd_handle *h, int mode, double *dst_dev)
  {
    return guarded([&] {
      if (!h || !dst_dev)
        throw Error(PD_ERR_INVALID, "null argument");
      if (mode == PD_VMULT_MAPPED_FINE && h->fe_kind != PD_FE_DGQ)
        throw Error(PD_ERR_UNSUPPORTED, "the fine-mesh matrix-free operators are defined for FE_DGQ only");
      solver_diagonal_inverse(h, mode, dst_dev);
    });
  }

  int
  pd_copy_array(pd_handle *h, const char *name, double *host_out, int64_t *count)
  {
    return guarded([&] {
      if (!h || !name)
        throw Error(PD_ERR_INVALID, "null argument");
      const std::string s(name);
      const double     *p = nullptr;
      int64_t           n = 0;
      if (s == "vol_qpt")
        p = h->vq_x.p, n = h->Q * h->dim;
      else if (s == "vol_jxw")
        p = h->vq_w.p, n = h->Q;
      else if (s == "face_qpt")
        p = h->fq_x.p, n = h->Qf * h->dim;
      else if (s == "face_normal")
        p = h->fq_n.p, n = h->Qf * h->dim;
      else if (s == "face_jxw")
        p = h->fq_w.p, n = h->Qf;
      else
        throw Error(PD_ERR_INVALID, "pd_copy_array: unknown array '" + s + "'");
      if (count)
        *count = n;
      if (host_out)
        {
          if (!h->quad_valid)
            throw Error(PD_ERR_STATE, "pd_copy_array: quadrature has not been built");
          PD_CUDA(cudaMemcpyAsync(host_out, p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
          PD_CUDA(cudaStreamSynchronize(h->stream));
        }
    });
  }

  int
  pd_assembly_path(const pd_handle *h)
  {
    return h ? h->last_assembly_path : -1;
  }

  int
  pd_tensor_path_stats(const pd_handle *h, int64_t *stats5)
  {
    return guarded([&] {
      if (!h || !stats5)
        throw Error(PD_ERR_INVALID, "null argument");
      stats5[0] = h->cartesian ? 1 : 0;
      stats5[1] = h->n_cell_bricks;
      stats5[2] = h->n_face_bricks;
      stats5[3] = h->n_diag_items;
      stats5[4] = h->n_apply_items;
    });
  }

  int64_t
  pd_launch_count(const pd_handle *h)
  {
    return h ? h->launches : 0;
  }

  int
  pd_last_kernel_ms(pd_handle *h, float *ms4)
  {
    return guarded([&] {
      if (!h || !ms4)
        throw Error(PD_ERR_INVALID, "null argument");
      if (!h->assembled)
        throw Error(PD_ERR_STATE, "pd_last_kernel_ms: pd_assemble has not been called");
      PD_CUDA(cudaEventSynchronize(h->ev[3]));
      PD_CUDA(cudaEventElapsedTime(&ms4[0], h->ev[0], h->ev[1]));
      PD_CUDA(cudaEventElapsedTime(&ms4[1], h->ev[1], h->ev[2]));
      PD_CUDA(cudaEventElapsedTime(&ms4[2], h->ev[2], h->ev[3]));
      PD_CUDA(cudaEventElapsedTime(&ms4[3], h->ev[4], h->ev[0]));
    });
  }

  // exposed so the CPU-only test-suite can check the rule tables the kernels use
  int
  pd_quadrature_rule_1d(int n, double *x, double *w)
  {
    return guarded([&] {
      Quad1D q;
      make_gauss_1d(n, q);
      for (int i = 0; i < n; ++i)
        {
          x[i] = q.x[i];
          w[i] = q.w[i];
        }
    });
  }

  int
  pd_dgq_nodes_1d(int degree, double *nodes)
  {
    return guarded([&] {
      Basis1D b;
      make_basis_1d(degree, b);
      for (int i = 0; i <= degree; ++i)
        nodes[i] = b.node[i];
    });
  }
}
