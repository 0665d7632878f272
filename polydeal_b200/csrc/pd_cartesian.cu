// -----------------------------------------------------------------------------
// pd_cartesian.cu -- SIP-DG assembly over agglomerates of AXIS-ALIGNED cells by
// per-sub-cell sum factorisation (the tensor path of pd_assemble).
//
// Same result as pd_assemble.cu (reference: PolyUtils::assemble_dg_matrix,
// include/poly_utils.h:2000-2195, kernels :1870-1926, through reinit / MappingBox,
// source/agglomeration_handler.cc:729-906, source/mapping_box.cc:393-439), another
// algorithm.  On a sub-cell [a, a + eta] that is an axis-aligned box, the agglomerated
// quadrature QGauss<dim>(n_q) is a TENSOR grid in the bounding-box coordinates and the basis
// is a tensor product, so every integral the reference sums point by point factorises into
// 1-D sums:
//   int_cell  d_x phi_i d_x phi_j = K^x[a,a'] M^y[b,b'] M^z[c,c'],   M^d[a,a'] = eta_d sum_q w_q l_a l_a'(xhat_q),
//                                                                    K^d[a,a'] = eta_d sum_q w_q l_a' l_a''(xhat_q) / h_d^2
//   int_face(normal x) (-1/2 dn phi_i phi_j - 1/2 phi_i dn phi_j + pen phi_i phi_j)
//                                  = F^x[a,a'] M^y[b,b'] M^z[c,c'],  F^x from l, l' at the face coordinate
// and the block of a polytope is  sum over its sub-cells / sub-faces of  X (x) Y (x) Z  -- a
// rank-structured sum with n^2 (2 + 1/S) multiply-adds per sub-cell instead of the 2 n^2 dim n_q^dim
// of the point-wise contraction (p = 3: 8.2 k instead of 1.57 M flops per sub-cell).  The sums over the
// quadrature points are the SAME sums the reference forms, regrouped; results agree to rounding
// (tests: per block entry <= 1e-12 against the CPU restatement of the reference, and against the DMMA kernels).
//
// Applies when every owned sub-cell is an axis-aligned box in standard orientation (checked on the
// device by k_check_axis_aligned at pd_create / pd_upload); distorted meshes take the DMMA kernels
// of pd_assemble.cu.  FE_DGQ and FE_AggloDGP (the latter is a sub-set of the same tensor index set).
//
//   k_cart_diag     one CTA per polytope: diagonal block = volume + boundary faces + own side
//                   (M11 / M22) of every interior interface, accumulated in registers in the fixed
//                   order sub-cells, then adjacency list -- deterministic, no atomics, no partials
//   k_cart_offdiag  one CTA per interior interface: M12 from cross matrices of the two bounding-box
//                   bases, written to (A,B) and, transposed (M21 = M12^T), to (B,A)
// Thread layout: thread (col, ig) owns the column col = (b,b',c,c') of the factorised block with all its
// N1^2 rows (a,a') in registers; per item it forms its Y (x) Z entry from two shared-memory loads and sweeps
// the rows with 16-byte broadcast loads of X -- the 1-D matrices are the only shared-memory state.  Small
// elements deal the items of a chunk to IG item groups so that a CTA still has ~256 threads.
// Roofline: the matrix is written once (8 n^2 bytes per block: config C 7.3 GB) => HBM-bound.
// -----------------------------------------------------------------------------
#include "pd_host.hpp"
#include "pd_internal.hpp"
#include "pd_device.cuh"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

namespace pd
{
  namespace
  {
    struct CartArgs
    {
      const double  *verts;
      const int32_t *cell_verts;
      const int64_t *subcell_ptr;
      const int32_t *subcell_idx;
      const double  *bbox;
      const int32_t *ifA, *ifB;
      const int64_t *if_sub_ptr;
      const int32_t *sub_cell, *sub_face;
      const double  *sub_sigma;
      const int64_t *padj_ptr, *padj;
      const int64_t *diag_base, *if_baseAB, *if_baseBA;
      const int32_t *dof_block, *row_stride;
      double        *values;
      double         stiffness, mass;
      uint32_t       flags;
      int32_t        np_own, n_ifaces, nq, nqf;
      Basis1D        basis;
      Quad1D         quad, quadf;
      unsigned char  dof_abc[64][3]; // (a, b, c) of every DoF of the element
      // bricks (see build_bricks): tensor-product sets of sub-cells / sub-faces whose sums factorise
      const int64_t *cbk_ptr; // [np_own + 1] cell bricks of a polytope
      const int32_t *cbk_iv;  // [n_cbk][2 dim]: (start, count) per axis into civ
      const int32_t *civ;     // representative cell of every interval
      const int64_t *fbk_ptr; // [n_ifaces + 1] face bricks of an interface / boundary face
      const int32_t *fbk_s;   // [n_fbk] representative sub-face (plane, orientation, penalty)
      const int32_t *fbk_iv;  // [n_fbk][2 (dim - 1)]: (start, count) per tangential axis into fiv
      const int32_t *fiv;     // representative (A-side) cell of every tangential interval
      // what the hot kernels read (refreshed from the representatives by k_refresh_bricks after every upload)
      const double2 *civ_box, *fiv_box;     // (lo, hi) of every interval
      const double  *fbk_plane, *fbk_sigma; // plane coordinate and penalty of a face brick
      const int32_t *fbk_lf;                // its local face number (axis, side) seen from the listing polytope
      // item list of a polytope: [pit_ptr[p], pit_diag_end[p]) cell bricks and own-side face bricks (the diagonal
      // block), [pit_diag_end[p], pit_ptr[p + 1]) the cross face bricks (matrix-free apply only)
      const int64_t *pit_ptr, *pit_diag_end;
      const int32_t *pit_brick, *pit_meta, *pit_q; // meta = kind (1 cell, 2 own face, 3 cross face) | side << 2 | boundary << 3
      // the 1-D matrices of every brick (k_brick_matrices, once per geometry): what the hot kernels load
      const int32_t *cbk_poly, *fbk_iface; // polytope of a cell brick, interface of a face brick
      double        *cmat;                 // [n_cbk][dim][2][NXP]: mass-like and stiffness-like sums on the polytope's box
      double        *fmat;                 // [n_fbk][3][dim][NXP]: own side A | own side B | cross (rows A, columns B); the
                                           // normal axis holds the face factor WITHOUT the stiffness coefficient
      int64_t        n_cbk, n_fbk;
    };

    // cells of the mesh that are NOT axis-aligned boxes in standard orientation (vertex v at lo + bit_k(v) (hi - lo))
    // -> flags[0]; with bricks (cpos != nullptr): sub-cells / sub-faces whose box no longer is the product of their
    // brick's representative intervals (vertices moved through pd_upload) -> flags[1]
    __global__ void __launch_bounds__(256)
    k_check_axis_aligned(const double *verts, const int32_t *cell_verts, const int32_t *subcell_idx, const int64_t n_sub,
                         const int dim, int *flags, const int32_t *cpos, const int32_t *cbk_iv, const int32_t *civ,
                         const int64_t n_subfaces, const int32_t *sub_cell, const int32_t *sub_face, const double *sub_sigma,
                         const int32_t *fpos, const int32_t *fbk_s, const int32_t *fbk_iv, const int32_t *fiv)
    {
      const int vpc = 1 << dim;
      int       bad = 0, stale = 0;
      auto      box = [&](const int32_t c, const int k, double &lo, double &hi) {
        const int32_t *cv = cell_verts + (int64_t)c * vpc;
        lo                = verts[(int64_t)cv[0] * dim + k];
        hi                = verts[(int64_t)cv[vpc - 1] * dim + k];
      };
      for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_sub; s += (int64_t)gridDim.x * blockDim.x)
        {
          const int32_t  c  = subcell_idx[s];
          const int32_t *cv = cell_verts + (int64_t)c * vpc;
          const double  *lo = verts + (int64_t)cv[0] * dim, *hi = verts + (int64_t)cv[vpc - 1] * dim;
          for (int k = 0; k < dim; ++k)
            bad |= !(hi[k] > lo[k]);
          for (int v = 1; v < vpc - 1; ++v)
            {
              const double *x = verts + (int64_t)cv[v] * dim;
              for (int k = 0; k < dim; ++k)
                bad |= x[k] != (((v >> k) & 1) ? hi[k] : lo[k]);
            }
          if (cpos)
            {
              const int32_t *ps = cpos + s * (1 + dim);
              for (int k = 0; k < dim; ++k)
                {
                  double rl, rh;
                  box(civ[cbk_iv[(int64_t)ps[0] * 2 * dim + 2 * k] + ps[1 + k]], k, rl, rh);
                  stale |= rl != lo[k] || rh != hi[k];
                }
            }
        }
      if (fpos)
        for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_subfaces; s += (int64_t)gridDim.x * blockDim.x)
          {
            const int32_t *ps = fpos + s * dim;
            const int32_t  rs = fbk_s[ps[0]];
            const int      lf = sub_face[s], fd = lf >> 1;
            stale |= lf != sub_face[rs] || sub_sigma[s] != sub_sigma[rs];
            double l0, h0, l1, h1;
            box(sub_cell[s], fd, l0, h0);
            box(sub_cell[rs], fd, l1, h1);
            stale |= ((lf & 1) ? h0 : l0) != ((lf & 1) ? h1 : l1);
            for (int k = 0; k < dim; ++k)
              if (k != fd)
                {
                  const int t = k < fd ? k : k - 1;
                  box(sub_cell[s], k, l0, h0);
                  box(fiv[fbk_iv[(int64_t)ps[0] * 2 * (dim - 1) + 2 * t] + ps[1 + t]], k, l1, h1);
                  stale |= l0 != l1 || h0 != h1;
                }
          }
      const int any_bad = __syncthreads_or(bad), any_stale = __syncthreads_or(stale);
      if (threadIdx.x == 0)
        {
          if (any_bad)
            atomicAdd(flags, 1);
          if (any_stale)
            atomicAdd(flags + 1, 1);
        }
    }

    template <int DIM, int DEGX>
    struct CartCfg
    {
      using C                 = Cfg<DIM, DEGX>;
      static constexpr int N1 = C::N1, NX = N1 * N1, NYZ = ipow(NX, DIM - 1), NF = ipow(N1, DIM);
      static constexpr int NXP = NX + (NX & 1); // slot stride: even, so that the X rows load as 16-byte pairs
      static constexpr int ISTR = DIM * 2 * NXP;
      // A work item (the diagonal block of a polytope / the M12 block of an interface) is computed by a GROUP of
      // threads: thread t owns the columns t, t + GROUP, ... (CPT of them) of the factorised block -- column =
      // (b,b'[,c,c']) -- with all NX rows (a,a') in registers.  Large elements (3-D p = 3: 256 columns): the whole
      // CTA is one group.  Small elements: a WARP is a group, a CTA runs WPC independent work items at once and only
      // warp-level barriers are used -- the work per block is a few hundred flops, what matters is how many blocks
      // are in flight per SM.
      // Which axis lives in the registers ("row axis"): x for even N1 -- a thread's N1 entries (a, 0..N1-1) are one
      // 16- / 32-byte store -- and FE_AggloDGP; the LAST axis for odd N1, where the lanes then run over (a', b') fastest
      // so that a warp's scalar stores cover runs of N1^(dim-1) consecutive entries of a row instead of one each.
      static constexpr int RA    = (N1 % 2 == 0 || C::DGP) ? 0 : DIM - 1;
      static constexpr int GROUP = NYZ >= 128 ? ((NYZ + 31) / 32) * 32 : 32;
      static constexpr int CPT   = (NYZ + GROUP - 1) / GROUP;
      static constexpr int NTHR  = GROUP >= 128 ? GROUP : 256;
      static constexpr int WPC   = NTHR / GROUP;
      static constexpr int CH    = 8; // items per chunk
      static constexpr int MINB  = GROUP >= 128 ? 4 : 2;
      // shared memory of a group: the chunk's 1-D matrices, later the staged block; + the kinds of the chunk's items
      static constexpr int GSM   = ((C::DGP && NF * (NF + 1) > CH * ISTR) ? NF * (NF + 1) : CH * ISTR) + CH / 2 + 2;
    };

    // a column of the factorised block: the index pairs (i_d, i_d') of the two column axes ax1 < ax2 (2-D: one)
    template <int DIM, int DEGX>
    struct ColumnIndex
    {
      using CC = CartCfg<DIM, DEGX>;
      int p1, p2;         // pair index i * N1 + i' along ax1 / ax2
      int q[2][2];        // (i, i') along ax1, ax2
      static constexpr int ax1 = CC::RA == 0 ? 1 : 0, ax2 = CC::RA == 0 ? 2 : 1;
      __device__ __forceinline__ explicit ColumnIndex(const int col)
      {
        constexpr int N1 = CC::N1, NX = CC::NX;
        if (CC::RA == 0)
          {
            p1 = DIM == 3 ? col % NX : col;
            p2 = DIM == 3 ? col / NX : 0;
          }
        else if (DIM == 3)
          { // lanes run over (a', b') fastest, then (a, b)
            const int ap = col % N1, bp = (col / N1) % N1, a = (col / NX) % N1, b = col / (NX * N1);
            p1 = a * N1 + ap;
            p2 = b * N1 + bp;
          }
        else
          {
            p1 = (col / N1) * N1 + col % N1; // (a, a'): a' fastest
            p2 = 0;
          }
        q[0][0] = p1 / N1, q[0][1] = p1 % N1, q[1][0] = p2 / N1, q[1][1] = p2 % N1;
      }
    };

    template <int GROUP>
    __device__ __forceinline__ void
    group_sync()
    {
      if (GROUP == 32)
        __syncwarp();
      else
        __syncthreads();
    }

    // rows a of the 1-D matrices of one item along one axis, [a'] = 0..N1-1:
    //  cell / tangential axis:  M = eta sum_q w l_a l_a',  K = eta sum_q w l_a' l_a'' (derivatives per real length)
    template <class C>
    __device__ __forceinline__ void
    axis_rows_mass(const Basis1D &B, const double *qx, const double *qw, const int nq, const double c_lo, const double eta,
                   const double b_lo, const double inv_h, const int a, double *M, double *K)
    {
      constexpr int N1 = C::N1; // accumulates: the caller zeroes M, K
      for (int q = 0; q < nq; ++q)
        {
          double       L[N1], dL[N1];
          const double xh = (c_lo + eta * qx[q] - b_lo) * inv_h;
          basis_1d<C>(B, xh, inv_h, L, dL);
          const double w = qw[q] * eta;
          double       la = 0., da = 0.;
#pragma unroll
          for (int k = 0; k < N1; ++k) // runtime row index without dynamic register indexing
            if (k == a)
              {
                la = L[k];
                da = dL[k];
              }
#pragma unroll
          for (int k = 0; k < N1; ++k)
            {
              M[k] += w * la * L[k];
              K[k] += w * da * dL[k];
            }
        }
    }

    // two bases (A rows, B columns) over the same interval: cross mass rows
    template <class C>
    __device__ __forceinline__ void
    axis_rows_cross(const Basis1D &B, const double *qx, const double *qw, const int nq, const double c_lo, const double eta,
                    const double a_lo, const double a_inv_h, const double b_lo, const double b_inv_h, const int a, double *M)
    {
      constexpr int N1 = C::N1; // accumulates: the caller zeroes M
      for (int q = 0; q < nq; ++q)
        {
          double       LA[N1], dLA[N1], LB[N1], dLB[N1];
          const double x = c_lo + eta * qx[q];
          basis_1d<C>(B, (x - a_lo) * a_inv_h, a_inv_h, LA, dLA);
          basis_1d<C>(B, (x - b_lo) * b_inv_h, b_inv_h, LB, dLB);
          const double w = qw[q] * eta;
          double       la = 0.;
#pragma unroll
          for (int k = 0; k < N1; ++k)
            if (k == a)
              la = LA[k];
#pragma unroll
          for (int k = 0; k < N1; ++k)
            M[k] += w * la * LB[k];
        }
    }

    template <int N1>
    __device__ __forceinline__ double
    pick(const double *v, const int a)
    {
      double r = 0.;
#pragma unroll
      for (int k = 0; k < N1; ++k)
        if (k == a)
          r = v[k];
      return r;
    }

    // box of a cell: lo corner = vertex 0, hi corner = vertex 2^dim - 1
    template <int DIM>
    __device__ __forceinline__ void
    cell_box(const CartArgs &A, const int32_t c, const int d, double &lo, double &hi)
    {
      const int32_t *cv = A.cell_verts + (int64_t)c * (1 << DIM);
      lo                = A.verts[(int64_t)cv[0] * DIM + d];
      hi                = A.verts[(int64_t)cv[(1 << DIM) - 1] * DIM + d];
    }

    // row a of the summed 1-D matrices of a brick along axis d: sum over its intervals (each read from the box of
    // its representative cell) of the mass-like and stiffness-like sums; row basis = column basis (box b_lo, inv_h)
    template <class C, int DIM>
    __device__ __forceinline__ void
    brick_rows_mass(const CartArgs &A, const double2 *boxes, const int start, const int count, const double *qx,
                    const double *qw, const int nq, const double b_lo, const double inv_h, const int a, double *M, double *K)
    {
#pragma unroll
      for (int k = 0; k < C::N1; ++k)
        M[k] = K[k] = 0.;
      for (int i = start; i < start + count; ++i)
        {
          const double2 bx = boxes[i];
          axis_rows_mass<C>(A.basis, qx, qw, nq, bx.x, bx.y - bx.x, b_lo, inv_h, a, M, K);
        }
    }
    // the same with a row basis (r_*) and another column basis (c_*)
    template <class C, int DIM>
    __device__ __forceinline__ void
    brick_rows_cross(const CartArgs &A, const double2 *boxes, const int start, const int count, const double *qx,
                     const double *qw, const int nq, const double r_lo, const double r_ih, const double c_lo, const double c_ih,
                     const int a, double *M)
    {
#pragma unroll
      for (int k = 0; k < C::N1; ++k)
        M[k] = 0.;
      for (int i = start; i < start + count; ++i)
        {
          const double2 bx = boxes[i];
          axis_rows_cross<C>(A.basis, qx, qw, nq, bx.x, bx.y - bx.x, r_lo, r_ih, c_lo, c_ih, a, M);
        }
    }

    // cached geometry of the bricks from their representatives (after pd_create / every pd_upload)
    __global__ void __launch_bounds__(256)
    k_refresh_bricks(const double *verts, const int32_t *cell_verts, const int dim, const int64_t n_cbk, const int32_t *cbk_iv,
                     const int32_t *civ, double2 *civ_box, const int64_t n_fbk, const int32_t *fbk_s, const int32_t *fbk_iv,
                     const int32_t *fiv, double2 *fiv_box, const int32_t *sub_cell, const int32_t *sub_face,
                     const double *sub_sigma, double *fbk_plane, double *fbk_sigma, int32_t *fbk_lf)
    {
      const int vpc = 1 << dim;
      auto      box = [&](const int32_t c, const int k) {
        const int32_t *cv = cell_verts + (int64_t)c * vpc;
        return make_double2(verts[(int64_t)cv[0] * dim + k], verts[(int64_t)cv[vpc - 1] * dim + k]);
      };
      const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      for (int64_t w = t0; w < n_cbk * dim; w += stride)
        {
          const int     k  = (int)(w % dim);
          const int32_t *iv = cbk_iv + (w / dim) * 2 * dim + 2 * k;
          for (int i = iv[0]; i < iv[0] + iv[1]; ++i)
            civ_box[i] = box(civ[i], k);
        }
      for (int64_t b = t0; b < n_fbk; b += stride)
        {
          const int32_t s  = fbk_s[b];
          const int     lf = sub_face[s], fd = lf >> 1;
          const double2 bx = box(sub_cell[s], fd);
          fbk_plane[b]     = (lf & 1) ? bx.y : bx.x;
          fbk_sigma[b]     = sub_sigma[s];
          fbk_lf[b]        = lf;
          for (int k = 0; k < dim; ++k)
            if (k != fd)
              {
                const int32_t *iv = fbk_iv + b * 2 * (dim - 1) + 2 * (k < fd ? k : k - 1);
                for (int i = iv[0]; i < iv[0] + iv[1]; ++i)
                  fiv_box[i] = box(fiv[i], k);
              }
        }
    }

    // X rows of a slot into registers, 16 bytes at a time (the address is the same for every thread of the
    // CTA: one broadcast wavefront per load)
    template <int NX>
    __device__ __forceinline__ void
    load_rows(const double *src, double *x)
    {
#pragma unroll
      for (int r = 0; r + 1 < NX; r += 2)
        {
          const double2 v = *reinterpret_cast<const double2 *>(src + r);
          x[r]            = v.x;
          x[r + 1]        = v.y;
        }
      if (NX & 1)
        x[NX - 1] = src[NX - 1];
    }

    // acc[k][r] += X1[r] yz1 + X2[r] yz2 over the items of a chunk, for the thread's columns k.
    // SL[item][d][2][NXP]: slot 0 = M-like, slot 1 = K-like.  kind[item]: 1 = cell (two terms), 2 = face (one term,
    // the face factor sits in slot 0 of its normal axis), 0 = switched off by the flags.
    template <int DIM, int DEGX>
    __device__ __forceinline__ void
    accumulate_chunk(const double *SL, const int *kind, const int cnt, const int t, const double stiffness, const double mass,
                     double (*acc)[CartCfg<DIM, DEGX>::NX])
    {
      using CC            = CartCfg<DIM, DEGX>;
      using CI            = ColumnIndex<DIM, DEGX>;
      constexpr int NX    = CC::NX, NXP = CC::NXP, ISTR = CC::ISTR, GROUP = CC::GROUP, CPT = CC::CPT, NYZ = CC::NYZ, RA = CC::RA;
      for (int it = 0; it < cnt; ++it)
        {
          const int kd = kind[it];
          if (kd == 0)
            continue;
          const double *s = SL + it * ISTR;
          double        x[NX], yz1[CPT], yz2[CPT];
#pragma unroll
          for (int k = 0; k < CPT; ++k)
            {
              const int col = t + k * GROUP;
              yz1[k] = yz2[k] = 0.;
              if (col < NYZ)
                {
                  const CI     ci(col);
                  const double m1 = s[(CI::ax1 * 2 + 0) * NXP + ci.p1];
                  const double m2 = DIM == 3 ? s[(CI::ax2 * 2 + 0) * NXP + ci.p2] : 1.;
                  yz1[k]          = m1 * m2;
                  if (kd == 1)
                    {
                      const double k1 = s[(CI::ax1 * 2 + 1) * NXP + ci.p1];
                      const double k2 = DIM == 3 ? s[(CI::ax2 * 2 + 1) * NXP + ci.p2] : 0.;
                      yz2[k]          = stiffness * (k1 * m2 + m1 * k2) + mass * yz1[k];
                      yz1[k] *= stiffness;
                    }
                }
            }
          if (kd == 1)
            {
              load_rows<NX>(s + (RA * 2 + 1) * NXP, x); // K along the row axis
#pragma unroll
              for (int k = 0; k < CPT; ++k)
#pragma unroll
                for (int r = 0; r < NX; ++r)
                  acc[k][r] += x[r] * yz1[k];
              load_rows<NX>(s + (RA * 2 + 0) * NXP, x); // M along the row axis
#pragma unroll
              for (int k = 0; k < CPT; ++k)
#pragma unroll
                for (int r = 0; r < NX; ++r)
                  acc[k][r] += x[r] * yz2[k];
            }
          else
            {
              load_rows<NX>(s + (RA * 2 + 0) * NXP, x);
#pragma unroll
              for (int k = 0; k < CPT; ++k)
#pragma unroll
                for (int r = 0; r < NX; ++r)
                  acc[k][r] += x[r] * yz1[k];
            }
        }
    }

    // registers -> OUT[full row index][full column index] (shared, NF x (NF + 1))
    template <int DIM, int DEGX>
    __device__ __forceinline__ void
    stage_block(double *OUT, const int t, const double (*acc)[CartCfg<DIM, DEGX>::NX])
    {
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, NF = CC::NF, GROUP = CC::GROUP, CPT = CC::CPT, NYZ = CC::NYZ;
#pragma unroll
      for (int k = 0; k < CPT; ++k)
        {
          const int col = t + k * GROUP;
          if (col >= NYZ)
            continue;
          static_assert(CartCfg<DIM, DEGX>::RA == 0 || !Cfg<DIM, DEGX>::DGP, "the staged path keeps x in the registers");
          const ColumnIndex<DIM, DEGX> ci(col);
          const int b = ci.q[0][0], bp = ci.q[0][1], c = ci.q[1][0], cp = ci.q[1][1];
#pragma unroll
          for (int r = 0; r < NX; ++r)
            {
              const int a = r / N1, ap = r % N1;
              const int i = a + N1 * (b + N1 * c), j = ap + N1 * (bp + N1 * cp);
              OUT[i * (NF + 1) + j] = acc[k][r];
            }
        }
    }

    // FE_DGQ: a thread's rows (a, a') for fixed a are N1 CONSECUTIVE entries of row i = (a,b,c) of the block (columns
    // (a',b',c'), a' fastest), so the block goes straight from the registers to the CSR rows -- N1 = 4: one 32-byte
    // sector per store pair, N1 = 2: 16 bytes -- without a shared-memory tile, a barrier or index arithmetic.
    // TRANSPOSED: the same registers as the block's transpose (M21 = M12^T): row (a',b',c'), columns (a,b,c).
    template <int DIM, int DEGX, bool TRANSPOSED>
    __device__ __forceinline__ void
    store_block_direct(double *base, const int stride, const int t, const double (*acc)[CartCfg<DIM, DEGX>::NX])
    {
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, GROUP = CC::GROUP, CPT = CC::CPT, NYZ = CC::NYZ;
#pragma unroll
      for (int k = 0; k < CPT; ++k)
        {
          const int col = t + k * GROUP;
          if (col >= NYZ)
            continue;
          const ColumnIndex<DIM, DEGX> ci(col);
          if (CC::RA != 0)
            {
              // registers = the pairs (c, c') of the last axis; the lanes run over (a', b') fastest: scalar stores, a warp
              // covers runs of N1^(dim-1) consecutive entries of a row.  (The transposed block is a work item of its own.)
              static_assert(CC::RA == 0 || !TRANSPOSED, "the transposed block is computed, not stored, on this path");
              const int rowq = ci.q[0][0] + (DIM == 3 ? N1 * ci.q[1][0] : 0), colq = ci.q[0][1] + (DIM == 3 ? N1 * ci.q[1][1] : 0);
              constexpr int NLAST = DIM == 3 ? N1 * N1 : N1; // stride of the last axis inside a row / column index
#pragma unroll
              for (int r = 0; r < NX; ++r)
                base[(int64_t)(rowq + NLAST * (r / N1)) * stride + colq + NLAST * (r % N1)] = acc[k][r];
              continue;
            }
          const int b = ci.q[0][0], bp = ci.q[0][1], c = ci.q[1][0], cp = ci.q[1][1];
          const int rowq = TRANSPOSED ? bp + N1 * cp : b + N1 * c;  // (b,c) part of the row index
          const int colq = TRANSPOSED ? b + N1 * c : bp + N1 * cp;  // (b,c) part of the column index
#pragma unroll
          for (int a = 0; a < N1; ++a)
            {
              double *dst = base + (int64_t)(a + N1 * rowq) * stride + N1 * colq;
              double  v[N1];
#pragma unroll
              for (int e = 0; e < N1; ++e)
                v[e] = TRANSPOSED ? acc[k][e * N1 + a] : acc[k][a * N1 + e];
              if (N1 == 4)
                // one 32-byte sector per store (st.global.v4.f64 -> STG.E.256 on sm_100a); dst is 32-byte aligned:
                // block bases and row strides are multiples of n = 64 doubles, the column offset a multiple of 4
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                             : "memory");
              else
                {
#pragma unroll
                  for (int e = 0; e < N1; e += 2)
                    *reinterpret_cast<double2 *>(dst + e) = make_double2(v[e], v[e + 1]);
                }
            }
        }
    }

    // The 1-D matrices of all bricks, one thread per (brick, axis): they depend on the geometry, the element and the
    // penalties only, so they are computed once per geometry (pd_create, pd_upload, pd_invalidate_quadrature) and
    // shared by the assembly and every matrix-free apply.  A thread evaluates the basis (both bases for a face
    // brick) once per quadrature point of its intervals and accumulates all its matrices as outer products.
    template <int DIM, int DEGX>
    __global__ void __launch_bounds__(128)
    k_brick_matrices(const CartArgs A)
    {
      using C           = Cfg<DIM, DEGX>;
      using CC          = CartCfg<DIM, DEGX>;
      constexpr int N1  = CC::N1, NX = CC::NX, NXP = CC::NXP;
      const int64_t nc  = A.n_cbk * DIM, nf = A.n_fbk * DIM;
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nc + nf; w += stride)
        {
          if (w < nc)
            {
              const int64_t brick = w / DIM;
              const int     d     = (int)(w % DIM);
              const double *bb    = A.bbox + (int64_t)A.cbk_poly[brick] * 2 * DIM;
              const double  b_lo = bb[d], inv_h = 1. / (bb[DIM + d] - bb[d]);
              const int32_t *iv = A.cbk_iv + brick * 2 * DIM + 2 * d;
              double         M[NX], K[NX];
#pragma unroll
              for (int k = 0; k < NX; ++k)
                M[k] = K[k] = 0.;
              for (int i = iv[0]; i < iv[0] + iv[1]; ++i)
                {
                  const double2 bx = A.civ_box[i];
                  for (int q = 0; q < A.nq; ++q)
                    {
                      double L[N1], dL[N1];
                      basis_1d<C>(A.basis, (bx.x + (bx.y - bx.x) * A.quad.x[q] - b_lo) * inv_h, inv_h, L, dL);
                      const double wq = A.quad.w[q] * (bx.y - bx.x);
#pragma unroll
                      for (int a = 0; a < N1; ++a)
#pragma unroll
                        for (int k = 0; k < N1; ++k)
                          {
                            M[a * N1 + k] += wq * L[a] * L[k];
                            K[a * N1 + k] += wq * dL[a] * dL[k];
                          }
                    }
                }
              double *dst = A.cmat + (brick * DIM + d) * 2 * NXP;
#pragma unroll
              for (int k = 0; k < NX; ++k)
                {
                  dst[k]       = M[k];
                  dst[NXP + k] = K[k];
                }
              continue;
            }
          const int64_t v     = w - nc;
          const int64_t brick = v / DIM;
          const int     d     = (int)(v % DIM);
          const int32_t f = A.fbk_iface[brick], pa = A.ifA[f], pb = A.ifB[f];
          const bool    two = pb >= 0; // an interior interface: both own sides and the cross matrices
          const int     lf = A.fbk_lf[brick], fd = lf >> 1, fs = lf & 1;
          const double *ba = A.bbox + (int64_t)pa * 2 * DIM, *bbx = A.bbox + (int64_t)(two ? pb : pa) * 2 * DIM;
          const double  a_lo = ba[d], a_ih = 1. / (ba[DIM + d] - ba[d]), b_lo = bbx[d], b_ih = 1. / (bbx[DIM + d] - bbx[d]);
          double        MA[NX], MB[NX], MX[NX]; // own side A | own side B | cross (rows A, columns B)
          if (d == fd)
            {
              const double x = A.fbk_plane[brick], pen = A.fbk_sigma[brick], nA = fs ? 1. : -1., cf = two ? 0.5 : 1.;
              double       LA[N1], dLA[N1], LB[N1], dLB[N1];
              basis_1d<C>(A.basis, (x - a_lo) * a_ih, a_ih, LA, dLA);
              basis_1d<C>(A.basis, (x - b_lo) * b_ih, b_ih, LB, dLB);
#pragma unroll
              for (int a = 0; a < N1; ++a)
#pragma unroll
                for (int k = 0; k < N1; ++k)
                  {
                    // M11 / boundary, M22 (outward normal of B = -nA), M12 (include/poly_utils.h:1891-1922), per unit stiffness
                    MA[a * N1 + k] = -cf * nA * (dLA[a] * LA[k] + LA[a] * dLA[k]) + pen * LA[a] * LA[k];
                    MB[a * N1 + k] = cf * nA * (dLB[a] * LB[k] + LB[a] * dLB[k]) + pen * LB[a] * LB[k];
                    MX[a * N1 + k] = 0.5 * nA * (dLA[a] * LB[k] - LA[a] * dLB[k]) - pen * LA[a] * LB[k];
                  }
            }
          else
            {
#pragma unroll
              for (int k = 0; k < NX; ++k)
                MA[k] = MB[k] = MX[k] = 0.;
              const int32_t *iv = A.fbk_iv + brick * 2 * (DIM - 1) + 2 * (d < fd ? d : d - 1);
              for (int i = iv[0]; i < iv[0] + iv[1]; ++i)
                {
                  const double2 bx = A.fiv_box[i];
                  for (int q = 0; q < A.nqf; ++q)
                    {
                      double       LA[N1], dLA[N1], LB[N1], dLB[N1];
                      const double x = bx.x + (bx.y - bx.x) * A.quadf.x[q];
                      basis_1d<C>(A.basis, (x - a_lo) * a_ih, a_ih, LA, dLA);
                      basis_1d<C>(A.basis, (x - b_lo) * b_ih, b_ih, LB, dLB);
                      const double wq = A.quadf.w[q] * (bx.y - bx.x);
#pragma unroll
                      for (int a = 0; a < N1; ++a)
#pragma unroll
                        for (int k = 0; k < N1; ++k)
                          {
                            MA[a * N1 + k] += wq * LA[a] * LA[k];
                            MB[a * N1 + k] += wq * LB[a] * LB[k];
                            MX[a * N1 + k] += wq * LA[a] * LB[k];
                          }
                    }
                }
            }
          double *dst = A.fmat + (brick * 3 * DIM + d) * NXP;
#pragma unroll
          for (int k = 0; k < NX; ++k)
            {
              dst[k] = MA[k];
              if (two)
                {
                  dst[DIM * NXP + k]     = MB[k];
                  dst[2 * DIM * NXP + k] = MX[k];
                }
            }
        }
    }

    // the chunk's 1-D matrices from k_brick_matrices into the group's shared memory (face factor x stiffness)
    template <int DIM, int DEGX, bool OFFDIAG>
    __device__ __forceinline__ void
    load_chunk(const CartArgs &A, const int64_t c0, const int cnt, const int t, double *SL, int *kind, const bool transpose = false)
    {
      using CC           = CartCfg<DIM, DEGX>;
      constexpr int NXP  = CC::NXP, ISTR = CC::ISTR, GROUP = CC::GROUP, N1 = CC::N1;
      for (int w = t; w < cnt * ISTR; w += GROUP)
        {
          const int it = w / ISTR, r = w % ISTR, d = r / (2 * NXP), slot = (r / NXP) & 1, k = r % NXP;
          double    v  = 0.;
          if (OFFDIAG)
            {
              const int64_t brick = c0 + it;
              if (r == 0)
                kind[it] = 2;
              if (slot == 0 && k < N1 * N1) // (B,A) block: M21 = M12^T, i.e. every 1-D factor transposed
                v = A.fmat[((brick * 3 + 2) * DIM + d) * NXP + (transpose ? (k % N1) * N1 + k / N1 : k)] *
                    (d == (A.fbk_lf[brick] >> 1) ? A.stiffness : 1.);
            }
          else
            {
              const int64_t brick = A.pit_brick[c0 + it];
              const int     meta  = A.pit_meta[c0 + it];
              const bool    on    = (meta & 3) == 1 ? (A.flags & PD_ASSEMBLE_VOLUME) != 0 :
                                    ((meta & 8) ? (A.flags & PD_ASSEMBLE_BOUNDARY) != 0 : (A.flags & PD_ASSEMBLE_INTERIOR) != 0);
              if (r == 0)
                kind[it] = on ? (meta & 3) : 0;
              if (on && (meta & 3) == 1)
                v = A.cmat[(brick * DIM + d) * 2 * NXP + slot * NXP + k];
              else if (on && slot == 0)
                v = A.fmat[((brick * 3 + ((meta >> 2) & 1)) * DIM + d) * NXP + k] * (d == (A.fbk_lf[brick] >> 1) ? A.stiffness : 1.);
            }
          SL[w] = v;
        }
    }

    template <int DIM, int DEGX>
    __global__ void __launch_bounds__(CartCfg<DIM, DEGX>::NTHR, CartCfg<DIM, DEGX>::MINB)
    k_cart_diag(const CartArgs A)
    {
      using C          = Cfg<DIM, DEGX>;
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, NF = CC::NF, CH = CC::CH, N = C::N, GROUP = CC::GROUP, CPT = CC::CPT;
      extern __shared__ __align__(16) double cart_smem[];
      const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
      double   *SL   = cart_smem + g * CC::GSM;
      int      *kind = reinterpret_cast<int *>(SL + CC::GSM - CH / 2 - 2);

      const int wpc = blockDim.x / GROUP;
      for (int64_t p = (int64_t)blockIdx.x * wpc + g; p < A.np_own; p += (int64_t)gridDim.x * wpc)
        {
          double acc[CPT][NX];
#pragma unroll
          for (int k = 0; k < CPT; ++k)
#pragma unroll
            for (int r = 0; r < NX; ++r)
              acc[k][r] = 0.;
          const int64_t base   = A.diag_base[p];
          const int     stride = A.row_stride[A.dof_block[p]];
          // the bricks of the polytope as one item list: its cell bricks, then the own-side face bricks of its adjacency
          const int64_t i0 = A.pit_ptr[p], i1 = A.pit_diag_end[p];
          for (int64_t c0 = i0; c0 < i1; c0 += CH)
            {
              const int cnt = (int)(i1 - c0 < CH ? i1 - c0 : CH);
              group_sync<GROUP>(); // the previous chunk / block has been consumed
              load_chunk<DIM, DEGX, false>(A, c0, cnt, t, SL, kind);
              group_sync<GROUP>();
              accumulate_chunk<DIM, DEGX>(SL, kind, cnt, t, A.stiffness, A.mass, acc);
            }
          if (!C::DGP)
            {
              store_block_direct<DIM, DEGX, false>(A.values + base, stride, t, acc);
              continue;
            }
          // ---- FE_AggloDGP (a sub-set of the tensor index set): registers -> shared tile -> the CSR rows
          group_sync<GROUP>();
          stage_block<DIM, DEGX>(SL, t, acc);
          group_sync<GROUP>();
          for (int idx = t; idx < N * N; idx += GROUP)
            {
              const int i = idx / N, j = idx - i * N;
              const int fi = A.dof_abc[i][0] + N1 * (A.dof_abc[i][1] + N1 * A.dof_abc[i][2]);
              const int fj = A.dof_abc[j][0] + N1 * (A.dof_abc[j][1] + N1 * A.dof_abc[j][2]);
              A.values[base + (int64_t)i * stride + j] = SL[fi * (NF + 1) + fj];
            }
        }
    }

    template <int DIM, int DEGX>
    __global__ void __launch_bounds__(CartCfg<DIM, DEGX>::NTHR, CartCfg<DIM, DEGX>::MINB)
    k_cart_offdiag(const CartArgs A)
    {
      using C          = Cfg<DIM, DEGX>;
      using CC         = CartCfg<DIM, DEGX>;
      constexpr int N1 = CC::N1, NX = CC::NX, NF = CC::NF, CH = CC::CH, N = C::N, GROUP = CC::GROUP, CPT = CC::CPT;
      extern __shared__ __align__(16) double cart_smem[];
      const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
      double   *SL   = cart_smem + g * CC::GSM;
      int      *kind = reinterpret_cast<int *>(SL + CC::GSM - CH / 2 - 2);

      const int wpc = blockDim.x / GROUP;
      // even N1: one work item per interface stores M12 and, transposed, M21; odd N1: two work items, the second one
      // computes M21 = M12^T from the transposed factors (cheap) so that both blocks are stored with the lanes along rows
      constexpr int     WPI     = CC::RA == 0 ? 1 : 2;
      for (int64_t w = (int64_t)blockIdx.x * wpc + g; w < (int64_t)A.n_ifaces * WPI; w += (int64_t)gridDim.x * wpc)
        {
          const int64_t f  = w / WPI;
          const bool    tr = WPI == 2 && (w & 1);
          const int32_t pa = A.ifA[f], pb = A.ifB[f];
          if (pb < 0)
            continue;
          // where the two blocks go: issued now, so that these dependent loads run under the matrix loads below
          const int64_t baseAB = A.if_baseAB[f], baseBA = A.if_baseBA[f];
          const int     strideA = A.row_stride[A.dof_block[pa]];
          const int     strideB = baseBA >= 0 ? A.row_stride[A.dof_block[pb]] : 0;
          if (tr && baseBA < 0)
            continue; // B a ghost polytope: its rows live on another rank
          double acc[CPT][NX];
#pragma unroll
          for (int k = 0; k < CPT; ++k)
#pragma unroll
            for (int r = 0; r < NX; ++r)
              acc[k][r] = 0.;
          const int64_t i0 = A.fbk_ptr[f], i1 = A.fbk_ptr[f + 1];
          for (int64_t c0 = i0; c0 < i1; c0 += CH)
            {
              const int cnt = (int)(i1 - c0 < CH ? i1 - c0 : CH);
              group_sync<GROUP>();
              load_chunk<DIM, DEGX, true>(A, c0, cnt, t, SL, kind, tr);
              group_sync<GROUP>();
              accumulate_chunk<DIM, DEGX>(SL, kind, cnt, t, 1., 0., acc);
            }
          if (!C::DGP)
            {
              if (WPI == 2)
                store_block_direct<DIM, DEGX, false>(A.values + (tr ? baseBA : baseAB), tr ? strideB : strideA, t, acc);
              else
                {
                  store_block_direct<DIM, DEGX, false>(A.values + baseAB, strideA, t, acc);
                  if (baseBA >= 0) // M21 = M12^T (B a ghost polytope: its rows live on another rank)
                    store_block_direct<DIM, DEGX, CC::RA == 0>(A.values + baseBA, strideB, t, acc);
                }
              continue;
            }
          group_sync<GROUP>();
          stage_block<DIM, DEGX>(SL, t, acc);
          group_sync<GROUP>();
          for (int idx = t; idx < N * N; idx += GROUP)
            {
              const int i = idx / N, j = idx - i * N;
              const int fi = A.dof_abc[i][0] + N1 * (A.dof_abc[i][1] + N1 * A.dof_abc[i][2]);
              const int fj = A.dof_abc[j][0] + N1 * (A.dof_abc[j][1] + N1 * A.dof_abc[j][2]);
              A.values[baseAB + (int64_t)i * strideA + j] = SL[fi * (NF + 1) + fj];
              if (baseBA >= 0) // M21 = M12^T; coalesced over j as well: entry (i, j) of (B,A) is M12(j, i)
                A.values[baseBA + (int64_t)i * strideB + j] = SL[fj * (NF + 1) + fi];
            }
        }
    }

    // -------------------------------------------------------------------------
    // matrix-free apply on agglomerates of axis-aligned cells: y_P = sum over the same items of
    // (X (x) Y (x) Z) x, by sum factorisation -- three 1-D contractions per item instead of a block.
    //   sub-cell      (sigma (K.M.M + M.K.M + M.M.K) + f M.M.M) x_P     7 N1^4 multiply-adds
    //   own face      (F in the slot of the normal axis, masses elsewhere) x_P
    //   cross face    the same with cross matrices of the two bounding-box bases, applied to x_Q
    // Thread (item j, line l): the j-th item of the chunk, one line of N1 coefficients per pass
    // (z-pass: line (a,b); y-pass: (a,c); x-pass: (b,c)); intermediates of an item live in shared
    // memory, the result line is accumulated in registers over the chunks and the item slots are
    // summed in a fixed order at the end.  FE_DGQ only.
    // -------------------------------------------------------------------------
    struct CartApplyArgs
    {
      CartArgs      g;
      const double *src;
      double       *dst;
      int           add;
    };

    template <int DIM, int DEGX>
    struct ApplyCfg
    {
      using C                   = Cfg<DIM, DEGX>;
      static constexpr int N1   = C::N1, NX = N1 * N1, NXP = NX + (NX & 1), NF = C::N;
      static constexpr int NL   = ipow(N1, DIM - 1); // lines per item
      // a WARP per polytope, WPC polytopes per CTA in flight, warp barriers only: the work per polytope is a few
      // thousand multiply-adds, what matters is how many polytopes an SM has in flight
      static constexpr int GROUP = 32;
      static constexpr int CH   = GROUP / NL;        // items a warp processes at once
      static constexpr int NTHR = 256, WPC = NTHR / GROUP;
      static constexpr int MSTR = DIM * 2 * NXP;     // 1-D matrices of an item
      static constexpr int WSTR = 2 * NF + 1;        // two intermediate tensors of an item (odd stride)
      static constexpr int GSM  = CH * (MSTR + 2 * WSTR) + 2 * CH + 2; // doubles of shared memory per warp
      static_assert(NL <= 32, "a line per lane");
    };

    template <int DIM, int DEGX>
    __global__ void __launch_bounds__(ApplyCfg<DIM, DEGX>::NTHR, 3)
    k_cart_apply(const CartApplyArgs P)
    {
      using AC         = ApplyCfg<DIM, DEGX>;
      constexpr int N1 = AC::N1, NXP = AC::NXP, NF = AC::NF, NL = AC::NL, CH = AC::CH, MSTR = AC::MSTR, WSTR = AC::WSTR,
                    GROUP = AC::GROUP;
      const CartArgs &A = P.g;
      extern __shared__ __align__(16) double smem[];
      const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
      double   *SM = smem + g * AC::GSM;   // [CH][MSTR]   1-D matrices, [d][slot][row a][a']
      double   *W  = SM + CH * MSTR;       // [CH][WSTR]   z-pass output (two tensors)
      double   *V  = W + CH * WSTR;        // [CH][WSTR]   y-pass output (two tensors)
      int64_t  *it_src  = reinterpret_cast<int64_t *>(V + CH * WSTR); // [CH] first source entry of the item's input block
      int      *it_kind = reinterpret_cast<int *>(it_src + CH);       // [CH] 1 cell, 2 own face, 3 cross face, 0 off
      const int  j = t / NL, l = t % NL;
      const bool active = t < CH * NL;
      const double sigma = A.stiffness, fmass = A.mass;

      const int wpc = blockDim.x / GROUP; // fewer warps per CTA on small problems: more CTAs, every SM busy
      for (int64_t p = (int64_t)blockIdx.x * wpc + g; p < A.np_own; p += (int64_t)gridDim.x * wpc)
        {
          double acc[N1];
#pragma unroll
          for (int k = 0; k < N1; ++k)
            acc[k] = 0.;
          // one item list per polytope: cell bricks, own-side face bricks, cross face bricks (x_Q of the neighbour)
          const int64_t i0 = A.pit_ptr[p], i1 = A.pit_ptr[p + 1];
          for (int64_t c0 = i0; c0 < i1; c0 += CH)
            {
              const int cnt = (int)(i1 - c0 < CH ? i1 - c0 : CH);
              __syncwarp();
              // ---- the 1-D matrices of the chunk's bricks (k_brick_matrices); seen from side B a cross matrix is
              // the transpose of the stored one
              for (int w = t; w < cnt * MSTR; w += GROUP)
                {
                  const int     it = w / MSTR, r = w % MSTR, d = r / (2 * NXP), slot = (r / NXP) & 1, k = r % NXP;
                  const int64_t brick = A.pit_brick[c0 + it];
                  const int     meta = A.pit_meta[c0 + it], knd = meta & 3, side = (meta >> 2) & 1;
                  const bool    enabled = knd == 1 ? (A.flags & PD_ASSEMBLE_VOLUME) != 0 :
                                          ((meta & 8) ? (A.flags & PD_ASSEMBLE_BOUNDARY) != 0 : (A.flags & PD_ASSEMBLE_INTERIOR) != 0);
                  if (r == 0)
                    {
                      it_kind[it] = enabled ? knd : 0;
                      it_src[it]  = (int64_t)A.dof_block[A.pit_q[c0 + it]] * NF;
                    }
                  double v = 0.;
                  if (enabled && knd == 1)
                    v = A.cmat[(brick * DIM + d) * 2 * NXP + slot * NXP + k];
                  else if (enabled && slot == 0 && k < N1 * N1)
                    {
                      const int kk = (knd == 3 && side) ? (k % N1) * N1 + k / N1 : k;
                      v = A.fmat[((brick * 3 + (knd == 3 ? 2 : side)) * DIM + d) * NXP + kk] *
                          (d == (A.fbk_lf[brick] >> 1) ? sigma : 1.);
                    }
                  SM[w] = v;
                }
              __syncwarp();
              const int     knd = (active && j < cnt) ? it_kind[j] : 0;
              const bool    on  = knd != 0;
              const int64_t src_base = on ? it_src[j] : 0;
              const double *m = SM + j * MSTR;
              double       *w1 = W + j * WSTR, *w2 = w1 + NF, *v1 = V + j * WSTR, *v2 = v1 + NF;
              double        in1[N1], in2[N1];
              if (DIM == 3)
                {
                  // ---- z-pass: line (a, b) = l
                  if (on)
                    {
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        in1[k] = P.src[src_base + l + NL * k];
#pragma unroll
                      for (int c = 0; c < N1; ++c)
                        {
                          double t1 = 0., t2 = 0.;
#pragma unroll
                          for (int k = 0; k < N1; ++k)
                            {
                              t1 += m[(2 * 2 + 0) * NXP + c * N1 + k] * in1[k];
                              if (knd == 1)
                                t2 += m[(2 * 2 + 1) * NXP + c * N1 + k] * in1[k];
                            }
                          w1[l + NL * c] = t1;
                          w2[l + NL * c] = t2;
                        }
                    }
                  __syncwarp();
                  // ---- y-pass: line (a, c): a = l % N1, c = l / N1
                  if (on)
                    {
                      const int a = l % N1, c = l / N1;
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        {
                          in1[k] = w1[a + N1 * k + NL * c];
                          in2[k] = w2[a + N1 * k + NL * c];
                        }
#pragma unroll
                      for (int b = 0; b < N1; ++b)
                        {
                          double aa = 0., ab = 0., ba = 0.;
#pragma unroll
                          for (int k = 0; k < N1; ++k)
                            {
                              const double my = m[(1 * 2 + 0) * NXP + b * N1 + k];
                              aa += my * in1[k];
                              if (knd == 1)
                                {
                                  ab += m[(1 * 2 + 1) * NXP + b * N1 + k] * in1[k];
                                  ba += my * in2[k];
                                }
                            }
                          v1[a + N1 * b + NL * c] = knd == 1 ? sigma * aa : 0.;
                          v2[a + N1 * b + NL * c] = knd == 1 ? sigma * (ab + ba) + fmass * aa : aa;
                        }
                    }
                  __syncwarp();
                }
              else
                {
                  // ---- 2-D y-pass: line a = l
                  if (on)
                    {
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        in1[k] = P.src[src_base + l + N1 * k];
#pragma unroll
                      for (int b = 0; b < N1; ++b)
                        {
                          double aa = 0., ab = 0.;
#pragma unroll
                          for (int k = 0; k < N1; ++k)
                            {
                              aa += m[(1 * 2 + 0) * NXP + b * N1 + k] * in1[k];
                              if (knd == 1)
                                ab += m[(1 * 2 + 1) * NXP + b * N1 + k] * in1[k];
                            }
                          v1[l + N1 * b] = knd == 1 ? sigma * aa : 0.;
                          v2[l + N1 * b] = knd == 1 ? sigma * ab + fmass * aa : aa;
                        }
                    }
                  __syncwarp();
                }
              // ---- x-pass: line (b, c) = l -> the result line, accumulated in registers
              if (on)
                {
#pragma unroll
                  for (int k = 0; k < N1; ++k)
                    {
                      in1[k] = v1[k + N1 * l];
                      in2[k] = v2[k + N1 * l];
                    }
#pragma unroll
                  for (int ap = 0; ap < N1; ++ap)
                    {
                      double tt = 0.;
#pragma unroll
                      for (int k = 0; k < N1; ++k)
                        {
                          tt += m[(0 * 2 + 0) * NXP + ap * N1 + k] * in2[k];
                          if (knd == 1)
                            tt += m[(0 * 2 + 1) * NXP + ap * N1 + k] * in1[k];
                        }
                      acc[ap] += tt;
                    }
                }
            }
          // ---- sum the item slots in a fixed order and write the rows of the polytope
          __syncwarp();
          if (active)
            {
#pragma unroll
              for (int k = 0; k < N1; ++k)
                W[j * (NF + 1) + k + N1 * l] = acc[k];
            }
          __syncwarp();
          for (int i = t; i < NF; i += GROUP)
            {
              double tt = 0.;
              for (int jj = 0; jj < CH; ++jj)
                tt += W[jj * (NF + 1) + i];
              double *o = P.dst + (int64_t)A.dof_block[p] * NF + i;
              *o        = P.add ? *o + tt : tt;
            }
        }
    }

    template <int DIM, int DEGX>
    void
    run_cart_apply(pd_handle *h, const CartApplyArgs &a)
    {
      using AC = ApplyCfg<DIM, DEGX>;
      auto         kern = k_cart_apply<DIM, DEGX>;
      static_assert(AC::WSTR >= AC::NF + 1, "the slot sums alias the intermediates");
      // warps per CTA: eight on large problems, fewer when that would leave SMs without a CTA
      int wpc = AC::WPC;
      while (wpc > 1 && (h->np_own + wpc - 1) / wpc < (int64_t)h->sm_count * 4)
        wpc /= 2;
      const size_t smem = sizeof(double) * (size_t)wpc * AC::GSM;
      PD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * AC::WPC * AC::GSM)));
      const int64_t ctas = (h->np_own + wpc - 1) / wpc;
      const int     grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)h->sm_count * 16));
      kern<<<grid, wpc * AC::GROUP, smem, h->stream>>>(a);
      ++h->launches;
      PD_CUDA(cudaGetLastError());
    }

    template <int DIM, int DEGX>
    void
    run_brick_matrices(pd_handle *h, const CartArgs &a)
    {
      using CC           = CartCfg<DIM, DEGX>;
      const int64_t work = (a.n_cbk + a.n_fbk) * DIM;
      const int     grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 127) / 128, (int64_t)h->sm_count * 32));
      k_brick_matrices<DIM, DEGX><<<grid, 128, 0, h->stream>>>(a);
      ++h->launches;
      PD_CUDA(cudaGetLastError());
    }

    template <int DIM, int DEGX>
    void
    run_cart(pd_handle *h, const CartArgs &a)
    {
      using CC = CartCfg<DIM, DEGX>;
      auto kd = k_cart_diag<DIM, DEGX>, ko = k_cart_offdiag<DIM, DEGX>;
      const size_t smem_max = sizeof(double) * (size_t)CC::WPC * CC::GSM;
      if (smem_max > 48 * 1024)
        {
          PD_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
          PD_CUDA(cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        }
      // work items per CTA (warp-granular kernels): eight on large problems, fewer when that would leave SMs idle
      auto launch = [&](auto kern, const int64_t n_items) {
        int wpc = CC::WPC;
        while (wpc > 1 && (n_items + wpc - 1) / wpc < (int64_t)h->sm_count * 4)
          wpc /= 2;
        const int64_t ctas = (n_items + wpc - 1) / wpc;
        const int     grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)h->sm_count * 16));
        kern<<<grid, wpc * CC::GROUP, sizeof(double) * (size_t)wpc * CC::GSM, h->stream>>>(a);
        ++h->launches;
      };
      // The two kernels write disjoint blocks and bound differently (the diagonal one by latency, the off-diagonal
      // one by the HBM write stream): they run CONCURRENTLY, the off-diagonal kernel on a second stream of the handle.
      // PD_CART_SERIAL=1 runs them one after the other (per-kernel timings for the rooflines).
      const bool        serial  = getenv("PD_CART_SERIAL") != nullptr;
      const bool        offdiag = (a.flags & PD_ASSEMBLE_INTERIOR) && h->n_ifaces > 0;
      const int64_t n_off_items = (int64_t)h->n_ifaces * (CC::RA == 0 ? 1 : 2);
      if (offdiag && !serial)
        {
          if (!h->aux_stream)
            {
              PD_CUDA(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
              PD_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
              PD_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
            }
          PD_CUDA(cudaEventRecord(h->ev_fork, h->stream));
          PD_CUDA(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
          cudaStream_t main_stream = h->stream;
          h->stream                = h->aux_stream;
          launch(ko, n_off_items);
          h->stream = main_stream;
          PD_CUDA(cudaEventRecord(h->ev_join, h->aux_stream));
          launch(kd, h->np_own);
          PD_CUDA(cudaEventRecord(h->ev[1], h->stream)); // (the diagonal kernel's end; the other kernel overlaps it)
          PD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        }
      else
        {
          launch(kd, h->np_own);
          PD_CUDA(cudaEventRecord(h->ev[1], h->stream));
          if (offdiag)
            launch(ko, n_off_items);
        }
      PD_CUDA(cudaEventRecord(h->ev[2], h->stream));
      PD_CUDA(cudaEventRecord(h->ev[3], h->stream));
      PD_CUDA(cudaGetLastError());
    }
  } // namespace

  // ---------------------------------------------------------------------------------------------------
  namespace
  {
    struct BrickBuilder
    {
      const pd_mesh_desc &d;
      const int           dim, vpc;
      std::vector<int64_t> cbk_ptr, fbk_ptr;
      std::vector<int32_t> cbk_iv, civ, cpos, fbk_s, fbk_iv, fiv, fpos;
      explicit BrickBuilder(const pd_mesh_desc &desc)
        : d(desc)
        , dim(desc.dim)
        , vpc(1 << desc.dim)
      {}
      void
      box(const int32_t c, const int k, double &lo, double &hi) const
      {
        const int32_t *cv = d.cell_verts + (size_t)c * vpc;
        lo                = d.verts[(size_t)cv[0] * dim + k];
        hi                = d.verts[(size_t)cv[vpc - 1] * dim + k];
      }
      // indices of the intervals of `cells` along axis k among the distinct ones; rep[i] = a cell with index i
      int
      index_axis(const int32_t *cells, const int n, const int k, std::vector<int> &idx, std::vector<int32_t> &rep) const
      {
        std::vector<std::pair<std::pair<double, double>, int>> key(n);
        for (int i = 0; i < n; ++i)
          {
            box(cells[i], k, key[i].first.first, key[i].first.second);
            key[i].second = i;
          }
        std::sort(key.begin(), key.end());
        idx.assign(n, 0);
        rep.clear();
        for (int i = 0; i < n; ++i)
          {
            if (i == 0 || key[i].first != key[i - 1].first)
              rep.push_back(cells[key[i].second]);
            idx[key[i].second] = (int)rep.size() - 1;
          }
        return (int)rep.size();
      }
      // Decompose the index set {(i0[e], i1[e], i2[e])} (all distinct) into tensor-product bricks.  Emits, per
      // brick, the index lists per axis through `emit(S0, S1, S2)`; ndim = 2 ignores i2.
      template <class Emit>
      void
      decompose(const int n, const std::vector<int> *ix, const int *nx, const int ndim, Emit &&emit) const
      {
        const int64_t full = (int64_t)nx[0] * (ndim > 1 ? nx[1] : 1) * (ndim > 2 ? nx[2] : 1);
        if (full == n)
          { // a full tensor grid: one brick
            std::vector<int> S[3];
            for (int k = 0; k < ndim; ++k)
              {
                S[k].resize(nx[k]);
                for (int i = 0; i < nx[k]; ++i)
                  S[k][i] = i;
              }
            for (int k = ndim; k < 3; ++k)
              S[k].assign(1, 0);
            emit(S[0], S[1], S[2]);
            return;
          }
        // pencils along axis 0: (i1, i2) -> sorted set of i0
        std::map<std::pair<int, int>, std::vector<int>> pencil;
        for (int e = 0; e < n; ++e)
          pencil[{ndim > 2 ? ix[2][e] : 0, ndim > 1 ? ix[1][e] : 0}].push_back(ix[0][e]);
        // slabs: for every i2, the i1's with the same pencil set
        std::map<std::pair<int, std::vector<int>>, std::vector<int>> slab; // (i2, S0) -> S1
        for (auto &pc : pencil)
          {
            std::sort(pc.second.begin(), pc.second.end());
            slab[{pc.first.first, pc.second}].push_back(pc.first.second);
          }
        // bricks: the i2's with the same (S0, S1)
        std::map<std::pair<std::vector<int>, std::vector<int>>, std::vector<int>> brick;
        for (auto &sl : slab)
          brick[{sl.first.second, sl.second}].push_back(sl.first.first);
        for (auto &bk : brick)
          emit(bk.first.first, bk.first.second, bk.second);
      }
      void
      build(const int32_t np_own)
      {
        cbk_ptr.assign(1, 0);
        cpos.assign((size_t)d.poly_subcell_ptr[np_own] * (1 + dim), 0);
        std::vector<int>     ix[3];
        std::vector<int32_t> rep[3];
        for (int32_t p = 0; p < np_own; ++p)
          {
            const int64_t  s0 = d.poly_subcell_ptr[p];
            const int      n  = (int)(d.poly_subcell_ptr[p + 1] - s0);
            const int32_t *cells = d.poly_subcell_idx + s0;
            if (n == 1)
              { // one cell: one brick (fine meshes)
                const int32_t b = (int32_t)(cbk_iv.size() / (2 * dim));
                for (int k = 0; k < dim; ++k)
                  {
                    cbk_iv.push_back((int32_t)civ.size());
                    cbk_iv.push_back(1);
                    civ.push_back(cells[0]);
                  }
                cpos[(size_t)s0 * (1 + dim)] = b;
                cbk_ptr.push_back((int64_t)(cbk_iv.size() / (2 * dim)));
                continue;
              }
            int nx[3] = {1, 1, 1};
            for (int k = 0; k < dim; ++k)
              nx[k] = index_axis(cells, n, k, ix[k], rep[k]);
            // (i0, i1, i2) -> position in the polytope's list, to write cpos
            std::map<std::array<int, 3>, int> where;
            const bool                        dense = (int64_t)nx[0] * nx[1] * nx[2] <= (int64_t)4 * n + 64;
            std::vector<int>                  grid;
            if (dense)
              grid.assign((size_t)nx[0] * nx[1] * nx[2], -1);
            for (int e = 0; e < n; ++e)
              {
                const int i0 = ix[0][e], i1 = dim > 1 ? ix[1][e] : 0, i2 = dim > 2 ? ix[2][e] : 0;
                if (dense)
                  grid[((size_t)i2 * nx[1] + i1) * nx[0] + i0] = e;
                else
                  where[{i0, i1, i2}] = e;
              }
            decompose(n, ix, nx, dim, [&](const std::vector<int> &S0, const std::vector<int> &S1, const std::vector<int> &S2) {
              const int32_t           b = (int32_t)(cbk_iv.size() / (2 * dim));
              const std::vector<int> *S[3] = {&S0, &S1, &S2};
              for (int k = 0; k < dim; ++k)
                {
                  cbk_iv.push_back((int32_t)civ.size());
                  cbk_iv.push_back((int32_t)S[k]->size());
                  for (const int i : *S[k])
                    civ.push_back(rep[k][i]);
                }
              for (size_t c = 0; c < S2.size(); ++c)
                for (size_t bb = 0; bb < S1.size(); ++bb)
                  for (size_t a = 0; a < S0.size(); ++a)
                    {
                      const int e = dense ? grid[((size_t)S2[c] * nx[1] + S1[bb]) * nx[0] + S0[a]] :
                                            where.at({S0[a], S1[bb], S2[c]});
                      int32_t  *ps = &cpos[(size_t)(s0 + e) * (1 + dim)];
                      ps[0]        = b;
                      ps[1]        = (int32_t)a;
                      if (dim > 1)
                        ps[2] = (int32_t)bb;
                      if (dim > 2)
                        ps[3] = (int32_t)c;
                    }
            });
            cbk_ptr.push_back((int64_t)(cbk_iv.size() / (2 * dim)));
          }
        // faces: the sub-faces of an interface grouped by (orientation, plane, penalty), each group into rectangles
        fbk_ptr.assign(1, 0);
        const int64_t nsf = d.n_ifaces ? d.iface_sub_ptr[d.n_ifaces] : 0;
        fpos.assign((size_t)nsf * dim, 0);
        const int td = dim - 1;
        for (int32_t f = 0; f < d.n_ifaces; ++f)
          {
            const int64_t s0 = d.iface_sub_ptr[f], s1 = d.iface_sub_ptr[f + 1];
            if (s1 - s0 == 1)
              { // one sub-face: one brick
                fpos[(size_t)s0 * dim] = (int32_t)fbk_s.size();
                fbk_s.push_back((int32_t)s0);
                for (int t = 0; t < td; ++t)
                  {
                    fbk_iv.push_back((int32_t)fiv.size());
                    fbk_iv.push_back(1);
                    fiv.push_back(d.sub_cell[s0]);
                  }
                fbk_ptr.push_back((int64_t)fbk_s.size());
                continue;
              }
            // group key: (local face, plane coordinate, penalty)
            std::vector<std::pair<std::tuple<int, double, double>, int64_t>> key;
            for (int64_t s = s0; s < s1; ++s)
              {
                const int lf = d.sub_face[s];
                double    lo, hi;
                box(d.sub_cell[s], lf >> 1, lo, hi);
                key.push_back({{lf, (lf & 1) ? hi : lo, d.sub_sigma[s]}, s});
              }
            std::sort(key.begin(), key.end());
            for (size_t g0 = 0; g0 < key.size();)
              {
                size_t g1 = g0;
                while (g1 < key.size() && key[g1].first == key[g0].first)
                  ++g1;
                const int            n  = (int)(g1 - g0);
                const int            fd = std::get<0>(key[g0].first) >> 1;
                std::vector<int32_t> cells(n);
                for (int e = 0; e < n; ++e)
                  cells[e] = d.sub_cell[key[g0 + e].second];
                int taxis[2] = {0, 0};
                for (int k = 0, t = 0; k < dim; ++k)
                  if (k != fd)
                    taxis[t++] = k;
                int nx[3] = {1, 1, 1};
                for (int t = 0; t < td; ++t)
                  nx[t] = index_axis(cells.data(), n, taxis[t], ix[t], rep[t]);
                std::map<std::pair<int, int>, int> where;
                for (int e = 0; e < n; ++e)
                  where[{ix[0][e], td > 1 ? ix[1][e] : 0}] = e;
                decompose(n, ix, nx, td, [&](const std::vector<int> &S0, const std::vector<int> &S1, const std::vector<int> &) {
                  const int32_t           b = (int32_t)fbk_s.size();
                  const std::vector<int> *S[2] = {&S0, &S1};
                  fbk_s.push_back((int32_t)key[g0 + where.at({S0[0], S1[0]})].second);
                  for (int t = 0; t < td; ++t)
                    {
                      fbk_iv.push_back((int32_t)fiv.size());
                      fbk_iv.push_back((int32_t)S[t]->size());
                      for (const int i : *S[t])
                        fiv.push_back(rep[t][i]);
                    }
                  for (size_t bb = 0; bb < S1.size(); ++bb)
                    for (size_t a = 0; a < S0.size(); ++a)
                      {
                        const int64_t s  = key[g0 + where.at({S0[a], S1[bb]})].second;
                        int32_t      *ps = &fpos[(size_t)s * dim];
                        ps[0]            = b;
                        ps[1]            = (int32_t)a;
                        if (td > 1)
                          ps[2] = (int32_t)bb;
                      }
                });
                g0 = g1;
              }
            fbk_ptr.push_back((int64_t)fbk_s.size());
          }
        // item lists: per owned polytope its cell bricks, the own-side face bricks of its adjacency (interfaces in
        // list order, A side then B side as in pd_create's adjacency), then the cross face bricks
        std::vector<std::vector<int64_t>> adj(np_own);
        for (int32_t f = 0; f < d.n_ifaces; ++f)
          {
            adj[d.iface_polyA[f]].push_back((int64_t)f * 2);
            if (d.iface_polyB[f] >= 0 && d.iface_polyB[f] < np_own)
              adj[d.iface_polyB[f]].push_back((int64_t)f * 2 + 1);
          }
        cbk_poly.resize((size_t)cbk_ptr[np_own]);
        for (int32_t p = 0; p < np_own; ++p)
          for (int64_t b = cbk_ptr[p]; b < cbk_ptr[p + 1]; ++b)
            cbk_poly[(size_t)b] = p;
        fbk_iface.resize(fbk_s.size());
        for (int32_t f = 0; f < d.n_ifaces; ++f)
          for (int64_t b = fbk_ptr[f]; b < fbk_ptr[f + 1]; ++b)
            fbk_iface[(size_t)b] = f;
        pit_ptr.assign(1, 0);
        pit_diag_end.clear();
        for (int32_t p = 0; p < np_own; ++p)
          {
            for (int64_t b = cbk_ptr[p]; b < cbk_ptr[p + 1]; ++b)
              {
                pit_brick.push_back((int32_t)b);
                pit_meta.push_back(1);
                pit_q.push_back(p);
              }
            for (const int64_t e : adj[p])
              {
                const int64_t f   = e >> 1;
                const int     bnd = d.iface_polyB[f] < 0;
                for (int64_t b = fbk_ptr[f]; b < fbk_ptr[f + 1]; ++b)
                  {
                    pit_brick.push_back((int32_t)b);
                    pit_meta.push_back(2 | ((int)(e & 1) << 2) | (bnd << 3));
                    pit_q.push_back(p);
                  }
              }
            pit_diag_end.push_back((int64_t)pit_brick.size());
            for (const int64_t e : adj[p])
              {
                const int64_t f = e >> 1;
                if (d.iface_polyB[f] < 0)
                  continue;
                for (int64_t b = fbk_ptr[f]; b < fbk_ptr[f + 1]; ++b)
                  {
                    pit_brick.push_back((int32_t)b);
                    pit_meta.push_back(3 | ((int)(e & 1) << 2));
                    pit_q.push_back((e & 1) ? d.iface_polyA[f] : d.iface_polyB[f]);
                  }
              }
            pit_ptr.push_back((int64_t)pit_brick.size());
          }
      }
      std::vector<int64_t> pit_ptr, pit_diag_end;
      std::vector<int32_t> pit_brick, pit_meta, pit_q, cbk_poly, fbk_iface;
    };
  } // namespace

  // host: decompose into bricks and upload (pd_create; pd_upload when the device scan finds the bricks stale)
  void
  build_cartesian_bricks(pd_handle *h, const pd_mesh_desc &d)
  {
    BrickBuilder B(d);
    B.build(h->np_own);
    auto put = [](auto &buf, const auto &v) {
      buf.alloc(std::max<size_t>(v.size(), 1));
      if (!v.empty())
        PD_CUDA(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    PD_CUDA(cudaStreamSynchronize(h->stream));
    put(h->cbk_ptr, B.cbk_ptr);
    put(h->cbk_iv, B.cbk_iv);
    put(h->civ, B.civ);
    put(h->cpos, B.cpos);
    put(h->fbk_ptr, B.fbk_ptr);
    put(h->fbk_s, B.fbk_s);
    put(h->fbk_iv, B.fbk_iv);
    put(h->fiv, B.fiv);
    put(h->fpos, B.fpos);
    put(h->pit_ptr, B.pit_ptr);
    put(h->pit_diag_end, B.pit_diag_end);
    put(h->pit_brick, B.pit_brick);
    put(h->pit_meta, B.pit_meta);
    put(h->pit_q, B.pit_q);
    put(h->cbk_poly, B.cbk_poly);
    put(h->fbk_iface, B.fbk_iface);
    {
      const int    nxp = h->n1 * h->n1 + ((h->n1 * h->n1) & 1);
      const size_t nc = std::max<size_t>(B.cbk_poly.size() * h->dim * 2 * nxp, 1), nf = std::max<size_t>(B.fbk_s.size() * 3 * h->dim * nxp, 1);
      h->cmat.alloc(nc);
      h->fmat.alloc(nf);
      PD_CUDA(cudaMemset(h->cmat.p, 0, nc * sizeof(double)));
      PD_CUDA(cudaMemset(h->fmat.p, 0, nf * sizeof(double)));
    }
    h->civ_box.alloc(std::max<size_t>(B.civ.size(), 1));
    h->fiv_box.alloc(std::max<size_t>(B.fiv.size(), 1));
    h->fbk_plane.alloc(std::max<size_t>(B.fbk_s.size(), 1));
    h->fbk_sigma.alloc(std::max<size_t>(B.fbk_s.size(), 1));
    h->fbk_lf.alloc(std::max<size_t>(B.fbk_s.size(), 1));
    h->n_cell_bricks = (int64_t)B.cbk_iv.size() / (2 * h->dim);
    h->n_face_bricks = (int64_t)B.fbk_s.size();
    h->n_diag_items  = 0;
    for (int32_t p = 0; p < h->np_own; ++p)
      h->n_diag_items += B.pit_diag_end[p] - B.pit_ptr[p];
    h->n_apply_items = B.pit_ptr[h->np_own];
    h->bricks_ready  = true;
    h->h_subcell_idx.assign(d.poly_subcell_idx, d.poly_subcell_idx + h->n_subcells);
    h->h_sub_cell.assign(d.sub_cell, d.sub_cell + h->n_subfaces);
    h->h_sub_face.assign(d.sub_face, d.sub_face + h->n_subfaces);
  }

  // Every owned sub-cell an axis-aligned box?  With bricks: do they still describe the mesh?  (device scan of the
  // uploaded arrays; one 8-byte read-back.)  Returns the axis-alignment; *bricks_stale reports the second answer.
  bool
  check_axis_aligned(pd_handle *h, bool *bricks_stale)
  {
    if (bricks_stale)
      *bricks_stale = false;
    if (h->n_subcells == 0)
      return false;
    if (!h->cart_flag.p)
      h->cart_flag.alloc(2);
    PD_CUDA(cudaMemsetAsync(h->cart_flag.p, 0, 2 * sizeof(int), h->stream));
    const bool with_bricks = h->bricks_ready;
    const int  grid        = (int)std::min<int64_t>((std::max(h->n_subcells, h->n_subfaces) + 255) / 256, (int64_t)h->sm_count * 8);
    k_check_axis_aligned<<<grid, 256, 0, h->stream>>>(h->verts.p, h->cell_verts.p, h->subcell_idx.p, h->n_subcells, h->dim,
                                                       h->cart_flag.p, with_bricks ? h->cpos.p : nullptr, h->cbk_iv.p, h->civ.p,
                                                       h->n_subfaces, h->sub_cell.p, h->sub_face.p, h->sub_sigma.p,
                                                       with_bricks ? h->fpos.p : nullptr, h->fbk_s.p, h->fbk_iv.p, h->fiv.p);
    int flags[2] = {1, 1};
    PD_CUDA(cudaMemcpyAsync(flags, h->cart_flag.p, sizeof flags, cudaMemcpyDeviceToHost, h->stream));
    PD_CUDA(cudaStreamSynchronize(h->stream));
    if (bricks_stale)
      *bricks_stale = flags[1] != 0;
    return flags[0] == 0;
  }

  // pd_create / pd_upload: which assembly path applies, bricks (re)built when needed
  void
  setup_cartesian(pd_handle *h, const pd_mesh_desc &d)
  {
    bool stale = false;
    // index arrays the bricks were built from: a changed list is a rebuild whatever the coordinates say
    if (h->bricks_ready &&
        (std::memcmp(h->h_subcell_idx.data(), d.poly_subcell_idx, sizeof(int32_t) * (size_t)h->n_subcells) != 0 ||
         (h->n_subfaces > 0 && (std::memcmp(h->h_sub_cell.data(), d.sub_cell, sizeof(int32_t) * (size_t)h->n_subfaces) != 0 ||
                                std::memcmp(h->h_sub_face.data(), d.sub_face, sizeof(int32_t) * (size_t)h->n_subfaces) != 0))))
      h->bricks_ready = false;
    h->cartesian = check_axis_aligned(h, &stale);
    if (!h->cartesian)
      return;
    if (!h->bricks_ready || stale)
      build_cartesian_bricks(h, d);
    h->brick_mat_valid = false; // coordinates / penalties may have changed: the 1-D matrices are rebuilt on next use
  }

  bool
  cartesian_assembly_selected(const pd_handle *h)
  {
    const char *e = getenv("PD_ASSEMBLE_KERNELS"); // "generic": the DMMA kernels everywhere (A/B measurements, tests)
    if (e && e[0] == 'g')
      return false;
    return h->cartesian;
  }

  static void
  fill_cart_args(pd_handle *h, const uint32_t flags, const pd_coefficients &coef, CartArgs &a);

#define PD_CART_DISPATCH(FN, h, ...)                                                                     \
  switch ((h)->fe_kind * 100 + (h)->dim * 10 + (h)->degree)                                                \
    {                                                                                                      \
      case 121: FN<2, DGP_BASE + 1>(__VA_ARGS__); break;                                                   \
      case 122: FN<2, DGP_BASE + 2>(__VA_ARGS__); break;                                                   \
      case 123: FN<2, DGP_BASE + 3>(__VA_ARGS__); break;                                                   \
      case 124: FN<2, DGP_BASE + 4>(__VA_ARGS__); break;                                                   \
      case 131: FN<3, DGP_BASE + 1>(__VA_ARGS__); break;                                                   \
      case 132: FN<3, DGP_BASE + 2>(__VA_ARGS__); break;                                                   \
      case 133: FN<3, DGP_BASE + 3>(__VA_ARGS__); break;                                                   \
      case 21: FN<2, 1>(__VA_ARGS__); break;                                                               \
      case 22: FN<2, 2>(__VA_ARGS__); break;                                                               \
      case 23: FN<2, 3>(__VA_ARGS__); break;                                                               \
      case 24: FN<2, 4>(__VA_ARGS__); break;                                                               \
      case 31: FN<3, 1>(__VA_ARGS__); break;                                                               \
      case 32: FN<3, 2>(__VA_ARGS__); break;                                                               \
      case 33: FN<3, 3>(__VA_ARGS__); break;                                                               \
      default:                                                                                             \
        throw CudaError{cudaErrorNotSupported, "no tensor kernel for this (dim, degree)", __LINE__};       \
    }

  // the cached geometry of the bricks and their 1-D matrices, rebuilt after every change of the geometry
  // (pd_create, pd_upload, pd_invalidate_quadrature: a benchmark step that starts from the flattened mesh pays for it)
  static void
  ensure_brick_matrices(pd_handle *h, const CartArgs &a)
  {
    if (h->brick_mat_valid)
      return;
    const int64_t work = std::max<int64_t>(h->n_cell_bricks * h->dim, h->n_face_bricks);
    const int     grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)h->sm_count * 8));
    k_refresh_bricks<<<grid, 256, 0, h->stream>>>(h->verts.p, h->cell_verts.p, h->dim, h->n_cell_bricks, h->cbk_iv.p, h->civ.p,
                                                   h->civ_box.p, h->n_face_bricks, h->fbk_s.p, h->fbk_iv.p, h->fiv.p, h->fiv_box.p,
                                                   h->sub_cell.p, h->sub_face.p, h->sub_sigma.p, h->fbk_plane.p, h->fbk_sigma.p,
                                                   h->fbk_lf.p);
    ++h->launches;
    PD_CUDA(cudaGetLastError());
    PD_CART_DISPATCH(run_brick_matrices, h, h, a);
    h->brick_mat_valid = true;
  }

  // matrix-free apply of the operator pd_set_operator describes on agglomerates of axis-aligned cells
  bool
  cartesian_apply_available(const pd_handle *h)
  {
    const char *e = getenv("PD_POLY_APPLY"); // "pointwise": the point-wise kernels everywhere (A/B measurements, tests)
    if (e && e[0] == 'p')
      return false;
    return h->cartesian && h->fe_kind == PD_FE_DGQ;
  }

  void
  launch_cart_apply(pd_handle *h, const double *src, double *dst, const bool add)
  {
    CartApplyArgs a{};
    fill_cart_args(h, h->op_flags, h->op_coef, a.g);
    ensure_brick_matrices(h, a.g);
    a.src = src;
    a.dst = dst;
    a.add = add ? 1 : 0;
    switch (h->dim * 10 + h->degree)
      {
        case 21: run_cart_apply<2, 1>(h, a); break;
        case 22: run_cart_apply<2, 2>(h, a); break;
        case 23: run_cart_apply<2, 3>(h, a); break;
        case 24: run_cart_apply<2, 4>(h, a); break;
        case 31: run_cart_apply<3, 1>(h, a); break;
        case 32: run_cart_apply<3, 2>(h, a); break;
        case 33: run_cart_apply<3, 3>(h, a); break;
        default:
          throw CudaError{cudaErrorNotSupported, "no tensor apply kernel for this (dim, degree)", __LINE__};
      }
  }

  void
  launch_assemble_cartesian(pd_handle *h, const uint32_t flags, const pd_coefficients &coef)
  {
    CartArgs a{};
    fill_cart_args(h, flags, coef, a);
    PD_CUDA(cudaEventRecord(h->ev[0], h->stream)); // the 1-D matrices count as part of the diagonal kernel's slot
    ensure_brick_matrices(h, a);
    PD_CART_DISPATCH(run_cart, h, h, a);
  }

  static void
  fill_cart_args(pd_handle *h, const uint32_t flags, const pd_coefficients &coef, CartArgs &a)
  {
    a.verts       = h->verts.p;
    a.cell_verts  = h->cell_verts.p;
    a.subcell_ptr = h->subcell_ptr.p;
    a.subcell_idx = h->subcell_idx.p;
    a.bbox        = h->bbox.p;
    a.ifA         = h->ifA.p;
    a.ifB         = h->ifB.p;
    a.if_sub_ptr  = h->if_sub_ptr.p;
    a.sub_cell    = h->sub_cell.p;
    a.sub_face    = h->sub_face.p;
    a.sub_sigma   = h->sub_sigma.p;
    a.padj_ptr    = h->padj_ptr.p;
    a.padj        = h->padj.p;
    a.diag_base   = h->diag_base.p;
    a.if_baseAB   = h->if_baseAB.p;
    a.if_baseBA   = h->if_baseBA.p;
    a.dof_block   = h->dof_block.p;
    a.row_stride  = h->row_stride.p;
    a.values      = h->values.p;
    a.stiffness   = coef.stiffness;
    a.mass        = coef.mass;
    a.flags       = flags;
    a.np_own      = h->np_own;
    a.n_ifaces    = h->n_ifaces;
    a.nq          = h->nq1;
    a.nqf         = h->nq1f;
    a.basis       = h->basis;
    a.quad        = h->quad;
    a.quadf       = h->quadf;
    a.cbk_ptr     = h->cbk_ptr.p;
    a.cbk_iv      = h->cbk_iv.p;
    a.civ         = h->civ.p;
    a.fbk_ptr     = h->fbk_ptr.p;
    a.fbk_s       = h->fbk_s.p;
    a.fbk_iv      = h->fbk_iv.p;
    a.fiv         = h->fiv.p;
    a.civ_box      = h->civ_box.p;
    a.fiv_box      = h->fiv_box.p;
    a.fbk_plane    = h->fbk_plane.p;
    a.fbk_sigma    = h->fbk_sigma.p;
    a.fbk_lf       = h->fbk_lf.p;
    a.pit_ptr      = h->pit_ptr.p;
    a.pit_diag_end = h->pit_diag_end.p;
    a.pit_brick    = h->pit_brick.p;
    a.pit_meta     = h->pit_meta.p;
    a.pit_q        = h->pit_q.p;
    a.cbk_poly     = h->cbk_poly.p;
    a.fbk_iface    = h->fbk_iface.p;
    a.cmat         = h->cmat.p;
    a.fmat         = h->fmat.p;
    a.n_cbk        = h->n_cell_bricks;
    a.n_fbk        = h->n_face_bricks;
    {
      // (a, b, c) of DoF i: FE_DGQ lexicographic; FE_AggloDGP in PolynomialSpace order (last coordinate outermost,
      // first fastest, total degree <= p: source/fe_agglodgp.cc:28-57)
      const int n1 = h->n1, p = h->degree;
      int       i  = 0;
      for (int c = 0; c < (h->dim == 3 ? n1 : 1); ++c)
        for (int b = 0; b < n1; ++b)
          for (int aa = 0; aa < n1; ++aa)
            {
              if (h->fe_kind == PD_FE_AGGLODGP && aa + b + c > p)
                continue;
              a.dof_abc[i][0] = (unsigned char)aa;
              a.dof_abc[i][1] = (unsigned char)b;
              a.dof_abc[i][2] = (unsigned char)c;
              ++i;
            }
    }
  }
} // namespace pd
