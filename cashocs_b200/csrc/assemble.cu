// Cell-batched P1 element kernels with precomputed cell->nnz scatter maps.
//
// Replaces the dolfin cell loop behind assemble_petsc_system
// (cashocs/_utils/linalg.py:107-188 -> fenics.assemble_system), SystemAssembler.assemble
// (cashocs/_forms/shape_form_handler.py:737, cashocs/_pde_problems/shape_gradient_problem.py:142)
// and the FFC-generated tabulate_tensor kernels.  One thread = one cell.  Cells arrive
// colour-sorted; cells of one colour share no dof, so the read-modify-write scatter into the
// CSR values needs no atomics and the summation order (colour 0, 1, 2, ...) is fixed: results
// are bit-reproducible run to run.
#include "common.cuh"

namespace cb {

template <int D>
struct CellData {
  int v[D + 1];
  double x[D + 1][D];
  Geo<D> g;
};

template <int D>
__device__ __forceinline__ void load_cell_data(const int32_t* __restrict__ cells,
                                               const double* __restrict__ coords, int64_t c,
                                               CellData<D>& cd) {
  load_cell<D>(cells, c, cd.v);
  load_coords<D>(coords, cd.v, cd.x);
  simplex_geometry(cd.x, cd.g);
}

// ------------------------------------------------------------------ scalar P1
template <int D, class P, int DEG, bool ZDEP>
__global__ void __launch_bounds__(128)
k_assemble_scalar(int64_t c_begin, int64_t c_end, const int32_t* __restrict__ cells,
                  const int32_t* __restrict__ cell2nnz, const double* __restrict__ coords,
                  int64_t nrows_owned, double c_stiff, double c_mass, double s_poly, P f,
                  cb_quad q, double s_nodal, const double* __restrict__ fnodal,
                  const uint8_t* __restrict__ bc_flag, const double* __restrict__ bc_val,
                  double* __restrict__ val, double* __restrict__ rhs, double* __restrict__ cell_out) {
  constexpr int K = D + 1;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  for (int64_t c = c_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < c_end;
       c += (int64_t)gridDim.x * blockDim.x) {
    CellData<D> cd;
    load_cell_data<D>(cells, coords, c, cd);
    const double vol = cd.g.vol;
    double Ae[K][K];
    double be[K];
    bool isbc[K];
    double gbc[K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
      be[a] = 0.0;
      isbc[a] = bc_flag ? (bc_flag[cd.v[a]] != 0) : false;
      gbc[a] = isbc[a] ? bc_val[cd.v[a]] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int b = 0; b < K; ++b) {
        double gg = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) gg += cd.g.G[a][i] * cd.g.G[b][i];
        Ae[a][b] = vol * (c_stiff * gg + c_mass * (a == b ? 2.0 * MREF : MREF));
      }
    }
    if (rhs) {
      if (s_poly != 0.0) {
        double acc[K];
#pragma unroll
        for (int a = 0; a < K; ++a) acc[a] = 0.0;
        for (int iq = 0; iq < q.nq; ++iq) {
          double xq[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int a = 0; a < K; ++a) {
#pragma unroll
            for (int i = 0; i < D; ++i) xq[i] += q.lam[iq][a] * cd.x[a][i];
          }
          const double fw = coef_eval<D, DEG, ZDEP>(f, xq[0], xq[1], xq[2]) * q.w[iq];
#pragma unroll
          for (int a = 0; a < K; ++a) acc[a] += fw * q.lam[iq][a];
        }
#pragma unroll
        for (int a = 0; a < K; ++a) be[a] += s_poly * vol * acc[a];
      }
      if (s_nodal != 0.0) {
        double fe[K], fsum = 0.0;
#pragma unroll
        for (int a = 0; a < K; ++a) { fe[a] = fnodal[cd.v[a]]; fsum += fe[a]; }
#pragma unroll
        for (int a = 0; a < K; ++a) be[a] += s_nodal * vol * MREF * (fsum + fe[a]);
      }
    }
    // symmetric Dirichlet elimination on the element tensor (dolfin SystemAssembler)
#pragma unroll
    for (int a = 0; a < K; ++a) {
      double lift = 0.0;
#pragma unroll
      for (int b = 0; b < K; ++b) lift += Ae[a][b] * gbc[b];
      be[a] = isbc[a] ? gbc[a] : be[a] - lift;
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int b = 0; b < K; ++b) {
        if (isbc[a] || isbc[b]) Ae[a][b] = (a == b) ? 1.0 : 0.0;
      }
    }
    if (val) {
      const int32_t* c2n = cell2nnz + c * (K * K);
#pragma unroll
      for (int a = 0; a < K; ++a) {
#pragma unroll
        for (int b = 0; b < K; ++b) {
          const int32_t pos = __ldg(c2n + a * K + b);
          if (pos >= 0) val[pos] += Ae[a][b];
        }
      }
    }
    if (cell_out) {  // gather path: local vector to scratch, summed per vertex by k_gather_cell_vectors
#pragma unroll
      for (int a = 0; a < K; ++a) cell_out[c * K + a] = be[a];
    } else if (rhs) {
#pragma unroll
      for (int a = 0; a < K; ++a) {
        if (cd.v[a] < nrows_owned) rhs[cd.v[a]] += be[a];
      }
    }
  }
}

// ------------------------------------------------------------------ vector P1 elasticity
// One thread per (cell, block (a, b)): the 16 (9 in 2-D) threads of a cell recompute the cheap cell
// geometry (loads are warp-broadcasts) and each owns one d x d block, i.e. one contiguous
// read-modify-write of d*d doubles through a coalesced read of the cell->nnz map.  Compared with
// one thread per cell this needs 4x fewer registers and 16x more parallelism.
template <int D>
__global__ void __launch_bounds__(256)
k_assemble_elasticity(int64_t c_begin, int64_t c_end, const int32_t* __restrict__ cells,
                      const int32_t* __restrict__ cell2nnz, const double* __restrict__ coords,
                      const double* __restrict__ mu, double lambda_lame, double damping,
                      const double* __restrict__ cell_scale, const uint8_t* __restrict__ bc_flag,
                      double* __restrict__ val) {
  constexpr int K = D + 1;
  constexpr int KK = K * K;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  const int64_t t_end = c_end * KK;
  for (int64_t t = c_begin * KK + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < t_end;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / KK;
    const int ab = (int)(t - c * KK);
    const int32_t pos = __ldg(cell2nnz + t);
    if (pos < 0) continue;
    const int a = ab / K, b = ab - a * K;
    CellData<D> cd;
    load_cell_data<D>(cells, coords, c, cd);
    double mubar = 0.0;
#pragma unroll
    for (int v = 0; v < K; ++v) mubar += mu[cd.v[v]];
    mubar *= (1.0 / K);
    const double scale = cell_scale ? cell_scale[c] : 1.0;
    const double vol = cd.g.vol;
    const double vm = scale * vol * mubar, vl = scale * vol * lambda_lame, dm = scale * damping * vol * MREF;
    double Ga[D], Gb[D];
    int va = 0, vb = 0;
#pragma unroll
    for (int v = 0; v < K; ++v) {
      if (v == a) {
        va = cd.v[v];
#pragma unroll
        for (int i = 0; i < D; ++i) Ga[i] = cd.g.G[v][i];
      }
      if (v == b) {
        vb = cd.v[v];
#pragma unroll
        for (int i = 0; i < D; ++i) Gb[i] = cd.g.G[v][i];
      }
    }
    unsigned bca = 0, bcb = 0;  // Dirichlet flags of the row dofs (a, i) and the column dofs (b, j)
    if (bc_flag) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if (bc_flag[(int64_t)va * D + i]) bca |= 1u << i;
        if (bc_flag[(int64_t)vb * D + i]) bcb |= 1u << i;
      }
    }
    double gg = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) gg += Ga[i] * Gb[i];
    const double mass = (a == b) ? 2.0 * dm : dm;
    double* blk = val + (int64_t)pos * (D * D);
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double e = vm * Ga[j] * Gb[i] + vl * Ga[i] * Gb[j];
        if (i == j) e += vm * gg + mass;
        if (((bca >> i) & 1u) || ((bcb >> j) & 1u)) e = (a == b && i == j) ? 1.0 : 0.0;
        blk[i * D + j] += e;
      }
    }
  }
}

// ------------------------------------------------------------------ gather (nnz-centric) matrix assembly
// One thread per block-nnz entry q = (row vertex, column vertex): it visits the cells sharing that
// edge (n2c list, ascending cell id, ~6 cells in a Kuhn mesh), recomputes the cheap cell geometry
// (cells of neighbouring entries coincide -> L1/L2 hits) and writes the finished d x d block once.
// No colours, no read-modify-write, one launch; fixed summation order -> bit-reproducible.
// Per-cell geometry cache for the gather kernels: one 128-byte line per cell,
//   geo[c] = { G[a][i] (a < D+1, i < D; 12 doubles in 3-D), w0, w1, w2, pad }
// scalar:     w0 = |K|
// elasticity: w0 = s |K| mubar, w1 = s |K| lambda, w2 = s delta |K| / ((D+1)(D+2))   (s = cell scale)
// Written by one streaming pass over the cells (k_cell_geometry); an entry thread then needs 2 G rows
// and the weights of each incident cell -- 7 to 9 loads from one line -- instead of re-deriving the
// geometry (12 coordinate gathers, a 3x3 inverse) of ~6 cells per entry.
constexpr int GEO_STRIDE = 16;

template <int D>
__global__ void __launch_bounds__(256)
k_cell_geometry(int64_t nc, const int32_t* __restrict__ cells, const double* __restrict__ coords,
                const double* __restrict__ mu, double lambda_lame, double damping,
                const double* __restrict__ cell_scale, double* __restrict__ geo) {
  constexpr int K = D + 1;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    CellData<D> cd;
    load_cell_data<D>(cells, coords, c, cd);
    double* g = geo + c * GEO_STRIDE;
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int i = 0; i < D; ++i) g[a * D + i] = cd.g.G[a][i];
    }
    const double vol = cd.g.vol;
    if (mu) {
      double mubar = 0.0;
#pragma unroll
      for (int v = 0; v < K; ++v) mubar += mu[cd.v[v]];
      mubar *= (1.0 / K);
      const double scale = cell_scale ? cell_scale[c] : 1.0;
      g[12] = scale * vol * mubar;
      g[13] = scale * vol * lambda_lame;
      g[14] = scale * damping * vol * MREF;
    } else {
      g[12] = vol;
    }
  }
}

// Diagonal entries visit every cell around their vertex (~24 in a Kuhn mesh) while off-diagonal ones visit the ~5
// cells around an edge, and every warp holds two or three diagonal entries: with one thread per entry the whole warp
// would idle through the diagonal lanes' long loops.  So a lane only walks the list of an OFF-diagonal entry itself;
// the diagonal entries of a warp are then processed one after the other by the whole warp -- lanes stride over the
// cell list, partial blocks are combined by a fixed xor-shuffle tree (bit-reproducible, like every other sum here).
template <int D>
__device__ __forceinline__ void elasticity_term(const double* __restrict__ geo, int32_t t, bool diag, unsigned bca,
                                                unsigned bcb, double (&acc)[D * D]) {
  constexpr int K = D + 1;
  constexpr int KK = K * K;
  const int64_t c = t / KK;
  const int ab = t - (int32_t)c * KK;
  const int a = ab / K, b = ab - a * K;
  const double* g = geo + c * GEO_STRIDE;
  double Ga[D], Gb[D];
#pragma unroll
  for (int i = 0; i < D; ++i) { Ga[i] = g[a * D + i]; Gb[i] = g[b * D + i]; }
  const double vm = g[12], vl = g[13], dm = g[14];
  double gg = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) gg += Ga[i] * Gb[i];
  const double mass = diag ? 2.0 * dm : dm;
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double e = vm * Ga[j] * Gb[i] + vl * Ga[i] * Gb[j];
      if (i == j) e += vm * gg + mass;
      if (((bca >> i) & 1u) || ((bcb >> j) & 1u)) e = (diag && i == j) ? 1.0 : 0.0;
      acc[i * D + j] += e;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256)
k_assemble_elasticity_gather(int64_t nnz, const int64_t* __restrict__ n2c_ptr, const int32_t* __restrict__ n2c_idx,
                             const int32_t* __restrict__ cells, const double* __restrict__ geo,
                             const uint8_t* __restrict__ bc_flag, double* __restrict__ val) {
  constexpr int K = D + 1;
  constexpr int KK = K * K;
  const int lane = threadIdx.x & 31;
  const int64_t nnz_up = (nnz + 31) & ~int64_t(31);  // warp-uniform trip count (the diagonal pass shuffles)
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz_up;
       q += (int64_t)gridDim.x * blockDim.x) {
    const bool valid = q < nnz;
    int64_t kb = 0, ke = 0;
    unsigned bca = 0, bcb = 0;
    bool diag = false;
    if (valid) {
      kb = n2c_ptr[q]; ke = n2c_ptr[q + 1];
      if (ke > kb) {
        const int32_t t = __ldg(n2c_idx + kb);
        const int64_t c = t / KK;
        const int ab = t - (int32_t)c * KK;
        const int a = ab / K, b = ab - a * K;
        diag = (a == b);
        if (bc_flag) {  // Dirichlet flags of the row vertex / column vertex (the same for every cell)
          const int64_t va = cells[c * K + a], vb = cells[c * K + b];
#pragma unroll
          for (int i = 0; i < D; ++i) {
            if (bc_flag[va * D + i]) bca |= 1u << i;
            if (bc_flag[vb * D + i]) bcb |= 1u << i;
          }
        }
      }
    }
    if (valid && !diag) {
      double acc[D * D];
#pragma unroll
      for (int i = 0; i < D * D; ++i) acc[i] = 0.0;
      for (int64_t k = kb; k < ke; ++k) elasticity_term<D>(geo, __ldg(n2c_idx + k), false, bca, bcb, acc);
      double* blk = val + q * (D * D);
#pragma unroll
      for (int i = 0; i < D * D; ++i) blk[i] = acc[i];
    }
    unsigned dmask = __ballot_sync(0xffffffffu, diag);
    while (dmask) {
      const int src = __ffs(dmask) - 1;
      dmask &= dmask - 1;
      const int64_t skb = __shfl_sync(0xffffffffu, kb, src), ske = __shfl_sync(0xffffffffu, ke, src);
      const unsigned sbc = __shfl_sync(0xffffffffu, bca, src);  // row vertex == column vertex
      double acc[D * D];
#pragma unroll
      for (int i = 0; i < D * D; ++i) acc[i] = 0.0;
      for (int64_t k = skb + lane; k < ske; k += 32) elasticity_term<D>(geo, __ldg(n2c_idx + k), true, sbc, sbc, acc);
#pragma unroll
      for (int i = 0; i < D * D; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
      }
      if (lane == src) {
        double* blk = val + q * (D * D);
#pragma unroll
        for (int i = 0; i < D * D; ++i) blk[i] = acc[i];
      }
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256)
k_assemble_scalar_gather(int64_t nnz, const int64_t* __restrict__ n2c_ptr, const int32_t* __restrict__ n2c_idx,
                         const int32_t* __restrict__ cells, const double* __restrict__ geo, double c_stiff,
                         double c_mass, const uint8_t* __restrict__ bc_flag, double* __restrict__ val) {
  constexpr int K = D + 1;
  constexpr int KK = K * K;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  const int lane = threadIdx.x & 31;
  const int64_t nnz_up = (nnz + 31) & ~int64_t(31);
  auto term = [&](int32_t t, bool diag, bool isbc) -> double {
    const int64_t c = t / KK;
    const int ab = t - (int32_t)c * KK;
    const int a = ab / K, b = ab - a * K;
    const double* g = geo + c * GEO_STRIDE;
    double gg = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) gg += g[a * D + i] * g[b * D + i];
    const double e = g[12] * (c_stiff * gg + c_mass * (diag ? 2.0 * MREF : MREF));
    return isbc ? (diag ? 1.0 : 0.0) : e;
  };
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nnz_up;
       q += (int64_t)gridDim.x * blockDim.x) {
    const bool valid = q < nnz;
    int64_t kb = 0, ke = 0;
    bool isbc = false, diag = false;
    if (valid) {
      kb = n2c_ptr[q]; ke = n2c_ptr[q + 1];
      if (ke > kb) {
        const int32_t t = __ldg(n2c_idx + kb);
        const int64_t c = t / KK;
        const int ab = t - (int32_t)c * KK;
        const int a = ab / K, b = ab - a * K;
        diag = (a == b);
        if (bc_flag) isbc = bc_flag[cells[c * K + a]] || bc_flag[cells[c * K + b]];
      }
    }
    if (valid && !diag) {
      double acc = 0.0;
      for (int64_t k = kb; k < ke; ++k) acc += term(__ldg(n2c_idx + k), false, isbc);
      val[q] = acc;
    }
    unsigned dmask = __ballot_sync(0xffffffffu, diag);
    while (dmask) {  // diagonal entries: the whole warp shares the long cell list (see the elasticity kernel)
      const int src = __ffs(dmask) - 1;
      dmask &= dmask - 1;
      const int64_t skb = __shfl_sync(0xffffffffu, kb, src), ske = __shfl_sync(0xffffffffu, ke, src);
      const bool sbc = __shfl_sync(0xffffffffu, (int)isbc, src) != 0;
      double acc = 0.0;
      for (int64_t k = skb + lane; k < ske; k += 32) acc += term(__ldg(n2c_idx + k), true, sbc);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == src) val[q] = acc;
    }
  }
}

// ------------------------------------------------------------------ row-centric matrix assembly
// The entry-centric gather kernels above are bound by L2 traffic: every one of the 16 (a, b) pairs of a cell pulls
// the cell's 128-byte geometry line again (161 M line requests at n = 119 for 10 M cells).  Here a group of GS lanes
// (16 for mesh stencils of <= 16 entries per row, else 32) owns one ROW i:
//   1. the lanes compute the geometry of the cells around vertex i (v2c list, ascending cell id) straight from the
//      coordinates -- no geometry cache pass, no cache buffer -- into shared memory: each (vertex, cell) pair once;
//   2. lane e owns entry (i, col[e]) of the row and walks the staged cells in order; a cell that contains the column
//      vertex contributes its term (fixed summation order: ascending cell id, bit-reproducible and the order of the
//      n2c lists);
//   3. Dirichlet rows / columns are eliminated at the end: an eliminated diagonal entry is the number of cells at
//      the vertex (what dolfin's SystemAssembler leaves there, summing one per cell), other eliminated entries 0.
// Rows with more cells than a chunk holds (CH = 32) or more entries than the group has lanes are processed in chunks.
template <int D>
struct RowCellStage {
  static constexpr int K = D + 1;
  static constexpr int CH = 32;
  int cv[CH][K];
  double G[CH][K * D];
  double w[CH][3];
  int loc[CH];
};

template <int D, bool ELAST, int GS>
__global__ void __launch_bounds__(8 * GS)
k_assemble_rows(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                const int32_t* __restrict__ cells, const double* __restrict__ coords,
                const double* __restrict__ mu, double p0 /* lambda | c_stiff */, double p1 /* damping | c_mass */,
                const double* __restrict__ cell_scale, const uint8_t* __restrict__ bc_flag, double* __restrict__ val) {
  constexpr int K = D + 1;
  constexpr int NB = ELAST ? D * D : 1;
  constexpr int CH = RowCellStage<D>::CH;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  __shared__ RowCellStage<D> stage_all[8];
  const int tile = threadIdx.x / GS, gl = threadIdx.x - tile * GS;
  RowCellStage<D>& st = stage_all[tile];
  const unsigned tmask = (GS == 32) ? 0xffffffffu : (0xffffu << (16 * ((threadIdx.x & 31) / 16)));
  const int64_t tiles_total = (int64_t)gridDim.x * 8;
  for (int64_t i = blockIdx.x * (int64_t)8 + tile; i < nrows; i += tiles_total) {
    const int64_t cb0 = v2c_ptr[i], cb1 = v2c_ptr[i + 1];
    const int64_t eb0 = rowptr[i], eb1 = rowptr[i + 1];
    for (int64_t e0 = eb0; e0 < eb1; e0 += GS) {  // entry chunks (one for mesh stencils)
      const bool has_entry = e0 + gl < eb1;
      const int j = has_entry ? __ldg(col + e0 + gl) : -1;
      const bool diag = (j == (int)i);
      double acc[NB];
#pragma unroll
      for (int q = 0; q < NB; ++q) acc[q] = 0.0;
      int count = 0;
      for (int64_t c0 = cb0; c0 < cb1; c0 += CH) {  // cell chunks (one for mesh vertices of valence <= 32)
        const int nch = (int)((cb1 - c0 < CH) ? (cb1 - c0) : CH);
        __syncwarp(tmask);  // the previous chunk is consumed
        for (int sidx = gl; sidx < nch; sidx += GS) {
          const int64_t c = __ldg(v2c_idx + c0 + sidx);
          CellData<D> cd;
          load_cell_data<D>(cells, coords, c, cd);
          int a = 0;
#pragma unroll
          for (int v = 0; v < K; ++v) {
            st.cv[sidx][v] = cd.v[v];
            a = (cd.v[v] == (int)i) ? v : a;
#pragma unroll
            for (int d = 0; d < D; ++d) st.G[sidx][v * D + d] = cd.g.G[v][d];
          }
          st.loc[sidx] = a;
          const double vol = cd.g.vol;
          if (ELAST) {
            double mubar = 0.0;
#pragma unroll
            for (int v = 0; v < K; ++v) mubar += mu[cd.v[v]];
            mubar *= (1.0 / K);
            const double scale = cell_scale ? cell_scale[c] : 1.0;
            st.w[sidx][0] = scale * vol * mubar;
            st.w[sidx][1] = scale * vol * p0;
            st.w[sidx][2] = scale * p1 * vol * MREF;
          } else {
            st.w[sidx][0] = vol;
          }
        }
        __syncwarp(tmask);
        if (has_entry) {
          for (int sidx = 0; sidx < nch; ++sidx) {
            int b = -1;
#pragma unroll
            for (int v = 0; v < K; ++v) b = (st.cv[sidx][v] == j) ? v : b;
            if (b < 0) continue;
            const int a = st.loc[sidx];
            ++count;
            double gg = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) gg += st.G[sidx][a * D + d] * st.G[sidx][b * D + d];
            if (ELAST) {
              const double vm = st.w[sidx][0], vl = st.w[sidx][1], dm = st.w[sidx][2];
              const double mass = diag ? 2.0 * dm : dm;
#pragma unroll
              for (int r = 0; r < D; ++r) {
#pragma unroll
                for (int q = 0; q < D; ++q) {
                  double e = vm * st.G[sidx][a * D + q] * st.G[sidx][b * D + r] +
                             vl * st.G[sidx][a * D + r] * st.G[sidx][b * D + q];
                  if (r == q) e += vm * gg + mass;
                  acc[ELAST ? r * D + q : 0] += e;
                }
              }
            } else {
              acc[0] += st.w[sidx][0] * (p0 * gg + p1 * (diag ? 2.0 * MREF : MREF));
            }
          }
        }
      }
      if (has_entry) {
        if (ELAST) {
          unsigned bca = 0, bcb = 0;
          if (bc_flag) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              if (bc_flag[i * D + d]) bca |= 1u << d;
              if (bc_flag[(int64_t)j * D + d]) bcb |= 1u << d;
            }
          }
          double* blk = val + (e0 + gl) * NB;
#pragma unroll
          for (int r = 0; r < D; ++r) {
#pragma unroll
            for (int q = 0; q < D; ++q) {
              double e = acc[ELAST ? r * D + q : 0];
              if (((bca >> r) & 1u) || ((bcb >> q) & 1u)) e = (diag && r == q) ? (double)count : 0.0;
              blk[r * D + q] = e;
            }
          }
        } else {
          double e = acc[0];
          if (bc_flag && (bc_flag[i] || bc_flag[j])) e = diag ? (double)count : 0.0;
          val[e0 + gl] = e;
        }
      }
    }
  }
}

template <int D, bool ELAST>
static int launch_assemble_rows(cudaStream_t s, const cb_mesh_view* m, const double* mu, double p0, double p1,
                                const double* cell_scale, const uint8_t* bc_flag, double* val) {
  // mesh stencils (<= 16 entries per row on average with some slack): two rows per warp
  const bool small = m->nnz <= 16 * m->nrows_owned;
  // One wave exactly: the grid is what is resident (asked from the occupancy calculator -- registers and the 36 KB of
  // staging per CTA allow 5 CTAs per SM, not the 6 the shared memory alone would; with 6 x 148 CTAs the sixth CTA of
  // every SM ran as a second wave of its own, 1.2 waves for the price of 2: ncu launch__waves_per_multiprocessor).
  auto one_wave = [&](auto kernel, int threads) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) {
      (void)cudaGetLastError();
      per_sm = 4;
    }
    return grid_for(m->nrows_owned, 8, per_sm);  // 8 rows per CTA and sweep
  };
  if (small) {
    auto kernel = k_assemble_rows<D, ELAST, 16>;
    kernel<<<one_wave(kernel, 128), 128, 0, s>>>(m->nrows_owned, m->rowptr, m->col, m->v2c_ptr, m->v2c_idx,
                                                  m->cells_orig, m->coords, mu, p0, p1, cell_scale, bc_flag, val);
  } else {
    auto kernel = k_assemble_rows<D, ELAST, 32>;
    kernel<<<one_wave(kernel, 256), 256, 0, s>>>(m->nrows_owned, m->rowptr, m->col, m->v2c_ptr, m->v2c_idx,
                                                  m->cells_orig, m->coords, mu, p0, p1, cell_scale, bc_flag, val);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// ------------------------------------------------------------------ shape derivative of the Poisson-type Lagrangian
// L = int c_s u + c_t/2 (u - u_d)^2 + kappa grad u . grad p + c u p - f p dx with P1 u, p, u_d and polynomial f:
//   dL[V] = int (density) div V - kappa (grad V + grad V^T) grad u . grad p - (grad f . V) p - c_t (u - u_d) (grad u_d . V)
// (the last term is the pull-back of cashocs/_forms/shape_form_handler.py:507-531 for a Function u_d; analytic forms:
// tests/test_shape_optimization.py:252-259 and :284-291).  Everything but the two f-integrals is in closed form.
struct ShapeDerivCoef {
  double c_int_u, kappa, reaction, c_track;
  int pull_back;
};

template <int D, class P, int DEG, bool ZDEP>
__global__ void __launch_bounds__(128)
k_shape_derivative_poisson(int64_t c_begin, int64_t c_end, const int32_t* __restrict__ cells,
                           const double* __restrict__ coords, int64_t nrows_owned,
                           const double* __restrict__ u, const double* __restrict__ p, P f,
                           P gf0, P gf1, P gf2, cb_quad q, ShapeDerivCoef co,
                           const double* __restrict__ ud, const uint8_t* __restrict__ bc_flag,
                           double* __restrict__ rhs, double* __restrict__ cell_out) {
  constexpr int K = D + 1;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  for (int64_t c = c_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < c_end;
       c += (int64_t)gridDim.x * blockDim.x) {
    CellData<D> cd;
    load_cell_data<D>(cells, coords, c, cd);
    const double vol = cd.g.vol;
    double ue[K], pe[K], usum = 0.0, psum = 0.0, up = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      ue[a] = u[cd.v[a]]; pe[a] = p[cd.v[a]];
      usum += ue[a]; psum += pe[a]; up += ue[a] * pe[a];
    }
    double gu[D], gp[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      gu[i] = 0.0; gp[i] = 0.0;
#pragma unroll
      for (int a = 0; a < K; ++a) { gu[i] += ue[a] * cd.g.G[a][i]; gp[i] += pe[a] * cd.g.G[a][i]; }
    }
    double gugp = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) gugp += gu[i] * gp[i];
    // quadrature: int f p  and  int (d_i f) phi_a p
    double int_fp = 0.0;
    double igf[K][D];
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int i = 0; i < D; ++i) igf[a][i] = 0.0;
    }
    if (coef_nonzero(f)) {
      for (int iq = 0; iq < q.nq; ++iq) {
        double xq[3] = {0.0, 0.0, 0.0};
        double pq = 0.0;
#pragma unroll
        for (int a = 0; a < K; ++a) {
          pq += q.lam[iq][a] * pe[a];
#pragma unroll
          for (int i = 0; i < D; ++i) xq[i] += q.lam[iq][a] * cd.x[a][i];
        }
        const double wp = q.w[iq] * pq;
        constexpr int GDEG = DEG > 0 ? DEG - 1 : 0;  // the gradient tables hold one degree less
        int_fp += wp * coef_eval<D, DEG, ZDEP>(f, xq[0], xq[1], xq[2]);
        double gfq[3];
        gfq[0] = coef_eval<D, GDEG, ZDEP>(gf0, xq[0], xq[1], xq[2]);
        gfq[1] = coef_eval<D, GDEG, ZDEP>(gf1, xq[0], xq[1], xq[2]);
        gfq[2] = (D == 3 && ZDEP) ? coef_eval<D, GDEG, ZDEP>(gf2, xq[0], xq[1], xq[2]) : 0.0;
#pragma unroll
        for (int a = 0; a < K; ++a) {
#pragma unroll
          for (int i = 0; i < D; ++i) igf[a][i] += wp * q.lam[iq][a] * gfq[i];
        }
      }
    }
    // Lagrangian density integrated over the cell (the factor of div V = G_a[i])
    double s0 = co.c_int_u * vol * usum * (1.0 / K) + co.kappa * vol * gugp - vol * int_fp;
    if (co.reaction != 0.0) s0 += co.reaction * vol * MREF * (usum * psum + up);
    double gud[D], mw[K];
#pragma unroll
    for (int i = 0; i < D; ++i) gud[i] = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) mw[a] = 0.0;
    if (co.c_track != 0.0) {
      double we[K], wsum = 0.0, ww = 0.0;
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const double uda = ud[cd.v[a]];
        we[a] = ue[a] - uda;
        wsum += we[a]; ww += we[a] * we[a];
#pragma unroll
        for (int i = 0; i < D; ++i) gud[i] += uda * cd.g.G[a][i];
      }
      s0 += 0.5 * co.c_track * vol * MREF * (wsum * wsum + ww);
      if (co.pull_back) {
#pragma unroll
        for (int a = 0; a < K; ++a) mw[a] = co.c_track * vol * MREF * (wsum + we[a]);  // c_t int (u - u_d) phi_a
      }
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
      if (!cell_out && cd.v[a] >= nrows_owned) continue;
      double Ggu = 0.0, Ggp = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) { Ggu += cd.g.G[a][i] * gu[i]; Ggp += cd.g.G[a][i] * gp[i]; }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double e = cd.g.G[a][i] * s0 - co.kappa * vol * (Ggu * gp[i] + Ggp * gu[i]) - vol * igf[a][i] - mw[a] * gud[i];
        const int64_t dof = (int64_t)cd.v[a] * D + i;
        if (bc_flag && bc_flag[dof]) e = 0.0;
        if (cell_out) cell_out[(c * K + a) * D + i] = e;
        else rhs[dof] += e;
      }
    }
  }
}

// ------------------------------------------------------------------ semilinear states: c u^3 p
// Residual and Jacobian of the cubic term of F(u)[v] = int kappa grad u . grad v + c0 u v + c u^3 v - f v dx
// (the nonlinearity of the reference's tests: tests/test_nonlinear_solvers.py:28-61,
// tests/test_optimal_control.py:364-376), which cashocs.newton_solve (nonlinear_solvers/newton_solver.py:285-362)
// re-assembles in every Newton step:
//   load_a  = c   |K| sum_q w_q u(x_q)^3 phi_a(x_q)                (k_cubic_load: one thread per cell -> scratch -> gather)
//   M_ab   += 3 c |K| sum_q w_q u(x_q)^2 phi_a(x_q) phi_b(x_q)     (k_add_weighted_mass_gather: one thread per nnz entry)
// with a rule exact for degree 4 (P1 u).  Entries of Dirichlet rows / columns are left alone: the symmetric
// elimination of the linear part already made them identity rows, and the increment has homogeneous conditions.
template <int D>
__global__ void __launch_bounds__(256)
k_cubic_load(int64_t nc, const int32_t* __restrict__ cells, const double* __restrict__ coords,
             const double* __restrict__ u, double c, cb_quad q, double* __restrict__ cell_out) {
  constexpr int K = D + 1;
  for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < nc;
       cell += (int64_t)gridDim.x * blockDim.x) {
    CellData<D> cd;
    load_cell_data<D>(cells, coords, cell, cd);
    double ue[K], acc[K];
#pragma unroll
    for (int a = 0; a < K; ++a) { ue[a] = u[cd.v[a]]; acc[a] = 0.0; }
    for (int iq = 0; iq < q.nq; ++iq) {
      double uq = 0.0;
#pragma unroll
      for (int a = 0; a < K; ++a) uq += q.lam[iq][a] * ue[a];
      const double w3 = q.w[iq] * uq * uq * uq;
#pragma unroll
      for (int a = 0; a < K; ++a) acc[a] += w3 * q.lam[iq][a];
    }
#pragma unroll
    for (int a = 0; a < K; ++a) cell_out[cell * K + a] = c * cd.g.vol * acc[a];
  }
}

template <int D>
__global__ void __launch_bounds__(256)
k_add_weighted_mass_gather(int64_t nnz, const int64_t* __restrict__ n2c_ptr, const int32_t* __restrict__ n2c_idx,
                           const int32_t* __restrict__ cells, const double* __restrict__ coords,
                           const double* __restrict__ u, double c3, cb_quad q, const uint8_t* __restrict__ bc_flag,
                           double* __restrict__ val) {
  constexpr int K = D + 1;
  constexpr int KK = K * K;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    const int64_t kb = n2c_ptr[e], ke = n2c_ptr[e + 1];
    bool isbc = false;
    for (int64_t k = kb; k < ke; ++k) {
      const int32_t t = __ldg(n2c_idx + k);
      const int64_t cell = t / KK;
      const int ab = t - (int32_t)cell * KK;
      const int a = ab / K, b = ab - a * K;
      CellData<D> cd;
      load_cell_data<D>(cells, coords, cell, cd);
      if (k == kb && bc_flag) isbc = bc_flag[cd.v[a]] || bc_flag[cd.v[b]];
      double ue[K];
#pragma unroll
      for (int v = 0; v < K; ++v) ue[v] = u[cd.v[v]];
      double s = 0.0;
      for (int iq = 0; iq < q.nq; ++iq) {
        double uq = 0.0, la = 0.0, lb = 0.0;
#pragma unroll
        for (int v = 0; v < K; ++v) {
          const double l = q.lam[iq][v];
          uq += l * ue[v];
          la = (v == a) ? l : la;
          lb = (v == b) ? l : lb;
        }
        s += q.w[iq] * uq * uq * la * lb;
      }
      acc += c3 * cd.g.vol * s;
    }
    if (!isbc) val[e] += acc;
  }
}

// ------------------------------------------------------------------ ident_zeros
__global__ void k_ident_zeros(int64_t nrows, int bs, const int64_t* __restrict__ rowptr,
                              const int32_t* __restrict__ col, double* __restrict__ val,
                              unsigned long long* __restrict__ n_fixed) {
  const int64_t total = nrows * bs;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t br = r / bs;
    const int i = (int)(r - br * bs);
    double s = 0.0;
    int64_t dpos = -1;
    for (int64_t k = rowptr[br]; k < rowptr[br + 1]; ++k) {
      for (int j = 0; j < bs; ++j) s += fabs(val[(k * bs + i) * bs + j]);
      if (col[k] == br) dpos = k;
    }
    if (s < 3.0e-16 && dpos >= 0) {
      val[(dpos * bs + i) * bs + i] = 1.0;
      if (n_fixed) atomicAdd(n_fixed, 1ULL);
    }
  }
}

template <typename F>
int for_each_colour(const cb_mesh_view* m, F&& launch) {
  for (int c = 0; c < m->ncolours; ++c) {
    const int64_t b = m->colour_off[c], e = m->colour_off[c + 1];
    if (e <= b) continue;
    const int grid = grid_for(e - b, 128, 12);
    launch(b, e, grid);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

}  // namespace cb

using namespace cb;

static int check_view(const cb_mesh_view* m) {
  CB_REQUIRE(m != nullptr, "mesh view is NULL");
  CB_REQUIRE(m->dim == 2 || m->dim == 3, "dim must be 2 or 3");
  CB_REQUIRE(m->ncolours > 0 && m->colour_off != nullptr, "colour offsets missing");
  CB_REQUIRE(m->cells && m->coords, "cells / coords missing");
  return CB_OK;
}

template <class P, int DEG, bool ZDEP>
static int launch_assemble_scalar(cudaStream_t s, const cb_mesh_view* m, double c_stiff, double c_mass, double s_poly,
                                  const P& ff, const cb_quad& qq, double s_nodal, const double* fnodal,
                                  const uint8_t* bc_flag, const double* bc_val, double* val, double* rhs) {
  if (!val && rhs && m->cells_orig && m->v2c_ptr && m->v2c_idx && m->cell_scratch) {
    // vector by gather: one launch over all cells (local vectors to scratch) + one per-vertex sum
    const int64_t nc = m->colour_off[m->ncolours];
    const int grid = grid_for(nc, 128, 12), vgrid = grid_for(m->nrows_owned, 256, 8);
    if (m->dim == 2) {
      k_assemble_scalar<2, P, DEG, ZDEP><<<grid, 128, 0, s>>>(0, nc, m->cells_orig, nullptr, m->coords, m->nrows_owned, c_stiff,
                                                  c_mass, s_poly, ff, qq, s_nodal, fnodal, bc_flag, bc_val, nullptr,
                                                  rhs, m->cell_scratch);
      CB_LAUNCH_CHECK();
      k_gather_cell_vectors<2, 1><<<vgrid, 256, 0, s>>>(m->nrows_owned, m->v2c_ptr, m->v2c_idx, m->cells_orig,
                                                       m->cell_scratch, rhs);
    } else {
      k_assemble_scalar<3, P, DEG, ZDEP><<<grid, 128, 0, s>>>(0, nc, m->cells_orig, nullptr, m->coords, m->nrows_owned, c_stiff,
                                                  c_mass, s_poly, ff, qq, s_nodal, fnodal, bc_flag, bc_val, nullptr,
                                                  rhs, m->cell_scratch);
      CB_LAUNCH_CHECK();
      k_gather_cell_vectors<3, 1><<<vgrid, 256, 0, s>>>(m->nrows_owned, m->v2c_ptr, m->v2c_idx, m->cells_orig,
                                                       m->cell_scratch, rhs);
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  if (val) CB_CUDA_CHECK(cudaMemsetAsync(val, 0, sizeof(double) * m->nnz, s));
  if (rhs) CB_CUDA_CHECK(cudaMemsetAsync(rhs, 0, sizeof(double) * m->nrows_owned, s));
  return for_each_colour(m, [&](int64_t b, int64_t e, int grid) {
    if (m->dim == 2)
      k_assemble_scalar<2, P, DEG, ZDEP><<<grid, 128, 0, s>>>(b, e, m->cells, m->cell2nnz, m->coords, m->nrows_owned,
                                                  c_stiff, c_mass, s_poly, ff, qq, s_nodal, fnodal,
                                                  bc_flag, bc_val, val, rhs, nullptr);
    else
      k_assemble_scalar<3, P, DEG, ZDEP><<<grid, 128, 0, s>>>(b, e, m->cells, m->cell2nnz, m->coords, m->nrows_owned,
                                                  c_stiff, c_mass, s_poly, ff, qq, s_nodal, fnodal,
                                                  bc_flag, bc_val, val, rhs, nullptr);
  });
}

template <class P, int DEG, bool ZDEP>
static int launch_shape_derivative(cudaStream_t s, const cb_mesh_view* m, const double* u, const double* p, const P& f,
                                   const P& g0, const P& g1, const P& g2, const cb_quad& q, const ShapeDerivCoef& co,
                                   const double* u_d, const uint8_t* bc_flag, double* rhs) {
  if (m->cells_orig && m->v2c_ptr && m->v2c_idx && m->cell_scratch) {
    const int64_t nc = m->colour_off[m->ncolours];
    const int grid = grid_for(nc, 128, 12), vgrid = grid_for(m->nrows_owned, 256, 8);
    if (m->dim == 2) {
      k_shape_derivative_poisson<2, P, DEG, ZDEP><<<grid, 128, 0, s>>>(0, nc, m->cells_orig, m->coords, m->nrows_owned, u, p, f,
                                                           g0, g1, g2, q, co, u_d, bc_flag, rhs, m->cell_scratch);
      CB_LAUNCH_CHECK();
      k_gather_cell_vectors<2, 2><<<vgrid, 256, 0, s>>>(m->nrows_owned, m->v2c_ptr, m->v2c_idx, m->cells_orig,
                                                       m->cell_scratch, rhs);
    } else {
      k_shape_derivative_poisson<3, P, DEG, ZDEP><<<grid, 128, 0, s>>>(0, nc, m->cells_orig, m->coords, m->nrows_owned, u, p, f,
                                                           g0, g1, g2, q, co, u_d, bc_flag, rhs, m->cell_scratch);
      CB_LAUNCH_CHECK();
      k_gather_cell_vectors<3, 3><<<vgrid, 256, 0, s>>>(m->nrows_owned, m->v2c_ptr, m->v2c_idx, m->cells_orig,
                                                       m->cell_scratch, rhs);
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  CB_CUDA_CHECK(cudaMemsetAsync(rhs, 0, sizeof(double) * m->nrows_owned * m->dim, s));
  return for_each_colour(m, [&](int64_t b, int64_t e, int grid) {
    if (m->dim == 2)
      k_shape_derivative_poisson<2, P, DEG, ZDEP><<<grid, 128, 0, s>>>(b, e, m->cells, m->coords, m->nrows_owned, u, p, f, g0, g1,
                                                           g2, q, co, u_d, bc_flag, rhs, nullptr);
    else
      k_shape_derivative_poisson<3, P, DEG, ZDEP><<<grid, 128, 0, s>>>(b, e, m->cells, m->coords, m->nrows_owned, u, p, f, g0, g1,
                                                           g2, q, co, u_d, bc_flag, rhs, nullptr);
  });
}

extern "C" int cb_assemble_scalar(void* stream, const cb_mesh_view* m, double c_stiff,
                                  double c_mass, double s_poly, const cb_poly* f, const cb_quad* q,
                                  double s_nodal, const double* fnodal, const uint8_t* bc_flag,
                                  const double* bc_val, double* val, double* rhs) {
  if (int e = check_view(m)) return e;
  CB_REQUIRE(val == nullptr || m->cell2nnz != nullptr, "cell2nnz missing");
  CB_REQUIRE(s_poly == 0.0 || (f && q), "polynomial source needs f and a quadrature rule");
  CB_REQUIRE(s_nodal == 0.0 || fnodal, "nodal source missing");
  CB_REQUIRE(bc_flag == nullptr || bc_val != nullptr, "bc_val missing");
  cudaStream_t s = as_stream(stream);
  if (val && m->rowptr && m->col && m->v2c_ptr && m->v2c_idx && m->cells_orig) {
    // matrix row by row: the cells around a vertex staged once in shared memory (k_assemble_rows)
    const int e = (m->dim == 2) ? launch_assemble_rows<2, false>(s, m, nullptr, c_stiff, c_mass, nullptr, bc_flag, val)
                                : launch_assemble_rows<3, false>(s, m, nullptr, c_stiff, c_mass, nullptr, bc_flag, val);
    if (e) return e;
    val = nullptr;
    if (!rhs) return CB_OK;
  }
  if (val && m->n2c_ptr && m->n2c_idx && m->cells_orig && m->cell_geo) {
    // matrix by gather: geometry cache pass + one thread per nnz entry
    const int64_t nc = m->colour_off[m->ncolours];
    const int cgrid = grid_for(nc, 256, 8), grid = grid_for(m->nnz, 256, 8);
    if (m->dim == 2) {
      k_cell_geometry<2><<<cgrid, 256, 0, s>>>(nc, m->cells_orig, m->coords, nullptr, 0.0, 0.0, nullptr, m->cell_geo);
      CB_LAUNCH_CHECK();
      k_assemble_scalar_gather<2><<<grid, 256, 0, s>>>(m->nnz, m->n2c_ptr, m->n2c_idx, m->cells_orig, m->cell_geo,
                                                      c_stiff, c_mass, bc_flag, val);
    } else {
      k_cell_geometry<3><<<cgrid, 256, 0, s>>>(nc, m->cells_orig, m->coords, nullptr, 0.0, 0.0, nullptr, m->cell_geo);
      CB_LAUNCH_CHECK();
      k_assemble_scalar_gather<3><<<grid, 256, 0, s>>>(m->nnz, m->n2c_ptr, m->n2c_idx, m->cells_orig, m->cell_geo,
                                                      c_stiff, c_mass, bc_flag, val);
    }
    CB_LAUNCH_CHECK();
    val = nullptr;
    if (!rhs) return CB_OK;
  }
  cb_poly fz;
  memset(&fz, 0, sizeof(fz));
  cb_quad qz;
  memset(&qz, 0, sizeof(qz));
  const cb_poly& ff = f ? *f : fz;
  const cb_quad& qq = q ? *q : qz;
  // the quadrature loop evaluates the source through the dense degree-4 table when it fits (common.cuh)
  DensePoly fd;
  if (dense_from_poly(ff, fd)) {
    bool zdep = false;
    const int deg = dense_degree(fd, &zdep);
#define CB_SCALAR_CASE(DEG, ZDEP) \
  return launch_assemble_scalar<DensePoly, DEG, ZDEP>(s, m, c_stiff, c_mass, s_poly, fd, qq, s_nodal, fnodal, bc_flag, bc_val, val, rhs)
    if (zdep && m->dim == 3) { if (deg <= 2) CB_SCALAR_CASE(2, true); CB_SCALAR_CASE(4, true); }
    if (deg <= 2) CB_SCALAR_CASE(2, false);
    CB_SCALAR_CASE(4, false);
#undef CB_SCALAR_CASE
  }
  return launch_assemble_scalar<cb_poly, 0, false>(s, m, c_stiff, c_mass, s_poly, ff, qq, s_nodal, fnodal, bc_flag, bc_val, val, rhs);
}

extern "C" int cb_assemble_elasticity(void* stream, const cb_mesh_view* m, const double* mu,
                                      double lambda_lame, double damping, const double* cell_scale,
                                      const uint8_t* bc_flag, double* val) {
  if (int e = check_view(m)) return e;
  CB_REQUIRE((m->cell2nnz || m->rowptr) && mu && val, "cell2nnz / mu / val missing");
  cudaStream_t s = as_stream(stream);
  if (m->rowptr && m->col && m->v2c_ptr && m->v2c_idx && m->cells_orig)
    return (m->dim == 2) ? launch_assemble_rows<2, true>(s, m, mu, lambda_lame, damping, cell_scale, bc_flag, val)
                         : launch_assemble_rows<3, true>(s, m, mu, lambda_lame, damping, cell_scale, bc_flag, val);
  if (m->n2c_ptr && m->n2c_idx && m->cells_orig && m->cell_geo) {
    const int64_t nc = m->colour_off[m->ncolours];
    const int cgrid = grid_for(nc, 256, 8), grid = grid_for(m->nnz, 256, 8);
    if (m->dim == 2) {
      k_cell_geometry<2><<<cgrid, 256, 0, s>>>(nc, m->cells_orig, m->coords, mu, lambda_lame, damping, cell_scale, m->cell_geo);
      CB_LAUNCH_CHECK();
      k_assemble_elasticity_gather<2><<<grid, 256, 0, s>>>(m->nnz, m->n2c_ptr, m->n2c_idx, m->cells_orig, m->cell_geo, bc_flag, val);
    } else {
      k_cell_geometry<3><<<cgrid, 256, 0, s>>>(nc, m->cells_orig, m->coords, mu, lambda_lame, damping, cell_scale, m->cell_geo);
      CB_LAUNCH_CHECK();
      k_assemble_elasticity_gather<3><<<grid, 256, 0, s>>>(m->nnz, m->n2c_ptr, m->n2c_idx, m->cells_orig, m->cell_geo, bc_flag, val);
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  CB_CUDA_CHECK(cudaMemsetAsync(val, 0, sizeof(double) * m->nnz * m->dim * m->dim, s));
  return for_each_colour(m, [&](int64_t b, int64_t e, int grid) {
    if (m->dim == 2)
      k_assemble_elasticity<2><<<grid_for((e - b) * 9, 256, 8), 256, 0, s>>>(b, e, m->cells, m->cell2nnz, m->coords, mu,
                                                                            lambda_lame, damping, cell_scale, bc_flag, val);
    else
      k_assemble_elasticity<3><<<grid_for((e - b) * 16, 256, 8), 256, 0, s>>>(b, e, m->cells, m->cell2nnz, m->coords, mu,
                                                                             lambda_lame, damping, cell_scale, bc_flag, val);
  });
}

extern "C" int cb_assemble_shape_derivative(void* stream, const cb_mesh_view* m, const double* u, const double* p,
                                            const cb_poly* f, const cb_poly* gradf, const cb_quad* q, double c_int_u,
                                            double kappa, double reaction, double c_track, const double* u_d,
                                            int pull_back, const uint8_t* bc_flag, double* rhs) {
  if (int e = check_view(m)) return e;
  CB_REQUIRE(u && p && f && gradf && q && rhs, "missing argument");
  CB_REQUIRE(c_track == 0.0 || u_d, "cb_assemble_shape_derivative: the tracking term needs u_d");
  cudaStream_t s = as_stream(stream);
  cb_poly gz;
  memset(&gz, 0, sizeof(gz));
  const ShapeDerivCoef co{c_int_u, kappa, reaction, c_track, pull_back};
  const cb_poly& g2 = (m->dim == 3) ? gradf[2] : gz;
  DensePoly fd, gd0, gd1, gd2;
  if (dense_from_poly(*f, fd) && dense_from_poly(gradf[0], gd0) && dense_from_poly(gradf[1], gd1) &&
      dense_from_poly(g2, gd2)) {
    bool zdep = false;
    int deg = dense_degree(fd, &zdep);
    for (const DensePoly* g : {&gd0, &gd1, &gd2}) {
      const int dg = dense_degree(*g, &zdep) + 1;
      if (g->planes && dg > deg) deg = dg;
    }
    if (gd2.planes) zdep = true;  // a z-derivative without z-dependence cannot come from one f, but honour it
    if (deg <= DP_DEG) {
#define CB_SHAPE_CASE(DEG, ZDEP) \
  return launch_shape_derivative<DensePoly, DEG, ZDEP>(s, m, u, p, fd, gd0, gd1, gd2, *q, co, u_d, bc_flag, rhs)
      if (zdep && m->dim == 3) { if (deg <= 2) CB_SHAPE_CASE(2, true); CB_SHAPE_CASE(4, true); }
      if (deg <= 2) CB_SHAPE_CASE(2, false);
      CB_SHAPE_CASE(4, false);
#undef CB_SHAPE_CASE
    }
  }
  return launch_shape_derivative<cb_poly, 0, false>(s, m, u, p, *f, gradf[0], gradf[1], g2, *q, co, u_d, bc_flag, rhs);
}

// the special case the C-ABI started with: kappa = 1, no reaction, J = c int u
extern "C" int cb_assemble_shape_derivative_poisson(void* stream, const cb_mesh_view* m,
                                                    const double* u, const double* p,
                                                    const cb_poly* f, const cb_poly* gradf,
                                                    const cb_quad* q, double c_int_u,
                                                    const uint8_t* bc_flag, double* rhs) {
  return cb_assemble_shape_derivative(stream, m, u, p, f, gradf, q, c_int_u, 1.0, 0.0, 0.0, nullptr, 0, bc_flag, rhs);
}

extern "C" int cb_assemble_cubic(void* stream, const cb_mesh_view* m, const double* u, double c, const cb_quad* q,
                                 const uint8_t* bc_flag, double* val, double* vec) {
  if (int e = check_view(m)) return e;
  CB_REQUIRE(u && q, "cb_assemble_cubic: missing argument");
  CB_REQUIRE(m->cells_orig && m->n2c_ptr && m->n2c_idx && m->v2c_ptr && m->v2c_idx && m->cell_scratch,
             "cb_assemble_cubic needs the gather data of the mesh view");
  cudaStream_t s = as_stream(stream);
  const int64_t nc = m->colour_off[m->ncolours];
  if (val) {
    const int grid = grid_for(m->nnz, 256, 8);
    if (m->dim == 2)
      k_add_weighted_mass_gather<2><<<grid, 256, 0, s>>>(m->nnz, m->n2c_ptr, m->n2c_idx, m->cells_orig, m->coords, u,
                                                        3.0 * c, *q, bc_flag, val);
    else
      k_add_weighted_mass_gather<3><<<grid, 256, 0, s>>>(m->nnz, m->n2c_ptr, m->n2c_idx, m->cells_orig, m->coords, u,
                                                        3.0 * c, *q, bc_flag, val);
    CB_LAUNCH_CHECK();
  }
  if (vec) {
    const int grid = grid_for(nc, 256, 8), vgrid = grid_for(m->nrows_owned, 256, 8);
    if (m->dim == 2) {
      k_cubic_load<2><<<grid, 256, 0, s>>>(nc, m->cells_orig, m->coords, u, c, *q, m->cell_scratch);
      CB_LAUNCH_CHECK();
      k_gather_cell_vectors<2, 1><<<vgrid, 256, 0, s>>>(m->nrows_owned, m->v2c_ptr, m->v2c_idx, m->cells_orig,
                                                       m->cell_scratch, vec);
    } else {
      k_cubic_load<3><<<grid, 256, 0, s>>>(nc, m->cells_orig, m->coords, u, c, *q, m->cell_scratch);
      CB_LAUNCH_CHECK();
      k_gather_cell_vectors<3, 1><<<vgrid, 256, 0, s>>>(m->nrows_owned, m->v2c_ptr, m->v2c_idx, m->cells_orig,
                                                       m->cell_scratch, vec);
    }
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

extern "C" int cb_ident_zeros(void* stream, int64_t nrows, int bs, const int64_t* rowptr,
                              const int32_t* col, double* val, int64_t* n_fixed) {
  cudaStream_t s = as_stream(stream);
  if (n_fixed) CB_CUDA_CHECK(cudaMemsetAsync(n_fixed, 0, sizeof(int64_t), s));
  k_ident_zeros<<<grid_for(nrows * bs, 128, 16), 128, 0, s>>>(nrows, bs, rowptr, col, val,
                                                              (unsigned long long*)n_fixed);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
