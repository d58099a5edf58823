// Smoothed-aggregation algebraic multigrid (the device counterpart of the reference's
// `pc_type hypre` / boomeramg default, cashocs/_utils/linalg.py:44-52, and of PETSc's `gamg`).
//
// Unknown-based smoothed aggregation with one scalar prolongator P_s shared by all bs components
// (P = P_s (x) I_bs): every level stays a block-CSR matrix with the fine block size, so all level
// operators run through the same TMA-staged SpMV.  The symbolic phase (aggregates, patterns of P,
// A*P, P^T*A*P) depends only on the sparsity pattern and is done once per mesh topology by the host
// layer; this file holds the numeric phase (values of P and of the Galerkin operators, re-done for
// every new matrix) and the V-cycle kernels.  Everything is gather-formulated (one thread owns one
// output entry, fixed summation order): no atomics, bit-reproducible.
#include "common.cuh"

namespace cb {

// S = trace(A_blk) / bs : scalar "strength" matrix with the pattern of A
__global__ void __launch_bounds__(256)
k_block_trace(int64_t nnzb, int bs, const double* __restrict__ val, double* __restrict__ s) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnzb;
       k += (int64_t)gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int i = 0; i < bs; ++i) t += val[(k * bs + i) * bs + i];
    s[k] = t / bs;
  }
}

__device__ __forceinline__ int64_t find_col(const int32_t* __restrict__ col, int64_t lo, int64_t hi, int32_t c) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// P[i, J] = [J == agg[i]] - omega / S^F_ii * sum_{k < nlocal : agg[k] == J} S^F_ik, where S^F is S
// filtered to the local columns with the dropped (ghost) couplings lumped into the diagonal:
// S^F_ii = S_ii + sum_{k >= nlocal} S_ik.  Row sums of S are preserved, so the rows of P keep
// summing to 1 - omega * rowsum(S) / S^F_ii (partition of unity for Laplace-like rows) although
// aggregates and prolongator rows never reach across a partition boundary.
__global__ void __launch_bounds__(128)
k_smooth_prolongator(int64_t n, int64_t nlocal, const int64_t* __restrict__ s_rowptr,
                     const int32_t* __restrict__ s_col, const double* __restrict__ s_val, double omega,
                     const int32_t* __restrict__ agg, const uint8_t* __restrict__ tentative,
                     const int64_t* __restrict__ p_rowptr, const int32_t* __restrict__ p_col,
                     double* __restrict__ p_val) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pb = p_rowptr[i], pe = p_rowptr[i + 1];
    if (tentative && tentative[i]) {  // keep the unsmoothed row: P[i, agg(i)] = 1
      for (int64_t k = pb; k < pe; ++k) p_val[k] = (p_col[k] == agg[i]) ? 1.0 : 0.0;
      continue;
    }
    for (int64_t k = pb; k < pe; ++k) p_val[k] = 0.0;
    double sii = 0.0, lumped = 0.0;
    for (int64_t k = s_rowptr[i]; k < s_rowptr[i + 1]; ++k) {
      const int32_t c = s_col[k];
      if (c == i) sii = s_val[k];
      else if (c >= nlocal) lumped += s_val[k];
    }
    const double sf = sii + lumped;
    const double w = (sf != 0.0) ? omega / sf : 0.0;
    for (int64_t k = s_rowptr[i]; k < s_rowptr[i + 1]; ++k) {
      const int32_t c = s_col[k];
      if (c >= nlocal) continue;
      const int64_t pos = find_col(p_col, pb, pe, agg[c]);
      p_val[pos] -= w * ((c == i) ? sf : s_val[k]);
    }
    p_val[find_col(p_col, pb, pe, agg[i])] += 1.0;
  }
}

// Numeric sparse product with a precomputed term list: for output entry q the terms
// t in [seg_ptr[q], seg_ptr[q+1]) name one lhs entry and one rhs entry; exactly one of the two
// operands carries bs x bs blocks (A or A*P), the other is scalar (P or P^T).  One thread owns one
// (entry, block element): pure gather, no search, no atomics, fixed summation order.
template <int BS, bool LHS_BLOCK>
__global__ void __launch_bounds__(256)
k_spgemm_terms(int64_t nnz_out, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ src_l,
               const int32_t* __restrict__ src_r, const double* __restrict__ lhs,
               const double* __restrict__ rhs, double* __restrict__ out) {
  constexpr int BB = BS * BS;
  const int64_t total = nnz_out * BB;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / BB;
    const int e = (int)(t - q * BB);
    double acc = 0.0;
    for (int64_t k = seg_ptr[q]; k < seg_ptr[q + 1]; ++k) {
      const int64_t l = src_l[k], r = src_r[k];
      if (LHS_BLOCK) acc += lhs[l * BB + e] * rhs[r];
      else acc += lhs[l] * rhs[r * BB + e];
    }
    out[t] = acc;
  }
}

// The same product without term lists (they cost 8 bytes per scalar product term, i.e. 2-4x the matrix
// itself, which does not fit for 10^9-nnz operators): thread (row i, block element e) walks row i of L
// and the rows of R it references, finds the target column in the known pattern of the output row by
// binary search and accumulates.  Same summation order as the term lists (L entries ascending, then R
// entries ascending) -> bitwise the same result.
template <int BS, bool LHS_BLOCK>
__global__ void __launch_bounds__(256)
k_spgemm_rowwise(int64_t nrows, const int64_t* __restrict__ l_rowptr, const int32_t* __restrict__ l_col,
                 const double* __restrict__ l_val, const int64_t* __restrict__ r_rowptr,
                 const int32_t* __restrict__ r_col, const double* __restrict__ r_val,
                 const int64_t* __restrict__ o_rowptr, const int32_t* __restrict__ o_col,
                 double* __restrict__ out) {
  constexpr int BB = BS * BS;
  const int64_t total = nrows * BB;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / BB;
    const int e = (int)(t - i * BB);
    const int64_t ob = o_rowptr[i], oe = o_rowptr[i + 1];
    for (int64_t q = ob; q < oe; ++q) out[q * BB + e] = 0.0;
    for (int64_t a = l_rowptr[i]; a < l_rowptr[i + 1]; ++a) {
      const int64_t k = l_col[a];
      const double lv = LHS_BLOCK ? l_val[a * BB + e] : l_val[a];
      for (int64_t b = r_rowptr[k]; b < r_rowptr[k + 1]; ++b) {
        const double rv = LHS_BLOCK ? r_val[b] : r_val[b * BB + e];
        const int64_t pos = find_col(o_col, ob, oe, r_col[b]);
        out[pos * BB + e] += lv * rv;
      }
    }
  }
}

// Dense inverse of the coarsest operator (n <= 256 scalar rows): Gauss-Jordan without pivoting
// (the matrix is SPD) by ONE CTA in global memory; `work` holds the n x 2n augmented matrix.
__global__ void __launch_bounds__(256)
k_dense_inverse(int n, int bs, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const double* __restrict__ val, double* __restrict__ work, double* __restrict__ inv) {
  const int w = 2 * n;
  for (int i = threadIdx.x; i < n * w; i += blockDim.x) work[i] = 0.0;
  __syncthreads();
  const int nb = n / bs;
  for (int br = threadIdx.x; br < nb; br += blockDim.x) {
    for (int64_t k = rowptr[br]; k < rowptr[br + 1]; ++k) {
      const int bc = col[k];
      if (bc >= nb) continue;
      for (int i = 0; i < bs; ++i)
        for (int j = 0; j < bs; ++j) work[(br * bs + i) * w + bc * bs + j] = val[(k * bs + i) * bs + j];
    }
    for (int i = 0; i < bs; ++i) work[(br * bs + i) * w + n + br * bs + i] = 1.0;
  }
  __syncthreads();
  for (int p = 0; p < n; ++p) {
    const double piv = 1.0 / work[p * w + p];
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += blockDim.x) work[p * w + j] *= piv;
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * w; idx += blockDim.x) {
      const int r = idx / w, j = idx - r * w;
      if (r == p || j == p) continue;
      work[idx] -= work[r * w + p] * work[p * w + j];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += blockDim.x)
      if (r != p) work[r * w + p] = 0.0;
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int r = idx / n, j = idx - r * n;
    inv[idx] = work[r * w + n + j];
  }
}

// x = Ainv * b (dense, row-major), one warp per row
__global__ void __launch_bounds__(256)
k_dense_apply(int n, const double* __restrict__ inv, const double* __restrict__ b, double* __restrict__ x,
              const int* __restrict__ done) {
  if (done && *done) return;
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += inv[(int64_t)row * n + j] * b[j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if (lane == 0) x[row] = s;
}

int amg_dense_apply(cudaStream_t s, int n, const double* inv, const double* b, double* x, const int* done) {
  k_dense_apply<<<(n * 32 + 255) / 256, 256, 0, s>>>(n, inv, b, x, done);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// bc[I, c] = sum_i Pt[I, i] * (b[i, c] - t[i, c])      (restriction of the residual; t == NULL: b is the residual)
template <int BS>
__global__ void __launch_bounds__(256)
k_restrict(int64_t nc, const int64_t* __restrict__ pt_rowptr, const int32_t* __restrict__ pt_col,
           const double* __restrict__ pt_val, const double* __restrict__ b, const double* __restrict__ t,
           double* __restrict__ bc, const int* __restrict__ done) {
  if (done && *done) return;
  // One warp per coarse node, lane = e * BS + c: a sweep covers EPW = 32 / BS entries of the P^T row (~100 per
  // row) and lane (e, c) owns component c of entry e, so the gather of one vector is ONE load instruction whose
  // addresses are EPW runs of BS contiguous doubles -- the kernel is bound by L1 wavefronts (sectors touched per
  // load instruction), and one-lane-per-entry with BS loads per vector touched every sector BS times.
  // Fixed-order shuffle tree over e at the end: bit-reproducible.
  constexpr int EPW = 32 / BS;
  const int lane = threadIdx.x & 31;
  const int e = lane / BS, c = lane - e * BS;
  const bool active = e < EPW;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t I = warp0; I < nc; I += nwarps) {
    double s = 0.0;
    const int64_t kb = pt_rowptr[I], ke = pt_rowptr[I + 1];
    if (active) {
      if (t) {
#pragma unroll 2
        for (int64_t k = kb + e; k < ke; k += EPW) {
          const int64_t i = (int64_t)__ldg(pt_col + k) * BS + c;
          s = fma(__ldg(pt_val + k), b[i] - t[i], s);
        }
      } else {
#pragma unroll 4
        for (int64_t k = kb + e; k < ke; k += EPW) {
          const int64_t i = (int64_t)__ldg(pt_col + k) * BS + c;
          s = fma(__ldg(pt_val + k), b[i], s);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      if (off < EPW) {
        const double o = __shfl_down_sync(0xffffffffu, s, off * BS);
        if (lane + off * BS < 32) s += o;
      }
    }
    if (lane < BS) bc[I * BS + lane] = s;
  }
}

// x[i, c] += sum_J P[i, J] * xc[J, c]                     (coarse-grid correction)
template <int BS>
__global__ void __launch_bounds__(256)
k_prolong(int64_t n, const int64_t* __restrict__ p_rowptr, const int32_t* __restrict__ p_col,
          const double* __restrict__ p_val, const double* __restrict__ xc, double* __restrict__ x,
          const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t total = n * BS;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / BS;
    const int c = (int)(q - i * BS);
    double s = 0.0;
    for (int64_t k = p_rowptr[i]; k < p_rowptr[i + 1]; ++k) s += p_val[k] * xc[(int64_t)p_col[k] * BS + c];
    x[q] += s;
  }
}

// block-prolongator variants (csrc/amg_rbm.cu)
int rbm_restrict_launch(cudaStream_t s, int64_t ncn, const int64_t* pt_rowptr, const int32_t* pt_col, const double* pt_val,
                        const double* b, const double* t, double* bc, const int* done);
int rbm_prolong_launch(cudaStream_t s, int64_t n, const int64_t* p_rowptr, const int32_t* p_col, const double* p_val,
                       const double* xc, double* x, const int* done);

int amg_restrict(cudaStream_t s, int bs, const cb_amg_level* L, const double* b, const double* t, double* bc,
                 const int* done, int p_block) {
  if (p_block) return rbm_restrict_launch(s, L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  const int grid = grid_for(L->nc * 32, 256, 8);
  if (bs == 1) k_restrict<1><<<grid, 256, 0, s>>>(L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  else if (bs == 2) k_restrict<2><<<grid, 256, 0, s>>>(L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  else k_restrict<3><<<grid, 256, 0, s>>>(L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int amg_prolong(cudaStream_t s, int bs, const cb_amg_level* L, const double* xc, double* x, const int* done, int p_block) {
  const int64_t n = L->A.nrows;
  if (p_block) return rbm_prolong_launch(s, n, L->p_rowptr, L->p_col, L->p_val, xc, x, done);
  const int grid = grid_for(n * bs, 256, 8);
  if (bs == 1) k_prolong<1><<<grid, 256, 0, s>>>(n, L->p_rowptr, L->p_col, L->p_val, xc, x, done);
  else if (bs == 2) k_prolong<2><<<grid, 256, 0, s>>>(n, L->p_rowptr, L->p_col, L->p_val, xc, x, done);
  else k_prolong<3><<<grid, 256, 0, s>>>(n, L->p_rowptr, L->p_col, L->p_val, xc, x, done);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

}  // namespace cb

using namespace cb;

extern "C" int cb_amg_block_trace(void* stream, const cb_mat* A, double* s_val) {
  CB_REQUIRE(A && s_val, "cb_amg_block_trace: missing argument");
  int64_t nnzb = 0;
  cudaStream_t s = as_stream(stream);
  CB_CUDA_CHECK(cudaMemcpyAsync(&nnzb, A->rowptr + A->nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  CB_CUDA_CHECK(cudaStreamSynchronize(s));
  k_block_trace<<<grid_for(nnzb, 256, 8), 256, 0, s>>>(nnzb, A->bs, A->val, s_val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_smooth_prolongator(void* stream, const cb_mat* S, double omega, const int32_t* agg,
                                         const uint8_t* tentative, const int64_t* p_rowptr,
                                         const int32_t* p_col, double* p_val) {
  CB_REQUIRE(S && S->bs == 1 && agg && p_rowptr && p_col && p_val, "cb_amg_smooth_prolongator: bad argument");
  k_smooth_prolongator<<<grid_for(S->nrows, 128, 16), 128, 0, as_stream(stream)>>>(
      S->nrows, S->nrows, S->rowptr, S->col, S->val, omega, agg, tentative, p_rowptr, p_col, p_val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_spgemm_terms(void* stream, int bs, int lhs_block, int64_t nnz_out,
                                   const int64_t* seg_ptr, const int32_t* src_l, const int32_t* src_r,
                                   const double* lhs, const double* rhs, double* out) {
  CB_REQUIRE(seg_ptr && src_l && src_r && lhs && rhs && out, "cb_amg_spgemm_terms: missing argument");
  CB_REQUIRE(bs >= 1 && bs <= 3, "cb_amg_spgemm_terms: bs must be 1..3");
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(nnz_out * bs * bs, 256, 16);
#define CB_TERMS(BS)                                                                                            \
  if (lhs_block) k_spgemm_terms<BS, true><<<grid, 256, 0, s>>>(nnz_out, seg_ptr, src_l, src_r, lhs, rhs, out); \
  else k_spgemm_terms<BS, false><<<grid, 256, 0, s>>>(nnz_out, seg_ptr, src_l, src_r, lhs, rhs, out);
  if (bs == 1) { CB_TERMS(1) } else if (bs == 2) { CB_TERMS(2) } else { CB_TERMS(3) }
#undef CB_TERMS
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_spgemm_rowwise(void* stream, int bs, int lhs_block, int64_t nrows, const int64_t* l_rowptr,
                                     const int32_t* l_col, const double* l_val, const int64_t* r_rowptr,
                                     const int32_t* r_col, const double* r_val, const int64_t* o_rowptr,
                                     const int32_t* o_col, double* out) {
  CB_REQUIRE(l_rowptr && l_col && l_val && r_rowptr && r_col && r_val && o_rowptr && o_col && out,
             "cb_amg_spgemm_rowwise: missing argument");
  CB_REQUIRE(bs >= 1 && bs <= 3, "cb_amg_spgemm_rowwise: bs must be 1..3");
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(nrows * bs * bs, 256, 16);
#define CB_ROWWISE(BS)                                                                                       \
  if (lhs_block) k_spgemm_rowwise<BS, true><<<grid, 256, 0, s>>>(nrows, l_rowptr, l_col, l_val, r_rowptr,    \
                                                                r_col, r_val, o_rowptr, o_col, out);        \
  else k_spgemm_rowwise<BS, false><<<grid, 256, 0, s>>>(nrows, l_rowptr, l_col, l_val, r_rowptr, r_col,      \
                                                       r_val, o_rowptr, o_col, out);
  if (bs == 1) { CB_ROWWISE(1) } else if (bs == 2) { CB_ROWWISE(2) } else { CB_ROWWISE(3) }
#undef CB_ROWWISE
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_dense_inverse(void* stream, const cb_mat* A, double* work, double* inv) {
  CB_REQUIRE(A && work && inv, "cb_amg_dense_inverse: missing argument");
  const int n = (int)(A->nrows * A->bs);
  CB_REQUIRE(n >= 1 && n <= 256, "cb_amg_dense_inverse: coarsest level must have <= 256 scalar rows");
  k_dense_inverse<<<1, 256, 0, as_stream(stream)>>>(n, A->bs, A->rowptr, A->col, A->val, work, inv);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
