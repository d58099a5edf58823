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

// P[i, J] = [J == agg[i]] - omega / S_ii * sum_{k : agg[k] == J} S_ik     (k restricted to k < nlocal)
__global__ void __launch_bounds__(128)
k_smooth_prolongator(int64_t n, int64_t nlocal, const int64_t* __restrict__ s_rowptr,
                     const int32_t* __restrict__ s_col, const double* __restrict__ s_val, double omega,
                     const int32_t* __restrict__ agg, const int64_t* __restrict__ p_rowptr,
                     const int32_t* __restrict__ p_col, double* __restrict__ p_val) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pb = p_rowptr[i], pe = p_rowptr[i + 1];
    for (int64_t k = pb; k < pe; ++k) p_val[k] = 0.0;
    double sii = 0.0;
    for (int64_t k = s_rowptr[i]; k < s_rowptr[i + 1]; ++k)
      if (s_col[k] == i) sii = s_val[k];
    const double w = (sii != 0.0) ? omega / sii : 0.0;
    for (int64_t k = s_rowptr[i]; k < s_rowptr[i + 1]; ++k) {
      const int32_t c = s_col[k];
      if (c >= nlocal) continue;
      const int64_t pos = find_col(p_col, pb, pe, agg[c]);
      p_val[pos] -= w * s_val[k];
    }
    p_val[find_col(p_col, pb, pe, agg[i])] += 1.0;
  }
}

// AP[i, J] = sum_k A[i, k] * P[k, J]   (A block bs x bs, P scalar); thread = (row, block entry)
template <int BS>
__global__ void __launch_bounds__(128)
k_spgemm_ap(int64_t n, int64_t nlocal, const int64_t* __restrict__ a_rowptr, const int32_t* __restrict__ a_col,
            const double* __restrict__ a_val, const int64_t* __restrict__ p_rowptr,
            const int32_t* __restrict__ p_col, const double* __restrict__ p_val,
            const int64_t* __restrict__ c_rowptr, const int32_t* __restrict__ c_col,
            double* __restrict__ c_val) {
  constexpr int BB = BS * BS;
  const int64_t total = n * BB;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / BB;
    const int e = (int)(t - i * BB);
    const int64_t cb = c_rowptr[i], ce = c_rowptr[i + 1];
    for (int64_t k = cb; k < ce; ++k) c_val[k * BB + e] = 0.0;
    for (int64_t ka = a_rowptr[i]; ka < a_rowptr[i + 1]; ++ka) {
      const int32_t k = a_col[ka];
      if (k >= nlocal) continue;
      const double a = a_val[ka * BB + e];
      for (int64_t kp = p_rowptr[k]; kp < p_rowptr[k + 1]; ++kp) {
        const int64_t pos = find_col(c_col, cb, ce, p_col[kp]);
        c_val[pos * BB + e] += a * p_val[kp];
      }
    }
  }
}

// Ac[I, J] = sum_i Pt[I, i] * AP[i, J]
template <int BS>
__global__ void __launch_bounds__(128)
k_spgemm_ptap(int64_t nc, const int64_t* __restrict__ pt_rowptr, const int32_t* __restrict__ pt_col,
              const double* __restrict__ pt_val, const int64_t* __restrict__ ap_rowptr,
              const int32_t* __restrict__ ap_col, const double* __restrict__ ap_val,
              const int64_t* __restrict__ c_rowptr, const int32_t* __restrict__ c_col,
              double* __restrict__ c_val) {
  constexpr int BB = BS * BS;
  const int64_t total = nc * BB;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = t / BB;
    const int e = (int)(t - I * BB);
    const int64_t cb = c_rowptr[I], ce = c_rowptr[I + 1];
    for (int64_t k = cb; k < ce; ++k) c_val[k * BB + e] = 0.0;
    for (int64_t kt = pt_rowptr[I]; kt < pt_rowptr[I + 1]; ++kt) {
      const int32_t i = pt_col[kt];
      const double w = pt_val[kt];
      for (int64_t ka = ap_rowptr[i]; ka < ap_rowptr[i + 1]; ++ka) {
        const int64_t pos = find_col(c_col, cb, ce, ap_col[ka]);
        c_val[pos * BB + e] += w * ap_val[ka * BB + e];
      }
    }
  }
}

// bc[I, c] = sum_i Pt[I, i] * (b[i, c] - t[i, c])      (restriction of the residual)
template <int BS>
__global__ void __launch_bounds__(256)
k_restrict(int64_t nc, const int64_t* __restrict__ pt_rowptr, const int32_t* __restrict__ pt_col,
           const double* __restrict__ pt_val, const double* __restrict__ b, const double* __restrict__ t,
           double* __restrict__ bc, const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t total = nc * BS;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t I = q / BS;
    const int c = (int)(q - I * BS);
    double s = 0.0;
    for (int64_t k = pt_rowptr[I]; k < pt_rowptr[I + 1]; ++k) {
      const int64_t i = (int64_t)pt_col[k] * BS + c;
      s += pt_val[k] * (b[i] - t[i]);
    }
    bc[q] = s;
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

int amg_restrict(cudaStream_t s, int bs, const cb_amg_level* L, const double* b, const double* t, double* bc,
                 const int* done) {
  const int grid = grid_for(L->nc * bs, 256, 8);
  if (bs == 1) k_restrict<1><<<grid, 256, 0, s>>>(L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  else if (bs == 2) k_restrict<2><<<grid, 256, 0, s>>>(L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  else k_restrict<3><<<grid, 256, 0, s>>>(L->nc, L->pt_rowptr, L->pt_col, L->pt_val, b, t, bc, done);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int amg_prolong(cudaStream_t s, int bs, const cb_amg_level* L, const double* xc, double* x, const int* done) {
  const int64_t n = L->A.nrows;
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
                                         const int64_t* p_rowptr, const int32_t* p_col, double* p_val) {
  CB_REQUIRE(S && S->bs == 1 && agg && p_rowptr && p_col && p_val, "cb_amg_smooth_prolongator: bad argument");
  k_smooth_prolongator<<<grid_for(S->nrows, 128, 16), 128, 0, as_stream(stream)>>>(
      S->nrows, S->nrows, S->rowptr, S->col, S->val, omega, agg, p_rowptr, p_col, p_val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_galerkin(void* stream, const cb_mat* A, const int64_t* p_rowptr, const int32_t* p_col,
                               const double* p_val, const int64_t* pt_rowptr, const int32_t* pt_col,
                               const double* pt_val, const int64_t* ap_rowptr, const int32_t* ap_col,
                               double* ap_val, const cb_mat* Ac) {
  CB_REQUIRE(A && Ac && p_rowptr && pt_rowptr && ap_rowptr && ap_val, "cb_amg_galerkin: missing argument");
  CB_REQUIRE(A->bs == Ac->bs && A->bs >= 1 && A->bs <= 3, "cb_amg_galerkin: block sizes");
  cudaStream_t s = as_stream(stream);
  const int bs = A->bs;
  const int64_t n = A->nrows, nc = Ac->nrows;
  const int g1 = grid_for(n * bs * bs, 128, 16), g2 = grid_for(nc * bs * bs, 128, 16);
  double* acv = const_cast<double*>(Ac->val);
#define CB_GALERKIN(BS)                                                                                   \
  k_spgemm_ap<BS><<<g1, 128, 0, s>>>(n, n, A->rowptr, A->col, A->val, p_rowptr, p_col, p_val, ap_rowptr, \
                                     ap_col, ap_val);                                                     \
  CB_LAUNCH_CHECK();                                                                                      \
  k_spgemm_ptap<BS><<<g2, 128, 0, s>>>(nc, pt_rowptr, pt_col, pt_val, ap_rowptr, ap_col, ap_val,         \
                                       Ac->rowptr, Ac->col, acv);                                         \
  CB_LAUNCH_CHECK();
  if (bs == 1) { CB_GALERKIN(1) } else if (bs == 2) { CB_GALERKIN(2) } else { CB_GALERKIN(3) }
#undef CB_GALERKIN
  return CB_OK;
}
