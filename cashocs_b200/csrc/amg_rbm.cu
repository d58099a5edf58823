// Numeric kernels of the smoothed-aggregation hierarchy that carries the six rigid-body modes (3-D vector
// problems, 3 x 3 blocks): every aggregate J is represented by two coarse nodes, 2J (translations) and 2J+1
// (rotations about the aggregate centroid, scaled by its radius).  Symbolic phase and data layout:
// cashocs_b200/_utils/amg_rbm.py.
//
// WORK IN PROGRESS (round 2): not yet called from cb_ksp_solve.  Every kernel below is restated one-to-one in
// numpy by tests/test_amg_rbm.py (emulate_*), where the arithmetic is checked against scipy over the same
// term lists; the kernels have been compiled for sm_100a but not yet executed on a GPU.
//
// Like csrc/amg.cu everything is gather-formulated: one thread (or warp) owns one output, fixed summation order,
// no atomics -> bit-reproducible.
#include "common.cuh"

namespace cb {

// C[a] = mean position of the translation nodes of aggregate a; S[a] = max_i (|x_i - C[a]| + s_i), 1 if that is 0
__global__ void __launch_bounds__(128)
k_rbm_aggregate_geometry(int64_t nc, const int64_t* __restrict__ member_ptr, const int32_t* __restrict__ member_idx,
                         const uint8_t* __restrict__ kind, const double* __restrict__ centre,
                         const double* __restrict__ scale, double* __restrict__ C, double* __restrict__ S) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < nc; a += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = member_ptr[a], e = member_ptr[a + 1];
    double c[3] = {0.0, 0.0, 0.0};
    int cnt = 0;
    for (int64_t k = b; k < e; ++k) {
      const int64_t i = member_idx[k];
      if (kind[i] == 0) {
        c[0] += centre[i * 3 + 0]; c[1] += centre[i * 3 + 1]; c[2] += centre[i * 3 + 2];
        ++cnt;
      }
    }
    const double inv = 1.0 / (cnt > 0 ? cnt : 1);
    c[0] *= inv; c[1] *= inv; c[2] *= inv;
    double rad = 0.0;
    for (int64_t k = b; k < e; ++k) {
      const int64_t i = member_idx[k];
      const double dx = centre[i * 3 + 0] - c[0], dy = centre[i * 3 + 1] - c[1], dz = centre[i * 3 + 2] - c[2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz) + scale[i];
      rad = r > rad ? r : rad;
    }
    C[a * 3 + 0] = c[0]; C[a * 3 + 1] = c[1]; C[a * 3 + 2] = c[2];
    S[a] = rad > 0.0 ? rad : 1.0;
  }
}

// Tentative prolongator, one thread per fine node.  Translation node i in J: T[i,2J] = I,
// T[i,2J+1] = cross(x_i - C_J) / S_J (the displacement w x r of a rotation w); rotation node: T[i,2J+1] = s_i/S_J I.
__global__ void __launch_bounds__(256)
k_rbm_tentative(int64_t n, const int64_t* __restrict__ t_rowptr, const int32_t* __restrict__ agg,
                const uint8_t* __restrict__ kind, const double* __restrict__ centre, const double* __restrict__ scale,
                const double* __restrict__ C, const double* __restrict__ S, double* __restrict__ T) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = agg[i];
    double* t = T + t_rowptr[i] * 9;
    const double is = 1.0 / S[a];
    if (kind[i] == 0) {
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = (e % 4 == 0) ? 1.0 : 0.0;
      const double r0 = (centre[i * 3 + 0] - C[a * 3 + 0]) * is, r1 = (centre[i * 3 + 1] - C[a * 3 + 1]) * is,
                   r2 = (centre[i * 3 + 2] - C[a * 3 + 2]) * is;
      double* u = t + 9;  // [[0, r2, -r1], [-r2, 0, r0], [r1, -r0, 0]]
      u[0] = 0.0; u[1] = r2;  u[2] = -r1;
      u[3] = -r2; u[4] = 0.0; u[5] = r0;
      u[6] = r1;  u[7] = -r0; u[8] = 0.0;
    } else {
      const double f = scale[i] * is;
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = (e % 4 == 0) ? f : 0.0;
    }
  }
}

// out[q] = sum over the terms k of segment q (ascending) of op(lhs[src_l[k]]) * rhs[src_r[k]], 3 x 3 blocks;
// op = transpose for the P^T (A P) product.  One thread per (output block, element).
template <bool TRANS_L>
__global__ void __launch_bounds__(256)
k_spgemm_terms_bb(int64_t nnz_out, const int64_t* __restrict__ seg_ptr, const int32_t* __restrict__ src_l,
                  const int32_t* __restrict__ src_r, const double* __restrict__ lhs, const double* __restrict__ rhs,
                  double* __restrict__ out) {
  const int64_t total = nnz_out * 9;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = t / 9;
    const int e = (int)(t - q * 9);
    const int r = e / 3, c = e - 3 * r;
    double acc = 0.0;
    for (int64_t k = seg_ptr[q]; k < seg_ptr[q + 1]; ++k) {
      const double* L = lhs + (int64_t)src_l[k] * 9;
      const double* R = rhs + (int64_t)src_r[k] * 9;
      if (TRANS_L) acc += L[r] * R[c] + L[3 + r] * R[3 + c] + L[6 + r] * R[6 + c];
      else acc += L[3 * r] * R[c] + L[3 * r + 1] * R[3 + c] + L[3 * r + 2] * R[6 + c];
    }
    out[t] = acc;
  }
}

// P = -omega * D^-1 (A T) in place (P holds A T on entry; D = diag(A), the smoothers' point Jacobi) ...
__global__ void __launch_bounds__(256)
k_rbm_scale_rows(int64_t n, const int64_t* __restrict__ p_rowptr, const double* __restrict__ dinv, double omega,
                 double* __restrict__ P) {
  const int64_t total = n * 3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / 3;
    const int r = (int)(t - i * 3);
    const double w = -omega * dinv[t];
    for (int64_t q = p_rowptr[i]; q < p_rowptr[i + 1]; ++q) {
      double* b = P + q * 9 + 3 * r;
      b[0] *= w; b[1] *= w; b[2] *= w;
    }
  }
}

// ... + T on T's positions (distinct targets, no race)
__global__ void __launch_bounds__(256)
k_rbm_add_tentative(int64_t nt, const int64_t* __restrict__ t2p, const double* __restrict__ T, double* __restrict__ P) {
  const int64_t total = nt * 9;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / 9;
    P[t2p[k] * 9 + (t - k * 9)] += T[t];
  }
}

// pt_val[k] = transpose(p_val[perm[k]]): the blocks of P^T in P^T's entry order (coalesced restriction)
__global__ void __launch_bounds__(256)
k_rbm_transpose_blocks(int64_t nnz, const int64_t* __restrict__ perm, const double* __restrict__ p_val,
                       double* __restrict__ pt_val) {
  const int64_t total = nnz * 9;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / 9;
    const int e = (int)(t - k * 9);
    const int r = e / 3, c = e - 3 * r;
    pt_val[t] = p_val[perm[k] * 9 + 3 * c + r];
  }
}

// bc[I] = sum_k Pt[k] (b - t)[pt_col[k]]: one warp per coarse node, lanes stride over the row, fixed shuffle tree
__global__ void __launch_bounds__(256)
k_rbm_restrict(int64_t ncn, const int64_t* __restrict__ pt_rowptr, const int32_t* __restrict__ pt_col,
               const double* __restrict__ pt_val, const double* __restrict__ b, const double* __restrict__ t,
               double* __restrict__ bc, const int* __restrict__ done) {
  if (done && *done) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t I = warp0; I < ncn; I += nwarps) {
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t k = pt_rowptr[I] + lane; k < pt_rowptr[I + 1]; k += 32) {
      const int64_t i = (int64_t)pt_col[k] * 3;
      const double v0 = b[i] - t[i], v1 = b[i + 1] - t[i + 1], v2 = b[i + 2] - t[i + 2];
      const double* m = pt_val + k * 9;
      s[0] += m[0] * v0 + m[1] * v1 + m[2] * v2;
      s[1] += m[3] * v0 + m[4] * v1 + m[5] * v2;
      s[2] += m[6] * v0 + m[7] * v1 + m[8] * v2;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s[c] += __shfl_down_sync(0xffffffffu, s[c], off);
    }
    if (lane == 0) { bc[I * 3] = s[0]; bc[I * 3 + 1] = s[1]; bc[I * 3 + 2] = s[2]; }
  }
}

// x[i] += sum_k P[k] xc[p_col[k]]: one thread per fine scalar row
__global__ void __launch_bounds__(256)
k_rbm_prolong(int64_t n, const int64_t* __restrict__ p_rowptr, const int32_t* __restrict__ p_col,
              const double* __restrict__ p_val, const double* __restrict__ xc, double* __restrict__ x,
              const int* __restrict__ done) {
  if (done && *done) return;
  const int64_t total = n * 3;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / 3;
    const int r = (int)(q - i * 3);
    double s = 0.0;
    for (int64_t k = p_rowptr[i]; k < p_rowptr[i + 1]; ++k) {
      const double* m = p_val + k * 9 + 3 * r;
      const double* v = xc + (int64_t)p_col[k] * 3;
      s += m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    }
    x[q] += s;
  }
}

int rbm_restrict_launch(cudaStream_t s, int64_t ncn, const int64_t* pt_rowptr, const int32_t* pt_col, const double* pt_val,
                        const double* b, const double* t, double* bc, const int* done) {
  k_rbm_restrict<<<grid_for(ncn * 32, 256, 8), 256, 0, s>>>(ncn, pt_rowptr, pt_col, pt_val, b, t, bc, done);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int rbm_prolong_launch(cudaStream_t s, int64_t n, const int64_t* p_rowptr, const int32_t* p_col, const double* p_val,
                       const double* xc, double* x, const int* done) {
  k_rbm_prolong<<<grid_for(n * 3, 256, 8), 256, 0, s>>>(n, p_rowptr, p_col, p_val, xc, x, done);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

}  // namespace cb

using namespace cb;

extern "C" int cb_amg_rbm_aggregate_geometry(void* stream, int64_t nc, const int64_t* member_ptr,
                                             const int32_t* member_idx, const uint8_t* kind, const double* centre,
                                             const double* scale, double* C, double* S) {
  CB_REQUIRE(member_ptr && member_idx && kind && centre && scale && C && S, "cb_amg_rbm_aggregate_geometry: missing argument");
  k_rbm_aggregate_geometry<<<grid_for(nc, 128, 16), 128, 0, as_stream(stream)>>>(nc, member_ptr, member_idx, kind, centre,
                                                                                   scale, C, S);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_rbm_tentative(void* stream, int64_t n, const int64_t* t_rowptr, const int32_t* agg,
                                    const uint8_t* kind, const double* centre, const double* scale, const double* C,
                                    const double* S, double* t_val) {
  CB_REQUIRE(t_rowptr && agg && kind && centre && scale && C && S && t_val, "cb_amg_rbm_tentative: missing argument");
  k_rbm_tentative<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(n, t_rowptr, agg, kind, centre, scale, C, S, t_val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_rbm_spgemm_terms(void* stream, int trans_l, int64_t nnz_out, const int64_t* seg_ptr,
                                       const int32_t* src_l, const int32_t* src_r, const double* lhs, const double* rhs,
                                       double* out) {
  CB_REQUIRE(seg_ptr && src_l && src_r && lhs && rhs && out, "cb_amg_rbm_spgemm_terms: missing argument");
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(nnz_out * 9, 256, 16);
  if (trans_l) k_spgemm_terms_bb<true><<<grid, 256, 0, s>>>(nnz_out, seg_ptr, src_l, src_r, lhs, rhs, out);
  else k_spgemm_terms_bb<false><<<grid, 256, 0, s>>>(nnz_out, seg_ptr, src_l, src_r, lhs, rhs, out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_rbm_prolongator_finish(void* stream, int64_t n, const int64_t* p_rowptr, int64_t nt,
                                             const int64_t* t2p, const double* t_val, const double* dinv, double omega,
                                             double* p_val) {
  CB_REQUIRE(p_rowptr && t2p && t_val && dinv && p_val, "cb_amg_rbm_prolongator_finish: missing argument");
  cudaStream_t s = as_stream(stream);
  k_rbm_scale_rows<<<grid_for(n * 3, 256, 8), 256, 0, s>>>(n, p_rowptr, dinv, omega, p_val);
  CB_LAUNCH_CHECK();
  k_rbm_add_tentative<<<grid_for(nt * 9, 256, 8), 256, 0, s>>>(nt, t2p, t_val, p_val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_rbm_transpose_blocks(void* stream, int64_t nnz, const int64_t* perm, const double* p_val,
                                           double* pt_val) {
  CB_REQUIRE(perm && p_val && pt_val, "cb_amg_rbm_transpose_blocks: missing argument");
  k_rbm_transpose_blocks<<<grid_for(nnz * 9, 256, 8), 256, 0, as_stream(stream)>>>(nnz, perm, p_val, pt_val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_amg_rbm_restrict(void* stream, int64_t ncn, const int64_t* pt_rowptr, const int32_t* pt_col,
                                   const double* pt_val, const double* b, const double* t, double* bc) {
  CB_REQUIRE(pt_rowptr && pt_col && pt_val && b && t && bc, "cb_amg_rbm_restrict: missing argument");
  return rbm_restrict_launch(as_stream(stream), ncn, pt_rowptr, pt_col, pt_val, b, t, bc, nullptr);
}

extern "C" int cb_amg_rbm_prolong(void* stream, int64_t n, const int64_t* p_rowptr, const int32_t* p_col,
                                  const double* p_val, const double* xc, double* x) {
  CB_REQUIRE(p_rowptr && p_col && p_val && xc && x, "cb_amg_rbm_prolong: missing argument");
  return rbm_prolong_launch(as_stream(stream), n, p_rowptr, p_col, p_val, xc, x, nullptr);
}
