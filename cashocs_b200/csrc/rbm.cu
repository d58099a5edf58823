// Set-up of the rigid-body coarse correction  M + Z (Z^T A Z)^-1 Z^T  of the AMG preconditioner (cb_ksp_opts.rbm_*):
// everything per matrix version runs here -- centred coordinates, the k Galerkin columns A z_j projected onto Z, the
// k x k inverse -- so the per-evaluation path holds no torch reductions, no cuBLAS call and no host round trip.
// Z: translations e_a and rotations about `centre` (k = 3 in 2-D, 6 in 3-D), evaluated on the fly from the coordinates.
// The reference reaches the same effect through BoomerAMG's near-null-space handling (cashocs/_utils/linalg.py:44-52).
#include "common.cuh"

namespace cb {

int spmv_dispatch(cudaStream_t s, const cb_mat* A, const double* x, double* y, const double* w,
                  double* result, double* scratch, const int* done = nullptr, bool use32 = false);

template <int BS>
__device__ __forceinline__ void rbm_mode(int j, const double (&c)[BS], double (&z)[BS]) {
#pragma unroll
  for (int a = 0; a < BS; ++a) z[a] = 0.0;
  if (j < BS) { z[j] = 1.0; return; }
  if constexpr (BS == 3) {
    if (j == 3) { z[0] = -c[1]; z[1] = c[0]; }
    else if (j == 4) { z[1] = -c[2]; z[2] = c[1]; }
    else { z[0] = c[2]; z[2] = -c[0]; }
  } else {
    z[0] = -c[1]; z[1] = c[0];
  }
}

// c = coords - centre for every local node (owned + ghost)
template <int BS>
__global__ void __launch_bounds__(256)
k_rbm_centre(int64_t nn, const double* __restrict__ coords, double c0, double c1, double c2, double* __restrict__ c) {
  const double ctr[3] = {c0, c1, c2};
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nn * BS; q += (int64_t)gridDim.x * blockDim.x)
    c[q] = coords[q] - ctr[q % BS];
}

// v = z_j on every local node
template <int BS>
__global__ void __launch_bounds__(256)
k_rbm_mode_vector(int64_t nn, int j, const double* __restrict__ c, double* __restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double ci[BS], z[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) ci[a] = c[i * BS + a];
    rbm_mode<BS>(j, ci, z);
#pragma unroll
    for (int a = 0; a < BS; ++a) v[i * BS + a] = z[a];
  }
}

// out[0..K) = Z^T w over the owned nodes (fixed-order grid reduction)
template <int BS>
__global__ void __launch_bounds__(256)
k_rbm_project_plain(int64_t nn, const double* __restrict__ w, const double* __restrict__ c, double* __restrict__ out,
                    double* __restrict__ scratch) {
  constexpr int K = BS == 3 ? 6 : 3;
  __shared__ double s_red[32 * K + 1];
  double v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn; i += (int64_t)gridDim.x * blockDim.x) {
    double wi[BS], ci[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) { wi[a] = w[i * BS + a]; ci[a] = c[i * BS + a]; }
#pragma unroll
    for (int a = 0; a < BS; ++a) v[a] += wi[a];
    if constexpr (BS == 3) {
      v[3] += ci[0] * wi[1] - ci[1] * wi[0];
      v[4] += ci[1] * wi[2] - ci[2] * wi[1];
      v[5] += ci[2] * wi[0] - ci[0] * wi[2];
    } else {
      v[2] += ci[0] * wi[1] - ci[1] * wi[0];
    }
  }
  int ops[K];
#pragma unroll
  for (int k = 0; k < K; ++k) ops[k] = RED_SUM;
  if (grid_reduce<K>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) out[k] = v[k];
    }
  }
}

// inv = G^-1 for a symmetric positive definite k x k matrix (k <= 6): Gauss-Jordan by one thread
__global__ void k_small_spd_inverse(int k, const double* __restrict__ G, double* __restrict__ inv) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a[6][12];
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) { a[i][j] = 0.5 * (G[i * k + j] + G[j * k + i]); a[i][k + j] = (i == j) ? 1.0 : 0.0; }
  for (int p = 0; p < k; ++p) {
    const double piv = 1.0 / a[p][p];
    for (int j = 0; j < 2 * k; ++j) a[p][j] *= piv;
    for (int r = 0; r < k; ++r) {
      if (r == p) continue;
      const double f = a[r][p];
      for (int j = 0; j < 2 * k; ++j) a[r][j] -= f * a[p][j];
    }
  }
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) inv[i * k + j] = a[i][k + j];
}

}  // namespace cb

using namespace cb;

extern "C" int cb_rbm_galerkin(void* stream, const cb_mat* A, const double* coords, const double* centre,
                               double* c_out, double* v_work, double* w_work, double* G_dev, double* scratch,
                               int64_t scratch_len) {
  CB_REQUIRE(A && coords && centre && c_out && v_work && w_work && G_dev, "cb_rbm_galerkin: missing argument");
  CB_REQUIRE(A->bs == 2 || A->bs == 3, "cb_rbm_galerkin: block size must be 2 or 3");
  CB_REQUIRE(scratch && scratch_len >= RED_SCRATCH_LEN, "cb_rbm_galerkin: scratch too small");
  cudaStream_t s = as_stream(stream);
  const int bs = A->bs, K = bs == 3 ? 6 : 3;
  const int64_t ncols = A->ncols, nrows = A->nrows;
  const int gv = grid_for(ncols * bs, 256, 8);
  int gr = grid_for(nrows, 256, 4);
  if (gr > RED_MAXP) gr = RED_MAXP;
  if (bs == 3) k_rbm_centre<3><<<gv, 256, 0, s>>>(ncols, coords, centre[0], centre[1], centre[2], c_out);
  else k_rbm_centre<2><<<gv, 256, 0, s>>>(ncols, coords, centre[0], centre[1], 0.0, c_out);
  CB_LAUNCH_CHECK();
  for (int j = 0; j < K; ++j) {
    if (bs == 3) k_rbm_mode_vector<3><<<gv, 256, 0, s>>>(ncols, j, c_out, v_work);
    else k_rbm_mode_vector<2><<<gv, 256, 0, s>>>(ncols, j, c_out, v_work);
    CB_LAUNCH_CHECK();
    if (int e = spmv_dispatch(s, A, v_work, w_work, nullptr, nullptr, nullptr)) return e;
    // column j of G = Z^T (A z_j)
    if (bs == 3) k_rbm_project_plain<3><<<gr, 256, 0, s>>>(nrows, w_work, c_out, G_dev + (int64_t)j * K, scratch);
    else k_rbm_project_plain<2><<<gr, 256, 0, s>>>(nrows, w_work, c_out, G_dev + (int64_t)j * K, scratch);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

extern "C" int cb_rbm_invert(void* stream, int k, const double* G_dev, double* einv_dev) {
  CB_REQUIRE(G_dev && einv_dev && k >= 1 && k <= 6, "cb_rbm_invert: k must be in 1..6");
  k_small_spd_inverse<<<1, 32, 0, as_stream(stream)>>>(k, G_dev, einv_dev);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
