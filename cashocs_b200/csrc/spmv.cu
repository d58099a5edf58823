// SpMV, scalar products and BLAS-1 reductions (PETSc Mat.mult / Vec.dot call sites:
// cashocs/_forms/shape_form_handler.py:776-778, cashocs/_forms/control_form_handler.py:184-207).
#include "spmv.cuh"

namespace cb {

int spmv_dispatch(cudaStream_t s, const cb_mat* A, const double* x, double* y, const double* w,
                  double* result, double* scratch, const int* done, bool use32 = false);

template <int BS, bool WRITE_Y, bool DOT, typename VT>
__global__ void __launch_bounds__(SpmvCfg<BS, VT>::THREADS)
k_spmv(int64_t nrows, int G, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
       const VT* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
       const double* __restrict__ w, double* __restrict__ result, double* __restrict__ scratch,
       const int* __restrict__ done, const double* __restrict__ dscale, double alpha,
       const double* __restrict__ bsub) {
  if (done && *done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[33];
  const SpmvSmem<BS, VT> sm = spmv_smem_init<BS, VT>(smem_raw);
  uint32_t phase = 0;
  const int R = spmv_rows_per_cta<BS>(G);
  const int64_t nrb = (nrows + R - 1) / R;
  double local = 0.0;
  for (int64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
    const int64_t row0 = rb * R;
    const int nr = (int)((nrows - row0 < R) ? (nrows - row0) : R);
    double acc;
    const int64_t i = spmv_row_block<BS, VT>(sm, rowptr, col, val, x, row0, nr, G, phase, acc);
    if (i >= 0) {
      if (dscale) acc *= alpha * dscale[i];      // power method: y = alpha D^-1 A x ...
      if (bsub) acc = bsub[i] - acc;             // residual pass of the V-cycle: y = b - A x
      if (WRITE_Y) y[i] = acc;
      if (DOT) local += acc * (w ? w[i] : acc);  // ... with |y|^2 (w == NULL)
    }
  }
  if (DOT) {
    double v[1] = {local};
    const int ops[1] = {RED_SUM};
    if (grid_reduce<1>(v, ops, scratch, s_red)) {
      if (threadIdx.x == 0) *result = v[0];
    }
  }
}

__global__ void __launch_bounds__(256)
k_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
      double* __restrict__ result, double* __restrict__ scratch) {
  __shared__ double s_red[33];
  double local = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    local += x[i] * y[i];
  double v[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) *result = v[0];
  }
}

__global__ void __launch_bounds__(256)
k_diag_inv(int64_t nrows, int bs, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
           const double* __restrict__ val, double* __restrict__ dinv, double* __restrict__ lmax,
           double* __restrict__ scratch) {
  __shared__ double s_red[33];
  const int64_t total = nrows * bs;
  double local = 0.0;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t br = r / bs;
    const int i = (int)(r - br * bs);
    double d = 0.0, s = 0.0;
    for (int64_t k = rowptr[br]; k < rowptr[br + 1]; ++k) {
      for (int j = 0; j < bs; ++j) s += fabs(val[(k * bs + i) * bs + j]);
      if (col[k] == br) d = val[(k * bs + i) * bs + i];
    }
    dinv[r] = 1.0 / d;
    const double ratio = s / fabs(d);
    local = ratio > local ? ratio : local;
  }
  if (lmax) {
    double v[1] = {local};
    const int ops[1] = {RED_MAX};
    if (grid_reduce<1>(v, ops, scratch, s_red)) {
      if (threadIdx.x == 0) *lmax = v[0];
    }
  }
}

template <int BS, typename VT>
int spmv_grid(int64_t nrows, int G) {
  using C = SpmvCfg<BS, VT>;
  const int R = spmv_rows_per_cta<BS>(G);
  const int64_t nrb = (nrows + R - 1) / R;
  int64_t cap = (int64_t)num_sms() * C::CTAS_PER_SM;
  if (cap > RED_MAXP) cap = RED_MAXP;
  return (int)(nrb < cap ? nrb : cap);
}

template <typename K>
int prefer_smem(K kernel) {
  CB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
  return CB_OK;
}

template <int BS, typename VT>
int launch_spmv(cudaStream_t s, const cb_mat* A, const VT* val, const double* x, double* y, const double* w,
                double* result, double* scratch, const int* done, const double* dscale = nullptr, double alpha = 1.0,
                const double* bsub = nullptr) {
  using C = SpmvCfg<BS, VT>;
  const int G = spmv_sanitize_split(A->row_split);
  const int grid = spmv_grid<BS, VT>(A->nrows, G);
  static bool configured = false;
  if (!configured) {
    if (int e = prefer_smem(k_spmv<BS, true, false, VT>)) return e;
    if (int e = prefer_smem(k_spmv<BS, true, true, VT>)) return e;
    if (int e = prefer_smem(k_spmv<BS, false, true, VT>)) return e;
    configured = true;
  }
  if (y && !w && !result)
    k_spmv<BS, true, false, VT><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(A->nrows, G, A->rowptr, A->col, val, x, y, nullptr, nullptr, nullptr, done, dscale, alpha, bsub);
  else if (y)
    k_spmv<BS, true, true, VT><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(A->nrows, G, A->rowptr, A->col, val, x, y, w, result, scratch, done, dscale, alpha, bsub);
  else
    k_spmv<BS, false, true, VT><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(A->nrows, G, A->rowptr, A->col, val, x, nullptr, w, result, scratch, done, dscale, alpha, nullptr);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int diag_inv_launch(cudaStream_t s, const cb_mat* A, double* dinv, double* lmax, double* scratch) {
  int grid = grid_for(A->nrows * A->bs, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  k_diag_inv<<<grid, 256, 0, s>>>(A->nrows, A->bs, A->rowptr, A->col, A->val, dinv, lmax, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// use32: read the fp32 copy of the values when the matrix carries one (preconditioner passes)
int spmv_dispatch(cudaStream_t s, const cb_mat* A, const double* x, double* y, const double* w,
                  double* result, double* scratch, const int* done, bool use32) {
  CB_REQUIRE(A && A->rowptr && A->col && A->val && x, "cb_spmv: missing argument");
  if (use32 && A->val32) {
    switch (A->bs) {
      case 1: return launch_spmv<1, float>(s, A, A->val32, x, y, w, result, scratch, done);
      case 2: return launch_spmv<2, float>(s, A, A->val32, x, y, w, result, scratch, done);
      case 3: return launch_spmv<3, float>(s, A, A->val32, x, y, w, result, scratch, done);
    }
  }
  switch (A->bs) {
    case 1: return launch_spmv<1, double>(s, A, A->val, x, y, w, result, scratch, done);
    case 2: return launch_spmv<2, double>(s, A, A->val, x, y, w, result, scratch, done);
    case 3: return launch_spmv<3, double>(s, A, A->val, x, y, w, result, scratch, done);
  }
  set_error("cb_spmv: unsupported block size %d", A->bs);
  return CB_ERR_UNSUPPORTED;
}

// y = b - A x in one pass (the residual the V-cycle restricts); use32 as above
int spmv_residual_dispatch(cudaStream_t s, const cb_mat* A, const double* x, const double* b, double* y, const int* done,
                           bool use32) {
  CB_REQUIRE(A && A->rowptr && A->col && A->val && x && b && y, "spmv_residual: missing argument");
  if (use32 && A->val32) {
    switch (A->bs) {
      case 1: return launch_spmv<1, float>(s, A, A->val32, x, y, nullptr, nullptr, nullptr, done, nullptr, 1.0, b);
      case 2: return launch_spmv<2, float>(s, A, A->val32, x, y, nullptr, nullptr, nullptr, done, nullptr, 1.0, b);
      case 3: return launch_spmv<3, float>(s, A, A->val32, x, y, nullptr, nullptr, nullptr, done, nullptr, 1.0, b);
    }
  }
  switch (A->bs) {
    case 1: return launch_spmv<1, double>(s, A, A->val, x, y, nullptr, nullptr, nullptr, done, nullptr, 1.0, b);
    case 2: return launch_spmv<2, double>(s, A, A->val, x, y, nullptr, nullptr, nullptr, done, nullptr, 1.0, b);
    case 3: return launch_spmv<3, double>(s, A, A->val, x, y, nullptr, nullptr, nullptr, done, nullptr, 1.0, b);
  }
  set_error("spmv_residual: unsupported block size %d", A->bs);
  return CB_ERR_UNSUPPORTED;
}

// dst = (float) src
__global__ void __launch_bounds__(256)
k_to_f32(int64_t n, const double* __restrict__ src, float* __restrict__ dst) {
  const int64_t n2 = n >> 1;
  const double2* s2 = reinterpret_cast<const double2*>(src);
  float2* d2 = reinterpret_cast<float2*>(dst);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 v = s2[i];
    d2[i] = make_float2((float)v.x, (float)v.y);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = (float)src[n - 1];
}

int to_f32_launch(cudaStream_t s, int64_t n, const double* src, float* dst) {
  k_to_f32<<<grid_for(n / 2 + 1, 256, 8), 256, 0, s>>>(n, src, dst);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

}  // namespace cb

using namespace cb;

extern "C" int64_t cb_reduce_scratch_len(void) { return RED_SCRATCH_LEN; }

extern "C" int cb_spmv(void* stream, const cb_mat* A, const double* x, double* y) {
  CB_REQUIRE(y != nullptr, "cb_spmv: y is NULL");
  return spmv_dispatch(as_stream(stream), A, x, y, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int cb_spmv_f32(void* stream, const cb_mat* A, const double* x, double* y) {
  CB_REQUIRE(y != nullptr && A && A->val32, "cb_spmv_f32: y or A->val32 is NULL");
  return spmv_dispatch(as_stream(stream), A, x, y, nullptr, nullptr, nullptr, nullptr, /*use32=*/true);
}

// One step of the power method for lambda_max(D^-1 A): y = alpha * dinv .* (A x), *norm2_dev = |y|^2 (this rank's rows).
extern "C" int cb_power_step(void* stream, const cb_mat* A, const double* dinv, double alpha, const double* x, double* y,
                             double* norm2_dev, double* scratch, int64_t scratch_len, int use32) {
  CB_REQUIRE(A && dinv && x && y && norm2_dev, "cb_power_step: missing argument");
  CB_REQUIRE(scratch && scratch_len >= RED_SCRATCH_LEN, "cb_power_step: scratch too small");
  cudaStream_t s = as_stream(stream);
  if (use32 && A->val32) {
    switch (A->bs) {
      case 1: return launch_spmv<1, float>(s, A, A->val32, x, y, nullptr, norm2_dev, scratch, nullptr, dinv, alpha);
      case 2: return launch_spmv<2, float>(s, A, A->val32, x, y, nullptr, norm2_dev, scratch, nullptr, dinv, alpha);
      case 3: return launch_spmv<3, float>(s, A, A->val32, x, y, nullptr, norm2_dev, scratch, nullptr, dinv, alpha);
    }
  }
  switch (A->bs) {
    case 1: return launch_spmv<1, double>(s, A, A->val, x, y, nullptr, norm2_dev, scratch, nullptr, dinv, alpha);
    case 2: return launch_spmv<2, double>(s, A, A->val, x, y, nullptr, norm2_dev, scratch, nullptr, dinv, alpha);
    case 3: return launch_spmv<3, double>(s, A, A->val, x, y, nullptr, norm2_dev, scratch, nullptr, dinv, alpha);
  }
  set_error("cb_power_step: unsupported block size %d", A->bs);
  return CB_ERR_UNSUPPORTED;
}

extern "C" int cb_scalar_product(void* stream, const cb_mat* A, const double* x, const double* y,
                                 double* result_dev, double* scratch, int64_t scratch_len) {
  CB_REQUIRE(scratch && scratch_len >= RED_SCRATCH_LEN, "cb_scalar_product: scratch too small");
  CB_REQUIRE(y && result_dev, "cb_scalar_product: missing argument");
  return spmv_dispatch(as_stream(stream), A, x, nullptr, y, result_dev, scratch, nullptr);
}

extern "C" int cb_to_f32(void* stream, int64_t n, const double* src, float* dst) {
  CB_REQUIRE(src && dst && n >= 0, "cb_to_f32: missing argument");
  CB_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
             "cb_to_f32: src must be 16-byte and dst 8-byte aligned");
  return to_f32_launch(as_stream(stream), n, src, dst);
}

extern "C" int cb_dot(void* stream, int64_t n, const double* x, const double* y, double* result_dev,
                      double* scratch, int64_t scratch_len) {
  CB_REQUIRE(scratch && scratch_len >= RED_SCRATCH_LEN, "cb_dot: scratch too small");
  int grid = grid_for(n, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  k_dot<<<grid, 256, 0, as_stream(stream)>>>(n, x, y, result_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_extract_diag_inv(void* stream, const cb_mat* A, double* dinv, double* lmax_dev,
                                   double* scratch, int64_t scratch_len) {
  CB_REQUIRE(A && dinv, "cb_extract_diag_inv: missing argument");
  CB_REQUIRE(!lmax_dev || (scratch && scratch_len >= RED_SCRATCH_LEN), "scratch too small");
  int grid = grid_for(A->nrows * A->bs, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  k_diag_inv<<<grid, 256, 0, as_stream(stream)>>>(A->nrows, A->bs, A->rowptr, A->col, A->val, dinv,
                                                  lmax_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
