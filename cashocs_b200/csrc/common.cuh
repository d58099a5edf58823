// Shared device helpers: error handling, deterministic grid reductions, small linear algebra.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "cashocs_b200.h"

namespace cb {

void set_error(const char* fmt, ...);

#define CB_CUDA_CHECK(expr)                                                             \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      cb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CB_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// kernels launched by this library (bench.py's gpu_launches); atomic so that callers may drive the library from several
// host threads, the per-thread twin serves the graph-capture bookkeeping of krylov.cu
extern std::atomic<long long> g_launch_count;
extern thread_local long long t_launch_count;
#define CB_LAUNCH_CHECK()                                            \
  do {                                                               \
    ++cb::t_launch_count;                                            \
    cb::g_launch_count.fetch_add(1, std::memory_order_relaxed);      \
    CB_CUDA_CHECK(cudaGetLastError());                               \
  } while (0)

#define CB_REQUIRE(cond, msg)                                \
  do {                                                       \
    if (!(cond)) {                                           \
      cb::set_error("%s:%d: %s", __FILE__, __LINE__, msg);   \
      return CB_ERR_ARG;                                     \
    }                                                        \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();
int configure_scratch_pool();

// Grid size for grid-stride kernels: a multiple of the SM count, capped by the work.
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t need = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------ deterministic reductions
// Scratch layout (doubles): [0, MAXP*NVMAX) block partials, then one 8-byte slot holding the
// arrival counter.  The grid never exceeds MAXP blocks for reducing kernels.
constexpr int RED_MAXP = 2048;
constexpr int RED_NVMAX = 8;
constexpr int64_t RED_SCRATCH_LEN = (int64_t)RED_MAXP * RED_NVMAX + 8;

enum RedOp { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2 };

__device__ __forceinline__ double red_combine(double a, double b, int op) {
  if (op == RED_SUM) return a + b;
  // NaN-propagating min/max so that a NaN determinant / quality is never silently dropped
  if (a != a) return a;
  if (b != b) return b;
  if (op == RED_MIN) return a < b ? a : b;
  return a > b ? a : b;
}

__device__ __forceinline__ double red_identity(int op) {
  if (op == RED_SUM) return 0.0;
  if (op == RED_MIN) return __longlong_as_double(0x7ff0000000000000LL);
  return __longlong_as_double(0xfff0000000000000LL);
}

// Fixed-order block tree reduction of NV values; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], const int (&ops)[NV], double* smem /* >= 32*NV */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double o = __shfl_down_sync(0xffffffffu, v[i], off);
      v[i] = red_combine(v[i], o, ops[i]);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[warp * NV + i] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < nwarps ? smem[lane * NV + i] : red_identity(ops[i]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        double o = __shfl_down_sync(0xffffffffu, x, off);
        x = red_combine(x, o, ops[i]);
      }
      v[i] = x;
    }
  }
  __syncthreads();
}

// Grid-wide deterministic reduction.  Every block contributes v; the last block to arrive
// re-reduces all block partials in a fixed order.  Returns true in ALL threads of that last
// block, with the final values valid in thread 0's v.  The counter resets itself.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], const int (&ops)[NV], double* scratch,
                                            double* smem /* >= 32*NV+1 */) {
  block_reduce<NV>(v, ops, smem);
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (int64_t)RED_MAXP * RED_NVMAX);
  __shared__ unsigned int s_ticket;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) __stcg(&scratch[(int64_t)blockIdx.x * NV + i], v[i]);
    __threadfence();
    s_ticket = atomicInc(counter, gridDim.x - 1);
  }
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return false;
  __threadfence();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = red_identity(ops[i]);
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = red_combine(v[i], __ldcg(&scratch[(int64_t)b * NV + i]), ops[i]);
  }
  block_reduce<NV>(v, ops, smem);
  return true;
}

// ------------------------------------------------------------------ simplex geometry
template <int D>
struct Geo {
  double G[D + 1][D];  // gradients of the barycentric coordinates
  double detJ;
  double vol;
};

// x[a][i]: vertex a, component i.
__device__ __forceinline__ void simplex_geometry(const double (&x)[3][2], Geo<2>& g) {
  const double a = x[1][0] - x[0][0], b = x[2][0] - x[0][0];
  const double c = x[1][1] - x[0][1], d = x[2][1] - x[0][1];
  const double det = a * d - b * c;
  const double inv = 1.0 / det;
  // Jinv = 1/det [[d, -b], [-c, a]] ; rows = grad lambda_1, grad lambda_2
  g.G[1][0] = d * inv;
  g.G[1][1] = -b * inv;
  g.G[2][0] = -c * inv;
  g.G[2][1] = a * inv;
  g.G[0][0] = -(g.G[1][0] + g.G[2][0]);
  g.G[0][1] = -(g.G[1][1] + g.G[2][1]);
  g.detJ = det;
  g.vol = fabs(det) * 0.5;
}

__device__ __forceinline__ void simplex_geometry(const double (&x)[4][3], Geo<3>& g) {
  double J[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int a = 0; a < 3; ++a) J[i][a] = x[a + 1][i] - x[0][i];
  }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  const double inv = 1.0 / det;
  // Jinv[a][i] = cofactor(J)[i][a] / det ; rows a = grad lambda_{a+1}
  g.G[1][0] = c00 * inv;
  g.G[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv;
  g.G[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
  g.G[2][0] = c01 * inv;
  g.G[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv;
  g.G[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv;
  g.G[3][0] = c02 * inv;
  g.G[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv;
  g.G[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
#pragma unroll
  for (int i = 0; i < 3; ++i) g.G[0][i] = -(g.G[1][i] + g.G[2][i] + g.G[3][i]);
  g.detJ = det;
  g.vol = fabs(det) * (1.0 / 6.0);
}

template <int D>
__device__ __forceinline__ void load_cell(const int32_t* __restrict__ cells, int64_t c, int (&v)[D + 1]) {
  if constexpr (D == 3) {
    const int4 t = __ldg(reinterpret_cast<const int4*>(cells) + c);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int a = 0; a < D + 1; ++a) v[a] = __ldg(cells + c * (D + 1) + a);
  }
}

template <int D>
__device__ __forceinline__ void load_coords(const double* __restrict__ coords, const int (&v)[D + 1],
                                            double (&x)[D + 1][D]) {
#pragma unroll
  for (int a = 0; a < D + 1; ++a) {
    if constexpr (D == 2) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(coords) + v[a]);
      x[a][0] = t.x; x[a][1] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) x[a][i] = __ldg(coords + (int64_t)v[a] * D + i);
    }
  }
}

__host__ __device__ __forceinline__ double poly_eval(const cb_poly& p, double x, double y, double z) {
  double s = 0.0;
  for (int t = 0; t < p.nterms; ++t) {
    double m = p.coef[t];
    for (int e = 0; e < p.ex[t]; ++e) m *= x;
    for (int e = 0; e < p.ey[t]; ++e) m *= y;
    for (int e = 0; e < p.ez[t]; ++e) m *= z;
    s += m;
  }
  return s;
}

// Dense coefficient table of a polynomial of total degree <= 4 (35 monomials): what the quadrature loops of the
// element kernels evaluate when the coefficient fits (the source terms of the reference's demos and tests are of
// degree <= 4).  The table lives in kernel-parameter space, the nested Horner scheme below is fully unrolled, so
// every coefficient is an immediate constant-bank operand of one DFMA -- no term loop, no exponent loops, no
// integer work (the generic monomial evaluator above costs ~10 instructions per multiply).  Rows x^0..x^n of one
// (j, k) = (power of y, power of z) are contiguous.  The kernels are instantiated for degree <= 2 / <= 4 and for
// z-independent sources (15 instead of 35 coefficients), chosen on the host from the table.
constexpr int DP_DEG = 4;
constexpr int DP_N = 35;
__host__ __device__ constexpr int dp_offset(int k, int j) {
  int off = 0;
  for (int kk = 0; kk < k; ++kk)
    for (int jj = 0; jj <= DP_DEG - kk; ++jj) off += DP_DEG - kk - jj + 1;
  for (int jj = 0; jj < j; ++jj) off += DP_DEG - k - jj + 1;
  return off;
}
struct DensePoly {
  double c[DP_N];
  uint32_t rows;    // bit k * 5 + j: row (j, k) holds a nonzero coefficient
  uint32_t planes;  // bit k: some row of the plane z^k does
};
// false when the polynomial has a monomial of total degree > DP_DEG (callers then keep the generic evaluator)
inline bool dense_from_poly(const cb_poly& p, DensePoly& d) {
  memset(&d, 0, sizeof(d));
  for (int t = 0; t < p.nterms; ++t) {
    const int i = p.ex[t], j = p.ey[t], k = p.ez[t];
    if (i < 0 || j < 0 || k < 0 || i + j + k > DP_DEG) return false;
    if (p.coef[t] == 0.0) continue;
    d.c[dp_offset(k, j) + i] += p.coef[t];
    d.rows |= 1u << (k * 5 + j);
    d.planes |= 1u << k;
  }
  return true;
}
// Horner scheme over the table, specialised at compile time on the degree (DEG <= DP_DEG: rows beyond it are not
// touched) and on whether the polynomial depends on z: straight-line DFMAs with constant-bank operands.
template <int D, int DEG, bool ZDEP>
__host__ __device__ __forceinline__ double dense_eval(const DensePoly& p, double x, double y, double z) {
  double sz = 0.0;
#pragma unroll
  for (int k = ((D == 3 && ZDEP) ? DEG : 0); k >= 0; --k) {
    double sy = 0.0;
#pragma unroll
    for (int j = DEG - k; j >= 0; --j) {
      const int o = dp_offset(k, j), n = DEG - k - j;
      double sx = p.c[o + n];
#pragma unroll
      for (int i = n - 1; i >= 0; --i) sx = fma(sx, x, p.c[o + i]);
      sy = (j == DEG - k) ? sx : fma(sy, y, sx);
    }
    sz = (k == ((D == 3 && ZDEP) ? DEG : 0)) ? sy : fma(sz, z, sy);
  }
  return sz;
}
// total degree and z-dependence of a table (host side: picks the kernel instantiation)
inline int dense_degree(const DensePoly& p, bool* zdep) {
  int deg = 0;
  for (int k = 0; k <= DP_DEG; ++k)
    for (int j = 0; j <= DP_DEG - k; ++j)
      for (int i = 0; i <= DP_DEG - k - j; ++i)
        if (p.c[dp_offset(k, j) + i] != 0.0) {
          if (i + j + k > deg) deg = i + j + k;
          if (k > 0 && zdep) *zdep = true;
        }
  return deg;
}
// one name for both representations, so the element kernels are templated on the coefficient type
template <int D, int DEG, bool ZDEP>
__device__ __forceinline__ double coef_eval(const cb_poly& p, double x, double y, double z) { return poly_eval(p, x, y, z); }
template <int D, int DEG, bool ZDEP>
__device__ __forceinline__ double coef_eval(const DensePoly& p, double x, double y, double z) {
  return dense_eval<D, DEG, ZDEP>(p, x, y, z);
}
__device__ __forceinline__ bool coef_nonzero(const cb_poly& p) { return p.nterms > 0; }
__device__ __forceinline__ bool coef_nonzero(const DensePoly& p) { return p.planes != 0u; }

// out[v, :] = sum over the cells c of vertex v (ascending c) of cell_vec[c, local index of v, :]
template <int D, int NC>
__global__ void __launch_bounds__(256)
k_gather_cell_vectors(int64_t nrows, const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                      const int32_t* __restrict__ cells, const double* __restrict__ cell_vec,
                      double* __restrict__ out) {
  constexpr int K = D + 1;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nrows;
       v += (int64_t)gridDim.x * blockDim.x) {
    double acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = 0.0;
    for (int64_t k = v2c_ptr[v]; k < v2c_ptr[v + 1]; ++k) {
      const int64_t c = v2c_idx[k];
      int cv[K];
      load_cell<D>(cells, c, cv);
      int a = 0;
#pragma unroll
      for (int j = 1; j < K; ++j) a = (cv[j] == (int)v) ? j : a;
      const double* src = cell_vec + (c * K + a) * NC;
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[i] += src[i];
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) out[v * NC + i] = acc[i];
  }
}

}  // namespace cb
