// p-Laplace projection of the shape derivative (_PLaplaceProjector, cashocs/_pde_problems/shape_gradient_problem.py:329-427):
//   F(U)[W] = int mu (eps + |grad U|^(p-2)) grad U : grad W dx + delta int U . W dx - dJ[W] = 0
// for vector-P1 U, W, solved by Newton's method for p = 2, 3, ..., p_target (each p starts from the previous solution).
// Two vertex-centric kernels, both walking the cells around a vertex in ascending order (no atomics, fixed summation
// order -> bit-reproducible): the residual F(U) and the Jacobian
//   J(U)[dU, W] = int mu [(eps + s^((p-2)/2)) grad dU : grad W + (p - 2) s^((p-4)/2) (grad U : grad dU)(grad U : grad W)] dx
//                 + delta int dU . W dx,   s = grad U : grad U,
// as a d x d block CSR matrix on the mesh's P1 pattern.  grad U is constant per cell and mu is P1, so
// int_K mu (...) dx = |K| mean(mu) (...) exactly -- no quadrature.  Dirichlet dofs are eliminated symmetrically with
// dolfin's SystemAssembler convention (an eliminated diagonal entry counts the cells at the vertex).
// Not on the hot path of bench.py: one thread per row is simple rather than fast.
#include "common.cuh"

namespace cb {

template <int D>
struct PLapCell {
  int v[D + 1];
  Geo<D> g;
  double H[D][D];   // H[i][j] = d_j U_i
  double s;         // grad U : grad U
  double mubar;
};

template <int D>
__device__ __forceinline__ void plap_load(const int32_t* __restrict__ cells, const double* __restrict__ coords,
                                          const double* __restrict__ U, const double* __restrict__ mu, int64_t c,
                                          PLapCell<D>& pc) {
  constexpr int K = D + 1;
  double x[K][D];
  load_cell<D>(cells, c, pc.v);
  load_coords<D>(coords, pc.v, x);
  simplex_geometry(x, pc.g);
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) pc.H[i][j] = 0.0;
  double m = 0.0;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    m += __ldg(mu + pc.v[a]);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double ua = __ldg(U + (int64_t)pc.v[a] * D + i);
#pragma unroll
      for (int j = 0; j < D; ++j) pc.H[i][j] += ua * pc.g.G[a][j];
    }
  }
  pc.mubar = m * (1.0 / K);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < D; ++j) s += pc.H[i][j] * pc.H[i][j];
  pc.s = s;
}

// kappa = s^((p-2)/2); the reference's UFL simplifies the exponent 0 to the constant 1
__device__ __forceinline__ double plap_kappa(double s, int p) { return p == 2 ? 1.0 : pow(s, 0.5 * (p - 2)); }

template <int D>
__global__ void __launch_bounds__(128)
k_plaplace_residual(int64_t nrows, const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                    const int32_t* __restrict__ cells, const double* __restrict__ coords, const double* __restrict__ U,
                    const double* __restrict__ mu, int p, double eps, double delta, const double* __restrict__ shift,
                    const uint8_t* __restrict__ bc_flag, double* __restrict__ out) {
  constexpr int K = D + 1;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) acc[i] = 0.0;
  for (int64_t k = v2c_ptr[r]; k < v2c_ptr[r + 1]; ++k) {
    PLapCell<D> pc;
    plap_load<D>(cells, coords, U, mu, v2c_idx[k], pc);
    int a = 0;
#pragma unroll
    for (int j = 1; j < K; ++j) a = (pc.v[j] == (int)r) ? j : a;
    const double w = pc.g.vol * pc.mubar * (eps + plap_kappa(pc.s, p));
    double usum[D];
#pragma unroll
    for (int i = 0; i < D; ++i) usum[i] = 0.0;
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
      for (int i = 0; i < D; ++i) usum[i] += __ldg(U + (int64_t)pc.v[b] * D + i);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double hg = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) hg += pc.H[i][j] * pc.g.G[a][j];
      const double ua = __ldg(U + r * D + i);
      acc[i] += w * hg + delta * pc.g.vol * MREF * (usum[i] + ua);
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const bool isbc = bc_flag && bc_flag[r * D + i];
    // Dirichlet rows of both assembled vectors hold u - g; their difference vanishes
    out[r * D + i] = isbc ? 0.0 : acc[i] - (shift ? shift[r * D + i] : 0.0);
  }
}

template <int D>
__global__ void __launch_bounds__(128)
k_plaplace_jacobian(int64_t nrows, const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                    const int32_t* __restrict__ cells, const double* __restrict__ coords, const double* __restrict__ U,
                    const double* __restrict__ mu, int p, double eps, double delta, const uint8_t* __restrict__ bc_flag,
                    const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, double* __restrict__ val) {
  constexpr int K = D + 1;
  constexpr int DD = D * D;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t e0 = rowptr[r], e1 = rowptr[r + 1];
  for (int64_t e = e0 * DD; e < e1 * DD; ++e) val[e] = 0.0;
  bool rbc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) rbc[i] = bc_flag && bc_flag[r * D + i];
  for (int64_t k = v2c_ptr[r]; k < v2c_ptr[r + 1]; ++k) {
    PLapCell<D> pc;
    plap_load<D>(cells, coords, U, mu, v2c_idx[k], pc);
    int a = 0;
#pragma unroll
    for (int j = 1; j < K; ++j) a = (pc.v[j] == (int)r) ? j : a;
    const double w0 = pc.g.vol * pc.mubar;
    const double w1 = w0 * (eps + plap_kappa(pc.s, p));
    // (p - 2) s^((p-4)/2); a cell with grad U = 0 contributes nothing (the reference would produce 0 * inf there)
    const double w2 = (p == 2 || pc.s == 0.0) ? 0.0 : w0 * (p - 2) * pow(pc.s, 0.5 * (p - 4));
    double hga[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) t += pc.H[i][j] * pc.g.G[a][j];
      hga[i] = t;
    }
#pragma unroll
    for (int b = 0; b < K; ++b) {
      const int32_t cb = pc.v[b];
      int64_t e = e0;
      while (e < e1 && col[e] != cb) ++e;
      if (e == e1) continue;  // cannot happen on the full cell pattern
      double gg = 0.0, hgb[D];
#pragma unroll
      for (int j = 0; j < D; ++j) gg += pc.g.G[a][j] * pc.g.G[b][j];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) t += pc.H[i][j] * pc.g.G[b][j];
        hgb[i] = t;
      }
      const double diag = w1 * gg + delta * pc.g.vol * MREF * (a == b ? 2.0 : 1.0);
      double* blk = val + e * DD;
#pragma unroll
      for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double t = w2 * hga[i] * hgb[j] + (i == j ? diag : 0.0);
          const bool cbc = bc_flag && bc_flag[(int64_t)cb * D + j];
          if (rbc[i] || cbc) t = (a == b && i == j && rbc[i]) ? 1.0 : 0.0;
          blk[i * D + j] += t;
        }
      }
    }
  }
}

// int mu (eps + |grad a|^(p-2)) grad a : grad b + delta a . b dx -- the "scalar product" of the reference when the
// p-Laplace projection is on (ShapeFormHandler.scalar_product, cashocs/_forms/shape_form_handler.py:761-770: the
// p-Laplace form with the gradient replaced by a and the test function by b; nonlinear in a).
template <int D>
__global__ void __launch_bounds__(256)
k_plaplace_product(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
                   const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ mu, int p,
                   double eps, double delta, double* __restrict__ result, double* __restrict__ scratch) {
  constexpr int K = D + 1;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  __shared__ double s_red[33];
  double local = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    PLapCell<D> pc;
    plap_load<D>(cells, coords, a, mu, c, pc);
    double hb[D][D], asum[D], bsum[D], ab = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      asum[i] = 0.0; bsum[i] = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) hb[i][j] = 0.0;
    }
#pragma unroll
    for (int v = 0; v < K; ++v) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double av = __ldg(a + (int64_t)pc.v[v] * D + i), bv = __ldg(b + (int64_t)pc.v[v] * D + i);
        asum[i] += av; bsum[i] += bv; ab += av * bv;
#pragma unroll
        for (int j = 0; j < D; ++j) hb[i][j] += bv * pc.g.G[v][j];
      }
    }
    double hh = 0.0, ss = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      ss += asum[i] * bsum[i];
#pragma unroll
      for (int j = 0; j < D; ++j) hh += pc.H[i][j] * hb[i][j];
    }
    local += pc.g.vol * pc.mubar * (eps + plap_kappa(pc.s, p)) * hh + delta * pc.g.vol * MREF * (ss + ab);
  }
  double v1[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v1, ops, scratch, s_red)) {
    if (threadIdx.x == 0) *result = v1[0];
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_plaplace_residual(void* stream, int dim, int64_t nrows, const int64_t* v2c_ptr, const int32_t* v2c_idx,
                                    const int32_t* cells, const double* coords, const double* U, const double* mu, int p,
                                    double eps, double delta, const double* shift, const uint8_t* bc_flag, double* out) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_plaplace_residual: dim must be 2 or 3");
  CB_REQUIRE(p >= 2, "cb_plaplace_residual: p must be at least 2");
  CB_REQUIRE(v2c_ptr && v2c_idx && cells && coords && U && mu && out, "cb_plaplace_residual: missing argument");
  if (nrows == 0) return CB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = (int)((nrows + 127) / 128);
  if (dim == 2)
    k_plaplace_residual<2><<<grid, 128, 0, s>>>(nrows, v2c_ptr, v2c_idx, cells, coords, U, mu, p, eps, delta, shift, bc_flag, out);
  else
    k_plaplace_residual<3><<<grid, 128, 0, s>>>(nrows, v2c_ptr, v2c_idx, cells, coords, U, mu, p, eps, delta, shift, bc_flag, out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_plaplace_jacobian(void* stream, int dim, int64_t nrows, const int64_t* v2c_ptr, const int32_t* v2c_idx,
                                    const int32_t* cells, const double* coords, const double* U, const double* mu, int p,
                                    double eps, double delta, const uint8_t* bc_flag, const int64_t* rowptr,
                                    const int32_t* col, double* val) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_plaplace_jacobian: dim must be 2 or 3");
  CB_REQUIRE(p >= 2, "cb_plaplace_jacobian: p must be at least 2");
  CB_REQUIRE(v2c_ptr && v2c_idx && cells && coords && U && mu && rowptr && col && val, "cb_plaplace_jacobian: missing argument");
  if (nrows == 0) return CB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = (int)((nrows + 127) / 128);
  if (dim == 2)
    k_plaplace_jacobian<2><<<grid, 128, 0, s>>>(nrows, v2c_ptr, v2c_idx, cells, coords, U, mu, p, eps, delta, bc_flag, rowptr, col, val);
  else
    k_plaplace_jacobian<3><<<grid, 128, 0, s>>>(nrows, v2c_ptr, v2c_idx, cells, coords, U, mu, p, eps, delta, bc_flag, rowptr, col, val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_plaplace_product(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                                   const double* a, const double* b, const double* mu, int p, double eps, double delta,
                                   double* result_dev, double* scratch, int64_t scratch_len) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_plaplace_product: dim must be 2 or 3");
  CB_REQUIRE(p >= 2, "cb_plaplace_product: p must be at least 2");
  CB_REQUIRE(coords && cells && a && b && mu && result_dev && scratch && scratch_len >= RED_SCRATCH_LEN,
             "cb_plaplace_product: missing argument");
  cudaStream_t s = as_stream(stream);
  int grid = grid_for(nc, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  if (dim == 2) k_plaplace_product<2><<<grid, 256, 0, s>>>(nc, coords, cells, a, b, mu, p, eps, delta, result_dev, scratch);
  else k_plaplace_product<3><<<grid, 256, 0, s>>>(nc, coords, cells, a, b, mu, p, eps, delta, result_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
