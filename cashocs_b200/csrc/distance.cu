// Distance to the boundary by the iterated eikonal linearisation (compute_boundary_distance_eikonal,
// cashocs/geometry/boundary_distance.py:191-282) and the distance-based Lame parameter of Stiffness
// (cashocs/_forms/shape_form_handler.py:141-186,216-235).
//
// Every sweep solves  int grad u . grad v dx = int (grad u_prev / |grad u_prev|) . grad v dx  with u = 0 on the
// chosen boundaries; for P1 functions the right-hand side is piecewise constant per cell, so it is exact without
// quadrature.  The stiffness matrix and the Krylov solve are the state path's (cb_assemble_scalar, cb_ksp_solve);
// the two kernels here are the right-hand side -- vertex-centric: one thread per vertex walks its cells in ascending
// order and recomputes their geometry (no cell-vector round trip, no atomics, fixed summation order) -- and the
// stopping criterion  int (|grad u| - 1)^2 dx  as a fused grid reduction.  mu(dist) is evaluated per vertex
// (fenics.interpolate of the reference's Expression to CG1).
//
// Compiled with -fmad=false: mu(dist) is a thresholded quantity (dist <= dist_min ? ...), like the mesh tests.
#include "common.cuh"

namespace cb {

template <int D>
__device__ __forceinline__ void cell_gradient(const double* __restrict__ coords, const int32_t* __restrict__ cells,
                                              const double* __restrict__ u, int64_t c, int (&v)[D + 1], Geo<D>& g,
                                              double (&gu)[D]) {
  constexpr int K = D + 1;
  double x[K][D];
  load_cell<D>(cells, c, v);
  load_coords<D>(coords, v, x);
  simplex_geometry(x, g);
#pragma unroll
  for (int i = 0; i < D; ++i) gu[i] = 0.0;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const double ua = __ldg(u + v[a]);
#pragma unroll
    for (int i = 0; i < D; ++i) gu[i] += ua * g.G[a][i];
  }
}

template <int D>
__global__ void __launch_bounds__(128)
k_eikonal_rhs(int64_t nrows, const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
              const int32_t* __restrict__ cells, const double* __restrict__ coords, const double* __restrict__ u,
              const uint8_t* __restrict__ bc_flag, double* __restrict__ out) {
  constexpr int K = D + 1;
  const int64_t vtx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (vtx >= nrows) return;
  if (bc_flag && bc_flag[vtx]) { out[vtx] = 0.0; return; }
  double acc = 0.0;
  for (int64_t k = v2c_ptr[vtx]; k < v2c_ptr[vtx + 1]; ++k) {
    int v[K];
    Geo<D> g;
    double gu[D];
    cell_gradient<D>(coords, cells, u, v2c_idx[k], v, g, gu);
    int a = 0;
#pragma unroll
    for (int j = 1; j < K; ++j) a = (v[j] == (int)vtx) ? j : a;
    double nrm2 = 0.0, dot = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) { nrm2 += gu[i] * gu[i]; dot += gu[i] * g.G[a][i]; }
    acc += g.vol * (dot / sqrt(nrm2));
  }
  out[vtx] = acc;
}

template <int D>
__global__ void __launch_bounds__(256)
k_eikonal_residual(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
                   const double* __restrict__ u, double* __restrict__ result, double* __restrict__ scratch) {
  constexpr int K = D + 1;
  __shared__ double s_red[33];
  double local = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc; c += (int64_t)gridDim.x * blockDim.x) {
    int v[K];
    Geo<D> g;
    double gu[D];
    cell_gradient<D>(coords, cells, u, c, v, g, gu);
    double nrm2 = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) nrm2 += gu[i] * gu[i];
    const double e = sqrt(nrm2) - 1.0;
    local += g.vol * (e * e);
  }
  double v1[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v1, ops, scratch, s_red)) {
    if (threadIdx.x == 0) *result = v1[0];
  }
}

__global__ void __launch_bounds__(256)
k_distance_mu(int64_t n, const double* __restrict__ dist, double dist_min, double dist_max, double mu_min,
              double mu_max, int smooth, double* __restrict__ mu) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double d = dist[i];
    double m;
    if (d <= dist_min) m = mu_min;
    else if (d <= dist_max) {
      const double w = dist_max - dist_min, dm = mu_max - mu_min;
      if (!smooth) m = mu_min + (d - dist_min) / w * dm;
      else
        m = mu_min + dm / w * (d - dist_min) - dm / (w * w) * (d - dist_min) * (d - dist_max)
            - 2.0 * dm / (w * w * w) * (d - dist_min) * ((d - dist_max) * (d - dist_max));
    } else m = mu_max;
    mu[i] = m;
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_eikonal_rhs(void* stream, int dim, int64_t nrows, const int64_t* v2c_ptr, const int32_t* v2c_idx,
                              const int32_t* cells, const double* coords, const double* u, const uint8_t* bc_flag,
                              double* out) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_eikonal_rhs: dim must be 2 or 3");
  CB_REQUIRE(v2c_ptr && v2c_idx && cells && coords && u && out, "cb_eikonal_rhs: missing argument");
  if (nrows == 0) return CB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = (int)((nrows + 127) / 128);
  if (dim == 2) k_eikonal_rhs<2><<<grid, 128, 0, s>>>(nrows, v2c_ptr, v2c_idx, cells, coords, u, bc_flag, out);
  else k_eikonal_rhs<3><<<grid, 128, 0, s>>>(nrows, v2c_ptr, v2c_idx, cells, coords, u, bc_flag, out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_eikonal_residual(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                                   const double* u, double* result_dev, double* scratch, int64_t scratch_len) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_eikonal_residual: dim must be 2 or 3");
  CB_REQUIRE(coords && cells && u && result_dev && scratch && scratch_len >= RED_SCRATCH_LEN,
             "cb_eikonal_residual: missing argument");
  cudaStream_t s = as_stream(stream);
  int grid = grid_for(nc, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  if (dim == 2) k_eikonal_residual<2><<<grid, 256, 0, s>>>(nc, coords, cells, u, result_dev, scratch);
  else k_eikonal_residual<3><<<grid, 256, 0, s>>>(nc, coords, cells, u, result_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_distance_mu(void* stream, int64_t n, const double* dist, double dist_min, double dist_max,
                              double mu_min, double mu_max, int smooth, double* mu) {
  CB_REQUIRE(dist && mu, "cb_distance_mu: missing argument");
  if (n == 0) return CB_OK;
  k_distance_mu<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(n, dist, dist_min, dist_max, mu_min, mu_max,
                                                                    smooth, mu);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
