// Boundary (facet) integrals of the geometric regularisation: the measure of the boundary
// int_Gamma 1 ds (cashocs/_forms/shape_regularization.py:307-316) and its shape derivative
// int_Gamma div_Gamma V ds (:248-269, t_div :63-74).  For the P1 deformation V = phi_a e_i the tangential
// divergence integrated over a flat facet F is exactly d|F| / d x_a[i], so the derivative vector is the
// gradient of the facet measures with respect to the vertex coordinates: one thread per facet writes its
// d local vertex gradients, a per-vertex gather (fixed order, no atomics) sums them.
//
// Curvature regularisation (:512-646) adds two facet kernels: the P1 mass matrix of the boundary, assembled row by
// row on the compressed boundary-vertex numbering (one thread per row walks its facets in ascending order: no atomics,
// fixed summation order), and the shape derivative of 1/2 int |kappa|^2 ds for a given mean-curvature vector kappa.
// On a flat facet t_grad(W, n) of a P1 function is sum_a W_a (x) gamma_a with gamma_a = (d|F|/dx_a) / |F| the surface
// gradient of the facet's hat functions, so both need nothing but the facet's own vertices.
//
// Compiled with -fmad=false like meshops.cu: the measure enters the cost functional and with it the
// Armijo accept / reject decisions.
#include "common.cuh"

namespace cb {

// |F| and d|F|/dx_a for a facet with D vertices in R^D (D = 2: segment, D = 3: triangle)
template <int D>
__device__ __forceinline__ double facet_measure(const double (&x)[D][D], double (&grad)[D][D]) {
  if constexpr (D == 2) {
    const double ex = x[1][0] - x[0][0], ey = x[1][1] - x[0][1];
    const double len = sqrt(ex * ex + ey * ey);
    const double inv = len > 0.0 ? 1.0 / len : 0.0;
    grad[1][0] = ex * inv; grad[1][1] = ey * inv;
    grad[0][0] = -grad[1][0]; grad[0][1] = -grad[1][1];
    return len;
  } else {
    double e1[3], e2[3], n[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { e1[i] = x[1][i] - x[0][i]; e2[i] = x[2][i] - x[0][i]; }
    n[0] = e1[1] * e2[2] - e1[2] * e2[1];
    n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
    const double nh[3] = {n[0] * inv, n[1] * inv, n[2] * inv};
    // d(area)/dx_a = 1/2 (opposite edge, oriented) x n_hat: a = 0 -> x1 - x2, 1 -> x2 - x0, 2 -> x0 - x1
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int b = (a + 1) % 3, c = (a + 2) % 3;
      const double e[3] = {x[b][0] - x[c][0], x[b][1] - x[c][1], x[b][2] - x[c][2]};
      grad[a][0] = 0.5 * (e[1] * nh[2] - e[2] * nh[1]);
      grad[a][1] = 0.5 * (e[2] * nh[0] - e[0] * nh[2]);
      grad[a][2] = 0.5 * (e[0] * nh[1] - e[1] * nh[0]);
    }
    return 0.5 * nn;
  }
}

template <int D>
__device__ __forceinline__ void load_facet(const double* __restrict__ coords, const int32_t* __restrict__ facets,
                                           int64_t f, double (&x)[D][D]) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
    const int64_t v = __ldg(facets + f * D + a);
#pragma unroll
    for (int i = 0; i < D; ++i) x[a][i] = __ldg(coords + v * D + i);
  }
}

template <int D>
__global__ void __launch_bounds__(256)
k_boundary_measure(int64_t nf, const double* __restrict__ coords, const int32_t* __restrict__ facets,
                   double* __restrict__ result, double* __restrict__ scratch) {
  __shared__ double s_red[33];
  double local = 0.0;
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf;
       f += (int64_t)gridDim.x * blockDim.x) {
    double x[D][D], g[D][D];
    load_facet<D>(coords, facets, f, x);
    local += facet_measure<D>(x, g);
  }
  double v1[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v1, ops, scratch, s_red)) {
    if (threadIdx.x == 0) *result = v1[0];
  }
}

// facet_vec[f, a, :] = d|F_f| / d x_a
template <int D>
__global__ void __launch_bounds__(256)
k_boundary_gradient(int64_t nf, const double* __restrict__ coords, const int32_t* __restrict__ facets,
                    double* __restrict__ facet_vec) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf;
       f += (int64_t)gridDim.x * blockDim.x) {
    double x[D][D], g[D][D];
    load_facet<D>(coords, facets, f, x);
    facet_measure<D>(x, g);
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int i = 0; i < D; ++i) facet_vec[(f * D + a) * D + i] = g[a][i];
  }
}

// unit normal of a facet (orientation irrelevant: it only enters through n n^T)
template <int D>
__device__ __forceinline__ void facet_normal(const double (&x)[D][D], double (&n)[D]) {
  if constexpr (D == 2) {
    const double ex = x[1][0] - x[0][0], ey = x[1][1] - x[0][1];
    const double len = sqrt(ex * ex + ey * ey);
    const double inv = len > 0.0 ? 1.0 / len : 0.0;
    n[0] = ey * inv; n[1] = -ex * inv;
  } else {
    double e1[3], e2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { e1[i] = x[1][i] - x[0][i]; e2[i] = x[2][i] - x[0][i]; }
    n[0] = e1[1] * e2[2] - e1[2] * e2[1];
    n[1] = e1[2] * e2[0] - e1[0] * e2[2];
    n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) n[i] *= inv;
  }
}

// Boundary mass matrix int_Gamma phi_a phi_b ds = |F| (1 + delta_ab) / (D (D + 1)) per facet, rows = compressed
// boundary vertices; bvert maps a compressed vertex to its mesh vertex, facets_c holds compressed ids.
template <int D>
__global__ void __launch_bounds__(128)
k_boundary_mass_rows(int64_t nb, const double* __restrict__ coords, const int32_t* __restrict__ bvert,
                     const int32_t* __restrict__ facets_c, const int64_t* __restrict__ v2f_ptr,
                     const int32_t* __restrict__ v2f_idx, const int64_t* __restrict__ rowptr,
                     const int32_t* __restrict__ col, double* __restrict__ val) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nb) return;
  const int64_t e0 = rowptr[r], e1 = rowptr[r + 1];
  for (int64_t e = e0; e < e1; ++e) val[e] = 0.0;
  for (int64_t t = v2f_ptr[r]; t < v2f_ptr[r + 1]; ++t) {
    const int64_t f = v2f_idx[t];
    double x[D][D], g[D][D];
    int32_t c[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      c[a] = __ldg(facets_c + f * D + a);
      const int64_t v = __ldg(bvert + c[a]);
#pragma unroll
      for (int i = 0; i < D; ++i) x[a][i] = __ldg(coords + v * D + i);
    }
    const double w = facet_measure<D>(x, g) / (double)(D * (D + 1));
#pragma unroll
    for (int a = 0; a < D; ++a) {
      for (int64_t e = e0; e < e1; ++e)
        if (col[e] == c[a]) { val[e] += (c[a] == (int32_t)r) ? 2.0 * w : w; break; }
    }
  }
}

// facet_vec[f, a, :] = (2 n n^T - I) K g_a + 1/2 tr(K) g_a,  K = sum_b kappa_b (x) g_b / |F|,  g_a = d|F|/dx_a:
// the facet's share of int inner((I - (P + P^T)) t_grad(V), t_grad(kappa)) + 1/2 t_div(V) t_div(kappa) ds for
// V = phi_a e_i (shape_regularization.py:554-575) with P = t_grad(x, n) = I - n n^T.
template <int D>
__global__ void __launch_bounds__(128)
k_boundary_curvature_derivative(int64_t nf, const double* __restrict__ coords, const int32_t* __restrict__ facets,
                                const double* __restrict__ kappa, double* __restrict__ facet_vec) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf;
       f += (int64_t)gridDim.x * blockDim.x) {
    double x[D][D], g[D][D], n[D], K[D][D];
    load_facet<D>(coords, facets, f, x);
    const double area = facet_measure<D>(x, g);
    facet_normal<D>(x, n);
    const double inv = area > 0.0 ? 1.0 / area : 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) K[i][j] = 0.0;
#pragma unroll
    for (int b = 0; b < D; ++b) {
      const int64_t v = __ldg(facets + f * D + b);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double kb = __ldg(kappa + v * D + i);
#pragma unroll
        for (int j = 0; j < D; ++j) K[i][j] += kb * g[b][j];
      }
    }
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      tr += K[i][i];
    }
    tr *= inv;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double w[D];
      double nw = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc += K[i][j] * g[a][j];
        w[i] = acc * inv;  // (K g_a)_i
        nw += n[i] * w[i];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) facet_vec[(f * D + a) * D + i] = 2.0 * nw * n[i] - w[i] + 0.5 * tr * g[a][i];
    }
  }
}

// opp[f] = the vertex of the cell behind boundary facet f that is not on the facet (orients the facet normal)
template <int D>
__global__ void __launch_bounds__(128)
k_facet_opposite(int64_t nf, const int32_t* __restrict__ facets, const int64_t* __restrict__ v2c_ptr,
                 const int32_t* __restrict__ v2c_idx, const int32_t* __restrict__ cells, int32_t* __restrict__ opp) {
  constexpr int K = D + 1;
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nf) return;
  int32_t fv[D];
#pragma unroll
  for (int a = 0; a < D; ++a) fv[a] = facets[f * D + a];
  int32_t found = -1;
  for (int64_t k = v2c_ptr[fv[0]]; k < v2c_ptr[fv[0] + 1] && found < 0; ++k) {
    int cv[K];
    load_cell<D>(cells, v2c_idx[k], cv);
    int hits = 0, other = -1;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      bool on = false;
#pragma unroll
      for (int a = 0; a < D; ++a) on = on || (cv[j] == fv[a]);
      if (on) ++hits; else other = cv[j];
    }
    if (hits == D) found = other;
  }
  opp[f] = found;
}

// L2 projections of the re-extension mode `normal` (shape_gradient_problem.py:262-297), facet shares of the right-hand
// sides with the OUTWARD unit normal n:  VEC_IN: int (g . n) phi_a ds  (one value per facet vertex);  otherwise
// int phi (n . e_i) phi_a ds  (D values per facet vertex).  P1 data on a flat facet: int f phi_a = |F|/(D(D+1)) (sum_b f_b + f_a).
template <int D, bool VEC_IN>
__global__ void __launch_bounds__(128)
k_boundary_normal_rhs(int64_t nf, const double* __restrict__ coords, const int32_t* __restrict__ facets,
                      const int32_t* __restrict__ opp, const double* __restrict__ in, double* __restrict__ facet_vec) {
  for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
    double x[D][D], g[D][D], n[D];
    load_facet<D>(coords, facets, f, x);
    const double w = facet_measure<D>(x, g) / (double)(D * (D + 1));
    facet_normal<D>(x, n);
    const int64_t o = opp[f];
    double side = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) side += n[i] * (__ldg(coords + o * D + i) - x[0][i]);
    if (side > 0.0) {
#pragma unroll
      for (int i = 0; i < D; ++i) n[i] = -n[i];
    }
    double val[D], sum = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const int64_t v = __ldg(facets + f * D + a);
      if constexpr (VEC_IN) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) t += __ldg(in + v * D + i) * n[i];
        val[a] = t;
      } else {
        val[a] = __ldg(in + v);
      }
      sum += val[a];
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double t = w * (sum + val[a]);
      if constexpr (VEC_IN) {
        facet_vec[f * D + a] = t;
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) facet_vec[(f * D + a) * D + i] = t * n[i];
      }
    }
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_boundary_measure(void* stream, int dim, int64_t nf, const double* coords, const int32_t* facets,
                                   double* result_dev, double* scratch, int64_t scratch_len) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_boundary_measure: dim must be 2 or 3");
  CB_REQUIRE(coords && result_dev && scratch && scratch_len >= RED_SCRATCH_LEN && (facets || nf == 0),
             "cb_boundary_measure: missing argument");
  cudaStream_t s = as_stream(stream);
  int grid = grid_for(nf, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  if (dim == 2) k_boundary_measure<2><<<grid, 256, 0, s>>>(nf, coords, facets, result_dev, scratch);
  else k_boundary_measure<3><<<grid, 256, 0, s>>>(nf, coords, facets, result_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_boundary_measure_gradient(void* stream, int dim, int64_t nf, int64_t nrows, const double* coords,
                                            const int32_t* facets, const int64_t* v2f_ptr, const int32_t* v2f_idx,
                                            double* facet_vec, double* out) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_boundary_measure_gradient: dim must be 2 or 3");
  CB_REQUIRE(coords && v2f_ptr && out && (nf == 0 || (facets && v2f_idx && facet_vec)),
             "cb_boundary_measure_gradient: missing argument");
  cudaStream_t s = as_stream(stream);
  if (nf > 0) {
    const int grid = grid_for(nf, 256, 8);
    if (dim == 2) k_boundary_gradient<2><<<grid, 256, 0, s>>>(nf, coords, facets, facet_vec);
    else k_boundary_gradient<3><<<grid, 256, 0, s>>>(nf, coords, facets, facet_vec);
    CB_LAUNCH_CHECK();
  }
  // a facet is a (dim-1)-simplex: the cell-vector gather with D = dim - 1 and dim components per vertex
  const int grid = grid_for(nrows, 256, 8);
  if (dim == 2) k_gather_cell_vectors<1, 2><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  else k_gather_cell_vectors<2, 3><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_boundary_mass(void* stream, int dim, int64_t nb, const double* coords, const int32_t* bvert,
                                const int32_t* facets_c, const int64_t* v2f_ptr, const int32_t* v2f_idx,
                                const int64_t* rowptr, const int32_t* col, double* val) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_boundary_mass: dim must be 2 or 3");
  CB_REQUIRE(nb == 0 || (coords && bvert && facets_c && v2f_ptr && v2f_idx && rowptr && col && val),
             "cb_boundary_mass: missing argument");
  if (nb == 0) return CB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = (int)((nb + 127) / 128);
  if (dim == 2) k_boundary_mass_rows<2><<<grid, 128, 0, s>>>(nb, coords, bvert, facets_c, v2f_ptr, v2f_idx, rowptr, col, val);
  else k_boundary_mass_rows<3><<<grid, 128, 0, s>>>(nb, coords, bvert, facets_c, v2f_ptr, v2f_idx, rowptr, col, val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_boundary_curvature_derivative(void* stream, int dim, int64_t nf, int64_t nrows, const double* coords,
                                                const int32_t* facets, const double* kappa, const int64_t* v2f_ptr,
                                                const int32_t* v2f_idx, double* facet_vec, double* out) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_boundary_curvature_derivative: dim must be 2 or 3");
  CB_REQUIRE(coords && kappa && v2f_ptr && out && (nf == 0 || (facets && v2f_idx && facet_vec)),
             "cb_boundary_curvature_derivative: missing argument");
  cudaStream_t s = as_stream(stream);
  if (nf > 0) {
    const int grid = grid_for(nf, 128, 8);
    if (dim == 2) k_boundary_curvature_derivative<2><<<grid, 128, 0, s>>>(nf, coords, facets, kappa, facet_vec);
    else k_boundary_curvature_derivative<3><<<grid, 128, 0, s>>>(nf, coords, facets, kappa, facet_vec);
    CB_LAUNCH_CHECK();
  }
  const int grid = grid_for(nrows, 256, 8);
  if (dim == 2) k_gather_cell_vectors<1, 2><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  else k_gather_cell_vectors<2, 3><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_boundary_facet_opposite(void* stream, int dim, int64_t nf, const int32_t* facets, const int64_t* v2c_ptr,
                                          const int32_t* v2c_idx, const int32_t* cells, int32_t* opp) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_boundary_facet_opposite: dim must be 2 or 3");
  CB_REQUIRE(nf == 0 || (facets && v2c_ptr && v2c_idx && cells && opp), "cb_boundary_facet_opposite: missing argument");
  if (nf == 0) return CB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = (int)((nf + 127) / 128);
  if (dim == 2) k_facet_opposite<2><<<grid, 128, 0, s>>>(nf, facets, v2c_ptr, v2c_idx, cells, opp);
  else k_facet_opposite<3><<<grid, 128, 0, s>>>(nf, facets, v2c_ptr, v2c_idx, cells, opp);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_boundary_normal_rhs(void* stream, int dim, int vector_in, int64_t nf, int64_t nrows, const double* coords,
                                      const int32_t* facets, const int32_t* opp, const double* in, const int64_t* v2f_ptr,
                                      const int32_t* v2f_idx, double* facet_vec, double* out) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_boundary_normal_rhs: dim must be 2 or 3");
  CB_REQUIRE(coords && in && v2f_ptr && out && (nf == 0 || (facets && opp && v2f_idx && facet_vec)),
             "cb_boundary_normal_rhs: missing argument");
  cudaStream_t s = as_stream(stream);
  if (nf > 0) {
    const int grid = grid_for(nf, 128, 8);
    if (dim == 2 && vector_in) k_boundary_normal_rhs<2, true><<<grid, 128, 0, s>>>(nf, coords, facets, opp, in, facet_vec);
    else if (dim == 2) k_boundary_normal_rhs<2, false><<<grid, 128, 0, s>>>(nf, coords, facets, opp, in, facet_vec);
    else if (vector_in) k_boundary_normal_rhs<3, true><<<grid, 128, 0, s>>>(nf, coords, facets, opp, in, facet_vec);
    else k_boundary_normal_rhs<3, false><<<grid, 128, 0, s>>>(nf, coords, facets, opp, in, facet_vec);
    CB_LAUNCH_CHECK();
  }
  const int grid = grid_for(nrows, 256, 8);
  if (dim == 2 && vector_in) k_gather_cell_vectors<1, 1><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  else if (dim == 2) k_gather_cell_vectors<1, 2><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  else if (vector_in) k_gather_cell_vectors<2, 1><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  else k_gather_cell_vectors<2, 3><<<grid, 256, 0, s>>>(nrows, v2f_ptr, v2f_idx, facets, facet_vec, out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
