// Per-cell mesh kernels: quality measures, a-priori determinant test, ALE move, functional
// reduction and the a-posteriori intersection test.  One thread = one cell, reductions fused
// (min / max are order independent -> exact; sums use the fixed-order grid reduction).
//
// Compiled with -fmad=false: these values are compared against thresholds (accept / reject
// decisions, cashocs/geometry/mesh_testing.py:117, shape_variable_abstractions.py:123-130), so
// the arithmetic follows the reference C++ (cashocs/geometry/quality.py:101-271) operation
// by operation without fused multiply-adds.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace cb {

constexpr double PI = 3.14159265358979323846;  // DOLFIN_PI

__device__ __forceinline__ double dot3(const double (&a)[3], const double (&b)[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void normalise3(double (&a)[3]) {
  const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  a[0] /= n; a[1] /= n; a[2] /= n;
}

// angles_triangle (quality.py:114-136)
__device__ __forceinline__ void cell_angles(const double (&x)[3][2], double (&ang)[3]) {
  double e0[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], 0.0};
  double e1[3] = {x[2][0] - x[0][0], x[2][1] - x[0][1], 0.0};
  double e2[3] = {x[2][0] - x[1][0], x[2][1] - x[1][1], 0.0};
  normalise3(e0); normalise3(e1); normalise3(e2);
  ang[0] = acos(dot3(e0, e1));
  ang[1] = acos(-dot3(e0, e2));
  ang[2] = acos(dot3(e1, e2));
}

// dihedral_angles (quality.py:140-177)
__device__ __forceinline__ void cell_angles(const double (&x)[4][3], double (&ang)[6]) {
  double e0[3], e1[3], e2[3], e3[3], e4[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    e0[i] = x[1][i] - x[0][i];
    e1[i] = x[2][i] - x[0][i];
    e2[i] = x[3][i] - x[0][i];
    e3[i] = x[2][i] - x[1][i];
    e4[i] = x[3][i] - x[1][i];
  }
  double n0[3], n1[3], n2[3], n3[3];
  cross3(e0, e1, n0); cross3(e0, e2, n1); cross3(e1, e2, n2); cross3(e3, e4, n3);
  normalise3(n0); normalise3(n1); normalise3(n2); normalise3(n3);
  const double m3[3] = {-n3[0], -n3[1], -n3[2]};
  ang[0] = acos(dot3(n0, n1));
  ang[1] = acos(-dot3(n0, n2));
  ang[2] = acos(dot3(n1, n2));
  ang[3] = acos(dot3(n0, n3));
  ang[4] = acos(dot3(n1, m3));
  ang[5] = acos(dot3(n2, n3));
}

template <int D>
__device__ __forceinline__ double quality_angles(const double (&x)[D + 1][D], bool skew) {
  constexpr int NA = (D == 2) ? 3 : 6;
  double ang[NA];
  cell_angles(x, ang);
  const double opt = (D == 2) ? PI / 3.0 : acos(1.0 / 3.0);
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const double hi = (ang[i] - opt) / (PI - opt);
    const double lo = skew ? (opt - ang[i]) / opt : 0.0;
    // std::max(a, b) = (a < b) ? b : a ; std::min_element starts from the first element and
    // replaces it only on '<' (so a NaN in slot 0 survives, later NaNs are skipped)
    const double m = (hi < lo) ? lo : hi;
    const double qi = 1.0 - m;
    if (i == 0 || qi < q) q = qi;
  }
  return q;
}

__device__ __forceinline__ double quality_condition(const double (&x)[3][2]) {
  Geo<2> g;
  simplex_geometry(x, g);
  double nj = 0.0, ni = 0.0;
#pragma unroll
  for (int a = 1; a < 3; ++a) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double e = x[a][i] - x[0][i];
      nj += e * e;
      ni += g.G[a][i] * g.G[a][i];
    }
  }
  return sqrt(2.0) / (sqrt(nj) * sqrt(ni));
}
__device__ __forceinline__ double quality_condition(const double (&x)[4][3]) {
  Geo<3> g;
  simplex_geometry(x, g);
  double nj = 0.0, ni = 0.0;
#pragma unroll
  for (int a = 1; a < 4; ++a) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double e = x[a][i] - x[0][i];
      nj += e * e;
      ni += g.G[a][i] * g.G[a][i];
    }
  }
  return sqrt(3.0) / (sqrt(nj) * sqrt(ni));
}

__device__ __forceinline__ double dist2(const double (&a)[2], const double (&b)[2]) {
  return sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]));
}
__device__ __forceinline__ double dist3(const double (&a)[3], const double (&b)[3]) {
  return sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}

// dolfin Cell::radius_ratio = d * inradius / circumradius [third party]
__device__ __forceinline__ double quality_radius_ratio(const double (&x)[3][2]) {
  const double a = dist2(x[1], x[2]), b = dist2(x[0], x[2]), c = dist2(x[0], x[1]);
  const double e1x = x[1][0] - x[0][0], e1y = x[1][1] - x[0][1];
  const double e2x = x[2][0] - x[0][0], e2y = x[2][1] - x[0][1];
  const double area = 0.5 * fabs(e1x * e2y - e1y * e2x);
  const double r_in = 2.0 * area / (a + b + c);
  const double r_circ = a * b * c / (4.0 * area);
  return 2.0 * r_in / r_circ;
}
__device__ __forceinline__ double tri_area3(const double (&p)[3], const double (&q)[3], const double (&r)[3]) {
  const double u[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]};
  const double v[3] = {r[0] - p[0], r[1] - p[1], r[2] - p[2]};
  double c[3];
  cross3(u, v, c);
  return 0.5 * sqrt(dot3(c, c));
}
__device__ __forceinline__ double quality_radius_ratio(const double (&x)[4][3]) {
  const double e1[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
  const double e2[3] = {x[2][0] - x[0][0], x[2][1] - x[0][1], x[2][2] - x[0][2]};
  const double e3[3] = {x[3][0] - x[0][0], x[3][1] - x[0][1], x[3][2] - x[0][2]};
  double c[3];
  cross3(e2, e3, c);
  const double vol = fabs(dot3(e1, c)) / 6.0;
  const double area = tri_area3(x[1], x[2], x[3]) + tri_area3(x[0], x[2], x[3]) +
                      tri_area3(x[0], x[1], x[3]) + tri_area3(x[0], x[1], x[2]);
  const double r_in = 3.0 * vol / area;
  const double la = dist3(x[1], x[0]) * dist3(x[3], x[2]);
  const double lb = dist3(x[2], x[0]) * dist3(x[3], x[1]);
  const double lc = dist3(x[3], x[0]) * dist3(x[2], x[1]);
  double s = (la + lb + lc) * (la + lb - lc) * (la - lb + lc) * (-la + lb + lc);
  if (s < 0.0) s = 0.0;
  const double r_circ = sqrt(s) / (24.0 * vol);
  return 3.0 * r_in / r_circ;
}

template <int D>
__global__ void __launch_bounds__(256)
k_quality(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
          int measure, double* __restrict__ q, double* __restrict__ red, double* __restrict__ scratch) {
  __shared__ double s_red[66];
  double qmin = __longlong_as_double(0x7ff0000000000000LL), qsum = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    int v[D + 1];
    double x[D + 1][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    double qc;
    if (measure == CB_Q_SKEWNESS) qc = quality_angles<D>(x, true);
    else if (measure == CB_Q_MAXIMUM_ANGLE) qc = quality_angles<D>(x, false);
    else if (measure == CB_Q_RADIUS_RATIOS) qc = quality_radius_ratio(x);
    else qc = quality_condition(x);
    if (q) q[c] = qc;
    qmin = red_combine(qmin, qc, RED_MIN);
    qsum += qc;
  }
  double v2[2] = {qmin, qsum};
  const int ops[2] = {RED_MIN, RED_SUM};
  if (grid_reduce<2>(v2, ops, scratch, s_red)) {
    if (threadIdx.x == 0) { red[0] = v2[0]; red[1] = v2[1]; }
  }
}

// grad W = sum_a W_a (x) G_a for a vertex field W
template <int D>
__device__ __forceinline__ void field_gradient(const double* __restrict__ W, const int (&v)[D + 1],
                                               const Geo<D>& g, double (&gw)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = 0; j < D; ++j) gw[i][j] = 0.0;
  }
#pragma unroll
  for (int a = 0; a < D + 1; ++a) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double wa = W[(int64_t)v[a] * D + i];
#pragma unroll
      for (int j = 0; j < D; ++j) gw[i][j] += wa * g.G[a][j];
    }
  }
}

// vol[c] = |K_c| and the maximum over the cells: CellVolume of the `inhomogeneous` elasticity scaling
// (cashocs/_forms/shape_form_handler.py:629-644 projects it to DG0 with l2_projection, cashocs/_utils/linalg.py:768-809,
// which is the identity for a cellwise constant; the form handler then divides by the maximum)
template <int D>
__global__ void __launch_bounds__(256)
k_cell_volumes(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
               double* __restrict__ vol, double* __restrict__ red, double* __restrict__ scratch) {
  __shared__ double s_red[33];
  double mx = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    int v[D + 1];
    double x[D + 1][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    Geo<D> g;
    simplex_geometry(x, g);
    vol[c] = g.vol;
    mx = g.vol > mx ? g.vol : mx;
  }
  double v1[1] = {mx};
  const int ops[1] = {RED_MAX};
  if (grid_reduce<1>(v1, ops, scratch, s_red)) {
    if (threadIdx.x == 0) red[0] = v1[0];
  }
}

template <int D>
__global__ void __launch_bounds__(256)
k_det_check(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
            const double* __restrict__ V, double step, double* __restrict__ det,
            double* __restrict__ red, double* __restrict__ scratch, int mode /* 0 det, 1 frobenius */) {
  __shared__ double s_red[66];
  double mn = __longlong_as_double(0x7ff0000000000000LL), mx = __longlong_as_double(0xfff0000000000000LL);
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    int v[D + 1];
    double x[D + 1][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    Geo<D> g;
    simplex_geometry(x, g);
    double gw[D][D];
    field_gradient<D>(V, v, g, gw);
    double val;
    if (mode == 0) {
      double F[D][D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) F[i][j] = step * gw[i][j] + (i == j ? 1.0 : 0.0);
      }
      if constexpr (D == 2) {
        val = F[0][0] * F[1][1] - F[0][1] * F[1][0];
      } else {
        val = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) -
              F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
              F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
      }
    } else {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
#pragma unroll
        for (int j = 0; j < D; ++j) s += gw[i][j] * gw[i][j];
      }
      val = step * sqrt(s);
    }
    if (det) det[c] = val;
    mn = red_combine(mn, val, RED_MIN);
    mx = red_combine(mx, val, RED_MAX);
  }
  double v2[2] = {mn, mx};
  const int ops[2] = {RED_MIN, RED_MAX};
  if (grid_reduce<2>(v2, ops, scratch, s_red)) {
    if (threadIdx.x == 0) { red[0] = v2[0]; red[1] = v2[1]; }
  }
}

__global__ void __launch_bounds__(256)
k_move(int64_t n, double* __restrict__ coords, double* __restrict__ old, const double* __restrict__ V,
       double step) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double c = coords[i];
    if (old) old[i] = c;
    coords[i] = c + step * V[i];
  }
}

template <int D>
__global__ void __launch_bounds__(256)
k_functional(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
             double c_u, const double* __restrict__ u, double c_track, const double* __restrict__ ud,
             double* __restrict__ result, double* __restrict__ scratch) {
  constexpr int K = D + 1;
  constexpr double MREF = 1.0 / ((D + 1) * (D + 2));
  __shared__ double s_red[33];
  double local = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    int v[K];
    double x[K][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    Geo<D> g;
    simplex_geometry(x, g);
    double usum = 0.0, esum = 0.0, e2 = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      const double ua = u[v[a]];
      usum += ua;
      if (ud) {
        const double e = ua - ud[v[a]];
        esum += e;
        e2 += e * e;
      }
    }
    double cellv = c_u * g.vol * usum * (1.0 / K);
    if (ud) cellv += 0.5 * c_track * g.vol * MREF * (e2 + esum * esum);
    local += cellv;
  }
  double v1[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v1, ops, scratch, s_red)) {
    if (threadIdx.x == 0) *result = v1[0];
  }
}

// ------------------------------------------------------------------ intersection test
struct GridInfo {
  double lo[3];
  double inv_h[3];
  int res[3];
};

__global__ void k_bbox(int64_t nv, int dim, const double* __restrict__ coords, double* __restrict__ out,
                       double* __restrict__ scratch, int axis) {
  __shared__ double s_red[66];
  double mn = __longlong_as_double(0x7ff0000000000000LL), mx = __longlong_as_double(0xfff0000000000000LL);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double c = coords[i * dim + axis];
    mn = red_combine(mn, c, RED_MIN);
    mx = red_combine(mx, c, RED_MAX);
  }
  double v2[2] = {mn, mx};
  const int ops[2] = {RED_MIN, RED_MAX};
  if (grid_reduce<2>(v2, ops, scratch, s_red)) {
    if (threadIdx.x == 0) { out[2 * axis] = v2[0]; out[2 * axis + 1] = v2[1]; }
  }
}

__device__ __forceinline__ int bin_coord(double x, double lo, double inv_h, int res) {
  int b = (int)floor((x - lo) * inv_h);
  if (b < 0) b = 0;
  if (b >= res) b = res - 1;
  return b;
}

template <int D>
__global__ void k_bin_count(int64_t nv, const double* __restrict__ coords, GridInfo gi,
                            int32_t* __restrict__ vbin, unsigned long long* __restrict__ cnt) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv;
       v += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = 0;
#pragma unroll
    for (int i = D - 1; i >= 0; --i) b = b * gi.res[i] + bin_coord(coords[v * D + i], gi.lo[i], gi.inv_h[i], gi.res[i]);
    vbin[v] = (int32_t)b;
    atomicAdd(&cnt[b], 1ULL);
  }
}

__global__ void k_bin_fill(int64_t nv, const int32_t* __restrict__ vbin, const int64_t* __restrict__ binptr,
                           unsigned int* __restrict__ cursor, int32_t* __restrict__ binverts) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = vbin[v];
    binverts[binptr[b] + atomicAdd(&cursor[b], 1u)] = (int32_t)v;
  }
}

// For every cell: visit the vertices binned inside its bounding box and count the ones lying in
// the closed simplex (own vertices always collide, like dolfin's exact predicates).
template <int D>
__global__ void __launch_bounds__(128)
k_collide(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells, GridInfo gi,
          const int64_t* __restrict__ binptr, const int32_t* __restrict__ binverts,
          const int32_t* __restrict__ valence, int32_t* __restrict__ counts,
          volatile int* __restrict__ abort_flag) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    // the outcome is a boolean (all counts == valence): once one vertex has collided with more
    // cells than its valence the test has failed, so stop (bounds the work on tangled meshes)
    if (*abort_flag) return;
    int v[D + 1];
    double x[D + 1][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    Geo<D> g;
    simplex_geometry(x, g);
    int blo[3] = {0, 0, 0}, bhi[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double mn = x[0][i], mx = x[0][i];
#pragma unroll
      for (int a = 1; a < D + 1; ++a) { mn = fmin(mn, x[a][i]); mx = fmax(mx, x[a][i]); }
      blo[i] = bin_coord(mn, gi.lo[i], gi.inv_h[i], gi.res[i]);
      bhi[i] = bin_coord(mx, gi.lo[i], gi.inv_h[i], gi.res[i]);
    }
    for (int bz = blo[2]; bz <= bhi[2]; ++bz) {
      for (int by = blo[1]; by <= bhi[1]; ++by) {
        for (int bx = blo[0]; bx <= bhi[0]; ++bx) {
          const int64_t b = ((D == 3 ? (int64_t)bz * gi.res[1] : 0) + by) * gi.res[0] + bx;
          for (int64_t t = binptr[b]; t < binptr[b + 1]; ++t) {
            const int32_t w = binverts[t];
            bool hit = false;
#pragma unroll
            for (int a = 0; a < D + 1; ++a) hit |= (w == v[a]);
            if (!hit) {
              // barycentric coordinates lambda_a = G_a . (p - x_0) (+1 for a = 0)
              double pr[D];
#pragma unroll
              for (int i = 0; i < D; ++i) pr[i] = coords[(int64_t)w * D + i] - x[0][i];
              bool inside = true;
              double l0 = 1.0;
#pragma unroll
              for (int a = 1; a < D + 1; ++a) {
                double l = 0.0;
#pragma unroll
                for (int i = 0; i < D; ++i) l += g.G[a][i] * pr[i];
                inside &= (l >= 0.0);
                l0 -= l;
              }
              inside &= (l0 >= 0.0);
              hit = inside;
            }
            if (hit) {
              if (atomicAdd(&counts[w], 1) >= valence[w]) *abort_flag = 1;
            }
          }
          if (*abort_flag) return;
        }
      }
    }
  }
}

__global__ void k_count_mismatch(int64_t nv, const int32_t* __restrict__ counts,
                                 const int32_t* __restrict__ valence, unsigned long long* __restrict__ out) {
  unsigned long long local = 0;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv;
       v += (int64_t)gridDim.x * blockDim.x)
    local += (counts[v] != valence[v]) ? 1ULL : 0ULL;
  if (local) atomicAdd(out, local);
}

inline int red_grid(int64_t n) {
  int g = grid_for(n, 256, 4);
  return g > RED_MAXP ? RED_MAXP : g;
}

}  // namespace cb

using namespace cb;

extern "C" int cb_mesh_quality(void* stream, int dim, int64_t nc, const double* coords,
                               const int32_t* cells, int measure, double* q, double* red_dev,
                               double* scratch, int64_t scratch_len) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_mesh_quality: dim must be 2 or 3");
  CB_REQUIRE(measure >= 0 && measure <= 3, "cb_mesh_quality: unknown measure");
  CB_REQUIRE(coords && cells && red_dev && scratch && scratch_len >= RED_SCRATCH_LEN, "cb_mesh_quality: missing argument");
  cudaStream_t s = as_stream(stream);
  if (dim == 2) k_quality<2><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, measure, q, red_dev, scratch);
  else k_quality<3><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, measure, q, red_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

static int det_launch(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells,
                      const double* V, double step, double* det, double* red_dev, double* scratch,
                      int64_t scratch_len, int mode) {
  CB_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  CB_REQUIRE(coords && cells && V && red_dev && scratch && scratch_len >= RED_SCRATCH_LEN, "missing argument");
  cudaStream_t s = as_stream(stream);
  if (dim == 2) k_det_check<2><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, V, step, det, red_dev, scratch, mode);
  else k_det_check<3><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, V, step, det, red_dev, scratch, mode);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_det_check(void* stream, int dim, int64_t nc, const double* coords,
                            const int32_t* cells, const double* V, double step, double* det,
                            double* red_dev, double* scratch, int64_t scratch_len) {
  return det_launch(stream, dim, nc, coords, cells, V, step, det, red_dev, scratch, scratch_len, 0);
}

extern "C" int cb_gradient_frobenius_max(void* stream, int dim, int64_t nc, const double* coords,
                                         const int32_t* cells, const double* D, double* red_dev,
                                         double* scratch, int64_t scratch_len) {
  return det_launch(stream, dim, nc, coords, cells, D, 1.0, nullptr, red_dev, scratch, scratch_len, 1);
}

extern "C" int cb_cell_volumes(void* stream, int dim, int64_t nc, const double* coords, const int32_t* cells, double* vol,
                               double* max_dev, double* scratch, int64_t scratch_len) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_cell_volumes: dim must be 2 or 3");
  CB_REQUIRE(coords && cells && vol && max_dev, "cb_cell_volumes: missing argument");
  CB_REQUIRE(scratch && scratch_len >= RED_SCRATCH_LEN, "cb_cell_volumes: scratch too small");
  cudaStream_t s = as_stream(stream);
  if (dim == 2) k_cell_volumes<2><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, vol, max_dev, scratch);
  else k_cell_volumes<3><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, vol, max_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_move_mesh(void* stream, int64_t n, double* coords, double* coords_old,
                            const double* V, double step) {
  CB_REQUIRE(coords && V, "cb_move_mesh: missing argument");
  k_move<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(n, coords, coords_old, V, step);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_functional(void* stream, int dim, int64_t nc, const double* coords,
                             const int32_t* cells, double c_u, const double* u, double c_track,
                             const double* ud, double* result_dev, double* scratch,
                             int64_t scratch_len) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_functional: dim must be 2 or 3");
  CB_REQUIRE(coords && cells && u && result_dev && scratch && scratch_len >= RED_SCRATCH_LEN, "cb_functional: missing argument");
  cudaStream_t s = as_stream(stream);
  if (dim == 2) k_functional<2><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, c_u, u, c_track, ud, result_dev, scratch);
  else k_functional<3><<<red_grid(nc), 256, 0, s>>>(nc, coords, cells, c_u, u, c_track, ud, result_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_intersection_test(void* stream, int dim, int64_t nv, int64_t nc,
                                    const double* coords, const int32_t* cells,
                                    const int32_t* valence, int32_t* counts, int64_t* mismatch_dev) {
  CB_REQUIRE(dim == 2 || dim == 3, "cb_intersection_test: dim must be 2 or 3");
  CB_REQUIRE(coords && cells && valence && counts && mismatch_dev, "cb_intersection_test: missing argument");
  configure_scratch_pool();
  cudaStream_t s = as_stream(stream);
  double* bb = nullptr;
  double* scratch = nullptr;
  CB_CUDA_CHECK(cudaMallocAsync(&bb, sizeof(double) * 8, s));
  CB_CUDA_CHECK(cudaMallocAsync(&scratch, sizeof(double) * RED_SCRATCH_LEN, s));
  CB_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double) * RED_SCRATCH_LEN, s));
  for (int a = 0; a < dim; ++a) {
    k_bbox<<<red_grid(nv), 256, 0, s>>>(nv, dim, coords, bb, scratch, a);
    CB_LAUNCH_CHECK();
  }
  double hbb[6];
  CB_CUDA_CHECK(cudaMemcpyAsync(hbb, bb, sizeof(double) * 2 * dim, cudaMemcpyDeviceToHost, s));
  CB_CUDA_CHECK(cudaStreamSynchronize(s));
  GridInfo gi;
  const double per_axis = pow((double)nc, 1.0 / dim);
  int res = (int)ceil(per_axis);
  if (res < 1) res = 1;
  if (res > 1024) res = 1024;
  int64_t nbins = 1;
  for (int i = 0; i < 3; ++i) {
    if (i < dim) {
      const double ext = hbb[2 * i + 1] - hbb[2 * i];
      gi.lo[i] = hbb[2 * i];
      gi.res[i] = (ext > 0.0) ? res : 1;
      gi.inv_h[i] = (ext > 0.0) ? gi.res[i] / ext : 0.0;
      if (!(ext == ext)) {  // NaN coordinates: every vertex mismatches
        gi.res[i] = 1; gi.inv_h[i] = 0.0;
      }
    } else {
      gi.lo[i] = 0.0; gi.res[i] = 1; gi.inv_h[i] = 0.0;
    }
    nbins *= gi.res[i];
  }
  int32_t* vbin = nullptr;
  int32_t* binverts = nullptr;
  unsigned long long* cnt = nullptr;
  int64_t* binptr = nullptr;
  unsigned int* cursor = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  CB_CUDA_CHECK(cudaMallocAsync(&vbin, sizeof(int32_t) * nv, s));
  CB_CUDA_CHECK(cudaMallocAsync(&binverts, sizeof(int32_t) * nv, s));
  CB_CUDA_CHECK(cudaMallocAsync(&cnt, sizeof(unsigned long long) * (nbins + 1), s));
  CB_CUDA_CHECK(cudaMallocAsync(&binptr, sizeof(int64_t) * (nbins + 1), s));
  CB_CUDA_CHECK(cudaMallocAsync(&cursor, sizeof(unsigned int) * nbins, s));
  CB_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (nbins + 1), s));
  CB_CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(unsigned int) * nbins, s));
  CB_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * nv, s));
  CB_CUDA_CHECK(cudaMemsetAsync(mismatch_dev, 0, sizeof(int64_t), s));
  int* abort_flag = reinterpret_cast<int*>(bb + 6);
  CB_CUDA_CHECK(cudaMemsetAsync(abort_flag, 0, sizeof(int), s));
  const int vgrid = grid_for(nv, 256, 8);
  if (dim == 2) k_bin_count<2><<<vgrid, 256, 0, s>>>(nv, coords, gi, vbin, cnt);
  else k_bin_count<3><<<vgrid, 256, 0, s>>>(nv, coords, gi, vbin, cnt);
  CB_LAUNCH_CHECK();
  CB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const long long*)cnt, (long long*)binptr, nbins + 1, s));
  CB_CUDA_CHECK(cudaMallocAsync(&tmp, tmp_bytes, s));
  CB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const long long*)cnt, (long long*)binptr, nbins + 1, s));
  k_bin_fill<<<vgrid, 256, 0, s>>>(nv, vbin, binptr, cursor, binverts);
  CB_LAUNCH_CHECK();
  const int cgrid = grid_for(nc, 128, 12);
  if (dim == 2) k_collide<2><<<cgrid, 128, 0, s>>>(nc, coords, cells, gi, binptr, binverts, valence, counts, abort_flag);
  else k_collide<3><<<cgrid, 128, 0, s>>>(nc, coords, cells, gi, binptr, binverts, valence, counts, abort_flag);
  CB_LAUNCH_CHECK();
  k_count_mismatch<<<vgrid, 256, 0, s>>>(nv, counts, valence, (unsigned long long*)mismatch_dev);
  CB_LAUNCH_CHECK();
  CB_CUDA_CHECK(cudaFreeAsync(tmp, s));
  CB_CUDA_CHECK(cudaFreeAsync(vbin, s));
  CB_CUDA_CHECK(cudaFreeAsync(binverts, s));
  CB_CUDA_CHECK(cudaFreeAsync(cnt, s));
  CB_CUDA_CHECK(cudaFreeAsync(binptr, s));
  CB_CUDA_CHECK(cudaFreeAsync(cursor, s));
  CB_CUDA_CHECK(cudaFreeAsync(bb, s));
  CB_CUDA_CHECK(cudaFreeAsync(scratch, s));
  return CB_OK;
}
