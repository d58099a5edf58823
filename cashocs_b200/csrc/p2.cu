// P2 (quadratic Lagrange) state / adjoint systems on affine simplices -- BASELINE configs[4].
//
// Replaces the dolfin cell loop + FFC tabulate_tensor kernels behind assemble_petsc_system
// (cashocs/_utils/linalg.py:107-188) for FunctionSpace(mesh, "CG", 2) states and adjoints, the
// shape-derivative assembly (cashocs/_pde_problems/shape_gradient_problem.py:142-144) when the
// Lagrangian holds P2 functions (the test functions stay vector P1: the reference forces the
// deformation space to CG1, shape_optimization_problem.py:251-256) and IntegralFunctional.evaluate
// (cashocs/_optimization/cost_functional.py:160-168).
//
// Everything cell-independent of the affine P2 element is precomputed by the host layer
// (cb_p2_view: ref_R / ref_M / ref_I, basis values at the quadrature points); the kernels only
// contract them with the cell geometry G_k . G_l and |K|.
//
//   * matrix: one WARP per matrix row.  The lanes first compute the geometry of the row's
//     incident cells (dof->cell adjacency) into shared memory, then every lane owns entries of
//     the row and sums over those cells in ascending cell order: no colours, no atomics, no
//     read-modify-write, no nnz->cell list (at 100 entries per cell it would dwarf the matrix).
//   * vectors: one thread per cell writes the local vector to scratch, one thread per dof sums
//     its incident cells in ascending cell order.
#include "common.cuh"

namespace cb {

template <int D>
struct P2 {
  static constexpr int K = D + 1;
  static constexpr int NE = (D == 2) ? 3 : 6;
  static constexpr int ND = K + NE;
};

__device__ __constant__ int c_edge2[3][2] = {{0, 1}, {0, 2}, {1, 2}};
__device__ __constant__ int c_edge3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

template <int D>
__device__ __forceinline__ void edge_vertices(int e, int& i, int& j) {
  if constexpr (D == 2) { i = c_edge2[e][0]; j = c_edge2[e][1]; }
  else { i = c_edge3[e][0]; j = c_edge3[e][1]; }
}

// ------------------------------------------------------------------ matrix (warp per row)
constexpr int P2_WARPS = 4;
constexpr int P2_MAX_PER_LANE = 8;  // 256 entries per row (pattern builder limit)

template <int D>
__global__ void __launch_bounds__(P2_WARPS * 32)
k_p2_matrix(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
            const int64_t* __restrict__ d2c_ptr, const int32_t* __restrict__ d2c_idx,
            const int32_t* __restrict__ cells, const int32_t* __restrict__ cell_dofs,
            const double* __restrict__ coords, const double* __restrict__ ref_R,
            const double* __restrict__ ref_M, double c_stiff, double c_mass,
            const uint8_t* __restrict__ bc_flag, double* __restrict__ val) {
  constexpr int K = P2<D>::K, ND = P2<D>::ND;
  constexpr int NG = K * K;
  __shared__ double s_gg[P2_WARPS][32][NG + 1];  // G_k.G_l and |K| of up to 32 cells of the row
  __shared__ int32_t s_dofs[P2_WARPS][32][ND];
  __shared__ int s_a[P2_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * P2_WARPS + w;
  const int64_t nwarps = (int64_t)gridDim.x * P2_WARPS;
  for (int64_t r = warp0; r < nrows; r += nwarps) {
    const int64_t rb = rowptr[r];
    const int rowlen = (int)(rowptr[r + 1] - rb);
    const int64_t cb0 = d2c_ptr[r];
    const int ncell = (int)(d2c_ptr[r + 1] - cb0);
    double acc[P2_MAX_PER_LANE];
    int32_t mycol[P2_MAX_PER_LANE];
#pragma unroll
    for (int j = 0; j < P2_MAX_PER_LANE; ++j) {
      acc[j] = 0.0;
      const int e = lane + 32 * j;
      mycol[j] = e < rowlen ? col[rb + e] : -1;
    }
    const bool bcr = bc_flag ? (bc_flag[r] != 0) : false;
    for (int base = 0; base < ncell; base += 32) {
      const int nb = min(32, ncell - base);
      __syncwarp();
      if (lane < nb) {
        const int64_t t = d2c_idx[cb0 + base + lane];
        int v[K];
        double x[K][D];
        load_cell<D>(cells, t, v);
        load_coords<D>(coords, v, x);
        Geo<D> g;
        simplex_geometry(x, g);
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int l = 0; l < K; ++l) {
            double d = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) d += g.G[k][i] * g.G[l][i];
            s_gg[w][lane][k * K + l] = d;
          }
        }
        s_gg[w][lane][NG] = g.vol;
        int a = 0;
#pragma unroll
        for (int q = 0; q < ND; ++q) {
          const int32_t dof = cell_dofs[t * ND + q];
          s_dofs[w][lane][q] = dof;
          a = (dof == (int32_t)r) ? q : a;
        }
        s_a[w][lane] = a;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < P2_MAX_PER_LANE; ++j) {
        const int32_t c = mycol[j];
        if (c < 0) continue;
        const bool bcc = bc_flag ? (bc_flag[c] != 0) : false;
        for (int tt = 0; tt < nb; ++tt) {
          int b = -1;
#pragma unroll
          for (int q = 0; q < ND; ++q) b = (s_dofs[w][tt][q] == c) ? q : b;
          if (b < 0) continue;
          const int a = s_a[w][tt];
          double e;
          if (bcr || bcc) {
            e = (c == (int32_t)r) ? 1.0 : 0.0;
          } else {
            // canonical (min, max) pair: entries (a, b) and (b, a) see the same products in the same
            // order (gg is bitwise symmetric), so the assembled matrix is bitwise symmetric
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            const double* R = ref_R + (int64_t)(lo * ND + hi) * NG;
            double sgg = 0.0;
#pragma unroll
            for (int q = 0; q < NG; ++q) sgg += __ldg(R + q) * s_gg[w][tt][q];
            e = s_gg[w][tt][NG] * (c_stiff * sgg + c_mass * __ldg(ref_M + lo * ND + hi));
          }
          acc[j] += e;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < P2_MAX_PER_LANE; ++j) {
      const int e = lane + 32 * j;
      if (e < rowlen) val[rb + e] = acc[j];
    }
  }
}

// ------------------------------------------------------------------ right-hand side (cell kernel)
// qtab [nq][K + 1 + ND]: barycentric point, weight (sum 1), P2 basis values.
template <int D>
__global__ void __launch_bounds__(128)
k_p2_rhs_cells(int64_t nc, const int32_t* __restrict__ cells, const int32_t* __restrict__ cell_dofs,
               const double* __restrict__ coords, const double* __restrict__ ref_R,
               const double* __restrict__ ref_M, const double* __restrict__ ref_I, double c_stiff,
               double c_mass, double s_poly, cb_poly f, int nq, const double* __restrict__ qtab,
               double s_const, double s_nodal, const double* __restrict__ fnodal,
               const uint8_t* __restrict__ bc_flag, const double* __restrict__ bc_val,
               double* __restrict__ cell_out) {
  constexpr int K = P2<D>::K, ND = P2<D>::ND, NG = K * K, QS = K + 1 + ND;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    int v[K];
    double x[K][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    Geo<D> g;
    simplex_geometry(x, g);
    const double vol = g.vol;
    int32_t dofs[ND];
    double be[ND];
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      dofs[a] = cell_dofs[c * ND + a];
      be[a] = s_const * vol * __ldg(ref_I + a);
    }
    if (s_poly != 0.0) {
      for (int iq = 0; iq < nq; ++iq) {
        const double* qt = qtab + (int64_t)iq * QS;
        double xq[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int m = 0; m < K; ++m) {
          const double lm = __ldg(qt + m);
#pragma unroll
          for (int i = 0; i < D; ++i) xq[i] += lm * x[m][i];
        }
        const double fw = s_poly * vol * __ldg(qt + K) * poly_eval(f, xq[0], xq[1], xq[2]);
#pragma unroll
        for (int a = 0; a < ND; ++a) be[a] += fw * __ldg(qt + K + 1 + a);
      }
    }
    if (s_nodal != 0.0) {
      double fe[ND];
#pragma unroll
      for (int a = 0; a < ND; ++a) fe[a] = fnodal[dofs[a]];
#pragma unroll
      for (int a = 0; a < ND; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < ND; ++b) s += __ldg(ref_M + a * ND + b) * fe[b];
        be[a] += s_nodal * vol * s;
      }
    }
    if (bc_flag) {
      bool isbc[ND];
      double gbc[ND];
      bool any = false, lift = false;
#pragma unroll
      for (int a = 0; a < ND; ++a) {
        isbc[a] = bc_flag[dofs[a]] != 0;
        gbc[a] = isbc[a] ? bc_val[dofs[a]] : 0.0;
        any |= isbc[a];
        lift |= gbc[a] != 0.0;
      }
      if (lift) {  // symmetric elimination: b_e -= A_e[:, i] g_i (dolfin SystemAssembler)
        double gg[NG];
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int l = 0; l < K; ++l) {
            double d = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) d += g.G[k][i] * g.G[l][i];
            gg[k * K + l] = d;
          }
        }
        for (int a = 0; a < ND; ++a) {
          double s = 0.0;
          for (int b = 0; b < ND; ++b) {
            if (gbc[b] == 0.0) continue;
            const double* R = ref_R + (int64_t)(a * ND + b) * NG;
            double sgg = 0.0;
#pragma unroll
            for (int q = 0; q < NG; ++q) sgg += __ldg(R + q) * gg[q];
            s += vol * (c_stiff * sgg + c_mass * __ldg(ref_M + a * ND + b)) * gbc[b];
          }
          be[a] -= s;
        }
      }
      if (any) {
#pragma unroll
        for (int a = 0; a < ND; ++a) be[a] = isbc[a] ? gbc[a] : be[a];
      }
    }
#pragma unroll
    for (int a = 0; a < ND; ++a) cell_out[c * ND + a] = be[a];
  }
}

// out[r] = sum over the cells t of dof r (ascending t) of cell_vec[t, local index of r]
template <int ND>
__global__ void __launch_bounds__(256)
k_p2_gather_dofs(int64_t nrows, const int64_t* __restrict__ d2c_ptr, const int32_t* __restrict__ d2c_idx,
                 const int32_t* __restrict__ cell_dofs, const double* __restrict__ cell_vec,
                 double* __restrict__ out) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int64_t k = d2c_ptr[r]; k < d2c_ptr[r + 1]; ++k) {
      const int64_t t = d2c_idx[k];
      int a = 0;
#pragma unroll
      for (int q = 1; q < ND; ++q) a = (cell_dofs[t * ND + q] == (int32_t)r) ? q : a;
      acc += cell_vec[t * ND + a];
    }
    out[r] = acc;
  }
}

// gradient of a P2 function at the K cell vertices (it is linear in between)
template <int D>
__device__ __forceinline__ void p2_vertex_gradients(const double (&ue)[P2<D>::ND], const Geo<D>& g,
                                                    double (&gv)[D + 1][D]) {
  constexpr int K = P2<D>::K, NE = P2<D>::NE;
#pragma unroll
  for (int m = 0; m < K; ++m) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < K; ++a) s += ue[a] * ((a == m) ? 3.0 : -1.0) * g.G[a][i];
      gv[m][i] = s;
    }
  }
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    int i0, j0;
    edge_vertices<D>(e, i0, j0);
    const double ue4 = 4.0 * ue[K + e];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      gv[i0][i] += ue4 * g.G[j0][i];
      gv[j0][i] += ue4 * g.G[i0][i];
    }
  }
}

// ------------------------------------------------------------------ shape derivative (P2 u, p; P1 vector test)
template <int D>
__global__ void __launch_bounds__(128)
k_p2_shape_derivative_cells(int64_t nc, const int32_t* __restrict__ cells, const int32_t* __restrict__ cell_dofs,
                            const double* __restrict__ coords, const double* __restrict__ u,
                            const double* __restrict__ p, cb_poly f, cb_poly gf0, cb_poly gf1, cb_poly gf2,
                            int nq, const double* __restrict__ qtab, double c_int_u,
                            const uint8_t* __restrict__ bc_flag, double* __restrict__ cell_out) {
  constexpr int K = P2<D>::K, ND = P2<D>::ND, QS = K + 1 + ND;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < nc;
       c += (int64_t)gridDim.x * blockDim.x) {
    int v[K];
    double x[K][D];
    load_cell<D>(cells, c, v);
    load_coords<D>(coords, v, x);
    Geo<D> g;
    simplex_geometry(x, g);
    const double vol = g.vol;
    double ue[ND], pe[ND];
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const int32_t dof = cell_dofs[c * ND + a];
      ue[a] = u[dof];
      pe[a] = p[dof];
    }
    double guv[K][D], gpv[K][D];
    p2_vertex_gradients<D>(ue, g, guv);
    p2_vertex_gradients<D>(pe, g, gpv);
    double s0 = 0.0;
    double S[D][D];   // int (gu_j gp_i + gp_j gu_i)
    double T2[K][D];  // int (d_i f) lam_a p
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = 0; j < D; ++j) S[i][j] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int i = 0; i < D; ++i) T2[a][i] = 0.0;
    }
    for (int iq = 0; iq < nq; ++iq) {
      const double* qt = qtab + (int64_t)iq * QS;
      double lam[K];
      double xq[3] = {0.0, 0.0, 0.0};
      double gu[D], gp[D];
#pragma unroll
      for (int i = 0; i < D; ++i) { gu[i] = 0.0; gp[i] = 0.0; }
#pragma unroll
      for (int m = 0; m < K; ++m) {
        lam[m] = __ldg(qt + m);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          xq[i] += lam[m] * x[m][i];
          gu[i] += lam[m] * guv[m][i];
          gp[i] += lam[m] * gpv[m][i];
        }
      }
      const double wq = __ldg(qt + K);
      double uq = 0.0, pq = 0.0;
#pragma unroll
      for (int a = 0; a < ND; ++a) {
        const double ph = __ldg(qt + K + 1 + a);
        uq += ph * ue[a];
        pq += ph * pe[a];
      }
      double gugp = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) gugp += gu[i] * gp[i];
      s0 += wq * (c_int_u * uq + gugp - poly_eval(f, xq[0], xq[1], xq[2]) * pq);
#pragma unroll
      for (int j = 0; j < D; ++j) {
#pragma unroll
        for (int i = 0; i < D; ++i) S[j][i] += wq * (gu[j] * gp[i] + gp[j] * gu[i]);
      }
      double gfq[3];
      gfq[0] = poly_eval(gf0, xq[0], xq[1], xq[2]);
      gfq[1] = poly_eval(gf1, xq[0], xq[1], xq[2]);
      gfq[2] = (D == 3) ? poly_eval(gf2, xq[0], xq[1], xq[2]) : 0.0;
      const double wp = wq * pq;
#pragma unroll
      for (int a = 0; a < K; ++a) {
#pragma unroll
        for (int i = 0; i < D; ++i) T2[a][i] += wp * lam[a] * gfq[i];
      }
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double t1 = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) t1 += g.G[a][j] * S[j][i];
        double e = vol * (g.G[a][i] * s0 - t1 - T2[a][i]);
        if (bc_flag && bc_flag[(int64_t)v[a] * D + i]) e = 0.0;
        cell_out[(c * K + a) * D + i] = e;
      }
    }
  }
}

// ------------------------------------------------------------------ cost functional
template <int D>
__global__ void __launch_bounds__(256)
k_p2_functional(int64_t nc, const double* __restrict__ coords, const int32_t* __restrict__ cells,
                const int32_t* __restrict__ cell_dofs, const double* __restrict__ ref_M,
                const double* __restrict__ ref_I, double c_u, const double* __restrict__ u, double c_track,
                const double* __restrict__ ud, double* __restrict__ result, double* __restrict__ scratch) {
  constexpr int K = P2<D>::K, ND = P2<D>::ND;
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
    double ue[ND], ee[ND];
    double su = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const int32_t dof = cell_dofs[c * ND + a];
      ue[a] = u[dof];
      ee[a] = ud ? ue[a] - ud[dof] : 0.0;
      su += __ldg(ref_I + a) * ue[a];
    }
    double cellv = c_u * su;
    if (ud) {
      double q = 0.0;
#pragma unroll
      for (int a = 0; a < ND; ++a) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < ND; ++b) s += __ldg(ref_M + a * ND + b) * ee[b];
        q += ee[a] * s;
      }
      cellv += 0.5 * c_track * q;
    }
    local += g.vol * cellv;
  }
  double v1[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v1, ops, scratch, s_red)) {
    if (threadIdx.x == 0) *result = v1[0];
  }
}

}  // namespace cb

using namespace cb;

static int check_p2(const cb_p2_view* m) {
  CB_REQUIRE(m != nullptr, "P2 view is NULL");
  CB_REQUIRE(m->dim == 2 || m->dim == 3, "dim must be 2 or 3");
  CB_REQUIRE(m->cells && m->cell_dofs && m->coords, "cells / cell_dofs / coords missing");
  CB_REQUIRE(m->ref_R && m->ref_M && m->ref_I, "reference tensors missing");
  CB_REQUIRE(m->nrows_owned > 0 && m->nrows_owned <= m->ndofs && m->nv_owned <= m->nv && m->nc_owned <= m->nc,
             "owned counts of the P2 view are inconsistent");
  return CB_OK;
}

extern "C" int cb_p2_assemble_matrix(void* stream, const cb_p2_view* m, double c_stiff, double c_mass,
                                     const uint8_t* bc_flag, double* val) {
  if (int e = check_p2(m)) return e;
  CB_REQUIRE(m->rowptr && m->col && m->d2c_ptr && m->d2c_idx && val, "pattern / adjacency / val missing");
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(m->nrows_owned * 32, P2_WARPS * 32, 8);
  if (m->dim == 2)
    k_p2_matrix<2><<<grid, P2_WARPS * 32, 0, s>>>(m->nrows_owned, m->rowptr, m->col, m->d2c_ptr, m->d2c_idx, m->cells,
                                                 m->cell_dofs, m->coords, m->ref_R, m->ref_M, c_stiff, c_mass,
                                                 bc_flag, val);
  else
    k_p2_matrix<3><<<grid, P2_WARPS * 32, 0, s>>>(m->nrows_owned, m->rowptr, m->col, m->d2c_ptr, m->d2c_idx, m->cells,
                                                 m->cell_dofs, m->coords, m->ref_R, m->ref_M, c_stiff, c_mass,
                                                 bc_flag, val);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_p2_assemble_rhs(void* stream, const cb_p2_view* m, double c_stiff, double c_mass,
                                  double s_poly, const cb_poly* f, int nq, const double* qtab,
                                  double s_const, double s_nodal, const double* fnodal,
                                  const uint8_t* bc_flag, const double* bc_val, double* rhs) {
  if (int e = check_p2(m)) return e;
  CB_REQUIRE(m->d2c_ptr && m->d2c_idx && m->cell_scratch && rhs, "adjacency / scratch / rhs missing");
  CB_REQUIRE(s_poly == 0.0 || (f && qtab && nq > 0), "polynomial source needs f and a quadrature table");
  CB_REQUIRE(s_nodal == 0.0 || fnodal, "nodal source missing");
  CB_REQUIRE(bc_flag == nullptr || bc_val != nullptr, "bc_val missing");
  cudaStream_t s = as_stream(stream);
  cb_poly fz;
  memset(&fz, 0, sizeof(fz));
  const cb_poly& ff = f ? *f : fz;
  const int grid = grid_for(m->nc, 128, 12), dgrid = grid_for(m->nrows_owned, 256, 8);
  if (m->dim == 2) {
    k_p2_rhs_cells<2><<<grid, 128, 0, s>>>(m->nc, m->cells, m->cell_dofs, m->coords, m->ref_R, m->ref_M, m->ref_I,
                                          c_stiff, c_mass, s_poly, ff, nq, qtab, s_const, s_nodal, fnodal,
                                          bc_flag, bc_val, m->cell_scratch);
    CB_LAUNCH_CHECK();
    k_p2_gather_dofs<6><<<dgrid, 256, 0, s>>>(m->nrows_owned, m->d2c_ptr, m->d2c_idx, m->cell_dofs, m->cell_scratch, rhs);
  } else {
    k_p2_rhs_cells<3><<<grid, 128, 0, s>>>(m->nc, m->cells, m->cell_dofs, m->coords, m->ref_R, m->ref_M, m->ref_I,
                                          c_stiff, c_mass, s_poly, ff, nq, qtab, s_const, s_nodal, fnodal,
                                          bc_flag, bc_val, m->cell_scratch);
    CB_LAUNCH_CHECK();
    k_p2_gather_dofs<10><<<dgrid, 256, 0, s>>>(m->nrows_owned, m->d2c_ptr, m->d2c_idx, m->cell_dofs, m->cell_scratch, rhs);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_p2_shape_derivative_poisson(void* stream, const cb_p2_view* m, const int64_t* v2c_ptr,
                                              const int32_t* v2c_idx, const double* u, const double* p,
                                              const cb_poly* f, const cb_poly* gradf, int nq,
                                              const double* qtab, double c_int_u, const uint8_t* bc_flag,
                                              double* rhs) {
  if (int e = check_p2(m)) return e;
  CB_REQUIRE(v2c_ptr && v2c_idx && u && p && f && gradf && qtab && nq > 0 && rhs && m->cell_scratch,
             "cb_p2_shape_derivative_poisson: missing argument");
  cudaStream_t s = as_stream(stream);
  cb_poly gz;
  memset(&gz, 0, sizeof(gz));
  const int grid = grid_for(m->nc, 128, 12), vgrid = grid_for(m->nv_owned, 256, 8);
  if (m->dim == 2) {
    k_p2_shape_derivative_cells<2><<<grid, 128, 0, s>>>(m->nc, m->cells, m->cell_dofs, m->coords, u, p, *f, gradf[0],
                                                       gradf[1], gz, nq, qtab, c_int_u, bc_flag, m->cell_scratch);
    CB_LAUNCH_CHECK();
    k_gather_cell_vectors<2, 2><<<vgrid, 256, 0, s>>>(m->nv_owned, v2c_ptr, v2c_idx, m->cells, m->cell_scratch, rhs);
  } else {
    k_p2_shape_derivative_cells<3><<<grid, 128, 0, s>>>(m->nc, m->cells, m->cell_dofs, m->coords, u, p, *f, gradf[0],
                                                       gradf[1], gradf[2], nq, qtab, c_int_u, bc_flag,
                                                       m->cell_scratch);
    CB_LAUNCH_CHECK();
    k_gather_cell_vectors<3, 3><<<vgrid, 256, 0, s>>>(m->nv_owned, v2c_ptr, v2c_idx, m->cells, m->cell_scratch, rhs);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int cb_p2_functional(void* stream, const cb_p2_view* m, double c_u, const double* u, double c_track,
                                const double* ud, double* result_dev, double* scratch, int64_t scratch_len) {
  if (int e = check_p2(m)) return e;
  CB_REQUIRE(u && result_dev && scratch && scratch_len >= RED_SCRATCH_LEN, "cb_p2_functional: missing argument");
  cudaStream_t s = as_stream(stream);
  int grid = grid_for(m->nc_owned, 256, 4);
  if (grid > RED_MAXP) grid = RED_MAXP;
  if (m->dim == 2)
    k_p2_functional<2><<<grid, 256, 0, s>>>(m->nc_owned, m->coords, m->cells, m->cell_dofs, m->ref_M, m->ref_I, c_u, u,
                                           c_track, ud, result_dev, scratch);
  else
    k_p2_functional<3><<<grid, 256, 0, s>>>(m->nc_owned, m->coords, m->cells, m->cell_dofs, m->ref_M, m->ref_I, c_u, u,
                                           c_track, ud, result_dev, scratch);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
