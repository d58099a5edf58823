// Preconditioned Krylov solvers with PETSc KSP semantics, device-resident control flow.
//
// Replaces LinearSolver.solve (cashocs/_utils/linalg.py:480-546) -> PETSc KSPSolve for
// ksp_type in {cg, minres, preonly} and pc_type in {none, jacobi, chebyshev(jacobi)}.
// Semantics restated from PETSc [third party]: zero initial guess, left preconditioning,
// preconditioned residual norm, KSPConvergedDefault.
//
// All scalars (alpha, beta, norms, iteration count, converged reason) live in a small state
// block in device memory.  Dot products are fused into the vector / SpMV kernels, reduced
// deterministically (fixed-order block partials, last-arriving block finalises) and the
// finalising block also runs the scalar recurrences and the convergence test, so the host
// only enqueues kernels and polls the state every `check_every` iterations; once `done` is
// set the remaining enqueued kernels return immediately, so the iterate is exactly the one
// of the converged iteration.  With a cb_dist the per-rank sums go through one
// ncclAllReduce per reduction point and a one-thread kernel runs the recurrences.
#include <math.h>

#include <utility>
#include <vector>

#include <stdlib.h>

#include "coarse_cycle.cuh"
#include "p2p.cuh"
#include "spmv.cuh"

namespace cb {

P2PView dist_p2p_view(const cb_dist* d);
HaloDev dist_halo_dev(const cb_dist* d);
int64_t dist_ghost_capacity(const cb_dist* d);
int64_t dist_recv_len(const cb_dist* d);

int spmv_residual_dispatch(cudaStream_t s, const cb_mat* A, const double* x, const double* b, double* y, const int* done,
                           bool use32);
int spmv_dispatch(cudaStream_t s, const cb_mat* A, const double* x, double* y, const double* w,
                  double* result, double* scratch, const int* done = nullptr, bool use32 = false);
int amg_restrict(cudaStream_t s, int bs, const cb_amg_level* L, const double* b, const double* t, double* bc,
                 const int* done, int p_block);
int amg_prolong(cudaStream_t s, int bs, const cb_amg_level* L, const double* xc, double* x, const int* done, int p_block);
int amg_dense_apply(cudaStream_t s, int n, const double* inv, const double* b, double* x, const int* done);
int dist_halo_exchange(cudaStream_t s, cb_dist* d, double* x, int bs, double* sendbuf);
int dist_allreduce_sum(cudaStream_t s, cb_dist* d, double* buf, int n);
bool dist_is_one_sided(const cb_dist* d);
int64_t dist_send_len(const cb_dist* d);

struct KspState {
  // sums[] is the landing zone of the fused reductions (allreduced in place when distributed)
  double sums[4];
  double beta, betaold, dpi, bcoef;
  double rnorm, rnorm0, ttol;
  double rtol, atol, dtol;
  // MINRES recurrences
  double oldb, mbeta, dbar, epsln, phibar, cs, sn, alfa, oldeps, delta, gamma, phi;
  int its, reason, done, max_it;
  int hist_len, flexible;
  double alpha;  // step length of the last x / r update (flexible beta)
};
constexpr int STATE_DOUBLES = 64;
static_assert(sizeof(KspState) <= STATE_DOUBLES * sizeof(double), "state block too small");

__device__ __forceinline__ void converged_default(KspState* st, double rnorm, int it) {
  if (it == 0) {
    st->ttol = fmax(st->rtol * rnorm, st->atol);
    st->rnorm0 = rnorm;
  }
  st->rnorm = rnorm;
  int reason = 0;
  if (rnorm != rnorm || isinf(rnorm)) reason = CB_KSP_DIVERGED_NANORINF;
  else if (rnorm <= st->ttol) reason = (rnorm < st->atol) ? CB_KSP_CONVERGED_ATOL : CB_KSP_CONVERGED_RTOL;
  else if (rnorm >= st->dtol * st->rnorm0) reason = CB_KSP_DIVERGED_DTOL;
  else if (it >= st->max_it) reason = CB_KSP_DIVERGED_ITS;
  if (reason) { st->reason = reason; st->done = 1; }
}

// ---- CG scalar recurrences -------------------------------------------------------------
// after (z.r, z.z) are known: called at it = 0 (init) and after every x/r update
__device__ void cg_after_zr(KspState* st, double* hist, int is_init) {
  const double zr = st->sums[0], zz = st->sums[1];
  if (!is_init) st->its += 1;
  st->betaold = st->beta;
  st->beta = zr;
  const double rnorm = sqrt(zz);
  if (hist && st->its < st->hist_len) hist[st->its] = rnorm;
  converged_default(st, rnorm, st->its);
  if (st->done) return;
  // PETSc KSPCG: beta == 0 -> converged (happy breakdown), beta < 0 -> indefinite PC
  if (zr == 0.0) { st->reason = CB_KSP_CONVERGED_ATOL; st->done = 1; return; }
  if (zr < 0.0) { st->reason = CB_KSP_DIVERGED_INDEFINITE_PC; st->done = 1; return; }
  // flexible (Polak-Ribiere): z_new.(r_new - r_old) = -alpha z_new.w since r_new = r_old - alpha w (sums[2] = z_new.w)
  if (is_init) st->bcoef = 0.0;
  else st->bcoef = st->flexible ? -st->alpha * st->sums[2] / st->betaold : st->beta / st->betaold;
}

// after p.Ap is known
__device__ void cg_after_pw(KspState* st) {
  const double dpi = st->sums[0];
  st->dpi = dpi;
  if (dpi != dpi) { st->reason = CB_KSP_DIVERGED_NANORINF; st->done = 1; st->its += 1; }
  else if (dpi <= 0.0) { st->reason = CB_KSP_DIVERGED_INDEFINITE_MAT; st->done = 1; st->its += 1; }
}

__global__ void k_state_step(KspState* st, double* hist, int which) {
  if (st->done) return;
  if (which == 0) cg_after_zr(st, hist, 1);
  else if (which == 1) cg_after_zr(st, hist, 0);
  else if (which == 2) cg_after_pw(st);
}

// ---- CG vector kernels -----------------------------------------------------------------
// r = b, x = 0, z = dinv .* r (or r), sums = (z.r, z.z)
template <bool FINALIZE>
__global__ void __launch_bounds__(256)
k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ z, double* __restrict__ x, bool apply_pc,
          KspState* st, double* hist, double* scratch) {
  __shared__ double s_red[66];
  double zr = 0.0, zz = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = b[i];
    r[i] = ri;
    x[i] = 0.0;
    if (apply_pc) {
      const double zi = dinv ? dinv[i] * ri : ri;
      z[i] = zi;
      zr += zi * ri;
      zz += zi * zi;
    }
  }
  if (!apply_pc) return;
  double v[2] = {zr, zz};
  const int ops[2] = {RED_SUM, RED_SUM};
  if (grid_reduce<2>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      st->sums[1] = v[1];
      if (FINALIZE) cg_after_zr(st, hist, 1);
    }
  }
}

// sums = (z.r, z.z) for an already applied preconditioner
__global__ void __launch_bounds__(256)
k_dots_zr(int64_t n, const double* __restrict__ z, const double* __restrict__ r, KspState* st,
          double* hist, double* scratch, int finalize_mode /* -1 none, 0 init, 1 iter */,
          const double* __restrict__ w = nullptr /* flexible beta: also z.w */) {
  if (st->done) return;
  __shared__ double s_red[99];
  double zr = 0.0, zz = 0.0, zw = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double zi = z[i];
    zr += zi * r[i];
    zz += zi * zi;
    if (w) zw += zi * w[i];
  }
  double v[3] = {zr, zz, zw};
  const int ops[3] = {RED_SUM, RED_SUM, RED_SUM};
  if (grid_reduce<3>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      st->sums[1] = v[1];
      st->sums[2] = v[2];
      if (finalize_mode >= 0) cg_after_zr(st, hist, finalize_mode == 0);
    }
  }
}

// ---- rigid-body coarse correction: z += Z (Z^T A Z)^-1 Z^T r ------------------------------------
// modes (c = node position relative to the common centre): translations e_k, rotations (-c_y, c_x, 0),
// (0, -c_z, c_y), (c_z, 0, -c_x) in 3-D; e_x, e_y and (-c_y, c_x) in 2-D
template <int BS>
struct Rbm {
  static constexpr int K = BS == 3 ? 6 : 3;
};

// out[0..K) = Z^T r
template <int BS>
__global__ void __launch_bounds__(256)
k_rbm_project(int64_t nn, const double* __restrict__ r, const double* __restrict__ xc, double* __restrict__ out,
              double* __restrict__ scratch, const KspState* st) {
  if (st->done) return;
  constexpr int K = Rbm<BS>::K;
  __shared__ double s_red[32 * K + 1];
  double v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn;
       i += (int64_t)gridDim.x * blockDim.x) {
    double ri[BS], c[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) { ri[a] = r[i * BS + a]; c[a] = xc[i * BS + a]; }
#pragma unroll
    for (int a = 0; a < BS; ++a) v[a] += ri[a];
    if constexpr (BS == 3) {
      v[3] += c[0] * ri[1] - c[1] * ri[0];
      v[4] += c[1] * ri[2] - c[2] * ri[1];
      v[5] += c[2] * ri[0] - c[0] * ri[2];
    } else {
      v[2] += c[0] * ri[1] - c[1] * ri[0];
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

// z += Z (einv * ztr);  sums = (z.r, z.z)
template <int BS>
__global__ void __launch_bounds__(256)
k_rbm_correct(int64_t nn, const double* __restrict__ r, double* __restrict__ z, const double* __restrict__ xc,
              const double* __restrict__ einv, const double* __restrict__ ztr, KspState* st, double* hist,
              double* scratch, int finalize_mode /* -1 none, 0 init, 1 iter */, const double* __restrict__ w = nullptr) {
  if (st->done) return;
  constexpr int K = Rbm<BS>::K;
  __shared__ double s_red[99];
  __shared__ double y[K];
  if (threadIdx.x < K) {
    double acc = 0.0;
    for (int j = 0; j < K; ++j) acc += einv[threadIdx.x * K + j] * ztr[j];
    y[threadIdx.x] = acc;
  }
  __syncthreads();
  double zr = 0.0, zz = 0.0, zw = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nn;
       i += (int64_t)gridDim.x * blockDim.x) {
    double c[BS], dz[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) c[a] = xc[i * BS + a];
    if constexpr (BS == 3) {
      dz[0] = y[0] - c[1] * y[3] + c[2] * y[5];
      dz[1] = y[1] + c[0] * y[3] - c[2] * y[4];
      dz[2] = y[2] + c[1] * y[4] - c[0] * y[5];
    } else {
      dz[0] = y[0] - c[1] * y[2];
      dz[1] = y[1] + c[0] * y[2];
    }
#pragma unroll
    for (int a = 0; a < BS; ++a) {
      const double zi = z[i * BS + a] + dz[a];
      z[i * BS + a] = zi;
      zr += zi * r[i * BS + a];
      zz += zi * zi;
      if (w) zw += zi * w[i * BS + a];
    }
  }
  double v[3] = {zr, zz, zw};
  const int ops[3] = {RED_SUM, RED_SUM, RED_SUM};
  if (grid_reduce<3>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      st->sums[1] = v[1];
      st->sums[2] = v[2];
      if (finalize_mode >= 0) cg_after_zr(st, hist, finalize_mode == 0);
    }
  }
}

// p = z + bcoef * p
__global__ void __launch_bounds__(256)
k_cg_p(int64_t n, const double* __restrict__ z, double* __restrict__ p, const KspState* st) {
  if (st->done) return;
  const double bc = st->bcoef;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = z[i] + bc * p[i];
}

// w = A p ; sums[0] = p.w
template <int BS, bool FINALIZE>
__global__ void __launch_bounds__(SpmvCfg<BS>::THREADS)
k_cg_spmv(int64_t nrows, int G, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
          const double* __restrict__ val, const double* __restrict__ p, double* __restrict__ w,
          KspState* st, double* scratch) {
  if (st->done) return;
  using C = SpmvCfg<BS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[33];
  const SpmvSmem<BS> sm = spmv_smem_init<BS>(smem_raw);
  uint32_t phase = 0;
  const int R = spmv_rows_per_cta<BS>(G);
  const int64_t nrb = (nrows + R - 1) / R;
  double local = 0.0;
  for (int64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
    const int64_t row0 = rb * R;
    const int nr = (int)((nrows - row0 < R) ? (nrows - row0) : R);
    double acc;
    const int64_t i = spmv_row_block<BS, double>(sm, rowptr, col, val, p, row0, nr, G, phase, acc);
    if (i >= 0) {
      w[i] = acc;
      local += acc * p[i];
    }
  }
  double v[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      if (FINALIZE) cg_after_pw(st);
    }
  }
}

// x += a p ; r -= a w ; [z = dinv .* r ; sums = (z.r, z.z)]
template <bool FINALIZE>
__global__ void __launch_bounds__(256)
k_cg_xr(int64_t n, const double* __restrict__ p, const double* __restrict__ w,
        const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
        double* __restrict__ z, bool fused_jacobi, KspState* st, double* hist, double* scratch,
        const double* __restrict__ pre_dinv = nullptr, double pre_scale = 0.0, double* __restrict__ pre_d = nullptr) {
  if (st->done) return;
  __shared__ double s_red[66];
  const double a = st->beta / st->dpi;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->alpha = a;  // nobody reads it in this kernel
  double zr = 0.0, zz = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += a * p[i];
    const double ri = r[i] - a * w[i];
    r[i] = ri;
    if (pre_dinv) {  // fused pre-smoothing of the AMG V-cycle (zero guess): z = dinv .* r / theta
      const double zi = pre_scale * pre_dinv[i] * ri;
      z[i] = zi;
      if (pre_d) pre_d[i] = zi;
    }
    if (fused_jacobi) {
      const double zi = dinv ? dinv[i] * ri : ri;
      z[i] = zi;
      zr += zi * ri;
      zz += zi * zi;
    }
  }
  if (!fused_jacobi) return;
  double v[2] = {zr, zz};
  const int ops[2] = {RED_SUM, RED_SUM};
  if (grid_reduce<2>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      st->sums[1] = v[1];
      if (FINALIZE) cg_after_zr(st, hist, 0);
    }
  }
}

// ---- Chebyshev(Jacobi) polynomial preconditioner ------------------------------------------
// d = dinv .* r / theta ; z = d
__global__ void __launch_bounds__(256)
k_cheb_first(int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
             double inv_theta, double* __restrict__ d, double* __restrict__ z, const KspState* st) {
  if (st->done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double di = dinv[i] * r[i] * inv_theta;
    d[i] = di;
    z[i] = di;
  }
}
// d = c1 d + c2 dinv .* (r - t) ; z += d
__global__ void __launch_bounds__(256)
k_cheb_step(int64_t n, const double* __restrict__ r, const double* __restrict__ t,
            const double* __restrict__ dinv, double c1, double c2, double* __restrict__ d,
            double* __restrict__ z, const KspState* st) {
  if (st->done) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double dprev = (c1 != 0.0) ? d[i] : 0.0;  // c1 == 0: first step from a non-zero guess
    const double di = c1 * dprev + c2 * dinv[i] * (r[i] - t[i]);
    d[i] = di;
    z[i] += di;
  }
}

// b (replicated level, global row order) from the all-gathered buffer [own rows | rows of the other ranks in rank order]
__global__ void __launch_bounds__(256)
k_unpack_allgather(int64_t n_total, int64_t n_own, int64_t lo, int bs, const double* __restrict__ tmp,
                   double* __restrict__ b, const KspState* st) {
  if (st->done) return;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n_total * bs;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / bs;
    const int64_t src = (i < lo) ? (n_own + i) : (i < lo + n_own ? i - lo : i);
    b[q] = tmp[src * bs + (q - i * bs)];
  }
}

// out[i, :] = in[gid[i], :]  (local copy of a replicated coarse vector in this rank's coarse numbering)
__global__ void __launch_bounds__(256)
k_gather_nodes(int64_t n, int bs, const int32_t* __restrict__ gid, const double* __restrict__ in,
               double* __restrict__ out, const KspState* st) {
  if (st->done) return;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n * bs;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = q / bs;
    out[q] = in[(int64_t)gid[i] * bs + (q - i * bs)];
  }
}

// First post-smoothing step of the AMG V-cycle without touching A: after the coarse-grid correction
// x1 = x0 + P e the residual is  b - A x1 = b - t - (A P) e  with t = A x0 already known, and (A P) --
// a by-product of the Galerkin product -- has fewer bytes than A.  One pass over AP:
//   z_i += c2 * dinv_i * (b_i - t_i - (AP e)_i)      [optionally d_i = that increment]
// z is only touched at the row a lane owns (the gather reads e, the coarse vector): in place.
// DOTS (fine level, last kernel of the cycle): also sums = (z.b, z.z) for the PCG recurrences.
template <int BS, bool DOTS, typename VT>
__global__ void __launch_bounds__(2 * SpmvCfg<BS, VT>::THREADS)
k_amg_post(int64_t nrows, int G, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
           const VT* __restrict__ val, const double* __restrict__ e, double* __restrict__ z,
           const double* __restrict__ b, const double* __restrict__ t, const double* __restrict__ dinv,
           double c2, double* __restrict__ dvec, KspState* st, double* hist, double* scratch,
           int finalize_mode /* -1 none, 0 init, 1 iter */, const double* __restrict__ w /* flexible beta: z.w */,
           int t_is_residual /* t holds b - A x0 (level_spmv with residual_of) instead of A x0 */) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SpmvSmem<BS, VT> sm = spmv_smem_init<BS, VT>(smem_raw);
  uint32_t phase = 0;
  const int R = (32 / (BS * G)) * (int)(blockDim.x >> 5);  // rows per CTA: the launcher picks the warps (cb_mat.cta_warps)
  const int64_t nrb = (nrows + R - 1) / R;
  double zr = 0.0, zz = 0.0, zw = 0.0;
  for (int64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
    const int64_t row0 = rb * R;
    const int nr = (int)((nrows - row0 < R) ? (nrows - row0) : R);
    double acc;
    const int64_t i = spmv_row_block<BS, VT>(sm, rowptr, col, val, e, row0, nr, G, phase, acc);
    if (i >= 0) {
      const double ri = t_is_residual ? t[i] : b[i] - t[i];
      const double di = c2 * dinv[i] * (ri - acc);
      if (dvec) dvec[i] = di;
      const double zi = z[i] + di;
      z[i] = zi;
      if (DOTS) { zr += zi * b[i]; zz += zi * zi; if (w) zw += zi * w[i]; }
    }
  }
  if (DOTS) {
    // the reduction scratch reuses the (now idle) tile buffer: 9 CTAs x 24.4 KB fill the SM's shared
    // memory to within 1.4 KB, a static array would cost the ninth resident CTA
    __syncthreads();
    double* s_red = reinterpret_cast<double*>(smem_raw);
    double v[3] = {zr, zz, zw};
    const int ops[3] = {RED_SUM, RED_SUM, RED_SUM};
    if (grid_reduce<3>(v, ops, scratch, s_red)) {
      if (threadIdx.x == 0) {
        st->sums[0] = v[0];
        st->sums[1] = v[1];
        st->sums[2] = v[2];
        if (finalize_mode >= 0) cg_after_zr(st, hist, finalize_mode == 0);
      }
    }
  }
}

// ---- MINRES (Paige & Saunders) ------------------------------------------------------------
__device__ void minres_after_beta(KspState* st, double* hist, int is_init) {
  const double b2 = st->sums[0];
  if (b2 < 0.0) { st->reason = CB_KSP_DIVERGED_INDEFINITE_PC; st->done = 1; return; }
  if (b2 != b2) { st->reason = CB_KSP_DIVERGED_NANORINF; st->done = 1; return; }
  const double beta = sqrt(b2);
  if (is_init) {
    st->oldb = 0.0; st->mbeta = beta; st->dbar = 0.0; st->epsln = 0.0; st->phibar = beta;
    st->cs = -1.0; st->sn = 0.0; st->its = 0;
    if (hist && st->hist_len > 0) hist[0] = beta;
    converged_default(st, beta, 0);
    return;
  }
  st->oldb = st->mbeta;
  st->mbeta = beta;
  const double alfa = st->alfa;
  st->oldeps = st->epsln;
  st->delta = st->cs * st->dbar + st->sn * alfa;
  const double gbar = st->sn * st->dbar - st->cs * alfa;
  st->epsln = st->sn * beta;
  st->dbar = -st->cs * beta;
  double gamma = sqrt(gbar * gbar + beta * beta);
  if (gamma < 2.2250738585072014e-308) gamma = 2.2250738585072014e-308;
  st->gamma = gamma;
  st->cs = gbar / gamma;
  st->sn = beta / gamma;
  st->phi = st->cs * st->phibar;
  st->phibar = st->sn * st->phibar;
  st->its += 1;
  const double rn = fabs(st->phibar);
  if (hist && st->its < st->hist_len) hist[st->its] = rn;
  // the x update of this iteration must still be applied: `done` is raised by k_minres_wx
  st->rnorm = rn;
}

__global__ void k_minres_state(KspState* st, double* hist, int which) {
  if (st->done) return;
  if (which == 0) minres_after_beta(st, hist, 1);
  else if (which == 1) minres_after_beta(st, hist, 0);
  else if (which == 2) st->alfa = st->sums[0];
}

// One reduction point of a partitioned Krylov iteration as ONE kernel of one warp: the one-sided all-reduce of
// st->sums[0..n) over NVLink (rank-ordered sum, identical on all ranks) and then the scalar recurrences / convergence
// test of that point.  Ranks stay in lock step even after convergence: the exchange always runs, only the recurrence
// is skipped once `done` is raised (identically everywhere).
__global__ void k_allreduce_state(int n, KspState* st, double* hist, P2PView v, int which, int minres) {
  const bool ok = p2p_allreduce_warp(v, n, st->sums, 0);
  if (threadIdx.x != 0) return;
  if (!ok) {  // a peer timed out (the host raises on the communicator's error word)
    if (!st->done) { st->reason = CB_KSP_DIVERGED_BREAKDOWN; st->done = 1; }
    return;
  }
  if (st->done) return;
  if (minres) {
    if (which == 0) minres_after_beta(st, hist, 1);
    else if (which == 1) minres_after_beta(st, hist, 0);
    else if (which == 2) st->alfa = st->sums[0];
  } else {
    if (which == 0) cg_after_zr(st, hist, 1);
    else if (which == 1) cg_after_zr(st, hist, 0);
    else if (which == 2) cg_after_pw(st);
  }
}

// r1 = b ; y = dinv .* r1 ; sums[0] = r1.y ; x = 0
template <bool FINALIZE>
__global__ void __launch_bounds__(256)
k_minres_init(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
              double* __restrict__ r2, double* __restrict__ y, double* __restrict__ x,
              KspState* st, double* hist, double* scratch) {
  __shared__ double s_red[33];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = b[i];
    r2[i] = ri;
    x[i] = 0.0;
    const double yi = dinv ? dinv[i] * ri : ri;
    y[i] = yi;
    s += ri * yi;
  }
  double v[1] = {s};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      if (FINALIZE) minres_after_beta(st, hist, 1);
    }
  }
}

// v = y / beta
__global__ void __launch_bounds__(256)
k_minres_v(int64_t n, const double* __restrict__ y, double* __restrict__ v, const KspState* st) {
  if (st->done) return;
  const double s = 1.0 / st->mbeta;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    v[i] = s * y[i];
}

// t = A v - (beta/oldb) r1 ; sums[0] = v.t
template <int BS, bool FINALIZE>
__global__ void __launch_bounds__(SpmvCfg<BS>::THREADS)
k_minres_spmv(int64_t nrows, int G, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
              const double* __restrict__ val, const double* __restrict__ v,
              const double* __restrict__ r1, double* __restrict__ t, KspState* st,
              double* scratch) {
  if (st->done) return;
  using C = SpmvCfg<BS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[33];
  const SpmvSmem<BS> sm = spmv_smem_init<BS>(smem_raw);
  uint32_t phase = 0;
  const double c = (st->its >= 1) ? st->mbeta / st->oldb : 0.0;
  const int R = spmv_rows_per_cta<BS>(G);
  const int64_t nrb = (nrows + R - 1) / R;
  double local = 0.0;
  for (int64_t rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
    const int64_t row0 = rb * R;
    const int nr = (int)((nrows - row0 < R) ? (nrows - row0) : R);
    double acc;
    const int64_t i = spmv_row_block<BS, double>(sm, rowptr, col, val, v, row0, nr, G, phase, acc);
    if (i >= 0) {
      if (c != 0.0) acc -= c * r1[i];
      t[i] = acc;
      local += acc * v[i];
    }
  }
  double vv[1] = {local};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(vv, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = vv[0];
      if (FINALIZE) st->alfa = vv[0];
    }
  }
}

// r2new = t - (alfa/beta) r2 ; ynew = dinv .* r2new ; sums[0] = r2new.ynew
template <bool FINALIZE>
__global__ void __launch_bounds__(256)
k_minres_r(int64_t n, double* __restrict__ t /* becomes r2new */, const double* __restrict__ r2,
           const double* __restrict__ dinv, double* __restrict__ ynew, KspState* st, double* hist,
           double* scratch) {
  if (st->done) return;
  __shared__ double s_red[33];
  const double c = st->alfa / st->mbeta;
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = t[i] - c * r2[i];
    t[i] = ri;
    const double yi = dinv ? dinv[i] * ri : ri;
    ynew[i] = yi;
    s += ri * yi;
  }
  double v[1] = {s};
  const int ops[1] = {RED_SUM};
  if (grid_reduce<1>(v, ops, scratch, s_red)) {
    if (threadIdx.x == 0) {
      st->sums[0] = v[0];
      if (FINALIZE) minres_after_beta(st, hist, 0);
    }
  }
}

// wnew = (v - oldeps w1 - delta w2) / gamma ; x += phi wnew ; then the convergence test
__global__ void __launch_bounds__(256)
k_minres_wx(int64_t n, const double* __restrict__ v, const double* __restrict__ w1 /* also wnew out */,
            const double* __restrict__ w2, double* __restrict__ wnew, double* __restrict__ x,
            KspState* st, unsigned int* ticket) {
  if (st->done) return;
  const double oldeps = st->oldeps, delta = st->delta, inv_gamma = 1.0 / st->gamma, phi = st->phi;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double wi = (v[i] - oldeps * w1[i] - delta * w2[i]) * inv_gamma;
    wnew[i] = wi;
    x[i] += phi * wi;
  }
  // last block to finish raises `done` if the recurrence says so (all blocks have read st by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) converged_default(st, st->rnorm, st->its);
  }
}

// x = dinv .* b (preonly + jacobi: DG0 projections, cashocs/geometry/mesh_testing.py:74-81)
__global__ void __launch_bounds__(256)
k_diag_solve(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv,
             double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = dinv[i] * b[i];
}

int diag_inv_launch(cudaStream_t s, const cb_mat* A, double* dinv, double* lmax, double* scratch);

struct Solver {
  cudaStream_t s;
  const cb_mat* A;
  cb_dist* dist;
  int64_t n, nc;
  int bs;
  KspState* st;
  double* hist_dev;
  double* scratch;
  double* sendbuf;
  int vgrid, sgrid, gsplit = 1;
  // optional live profiling of the SpMV launches (bench.py roofline): CUDA events around every
  // profile_every-th launch on the launching stream
  int profile_every = 0, spmv_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof;

  const double* flex_w = nullptr;  // flexible beta: the fused dots of the iteration also accumulate z.w
  bool capturing = false;  // inside a CUDA-graph capture: no event timing
  bool prof_begin() {
    const bool on = !capturing && profile_every > 0 && (spmv_launches % profile_every) == 0 && prof.size() < 1024;
    ++spmv_launches;
    if (!on) return false;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return false;
    prof.push_back({a, b});
    cudaEventRecord(a, s);
    return true;
  }
  void prof_end(bool on) {
    if (on) cudaEventRecord(prof.back().second, s);
  }
  double prof_collect() {  // after a stream synchronisation
    double total = 0.0;
    for (auto& e : prof) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) total += ms;
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    prof.clear();
    return total;
  }

  template <int BS>
  int cg_spmv_bs(const double* p, double* w) {
    if (dist) k_cg_spmv<BS, false><<<sgrid, SpmvCfg<BS>::THREADS, SpmvCfg<BS>::SMEM_BYTES, s>>>(A->nrows, gsplit, A->rowptr, A->col, A->val, p, w, st, scratch);
    else k_cg_spmv<BS, true><<<sgrid, SpmvCfg<BS>::THREADS, SpmvCfg<BS>::SMEM_BYTES, s>>>(A->nrows, gsplit, A->rowptr, A->col, A->val, p, w, st, scratch);
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  int cg_spmv(double* p, double* w) {
    if (dist) { if (int e = dist_halo_exchange(s, dist, p, bs, sendbuf)) return e; }
    const bool pr = prof_begin();
    int e = bs == 1 ? cg_spmv_bs<1>(p, w) : bs == 2 ? cg_spmv_bs<2>(p, w) : cg_spmv_bs<3>(p, w);
    prof_end(pr);
    if (e) return e;
    return reduce_and_step(1, 2, false);
  }
  template <int BS>
  int minres_spmv_bs(const double* v, const double* r1, double* t) {
    if (dist) k_minres_spmv<BS, false><<<sgrid, SpmvCfg<BS>::THREADS, SpmvCfg<BS>::SMEM_BYTES, s>>>(A->nrows, gsplit, A->rowptr, A->col, A->val, v, r1, t, st, scratch);
    else k_minres_spmv<BS, true><<<sgrid, SpmvCfg<BS>::THREADS, SpmvCfg<BS>::SMEM_BYTES, s>>>(A->nrows, gsplit, A->rowptr, A->col, A->val, v, r1, t, st, scratch);
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  int plain_spmv(double* x, double* y) {
    if (dist) { if (int e = dist_halo_exchange(s, dist, x, bs, sendbuf)) return e; }
    return spmv_dispatch(s, A, x, y, nullptr, nullptr, nullptr, &st->done);
  }

  static void cheb_coefficients(double lmax, double ratio, int degree, double& inv_theta,
                                std::pair<double, double>* coef) {
    const double lmin = lmax / ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
    const double sigma = theta / delta;
    double rho = 1.0 / sigma;
    inv_theta = 1.0 / theta;
    for (int j = 1; j < degree; ++j) {
      const double rho_new = 1.0 / (2.0 * sigma - rho);
      coef[j] = {rho_new * rho, 2.0 * rho_new / delta};
      rho = rho_new;
    }
  }

  // t_l = A_l x_l (ghost update of x_l first on a partitioned mesh)
  // L.t = A_l x  -- or, with `residual_of` = the level's right-hand side, L.t = b - A_l x in the same pass (what the
  // restriction and the first post-smoothing step read: one vector instead of two)
  int level_spmv(const cb_amg_level& L, double* xl, const double* residual_of = nullptr) {
    if (L.dist) { if (int e = dist_halo_exchange(s, L.dist, xl, bs, L.sendbuf)) return e; }
    if (residual_of) return spmv_residual_dispatch(s, &L.A, xl, residual_of, L.t, &st->done, /*use32=*/true);
    return spmv_dispatch(s, &L.A, xl, L.t, nullptr, nullptr, nullptr, &st->done, /*use32=*/true);
  }

  // `degree` Chebyshev-Jacobi steps on A_l x = b (zero or current initial guess); local operator
  int amg_smooth(const cb_amg_level& L, const double* bl, double* xl, int degree, double ratio, bool zero_guess,
                 bool first_done = false) {
    const int64_t nl = L.A.nrows * bs;
    int g = grid_for(nl, 256, 4);
    std::pair<double, double> coef[64];
    double inv_theta = 0.0;
    cheb_coefficients(L.lmax, ratio, degree, inv_theta, coef);
    int j0 = 0;
    if (zero_guess) {
      if (!first_done) {  // else: fused into k_cg_xr by the caller
        k_cheb_first<<<g, 256, 0, s>>>(nl, bl, L.dinv, inv_theta, L.d, xl, st);
        CB_LAUNCH_CHECK();
      }
      j0 = 1;
    }
    for (int j = j0; j < degree; ++j) {
      if (int e = level_spmv(L, xl)) return e;
      const double c1 = (j == 0) ? 0.0 : coef[j].first;
      const double c2 = (j == 0) ? inv_theta : coef[j].second;
      k_cheb_step<<<g, 256, 0, s>>>(nl, bl, L.t, L.dinv, c1, c2, L.d, xl, st);
      CB_LAUNCH_CHECK();
    }
    return CB_OK;
  }

  // the residual pass of the V-cycle writes lev.t = b - A x (scalar-prolongator hierarchies): restriction and the
  // A*P post-smoothing step then read one vector instead of two
  bool t_residual = false;
  template <int BS, typename VT>
  int amg_post_bs(const cb_mat* AP, const VT* val, const double* e, double* z, const double* b, const double* t,
                  const double* dinv, double c2, double* dvec, int dots_mode) {
    using C = SpmvCfg<BS, VT>;
    const int G = spmv_sanitize_split(AP->row_split);
    // short rows (A*P): more rows per CTA keep the bulk copies as large as those of A (cb_mat.cta_warps)
    int warps = AP->cta_warps > 0 ? AP->cta_warps : C::WARPS;
    if (warps < C::WARPS) warps = C::WARPS;
    if (warps > 2 * C::WARPS) warps = 2 * C::WARPS;
    const int threads = 32 * warps;
    const int R = (32 / (BS * G)) * warps;
    const int64_t nrb = (AP->nrows + R - 1) / R;
    int64_t cap = (int64_t)num_sms() * C::CTAS_PER_SM;
    if (cap > RED_MAXP) cap = RED_MAXP;
    const int grid = (int)(nrb < cap ? nrb : cap);
    static bool configured = false;
    if (!configured) {
      const int co = (int)cudaSharedmemCarveoutMaxShared;
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_amg_post<BS, true, VT>, cudaFuncAttributePreferredSharedMemoryCarveout, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_amg_post<BS, false, VT>, cudaFuncAttributePreferredSharedMemoryCarveout, co));
      configured = true;
    }
    if (dots_mode > -2)
      k_amg_post<BS, true, VT><<<grid, threads, C::SMEM_BYTES, s>>>(AP->nrows, G, AP->rowptr, AP->col, val, e, z, b, t,
                                                                     dinv, c2, dvec, st, hist_dev, scratch, dots_mode, flex_w, t_residual ? 1 : 0);
    else
      k_amg_post<BS, false, VT><<<grid, threads, C::SMEM_BYTES, s>>>(AP->nrows, G, AP->rowptr, AP->col, val, e, z, b, t,
                                                                      dinv, c2, dvec, st, hist_dev, scratch, -1, nullptr, t_residual ? 1 : 0);
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  int amg_post(const cb_mat* AP, const double* e, double* z, const double* b, const double* t, const double* dinv,
               double c2, double* dvec, int dots_mode) {
    if (AP->val32)
      return AP->bs == 1 ? amg_post_bs<1, float>(AP, AP->val32, e, z, b, t, dinv, c2, dvec, dots_mode)
           : AP->bs == 2 ? amg_post_bs<2, float>(AP, AP->val32, e, z, b, t, dinv, c2, dvec, dots_mode)
                         : amg_post_bs<3, float>(AP, AP->val32, e, z, b, t, dinv, c2, dvec, dots_mode);
    return AP->bs == 1 ? amg_post_bs<1, double>(AP, AP->val, e, z, b, t, dinv, c2, dvec, dots_mode)
         : AP->bs == 2 ? amg_post_bs<2, double>(AP, AP->val, e, z, b, t, dinv, c2, dvec, dots_mode)
                       : amg_post_bs<3, double>(AP, AP->val, e, z, b, t, dinv, c2, dvec, dots_mode);
  }

  // Fusion hooks of the V-cycle with the PCG loop around it:
  //  * level-0 pre-smoothing from a zero guess is z = dinv .* r / theta -- k_cg_xr can write it while it
  //    updates r (amg_presmooth_fusable / presmoothed);
  //  * with one post-smoothing step through A*P the last kernel of the cycle touches every z_i, so it
  //    also accumulates z.r and z.z (dots_mode > -2; dots_fused reports whether that happened).
  bool dots_fused = false;
  bool amg_presmooth_fusable(const cb_amg* h, const double** dinv0, double* scale, double** d0) const {
    if (h->nlevels < 2) return false;
    const cb_amg_level& lev = h->levels[0];
    std::pair<double, double> coef[64];
    double inv_theta = 0.0;
    cheb_coefficients(lev.lmax, h->smooth_ratio, h->smooth_degree, inv_theta, coef);
    *dinv0 = lev.dinv;
    *scale = inv_theta;
    *d0 = h->smooth_degree > 1 ? lev.d : nullptr;
    return true;
  }

  // Chebyshev degree of the pre / post smoother of level l: the coarse levels (a tenth of the fine level's bytes) may
  // smooth harder than the fine one
  static int level_degree(const cb_amg* h, int l) {
    return (l > 0 && h->level_smooth_degree > 0) ? h->level_smooth_degree : h->smooth_degree;
  }

  // ---- levels >= fused_first as one persistent kernel (coarse_cycle.cuh)
  int fused_first = 0;                 // 0: off
  CoarseCycleDev* fused_dev = nullptr; // device copy of the descriptor (workspace)
  int fused_grid = 1;

  // Decide which levels the persistent kernel takes and upload its descriptor (once per solve, outside the graph).
  // `buf` is a workspace slice of cc_workspace_doubles() doubles: descriptor, then the three barrier words.
  static int64_t cc_workspace_doubles() { return (int64_t)((sizeof(CoarseCycleDev) + 7) / 8 + 8); }
  int setup_fused_cycle(const cb_amg* h, double* buf) {
    fused_first = 0;
    const char* env = getenv("CASHOCS_B200_FUSED_COARSE");
    if (env && env[0] == '0') return CB_OK;
    const int L = h->nlevels;
    if (L < 2 || h->p_block) return CB_OK;
    // levels with more rows than this keep their own (TMA-staged) kernels: they are bandwidth-, not latency-bound
    int64_t max_rows = 20000;
    if (const char* mr = getenv("CASHOCS_B200_FUSED_MAX_ROWS")) max_rows = atoll(mr);
    // A level joins the persistent kernel if it is local / replicated, or distributed with the one-sided transport on
    // (its ghost updates and the all-gather at the agglomeration point then run inside the kernel).
    auto level_ok = [&](int l) -> bool {
      const cb_amg_level& lev = h->levels[l];
      if (lev.A.nrows > max_rows) return false;
      if (lev.dist) {
        if (!dist_is_one_sided(lev.dist)) return false;
        if (dist_recv_len(lev.dist) * bs > dist_ghost_capacity(lev.dist)) return false;
      } else if (lev.A.ncols != lev.A.nrows) {
        return false;
      }
      if (l < L - 1) {
        if (!lev.AP.val || lev.AP.val32) return false;
        const cb_amg_level& nxt = h->levels[l + 1];
        if (nxt.rep_dist) {  // agglomeration point: needs the one-sided all-gather plan and the coarse numbering map
          if (!nxt.rep_gather || !dist_is_one_sided(nxt.rep_gather) || !nxt.rep_tmp || !lev.coarse_gid) return false;
          if (dist_recv_len(nxt.rep_gather) * bs > dist_ghost_capacity(nxt.rep_gather)) return false;
        }
      }
      return true;
    };
    int first = L;
    for (int l = L - 1; l >= 1; --l) {
      if (!level_ok(l)) break;
      first = l;
    }
    if (first >= L || L - first > CC_MAX_LEVELS) return CB_OK;
    CoarseCycleDev cy;
    memset(&cy, 0, sizeof(cy));
    cy.nlevels = L - first;
    const cb_amg_level& lastlev = h->levels[L - 1];
    cy.dense_n = h->coarse_inv ? (int)(lastlev.A.nrows * bs) : 0;
    cy.coarse_inv = h->coarse_inv;
    for (int l = first; l < L; ++l) {
      const cb_amg_level& lev = h->levels[l];
      CoarseLevelDev& d = cy.lev[l - first];
      const bool last = (l == L - 1);
      d.n = lev.A.nrows;
      d.nc = last ? 0 : lev.nc;
      d.rowptr = lev.A.rowptr; d.col = lev.A.col; d.val = lev.A.val; d.dinv = lev.dinv;
      d.p_rowptr = lev.p_rowptr; d.p_col = lev.p_col; d.p_val = lev.p_val;
      d.pt_rowptr = lev.pt_rowptr; d.pt_col = lev.pt_col; d.pt_val = lev.pt_val;
      d.ap_rowptr = lev.AP.rowptr; d.ap_col = lev.AP.col; d.ap_val = lev.AP.val;
      d.x = lev.x; d.b = lev.b; d.t = lev.t; d.d = lev.d;
      d.degree = (last && !h->coarse_inv) ? h->coarse_degree : level_degree(h, l);
      if (d.degree > CC_MAX_DEGREE) return CB_OK;
      std::pair<double, double> coef[64];
      cheb_coefficients(lev.lmax, (last && !h->coarse_inv) ? h->coarse_ratio : h->smooth_ratio, d.degree, d.inv_theta, coef);
      for (int j = 1; j < d.degree; ++j) { d.c1[j] = coef[j].first; d.c2[j] = coef[j].second; }
      if (lev.dist) {
        d.has_halo = 1;
        d.halo = dist_halo_dev(lev.dist);
      }
      if (!last && h->levels[l + 1].rep_dist) {
        const cb_amg_level& nxt = h->levels[l + 1];
        d.next_rep = 1;
        d.gather = dist_halo_dev(nxt.rep_gather);
        d.rep_tmp = nxt.rep_tmp;
        d.rep_off = nxt.rep_off;
        d.coarse_gid = lev.coarse_gid;
      }
    }
    fused_dev = reinterpret_cast<CoarseCycleDev*>(buf);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(buf + (sizeof(CoarseCycleDev) + 7) / 8);
    cy.barrier = bar;
    CB_CUDA_CHECK(cudaMemsetAsync(bar, 0, 8 * sizeof(double), s));
    CB_CUDA_CHECK(cudaMemcpyAsync(fused_dev, &cy, sizeof(cy), cudaMemcpyHostToDevice, s));
    CB_CUDA_CHECK(cudaStreamSynchronize(s));  // `cy` lives on this stack frame
    const int64_t rows = cy.lev[0].n;
    int64_t g = (rows + 31) / 32;  // 32 warps per CTA, one block row per warp and sweep
    if (g < 1) g = 1;
    if (g > num_sms()) g = num_sms();
    fused_grid = (int)g;
    fused_first = first;
    return CB_OK;
  }
  int launch_fused_cycle() {
    if (bs == 1) k_coarse_cycle<1><<<fused_grid, CC_THREADS, 0, s>>>(fused_dev, &st->done);
    else if (bs == 2) k_coarse_cycle<2><<<fused_grid, CC_THREADS, 0, s>>>(fused_dev, &st->done);
    else k_coarse_cycle<3><<<fused_grid, CC_THREADS, 0, s>>>(fused_dev, &st->done);
    CB_LAUNCH_CHECK();
    return CB_OK;
  }

  // z = M^-1 r : one symmetric V-cycle (pre/post Chebyshev smoothing, Chebyshev coarsest solve)
  int amg_vcycle(const cb_amg* h, const double* r, double* z, bool presmoothed = false, int dots_mode = -2) {
    const int L = h->nlevels;
    dots_fused = false;
    t_residual = !h->p_block;
    int up_from = L - 2;  // first level of the upward sweep
    for (int l = 0; l < L; ++l) {
      const cb_amg_level& lev = h->levels[l];
      const double* bl = (l == 0) ? r : lev.b;
      double* xl = (l == 0) ? z : lev.x;
      if (fused_first > 0 && l == fused_first) {  // everything from here down and back up is one kernel
        if (int e = launch_fused_cycle()) return e;
        up_from = l - 1;
        break;
      }
      if (l == L - 1) {
        if (h->coarse_inv) {
          if (int e = amg_dense_apply(s, (int)(lev.A.nrows * bs), h->coarse_inv, bl, xl, &st->done)) return e;
        } else {
          if (int e = amg_smooth(lev, bl, xl, h->coarse_degree, h->coarse_ratio, true)) return e;
        }
        break;
      }
      if (int e = amg_smooth(lev, bl, xl, level_degree(h, l), h->smooth_ratio, true, l == 0 && presmoothed)) return e;
      if (int e = level_spmv(lev, xl, t_residual ? bl : nullptr)) return e;
      const double* rb = t_residual ? lev.t : bl;       // what the restriction reads: the residual itself ...
      const double* rt = t_residual ? nullptr : lev.t;  // ... or b and A x
      const cb_amg_level& nxt = h->levels[l + 1];
      if (nxt.rep_dist) {
        // agglomeration: the next level is replicated on every rank; this rank contributes its
        // rows of the restricted residual, one all-reduce assembles the global vector everywhere
        if (nxt.rep_gather && dist_is_one_sided(nxt.rep_gather)) {
          // every coarse row is produced by exactly one rank: one-sided all-gather (one exchange kernel) + unpack
          if (int e = amg_restrict(s, bs, &lev, rb, rt, nxt.rep_tmp, &st->done, h->p_block)) return e;
          if (int e = dist_halo_exchange(s, nxt.rep_gather, nxt.rep_tmp, bs, nullptr)) return e;
          const int64_t n_own = lev.nc;
          k_unpack_allgather<<<grid_for(nxt.A.nrows * bs, 256, 4), 256, 0, s>>>(nxt.A.nrows, n_own, nxt.rep_off, bs,
                                                                              nxt.rep_tmp, nxt.b, st);
          CB_LAUNCH_CHECK();
        } else {
          const size_t len = sizeof(double) * (size_t)(nxt.A.nrows * bs);
          CB_CUDA_CHECK(cudaMemsetAsync(nxt.b, 0, len, s));
          if (int e = amg_restrict(s, bs, &lev, rb, rt, nxt.b + nxt.rep_off * bs, &st->done, h->p_block)) return e;
          if (int e = dist_allreduce_sum(s, nxt.rep_dist, nxt.b, (int)(nxt.A.nrows * bs))) return e;
        }
      } else {
        if (int e = amg_restrict(s, bs, &lev, rb, rt, nxt.b, &st->done, h->p_block)) return e;
      }
    }
    for (int l = up_from; l >= 0; --l) {
      const cb_amg_level& lev = h->levels[l];
      const cb_amg_level& nxt = h->levels[l + 1];
      const double* bl = (l == 0) ? r : lev.b;
      double* xl = (l == 0) ? z : lev.x;
      const double* xc = nxt.rep_dist ? nxt.x + nxt.rep_off * bs : nxt.x;
      if (int e = amg_prolong(s, bs, &lev, xc, xl, &st->done, h->p_block)) return e;
      if (lev.AP.val) {
        // first post-smoothing step through A P instead of A (see k_spmv_post): needs the coarse
        // correction in this rank's coarse numbering (owned + ghost coarse nodes)
        const double* xcl = nxt.x;
        if (nxt.rep_dist) {
          k_gather_nodes<<<grid_for(lev.AP.ncols * bs, 256, 4), 256, 0, s>>>(lev.AP.ncols, bs, lev.coarse_gid, nxt.x,
                                                                          lev.xc_local, st);
          CB_LAUNCH_CHECK();
          xcl = lev.xc_local;
        } else if (nxt.dist) {
          if (int e = dist_halo_exchange(s, nxt.dist, nxt.x, bs, nxt.sendbuf)) return e;
        }
        const int degree = level_degree(h, l);
        std::pair<double, double> coef[64];
        double inv_theta = 0.0;
        cheb_coefficients(lev.lmax, h->smooth_ratio, degree, inv_theta, coef);
        const bool fuse_dots = (l == 0 && degree == 1 && dots_mode > -2);
        if (int e = amg_post(&lev.AP, xcl, xl, bl, lev.t, lev.dinv, inv_theta, degree > 1 ? lev.d : nullptr,
                             fuse_dots ? dots_mode : -2)) return e;
        if (fuse_dots) dots_fused = true;
        const int64_t nl = lev.A.nrows * bs;
        for (int j = 1; j < degree; ++j) {
          if (int e = level_spmv(lev, xl)) return e;
          k_cheb_step<<<grid_for(nl, 256, 4), 256, 0, s>>>(nl, bl, lev.t, lev.dinv, coef[j].first, coef[j].second,
                                                          lev.d, xl, st);
          CB_LAUNCH_CHECK();
        }
      } else {
        if (int e = amg_smooth(lev, bl, xl, level_degree(h, l), h->smooth_ratio, false)) return e;
      }
    }
    return CB_OK;
  }
  int reduce_and_step(int nsums, int which, bool minres) {
    if (!dist) return CB_OK;
    if (dist_is_one_sided(dist) && nsums <= RED_W) {
      k_allreduce_state<<<1, 32, 0, s>>>(nsums, st, hist_dev, dist_p2p_view(dist), which, minres ? 1 : 0);
      CB_LAUNCH_CHECK();
      return CB_OK;
    }
    if (int e = dist_allreduce_sum(s, dist, st->sums, nsums)) return e;
    if (minres) k_minres_state<<<1, 1, 0, s>>>(st, hist_dev, which);
    else k_state_step<<<1, 1, 0, s>>>(st, hist_dev, which);
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
};

}  // namespace cb

using namespace cb;

extern "C" int64_t cb_ksp_workspace_len(int64_t n, int64_t nc, const cb_ksp_opts* opts) {
  (void)opts;
  // r, w, dinv, d, t : n ; z, p, (minres: 3 more with ghosts) : nc ; + state, scratch, send buffer
  return 6 * n + 6 * nc + STATE_DOUBLES + RED_SCRATCH_LEN + nc + 64 + Solver::cc_workspace_doubles();
}

extern "C" int cb_ksp_solve(void* stream, const cb_mat* A, const double* b, double* x,
                            const cb_ksp_opts* o, cb_dist* dist, double* work, int64_t work_len,
                            cb_ksp_result* result, double* hist, int64_t hist_len) {
  CB_REQUIRE(A && b && x && o && work && result, "cb_ksp_solve: missing argument");
  CB_REQUIRE(A->bs >= 1 && A->bs <= 3, "cb_ksp_solve: block size must be 1..3");
  const int64_t n = A->nrows * A->bs, nc = A->ncols * A->bs;
  CB_REQUIRE(work_len >= cb_ksp_workspace_len(n, nc, o), "cb_ksp_solve: workspace too small");
  CB_REQUIRE(o->pc_type != CB_PC_CHEBYSHEV || o->cheb_degree >= 1, "chebyshev degree must be >= 1");
  cudaStream_t s = as_stream(stream);
  configure_scratch_pool();

  // carve the workspace
  double* wp = work;
  auto take = [&](int64_t len) { double* p = wp; wp += (len + 1) & ~int64_t(1); return p; };
  KspState* st = reinterpret_cast<KspState*>(take(STATE_DOUBLES));
  double* scratch = take(RED_SCRATCH_LEN);
  double* dinv = take(n);
  double* r = take(n);
  double* w = take(n);
  double* cd = take(n);
  double* ct = take(n);
  double* z = take(nc);
  double* p = take(nc);
  double* m1 = take(nc);
  double* m2 = take(nc);
  double* m3 = take(nc);
  double* sendbuf = take(nc);
  double* cc_buf = take(Solver::cc_workspace_doubles());
  double* hist_dev = nullptr;
  const int64_t hlen = (o->monitor && hist && hist_len > 0) ? hist_len : 0;
  if (hlen) CB_CUDA_CHECK(cudaMallocAsync(&hist_dev, sizeof(double) * hlen, s));
  Solver S;  // declared before the guard: the guard's destructor (which runs first) still sees S.prof
  // everything this call acquires is released on EVERY return path (error returns included)
  struct Cleanup {
    cudaStream_t stream;
    double* hist_dev = nullptr;
    cudaGraphExec_t gexec = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>>* events = nullptr;
    ~Cleanup() {
      if (gexec) cudaGraphExecDestroy(gexec);
      if (hist_dev) cudaFreeAsync(hist_dev, stream);
      if (events) {
        for (auto& e : *events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        events->clear();
      }
    }
  } cleanup{s};
  cleanup.hist_dev = hist_dev;

  KspState h0;
  memset(&h0, 0, sizeof(h0));
  h0.rtol = o->rtol; h0.atol = o->atol; h0.dtol = o->dtol; h0.max_it = o->max_it;
  h0.hist_len = (int)hlen;
  h0.flexible = (o->flexible != 0 && o->ksp_type == CB_KSP_CG && (o->pc_type == CB_PC_AMG || o->pc_type == CB_PC_CHEBYSHEV)) ? 1 : 0;
  CB_CUDA_CHECK(cudaMemcpyAsync(st, &h0, sizeof(h0), cudaMemcpyHostToDevice, s));
  CB_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(double) * RED_SCRATCH_LEN, s));

  S.s = s; S.A = A; S.dist = dist; S.n = n; S.nc = nc; S.bs = A->bs; S.st = st;
  S.hist_dev = hist_dev; S.scratch = scratch; S.sendbuf = sendbuf;
  S.profile_every = o->profile_every;
  cleanup.events = &S.prof;
  S.vgrid = grid_for(n, 256, 4);
  if (S.vgrid > RED_MAXP) S.vgrid = RED_MAXP;
  {
    S.gsplit = spmv_sanitize_split(A->row_split);
    const int R = A->bs == 1 ? spmv_rows_per_cta<1>(S.gsplit) : A->bs == 2 ? spmv_rows_per_cta<2>(S.gsplit) : spmv_rows_per_cta<3>(S.gsplit);
    const int per_sm = A->bs == 1 ? SpmvCfg<1>::CTAS_PER_SM : A->bs == 2 ? SpmvCfg<2>::CTAS_PER_SM : SpmvCfg<3>::CTAS_PER_SM;
    const int64_t nrb = (A->nrows + R - 1) / R;
    int64_t cap = (int64_t)num_sms() * per_sm;
    if (cap > RED_MAXP) cap = RED_MAXP;
    S.sgrid = (int)(nrb < cap ? nrb : cap);
    static bool configured = false;
    if (!configured) {
      const int co = (int)cudaSharedmemCarveoutMaxShared;
      const auto attr = cudaFuncAttributePreferredSharedMemoryCarveout;
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_cg_spmv<1, true>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_cg_spmv<2, true>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_cg_spmv<3, true>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_cg_spmv<1, false>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_cg_spmv<2, false>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_cg_spmv<3, false>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_minres_spmv<1, true>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_minres_spmv<2, true>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_minres_spmv<3, true>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_minres_spmv<1, false>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_minres_spmv<2, false>, attr, co));
      CB_CUDA_CHECK(cudaFuncSetAttribute(k_minres_spmv<3, false>, attr, co));
      configured = true;
    }
  }
  const int vg = S.vgrid;
  const bool use_pc = o->pc_type != CB_PC_NONE;
  double lmax_h = 0.0;
  if (use_pc) {
    double* lmax_dev = (o->pc_type == CB_PC_CHEBYSHEV) ? &st->sums[3] : nullptr;
    if (o->pc_type != CB_PC_AMG) {
      if (int e = diag_inv_launch(s, A, dinv, lmax_dev, scratch)) return e;
    }
    if (lmax_dev) {
      if (dist) { if (int e = cb_dist_allreduce(stream, dist, lmax_dev, 1, 2)) return e; }
      CB_CUDA_CHECK(cudaMemcpyAsync(&lmax_h, lmax_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
      CB_CUDA_CHECK(cudaStreamSynchronize(s));
    }
  }
  const double* dinv_or_null = use_pc ? dinv : nullptr;

  KspState hs;
  auto poll = [&]() -> int {
    CB_CUDA_CHECK(cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost, s));
    CB_CUDA_CHECK(cudaStreamSynchronize(s));
    return CB_OK;
  };
  int rc = CB_OK;
  const int check_every = o->check_every > 0 ? o->check_every : (o->pc_type == CB_PC_AMG ? 4 : 25);

  if (o->ksp_type == CB_KSP_PREONLY) {
    CB_REQUIRE(o->pc_type == CB_PC_JACOBI, "preonly supports pc_type jacobi only (no direct solver on the device)");
    k_diag_solve<<<vg, 256, 0, s>>>(n, b, dinv, x);
    CB_LAUNCH_CHECK();
    CB_CUDA_CHECK(cudaStreamSynchronize(s));
    result->reason = CB_KSP_CONVERGED_ITS; result->its = 1; result->rnorm = 0.0; result->rnorm0 = 0.0;
    result->spmv_ms = 0.0; result->spmv_profiled = 0; result->spmv_launches = 0;
    return CB_OK;
  }

  if (o->ksp_type == CB_KSP_CG) {
    const bool amg = o->pc_type == CB_PC_AMG;
    const bool cheb = o->pc_type == CB_PC_CHEBYSHEV || amg;  // "general" (non-fused) preconditioner path
    if (amg) {
      CB_REQUIRE(o->amg && o->amg->nlevels >= 1 && o->amg->levels, "pc_type amg needs a hierarchy (cb_ksp_opts.amg)");
      CB_REQUIRE(o->amg->smooth_degree >= 1 && o->amg->smooth_degree <= 64 && o->amg->coarse_degree >= 1 &&
                 o->amg->coarse_degree <= 64 && o->amg->level_smooth_degree >= 0 && o->amg->level_smooth_degree <= 64,
                 "amg smoother degrees must be in 1..64");
    }
    if (amg) { if ((rc = S.setup_fused_cycle(o->amg, cc_buf))) return rc; }
    // Chebyshev coefficients (fixed polynomial -> a linear SPD preconditioner)
    std::pair<double, double> coef[64];
    double inv_theta = 0.0;
    if (cheb && !amg) {
      CB_REQUIRE(o->cheb_degree <= 64, "chebyshev degree > 64");
      const double ratio = o->cheb_ratio > 1.0 ? o->cheb_ratio : 30.0;
      const double lmax = lmax_h, lmin = lmax / ratio;
      const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
      const double sigma = theta / delta;
      double rho = 1.0 / sigma;
      inv_theta = 1.0 / theta;
      for (int j = 1; j < o->cheb_degree; ++j) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        coef[j] = {rho_new * rho, 2.0 * rho_new / delta};
        rho = rho_new;
      }
    }
    const double* pre_dinv = nullptr;
    double pre_scale = 0.0;
    double* pre_d = nullptr;
    const bool pre_fuse = amg && S.amg_presmooth_fusable(o->amg, &pre_dinv, &pre_scale, &pre_d);
    bool presmoothed = false;
    int dots_mode = -2;
    const bool rbm = amg && o->rbm_coords != nullptr;
    const bool flexible = cheb && o->flexible != 0;
    if (rbm) CB_REQUIRE(S.bs >= 2 && o->rbm_einv && o->rbm_work, "rigid-body correction: bs must be 2 or 3, einv / work set");
    auto apply_cheb = [&]() -> int {  // z = P(r)
      if (rbm) {
        // V-cycle without the fused dots (z still changes), then z += Z (Z^T A Z)^-1 Z^T r with the dots
        if (int e = S.amg_vcycle(o->amg, r, z, presmoothed, -2)) return e;
        const int64_t nn = n / S.bs;
        if (S.bs == 3) k_rbm_project<3><<<vg, 256, 0, s>>>(nn, r, o->rbm_coords, o->rbm_work, scratch, st);
        else k_rbm_project<2><<<vg, 256, 0, s>>>(nn, r, o->rbm_coords, o->rbm_work, scratch, st);
        CB_LAUNCH_CHECK();
        if (dist) { if (int e = dist_allreduce_sum(s, dist, o->rbm_work, S.bs == 3 ? 6 : 3)) return e; }
        if (S.bs == 3) k_rbm_correct<3><<<vg, 256, 0, s>>>(nn, r, z, o->rbm_coords, o->rbm_einv, o->rbm_work, st, hist_dev,
                                                        scratch, dots_mode, S.flex_w);
        else k_rbm_correct<2><<<vg, 256, 0, s>>>(nn, r, z, o->rbm_coords, o->rbm_einv, o->rbm_work, st, hist_dev, scratch,
                                                 dots_mode, S.flex_w);
        CB_LAUNCH_CHECK();
        S.dots_fused = true;
        return CB_OK;
      }
      if (amg) return S.amg_vcycle(o->amg, r, z, presmoothed, dots_mode);
      k_cheb_first<<<vg, 256, 0, s>>>(n, r, dinv, inv_theta, cd, z, st);
      CB_LAUNCH_CHECK();
      for (int j = 1; j < o->cheb_degree; ++j) {
        if (int e = S.plain_spmv(z, ct)) return e;
        k_cheb_step<<<vg, 256, 0, s>>>(n, r, ct, dinv, coef[j].first, coef[j].second, cd, z, st);
        CB_LAUNCH_CHECK();
      }
      return CB_OK;
    };
    CB_CUDA_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * nc, s));
    if (cheb) CB_CUDA_CHECK(cudaMemsetAsync(z, 0, sizeof(double) * nc, s));
    if (!cheb) {
      if (dist) k_cg_init<false><<<vg, 256, 0, s>>>(n, b, dinv_or_null, r, z, x, true, st, hist_dev, scratch);
      else k_cg_init<true><<<vg, 256, 0, s>>>(n, b, dinv_or_null, r, z, x, true, st, hist_dev, scratch);
      CB_LAUNCH_CHECK();
      if ((rc = S.reduce_and_step(2, 0, false))) return rc;
    } else {
      k_cg_init<true><<<vg, 256, 0, s>>>(n, b, dinv_or_null, r, z, x, false, st, hist_dev, scratch);
      CB_LAUNCH_CHECK();
      dots_mode = amg ? (dist ? -1 : 0) : -2;
      if ((rc = apply_cheb())) return rc;
      if (!S.dots_fused) {
        k_dots_zr<<<vg, 256, 0, s>>>(n, z, r, st, hist_dev, scratch, dist ? -1 : 0);
        CB_LAUNCH_CHECK();
      }
      if ((rc = S.reduce_and_step(2, 0, false))) return rc;
    }
    // one PCG iteration; every argument is iteration-independent, so the body can be captured once
    // into a CUDA graph and replayed (the one-sided halo / all-reduce kernels keep their epochs in
    // device memory for exactly that reason)
    auto iteration = [&]() -> int {
      S.flex_w = flexible ? w : nullptr;  // (not during the initial application: w is not defined yet)
      k_cg_p<<<vg, 256, 0, s>>>(n, z, p, st);
      CB_LAUNCH_CHECK();
      if (int e = S.cg_spmv(p, w)) return e;
      if (!cheb) {
        if (dist) k_cg_xr<false><<<vg, 256, 0, s>>>(n, p, w, dinv_or_null, x, r, z, true, st, hist_dev, scratch);
        else k_cg_xr<true><<<vg, 256, 0, s>>>(n, p, w, dinv_or_null, x, r, z, true, st, hist_dev, scratch);
        CB_LAUNCH_CHECK();
        if (int e = S.reduce_and_step(2, 1, false)) return e;
      } else {
        k_cg_xr<false><<<vg, 256, 0, s>>>(n, p, w, dinv_or_null, x, r, z, false, st, hist_dev, scratch,
                                          pre_fuse ? pre_dinv : nullptr, pre_scale, pre_d);
        CB_LAUNCH_CHECK();
        presmoothed = pre_fuse;
        dots_mode = amg ? (dist ? -1 : 1) : -2;
        if (int e = apply_cheb()) return e;
        if (!S.dots_fused) {
          k_dots_zr<<<vg, 256, 0, s>>>(n, z, r, st, hist_dev, scratch, dist ? -1 : 1, S.flex_w);
          CB_LAUNCH_CHECK();
        }
        if (int e = S.reduce_and_step(flexible ? 3 : 2, 1, false)) return e;
      }
      return CB_OK;
    };
    // graph replay needs the one-sided transport on partitioned meshes (its exchanges are plain kernels with
    // device-side epochs); with the NCCL send/recv fallback the loop stays eager unless forced (use_graph > 0)
    const bool want_graph = o->use_graph > 0 || (o->use_graph == 0 && (!dist || dist_is_one_sided(dist)));
    // chunks run eagerly before the capture: warm-up (lazy kernel attributes, NCCL buffers) and, when
    // the SpMV launches are being timed with events, enough of them for a sample
    const int eager_chunks = o->profile_every > 0 ? (2 * o->profile_every + check_every - 1) / check_every : 1;
    cudaGraphExec_t gexec = nullptr;
    long long graph_kernels = 0;
    int enq = 0, chunk = 0;
    while (true) {
      if (want_graph && chunk >= eager_chunks) {
        if (!gexec) {
          // capture on a private stream (the caller's may be the legacy default stream, which cannot
          // be captured); the launches go to `s` / `S.s`, so both are redirected for the capture
          cudaGraph_t graph = nullptr;
          const long long before = t_launch_count;
          cudaStream_t user_stream = s, cs = nullptr;
          CB_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
          S.capturing = true;
          s = cs;
          S.s = cs;
          cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
          int brc = CB_OK;
          if (ce == cudaSuccess) {
            brc = iteration();
            // always end the capture, also after a failed iteration(): the thread must not stay in capture mode
            const cudaError_t ee = cudaStreamEndCapture(cs, &graph);
            if (brc == CB_OK) ce = ee;
            else (void)cudaGetLastError();
          }
          s = user_stream;
          S.s = user_stream;
          S.capturing = false;
          cudaStreamDestroy(cs);
          graph_kernels = t_launch_count - before;  // launches of this thread only (captured, not executed)
          t_launch_count = before;
          g_launch_count.fetch_sub(graph_kernels, std::memory_order_relaxed);
          if (brc) { if (graph) cudaGraphDestroy(graph); return brc; }
          CB_CUDA_CHECK(ce);
          const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
          cudaGraphDestroy(graph);
          CB_CUDA_CHECK(ie);
          cleanup.gexec = gexec;
        }
        for (int c = 0; c < check_every && enq < o->max_it; ++c, ++enq) {
          const cudaError_t le = cudaGraphLaunch(gexec, s);
          CB_CUDA_CHECK(le);
          g_launch_count.fetch_add(graph_kernels, std::memory_order_relaxed);
          S.spmv_launches += 1;
        }
      } else {
        for (int c = 0; c < check_every && enq < o->max_it; ++c, ++enq) {
          if ((rc = iteration())) return rc;
        }
      }
      ++chunk;
      if ((rc = poll())) return rc;
      if (hs.done || enq >= o->max_it) break;
    }
  } else if (o->ksp_type == CB_KSP_MINRES) {
    CB_REQUIRE(o->pc_type != CB_PC_CHEBYSHEV, "minres supports pc_type none / jacobi");
    // rotating roles: (r1, r2, y) and (w1, w2, w)
    double* R3[3] = {r, w, ct};       // n each
    double* W3[3] = {m1, m2, m3};     // n each (allocated nc)
    double* v = p;                    // SpMV input (ghost tail)
    double* y = z;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(cd);
    CB_CUDA_CHECK(cudaMemsetAsync(cd, 0, 16, s));
    CB_CUDA_CHECK(cudaMemsetAsync(m1, 0, sizeof(double) * n, s));
    CB_CUDA_CHECK(cudaMemsetAsync(m2, 0, sizeof(double) * n, s));
    CB_CUDA_CHECK(cudaMemsetAsync(m3, 0, sizeof(double) * n, s));
    CB_CUDA_CHECK(cudaMemsetAsync(R3[0], 0, sizeof(double) * n, s));
    CB_CUDA_CHECK(cudaMemsetAsync(v, 0, sizeof(double) * nc, s));
    // r2 = R3[1]
    if (dist) k_minres_init<false><<<vg, 256, 0, s>>>(n, b, dinv_or_null, R3[1], y, x, st, hist_dev, scratch);
    else k_minres_init<true><<<vg, 256, 0, s>>>(n, b, dinv_or_null, R3[1], y, x, st, hist_dev, scratch);
    CB_LAUNCH_CHECK();
    if ((rc = S.reduce_and_step(1, 0, true))) return rc;
    int enq = 0;
    while (true) {
      for (int c = 0; c < check_every && enq < o->max_it; ++c, ++enq) {
        double* r1 = R3[enq % 3];
        double* r2 = R3[(enq + 1) % 3];
        double* t = R3[(enq + 2) % 3];
        double* w1 = W3[enq % 3];
        double* w2 = W3[(enq + 1) % 3];
        double* wn = W3[(enq + 2) % 3];  // slot of the w from three iterations ago (free)
        k_minres_v<<<vg, 256, 0, s>>>(n, y, v, st);
        CB_LAUNCH_CHECK();
        if (dist) { if ((rc = dist_halo_exchange(s, dist, v, S.bs, sendbuf))) return rc; }
        rc = S.bs == 1 ? S.minres_spmv_bs<1>(v, r1, t) : S.bs == 2 ? S.minres_spmv_bs<2>(v, r1, t) : S.minres_spmv_bs<3>(v, r1, t);
        if (rc) return rc;
        if ((rc = S.reduce_and_step(1, 2, true))) return rc;
        if (dist) k_minres_r<false><<<vg, 256, 0, s>>>(n, t, r2, dinv_or_null, y, st, hist_dev, scratch);
        else k_minres_r<true><<<vg, 256, 0, s>>>(n, t, r2, dinv_or_null, y, st, hist_dev, scratch);
        CB_LAUNCH_CHECK();
        if ((rc = S.reduce_and_step(1, 1, true))) return rc;
        k_minres_wx<<<vg, 256, 0, s>>>(n, v, w1, w2, wn, x, st, ticket);
        CB_LAUNCH_CHECK();
      }
      if ((rc = poll())) return rc;
      if (hs.done || enq >= o->max_it) break;
    }
  } else {
    set_error("cb_ksp_solve: unsupported ksp_type %d", o->ksp_type);
    return CB_ERR_UNSUPPORTED;
  }

  if (!hs.done) {  // max_it exhausted without the device raising done (cannot happen, but be safe)
    hs.reason = CB_KSP_DIVERGED_ITS;
  }
  result->reason = hs.reason;
  result->its = hs.its;
  result->rnorm = hs.rnorm;
  result->rnorm0 = hs.rnorm0;
  result->spmv_launches = S.spmv_launches;
  result->spmv_profiled = (int)S.prof.size();
  result->spmv_ms = S.prof_collect();
  if (hist_dev) {
    const int64_t ncopy = (hs.its + 1 < hlen) ? hs.its + 1 : hlen;
    CB_CUDA_CHECK(cudaMemcpyAsync(hist, hist_dev, sizeof(double) * ncopy, cudaMemcpyDeviceToHost, s));
    CB_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return CB_OK;
}
