// The levels >= 1 of the AMG V-cycle as ONE persistent kernel.
//
// Below the fine level a cycle is a chain of ~20 small dependent operations (smoother steps, residuals, restrictions,
// the dense coarsest solve, prolongations, post-smoothing) on operators that hold a tenth of the fine level's bytes and
// less: launched one by one they cost more in launch latency and pipeline drains than in work -- 0.1 ms of a 0.4 ms
// scalar PCG iteration on one B200, most of an iteration on eight.  Here one grid of <= 1 CTA per SM walks through all
// of them; the phases are separated by a grid barrier (an epoch counter in device memory, so a captured CUDA graph
// replays the kernel unchanged), every phase is a warp-per-row gather with a fixed summation order: bit-reproducible.
//
// Handles levels whose operators are local to this rank (one GPU: all of them; partitioned meshes: the replicated
// levels below the agglomeration point) AND distributed levels when the one-sided NVLink transport is on: the ghost
// update of a level's iterate and the all-gather of the restricted residual at the agglomeration point are then phases
// of the same kernel (p2p_halo_exchange_grid: boundary values are stored straight into the neighbours' landing zones,
// flags released by the last CTA, every CTA waits and copies its share) -- compute and collective in one launch, no
// NCCL call, no host involvement.  Scalar prolongator (P (x) I), Chebyshev-Jacobi smoothers of any degree,
// first post-smoothing step through A*P, dense inverse or Chebyshev sweep on the last level -- the same arithmetic,
// operation by operation, as the launch-per-operation path of Solver::amg_vcycle (which stays for distributed levels
// and as the cross-check of the tests, CASHOCS_B200_FUSED_COARSE=0).
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace cb {

constexpr int CC_MAX_LEVELS = 12;
constexpr int CC_MAX_DEGREE = 16;

struct CoarseLevelDev {
  int64_t n;        // block rows of this level
  int64_t nc;       // block rows of the next level (0 on the last level)
  const int64_t* rowptr; const int32_t* col; const double* val;        // A_l
  const double* dinv;
  const int64_t* p_rowptr; const int32_t* p_col; const double* p_val;  // P  : n rows
  const int64_t* pt_rowptr; const int32_t* pt_col; const double* pt_val;  // P^T: nc rows
  const int64_t* ap_rowptr; const int32_t* ap_col; const double* ap_val;  // A*P: n rows (nullable)
  double* x; double* b; double* t; double* d;
  int degree;            // Chebyshev degree of the pre / post smoother (last level without dense inverse: of the solve)
  double inv_theta;      // first step: d = inv_theta * dinv .* b
  double c1[CC_MAX_DEGREE], c2[CC_MAX_DEGREE];  // step j >= 1: d = c1[j] d + c2[j] dinv .* (b - t)
  // ---- partitioned meshes (all zero on one GPU)
  int has_halo;          // distributed level: x carries ghost nodes that are refreshed before every pass over A
  int next_rep;          // the next level is the replicated entry level: its b is all-gathered from the ranks' rows
  HaloDev halo;          // ghost update of this level's x
  HaloDev gather;        // next_rep: all-gather plan ("ghosts" = the other ranks' rows, rank order)
  double* rep_tmp;       // next_rep: [nx.n * BS] this rank's rows of the restricted residual first, then the others'
  int64_t rep_off;       // next_rep: rows owned by lower ranks
  const int32_t* coarse_gid;  // next_rep: local coarse node (column of A*P) -> row of the replicated level
};

struct CoarseCycleDev {
  int nlevels;              // number of levels handled here (the first one is level `first` of the hierarchy)
  int dense_n;              // > 0: scalar size of the dense inverse of the last level
  const double* coarse_inv;
  unsigned long long* barrier;  // [3]: arrival counter (monotonic), arrivals expected before this launch, exit ticket
  CoarseLevelDev lev[CC_MAX_LEVELS];
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Grid barrier for a grid whose CTAs are all resident (<= 1 per SM): wait until the monotonic arrival counter reaches
// `target`.
__device__ __forceinline__ void cc_grid_barrier(unsigned long long* counter, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1ULL);
    while (ld_acquire_gpu(counter) < target) {}
  }
  __syncthreads();
}

// Vectors change between the phases of ONE kernel, written by other SMs: every vector read goes to L2 (ld.global.cg),
// never through this SM's L1 / read-only path, whose lines may predate the last grid barrier.  Matrix data is
// constant for the lifetime of the kernel and takes the default path.
#define CC_LD(ptr) __ldcg(ptr)

// (A x)_row for block row `row`, computed by one warp: lanes stride over the blocks of the row, fixed shuffle tree.
// The result is valid in ALL lanes.
template <int BS>
__device__ __forceinline__ void cc_row_times(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                             const double* __restrict__ val, const double* x, int64_t row,
                                             int lane, double (&out)[BS], const int32_t* __restrict__ gid = nullptr) {
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) acc[i] = 0.0;
  const int64_t kb = rowptr[row], ke = rowptr[row + 1];
  for (int64_t k = kb + lane; k < ke; k += 32) {
    const int64_t cidx = gid ? (int64_t)gid[col[k]] : (int64_t)col[k];
    const double* xb = x + cidx * BS;
    const double* a = val + k * (BS * BS);
    double xv[BS];
#pragma unroll
    for (int j = 0; j < BS; ++j) xv[j] = CC_LD(xb + j);
#pragma unroll
    for (int i = 0; i < BS; ++i) {
#pragma unroll
      for (int j = 0; j < BS; ++j) acc[i] = fma(a[i * BS + j], xv[j], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < BS; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    out[i] = acc[i];
  }
}

// sum_k w[k] * v[col[k], :] over one row of a scalar-weighted transfer operator (P or P^T), by one warp
template <int BS, bool RESIDUAL>
__device__ __forceinline__ void cc_row_transfer(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                const double* __restrict__ w, const double* v,
                                                const double* v2, int64_t row, int lane, double (&out)[BS]) {
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) acc[i] = 0.0;
  const int64_t kb = rowptr[row], ke = rowptr[row + 1];
  for (int64_t k = kb + lane; k < ke; k += 32) {
    const int64_t c = (int64_t)col[k] * BS;
    const double wk = w[k];
#pragma unroll
    for (int i = 0; i < BS; ++i) acc[i] += wk * (RESIDUAL ? (CC_LD(v + c + i) - CC_LD(v2 + c + i)) : CC_LD(v + c + i));
  }
#pragma unroll
  for (int i = 0; i < BS; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    out[i] = acc[i];
  }
}

constexpr int CC_THREADS = 1024;  // one CTA per SM: 32 warps keep enough row gathers in flight

template <int BS>
__global__ void __launch_bounds__(CC_THREADS, 1)
k_coarse_cycle(const CoarseCycleDev* __restrict__ cyp, const int* __restrict__ done) {
  if (done && *done) return;
  __shared__ CoarseCycleDev cy;
  {
    const int words = (int)(sizeof(CoarseCycleDev) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(cyp);
    int* dst = reinterpret_cast<int*>(&cy);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t tid0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  unsigned long long* counter = cy.barrier;
  // the arrival counter only grows; barrier[1] holds its value at the start of this launch (written by the previous one)
  unsigned long long target = *reinterpret_cast<volatile unsigned long long*>(cy.barrier + 1);
  auto barrier = [&]() { target += gridDim.x; cc_grid_barrier(counter, target); };

  const int L = cy.nlevels;
  // t = A x of a level: x complete (barrier), ghosts refreshed in place when the level is distributed, rows, barrier
  auto spmv_phase = [&](const CoarseLevelDev& lv) {
    barrier();
    if (lv.has_halo) {
      p2p_halo_exchange_grid<BS>(lv.halo, lv.x);
      barrier();
    }
    for (int64_t r = warp0; r < lv.n; r += nwarps) {
      double ax[BS];
      cc_row_times<BS>(lv.rowptr, lv.col, lv.val, lv.x, r, lane, ax);
      if (lane < BS) lv.t[r * BS + lane] = ax[lane];
    }
    barrier();
  };
  auto cheb_step = [&](const CoarseLevelDev& lv, int j) {
    const int64_t ns = lv.n * BS;
    for (int64_t i = tid0; i < ns; i += nthreads) {
      const double di = lv.c1[j] * CC_LD(lv.d + i) + lv.c2[j] * lv.dinv[i] * (CC_LD(lv.b + i) - CC_LD(lv.t + i));
      lv.d[i] = di;
      lv.x[i] = CC_LD(lv.x + i) + di;
    }
  };
  // ---- down: pre-smoothing from a zero guess, residual, restriction.  b of the first level was written by the caller.
  for (int l = 0; l < L; ++l) {
    const CoarseLevelDev& lv = cy.lev[l];
    const int64_t ns = lv.n * BS;
    const bool last = (l == L - 1);
    if (last && cy.dense_n > 0) {
      // x = inv * b, one warp per scalar row
      const int n = cy.dense_n;
      for (int64_t r = warp0; r < n; r += nwarps) {
        double s = 0.0;
        for (int j = lane; j < n; j += 32) s += cy.coarse_inv[r * n + j] * CC_LD(lv.b + j);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) lv.x[r] = s;
      }
      barrier();
      break;
    }
    // first Chebyshev step: d = x = inv_theta * dinv .* b   (b is complete: see the end of the previous level)
    for (int64_t i = tid0; i < ns; i += nthreads) {
      const double di = lv.inv_theta * lv.dinv[i] * CC_LD(lv.b + i);
      lv.d[i] = di;
      lv.x[i] = di;
    }
    for (int j = 1; j < lv.degree; ++j) {
      spmv_phase(lv);
      cheb_step(lv, j);
    }
    if (last) {  // Chebyshev "solve" of the last level: done
      barrier();
      break;
    }
    spmv_phase(lv);
    const CoarseLevelDev& nx = cy.lev[l + 1];
    double* rc_out = lv.next_rep ? lv.rep_tmp : nx.b;
    for (int64_t r = warp0; r < lv.nc; r += nwarps) {
      double rc[BS];
      cc_row_transfer<BS, true>(lv.pt_rowptr, lv.pt_col, lv.pt_val, lv.b, lv.t, r, lane, rc);
      if (lane < BS) rc_out[r * BS + lane] = rc[lane];
    }
    barrier();  // restricted residual complete
    if (lv.next_rep) {
      // every coarse row is produced by exactly one rank (aggregates never cross ranks): all-gather over NVLink,
      // then reorder [own rows | other ranks' rows] into the global row order of the replicated level
      p2p_halo_exchange_grid<BS>(lv.gather, lv.rep_tmp);
      barrier();
      const int64_t n_own = lv.nc, lo = lv.rep_off, total = nx.n * BS;
      for (int64_t q = tid0; q < total; q += nthreads) {
        const int64_t i = q / BS;
        const int64_t src = (i < lo) ? (n_own + i) : (i < lo + n_own ? i - lo : i);
        nx.b[q] = CC_LD(lv.rep_tmp + src * BS + (q - i * BS));
      }
      barrier();
    }
  }
  // ---- up: coarse-grid correction + first post-smoothing step through A*P, further steps through A
  for (int l = L - 2; l >= 0; --l) {
    const CoarseLevelDev& lv = cy.lev[l];
    const CoarseLevelDev& nx = cy.lev[l + 1];
    const double c2 = lv.inv_theta;
    if (nx.has_halo) {  // A*P gathers from the ghost coarse nodes too
      p2p_halo_exchange_grid<BS>(nx.halo, nx.x);
      barrier();
    }
    const double* xc_own = lv.next_rep ? nx.x + lv.rep_off * BS : nx.x;  // P maps to this rank's coarse nodes only
    const int32_t* gid = lv.next_rep ? lv.coarse_gid : nullptr;
    for (int64_t r = warp0; r < lv.n; r += nwarps) {
      double pe[BS], ape[BS];
      cc_row_transfer<BS, false>(lv.p_rowptr, lv.p_col, lv.p_val, xc_own, nullptr, r, lane, pe);
      cc_row_times<BS>(lv.ap_rowptr, lv.ap_col, lv.ap_val, nx.x, r, lane, ape, gid);
      if (lane < BS) {
        const int64_t i = r * BS + lane;
        const double xi = CC_LD(lv.x + i) + pe[lane];
        const double di = c2 * lv.dinv[i] * (CC_LD(lv.b + i) - CC_LD(lv.t + i) - ape[lane]);
        lv.d[i] = di;
        lv.x[i] = xi + di;
      }
    }
    for (int j = 1; j < lv.degree; ++j) {
      spmv_phase(lv);
      cheb_step(lv, j);
    }
    if (l > 0) barrier();  // x of this level complete before the level above gathers from it
  }
  // the last CTA to leave publishes the counter value the next launch starts from (every CTA passed the same number
  // of barriers, so `target` is the same everywhere; everybody read barrier[1] before the first barrier)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(cy.barrier + 2);
    if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) cy.barrier[1] = target;
  }
}

}  // namespace cb
