// Block-CSR SpMV building block shared by spmv.cu and krylov.cu (sm_100a).
//
// Design ("TMA-staged row blocks"): a CTA owns R consecutive block rows.  Their nonzeros are
// one contiguous span of `val` / `col`; one elected thread streams that span into shared
// memory with two 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, completion on an
// mbarrier), so the matrix -- 85-95 % of all bytes -- never passes through the LSU or the
// register file and every DRAM sector of it is fetched exactly once.  BS threads then own one
// block row: thread (r, j) multiplies column j of every block of row r with x[BS*col + j] (one
// gather per block and thread; for mesh matrices neighbouring rows hit neighbouring columns, so
// a warp's gather touches 2-3 lines and is served by L1/L2) and the BS partial vectors are
// transposed-reduced with two warp shuffles per row.  No shared-memory partial products, no
// atomics, fixed summation order per row: bit-reproducible.
// Several CTAs are resident per SM (single-buffered): while one waits on its mbarrier the
// others compute, which keeps ~150-200 KB of matrix data in flight per SM.
#pragma once
#include "common.cuh"

namespace cb {

// VT is the storage type of the matrix values: double for operators the Krylov method itself applies, float for
// the copies the preconditioner reads (cb_mat.val32: half the bytes per pass; the products and sums are fp64
// either way, so a V-cycle over rounded entries is still a fixed linear operator).
template <int BS, typename VT = double>
struct SpmvCfg {
  static constexpr int ROWS_PER_WARP = 32 / BS;           // 32, 16, 10 (2 idle lanes for BS = 3)
#ifndef CB_SPMV_WARPS3
#define CB_SPMV_WARPS3 2
#endif
#ifndef CB_SPMV_CTAS3
#define CB_SPMV_CTAS3 9
#endif
#ifndef CB_SPMV_WARPS1
#define CB_SPMV_WARPS1 4
#endif
#ifndef CB_SPMV_CTAS1
#define CB_SPMV_CTAS1 8
#endif
#ifndef CB_SPMV_TILE
#define CB_SPMV_TILE 16
#endif
// float tiles are half as large, so twice as many CTAs keep the same number of bytes in flight per SM
#ifndef CB_SPMV_CTAS3_F32
#define CB_SPMV_CTAS3_F32 16
#endif
#ifndef CB_SPMV_CTAS1_F32
#define CB_SPMV_CTAS1_F32 10
#endif
  // many small CTAs beat few large ones: every CTA is one independent bulk-copy stream, and the
  // DRAM pipe wants ~9 x 24 KB in flight per SM (measured: 4 x 47 KB -> 5.3 TB/s, 9 x 24 KB -> 6.0 TB/s)
  static constexpr int WARPS = (BS == 3) ? CB_SPMV_WARPS3 : (BS == 1 ? CB_SPMV_WARPS1 : 4);
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int R = ROWS_PER_WARP * WARPS;          // block rows per CTA: 128, 64, 40
  static constexpr int T = CB_SPMV_TILE * R;               // tile capacity in blocks
  static constexpr int BB = BS * BS;
  static constexpr bool F32 = sizeof(VT) == 4;
  static constexpr int EPB = 16 / (int)sizeof(VT);         // values per 16-byte bulk-copy granule
  static constexpr int VAL_ELEMS = T * BB + EPB;           // + alignment slack (head)
  static constexpr int COL_INTS = T + 8;                   // +8: alignment slack (head up to 3)
  static constexpr size_t SMEM_BYTES = sizeof(VT) * VAL_ELEMS + sizeof(int32_t) * COL_INTS + 16;
  static constexpr int CTAS_PER_SM = F32 ? ((BS == 3) ? CB_SPMV_CTAS3_F32 : (BS == 2 ? 8 : CB_SPMV_CTAS1_F32))
                                         : ((BS == 3) ? CB_SPMV_CTAS3 : (BS == 2 ? 5 : CB_SPMV_CTAS1));
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-CTA state for the streaming SpMV (dynamic shared memory carve-up).
template <int BS, typename VT = double>
struct SpmvSmem {
  VT* val;
  int32_t* col;
  uint64_t* bar;
  __device__ __forceinline__ explicit SpmvSmem(unsigned char* base) {
    using C = SpmvCfg<BS, VT>;
    val = reinterpret_cast<VT*>(base);
    col = reinterpret_cast<int32_t*>(base + sizeof(VT) * C::VAL_ELEMS);
    bar = reinterpret_cast<uint64_t*>(base + sizeof(VT) * C::VAL_ELEMS + sizeof(int32_t) * C::COL_INTS);
  }
};

// One-time per-CTA setup: carve dynamic shared memory, initialise the mbarrier.
template <int BS, typename VT = double>
__device__ __forceinline__ SpmvSmem<BS, VT> spmv_smem_init(unsigned char* base) {
  SpmvSmem<BS, VT> sm(base);
  if (threadIdx.x == 0) {
    mbar_init(sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();
  return sm;
}

// Stream tile [s, s+nblk) of (val, col) into shared memory.  Called by ONE thread.  The bulk
// copies cover the 16-byte aligned interior; the unaligned head / tail elements (at most one
// double, three ints each side) are moved with plain loads by the same thread before it
// arrives on the barrier (arrive = release, so waiters see them).  Returns nothing; the
// element offsets (val_head, col_head) are recomputed by every thread with tile_heads().
template <int BS, typename VT>
__device__ __forceinline__ void issue_tile(const SpmvSmem<BS, VT>& sm, const VT* __restrict__ val,
                                           const int32_t* __restrict__ col, int64_t s, int nblk) {
  constexpr int BB = BS * BS;
  constexpr int EPB = 16 / (int)sizeof(VT);
  const VT* vsrc = val + s * BB;
  const int32_t* csrc = col + s;
  const int vhead = (int)((reinterpret_cast<uintptr_t>(vsrc) & 15) / sizeof(VT));  // 0..EPB-1 values
  const int chead = (int)((reinterpret_cast<uintptr_t>(csrc) & 15) >> 2);          // 0..3 ints
  const int nv = nblk * BB, ncol = nblk;
  // aligned interior in elements relative to the tile start
  const int vb0 = vhead ? (EPB - vhead) : 0;       // first element at a 16-byte boundary
  const int vb = vb0 < nv ? vb0 : nv;
  const int ve = vb + ((nv - vb) & ~(EPB - 1));    // whole granules
  const int cb = chead ? (4 - chead) : 0;
  const int cbeg = cb < ncol ? cb : ncol;
  const int cend = cbeg + ((ncol - cbeg) & ~3);
  uint32_t bytes = 0;
  fence_proxy_async();
  if (ve > vb) {
    bulk_g2s(sm.val + vhead + vb, vsrc + vb, (uint32_t)(ve - vb) * (uint32_t)sizeof(VT), sm.bar);
    bytes += (uint32_t)(ve - vb) * (uint32_t)sizeof(VT);
  }
  if (cend > cbeg) {
    bulk_g2s(sm.col + chead + cbeg, csrc + cbeg, (uint32_t)(cend - cbeg) * 4u, sm.bar);
    bytes += (uint32_t)(cend - cbeg) * 4u;
  }
  for (int e = 0; e < vb; ++e) sm.val[vhead + e] = vsrc[e];
  for (int e = ve; e < nv; ++e) sm.val[vhead + e] = vsrc[e];
  for (int e = 0; e < cbeg; ++e) sm.col[chead + e] = csrc[e];
  for (int e = cend; e < ncol; ++e) sm.col[chead + e] = csrc[e];
  mbar_arrive_expect_tx(sm.bar, bytes);
}

// Rows per CTA for a row split of G thread groups per row (G in {1,2,4,8}; long rows of coarse
// AMG operators are shared by G*BS lanes).
template <int BS>
__host__ __device__ __forceinline__ int spmv_rows_per_cta(int G) {
  return (32 / (BS * G)) * SpmvCfg<BS>::WARPS;
}
__host__ __device__ __forceinline__ int spmv_sanitize_split(int g) { return g >= 8 ? 8 : (g >= 4 ? 4 : (g >= 2 ? 2 : 1)); }

// Computes block rows [row0, row0+nr) of y = A x, nr <= spmv_rows_per_cta<BS>(G).  Lane layout
// inside a warp: lane = (r*G + g)*BS + j  (row r, split group g, block column j).  On return the
// g == 0 lanes hold in `acc_out` the result for scalar row (row0 + warp*rpw + r)*BS + j, which
// is returned (-1 for all other lanes).  `phase` is the CTA's mbarrier parity (in/out).  All
// threads of the CTA must call this together.
template <int BS, typename VT>
__device__ __forceinline__ int64_t spmv_row_block(const SpmvSmem<BS, VT>& sm, const int64_t* __restrict__ rowptr,
                                                  const int32_t* __restrict__ col,
                                                  const VT* __restrict__ val,
                                                  const double* __restrict__ x, int64_t row0, int nr, int G,
                                                  uint32_t& phase, double& acc_out) {
  using C = SpmvCfg<BS, VT>;
  constexpr int BB = BS * BS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rpw = 32 / (BS * G);
  const int q = lane / BS, j = lane - q * BS;
  const int r = q / G, g = q - r * G;
  const int lr = warp * rpw + r;
  const bool active = (q < rpw * G) && (lr < nr);
  const int64_t k0 = __ldg(rowptr + row0), k1 = __ldg(rowptr + row0 + nr);
  int64_t rb = 0, re = 0;
  if (active) { rb = __ldg(rowptr + row0 + lr); re = __ldg(rowptr + row0 + lr + 1); }
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; ++i) acc[i] = 0.0;
  for (int64_t s = k0; s < k1; s += C::T) {
    const int nblk = (int)((k1 - s < C::T) ? (k1 - s) : C::T);
    if (tid == 0) issue_tile<BS, VT>(sm, val, col, s, nblk);
    const int vhead = (int)((reinterpret_cast<uintptr_t>(val + s * BB) & 15) / sizeof(VT));
    const int chead = (int)((reinterpret_cast<uintptr_t>(col + s) & 15) >> 2);
    mbar_wait(sm.bar, phase);
    phase ^= 1u;
    if (active) {
      const int b = (int)((rb > s ? rb : s) - s);
      const int e = (int)((re < s + nblk ? re : s + nblk) - s);
      const VT* sv = sm.val + vhead;
      const int32_t* sc = sm.col + chead;
#pragma unroll 4
      for (int k = b + g; k < e; k += G) {
        const int32_t c = sc[k];
        const double xv = __ldg(x + (int64_t)c * BS + j);
#pragma unroll
        for (int i = 0; i < BS; ++i) acc[i] = fma((double)sv[k * BB + i * BS + j], xv, acc[i]);
      }
    }
    __syncthreads();  // the tile is consumed: the next bulk copy may overwrite it
  }
  // sum the G split groups of a row (fixed tree order), result in the g == 0 lanes
  for (int stride = G >> 1; stride > 0; stride >>= 1) {
#pragma unroll
    for (int i = 0; i < BS; ++i) {
      const double o = __shfl_down_sync(0xffffffffu, acc[i], stride * BS);
      acc[i] += o;
    }
  }
  // transpose-reduce the BS partial vectors of a row: lane j ends with y_j
  double y = acc[0];
  if constexpr (BS > 1) {
    const int base = lane - j;
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < BS; ++i) mine = (i == j) ? acc[i] : mine;
    y = mine;
#pragma unroll
    for (int d = 1; d < BS; ++d) {
      const int want = (j - d + BS) % BS;  // element the lane d positions "below" needs from us
      double send = 0.0;
#pragma unroll
      for (int i = 0; i < BS; ++i) send = (i == want) ? acc[i] : send;
      int src = base + (j + d) % BS;
      src = src < 32 ? src : lane;
      const double recv = __shfl_sync(0xffffffffu, send, src);
      y += recv;
    }
  }
  acc_out = y;
  return (active && g == 0) ? (row0 + lr) * BS + j : -1;
}

}  // namespace cb
