// Block-CSR SpMV building blocks shared by spmv.cu and krylov.cu.
//
// Streaming design ("CSR-stream"): a CTA owns R consecutive block rows.  The nonzeros of those
// rows are one contiguous span of `val` / `col`, which the CTA streams through shared memory
// tile by tile with perfectly coalesced loads (every DRAM sector of the matrix is used once);
// the products val * x[col] are written to shared memory and then reduced row by row in a
// fixed order.  x is gathered through L1/L2 (for mesh matrices consecutive rows share almost
// all of their columns, so the gathers hit in cache and DRAM traffic for x stays ~8 B/row).
#pragma once
#include "common.cuh"

namespace cb {

template <int BS>
struct SpmvCfg {
  static constexpr int THREADS = 256;
  static constexpr int R = (BS == 1) ? 128 : (BS == 2 ? 64 : 32);  // block rows per CTA
  static constexpr int T = 16 * R;                                  // blocks per smem tile
  static constexpr int TILE_DOUBLES = T * BS * BS;
};

// Computes rows [row0, row0+nr) of y = A x.  Result for scalar row (lr, i) is returned in
// `acc` of thread t = lr*BS + i (t < nr*BS).  smem_tile: TILE_DOUBLES doubles, smem_rp: R+1.
template <int BS>
__device__ __forceinline__ double spmv_row_block(const int64_t* __restrict__ rowptr,
                                                 const int32_t* __restrict__ col,
                                                 const double* __restrict__ val,
                                                 const double* __restrict__ x, int64_t row0, int nr,
                                                 double* smem_tile, int64_t* smem_rp) {
  using C = SpmvCfg<BS>;
  constexpr int BB = BS * BS;
  const int tid = threadIdx.x;
  if (tid <= nr) smem_rp[tid] = rowptr[row0 + tid];
  __syncthreads();
  const int64_t k0 = smem_rp[0], k1 = smem_rp[nr];
  const int lr = tid / BS, li = tid - lr * BS;
  const bool rowthread = tid < nr * BS;
  int64_t rb = 0, re = 0;
  if (rowthread) { rb = smem_rp[lr]; re = smem_rp[lr + 1]; }
  double acc = 0.0;
  if (k0 == k1) __syncthreads();  // keep smem_rp stable until every thread has read it
  for (int64_t s = k0; s < k1; s += C::T) {
    const int64_t tile_blocks = (k1 - s < C::T) ? (k1 - s) : C::T;
    const int tile_elems = (int)tile_blocks * BB;
    const double* vsrc = val + s * BB;
#pragma unroll 4
    for (int e = tid; e < tile_elems; e += C::THREADS) {
      const int blk = e / BB;
      const int j = (e - blk * BB) % BS;
      const int32_t cidx = __ldg(col + s + blk);
      const double a = __ldcs(vsrc + e);  // streamed once: evict-first
      smem_tile[e] = a * __ldg(x + (int64_t)cidx * BS + j);
    }
    __syncthreads();
    if (rowthread) {
      const int64_t b = (rb > s ? rb : s) - s;
      const int64_t e2 = (re < s + tile_blocks ? re : s + tile_blocks) - s;
      for (int64_t k = b; k < e2; ++k) {
#pragma unroll
        for (int j = 0; j < BS; ++j) acc += smem_tile[k * BB + li * BS + j];
      }
    }
    __syncthreads();
  }
  return acc;
}

}  // namespace cb
