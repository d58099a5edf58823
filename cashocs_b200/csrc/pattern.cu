// Sparsity pattern of the dofmap graph, cell -> nnz scatter map and cell colouring.
//
// Replaces dolfin's sparsity-pattern builder behind fenics.assemble_system with the
// cashocs form modification (cashocs/_utils/forms.py:274-288, cashocs/_utils/linalg.py:149-156):
// the pattern is every (test dof, trial dof) pair of every cell, diagonal included, columns
// sorted ascending within a row.  Everything here runs once per mesh topology.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cb {

constexpr int MAX_ROW_NNZ = 256;  // capacity of the per-row neighbour buffer

__global__ void k_count_incidence(int64_t n_entries, const int32_t* __restrict__ cell_dofs,
                                  unsigned long long* __restrict__ cnt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_entries;
       i += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&cnt[cell_dofs[i]], 1ULL);
}

__global__ void k_fill_incidence(int64_t nc, int k, const int32_t* __restrict__ cell_dofs,
                                 const int64_t* __restrict__ ptr, unsigned int* __restrict__ cursor,
                                 int32_t* __restrict__ idx) {
  const int64_t n_entries = nc * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_entries;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t dof = cell_dofs[i];
    const unsigned int slot = atomicAdd(&cursor[dof], 1u);
    idx[ptr[dof] + slot] = (int32_t)(i / k);
  }
}

// Arrival order of the atomics is arbitrary -> sort every (short) list so the adjacency,
// and everything derived from it, is bit-reproducible.
__global__ void k_sort_incidence(int64_t nd, const int64_t* __restrict__ ptr, int32_t* __restrict__ idx) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nd;
       v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = ptr[v], e = ptr[v + 1];
    for (int64_t i = b + 1; i < e; ++i) {
      const int32_t key = idx[i];
      int64_t j = i - 1;
      while (j >= b && idx[j] > key) {
        idx[j + 1] = idx[j];
        --j;
      }
      idx[j + 1] = key;
    }
  }
}

// Sorted unique neighbour dofs of row `row`; returns the count (or -1 on overflow).
__device__ int collect_row(int64_t row, int k, const int32_t* __restrict__ cell_dofs,
                           const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                           int32_t* buf) {
  int n = 0;
  for (int64_t t = v2c_ptr[row]; t < v2c_ptr[row + 1]; ++t) {
    const int64_t c = v2c_idx[t];
    for (int a = 0; a < k; ++a) {
      const int32_t dof = cell_dofs[c * k + a];
      // binary search for the insertion point in buf[0..n)
      int lo = 0, hi = n;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (buf[mid] < dof) lo = mid + 1; else hi = mid;
      }
      if (lo < n && buf[lo] == dof) continue;
      if (n >= MAX_ROW_NNZ) return -1;
      for (int j = n; j > lo; --j) buf[j] = buf[j - 1];
      buf[lo] = dof;
      ++n;
    }
  }
  return n;
}

__global__ void k_row_lengths(int64_t nrows, int k, const int32_t* __restrict__ cell_dofs,
                              const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                              int64_t* __restrict__ rowlen, int* __restrict__ err) {
  int32_t buf[MAX_ROW_NNZ];
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int n = collect_row(r, k, cell_dofs, v2c_ptr, v2c_idx, buf);
    if (n < 0) { *err = 1; rowlen[r] = 0; } else rowlen[r] = n;
  }
}

__global__ void k_row_fill(int64_t nrows, int k, const int32_t* __restrict__ cell_dofs,
                           const int64_t* __restrict__ v2c_ptr, const int32_t* __restrict__ v2c_idx,
                           const int64_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  int32_t buf[MAX_ROW_NNZ];
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < nrows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int n = collect_row(r, k, cell_dofs, v2c_ptr, v2c_idx, buf);
    const int64_t b = rowptr[r];
    for (int j = 0; j < n; ++j) col[b + j] = buf[j];
  }
}

__global__ void k_cell2nnz(int64_t nc, int k, int64_t nrows, const int32_t* __restrict__ cell_dofs,
                           const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                           int32_t* __restrict__ cell2nnz) {
  const int64_t total = nc * k * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / (k * k);
    const int ab = (int)(i - c * k * k);
    const int a = ab / k, b = ab - a * k;
    const int32_t row = cell_dofs[c * k + a];
    const int32_t cdof = cell_dofs[c * k + b];
    int32_t pos = -1;
    if (row < nrows) {
      int64_t lo = rowptr[row], hi = rowptr[row + 1];
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (col[mid] < cdof) lo = mid + 1; else hi = mid;
      }
      pos = (int32_t)lo;
    }
    cell2nnz[i] = pos;
  }
}

}  // namespace cb

using namespace cb;

extern "C" int cb_v2c_build(void* stream, int64_t nd, int64_t nc, int k, const int32_t* cell_dofs,
                            int64_t* v2c_ptr, int32_t* v2c_idx) {
  CB_REQUIRE(nd > 0 && nc > 0 && k > 0, "cb_v2c_build: empty mesh");
  configure_scratch_pool();
  cudaStream_t s = as_stream(stream);
  unsigned long long* cnt = nullptr;
  unsigned int* cursor = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  CB_CUDA_CHECK(cudaMallocAsync(&cnt, sizeof(unsigned long long) * (nd + 1), s));
  CB_CUDA_CHECK(cudaMallocAsync(&cursor, sizeof(unsigned int) * nd, s));
  CB_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (nd + 1), s));
  CB_CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(unsigned int) * nd, s));
  k_count_incidence<<<grid_for(nc * k, 256, 8), 256, 0, s>>>(nc * k, cell_dofs, cnt);
  CB_LAUNCH_CHECK();
  CB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const long long*)cnt,
                                              (long long*)v2c_ptr, nd + 1, s));
  CB_CUDA_CHECK(cudaMallocAsync(&tmp, tmp_bytes, s));
  CB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const long long*)cnt,
                                              (long long*)v2c_ptr, nd + 1, s));
  k_fill_incidence<<<grid_for(nc * k, 256, 8), 256, 0, s>>>(nc, k, cell_dofs, v2c_ptr, cursor, v2c_idx);
  CB_LAUNCH_CHECK();
  k_sort_incidence<<<grid_for(nd, 128, 16), 128, 0, s>>>(nd, v2c_ptr, v2c_idx);
  CB_LAUNCH_CHECK();
  CB_CUDA_CHECK(cudaFreeAsync(tmp, s));
  CB_CUDA_CHECK(cudaFreeAsync(cnt, s));
  CB_CUDA_CHECK(cudaFreeAsync(cursor, s));
  return CB_OK;
}

extern "C" int cb_pattern_rowptr(void* stream, int64_t nrows, int k, const int32_t* cell_dofs,
                                 const int64_t* v2c_ptr, const int32_t* v2c_idx, int64_t* rowptr) {
  CB_REQUIRE(nrows > 0, "cb_pattern_rowptr: no rows");
  configure_scratch_pool();
  cudaStream_t s = as_stream(stream);
  int64_t* rowlen = nullptr;
  int* err = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  CB_CUDA_CHECK(cudaMallocAsync(&rowlen, sizeof(int64_t) * (nrows + 1), s));
  CB_CUDA_CHECK(cudaMallocAsync(&err, sizeof(int), s));
  CB_CUDA_CHECK(cudaMemsetAsync(rowlen, 0, sizeof(int64_t) * (nrows + 1), s));
  CB_CUDA_CHECK(cudaMemsetAsync(err, 0, sizeof(int), s));
  k_row_lengths<<<grid_for(nrows, 128, 8), 128, 0, s>>>(nrows, k, cell_dofs, v2c_ptr, v2c_idx, rowlen, err);
  CB_LAUNCH_CHECK();
  CB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const long long*)rowlen,
                                              (long long*)rowptr, nrows + 1, s));
  CB_CUDA_CHECK(cudaMallocAsync(&tmp, tmp_bytes, s));
  CB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const long long*)rowlen,
                                              (long long*)rowptr, nrows + 1, s));
  int herr = 0;
  CB_CUDA_CHECK(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CB_CUDA_CHECK(cudaStreamSynchronize(s));
  CB_CUDA_CHECK(cudaFreeAsync(tmp, s));
  CB_CUDA_CHECK(cudaFreeAsync(rowlen, s));
  CB_CUDA_CHECK(cudaFreeAsync(err, s));
  if (herr) {
    set_error("cb_pattern_rowptr: a row has more than %d neighbours", MAX_ROW_NNZ);
    return CB_ERR_CAPACITY;
  }
  return CB_OK;
}

extern "C" int cb_pattern_fill(void* stream, int64_t nrows, int64_t nc, int k,
                               const int32_t* cell_dofs, const int64_t* v2c_ptr,
                               const int32_t* v2c_idx, const int64_t* rowptr, int32_t* col,
                               int32_t* cell2nnz) {
  cudaStream_t s = as_stream(stream);
  k_row_fill<<<grid_for(nrows, 128, 8), 128, 0, s>>>(nrows, k, cell_dofs, v2c_ptr, v2c_idx, rowptr, col);
  CB_LAUNCH_CHECK();
  if (cell2nnz) {
    k_cell2nnz<<<grid_for(nc * k * k, 256, 8), 256, 0, s>>>(nc, k, nrows, cell_dofs, rowptr, col, cell2nnz);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

// Greedy first-fit colouring in cell order with a 128-bit "colours in use" mask per dof.
// Serial on the host: O(nc*k), deterministic, runs once per mesh topology.
extern "C" int cb_colour_cells_host(int64_t nd, int64_t nc, int k, const int32_t* cell_dofs_host,
                                    int32_t* colour_host) {
  if (nd <= 0 || nc <= 0 || k <= 0) {
    set_error("cb_colour_cells_host: empty mesh");
    return CB_ERR_ARG;
  }
  std::vector<uint64_t> lo((size_t)nd, 0), hi((size_t)nd, 0);
  int ncol = 0;
  for (int64_t c = 0; c < nc; ++c) {
    uint64_t flo = 0, fhi = 0;
    for (int a = 0; a < k; ++a) {
      const int32_t d = cell_dofs_host[c * k + a];
      flo |= lo[d];
      fhi |= hi[d];
    }
    int colour;
    if (~flo) colour = __builtin_ctzll(~flo);
    else if (~fhi) colour = 64 + __builtin_ctzll(~fhi);
    else {
      set_error("cb_colour_cells_host: more than 128 colours needed");
      return CB_ERR_CAPACITY;
    }
    colour_host[c] = colour;
    ncol = std::max(ncol, colour + 1);
    for (int a = 0; a < k; ++a) {
      const int32_t d = cell_dofs_host[c * k + a];
      if (colour < 64) lo[d] |= (1ULL << colour); else hi[d] |= (1ULL << (colour - 64));
    }
  }
  return ncol;
}
