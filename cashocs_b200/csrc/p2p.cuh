// Device side of the one-sided NVLink transport (csrc/dist.cu owns the windows; krylov.cu fuses the exchanges with the
// Krylov recurrences).  Every rank exports one cudaMalloc'ed window per halo plan through CUDA IPC; neighbours write
// their boundary values straight into its landing zone with plain stores over NVLink, then release a per-source flag
// holding the exchange epoch; the receiver acquires the flags and copies the landing zone behind the owned part of the
// vector.  Landing zone and reduction slots are double-buffered by epoch parity: a neighbour can be at most one exchange
// ahead (it needs my next message to get further), so parity e+1 never overwrites data still read.
// An exchange is ONE kernel: put phase, flags, wait, get phase -- the epochs live in the window, so a captured CUDA
// graph replays it unchanged.
#pragma once
#include "common.cuh"

namespace cb {

constexpr int RED_W = 8;  // doubles per reduction message

// Window layout in doubles.  The header (flags, reduction slots) sits at FIXED offsets so a peer can address it without
// knowing this rank's ghost count; the landing zone follows, its parity stride (`ghost_cap`) is the RECEIVER's and is
// told to every sender at setup (d_remote_cap).
// uint64 header words: halo_flag[nranks], red_flag[nranks], ticket, error, halo_epoch, red_epoch
struct WindowLayout {
  int nranks;
  __host__ __device__ int64_t flags_off() const { return 0; }                       // 2*nranks + 4 uint64
  __host__ __device__ int64_t red_off() const { return 2LL * nranks + 4; }          // [2][nranks][RED_W]
  __host__ __device__ int64_t land_off() const { return red_off() + 2LL * nranks * RED_W; }
  __host__ __device__ int64_t total_doubles(int64_t ghost_cap) const { return land_off() + 2 * ghost_cap; }
};

// what a kernel needs of a cb_dist with the one-sided transport switched on
struct P2PView {
  unsigned char* const* peer_window;  // device array [nranks]
  unsigned char* window;              // my window
  unsigned long long* error;          // error word SHARED by a plan and every plan cloned from it (device)
  unsigned long long timeout_cycles;  // a wait gives up after this many cycles and raises *error
  WindowLayout lay;
  int rank;
};

// device-side description of one halo plan (what k_halo_exchange / the persistent V-cycle kernel need of a cb_dist)
struct HaloDev {
  int nneigh;
  const int64_t* send_off;    // [nneigh+1]
  const int64_t* remote_off;  // [nneigh] node offset of my data in the peer's ghost range
  const int32_t* neigh;       // [nneigh]
  const int32_t* send_idx;    // [nsend] owned local nodes to send
  const int64_t* remote_cap;  // [nneigh] parity stride of the peer's landing zone
  const int64_t* recv_off;    // [nneigh+1]
  int64_t n_owned;
  int64_t ghost_cap;
  P2PView v;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Spin until *flag >= epoch.  A peer that never answers must not hang the GPU: after `timeout_cycles` (default 60 s of
// SM clock, CASHOCS_B200_P2P_TIMEOUT_S) the wait raises the shared error word and returns false; so does every later
// wait on any plan of the same communicator, until the host has seen the error (cb_dist_p2p_error) -- the host layer
// checks it after every solve and every stand-alone exchange and raises.
__device__ __forceinline__ bool wait_flag(const unsigned long long* flag, unsigned long long epoch, const P2PView& v) {
  if (ld_acquire_sys(flag) >= epoch) return true;  // the common case: the message is already there
  const long long t0 = clock64();
  unsigned int backoff = 32, spins = 0;
  while (ld_acquire_sys(flag) < epoch) {
    // the error word is HOST memory (a PCIe round trip per load): look at it and at the clock only now and then
    if ((++spins & 1023u) == 0u) {
      if (*reinterpret_cast<volatile unsigned long long*>(v.error)) return false;  // an earlier wait already failed
      if ((unsigned long long)(clock64() - t0) > v.timeout_cycles) {
        *reinterpret_cast<volatile unsigned long long*>(v.error) = 1ULL;  // plain store: PCIe atomics are optional
        __threadfence_system();
        return false;
      }
    }
    __nanosleep(backoff);
    if (backoff < 256) backoff <<= 1;
  }
  return true;
}

__device__ __forceinline__ unsigned long long* window_header(unsigned char* window, const WindowLayout& lay) {
  return reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(window) + lay.flags_off());
}

// Halo exchange by ALL CTAs of the calling grid (every thread must call it; the grid's CTAs must be co-resident).
// Put phase: gather the boundary values and store them straight into the neighbours' landing zones over NVLink; the
// last CTA to finish releases the per-source epoch flags.  Get phase: every CTA waits for the neighbours' flags of this
// epoch and copies its share of the landing zone behind the owned part of x.  x must be complete on entry (the
// caller separates it from its producers by a kernel boundary or a grid barrier); on return THIS CTA's share of the
// ghosts is in place.  Returns false after a time-out.
template <int BS>
__device__ __forceinline__ bool p2p_halo_exchange_grid(const HaloDev& h, double* x) {
  // the exchange epoch lives in the window (not in a kernel argument) so that a captured CUDA graph
  // can replay the exchange: every thread reads it here, the last CTA of the put phase bumps it
  unsigned long long* hdr = window_header(h.v.window, h.v.lay);
  const int nr = h.v.lay.nranks;
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(hdr + 2 * nr + 2) + 1ULL;
  const int nneigh = h.nneigh;
  const int64_t nsend = h.send_off[nneigh];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nsend * BS;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = i / BS;
    const int c = (int)(i - q * BS);
    int j = 0;
    while (j + 1 < nneigh && q >= h.send_off[j + 1]) ++j;
    double* dst = reinterpret_cast<double*>(h.v.peer_window[h.neigh[j]]) + h.v.lay.land_off() +
                  (int64_t)(epoch & 1ULL) * h.remote_cap[j] + (h.remote_off[j] + (q - h.send_off[j])) * BS + c;
    *dst = __ldcg(x + (int64_t)h.send_idx[q] * BS + c);
  }
  // the last CTA to finish releases the flags (all stores of all CTAs are fenced before)
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  __shared__ int s_ok;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(hdr + 2 * nr);
  if (threadIdx.x == 0) { s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1); s_ok = 1; }
  __syncthreads();
  if (s_last) {
    for (int j = threadIdx.x; j < nneigh; j += blockDim.x) {
      __threadfence_system();
      st_release_sys(window_header(h.v.peer_window[h.neigh[j]], h.v.lay) + h.v.rank, epoch);
    }
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(hdr + 2 * nr + 2) = epoch;
  }
  // get phase
  for (int j = threadIdx.x; j < nneigh; j += blockDim.x) {
    if (h.recv_off[j + 1] > h.recv_off[j]) {
      if (!wait_flag(hdr + h.neigh[j], epoch, h.v)) s_ok = 0;
    }
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  if (ok) {
    const double* src = reinterpret_cast<const double*>(h.v.window) + h.v.lay.land_off() + (int64_t)(epoch & 1ULL) * h.ghost_cap;
    const int64_t n = h.recv_off[nneigh] * BS;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      x[h.n_owned * BS + i] = __ldcv(src + i);
  }
  __syncthreads();  // s_last / s_ok may be reused by the next call
  return ok;
}

// All-reduce of n <= RED_W doubles by ONE warp of one block: every rank writes its message into every rank's slots,
// waits for the nranks messages and sums them in rank order (bit-identical on all ranks, deterministic).  Returns
// false after a time-out (buf is then untouched).  op: 0 sum, 1 min, 2 max.  Call with the first 32 threads of a block.
__device__ __forceinline__ bool p2p_allreduce_warp(const P2PView& v, int n, double* buf, int op) {
  unsigned long long* hdr = window_header(v.window, v.lay);
  const int nr = v.lay.nranks;
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(hdr + 2 * nr + 3) + 1ULL;
  __syncwarp();  // everybody has read the epoch before lane 0 bumps it
  const int r = threadIdx.x;
  for (int q = r; q < nr; q += 32) {
    double* slots = reinterpret_cast<double*>(v.peer_window[q]) + v.lay.red_off() + ((int64_t)(epoch & 1ULL) * nr + v.rank) * RED_W;
    for (int i = 0; i < n; ++i) slots[i] = buf[i];
    __threadfence_system();
    st_release_sys(window_header(v.peer_window[q], v.lay) + nr + v.rank, epoch);
  }
  if (r == 0) hdr[2 * nr + 3] = epoch;
  bool ok = true;
  for (int q = r; q < nr; q += 32)
    if (!wait_flag(hdr + nr + q, epoch, v)) ok = false;
  ok = __all_sync(0xffffffffu, ok);
  if (!ok) return false;
  const double* slots = reinterpret_cast<const double*>(v.window) + v.lay.red_off() + (int64_t)(epoch & 1ULL) * nr * RED_W;
  double acc = 0.0;
  if (r < n) {
    acc = __ldcv(slots + r);
    for (int q = 1; q < nr; ++q) {
      const double x = __ldcv(slots + (int64_t)q * RED_W + r);
      acc = (op == 0) ? acc + x : red_combine(acc, x, op == 1 ? RED_MIN : RED_MAX);
    }
  }
  __syncwarp();
  if (r < n) buf[r] = acc;
  __syncwarp();
  return true;
}

}  // namespace cb
