// Multi-GPU plumbing: one process per GPU, NCCL over NVLink for the ghost-vertex halo
// exchange and the Krylov all-reduces (the reference does this with MPI inside dolfin/PETSc:
// VecGhostUpdate + MPI_Allreduce per KSP iteration).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") -- inside a Python process that
// imported torch this is the torch-bundled NCCL -- so the library has no link-time
// dependency and single-GPU use never touches NCCL.
#include <dlfcn.h>
#include <stdlib.h>

#include <vector>

#include "p2p.cuh"

namespace cb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
enum { ncclFloat64 = 8 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return CB_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return CB_ERR_NCCL;
  }
#define CB_SYM(field, name)                                             \
  *(void**)(&g_nccl.field) = dlsym(h, name);                            \
  if (!g_nccl.field) { set_error("NCCL symbol %s missing", name); return CB_ERR_NCCL; }
  CB_SYM(GetUniqueId, "ncclGetUniqueId");
  CB_SYM(CommInitRank, "ncclCommInitRank");
  CB_SYM(CommDestroy, "ncclCommDestroy");
  CB_SYM(AllReduce, "ncclAllReduce");
  CB_SYM(Send, "ncclSend");
  CB_SYM(Recv, "ncclRecv");
  CB_SYM(GroupStart, "ncclGroupStart");
  CB_SYM(GroupEnd, "ncclGroupEnd");
  CB_SYM(GetErrorString, "ncclGetErrorString");
#undef CB_SYM
  g_nccl.handle = h;
  return CB_OK;
}

#define CB_NCCL_CHECK(expr)                                                               \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != 0) {                                                                        \
      cb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return CB_ERR_NCCL;                                                                 \
    }                                                                                     \
  } while (0)

}  // namespace cb

struct cb_dist {
  cb::ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  std::vector<int> neigh;
  std::vector<int64_t> send_off, recv_off;
  const int32_t* send_idx = nullptr;  // device
  int64_t n_owned = 0, n_ghost = 0;
  bool owns_comm = true;               // false for plans cloned from a parent (AMG coarse levels)
  // ---- one-sided transport over NVLink peer memory (cb_dist_p2p_*), optional; window layout: p2p.cuh
  bool p2p = false;
  // Error word of the communicator: pinned, mapped host memory owned by the root plan and shared by every plan cloned
  // from it, so the host reads it without a copy or a synchronisation after any exchange.
  unsigned long long* error_host = nullptr;
  unsigned long long* error_dev = nullptr;
  unsigned long long timeout_cycles = 0;
  cb_dist* root = nullptr;                         // plan that owns communicator and error word (NULL: this one)
  unsigned char* window = nullptr;                 // my window (cudaMalloc, exported by IPC)
  std::vector<unsigned char*> peer_window;         // [nranks] mapped peer windows (self = window)
  int64_t ghost_cap = 0;                           // doubles per parity of the landing zone
  // device copies of the plan for the put kernel
  int64_t* d_send_off = nullptr;                   // [nneigh+1]
  int64_t* d_remote_off = nullptr;                 // [nneigh] node offset of my data in the peer's ghost range
  int64_t* d_remote_cap = nullptr;                 // [nneigh] parity stride (doubles) of the peer's landing zone
  int64_t* d_recv_off = nullptr;                   // [nneigh+1]
  int32_t* d_neigh = nullptr;                      // [nneigh]
  unsigned char** d_peer_window = nullptr;         // [nranks]
};

namespace cb {

template <int BS>
__global__ void k_pack(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ x,
                       double* __restrict__ buf) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * BS;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / BS;
    const int c = (int)(i - v * BS);
    buf[i] = x[(int64_t)idx[v] * BS + c];
  }
}

int64_t dist_send_len(const cb_dist* d) { return d ? d->send_off.back() : 0; }
bool dist_is_one_sided(const cb_dist* d);

// ------------------------------------------------------------------ one-sided NVLink transport (device side: p2p.cuh)
// Halo exchange as ONE kernel.  Put phase: gather the boundary values and store them straight into the neighbours'
// landing zones over NVLink; the last block to finish releases the per-source epoch flags.  Get phase: every block waits
// for the neighbours' flags of this epoch and copies its share of the landing zone behind the owned part of x.  The
// grid is at most 2 blocks per SM, so all blocks are co-resident and a block spinning on a flag cannot starve the
// block that has to release ours.
template <int BS>
__global__ void __launch_bounds__(256)
k_halo_exchange(HaloDev h, double* __restrict__ x) {
  p2p_halo_exchange_grid<BS>(h, x);
}

// all-reduce of n <= RED_W doubles as one kernel of one warp (p2p_allreduce_warp)
__global__ void k_allreduce(int n, double* __restrict__ buf, P2PView v, int op) {
  p2p_allreduce_warp(v, n, buf, op);
}

P2PView dist_p2p_view(const cb_dist* d) {
  P2PView v;
  v.peer_window = d->d_peer_window;
  v.window = d->window;
  v.error = d->error_dev;
  v.timeout_cycles = d->timeout_cycles;
  v.lay = WindowLayout{d->nranks};
  v.rank = d->rank;
  return v;
}

HaloDev dist_halo_dev(const cb_dist* d) {
  HaloDev h;
  h.nneigh = (int)d->neigh.size();
  h.send_off = d->d_send_off;
  h.remote_off = d->d_remote_off;
  h.neigh = d->d_neigh;
  h.send_idx = d->send_idx;
  h.remote_cap = d->d_remote_cap;
  h.recv_off = d->d_recv_off;
  h.n_owned = d->n_owned;
  h.ghost_cap = d->ghost_cap;
  h.v = dist_p2p_view(d);
  return h;
}

// largest block size (doubles per node) a plan's landing zone holds
int64_t dist_ghost_capacity(const cb_dist* d) { return d ? d->ghost_cap : 0; }
int64_t dist_recv_len(const cb_dist* d) { return d ? d->recv_off.back() : 0; }

static int p2p_halo_exchange(cudaStream_t s, cb_dist* d, double* x, int bs) {
  const int64_t nsend = d->send_off.back(), nrecv = d->recv_off.back();
  if (nrecv * bs > d->ghost_cap) {
    set_error("p2p halo exchange: landing zone too small");
    return CB_ERR_ARG;
  }
  const int64_t work = (nsend > nrecv ? nsend : nrecv) * bs;
  const int grid = grid_for(work > 0 ? work : 1, 256, 2);
  const HaloDev h = dist_halo_dev(d);
  if (bs == 1) k_halo_exchange<1><<<grid, 256, 0, s>>>(h, x);
  else if (bs == 2) k_halo_exchange<2><<<grid, 256, 0, s>>>(h, x);
  else k_halo_exchange<3><<<grid, 256, 0, s>>>(h, x);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

static int p2p_allreduce(cudaStream_t s, cb_dist* d, double* buf, int n, int op) {
  k_allreduce<<<1, 32, 0, s>>>(n, buf, dist_p2p_view(d), op);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

bool dist_is_one_sided(const cb_dist* d) { return d && d->p2p; }

int dist_halo_exchange(cudaStream_t s, cb_dist* d, double* x, int bs, double* sendbuf) {
  if (!d || d->neigh.empty()) return CB_OK;
  if (d->p2p) return p2p_halo_exchange(s, d, x, bs);
  const int64_t nsend = d->send_off.back();
  if (nsend > 0) {
    const int grid = grid_for(nsend * bs, 256, 4);
    if (bs == 1) k_pack<1><<<grid, 256, 0, s>>>(nsend, d->send_idx, x, sendbuf);
    else if (bs == 2) k_pack<2><<<grid, 256, 0, s>>>(nsend, d->send_idx, x, sendbuf);
    else k_pack<3><<<grid, 256, 0, s>>>(nsend, d->send_idx, x, sendbuf);
    CB_LAUNCH_CHECK();
  }
  CB_NCCL_CHECK(g_nccl.GroupStart());
  for (size_t j = 0; j < d->neigh.size(); ++j) {
    const int64_t ns = d->send_off[j + 1] - d->send_off[j];
    const int64_t nr = d->recv_off[j + 1] - d->recv_off[j];
    if (ns > 0)
      CB_NCCL_CHECK(g_nccl.Send(sendbuf + d->send_off[j] * bs, (size_t)(ns * bs), ncclFloat64, d->neigh[j], d->comm, s));
    if (nr > 0)
      CB_NCCL_CHECK(g_nccl.Recv(x + (d->n_owned + d->recv_off[j]) * bs, (size_t)(nr * bs), ncclFloat64, d->neigh[j], d->comm, s));
  }
  CB_NCCL_CHECK(g_nccl.GroupEnd());
  return CB_OK;
}

int dist_allreduce_sum(cudaStream_t s, cb_dist* d, double* buf, int n) {
  if (!d) return CB_OK;
  if (d->p2p && n <= RED_W && d->nranks <= 32) return p2p_allreduce(s, d, buf, n, 0);
  CB_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, d->comm, s));
  return CB_OK;
}

}  // namespace cb

using namespace cb;

// The communicator's error word (pinned + mapped: device kernels raise it, the host polls it for free) and the wait
// limit of the one-sided kernels: CASHOCS_B200_P2P_TIMEOUT_S seconds (default 60) of SM clock.
static int dist_init_error_word(cb_dist* d) {
  CB_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&d->error_host), sizeof(unsigned long long), cudaHostAllocMapped));
  *d->error_host = 0ULL;
  CB_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d->error_dev), d->error_host, 0));
  double seconds = 60.0;
  if (const char* env = getenv("CASHOCS_B200_P2P_TIMEOUT_S")) {
    const double v = atof(env);
    if (v > 0.0) seconds = v;
  }
  int dev = 0, khz = 0;
  CB_CUDA_CHECK(cudaGetDevice(&dev));
  CB_CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  if (khz <= 0) khz = 1965000;
  d->timeout_cycles = (unsigned long long)(seconds * 1e3 * (double)khz);
  return CB_OK;
}

extern "C" int cb_dist_unique_id(char* id128) {
  if (int e = load_nccl()) return e;
  ncclUniqueId id;
  CB_NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return CB_OK;
}

extern "C" int cb_dist_create(cb_dist** out, const char* id128, int rank, int nranks, int nneigh,
                              const int32_t* neigh_rank, const int64_t* send_off,
                              const int32_t* send_idx_dev, const int64_t* recv_off, int64_t n_owned,
                              int64_t n_ghost) {
  CB_REQUIRE(out && id128, "cb_dist_create: missing argument");
  if (int e = load_nccl()) return e;
  cb_dist* d = new cb_dist();
  d->rank = rank;
  d->nranks = nranks;
  d->neigh.assign(neigh_rank, neigh_rank + nneigh);
  d->send_off.assign(send_off, send_off + nneigh + 1);
  d->recv_off.assign(recv_off, recv_off + nneigh + 1);
  d->send_idx = send_idx_dev;
  d->n_owned = n_owned;
  d->n_ghost = n_ghost;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&d->comm, nranks, id, rank);
  if (r != 0) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete d;
    return CB_ERR_NCCL;
  }
  if (int e = dist_init_error_word(d)) { delete d; return e; }
  *out = d;
  return CB_OK;
}

extern "C" int cb_dist_clone_plan(cb_dist* parent, cb_dist** out, int nneigh, const int32_t* neigh_rank,
                                  const int64_t* send_off, const int32_t* send_idx_dev,
                                  const int64_t* recv_off, int64_t n_owned, int64_t n_ghost) {
  CB_REQUIRE(parent && out, "cb_dist_clone_plan: missing argument");
  cb_dist* d = new cb_dist();
  d->comm = parent->comm;
  d->owns_comm = false;
  d->root = parent->root ? parent->root : parent;
  d->error_host = d->root->error_host;
  d->error_dev = d->root->error_dev;
  d->timeout_cycles = d->root->timeout_cycles;
  d->rank = parent->rank;
  d->nranks = parent->nranks;
  d->neigh.assign(neigh_rank, neigh_rank + nneigh);
  d->send_off.assign(send_off, send_off + nneigh + 1);
  d->recv_off.assign(recv_off, recv_off + nneigh + 1);
  d->send_idx = send_idx_dev;
  d->n_owned = n_owned;
  d->n_ghost = n_ghost;
  *out = d;
  return CB_OK;
}

// Allocate this plan's window and return its 64-byte CUDA IPC handle (host layer all-gathers them).
extern "C" int cb_dist_p2p_export(cb_dist* d, char* handle64) {
  CB_REQUIRE(d && handle64, "cb_dist_p2p_export: missing argument");
  d->ghost_cap = ((d->n_ghost * 3 + 7) / 8) * 8 + 8;
  const WindowLayout lay{d->nranks};
  const size_t bytes = sizeof(double) * (size_t)lay.total_doubles(d->ghost_cap);
  CB_CUDA_CHECK(cudaMalloc(&d->window, bytes));
  CB_CUDA_CHECK(cudaMemset(d->window, 0, bytes));
  CB_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CB_CUDA_CHECK(cudaIpcGetMemHandle(&h, d->window));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return CB_OK;
}

extern "C" int64_t cb_dist_p2p_ghost_cap(cb_dist* d) { return d ? d->ghost_cap : 0; }

static void dist_release_p2p(cb_dist* d);

// Map the peers' windows (handles [nranks*64], rank order) and switch this plan to the one-sided
// transport.  remote_off [nneigh]: node offset of my data inside neighbour j's ghost range.
extern "C" int cb_dist_p2p_open(cb_dist* d, const char* handles, const int64_t* remote_off,
                                const int64_t* remote_cap) {
  CB_REQUIRE(d && handles && d->window, "cb_dist_p2p_open: export the window first");
  const int nneigh = (int)d->neigh.size();
  CB_REQUIRE(nneigh == 0 || (remote_off && remote_cap), "cb_dist_p2p_open: remote offsets missing");
  d->peer_window.assign(d->nranks, nullptr);
  for (int r = 0; r < d->nranks; ++r) {
    if (r == d->rank) { d->peer_window[r] = d->window; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      (void)cudaGetLastError();
      dist_release_p2p(d);
      return CB_ERR_CUDA;
    }
    d->peer_window[r] = static_cast<unsigned char*>(ptr);
  }
  CB_CUDA_CHECK(cudaMalloc(&d->d_send_off, sizeof(int64_t) * (nneigh + 1)));
  CB_CUDA_CHECK(cudaMalloc(&d->d_recv_off, sizeof(int64_t) * (nneigh + 1)));
  CB_CUDA_CHECK(cudaMalloc(&d->d_remote_off, sizeof(int64_t) * (nneigh + 1)));
  CB_CUDA_CHECK(cudaMalloc(&d->d_remote_cap, sizeof(int64_t) * (nneigh + 1)));
  CB_CUDA_CHECK(cudaMalloc(&d->d_neigh, sizeof(int32_t) * (nneigh + 1)));
  CB_CUDA_CHECK(cudaMalloc(&d->d_peer_window, sizeof(unsigned char*) * d->nranks));
  CB_CUDA_CHECK(cudaMemcpy(d->d_send_off, d->send_off.data(), sizeof(int64_t) * (nneigh + 1), cudaMemcpyHostToDevice));
  CB_CUDA_CHECK(cudaMemcpy(d->d_recv_off, d->recv_off.data(), sizeof(int64_t) * (nneigh + 1), cudaMemcpyHostToDevice));
  if (nneigh) {
    CB_CUDA_CHECK(cudaMemcpy(d->d_remote_off, remote_off, sizeof(int64_t) * nneigh, cudaMemcpyHostToDevice));
    CB_CUDA_CHECK(cudaMemcpy(d->d_remote_cap, remote_cap, sizeof(int64_t) * nneigh, cudaMemcpyHostToDevice));
    std::vector<int32_t> ng(d->neigh.begin(), d->neigh.end());
    CB_CUDA_CHECK(cudaMemcpy(d->d_neigh, ng.data(), sizeof(int32_t) * nneigh, cudaMemcpyHostToDevice));
  }
  CB_CUDA_CHECK(cudaMemcpy(d->d_peer_window, d->peer_window.data(), sizeof(unsigned char*) * d->nranks, cudaMemcpyHostToDevice));
  d->p2p = true;
  return CB_OK;
}

// Back to NCCL for this plan (collective decision of the host layer when a rank could not map a peer).
extern "C" int cb_dist_p2p_disable(cb_dist* d) {
  if (d) d->p2p = false;
  return CB_OK;
}

// 1 if a peer timed out inside a one-sided wait on ANY plan of this communicator (results are then invalid), else 0.
// A plain host load of the pinned error word: no copy, no synchronisation -- cheap enough to be called after every
// exchange.  The word stays raised: a timed-out communicator is dead, the host layer raises.
extern "C" int cb_dist_p2p_error(cb_dist* d) {
  if (!d || !d->error_host) return 0;
  return *reinterpret_cast<volatile unsigned long long*>(d->error_host) != 0ULL;
}

static void dist_release_p2p(cb_dist* d) {
  for (size_t r = 0; r < d->peer_window.size(); ++r)
    if ((int)r != d->rank && d->peer_window[r]) cudaIpcCloseMemHandle(d->peer_window[r]);
  d->peer_window.clear();
  if (d->d_send_off) cudaFree(d->d_send_off);
  if (d->d_recv_off) cudaFree(d->d_recv_off);
  if (d->d_remote_off) cudaFree(d->d_remote_off);
  if (d->d_remote_cap) cudaFree(d->d_remote_cap);
  if (d->d_neigh) cudaFree(d->d_neigh);
  if (d->d_peer_window) cudaFree(d->d_peer_window);
  d->d_send_off = d->d_recv_off = d->d_remote_off = d->d_remote_cap = nullptr;
  d->d_neigh = nullptr;
  d->d_peer_window = nullptr;
}

extern "C" int cb_dist_destroy(cb_dist* d) {
  if (!d) return CB_OK;
  dist_release_p2p(d);  // by pointer, whether or not the one-sided transport ended up enabled
  if (d->window) cudaFree(d->window);
  if (d->comm && d->owns_comm) g_nccl.CommDestroy(d->comm);
  if (d->error_host && !d->root) cudaFreeHost(d->error_host);
  delete d;
  return CB_OK;
}

extern "C" int cb_dist_halo_exchange(void* stream, cb_dist* d, double* x, int bs, double* sendbuf) {
  CB_REQUIRE(d && x && sendbuf, "cb_dist_halo_exchange: missing argument");
  CB_REQUIRE(bs >= 1 && bs <= 3, "cb_dist_halo_exchange: bs must be 1..3");
  return dist_halo_exchange(as_stream(stream), d, x, bs, sendbuf);
}

extern "C" int cb_dist_allreduce(void* stream, cb_dist* d, double* buf_dev, int n, int op) {
  CB_REQUIRE(d && buf_dev, "cb_dist_allreduce: missing argument");
  if (d->p2p && n <= RED_W && d->nranks <= 32) return p2p_allreduce(as_stream(stream), d, buf_dev, n, op);
  const int nop = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
  CB_NCCL_CHECK(g_nccl.AllReduce(buf_dev, buf_dev, (size_t)n, ncclFloat64, nop, d->comm, as_stream(stream)));
  return CB_OK;
}
