// Multi-GPU plumbing: one process per GPU, NCCL over NVLink for the ghost-vertex halo
// exchange and the Krylov all-reduces (the reference does this with MPI inside dolfin/PETSc:
// VecGhostUpdate + MPI_Allreduce per KSP iteration).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") -- inside a Python process that
// imported torch this is the torch-bundled NCCL -- so the library has no link-time
// dependency and single-GPU use never touches NCCL.
#include <dlfcn.h>

#include <vector>

#include "common.cuh"

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

int dist_halo_exchange(cudaStream_t s, cb_dist* d, double* x, int bs, double* sendbuf) {
  if (!d || d->neigh.empty()) return CB_OK;
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
  CB_NCCL_CHECK(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, d->comm, s));
  return CB_OK;
}

}  // namespace cb

using namespace cb;

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

extern "C" int cb_dist_destroy(cb_dist* d) {
  if (!d) return CB_OK;
  if (d->comm && d->owns_comm) g_nccl.CommDestroy(d->comm);
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
  const int nop = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
  CB_NCCL_CHECK(g_nccl.AllReduce(buf_dev, buf_dev, (size_t)n, ncclFloat64, nop, d->comm, as_stream(stream)));
  return CB_OK;
}
