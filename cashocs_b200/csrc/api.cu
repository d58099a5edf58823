// Error reporting and device queries of the C-ABI.
#include <stdarg.h>

#include "common.cuh"

namespace cb {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launch_count{0};
thread_local long long t_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// cudaMallocAsync scratch is returned to the driver at every synchronisation unless the pool's
// release threshold is raised; keep it so repeated calls reuse the same blocks.
int configure_scratch_pool() {
  static int done_for_device = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return CB_ERR_CUDA;
  if (done_for_device == dev) return CB_OK;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess) return CB_ERR_CUDA;
  unsigned long long threshold = ~0ULL;
  if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold) != cudaSuccess) return CB_ERR_CUDA;
  done_for_device = dev;
  return CB_OK;
}

int num_sms() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

}  // namespace cb

extern "C" const char* cb_last_error(void) { return cb::g_err; }
extern "C" int cb_version(void) { return 200; }
// sizeof() of the structs of the header as THIS library was compiled: a binding (ctypes / cgo / JNI) checks its own
// mirror against these before the first call instead of trusting a hand-counted constant
extern "C" int64_t cb_struct_size(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(cb_poly);
    case 1: return (int64_t)sizeof(cb_quad);
    case 2: return (int64_t)sizeof(cb_mat);
    case 3: return (int64_t)sizeof(cb_mesh_view);
    case 4: return (int64_t)sizeof(cb_ksp_opts);
    case 5: return (int64_t)sizeof(cb_amg_level);
    case 6: return (int64_t)sizeof(cb_amg);
    case 7: return (int64_t)sizeof(cb_ksp_result);
    case 8: return (int64_t)sizeof(cb_p2_view);
  }
  return -1;
}
extern "C" int cb_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cb::set_error("no CUDA device"); return CB_ERR_CUDA; }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cb::set_error("cudaDeviceGetAttribute failed"); return CB_ERR_CUDA; }
  return n;
}
extern "C" long long cb_launch_count(void) { return cb::g_launch_count.load(); }
