#include "common.cuh"

#include <cstdlib>

namespace gfs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

cudaError_t stream_wait(cudaStream_t st) {
  static const int mode = [] {
    const char* e = getenv("GFS_SYNC");
    return (e && (e[0] == 'b' || e[0] == 'B')) ? 1 : 0;
  }();
  if (mode == 0) return cudaStreamSynchronize(st);
  // one blocking-sync event per (thread, device): events belong to the device that was current when they were made
  struct Ev { cudaEvent_t e = nullptr; int dev = -1; };
  static thread_local Ev ev;
  int dev = 0;
  cudaError_t rc = cudaGetDevice(&dev);
  if (rc != cudaSuccess) return rc;
  if (!ev.e || ev.dev != dev) {
    if (ev.e) cudaEventDestroy(ev.e);
    ev.e = nullptr;
    rc = cudaEventCreateWithFlags(&ev.e, cudaEventBlockingSync | cudaEventDisableTiming);
    if (rc != cudaSuccess) return rc;
    ev.dev = dev;
  }
  rc = cudaEventRecord(ev.e, st);
  if (rc != cudaSuccess) return rc;
  return cudaEventSynchronize(ev.e);
}

bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int PinnedBuf::reserve(size_t n) {
  if (n <= cap) return GFS_OK;
  release();
  GFS_CUDA(cudaMallocHost(&p, n));
  cap = n;
  return GFS_OK;
}
void PinnedBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}
int DevBuf::reserve(size_t n) {
  if (n <= cap) return GFS_OK;
  release();
  GFS_CUDA(cudaMalloc(&p, n));
  cap = n;
  return GFS_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

}  // namespace gfs

extern "C" {

const char* gfs_last_error(void) { return gfs::g_err; }

int gfs_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    gfs::set_error("no CUDA device (%s); libgfs_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    return GFS_ERR_NODEVICE;
  }
  return GFS_OK;
}

const char* gfs_version(void) { return "gfs_b200 0.1 sm_100a"; }
}
