#include "common.cuh"

namespace gfs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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
