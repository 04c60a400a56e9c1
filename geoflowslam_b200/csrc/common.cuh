// Shared helpers for the libgfs_b200 CUDA sources.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/gfs_b200.h"

namespace gfs {

void set_error(const char* fmt, ...);

#define GFS_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      gfs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));          \
      return GFS_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

#define GFS_REQUIRE(cond, code, msg)                \
  do {                                              \
    if (!(cond)) {                                  \
      gfs::set_error("%s: %s", __func__, msg);      \
      return code;                                  \
    }                                               \
  } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Host wait for everything enqueued on `st`.  GFS_SYNC=spin (default): cudaStreamSynchronize, the CUDA runtime's spin wait
// (lowest latency; right when the process has cores to itself).  GFS_SYNC=block: a blocking-sync event -- the thread sleeps, which
// keeps eight ranks x four host threads from spinning on every logical core of the box (bench.py sets it for multi-rank runs).
cudaError_t stream_wait(cudaStream_t st);

// true when p is device-accessible pinned host memory (so async copies need no staging)
bool is_pinned_host(const void* p);

// Pinned staging buffer that grows on demand.
struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n);
  void release();
};
// Device buffer that grows on demand.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n);
  void release();
};

}  // namespace gfs
