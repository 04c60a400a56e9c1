// Canonical CUDA-guide style probe: 2-D tensor, element type / box / swizzle from argv, direct __grid_constant__ param.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <vector>
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes, uint8_t* out) {
  __shared__ alignas(1024) uint8_t sm[16384];
  __shared__ alignas(8) unsigned long long bar_;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(sm)), "l"(&tm), "r"(x), "r"(y), "r"(bar) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  const int es_ = atoi(argv[1]), boxW = atoi(argv[2]), boxH = atoi(argv[3]);   // element size 1 / 4, box in elements
  const int W = 1024, H = 256;
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  std::vector<uint8_t> h((size_t)W * H * es_);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t st[1] = {(cuuint64_t)W * es_};
  cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH}, es[2] = {1, 1};
  CUresult r = ((Enc)f)(&tm, es_ == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int bytes = boxW * boxH * es_;
  printf("elem %d box %dx%d (%d B) encode -> %d\n", es_, boxW, boxH, bytes, (int)r);
  const int X = argc > 4 ? atoi(argv[4]) : 16;
  k<<<1, 128>>>(tm, X, 8, bytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  run: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<uint8_t> rr(bytes); cudaMemcpy(rr.data(), o, rr.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int yy = 0; yy < boxH; yy++) for (int xx = 0; xx < boxW * es_; xx++) bad += rr[yy * boxW * es_ + xx] != h[((size_t)(8 + yy) * W + X) * es_ + xx];
  printf("  mismatches %d\n", bad);
  return 0;
}
