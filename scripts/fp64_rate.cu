// Measures sustained DFMA and FFMA throughput (independent chains) -- context for the fp64 GICP / BA kernels.
#include <cstdio>
template <class T> __global__ void k(T* out, int iters) {
  T a[8]; for (int i = 0; i < 8; i++) a[i] = (T)(threadIdx.x + i);
  const T b = (T)1.0000001, c = (T)0.5;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a[i] * b + c;
  T s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class T> double run(const char* name) {
  T* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(T));
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<T><<<148 * 8, 256>>>(d, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<T><<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 8 * iters * 148.0 * 8 * 256;
  printf("%s: %.2f TFLOP/s (%.3f ms)\n", name, flops / ms / 1e9, ms);
  cudaFree(d); return flops / ms / 1e9;
}
int main() { run<float>("fp32 FFMA"); run<double>("fp64 DFMA"); return 0; }
