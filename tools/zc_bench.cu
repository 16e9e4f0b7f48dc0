// zc_bench.cu -- zero-copy (pinned host memory read by a kernel over PCIe) bandwidth vs. per-lane access width, next to
// cudaMemcpyAsync of the same buffer.  Decides how mgplr_step_env_host should read its int64 actions.
//   nvcc -O2 -o tools/zc_bench tools/zc_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <typename T>
__global__ void k_read(const T *__restrict__ src, size_t n, unsigned long long *sink) {
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const T v = src[i];
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&v);
    for (int k = 0; k < (int)(sizeof(T) / 8); k++) acc += w[k];
  }
  if (acc == 0x1234567ull) *sink = acc;
}

int main() {
  const size_t bytes = 4u << 20;  // 524 288 int64 actions
  void *h, *d; unsigned long long *sink;
  CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped));
  CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&sink, 8));
  for (size_t i = 0; i < bytes / 8; i++) ((uint64_t *)h)[i] = i % 7;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms;
  for (int grid : {148, 592, 2368}) {
    for (int width : {8, 16}) {
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        for (int k = 0; k < 20; k++) {
          if (width == 8) k_read<uint64_t><<<grid, 128>>>((const uint64_t *)h, bytes / 8, sink);
          else k_read<ulonglong2><<<grid, 128>>>((const ulonglong2 *)h, bytes / 16, sink);
        }
        cudaEventRecord(b); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, a, b);
      }
      printf("zero-copy read, %2d B per lane, grid %4d x 128: %.1f us per 4 MB = %.1f GB/s\n", width, grid, ms * 1e3 / 20, bytes / (ms * 1e-3 / 20) * 1e-9);
    }
  }
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(a);
    for (int k = 0; k < 20; k++) CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0));
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b);
  }
  printf("cudaMemcpyAsync H2D: %.1f us per 4 MB = %.1f GB/s\n", ms * 1e3 / 20, bytes / (ms * 1e-3 / 20) * 1e-9);
  return 0;
}
