// HBM stream probe for the step kernel's traffic mix: the step launch writes 300 B and reads ~60 B per env-step,
// so its ceiling is the WRITE-dominated stream rate, not the copy rate MEASURED_PEAKS.json quotes.  Prints GB/s for
//   write-only (STG.128), write-only through TMA bulk stores from shared memory (the step kernel's store path),
//   read-only (LDG.128), copy (the MEASURED_PEAKS mix), and a 5:1 write:read mix.
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bw_probe tools/bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void k_write(float4 *dst, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(dst + i, v);
}
__global__ void k_read(const float4 *src, size_t n4, float *sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) { const float4 v = __ldcs(src + i); acc += v.x + v.w; }
  if (acc == 123.456f) *sink = acc;
}
__global__ void k_copy(float4 *dst, const float4 *src, size_t n4) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) __stcs(dst + i, __ldcs(src + i));
}
// 5 parts written, 1 part read (the step kernel's mix): thread i reads src[i] and writes 5 vectors
__global__ void k_mix(float4 *dst, const float4 *src, size_t n4_read, float *sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4_read; i += stride) {
    const float4 v = __ldcs(src + i);
    acc += v.x;
#pragma unroll
    for (int k = 0; k < 5; k++) __stcs(dst + (size_t)k * n4_read + i, v);
  }
  if (acc == 123.456f) *sink = acc;
}
// each warp streams 9600-byte tiles out of shared memory with one cp.async.bulk store per tile (persistent, 4 warps per CTA)
__global__ void __launch_bounds__(128) k_write_tma(uint8_t *dst, size_t n_tiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *buf = smem + warp * 9600;
  for (int i = lane; i < 9600 / 4; i += 32) reinterpret_cast<float *>(buf)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t total = (size_t)gridDim.x * 4;
  for (size_t t = (size_t)blockIdx.x * 4 + warp; t < n_tiles; t += total) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * 9600),
                   "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(9600u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
}

int main(int argc, char **argv) {
  const size_t bytes = (argc > 1 ? atoll(argv[1]) : 4096ll) << 20;   // MiB per buffer
  const int reps = 10;
  uint8_t *a, *b; float *sink;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t n4 = bytes / 16;
  const int grid = 148 * 16;
  float ms;
#define TIME(name, moved, launch) \
  launch; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); for (int r = 0; r < reps; r++) { launch; } cudaEventRecord(e1); \
  CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); \
  printf("%-34s %8.1f GB/s\n", name, (double)(moved) * reps / (ms * 1e-3) * 1e-9);
  TIME("write-only STG.128", bytes, (k_write<<<grid, 256>>>((float4 *)a, n4)));
  CK(cudaFuncSetAttribute(k_write_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 9600));
  TIME("write-only TMA bulk 9600 B tiles", (bytes / 9600) * 9600, (k_write_tma<<<148 * 4, 128, 4 * 9600>>>(a, bytes / 9600)));
  TIME("read-only LDG.128", bytes, (k_read<<<grid, 256>>>((const float4 *)a, n4, sink)));
  TIME("copy (read + write counted)", 2 * bytes, (k_copy<<<grid, 256>>>((float4 *)b, (const float4 *)a, n4)));
  TIME("cudaMemcpyAsync D2D (r + w)", 2 * bytes, (cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)));
  TIME("cudaMemsetAsync", bytes, (cudaMemsetAsync(a, 0, bytes)));
  const size_t n4r = n4 / 5;
  TIME("mix 5 written : 1 read", 6 * n4r * 16, (k_mix<<<grid, 256>>>((float4 *)b, (const float4 *)a, n4r, sink)));
  return 0;
}
