// Microbenchmark: tcgen05.ld / tcgen05.st throughput (32x32b.x32: one warp moves 32 lanes x 32 columns x 4 B = 4 KB)
// with 1, 2 or 4 warps per SM sub-partition issuing concurrently. One CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tmem ubench_tmem.cu
#include <cstdio>
#include <cstdlib>

#include "../long_context_biomedical_imaging_b200/csrc/sm100_ptx.cuh"

using namespace lcbi;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

constexpr int kReps = 256;

template <bool STORE>
__global__ void __launch_bounds__(512, 1) ubench_kernel(int active_warps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = tid + i;
  tmem_st_x32(addr, r);
  tmem_st_x32(addr + 32, r);
  tmem_st_wait();
  __syncthreads();
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    t0 = clock64();
    for (int it = 0; it < kReps; it += 2) {
      if constexpr (STORE) {
        tmem_st_x32(addr, r);
        tmem_st_x32(addr + 32, r);
      } else {
        uint32_t a[32], b[32];
        tmem_ld_x32(addr, a);
        tmem_ld_x32(addr + 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i] ^ b[i]);
      }
    }
    if constexpr (STORE) tmem_st_wait();
    t1 = clock64();
  }
  if (acc == 123.456f) sink[0] = acc;
  if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* out;
  float* sink;
  CK(cudaMalloc(&out, 16));
  CK(cudaMalloc(&sink, 16));
  printf("%-6s %-6s %14s %16s\n", "op", "warps", "cyc/4KB/warp", "B/cyc/SM");
  for (int store = 0; store < 2; ++store)
    for (int w : {1, 4, 8, 16}) {
      CK(cudaMemset(out, 0, 16));
      if (store) ubench_kernel<true><<<148, 512>>>(w, out, sink);
      else ubench_kernel<false><<<148, 512>>>(w, out, sink);
      CK(cudaDeviceSynchronize());
      long long h;
      CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
      const double cyc = double(h) / kReps;
      printf("%-6s %-6d %14.1f %16.1f\n", store ? "st" : "ld", w, cyc, 4096.0 * w / cyc);
    }
  return 0;
}
