// Micro-benchmark: per-SMSP throughput of ex2.approx, cvt.rn.bf16x2.f32, prmt-based bf16 packing, fma, at 2 warps/SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[8];
  uint32_t acc = 0;
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i * 0.1f - 3.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); x[i] -= 1.5f; }
      if (MODE == 1) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7])); acc ^= r; x[i] += 0.25f; }
      if (MODE == 2) { uint32_t r; asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 7])); acc ^= r; x[i] -= 1.5f; }
      if (MODE == 3) { uint32_t a = __float_as_uint(x[i]) + 0x8000u, b = __float_as_uint(x[(i + 1) & 7]) + 0x8000u, r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b)); acc ^= r; x[i] += 0.25f; }
      if (MODE == 4) { x[i] = fmaf(x[i], 1.0001f, 0.5f); }
      if (MODE == 5) { uint32_t r; asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); uint32_t a = __float_as_uint(x[i]) + 0x8000u, b = __float_as_uint(x[(i + 1) & 7]) + 0x8000u; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b)); acc ^= r; x[i] -= 1.5f; }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double per = (double)h[0] / (iters * 8.0);
  int warps_per_smsp = threads / 128;
  printf("%-34s threads=%4d  cycles per (op-group, per warp) = %6.2f   per SMSP op-group issue interval = %6.2f\n", name, threads, per, per / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("ex2 (+fadd)", threads);
    run<1>("cvt.rn.bf16x2 (+xor,+fadd)", threads);
    run<2>("ex2 + cvt", threads);
    run<3>("prmt pack (+2 iadd,+xor,+fadd)", threads);
    run<4>("ffma", threads);
    run<5>("ex2 + prmt pack", threads);
  }
  return 0;
}
