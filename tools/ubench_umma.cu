// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1, M = 128, K = 16) as a function of N, operand
// source (A from smem or TMEM), operand majorness and accumulator dependence. One CTA per SM, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_umma ubench_umma.cu
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

constexpr int kOuter = 64;

template <int N, bool TS, bool AMN, bool BMN, int NACC, int KS, bool BG, bool COMMIT = false>
__global__ void __launch_bounds__(256, 1) ubench_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, sink_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&sink_bar, 1u << 19);   // absorbs the per-group commits of the COMMIT variants, never completes
    fence_mbar_init();
    stop_flag = 0;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, N, AMN ? 1 : 0, BMN ? 1 : 0);
      const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 65536);
      // K-major SW128: k-step advance 32 B; MN-major SW128: 2048 B per k-step, LBO = 16 KB between 64-element MN blocks
      const uint64_t da0 = AMN ? make_smem_desc(a_addr, 16384, 1024, kLayoutSW128) : make_smem_desc(a_addr, 16, 1024, kLayoutSW128);
      const uint64_t db0 = BMN ? make_smem_desc(b_addr, 16384, 1024, kLayoutSW128) : make_smem_desc(b_addr, 16, 1024, kLayoutSW128);
      constexpr uint32_t a_step = AMN ? 2048 : 32, b_step = BMN ? 2048 : 32;
      const long long t0 = clock64();
      for (int it = 0; it < kOuter; ++it) {
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc) {
          const uint32_t d = tmem + (acc * N) % 448;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            const uint64_t da = da0 + ((kk * a_step) >> 4), db = db0 + ((kk * b_step) >> 4);
            if constexpr (TS) umma_ts(d, tmem + 448 + (kk & 3) * 8, db, idesc, 1u);
            else umma_ss(d, da, db, idesc, 1u);
          }
          if constexpr (COMMIT) umma_commit(&sink_bar);
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      stop_flag = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  } else if (BG && warp >= 4) {
    // background shared-memory store traffic in a region the MMAs do not read
    uint8_t* dst = smem + 98304 + (tid & 127) * 16;
    while (!stop_flag) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(dst + i * 2048)), "r"(i) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, bool TS, bool AMN, bool BMN, int NACC, int KS, bool BG, bool COMMIT = false>
void run(long long* out) {
  const int smem_bytes = 128 * 1024 + 1024;
  auto kern = ubench_kernel<N, TS, AMN, BMN, NACC, KS, BG, COMMIT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CK(cudaMemset(out, 0, 16));
  kern<<<148, 256, smem_bytes>>>(out);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
  const double n = double(kOuter) * NACC * KS;
  printf("%-6d %-3d %-4d %-4d %-4d %-3d %-3d %-3d %10.1f %10.1f\n", N, int(TS), int(AMN), int(BMN), NACC, KS, int(BG), int(COMMIT),
         double(h[0]) / n, double(h[1]) / n);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* out;
  CK(cudaMalloc(&out, 16));
  printf("%-6s %-3s %-4s %-4s %-4s %-3s %-3s  %10s %10s\n", "N", "TS", "Amn", "Bmn", "nacc", "ks", "bg c", "cyc/issue", "cyc/mma");
  //   N   TS     Amn    Bmn   nacc ks bg
  run<64, false, false, false, 1, 4, false>(out);    // SS, dependent chain
  run<128, false, false, false, 1, 4, false>(out);
  run<256, false, false, false, 1, 4, false>(out);
  run<64, false, false, false, 2, 4, false>(out);    // SS, two accumulators
  run<128, false, false, false, 2, 4, false>(out);
  run<64, true, false, false, 1, 4, false>(out);     // TS, dependent chain
  run<128, true, false, false, 1, 4, false>(out);
  run<256, true, false, false, 1, 4, false>(out);
  run<64, true, false, false, 2, 4, false>(out);     // TS, two accumulators
  run<128, true, false, false, 2, 4, false>(out);
  run<64, true, false, false, 4, 4, false>(out);     // TS, four accumulators
  run<64, true, false, false, 4, 1, false>(out);     // TS, four accumulators, a different one every instruction
  run<64, true, false, true, 2, 4, false>(out);      // TS, B MN-major (dV / dK / PV)
  run<128, true, false, true, 2, 4, false>(out);
  run<64, false, true, true, 1, 8, false>(out);      // SS, A and B MN-major (dQ)
  run<128, false, true, true, 1, 8, false>(out);
  run<64, false, false, false, 2, 4, true>(out);     // with background smem stores
  run<64, true, false, false, 2, 4, true>(out);
  run<128, true, false, false, 2, 4, true>(out);
  run<64, true, false, false, 2, 4, false, true>(out);     // a commit after every 4 k-steps
  run<64, false, false, false, 2, 4, false, true>(out);
  run<64, true, false, false, 4, 1, false, true>(out);     // a commit after every instruction
  run<8, true, false, false, 2, 4, false>(out);      // tiny N: exposes the per-instruction issue cost
  run<16, true, false, false, 2, 4, false>(out);
  run<32, true, false, false, 2, 4, false>(out);
  run<8, false, false, false, 2, 4, false>(out);
  run<32, false, false, false, 2, 4, false>(out);
  run<8, true, false, false, 2, 4, false, true>(out);
  return 0;
}
