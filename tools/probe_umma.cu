// Standalone sm_100a probe: validates the UMMA/TMA/TMEM conventions the attention kernels rely on.
//   v1: D[128x128] = A[128x64] * B[128x64]^T        (both K-major, SW128, via TMA)       -> S = Q K^T
//   v2: D[128x64]  = A[128x128] * B[128(k)x64(n)]   (A K-major 2 atoms, B MN-major)      -> O = P V (P in smem)
//   v3: same as v2 with A (bf16) stored to TMEM by tcgen05.st and a TS MMA                -> O = P V (P in TMEM)
//   v4: D[128x64]  = AT[128(k)x128(m)]^T * B[128(k)x64(n)]  (A MN-major, B MN-major)      -> dQ = dS K
//   v5: v1 with A written to smem by threads using the SW128 address swizzle (no TMA)
//   v6: v1 with N=64 accumulate chain: D = A*B^T issued twice (second with accumulate=1) -> 2x
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_umma probe_umma.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../long_context_biomedical_imaging_b200/csrc/sm100_ptx.cuh"
#include "../long_context_biomedical_imaging_b200/csrc/tma_host.h"

using namespace lcbi;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

struct ProbeArgs {
  int variant;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;  // bytes
  int n_out;                             // 128 or 64
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __nv_bfloat16* __restrict__ Araw, float* __restrict__ D, ProbeArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;           // 32 KB
  uint8_t* sB = smem + 32768;   // 32 KB
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int v = args.variant;

  // ---------------- operand staging ----------------
  if (tid == 0) {
    uint32_t bytes = 0;
    if (v == 1 || v == 6) {
      tma_load_2d(sA, &tmA, &bar_tma, 0, 0); bytes += 16384;
      tma_load_2d(sB, &tmB, &bar_tma, 0, 0); bytes += (v == 6) ? 8192 : 16384;
    } else if (v == 2 || v == 4) {
      tma_load_2d(sA, &tmA, &bar_tma, 0, 0);
      tma_load_2d(sA + 16384, &tmA, &bar_tma, 64, 0); bytes += 32768;
      tma_load_2d(sB, &tmB, &bar_tma, 0, 0); bytes += 16384;
    } else if (v == 3 || v == 5) {
      tma_load_2d(sB, &tmB, &bar_tma, 0, 0); bytes += 16384;
    }
    mbar_expect_tx(&bar_tma, bytes);
  }
  if (v == 3) {
    // thread t owns row t of A[128x128]; pack pairs and store to TMEM columns [64,128)
    uint32_t regs[64];
    const __nv_bfloat16* row = Araw + tid * 128;
    for (int j = 0; j < 64; ++j) {
      uint32_t lo = __bfloat16_as_ushort(row[2 * j]), hi = __bfloat16_as_ushort(row[2 * j + 1]);
      regs[j] = lo | (hi << 16);
    }
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64;
    tmem_st_x32(taddr, regs);
    tmem_st_x32(taddr + 32, regs + 32);
    tmem_st_wait();
  }
  if (v == 5) {
    // A[128x64] K-major written by threads: thread t owns row t
    const __nv_bfloat16* row = Araw + tid * 64;
    for (int c16 = 0; c16 < 8; ++c16) {
      uint4 val = *reinterpret_cast<const uint4*>(row + c16 * 8);
      *reinterpret_cast<uint4*>(sA + sw128_offset(tid, c16)) = val;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---------------- MMA issue ----------------
  if (tid == 0) {
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (v == 1 || v == 5 || v == 6) {
      const uint32_t N = (v == 6) ? 64 : 128;
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const int reps = (v == 6) ? 2 : 1;
      for (int rep = 0; rep < reps; ++rep)
        for (int k = 0; k < 4; ++k) {
          uint64_t da = make_smem_desc(a0 + k * 32, args.lbo_a, args.sbo_a, kLayoutSW128);
          uint64_t db = make_smem_desc(b0 + k * 32, args.lbo_b, args.sbo_b, kLayoutSW128);
          umma_ss(tmem, da, db, idesc, (k > 0 || rep > 0) ? 1u : 0u);
        }
    } else if (v == 2) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
      for (int k = 0; k < 8; ++k) {
        uint64_t da = make_smem_desc(a0 + (k >> 2) * 16384 + (k & 3) * 32, args.lbo_a, args.sbo_a, kLayoutSW128);
        uint64_t db = make_smem_desc(b0 + k * 2048, args.lbo_b, args.sbo_b, kLayoutSW128);
        umma_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
      }
    } else if (v == 3) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
      for (int k = 0; k < 8; ++k) {
        uint64_t db = make_smem_desc(b0 + k * 2048, args.lbo_b, args.sbo_b, kLayoutSW128);
        umma_ts(tmem, tmem + 64 + k * 8, db, idesc, k > 0 ? 1u : 0u);
      }
    } else if (v == 4) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int k = 0; k < 8; ++k) {
        uint64_t da = make_smem_desc(a0 + k * 2048, args.lbo_a, args.sbo_a, kLayoutSW128);
        uint64_t db = make_smem_desc(b0 + k * 2048, args.lbo_b, args.sbo_b, kLayoutSW128);
        umma_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  // ---------------- read back D ----------------
  const int ncols = args.n_out;
  for (int c = 0; c < ncols; c += 32) {
    uint32_t r[32];
    tmem_ld_x32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[tid * ncols + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

static int make2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t bc, uint32_t br) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {cols * 2};
  uint32_t box[2] = {bc, br};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  srand(1);
  // generic buffers: A up to 128x128, B up to 128x128
  std::vector<float> A(128 * 128), B(128 * 128);
  for (auto& x : A) x = bf16r((rand() % 2001 - 1000) / 1000.0f);
  for (auto& x : B) x = bf16r((rand() % 2001 - 1000) / 1000.0f);
  std::vector<__nv_bfloat16> Ah(128 * 128), Bh(128 * 128);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, 128 * 128 * 2));
  CK(cudaMalloc(&dB, 128 * 128 * 2));
  CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));

  struct Case { int variant; uint32_t lbo_a, sbo_a, lbo_b, sbo_b; const char* name; };
  Case cases[] = {
      {1, 16, 1024, 16, 1024, "v1 QK^T K-major SW128 (lbo=16,sbo=1024)"},
      {1, 0, 1024, 0, 1024, "v1 QK^T K-major SW128 (lbo=0)"},
      {5, 16, 1024, 16, 1024, "v5 manual-swizzle A"},
      {6, 16, 1024, 16, 1024, "v6 N=64 accumulate x2"},
      {2, 16, 1024, 16384, 1024, "v2 PV smem, B MN-major (lbo=16384,sbo=1024)"},
      {2, 16, 1024, 1024, 16384, "v2 PV smem, B MN-major (lbo=1024,sbo=16384) [swapped]"},
      {2, 16, 1024, 0, 1024, "v2 PV smem, B MN-major (lbo=0,sbo=1024)"},
      {3, 0, 0, 16384, 1024, "v3 PV tmem-A, B MN-major"},
      {4, 16384, 1024, 16384, 1024, "v4 A MN-major (lbo=16384,sbo=1024)"},
      {4, 1024, 16384, 16384, 1024, "v4 A MN-major (lbo=1024,sbo=16384) [swapped]"},
  };
  int nfail = 0;
  for (const Case& c : cases) {
    int v = c.variant;
    // layouts in global memory
    int a_rows, a_cols, b_rows, b_cols, n_out;
    if (v == 1 || v == 5) { a_rows = 128; a_cols = 64; b_rows = 128; b_cols = 64; n_out = 128; }
    else if (v == 6)      { a_rows = 128; a_cols = 64; b_rows = 64;  b_cols = 64; n_out = 64; }
    else                  { a_rows = 128; a_cols = 128; b_rows = 128; b_cols = 64; n_out = 64; }
    for (int i = 0; i < a_rows * a_cols; ++i) Ah[i] = __float2bfloat16(A[i]);
    for (int i = 0; i < b_rows * b_cols; ++i) Bh[i] = __float2bfloat16(B[i]);
    CK(cudaMemcpy(dA, Ah.data(), a_rows * a_cols * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bh.data(), b_rows * b_cols * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * 128 * 4));
    CUtensorMap tmA, tmB;
    int r1 = make2d(&tmA, dA, a_cols, a_rows, 64, 128);
    int r2 = make2d(&tmB, dB, b_cols, b_rows, 64, (uint32_t)b_rows);
    if (r1 || r2) { printf("tensor map encode failed %d %d\n", r1, r2); return 3; }
    ProbeArgs pa{v, c.lbo_a, c.sbo_a, c.lbo_b, c.sbo_b, n_out};
    probe_kernel<<<1, 128, 66560>>>(tmA, tmB, dA, dD, pa);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 4; }
    printf("ran %s\n", c.name);
    std::vector<float> Dh(128 * n_out);
    CK(cudaMemcpy(Dh.data(), dD, 128 * n_out * 4, cudaMemcpyDeviceToHost));
    // reference
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < n_out; ++n) {
        double acc = 0;
        if (v == 1 || v == 5 || v == 6) {
          for (int k = 0; k < 64; ++k) acc += (double)A[m * 64 + k] * B[n * 64 + k];
          if (v == 6) acc *= 2;
        } else if (v == 2 || v == 3) {
          for (int k = 0; k < 128; ++k) acc += (double)A[m * 128 + k] * B[k * 64 + n];
        } else {  // v4: A stored as AT[k][m]
          for (int k = 0; k < 128; ++k) acc += (double)A[k * 128 + m] * B[k * 64 + n];
        }
        maxerr = fmax(maxerr, fabs(acc - Dh[m * n_out + n]));
        maxref = fmax(maxref, fabs(acc));
      }
    bool ok = maxerr < 1e-3 * fmax(1.0, maxref);
    printf("[%s] max_err=%.6f max_ref=%.3f D[0][0..3]=%.4f %.4f %.4f %.4f  %s\n", c.name, maxerr, maxref,
           Dh[0], Dh[1], Dh[2], Dh[3], ok ? "PASS" : "FAIL");
    if (!ok) ++nfail;
  }
  printf("probe done, %d failing cases (some are expected alternates)\n", nfail);
  return 0;
}
