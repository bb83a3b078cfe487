// Standalone sm_100a probe for the layouts the tcgen05 window-attention kernels rely on (head_dim 16 / 32):
//   g*: TMA tile::gather4 (4 arbitrary rows of a 2-D tensor per instruction) and a rank-5 box load with
//       SWIZZLE_32B / SWIZZLE_64B: where do the bytes land in shared memory, what do out-of-bounds rows read as
//   k*: UMMA SS, A and B K-major with 32-byte rows (SW32, K = 16) and 64-byte rows (SW64, K = 32)     -> S = Q K^T
//   m*: UMMA TS, A (bf16) from TMEM, B MN-major with 32-byte / 64-byte rows (SW32 / SW64)              -> O = P V
// Every case self-checks against a CPU reference and prints PASS / FAIL; alternates of the descriptor fields are tried.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_small_swizzle probe_small_swizzle.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

constexpr uint32_t kLayoutSW64 = 4, kLayoutSW32 = 6;

// byte offset of 16-byte chunk c of row r in a dense tile of `row_bytes`-byte rows with the matching TMA / UMMA swizzle
// (Swizzle<B,4,3>: address bits [4,4+B) ^= bits [7,7+B); B = 1 / 2 / 3 for 32 / 64 / 128-byte rows)
__host__ __device__ inline uint32_t swz(uint32_t r, uint32_t c, uint32_t row_bytes) {
  const uint32_t lin = r * row_bytes + c * 16;
  const uint32_t mask = (row_bytes / 16 - 1);
  return lin ^ (((lin >> 7) & mask) << 4);
}

// ------------------------------------------------------------------------------------------------ TMA probes
__device__ __forceinline__ void tma_gather4(void* dst, const void* tmap, uint64_t* bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

struct TmaArgs { int mode; int dst_off; int rows[4]; int col; int bytes; int c[5]; };

__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, TmaArgs a, uint8_t* out, int out_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) buf[i] = 0xEE;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, a.bytes);
    if (a.mode == 0) tma_gather4(buf + a.dst_off, &tm, &bar, a.col, a.rows[0], a.rows[1], a.rows[2], a.rows[3]);
    else tma_load_5d(buf + a.dst_off, &tm, &bar, a.c[0], a.c[1], a.c[2], a.c[3], a.c[4]);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) out[i] = buf[i];
}

// ------------------------------------------------------------------------------------------------ UMMA probes
struct MmaArgs {
  int variant;            // 0: SS K-major (A [128 x K], B [N x K]); 1: TS, A from TMEM [128 x 64], B MN-major [64 x N]
  int row_bytes;          // 32 or 64 (K-major: K*2; MN-major: N*2)
  uint32_t layout;        // 6 = SW32, 4 = SW64
  uint32_t lbo, sbo;      // bytes
  int N;                  // output columns
};

__global__ void __launch_bounds__(128, 1)
mma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, MmaArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // up to 8 KB
  uint8_t* sB = smem + 8192;     // up to 8 KB
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
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
  const int chunks = a.row_bytes / 16;
  if (a.variant == 0) {
    // A: 128 rows x K (K = row_bytes / 2), thread t owns row t; B: N rows x K, threads 0..N-1
    const int K = a.row_bytes / 2;
    for (int c = 0; c < chunks; ++c) {
      *reinterpret_cast<uint4*>(sA + swz(tid, c, a.row_bytes)) = *reinterpret_cast<const uint4*>(A + tid * K + c * 8);
      if (tid < a.N) *reinterpret_cast<uint4*>(sB + swz(tid, c, a.row_bytes)) = *reinterpret_cast<const uint4*>(B + tid * K + c * 8);
    }
  } else {
    // A: [128 x 64] bf16 into TMEM columns [64, 96) (packed pairs), thread t owns row t
    uint32_t regs[32];
    for (int j = 0; j < 32; ++j) {
      uint32_t lo = __bfloat16_as_ushort(A[tid * 64 + 2 * j]), hi = __bfloat16_as_ushort(A[tid * 64 + 2 * j + 1]);
      regs[j] = lo | (hi << 16);
    }
    tmem_st_x32(tmem + ((uint32_t)(warp * 32) << 16) + 64, regs);
    tmem_st_wait();
    // B: [64 (k) x N] row-major = MN-major operand, rows of row_bytes; threads 0..63 own a k-row
    if (tid < 64)
      for (int c = 0; c < chunks; ++c)
        *reinterpret_cast<uint4*>(sB + swz(tid, c, a.row_bytes)) = *reinterpret_cast<const uint4*>(B + tid * a.N + c * 8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (a.variant == 0) {
      const uint32_t idesc = make_idesc_bf16(128, a.N, 0, 0);
      const int ksteps = a.row_bytes / 32;     // K / 16
      for (int k = 0; k < ksteps; ++k) {
        uint64_t da = make_smem_desc(a0 + k * 32, a.lbo, a.sbo, a.layout);
        uint64_t db = make_smem_desc(b0 + k * 32, a.lbo, a.sbo, a.layout);
        umma_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
      }
    } else {
      const uint32_t idesc = make_idesc_bf16(128, a.N, 0, 1);
      for (int k = 0; k < 4; ++k) {            // 64 keys = 4 k-steps of 16 rows
        uint64_t db = make_smem_desc(b0 + k * 16 * a.row_bytes, a.lbo, a.sbo, a.layout);
        umma_ts(tmem, tmem + 64 + k * 8, db, idesc, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < a.N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[tid * a.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  srand(3);
  // ================================================================ TMA
  {
    // 2-D tensor: 40 rows x 48 two-byte elements (96-byte rows); element BIT PATTERN = row * 64 + col + 1 (TMA copies
    // bits; 0 is reserved for "zero fill")
    const int R = 40, C = 48;
    std::vector<uint16_t> h(R * C);
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 64 + c + 1);
    uint16_t* d;
    uint8_t* dout;
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMalloc(&dout, 32768));
    CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    struct G { const char* name; CUtensorMapSwizzle swz; uint32_t box_rows; int dst_off; int rows[4]; };
    G cases[] = {
        {"g1 gather4 no-swizzle box{16,1} dst+0", CU_TENSOR_MAP_SWIZZLE_NONE, 1, 0, {1, 5, 2, 9}},
        {"g3 gather4 SW32 box{16,1} dst+0", CU_TENSOR_MAP_SWIZZLE_32B, 1, 0, {1, 5, 2, 9}},
        {"g4 gather4 SW32 box{16,1} dst+128", CU_TENSOR_MAP_SWIZZLE_32B, 1, 128, {1, 5, 2, 9}},
        {"g5 gather4 SW32 box{16,1} rows incl. OOB (40, -1)", CU_TENSOR_MAP_SWIZZLE_32B, 1, 0, {3, 40, -1, 7}},
    };
    for (const G& g : cases) {
      CUtensorMap tm;
      uint64_t dims[2] = {(uint64_t)C, (uint64_t)R};
      uint64_t str[1] = {(uint64_t)C * 2};
      uint32_t box[2] = {16, g.box_rows};
      int rc = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, g.swz);
      if (rc) { printf("[%s] tensor map encode failed rc=%d\n", g.name, rc); continue; }
      TmaArgs a{};
      a.mode = 0; a.dst_off = g.dst_off; a.col = 16; a.bytes = 128;
      for (int i = 0; i < 4; ++i) a.rows[i] = g.rows[i];
      CK(cudaMemset(dout, 0, 512));
      tma_probe_kernel<<<1, 128, 4096>>>(tm, a, dout, 512);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", g.name, cudaGetErrorString(e)); return 4; }
      std::vector<uint8_t> o(512);
      CK(cudaMemcpy(o.data(), dout, 512, cudaMemcpyDeviceToHost));
      // expectation A: rows land densely at dst_off + i*32, swizzle by ABSOLUTE smem address bits
      bool okA = true, okB = true;
      const bool sw = g.swz == CU_TENSOR_MAP_SWIZZLE_32B;
      for (int i = 0; i < 4; ++i)
        for (int c = 0; c < 16; ++c) {
          const int row = g.rows[i];
          const uint16_t want = (row < 0 || row >= R) ? 0 : (uint16_t)(row * 64 + 16 + c + 1);
          const uint32_t linA = g.dst_off + i * 32 + c * 2;
          const uint32_t offA = sw ? (linA ^ (((linA >> 7) & 1) << 4)) : linA;      // absolute-address swizzle
          const uint32_t linB = i * 32 + c * 2;
          const uint32_t offB = g.dst_off + (sw ? (linB ^ (((linB >> 7) & 1) << 4)) : linB);   // box-relative swizzle
          uint16_t va, vb;
          memcpy(&va, &o[offA], 2);
          memcpy(&vb, &o[offB], 2);
          okA = okA && (va == want);
          okB = okB && (vb == want);
        }
      printf("[%s] absolute-address layout %s, box-relative layout %s; raw (row*64+col+1) per 16-byte chunk:", g.name,
             okA ? "PASS" : "FAIL", okB ? "PASS" : "FAIL");
      for (int ch = 0; ch < 16; ++ch) {
        uint16_t v;
        memcpy(&v, &o[ch * 16], 2);
        printf(" %d:%d", v ? (v - 1) / 64 : -1, v ? (v - 1) % 64 : -1);
      }
      printf("\n");
    }
    // rank-5 box: tensor (48, 9, 8, 6, 2), box (16, 7, 7, 3, 1) at coords (16, 4, 3, 4, 1): partially out of bounds
    {
      const int X = 9, Y = 8, Z = 6, Bn = 2, C5 = 48;
      std::vector<uint16_t> h5((size_t)Bn * Z * Y * X * C5);
      for (size_t i = 0; i < h5.size(); ++i) h5[i] = (uint16_t)(i % 30011 + 1);
      uint16_t* d5;
      CK(cudaMalloc(&d5, h5.size() * 2));
      CK(cudaMemcpy(d5, h5.data(), h5.size() * 2, cudaMemcpyHostToDevice));
      for (int use_sw = 0; use_sw < 2; ++use_sw) {
        CUtensorMap tm;
        uint64_t dims[5] = {(uint64_t)C5, (uint64_t)X, (uint64_t)Y, (uint64_t)Z, (uint64_t)Bn};
        uint64_t str[4] = {(uint64_t)C5 * 2, (uint64_t)C5 * X * 2, (uint64_t)C5 * X * Y * 2, (uint64_t)C5 * X * Y * Z * 2};
        uint32_t box[5] = {16, 7, 7, 3, 1};
        int rc = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d5, dims, str, box,
                           use_sw ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) { printf("[g6 rank-5 box sw=%d] encode failed rc=%d\n", use_sw, rc); continue; }
        TmaArgs a{};
        a.mode = 1; a.dst_off = 0; a.bytes = 7 * 7 * 3 * 32;
        a.c[0] = 16; a.c[1] = 4; a.c[2] = 3; a.c[3] = 4; a.c[4] = 1;
        tma_probe_kernel<<<1, 128, 8192>>>(tm, a, dout, 7 * 7 * 3 * 32);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("[g6] CUDA error: %s\n", cudaGetErrorString(e)); return 4; }
        std::vector<uint8_t> o(7 * 7 * 3 * 32);
        CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
        bool ok = true;
        for (int z = 0; z < 3; ++z)
          for (int y = 0; y < 7; ++y)
            for (int x = 0; x < 7; ++x)
              for (int c = 0; c < 16; ++c) {
                const int gx = 4 + x, gy = 3 + y, gz = 4 + z;
                uint16_t want = 0;
                if (gx < X && gy < Y && gz < Z) {
                  size_t idx = ((((size_t)1 * Z + gz) * Y + gy) * X + gx) * C5 + 16 + c;
                  want = h5[idx];
                }
                const uint32_t row = (z * 7 + y) * 7 + x;
                const uint32_t lin = row * 32 + c * 2;
                const uint32_t off = use_sw ? (lin ^ (((lin >> 7) & 1) << 4)) : lin;
                uint16_t v;
                memcpy(&v, &o[off], 2);
                ok = ok && (v == want);
              }
        printf("[g6 rank-5 box (16,7,7,3,1) partially OOB, %s] dense rows, zero fill: %s\n", use_sw ? "SW32" : "no swizzle",
               ok ? "PASS" : "FAIL");
      }
    }
  }
  // ================================================================ UMMA
  {
    std::vector<float> A(128 * 64), B(64 * 64);
    for (auto& x : A) x = bf16r((rand() % 2001 - 1000) / 1000.0f);
    for (auto& x : B) x = bf16r((rand() % 2001 - 1000) / 1000.0f);
    std::vector<__nv_bfloat16> Ah(A.size()), Bh(B.size());
    for (size_t i = 0; i < A.size(); ++i) Ah[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); ++i) Bh[i] = __float2bfloat16(B[i]);
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&dD, 128 * 64 * 4));
    CK(cudaMemcpy(dA, Ah.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bh.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    struct M { const char* name; MmaArgs a; };
    M cases[] = {
        {"k1 SS K-major SW32 K=16 N=64 (lbo=16,sbo=256)", {0, 32, kLayoutSW32, 16, 256, 64}},
        {"k2 SS K-major SW32 K=16 N=64 (lbo=0,sbo=256)", {0, 32, kLayoutSW32, 0, 256, 64}},
        {"k3 SS K-major SW64 K=32 N=64 (lbo=16,sbo=512)", {0, 64, kLayoutSW64, 16, 512, 64}},
        {"k4 SS K-major SW64 K=32 N=64 (lbo=0,sbo=512)", {0, 64, kLayoutSW64, 0, 512, 64}},
        {"m1 TS B MN-major SW32 N=16 (lbo=16,sbo=256)", {1, 32, kLayoutSW32, 16, 256, 16}},
        {"m2 TS B MN-major SW32 N=16 (lbo=256,sbo=16)", {1, 32, kLayoutSW32, 256, 16, 16}},
        {"m3 TS B MN-major SW32 N=16 (lbo=0,sbo=256)", {1, 32, kLayoutSW32, 0, 256, 16}},
        {"m4 TS B MN-major SW64 N=32 (lbo=16,sbo=512)", {1, 64, kLayoutSW64, 16, 512, 32}},
        {"m5 TS B MN-major SW64 N=32 (lbo=512,sbo=16)", {1, 64, kLayoutSW64, 512, 16, 32}},
        {"m6 TS B MN-major SW64 N=32 (lbo=0,sbo=512)", {1, 64, kLayoutSW64, 0, 512, 32}},
    };
    for (const M& m : cases) {
      CK(cudaMemset(dD, 0, 128 * 64 * 4));
      mma_probe_kernel<<<1, 128, 32768>>>(dA, dB, dD, m.a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", m.name, cudaGetErrorString(e)); return 4; }
      std::vector<float> Dh(128 * m.a.N);
      CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxref = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < m.a.N; ++n) {
          double acc = 0;
          if (m.a.variant == 0) {
            const int K = m.a.row_bytes / 2;
            for (int k = 0; k < K; ++k) acc += (double)A[i * K + k] * B[n * K + k];
          } else {
            for (int k = 0; k < 64; ++k) acc += (double)A[i * 64 + k] * B[k * m.a.N + n];
          }
          maxerr = fmax(maxerr, fabs(acc - Dh[i * m.a.N + n]));
          maxref = fmax(maxref, fabs(acc));
        }
      const bool ok = maxerr < 1e-3 * fmax(1.0, maxref);
      printf("[%s] max_err=%.6f max_ref=%.3f  %s\n", m.name, maxerr, maxref, ok ? "PASS" : "FAIL");
    }
  }
  printf("probe done\n");
  return 0;
}
