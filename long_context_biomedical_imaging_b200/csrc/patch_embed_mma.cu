// Patch embedding for LARGE reduction lengths (K = Cin * prod(patch) >= 64: cfg1 K = 256, cfg3 K = 512) on the tensor
// cores at fp32 accuracy.
//
// Same contract as patch_embed.cu (the reference's MONAI PatchEmbeddingBlock / PatchEmbed call sites,
// /root/reference/model/models/backbone_vit.py:351-361,383 and backbone_swin.py:800-806,885). The fp32 CUDA-core
// kernel there reaches 8.7 TFLOP/s at cfg3 (2.5 ms for 21.7 GFLOP), more than the attention forward of eleven layers.
// Here both fp32 operands are split into two bf16 terms, x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and
//   x * w  ~=  hi_x hi_w + hi_x lo_w + lo_x hi_w          (dropped: lo_x lo_w ~ 2^-18 |x w|, split error ~ 2^-17)
// is accumulated in fp32 by three mma.sync.m16n8k16 per tile pair: relative error ~ 2e-5, inside the 1e-4 parity
// budget of the fp32 path, so no precision switch is needed. Taken when K % 32 == 0, K >= 64 and the image is an exact
// multiple of the patch (no trailing zero pad); everything else stays on patch_embed.cu.
//
// Forward: CTA = 128 patches x 128 output features, 8 warps (4 along M x 2 along N, 32 x 64 each), K in chunks of 32:
// the im2col gather (row base + per-k offset tables), the hi/lo split and the staging into padded smem rows are done
// by all threads for chunk c+1 in registers while chunk c is multiplied (register double buffering).
// Backward (dW, dbias): CTA = 128 features x 128 k x one slab of patches; dW[n][k] += sum_m dOut[m][n] * A[m][k] with
// both operands read through transposing ldmatrix (the reduction index m is the row index of both smem tiles),
// fp32 atomics over the slabs; dbias from the same dOut tiles.
#include <cuda_bf16.h>

#include "lcbi_kernels.h"
#include "window_common.cuh"

namespace lcbi {

namespace {

constexpr int BM = 128, BN = 128, BK = 32, kThreads = 256;
constexpr int kRowBytes = BK * 2 + 16;        // 80-byte smem rows: ldmatrix conflict-free

struct PEGeomMma {
  int B, Cin, D, H, W;
  int Pd, Ph, Pw;
  int Gd, Gh, Gw;
  int N, K;
  int64_t M;
};

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float ld_px(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_px(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// image offset of the first element of patch m (no padding on this path)
__device__ __forceinline__ int64_t patch_origin(const PEGeomMma& g, int64_t m) {
  const int gw = static_cast<int>(m % g.Gw);
  const int gh = static_cast<int>((m / g.Gw) % g.Gh);
  const int gd = static_cast<int>((m / (static_cast<int64_t>(g.Gw) * g.Gh)) % g.Gd);
  const int64_t b = m / (static_cast<int64_t>(g.Gw) * g.Gh * g.Gd);
  return ((b * g.Cin * g.D + static_cast<int64_t>(gd) * g.Pd) * g.H + static_cast<int64_t>(gh) * g.Ph) * g.W +
         static_cast<int64_t>(gw) * g.Pw;
}
// offset of reduction index k inside its patch, relative to the patch origin
__device__ __forceinline__ int64_t k_offset(const PEGeomMma& g, int k) {
  const int kw = k % g.Pw;
  const int kh = (k / g.Pw) % g.Ph;
  const int kd = (k / (g.Pw * g.Ph)) % g.Pd;
  const int c = k / (g.Pw * g.Ph * g.Pd);
  return ((static_cast<int64_t>(c) * g.D + kd) * g.H + kh) * g.W + kw;
}

struct FwdSmemMma {
  uint8_t a_hi[BM * kRowBytes], a_lo[BM * kRowBytes];   // [patch][k] bf16
  uint8_t w_hi[BN * kRowBytes], w_lo[BN * kRowBytes];   // [feature][k] bf16
  int64_t row_base[BM];
  int64_t k_off[BK];
};

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kThreads)
patch_embed_fwd_mma_kernel(const TIn* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                           const float* __restrict__ pos, TOut* __restrict__ out, PEGeomMma g) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FwdSmemMma& sm = *reinterpret_cast<FwdSmemMma*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;     // this warp's 32 x 64 sub-tile

  if (tid < BM) sm.row_base[tid] = (m0 + tid < g.M) ? patch_origin(g, m0 + tid) : -1;

  // staging assignment: thread -> (row = tid / 2, 16 consecutive k at (tid & 1) * 16) for both operand tiles
  const int srow = tid >> 1, sk = (tid & 1) * 16;
  float ra[16], rw[16];
  auto fetch = [&](int k0) {       // global -> registers (chunk k0); k_off of the chunk must be in smem
    const int64_t base = sm.row_base[srow];
#pragma unroll
    for (int i = 0; i < 16; ++i) ra[i] = base >= 0 ? ld_px(img + base + sm.k_off[sk + i]) : 0.f;
    const int n = n0 + srow;
    if (n < g.N) {
      const float4* src = reinterpret_cast<const float4*>(w + static_cast<int64_t>(n) * g.K + k0 + sk);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(src + i);
        rw[i * 4] = v.x; rw[i * 4 + 1] = v.y; rw[i * 4 + 2] = v.z; rw[i * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) rw[i] = 0.f;
    }
  };
  auto stage = [&]() {             // registers -> hi / lo bf16 tiles
    uint32_t ah[8], al[8], wh[8], wl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(ra[2 * i], h0, l0);
      split_bf16(ra[2 * i + 1], h1, l1);
      ah[i] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      al[i] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
      split_bf16(rw[2 * i], h0, l0);
      split_bf16(rw[2 * i + 1], h1, l1);
      wh[i] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      wl[i] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    const int off = srow * kRowBytes + sk * 2;
    *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
    *reinterpret_cast<uint4*>(sm.a_hi + off + 16) = make_uint4(ah[4], ah[5], ah[6], ah[7]);
    *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(al[0], al[1], al[2], al[3]);
    *reinterpret_cast<uint4*>(sm.a_lo + off + 16) = make_uint4(al[4], al[5], al[6], al[7]);
    *reinterpret_cast<uint4*>(sm.w_hi + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    *reinterpret_cast<uint4*>(sm.w_hi + off + 16) = make_uint4(wh[4], wh[5], wh[6], wh[7]);
    *reinterpret_cast<uint4*>(sm.w_lo + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    *reinterpret_cast<uint4*>(sm.w_lo + off + 16) = make_uint4(wl[4], wl[5], wl[6], wl[7]);
  };

  float acc[2][8][4] = {};
  const uint32_t a_hi = smem_u32_generic(sm.a_hi), a_lo = smem_u32_generic(sm.a_lo);
  const uint32_t w_hi = smem_u32_generic(sm.w_hi), w_lo = smem_u32_generic(sm.w_lo);

  if (tid < BK) sm.k_off[tid] = k_offset(g, tid);
  __syncthreads();
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    __syncthreads();                                   // the previous chunk's tiles and k_off are no longer read
    stage();
    if (k0 + BK < g.K && tid < BK) sm.k_off[tid] = k_offset(g, k0 + BK + tid);
    __syncthreads();
    if (k0 + BK < g.K) fetch(k0 + BK);                 // in flight while this chunk is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      uint32_t fah[2][4], fal[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        load_a_frag<kRowBytes>(fah[mt], a_hi, wm + mt * 16, kk, lane);
        load_a_frag<kRowBytes>(fal[mt], a_lo, wm + mt * 16, kk, lane);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bh0, bh1, bl0, bl1;
        load_b_frag_nt<kRowBytes>(bh0, bh1, w_hi, wn + nt * 8, kk, lane);
        load_b_frag_nt<kRowBytes>(bl0, bl1, w_lo, wn + nt * 8, kk, lane);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16_16816(acc[mt][nt], fal[mt], bh0, bh1);   // small terms first
          mma_bf16_16816(acc[mt][nt], fah[mt], bl0, bl1);
          mma_bf16_16816(acc[mt][nt], fah[mt], bh0, bh1);
        }
      }
    }
  }

  // epilogue: + bias (+ position embedding), fragment rows lane / 4 (+ 8), columns 2 * (lane % 4) (+ 1)
  const int64_t np_total = static_cast<int64_t>(g.Gd) * g.Gh * g.Gw;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int64_t m = m0 + wm + mt * 16 + (lane >> 2) + half * 8;
      if (m >= g.M) continue;
      const float* pr = pos != nullptr ? pos + (m % np_total) * g.N : nullptr;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int n = n0 + wn + nt * 8 + (lane & 3) * 2;
        if (n >= g.N) continue;          // N is even on this path
        float v0 = acc[mt][nt][half * 2] + bias[n], v1 = acc[mt][nt][half * 2 + 1] + bias[n + 1];
        if (pr != nullptr) {
          v0 += pr[n];
          v1 += pr[n + 1];
        }
        if constexpr (sizeof(TOut) == 4) {
          *reinterpret_cast<float2*>(out + m * g.N + n) = make_float2(v0, v1);
        } else {
          *reinterpret_cast<__nv_bfloat162*>(out + m * g.N + n) = __floats2bfloat162_rn(v0, v1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: dW[n][k] += sum_m dOut[m][n] * A[m][k], dbias[n] += sum_m dOut[m][n]
// ------------------------------------------------------------------------------------------------
constexpr int kBwdChunk = 32;                 // patches staged per step
constexpr int kBwdSlab = 512;                 // patches per CTA
constexpr int kBwdRowBytes = 128 * 2 + 16;    // [m][128 columns] bf16 rows, padded

struct BwdSmemMma {
  uint8_t g_hi[kBwdChunk * kBwdRowBytes], g_lo[kBwdChunk * kBwdRowBytes];   // dOut tile  [m][n]
  uint8_t a_hi[kBwdChunk * kBwdRowBytes], a_lo[kBwdChunk * kBwdRowBytes];   // patch tile [m][k]
  int64_t row_base[kBwdChunk];
  int64_t k_off[128];
};

template <int STRIDE_BYTES>
__device__ __forceinline__ void load_a_frag_rows_are_k(uint32_t (&a)[4], uint32_t tile_base, int k0, int m0, int lane) {
  // A[m][k] = X[k0 + k][m0 + m] for X stored row-major with the reduction index as the row index
  const uint32_t addr = tile_base + (k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * STRIDE_BYTES + (m0 + ((lane >> 3) & 1) * 8) * 2;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

__device__ __forceinline__ float ld_g(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_g(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TIn, typename TG>
__global__ void __launch_bounds__(kThreads)
patch_embed_bwd_w_mma_kernel(const TIn* __restrict__ img, const TG* __restrict__ dout, float* __restrict__ dw,
                             float* __restrict__ dbias, PEGeomMma g) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BwdSmemMma& sm = *reinterpret_cast<BwdSmemMma*>(smem_raw);
  constexpr bool kGradIsBf16 = sizeof(TG) == 2;      // then dOut has no low part: two MMAs per tile pair instead of three
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m_begin = static_cast<int64_t>(blockIdx.x) * kBwdSlab;
  const int64_t m_end = (m_begin + kBwdSlab < g.M) ? m_begin + kBwdSlab : g.M;
  const int n0 = blockIdx.y * 128, kt0 = blockIdx.z * 128;
  const int wn = (warp & 3) * 32, wk = (warp >> 2) * 64;       // this warp's 32 (n) x 64 (k) sub-tile of dW
  const bool do_bias = dbias != nullptr && blockIdx.z == 0;

  if (tid < 128) sm.k_off[tid] = (kt0 + tid < g.K) ? k_offset(g, kt0 + tid) : -1;

  // staging assignment: thread -> (row = tid / 8, 16 consecutive columns at (tid & 7) * 16) of both tiles
  const int srow = tid >> 3, sc = (tid & 7) * 16;
  float rg[16], ra[16];
  auto fetch = [&](int64_t mc) {     // global -> registers for the chunk starting at patch mc (row_base in smem)
    const int64_t m = mc + srow;
    const int64_t base = sm.row_base[srow];
    if (m < m_end) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = n0 + sc + i;
        rg[i] = n < g.N ? ld_g(dout + m * g.N + n) : 0.f;
        const int64_t ko = sm.k_off[sc + i];
        ra[i] = ko >= 0 ? ld_px(img + base + ko) : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) rg[i] = ra[i] = 0.f;
    }
  };
  auto pack_pair = [](float x0, float x1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(x0, h0, l0);
    split_bf16(x1, h1, l1);
    hi = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    lo = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  };
  auto stage = [&]() {
    uint32_t gh[8], gl[8], ah[8], al[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pack_pair(rg[2 * i], rg[2 * i + 1], gh[i], gl[i]);
      pack_pair(ra[2 * i], ra[2 * i + 1], ah[i], al[i]);
    }
    const int off = srow * kBwdRowBytes + sc * 2;
    *reinterpret_cast<uint4*>(sm.g_hi + off) = make_uint4(gh[0], gh[1], gh[2], gh[3]);
    *reinterpret_cast<uint4*>(sm.g_hi + off + 16) = make_uint4(gh[4], gh[5], gh[6], gh[7]);
    if (!kGradIsBf16) {
      *reinterpret_cast<uint4*>(sm.g_lo + off) = make_uint4(gl[0], gl[1], gl[2], gl[3]);
      *reinterpret_cast<uint4*>(sm.g_lo + off + 16) = make_uint4(gl[4], gl[5], gl[6], gl[7]);
    }
    *reinterpret_cast<uint4*>(sm.a_hi + off) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
    *reinterpret_cast<uint4*>(sm.a_hi + off + 16) = make_uint4(ah[4], ah[5], ah[6], ah[7]);
    *reinterpret_cast<uint4*>(sm.a_lo + off) = make_uint4(al[0], al[1], al[2], al[3]);
    *reinterpret_cast<uint4*>(sm.a_lo + off + 16) = make_uint4(al[4], al[5], al[6], al[7]);
  };

  float acc[2][8][4] = {};
  float acc_bias[2][4] = {};                           // dOut^T times a block of ones: every column is dbias
  const bool bias_warp = do_bias && wk == 0;
  const uint32_t g_hi = smem_u32_generic(sm.g_hi), g_lo = smem_u32_generic(sm.g_lo);
  const uint32_t a_hi = smem_u32_generic(sm.a_hi), a_lo = smem_u32_generic(sm.a_lo);

  if (tid < kBwdChunk) sm.row_base[tid] = (m_begin + tid < m_end) ? patch_origin(g, m_begin + tid) : 0;
  __syncthreads();
  fetch(m_begin);
  for (int64_t mc = m_begin; mc < m_end; mc += kBwdChunk) {
    __syncthreads();                                   // the previous chunk's tiles and row_base are no longer read
    stage();
    const int64_t mn = mc + kBwdChunk;
    if (mn < m_end && tid < kBwdChunk) sm.row_base[tid] = (mn + tid < m_end) ? patch_origin(g, mn + tid) : 0;
    __syncthreads();
    if (mn < m_end) fetch(mn);                         // in flight while this chunk is multiplied
#pragma unroll
    for (int mm = 0; mm < kBwdChunk; mm += 16) {       // reduction over patches
      uint32_t fgh[2][4], fgl[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        load_a_frag_rows_are_k<kBwdRowBytes>(fgh[nt], g_hi, mm, wn + nt * 16, lane);
        if (!kGradIsBf16) load_a_frag_rows_are_k<kBwdRowBytes>(fgl[nt], g_lo, mm, wn + nt * 16, lane);
        if (bias_warp) {
          constexpr uint32_t kOnes = 0x3F803F80u;      // two bf16 1.0
          if (!kGradIsBf16) mma_bf16_16816(acc_bias[nt], fgl[nt], kOnes, kOnes);
          mma_bf16_16816(acc_bias[nt], fgh[nt], kOnes, kOnes);
        }
      }
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        uint32_t bh0, bh1, bl0, bl1;
        load_b_frag_t<kBwdRowBytes>(bh0, bh1, a_hi, mm, wk + kt * 8, lane);
        load_b_frag_t<kBwdRowBytes>(bl0, bl1, a_lo, mm, wk + kt * 8, lane);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          if (!kGradIsBf16) mma_bf16_16816(acc[nt][kt], fgl[nt], bh0, bh1);
          mma_bf16_16816(acc[nt][kt], fgh[nt], bl0, bl1);
          mma_bf16_16816(acc[nt][kt], fgh[nt], bh0, bh1);
        }
      }
    }
  }

  // dW: fragment rows (n) lane / 4 (+ 8), columns (k) 2 * (lane % 4) (+ 1)
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + wn + nt * 16 + (lane >> 2) + half * 8;
      if (n >= g.N) continue;
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        const int k = kt0 + wk + kt * 8 + (lane & 3) * 2;
        if (k < g.K) atomicAdd(dw + static_cast<int64_t>(n) * g.K + k, acc[nt][kt][half * 2]);
        if (k + 1 < g.K) atomicAdd(dw + static_cast<int64_t>(n) * g.K + k + 1, acc[nt][kt][half * 2 + 1]);
      }
    }
  }
  if (bias_warp && (lane & 3) == 0) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int n = n0 + wn + nt * 16 + (lane >> 2);
      if (n < g.N) atomicAdd(dbias + n, acc_bias[nt][0]);
      if (n + 8 < g.N) atomicAdd(dbias + n + 8, acc_bias[nt][2]);
    }
  }
}

}  // namespace

bool patch_embed_mma_applicable(int Cin, const int* img_dims, const int* patch, const int* grid, int N) {
  const int K = Cin * patch[0] * patch[1] * patch[2];
  if (K < 64 || K % BK != 0 || N % 2 != 0) return false;
  for (int i = 0; i < 3; ++i)
    if (grid[i] * patch[i] > img_dims[i]) return false;      // trailing zero pad: stays on the general kernel
  return true;
}

int patch_embed_fwd_mma_launch(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                               void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                               const int* grid, int N, cudaStream_t stream) {
  PEGeomMma g;
  g.B = B; g.Cin = Cin; g.D = img_dims[0]; g.H = img_dims[1]; g.W = img_dims[2];
  g.Pd = patch[0]; g.Ph = patch[1]; g.Pw = patch[2];
  g.Gd = grid[0]; g.Gh = grid[1]; g.Gw = grid[2];
  g.N = N; g.K = Cin * patch[0] * patch[1] * patch[2];
  g.M = static_cast<int64_t>(B) * grid[0] * grid[1] * grid[2];
  if ((reinterpret_cast<uintptr_t>(w) & 15) != 0) return LCBI_ERR_BAD_ARG;
  dim3 grd(static_cast<unsigned>((g.M + BM - 1) / BM), (N + BN - 1) / BN);
  const int smem = static_cast<int>(sizeof(FwdSmemMma));
#define LCBI_PE_MMA(TI, TO)                                                                                          \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(patch_embed_fwd_mma_kernel<TI, TO>,                                         \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                         \
    if (e != cudaSuccess) return set_cuda_error(e);                                                                  \
    patch_embed_fwd_mma_kernel<TI, TO><<<grd, kThreads, smem, stream>>>(static_cast<const TI*>(img), w, bias, pos,   \
                                                                        static_cast<TO*>(out), g);                   \
  } while (0)
  if (img_is_bf16) {
    if (out_is_bf16) LCBI_PE_MMA(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_MMA(__nv_bfloat16, float);
  } else {
    if (out_is_bf16) LCBI_PE_MMA(float, __nv_bfloat16); else LCBI_PE_MMA(float, float);
  }
#undef LCBI_PE_MMA
  return set_cuda_error(cudaGetLastError());
}

// dw and dbias must have been zeroed by the caller (they are accumulated with fp32 atomics over the patch slabs)
int patch_embed_bwd_w_mma_launch(const void* img, int img_is_bf16, const void* dout, int dout_is_bf16, float* dw,
                                 float* dbias, int B, int Cin, const int* img_dims, const int* patch, const int* grid,
                                 int N, cudaStream_t stream) {
  PEGeomMma g;
  g.B = B; g.Cin = Cin; g.D = img_dims[0]; g.H = img_dims[1]; g.W = img_dims[2];
  g.Pd = patch[0]; g.Ph = patch[1]; g.Pw = patch[2];
  g.Gd = grid[0]; g.Gh = grid[1]; g.Gw = grid[2];
  g.N = N; g.K = Cin * patch[0] * patch[1] * patch[2];
  g.M = static_cast<int64_t>(B) * grid[0] * grid[1] * grid[2];
  dim3 grd(static_cast<unsigned>((g.M + kBwdSlab - 1) / kBwdSlab), (N + 127) / 128, (g.K + 127) / 128);
  const int smem = static_cast<int>(sizeof(BwdSmemMma));
#define LCBI_PE_BMMA(TI, TG)                                                                                         \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(patch_embed_bwd_w_mma_kernel<TI, TG>,                                       \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                         \
    if (e != cudaSuccess) return set_cuda_error(e);                                                                  \
    patch_embed_bwd_w_mma_kernel<TI, TG><<<grd, kThreads, smem, stream>>>(static_cast<const TI*>(img),               \
                                                                          static_cast<const TG*>(dout), dw, dbias, g); \
  } while (0)
  if (img_is_bf16) {
    if (dout_is_bf16) LCBI_PE_BMMA(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_BMMA(__nv_bfloat16, float);
  } else {
    if (dout_is_bf16) LCBI_PE_BMMA(float, __nv_bfloat16); else LCBI_PE_BMMA(float, float);
  }
#undef LCBI_PE_BMMA
  return set_cuda_error(cudaGetLastError());
}

}  // namespace lcbi
