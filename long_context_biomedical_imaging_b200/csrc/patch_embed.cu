// Patch-embedding projection (strided conv with kernel == stride == patch) as an implicit GEMM.
//
// Replaces MONAI 1.3.0 PatchEmbeddingBlock / PatchEmbed as called by the reference
// (/root/reference/model/models/backbone_vit.py:351-361,383 and backbone_swin.py:800-806,885):
//   out[b, p, n] = bias[n] + sum_k img[b, c, (pd,ph,pw)*patch + (kd,kh,kw)] * W[n, k]   (+ pos[p, n] for ViT)
// with k = ((c*Pd + kd)*Ph + kh)*Pw + kw (the conv weight's own memory order), patches p in raster order over
// the conv output grid, and image positions past the far edge read as zero (PatchEmbed's trailing pad;
// PatchEmbeddingBlock floors instead, which the host expresses through the patch-grid size it passes).
// Output is token-major / channel-last (B, Np, N): no flatten/transpose copy afterwards.
//
// M = B*Np, K = Cin*prod(patch), N = hidden. For the reference's long-context configs K is 4..16 and the
// kernel is bound by the HBM write of the output, so this is an fp32 CUDA-core GEMM with coalesced 16-byte
// stores; the im2col gather is fused into the A-tile load.
#include <cuda_bf16.h>

#include "lcbi_kernels.h"
#include "window_common.cuh"      // FastDiv

namespace lcbi {

namespace {

constexpr int TM = 32, TN = 128, KC = 32, kThreads = 256;

struct PEGeom {
  int B, Cin, D, H, W;      // image (D == 1 for 2-D)
  int Pd, Ph, Pw;           // patch
  int Gd, Gh, Gw;           // patch grid
  int N, K;                 // hidden, Cin*Pd*Ph*Pw
  int64_t M;                // B*Gd*Gh*Gw
  FastDiv d_Gw, d_Gh, d_Gd, d_Pw, d_Ph, d_Pd;   // multiply-high divisors of the (patch, k) -> pixel map
};

__device__ __forceinline__ float load_px(const float* p) { return *p; }
__device__ __forceinline__ float load_px(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// image offset of (patch m, reduction index k), or -1 when it falls into the zero padding
__device__ __forceinline__ int64_t px_offset(const PEGeom& g, int64_t m, int k) {
  if (g.M < (1ll << 31)) {
    // the usual case: every quotient fits 31 bits, six multiply-high divisions instead of ~10 runtime divisions (which
    // were a large share of the instructions of the K <= 16 kernels: 4..16 pixels per patch leave little else to do)
    int q, gw, gh, gd, b, kw, kh, kd, c;
    fdivmod(static_cast<int>(m), g.d_Gw, q, gw);
    fdivmod(q, g.d_Gh, q, gh);
    fdivmod(q, g.d_Gd, b, gd);
    fdivmod(k, g.d_Pw, q, kw);
    fdivmod(q, g.d_Ph, q, kh);
    fdivmod(q, g.d_Pd, c, kd);
    const int z = gd * g.Pd + kd, y = gh * g.Ph + kh, x = gw * g.Pw + kw;
    if (z >= g.D || y >= g.H || x >= g.W) return -1;
    return (((static_cast<int64_t>(b) * g.Cin + c) * g.D + z) * g.H + y) * g.W + x;
  }
  const int gw = static_cast<int>(m % g.Gw);
  const int gh = static_cast<int>((m / g.Gw) % g.Gh);
  const int gd = static_cast<int>((m / (static_cast<int64_t>(g.Gw) * g.Gh)) % g.Gd);
  const int b = static_cast<int>(m / (static_cast<int64_t>(g.Gw) * g.Gh * g.Gd));
  const int kw = k % g.Pw;
  const int kh = (k / g.Pw) % g.Ph;
  const int kd = (k / (g.Pw * g.Ph)) % g.Pd;
  const int c = k / (g.Pw * g.Ph * g.Pd);
  const int z = gd * g.Pd + kd, y = gh * g.Ph + kh, x = gw * g.Pw + kw;
  if (z >= g.D || y >= g.H || x >= g.W) return -1;
  return (((static_cast<int64_t>(b) * g.Cin + c) * g.D + z) * g.H + y) * g.W + x;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kThreads)
patch_embed_fwd_kernel(const TIn* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ pos, TOut* __restrict__ out, PEGeom g) {
  __shared__ float As[KC][TM + 1];
  __shared__ __align__(16) float Ws[KC][TN];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * TM;
  const int n0 = blockIdx.y * TN;
  float acc[4][4] = {};

  for (int k0 = 0; k0 < g.K; k0 += KC) {
    const int kmax = min(KC, g.K - k0);       // only the real reduction extent is staged (K is 4..16 for the
                                              // long-context configs: filling all KC rows would be mostly waste)
    // A tile: element e -> (kk = e / TM, mm = e % TM); consecutive threads walk adjacent patches
    for (int e = tid; e < TM * kmax; e += kThreads) {
      const int kk = e / TM, mm = e % TM;
      float v = 0.f;
      if (m0 + mm < g.M) {
        const int64_t off = px_offset(g, m0 + mm, k0 + kk);
        if (off >= 0) v = load_px(img + off);
      }
      As[kk][mm] = v;
    }
    // W tile (transposed on the fly): W[n][k] -> Ws[k][n]; consecutive threads walk k (contiguous in W)
    for (int e = tid; e < TN * kmax; e += kThreads) {
      const int kk = e % kmax, nn = e / kmax;
      float v = 0.f;
      if (n0 + nn < g.N) v = w[static_cast<int64_t>(n0 + nn) * g.K + k0 + kk];
      Ws[kk][nn] = v;
    }
    __syncthreads();
    for (int kk = 0; kk < kmax; ++kk) {
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float a0 = As[kk][ty * 4 + 0], a1 = As[kk][ty * 4 + 1], a2 = As[kk][ty * 4 + 2], a3 = As[kk][ty * 4 + 3];
      acc[0][0] = fmaf(a0, b4.x, acc[0][0]); acc[0][1] = fmaf(a0, b4.y, acc[0][1]);
      acc[0][2] = fmaf(a0, b4.z, acc[0][2]); acc[0][3] = fmaf(a0, b4.w, acc[0][3]);
      acc[1][0] = fmaf(a1, b4.x, acc[1][0]); acc[1][1] = fmaf(a1, b4.y, acc[1][1]);
      acc[1][2] = fmaf(a1, b4.z, acc[1][2]); acc[1][3] = fmaf(a1, b4.w, acc[1][3]);
      acc[2][0] = fmaf(a2, b4.x, acc[2][0]); acc[2][1] = fmaf(a2, b4.y, acc[2][1]);
      acc[2][2] = fmaf(a2, b4.z, acc[2][2]); acc[2][3] = fmaf(a2, b4.w, acc[2][3]);
      acc[3][0] = fmaf(a3, b4.x, acc[3][0]); acc[3][1] = fmaf(a3, b4.y, acc[3][1]);
      acc[3][2] = fmaf(a3, b4.z, acc[3][2]); acc[3][3] = fmaf(a3, b4.w, acc[3][3]);
    }
    __syncthreads();
  }

  const int n = n0 + tx * 4;
  if (n >= g.N) return;
  const int64_t np_total = static_cast<int64_t>(g.Gd) * g.Gh * g.Gw;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (n + j < g.N) ? bias[n + j] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) break;
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = acc[i][j] + bv[j];
    if (pos != nullptr) {
      const float* pr = pos + (m % np_total) * g.N + n;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) r[j] += pr[j];
    }
    TOut* o = out + m * g.N + n;
    if (n + 3 < g.N && (g.N & 3) == 0) {
      if constexpr (sizeof(TOut) == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(o) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < g.N) o[j] = static_cast<TOut>(r[j]);
    }
  }
}

// dW[n][k] += sum_m dOut[m][n] * A[m][k];  dbias[n] += sum_m dOut[m][n]   (split over M, fp32 atomics)
// grid: (M slabs, N tiles of TN, K tiles of KC). Each CTA reduces MS patches.
constexpr int MS = 256;
template <typename TIn, typename TG>
__global__ void __launch_bounds__(kThreads)
patch_embed_bwd_w_kernel(const TIn* __restrict__ img, const TG* __restrict__ dout, float* __restrict__ dw,
                         float* __restrict__ dbias, PEGeom g) {
  __shared__ float As[TM][KC + 1];            // [m][k]
  __shared__ __align__(16) float Gs[TM][TN];  // [m][n]
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;   // tx -> 4 n, ty -> 4 k
  const int64_t m_begin = static_cast<int64_t>(blockIdx.x) * MS;
  const int n0 = blockIdx.y * TN, k0 = blockIdx.z * KC;
  float acc[4][4] = {};   // [k][n]
  float bsum[4] = {};
  const int64_t m_end = min(m_begin + MS, g.M);
  for (int64_t m0 = m_begin; m0 < m_end; m0 += TM) {
    for (int e = tid; e < TM * KC; e += kThreads) {
      const int kk = e / TM, mm = e % TM;
      float v = 0.f;
      if (k0 + kk < g.K && m0 + mm < m_end) {
        const int64_t off = px_offset(g, m0 + mm, k0 + kk);
        if (off >= 0) v = load_px(img + off);
      }
      As[mm][kk] = v;
    }
    for (int e = tid; e < TM * TN; e += kThreads) {
      const int nn = e % TN, mm = e / TN;
      float v = 0.f;
      if (n0 + nn < g.N && m0 + mm < m_end) v = static_cast<float>(dout[(m0 + mm) * g.N + n0 + nn]);
      Gs[mm][nn] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int mm = 0; mm < TM; ++mm) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[mm][tx * 4]);
      const float a0 = As[mm][ty * 4 + 0], a1 = As[mm][ty * 4 + 1], a2 = As[mm][ty * 4 + 2], a3 = As[mm][ty * 4 + 3];
      acc[0][0] = fmaf(a0, g4.x, acc[0][0]); acc[0][1] = fmaf(a0, g4.y, acc[0][1]);
      acc[0][2] = fmaf(a0, g4.z, acc[0][2]); acc[0][3] = fmaf(a0, g4.w, acc[0][3]);
      acc[1][0] = fmaf(a1, g4.x, acc[1][0]); acc[1][1] = fmaf(a1, g4.y, acc[1][1]);
      acc[1][2] = fmaf(a1, g4.z, acc[1][2]); acc[1][3] = fmaf(a1, g4.w, acc[1][3]);
      acc[2][0] = fmaf(a2, g4.x, acc[2][0]); acc[2][1] = fmaf(a2, g4.y, acc[2][1]);
      acc[2][2] = fmaf(a2, g4.z, acc[2][2]); acc[2][3] = fmaf(a2, g4.w, acc[2][3]);
      acc[3][0] = fmaf(a3, g4.x, acc[3][0]); acc[3][1] = fmaf(a3, g4.y, acc[3][1]);
      acc[3][2] = fmaf(a3, g4.z, acc[3][2]); acc[3][3] = fmaf(a3, g4.w, acc[3][3]);
      if (ty == 0) { bsum[0] += g4.x; bsum[1] += g4.y; bsum[2] += g4.z; bsum[3] += g4.w; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= g.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.N) atomicAdd(&dw[static_cast<int64_t>(n) * g.K + k], acc[i][j]);
    }
  }
  if (ty == 0 && blockIdx.z == 0 && dbias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.N) atomicAdd(&dbias[n], bsum[j]);
    }
  }
}

// dpos[p][n] = sum_b dOut[b][p][n]
template <typename TG>
__global__ void patch_embed_bwd_pos_kernel(const TG* __restrict__ dout, float* __restrict__ dpos, int B, int64_t np,
                                           int N) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= np * N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += static_cast<float>(dout[b * np * N + idx]);
  dpos[idx] = s;
}

// dimg[pixel] = sum_n dOut[m(pixel)][n] * W[n][k(pixel)]  (every pixel belongs to at most one patch)
template <typename TG>
__global__ void patch_embed_bwd_x_kernel(const TG* __restrict__ dout, const float* __restrict__ w,
                                         float* __restrict__ dimg, PEGeom g) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(g.B) * g.Cin * g.D * g.H * g.W;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % g.W);
  const int y = static_cast<int>((idx / g.W) % g.H);
  const int z = static_cast<int>((idx / (static_cast<int64_t>(g.W) * g.H)) % g.D);
  const int c = static_cast<int>((idx / (static_cast<int64_t>(g.W) * g.H * g.D)) % g.Cin);
  const int b = static_cast<int>(idx / (static_cast<int64_t>(g.W) * g.H * g.D * g.Cin));
  const int gd = z / g.Pd, gh = y / g.Ph, gw = x / g.Pw;
  float s = 0.f;
  if (gd < g.Gd && gh < g.Gh && gw < g.Gw) {
    const int k = ((c * g.Pd + z % g.Pd) * g.Ph + y % g.Ph) * g.Pw + x % g.Pw;
    const int64_t m = ((static_cast<int64_t>(b) * g.Gd + gd) * g.Gh + gh) * g.Gw + gw;
    const TG* row = dout + m * g.N;
    for (int n = 0; n < g.N; ++n) s = fmaf(static_cast<float>(row[n]), w[static_cast<int64_t>(n) * g.K + k], s);
  }
  dimg[idx] = s;
}

// ------------------------------------------------------------------------------------------------
// Small-K backward (K <= 16): one pass over dOut produces dW, dbias and (when the batch is 1, so that the sum over
// the batch is the identity) dpos. CTA = kBwRows patches x 128 channels; thread = 4 channels; warp w takes rows
// w, w+8, ... of the slab. The slab's pixels are staged once in shared memory ([row][k], K floats per row), every
// dOut row is read with one coalesced 16-byte load per thread, partial dW[K][4] / dbias[4] live in registers and
// are reduced across the 8 warps through shared memory, then added to global with fp32 atomics.
// ------------------------------------------------------------------------------------------------
constexpr int kBwRows = 256;
constexpr int kBwKMax = 16;

template <typename TIn, typename TG>
__global__ void __launch_bounds__(256)
patch_embed_bwd_smallk_kernel(const TIn* __restrict__ img, const TG* __restrict__ dout, float* __restrict__ dw,
                              float* __restrict__ dbias, float* __restrict__ dpos_copy, PEGeom g) {
  extern __shared__ __align__(16) float pe_smem[];
  float (*px)[kBwKMax + 1] = reinterpret_cast<float (*)[kBwKMax + 1]>(pe_smem);           // [kBwRows][17]
  float* red = pe_smem + kBwRows * (kBwKMax + 1);                                          // [8][K+1][132]
  const int red_k = 132, red_w = (g.K + 1) * 132;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y * 128 + lane * 4;
  const bool live = n < g.N;                 // N is a multiple of 4 on this path (checked by the launcher)
  float acc[kBwKMax][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < kBwKMax; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  const int64_t n_slabs = (g.M + kBwRows - 1) / kBwRows;
  // persistent over the row slabs: the partial sums stay in registers, so each CTA issues its atomics once
  for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
    const int64_t m_begin = slab * kBwRows;
    const int rows = static_cast<int>(min(static_cast<int64_t>(kBwRows), g.M - m_begin));
    __syncthreads();                         // previous slab's pixels fully consumed
    for (int e = tid; e < rows * g.K; e += 256) {
      const int r = e / g.K, k = e % g.K;
      const int64_t off = px_offset(g, m_begin + r, k);
      px[r][k] = off >= 0 ? load_px(img + off) : 0.f;
    }
    __syncthreads();
    if (live) {
      auto load_row = [&](int r) {
        const TG* src = dout + (m_begin + r) * g.N + n;
        if constexpr (sizeof(TG) == 4) {
          return *reinterpret_cast<const float4*>(src);
        } else {
          const uint2 raw = *reinterpret_cast<const uint2*>(src);
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
          return make_float4(lo.x, lo.y, hi.x, hi.y);
        }
      };
      auto use_row = [&](int r, const float4& gv) {
        if (dpos_copy != nullptr) *reinterpret_cast<float4*>(dpos_copy + (m_begin + r) * g.N + n) = gv;
        bsum[0] += gv.x; bsum[1] += gv.y; bsum[2] += gv.z; bsum[3] += gv.w;
#pragma unroll
        for (int k = 0; k < kBwKMax; ++k) {
          if (k < g.K) {
            const float a = px[r][k];
            acc[k][0] = fmaf(a, gv.x, acc[k][0]); acc[k][1] = fmaf(a, gv.y, acc[k][1]);
            acc[k][2] = fmaf(a, gv.z, acc[k][2]); acc[k][3] = fmaf(a, gv.w, acc[k][3]);
          }
        }
      };
      int r = warp;
      for (; r + 24 < rows; r += 32) {       // four rows in flight per thread
        const float4 g0 = load_row(r), g1 = load_row(r + 8), g2 = load_row(r + 16), g3 = load_row(r + 24);
        use_row(r, g0); use_row(r + 8, g1); use_row(r + 16, g2); use_row(r + 24, g3);
      }
      for (; r < rows; r += 8) use_row(r, load_row(r));
    }
  }
#pragma unroll
  for (int k = 0; k < kBwKMax; ++k)
    if (k < g.K)
      *reinterpret_cast<float4*>(red + warp * red_w + k * red_k + lane * 4) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
  *reinterpret_cast<float4*>(red + warp * red_w + g.K * red_k + lane * 4) = make_float4(bsum[0], bsum[1], bsum[2], bsum[3]);
  __syncthreads();
  for (int e = tid; e < (g.K + 1) * 128; e += 256) {
    const int k = e / 128, c = e % 128;
    const int col = blockIdx.y * 128 + c;
    if (col >= g.N) continue;
    float sum = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) sum += red[w8 * red_w + k * red_k + c];
    if (k == g.K) {
      if (dbias != nullptr) atomicAdd(&dbias[col], sum);
    } else {
      atomicAdd(&dw[static_cast<int64_t>(col) * g.K + k], sum);
    }
  }
}

// dpos[p][n] = sum_b dOut[b][p][n], four channels per thread
template <typename TG>
__global__ void patch_embed_bwd_pos4_kernel(const TG* __restrict__ dout, float* __restrict__ dpos, int B, int64_t np_n4) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= np_n4) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const TG* src = dout + (b * np_n4 + idx) * 4;
    if constexpr (sizeof(TG) == 4) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    } else {
      const uint2 raw = *reinterpret_cast<const uint2*>(src);
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      s.x += lo.x; s.y += lo.y; s.z += hi.x; s.w += hi.y;
    }
  }
  *reinterpret_cast<float4*>(dpos + idx * 4) = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward for the long-context configs, K = Cin * prod(patch) in {4, 8, 16} (cfg5 / cfg4 / cfg2): pure streaming.
// 4..16 FMAs per output value against 2..4 bytes written, so the only thing that matters is that the output (and the
// position embedding) move as whole cache lines and that nothing else is on the critical path:
//   * a CTA takes 64 consecutive patches of one patch row: their pixels are Pd*Ph contiguous runs of the image (coalesced
//     loads, no per-element index arithmetic beyond one division), staged patch-major in shared memory;
//   * a thread owns 4 consecutive features: its 4 x K weights and 4 biases live in registers for the whole CTA;
//   * threads are laid out (patch, feature group) in the memory order of `out`, so a warp's 16-byte (fp32) or 8-byte
//     (bf16) stores and its position-embedding loads form one contiguous stream.
// The generic tiled kernel above re-derived (b, z, y, x) with ~10 divisions per loaded pixel and wrote through a
// 32 x 128 tile: 7..9 % of the HBM roofline at cfg2 / cfg4, 47 % at cfg5.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStreamRows = 64;

// two fp32 FMAs per instruction (sm_100 FFMA2), IEEE rounding, no flush-to-zero
__device__ __forceinline__ float2 pe_fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

template <int KK, typename TIn, typename TOut>
__global__ void __launch_bounds__(256, KK == 16 ? 2 : (KK == 8 ? 3 : 5))
patch_embed_fwd_stream_kernel(const TIn* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                              const float* __restrict__ pos, TOut* __restrict__ out, PEGeom g, int chunks_per_row, int n_items) {
  constexpr int KS = KK + 4;                                  // padded row: conflict-free stores, 16-byte aligned reads
  constexpr int kPre = kStreamRows * KK / 256;                // pixels per thread and item
  __shared__ __align__(16) float s_a[kStreamRows * KS];
  const int tid = threadIdx.x;

  // ---- this thread's 4 features: weights (as (k, k+1) pairs for FFMA2) and bias stay in registers for all items
  const int tpp = g.N >> 2;                                   // threads per patch
  const int ppp = 256 / tpp;                                  // patches per pass
  const int fg = tid % tpp, rs = tid / tpp;
  float2 wr[4][KK / 2];
  float bv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bv[i] = bias[fg * 4 + i];
#pragma unroll
    for (int k4 = 0; k4 < KK / 4; ++k4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(w + static_cast<int64_t>(fg * 4 + i) * KK + k4 * 4));
      wr[i][k4 * 2] = make_float2(t.x, t.y);
      wr[i][k4 * 2 + 1] = make_float2(t.z, t.w);
    }
  }

  // Pixels of an item: KK / Pw image rows, 64 * Pw contiguous pixels each (a ragged last chunk loads the same extent:
  // what lies past the image reads as zero, what lies past the patch grid is never consumed). The split of a thread's
  // element index into (image row of the patch, pixel) does not depend on the item and is taken once.
  const int run = kStreamRows * g.Pw;
  int e_row[kPre], e_j[kPre], e_smem[kPre];                   // e_row = channel << 16 | kz << 8 | ky
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int e = tid + u * 256;
    const int kr = e / run, j = e - kr * run;
    const int ky = kr % g.Ph, kz = (kr / g.Ph) % g.Pd, cin = kr / (g.Ph * g.Pd);
    const int r = j / g.Pw, px = j - r * g.Pw;
    e_row[u] = (cin << 16) | (kz << 8) | ky;
    e_j[u] = j;
    e_smem[u] = r * KS + kr * g.Pw + px;
  }
  float pre[kPre];
  auto fetch = [&](int item) {
    const int prow = item / chunks_per_row, x0 = (item - prow * chunks_per_row) * kStreamRows;
    const int gy = prow % g.Gh, t = prow / g.Gh;
    const int gz = t % g.Gd, b = t / g.Gd;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int z = gz * g.Pd + ((e_row[u] >> 8) & 0xff), y = gy * g.Ph + (e_row[u] & 0xff), x = x0 * g.Pw + e_j[u];
      float v = 0.f;
      if (z < g.D && y < g.H && x < g.W)
        v = load_px(img + (((static_cast<int64_t>(b) * g.Cin + (e_row[u] >> 16)) * g.D + z) * g.H + y) * g.W + x);
      pre[u] = v;
    }
  };

  int item = blockIdx.x;
  if (item < n_items) fetch(item);
  for (; item < n_items; item += gridDim.x) {
    const int prow = item / chunks_per_row, x0 = (item - prow * chunks_per_row) * kStreamRows;
    const int rv = min(kStreamRows, g.Gw - x0);               // patches in the chunk
#pragma unroll
    for (int u = 0; u < kPre; ++u) s_a[e_smem[u]] = pre[u];
    __syncthreads();
    if (item + static_cast<int>(gridDim.x) < n_items) fetch(item + gridDim.x);     // in flight behind this item's stores

    if (rs < ppp) {
      const int64_t m0 = static_cast<int64_t>(prow) * g.Gw + x0;                       // first patch (output row) of the chunk
      const int64_t p0 = static_cast<int64_t>(prow % (g.Gh * g.Gd)) * g.Gw + x0;       // ... and its position-embedding row
#pragma unroll 2
      for (int r = rs; r < rv; r += ppp) {
        float2 a[KK / 2];
#pragma unroll
        for (int k4 = 0; k4 < KK / 4; ++k4) {
          const float4 t = *reinterpret_cast<const float4*>(s_a + r * KS + k4 * 4);
          a[k4 * 2] = make_float2(t.x, t.y);
          a[k4 * 2 + 1] = make_float2(t.z, t.w);
        }
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos != nullptr) q = __ldg(reinterpret_cast<const float4*>(pos + (p0 + r) * g.N + fg * 4));
        const float q4[4] = {q.x, q.y, q.z, q.w};
        float acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 s2 = make_float2(bv[i] + q4[i], 0.f);        // even k in .x, odd k in .y
#pragma unroll
          for (int k2 = 0; k2 < KK / 2; ++k2) s2 = pe_fma2(a[k2], wr[i][k2], s2);
          acc[i] = s2.x + s2.y;
        }
        TOut* o = out + (m0 + r) * g.N + fg * 4;
        if constexpr (sizeof(TOut) == 4) {
          *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
          __nv_bfloat162 lo = __floats2bfloat162_rn(acc[0], acc[1]), hi = __floats2bfloat162_rn(acc[2], acc[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(o) = pk;
        }
      }
    }
    __syncthreads();
  }
}

template <int KK>
void launch_fwd_stream(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos, void* out,
                       int out_is_bf16, const PEGeom& g, cudaStream_t stream) {
  const int chunks = (g.Gw + kStreamRows - 1) / kStreamRows;
  const int n_items = static_cast<int>(static_cast<int64_t>(g.B) * g.Gd * g.Gh * chunks);
  const int sms = current_device_sm_count();
  // persistent: exactly as many CTAs as are resident at once (a partial second wave would run at half occupancy), each
  // walking items with a register prefetch of the next item's pixels
#define LCBI_PE_STREAM(TI, TO)                                                                                        \
  do {                                                                                                                \
    int per_sm = 0;                                                                                                   \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, patch_embed_fwd_stream_kernel<KK, TI, TO>, 256, 0);        \
    const int want = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);                                               \
    /* wide rows (cfg5: 196 KB of output per item) are long enough per item: one item per CTA, scheduled by the */   \
    /* hardware, measured faster than the persistent walk (318 vs 366 us) */                                        \
    const unsigned blocks = static_cast<unsigned>((n_items < want || g.N >= 256) ? n_items : want);                   \
    patch_embed_fwd_stream_kernel<KK, TI, TO><<<blocks, 256, 0, stream>>>(static_cast<const TI*>(img), w, bias, pos, \
                                                                          static_cast<TO*>(out), g, chunks, n_items); \
  } while (0)
  if (img_is_bf16) {
    if (out_is_bf16) LCBI_PE_STREAM(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_STREAM(__nv_bfloat16, float);
  } else {
    if (out_is_bf16) LCBI_PE_STREAM(float, __nv_bfloat16); else LCBI_PE_STREAM(float, float);
  }
#undef LCBI_PE_STREAM
}

// ---------------------------------------------------------------------------------------------------------------------
// Weight / bias gradient for NARROW token rows (N < 256: the Swin configs, 48 / 96 features) and K in {4, 8, 16}: the
// mirror of the streaming forward. patch_embed_bwd_smallk_kernel maps a warp to one 128-feature block of one row, so at
// N = 48 twelve of its 32 lanes load 96 bytes per instruction and the kernel waits on memory (ncu: 63 % of the stall
// samples on the dOut loads, 96 us for 25 MB at cfg4). Here threads are laid out (patch, feature group) in the memory
// order of dOut - a warp's loads are one contiguous stream - the item's pixels are staged patch-major in shared memory
// with the next item prefetched into registers, each thread keeps its 4 x K partial dW and 4 partial dbias in registers
// over all items of the CTA, and the CTA folds them in shared memory before one global add per (feature, k).
// ---------------------------------------------------------------------------------------------------------------------
template <int KK, typename TIn, typename TG>
__global__ void __launch_bounds__(256, KK == 16 ? 2 : 3)
patch_embed_bwd_stream_kernel(const TIn* __restrict__ img, const TG* __restrict__ dout, float* __restrict__ dw,
                              float* __restrict__ dbias, PEGeom g, int chunks_per_row, int n_items) {
  constexpr int KS = KK + 4;
  constexpr int kPre = kStreamRows * KK / 256;
  __shared__ __align__(16) float s_a[kStreamRows * KS];
  extern __shared__ float s_fold[];                           // [N][KK + 1]: the CTA's dW rows and dbias
  const int tid = threadIdx.x;
  const int tpp = g.N >> 2, ppp = 256 / tpp;
  const int fg = tid % tpp, rs = tid / tpp;
  for (int i = tid; i < g.N * (KK + 1); i += 256) s_fold[i] = 0.f;

  const int run = kStreamRows * g.Pw;
  int e_row[kPre], e_j[kPre], e_smem[kPre];                   // e_row = channel << 16 | kz << 8 | ky
#pragma unroll
  for (int u = 0; u < kPre; ++u) {
    const int e = tid + u * 256;
    const int kr = e / run, j = e - kr * run;
    const int ky = kr % g.Ph, kz = (kr / g.Ph) % g.Pd, cin = kr / (g.Ph * g.Pd);
    const int r = j / g.Pw, px = j - r * g.Pw;
    e_row[u] = (cin << 16) | (kz << 8) | ky;
    e_j[u] = j;
    e_smem[u] = r * KS + kr * g.Pw + px;
  }
  float pre[kPre];
  auto fetch = [&](int item) {
    const int prow = item / chunks_per_row, x0 = (item - prow * chunks_per_row) * kStreamRows;
    const int gy = prow % g.Gh, t = prow / g.Gh;
    const int gz = t % g.Gd, b = t / g.Gd;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int z = gz * g.Pd + ((e_row[u] >> 8) & 0xff), y = gy * g.Ph + (e_row[u] & 0xff), x = x0 * g.Pw + e_j[u];
      float v = 0.f;
      if (z < g.D && y < g.H && x < g.W)
        v = load_px(img + (((static_cast<int64_t>(b) * g.Cin + (e_row[u] >> 16)) * g.D + z) * g.H + y) * g.W + x);
      pre[u] = v;
    }
  };

  float acc[4][KK], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[i][k] = 0.f;

  int item = blockIdx.x;
  if (item < n_items) fetch(item);
  for (; item < n_items; item += gridDim.x) {
    const int prow = item / chunks_per_row, x0 = (item - prow * chunks_per_row) * kStreamRows;
    const int rv = min(kStreamRows, g.Gw - x0);
#pragma unroll
    for (int u = 0; u < kPre; ++u) s_a[e_smem[u]] = pre[u];
    __syncthreads();
    if (item + static_cast<int>(gridDim.x) < n_items) fetch(item + gridDim.x);
    if (rs < ppp) {
      const TG* src = dout + (static_cast<int64_t>(prow) * g.Gw + x0) * g.N + fg * 4;
#pragma unroll 2
      for (int r = rs; r < rv; r += ppp) {
        float gv[4];
        if constexpr (sizeof(TG) == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src + static_cast<int64_t>(r) * g.N));
          gv[0] = t.x; gv[1] = t.y; gv[2] = t.z; gv[3] = t.w;
        } else {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src + static_cast<int64_t>(r) * g.N));
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
          gv[0] = lo.x; gv[1] = lo.y; gv[2] = hi.x; gv[3] = hi.y;
        }
        float a[KK];
#pragma unroll
        for (int k4 = 0; k4 < KK / 4; ++k4) {
          const float4 t = *reinterpret_cast<const float4*>(s_a + r * KS + k4 * 4);
          a[k4 * 4 + 0] = t.x; a[k4 * 4 + 1] = t.y; a[k4 * 4 + 2] = t.z; a[k4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          bsum[i] += gv[i];
#pragma unroll
          for (int k = 0; k < KK; ++k) acc[i][k] = fmaf(gv[i], a[k], acc[i][k]);
        }
      }
    }
    __syncthreads();
  }
  // fold the patch lanes of the CTA (ppp threads per feature group) in shared memory, then one global add per value
  if (rs < ppp) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* row = s_fold + (fg * 4 + i) * (KK + 1);
#pragma unroll
      for (int k = 0; k < KK; ++k) atomicAdd(row + k, acc[i][k]);
      atomicAdd(row + KK, bsum[i]);
    }
  }
  __syncthreads();
  for (int i = tid; i < g.N * (KK + 1); i += 256) {
    const int col = i / (KK + 1), k = i - col * (KK + 1);
    const float v = s_fold[i];
    if (k == KK) {
      if (dbias != nullptr) atomicAdd(dbias + col, v);
    } else {
      atomicAdd(dw + static_cast<int64_t>(col) * KK + k, v);
    }
  }
}

template <int KK>
void launch_bwd_stream(const void* img, int img_is_bf16, const void* dout, int dout_is_bf16, float* dw, float* dbias,
                       const PEGeom& g, cudaStream_t stream) {
  const int chunks = (g.Gw + kStreamRows - 1) / kStreamRows;
  const int n_items = static_cast<int>(static_cast<int64_t>(g.B) * g.Gd * g.Gh * chunks);
  const int sms = current_device_sm_count();
  const size_t fold_bytes = static_cast<size_t>(g.N) * (KK + 1) * sizeof(float);
#define LCBI_PE_BSTREAM(TI, TG)                                                                                        \
  do {                                                                                                                 \
    int per_sm = 0;                                                                                                    \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, patch_embed_bwd_stream_kernel<KK, TI, TG>, 256, fold_bytes); \
    const int want = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);                                                \
    const unsigned blocks = static_cast<unsigned>(n_items < want ? n_items : want);                                    \
    patch_embed_bwd_stream_kernel<KK, TI, TG><<<blocks, 256, fold_bytes, stream>>>(                                    \
        static_cast<const TI*>(img), static_cast<const TG*>(dout), dw, dbias, g, chunks, n_items);                     \
  } while (0)
  if (img_is_bf16) {
    if (dout_is_bf16) LCBI_PE_BSTREAM(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_BSTREAM(__nv_bfloat16, float);
  } else {
    if (dout_is_bf16) LCBI_PE_BSTREAM(float, __nv_bfloat16); else LCBI_PE_BSTREAM(float, float);
  }
#undef LCBI_PE_BSTREAM
}

int fill_geom(PEGeom& g, const int* img_dims, const int* patch, const int* grid, int B, int Cin, int N) {
  if (B <= 0 || Cin <= 0 || N <= 0) return LCBI_ERR_BAD_ARG;
  for (int i = 0; i < 3; ++i)
    if (img_dims[i] <= 0 || patch[i] <= 0 || grid[i] <= 0) return LCBI_ERR_BAD_ARG;
  g.B = B; g.Cin = Cin; g.D = img_dims[0]; g.H = img_dims[1]; g.W = img_dims[2];
  g.Pd = patch[0]; g.Ph = patch[1]; g.Pw = patch[2];
  g.Gd = grid[0]; g.Gh = grid[1]; g.Gw = grid[2];
  g.N = N;
  g.K = Cin * patch[0] * patch[1] * patch[2];
  g.M = static_cast<int64_t>(B) * grid[0] * grid[1] * grid[2];
  g.d_Gw = make_fastdiv(g.Gw); g.d_Gh = make_fastdiv(g.Gh); g.d_Gd = make_fastdiv(g.Gd);
  g.d_Pw = make_fastdiv(g.Pw); g.d_Ph = make_fastdiv(g.Ph); g.d_Pd = make_fastdiv(g.Pd);
  return LCBI_OK;
}

}  // namespace

int patch_embed_fwd_launch(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                           void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                           const int* grid, int N, cudaStream_t stream) {
  PEGeom g;
  int rc = fill_geom(g, img_dims, patch, grid, B, Cin, N);
  if (rc) return rc;
  // K in {4, 8, 16} whole image rows, N a multiple of 4 up to 1024, 16-byte aligned rows of w / bias / pos / out
  const bool stream_ok = (g.K == 4 || g.K == 8 || g.K == 16) && g.K % g.Pw == 0 && (N & 3) == 0 && N <= 1024 &&
                         g.Pd < 256 && g.Ph < 256 && Cin < 32768 &&
                         static_cast<int64_t>(B) * g.Gd * g.Gh * ((g.Gw + kStreamRows - 1) / kStreamRows) < (1ll << 31) &&
                         ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(pos) |
                           reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (stream_ok) {
    if (g.K == 4) launch_fwd_stream<4>(img, img_is_bf16, w, bias, pos, out, out_is_bf16, g, stream);
    else if (g.K == 8) launch_fwd_stream<8>(img, img_is_bf16, w, bias, pos, out, out_is_bf16, g, stream);
    else launch_fwd_stream<16>(img, img_is_bf16, w, bias, pos, out, out_is_bf16, g, stream);
    return set_cuda_error(cudaGetLastError());
  }
  if (patch_embed_mma_applicable(Cin, img_dims, patch, grid, N))
    return patch_embed_fwd_mma_launch(img, img_is_bf16, w, bias, pos, out, out_is_bf16, B, Cin, img_dims, patch, grid, N,
                                      stream);
  dim3 grd(static_cast<unsigned>((g.M + TM - 1) / TM), (N + TN - 1) / TN);
#define LCBI_PE_FWD(TI, TO) \
  patch_embed_fwd_kernel<TI, TO><<<grd, kThreads, 0, stream>>>(static_cast<const TI*>(img), w, bias, pos, \
                                                              static_cast<TO*>(out), g)
  if (img_is_bf16) {
    if (out_is_bf16) LCBI_PE_FWD(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_FWD(__nv_bfloat16, float);
  } else {
    if (out_is_bf16) LCBI_PE_FWD(float, __nv_bfloat16); else LCBI_PE_FWD(float, float);
  }
#undef LCBI_PE_FWD
  return set_cuda_error(cudaGetLastError());
}

int patch_embed_bwd_launch(const void* img, int img_is_bf16, const float* w, const void* dout, int dout_is_bf16,
                           float* dw, float* dbias, float* dpos, float* dimg, int B, int Cin, const int* img_dims,
                           const int* patch, const int* grid, int N, cudaStream_t stream, void* workspace,
                           size_t workspace_bytes) {
  PEGeom g;
  int rc = fill_geom(g, img_dims, patch, grid, B, Cin, N);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * static_cast<size_t>(N) * g.K, stream);
  if (e != cudaSuccess) return set_cuda_error(e);
  if (dbias) {
    e = cudaMemsetAsync(dbias, 0, sizeof(float) * N, stream);
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  const int64_t np = static_cast<int64_t>(g.Gd) * g.Gh * g.Gw;
  bool dpos_done = false;
  const bool narrow_stream = (g.K == 4 || g.K == 8 || g.K == 16) && g.K % g.Pw == 0 && (N & 3) == 0 && N < 256 &&
                             g.Pd < 256 && g.Ph < 256 && Cin < 32768 &&
                             static_cast<int64_t>(B) * g.Gd * g.Gh * ((g.Gw + kStreamRows - 1) / kStreamRows) < (1ll << 31) &&
                             (reinterpret_cast<uintptr_t>(dout) & 15) == 0;
  if (narrow_stream) {
    if (g.K == 4) launch_bwd_stream<4>(img, img_is_bf16, dout, dout_is_bf16, dw, dbias, g, stream);
    else if (g.K == 8) launch_bwd_stream<8>(img, img_is_bf16, dout, dout_is_bf16, dw, dbias, g, stream);
    else launch_bwd_stream<16>(img, img_is_bf16, dout, dout_is_bf16, dw, dbias, g, stream);
  } else if (g.K <= kBwKMax && (N & 3) == 0) {
    // single pass over dOut; with B == 1 the position-embedding gradient is a cast/copy of dOut and rides along
    float* dpos_copy = (dpos != nullptr && B == 1) ? dpos : nullptr;
    dpos_done = dpos_copy != nullptr;
    const int col_blocks = (N + 127) / 128;
    int64_t slabs = (g.M + kBwRows - 1) / kBwRows;
    // ~6 persistent CTAs per SM in total for latency hiding, but never more than 2*148 CTAs adding into the same
    // dW/dbias addresses (measured: atomic contention dominates the narrow-N Swin cases beyond that)
    int64_t want = (6 * 148 + col_blocks - 1) / col_blocks;
    if (want > 2 * 148) want = 2 * 148;
    dim3 sg(static_cast<unsigned>(slabs < want ? slabs : want), col_blocks);
    const size_t sk_smem = (static_cast<size_t>(kBwRows) * (kBwKMax + 1) + static_cast<size_t>(8) * (g.K + 1) * 132) * 4;
#define LCBI_PE_BSK(TI, TG)                                                                                      \
  do {                                                                                                           \
    if (sk_smem > 48 * 1024)                                                                                     \
      cudaFuncSetAttribute(patch_embed_bwd_smallk_kernel<TI, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           static_cast<int>(sk_smem));                                                           \
    patch_embed_bwd_smallk_kernel<TI, TG><<<sg, 256, sk_smem, stream>>>(                                         \
        static_cast<const TI*>(img), static_cast<const TG*>(dout), dw, dbias, dpos_copy, g);                     \
  } while (0)
    if (img_is_bf16) {
      if (dout_is_bf16) LCBI_PE_BSK(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_BSK(__nv_bfloat16, float);
    } else {
      if (dout_is_bf16) LCBI_PE_BSK(float, __nv_bfloat16); else LCBI_PE_BSK(float, float);
    }
#undef LCBI_PE_BSK
  } else if (workspace != nullptr && !dout_is_bf16 && patch_embed_tc_applicable(img_is_bf16, Cin, img_dims, patch, grid, N)) {
    rc = patch_embed_tc_bwd_w_launch(img, static_cast<const float*>(dout), dw, dbias, B, Cin, img_dims, patch, grid, N, workspace,
                                     workspace_bytes, stream);
    if (rc) return rc;
  } else if (patch_embed_mma_applicable(Cin, img_dims, patch, grid, N)) {
    rc = patch_embed_bwd_w_mma_launch(img, img_is_bf16, dout, dout_is_bf16, dw, dbias, B, Cin, img_dims, patch, grid, N,
                                      stream);
    if (rc) return rc;
  } else {
    dim3 grd(static_cast<unsigned>((g.M + MS - 1) / MS), (N + TN - 1) / TN, (g.K + KC - 1) / KC);
#define LCBI_PE_BWD(TI, TG) \
  patch_embed_bwd_w_kernel<TI, TG><<<grd, kThreads, 0, stream>>>(static_cast<const TI*>(img), \
                                                                static_cast<const TG*>(dout), dw, dbias, g)
    if (img_is_bf16) {
      if (dout_is_bf16) LCBI_PE_BWD(__nv_bfloat16, __nv_bfloat16); else LCBI_PE_BWD(__nv_bfloat16, float);
    } else {
      if (dout_is_bf16) LCBI_PE_BWD(float, __nv_bfloat16); else LCBI_PE_BWD(float, float);
    }
#undef LCBI_PE_BWD
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);
  if (dpos && !dpos_done) {
    if ((N & 3) == 0) {
      const int64_t total4 = np * N / 4;
      const unsigned blocks = static_cast<unsigned>((total4 + 255) / 256);
      if (dout_is_bf16)
        patch_embed_bwd_pos4_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dout), dpos, B, total4);
      else
        patch_embed_bwd_pos4_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(dout), dpos, B, total4);
    } else {
      const int64_t total = np * N;
      const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
      if (dout_is_bf16)
        patch_embed_bwd_pos_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dout), dpos, B, np, N);
      else
        patch_embed_bwd_pos_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(dout), dpos, B, np, N);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  if (dimg) {
    const int64_t total = static_cast<int64_t>(B) * Cin * g.D * g.H * g.W;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    if (dout_is_bf16)
      patch_embed_bwd_x_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dout), w, dimg, g);
    else
      patch_embed_bwd_x_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(dout), w, dimg, g);
    e = cudaGetLastError();
  }
  return set_cuda_error(e);
}

}  // namespace lcbi
