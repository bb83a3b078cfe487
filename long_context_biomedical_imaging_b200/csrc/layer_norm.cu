// LayerNorm over the channel axis of token rows, forward and backward, for sm_100a.
//
// The normalisations that bracket the attention kernels in the reference's encoder blocks: `self.norm1(x)` in front
// of the qkv projection and `self.norm2(x)` in front of the MLP (/root/reference/model/models/backbone_vit.py:260-263,
// backbone_swin.py:437,489-490), torch.nn.LayerNorm(hidden) with eps 1e-5 evaluated in fp32 (layer_norm is an
// autocast-to-fp32 op). Here the forward writes the normalised rows directly in the dtype the following Linear
// consumes (bf16 under bf16 autocast: the same round-to-nearest the autocast cast applies to the fp32 result, minus
// one full-tensor cast kernel), and the backward is two HBM-bound kernels: dx row by row, and d(gamma) / d(beta) as
// a column reduction over row slabs followed by a small deterministic sum of the slab partials (no atomics).
//   fwd       : y = (x - mean) * rstd * gamma + beta                       warp = row
//   bwd (dx)  : dx = rstd * (g dy - mean_c(g dy) - xhat mean_c(g dy xhat))  warp = row
//   bwd (g,b) : dgamma_c = sum_r dy xhat, dbeta_c = sum_r dy               CTA = 128 columns x one slab of rows
// Rows are C contiguous elements (C % 4 == 0), fp32 or bf16 in, fp32 or bf16 out; statistics in fp32, two-pass
// variance (mean first, then centred squares) like the reference's kernel.
#include <cuda_bf16.h>

#include <type_traits>

#include "lcbi_kernels.h"

namespace lcbi {

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kCache = 8;   // float4 per lane kept in registers: rows up to 1024 channels are read from memory once
constexpr int kCacheDx = 6; // the dx kernel keeps two values per element: 768 channels in registers, 3 CTAs per SM

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&a);
  o.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = o;
}
// the value a float4 has after a round trip through storage type T (identity for fp32)
template <typename T>
__device__ __forceinline__ float4 load4_rounded(float4 v) {
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  } else {
    return v;
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// With `delta` (the residual-branch output: attention or MLP result) the kernel first forms the new residual stream
// xsum = x + delta, rounded to x's dtype exactly like the reference's `x = x + self.attn(...)` (backbone_vit.py:261-262),
// writes it out, and normalises THAT: the block's residual add costs no pass of its own. The later passes re-read xsum.
template <typename TX, typename TY, typename TD>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ln_fwd_kernel(const TX* __restrict__ x, const TD* __restrict__ delta, TX* __restrict__ xsum,
              const float* __restrict__ gamma, const float* __restrict__ beta,
              TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int C,
              float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (r >= rows) return;
  const TX* xin = x + r * C;
  // what the later passes read for rows longer than the register cache: the sum this thread wrote, if there is one
  const TX* xr = delta != nullptr ? xsum + r * C : xin;
  const int nvec = C >> 2;
  auto fetch = [&](int v) {
    float4 t = load4(xin + 4 * v);
    if (delta != nullptr) {
      const float4 d = load4(delta + r * C + 4 * v);
      t = make_float4(t.x + d.x, t.y + d.y, t.z + d.z, t.w + d.w);
      store4(xsum + r * C + 4 * v, t);
      t = load4_rounded<TX>(t);
    }
    return t;
  };
  float4 cache[kCache];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kCache; ++i) {
    const int v = lane + 32 * i;
    cache[i] = v < nvec ? fetch(v) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (cache[i].x + cache[i].y) + (cache[i].z + cache[i].w);
  }
  for (int v = lane + 32 * kCache; v < nvec; v += 32) {
    const float4 t = fetch(v);
    s += (t.x + t.y) + (t.z + t.w);
  }
  const float mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kCache; ++i) {
    if (lane + 32 * i < nvec) {
      const float a = cache[i].x - mean, b = cache[i].y - mean, c = cache[i].z - mean, d = cache[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
  for (int v = lane + 32 * kCache; v < nvec; v += 32) {
    const float4 t = load4(xr + 4 * v);
    const float a = t.x - mean, b = t.y - mean, c = t.z - mean, d = t.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
  if (lane == 0) {
    mean_out[r] = mean;
    rstd_out[r] = rstd;
  }
  TY* yr = y + r * C;
  auto emit = [&](int v, float4 t) {
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gamma != nullptr) g = load4(gamma + 4 * v);
    if (beta != nullptr) b = load4(beta + 4 * v);
    store4(yr + 4 * v, make_float4((t.x - mean) * rstd * g.x + b.x, (t.y - mean) * rstd * g.y + b.y,
                                   (t.z - mean) * rstd * g.z + b.z, (t.w - mean) * rstd * g.w + b.w));
  };
#pragma unroll
  for (int i = 0; i < kCache; ++i)
    if (lane + 32 * i < nvec) emit(lane + 32 * i, cache[i]);
  for (int v = lane + 32 * kCache; v < nvec; v += 32) emit(v, load4(xr + 4 * v));
}

// Backward of the fused form: `dres` (gradient arriving at xsum from everything downstream of the residual stream,
// may be null) is added to the LayerNorm's dx, and the total is written twice, as the gradient of x (TX) and as the
// gradient of delta (TD, may be null) — the add's backward and the dtype cast cost no passes of their own either.
template <typename TX, typename TY, typename TD>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3)
ln_bwd_dx_kernel(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ mean_in, const float* __restrict__ rstd_in, TX* __restrict__ dx,
                 const TX* __restrict__ dres, TD* __restrict__ ddelta, int64_t rows, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (r >= rows) return;
  const TX* xr = x + r * C;
  const TY* dyr = dy + r * C;
  const int nvec = C >> 2;
  const float mean = mean_in[r], rstd = rstd_in[r];
  // per element: w = gamma * dy, h = xhat
  auto wh = [&](int v, float4& w, float4& h) {
    const float4 t = load4(xr + 4 * v), d = load4(dyr + 4 * v);
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gamma != nullptr) g = load4(gamma + 4 * v);
    w = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
    h = make_float4((t.x - mean) * rstd, (t.y - mean) * rstd, (t.z - mean) * rstd, (t.w - mean) * rstd);
  };
  float4 cw[kCacheDx], ch[kCacheDx];
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int i = 0; i < kCacheDx; ++i) {
    if (lane + 32 * i < nvec) {
      wh(lane + 32 * i, cw[i], ch[i]);
      a += (cw[i].x + cw[i].y) + (cw[i].z + cw[i].w);
      b += (cw[i].x * ch[i].x + cw[i].y * ch[i].y) + (cw[i].z * ch[i].z + cw[i].w * ch[i].w);
    }
  }
  for (int v = lane + 32 * kCacheDx; v < nvec; v += 32) {
    float4 w, h;
    wh(v, w, h);
    a += (w.x + w.y) + (w.z + w.w);
    b += (w.x * h.x + w.y * h.y) + (w.z * h.z + w.w * h.w);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  a = warp_sum(a) * inv_c;
  b = warp_sum(b) * inv_c;
  TX* dxr = dx + r * C;
  auto emit = [&](int v, const float4& w, const float4& h) {
    float4 o = make_float4(rstd * (w.x - a - h.x * b), rstd * (w.y - a - h.y * b), rstd * (w.z - a - h.z * b),
                           rstd * (w.w - a - h.w * b));
    if (dres != nullptr) {
      const float4 e = load4(dres + r * C + 4 * v);
      o = make_float4(o.x + e.x, o.y + e.y, o.z + e.z, o.w + e.w);
    }
    store4(dxr + 4 * v, o);
    if (ddelta != nullptr) store4(ddelta + r * C + 4 * v, o);
  };
#pragma unroll
  for (int i = 0; i < kCacheDx; ++i)
    if (lane + 32 * i < nvec) emit(lane + 32 * i, cw[i], ch[i]);
  for (int v = lane + 32 * kCacheDx; v < nvec; v += 32) {
    float4 w, h;
    wh(v, w, h);
    emit(v, w, h);
  }
}

// d(gamma), d(beta) partials: CTA (32 x 8 threads) = 128 columns x the rows of slab blockIdx.y; thread (tx, ty) walks
// rows ty, ty + 8, ... of the slab with four columns in registers, then the 8 row lanes are summed through shared
// memory and written to partial[slab][2][C].
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
ln_bwd_params_kernel(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, float* __restrict__ partial, int64_t rows, int C,
                     int64_t rows_per_slab) {
  __shared__ float4 sg[8][32], sb[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_slab;
  const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < C) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float mean = mean_in[r], rstd = rstd_in[r];
      const float4 t = load4(x + r * C + col), d = load4(dy + r * C + col);
      g.x += d.x * (t.x - mean) * rstd; g.y += d.y * (t.y - mean) * rstd;
      g.z += d.z * (t.z - mean) * rstd; g.w += d.w * (t.w - mean) * rstd;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
  }
  sg[ty][tx] = g;
  sb[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && col < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 g2 = sg[k][tx], b2 = sb[k][tx];
      g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
      b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
    }
    float* out = partial + static_cast<int64_t>(blockIdx.y) * 2 * C;
    store4(out + col, g);
    store4(out + C + col, b);
  }
}

// sum of the slab partials, one warp per output column (dgamma columns first, then dbeta): lanes take slabs
// lane, lane + 32, ... and are combined by a fixed shuffle tree, so the result does not depend on timing
__global__ void ln_bwd_params_finish_kernel(const float* __restrict__ partial, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int C, int slabs) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= 2 * C) return;
  float s = 0.f;
  for (int k = threadIdx.x & 31; k < slabs; k += 32) s += partial[static_cast<int64_t>(k) * 2 * C + c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) {
    if (c < C) {
      if (dgamma != nullptr) dgamma[c] = s;
    } else if (dbeta != nullptr) {
      dbeta[c - C] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Narrow rows (C <= 128: Swin stage-1 tokens, 48 / 96 / 128 channels). One warp per row would leave most lanes idle
// and give every warp a few hundred bytes of work, so LPR lanes (a power of two >= C / 4) share a row, a warp holds
// 32 / LPR rows side by side, and every warp walks kNarrowIter such row groups with all their loads issued up front.
// ------------------------------------------------------------------------------------------------
constexpr int kNarrowIter = 4;

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TX, typename TY, int LPR>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ln_fwd_narrow_kernel(const TX* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows,
                     int C, float eps) {
  constexpr int kRpw = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const bool active = sub < (C >> 2);
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t row0 = warp * (kRpw * kNarrowIter) + lane / LPR;
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active && gamma != nullptr) g = load4(gamma + 4 * sub);
  if (active && beta != nullptr) b = load4(beta + 4 * sub);
  float4 v[kNarrowIter];
#pragma unroll
  for (int it = 0; it < kNarrowIter; ++it) {
    const int64_t r = row0 + it * kRpw;
    v[it] = (active && r < rows) ? load4(x + r * C + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
#pragma unroll
  for (int it = 0; it < kNarrowIter; ++it) {
    const int64_t r = row0 + it * kRpw;
    const float mean = group_sum<LPR>((v[it].x + v[it].y) + (v[it].z + v[it].w)) * inv_c;
    const float dx0 = v[it].x - mean, dx1 = v[it].y - mean, dx2 = v[it].z - mean, dx3 = v[it].w - mean;
    const float q = active ? (dx0 * dx0 + dx1 * dx1) + (dx2 * dx2 + dx3 * dx3) : 0.f;
    const float rstd = rsqrtf(group_sum<LPR>(q) * inv_c + eps);
    if (r < rows) {
      if (sub == 0) {
        mean_out[r] = mean;
        rstd_out[r] = rstd;
      }
      if (active)
        store4(y + r * C + 4 * sub, make_float4(dx0 * rstd * g.x + b.x, dx1 * rstd * g.y + b.y, dx2 * rstd * g.z + b.z,
                                               dx3 * rstd * g.w + b.w));
    }
  }
}

template <typename TX, typename TY, int LPR>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
ln_bwd_dx_narrow_kernel(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ mean_in, const float* __restrict__ rstd_in, TX* __restrict__ dx,
                        int64_t rows, int C) {
  constexpr int kRpw = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const bool active = sub < (C >> 2);
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  const int64_t row0 = warp * (kRpw * kNarrowIter) + lane / LPR;
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
  if (active && gamma != nullptr) g = load4(gamma + 4 * sub);
  float4 xv[kNarrowIter], dv[kNarrowIter];
  float mean[kNarrowIter], rstd[kNarrowIter];
#pragma unroll
  for (int it = 0; it < kNarrowIter; ++it) {
    const int64_t r = row0 + it * kRpw;
    const bool ok = active && r < rows;
    xv[it] = ok ? load4(x + r * C + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    dv[it] = ok ? load4(dy + r * C + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    mean[it] = r < rows ? mean_in[r] : 0.f;
    rstd[it] = r < rows ? rstd_in[r] : 0.f;
  }
  const float inv_c = 1.0f / static_cast<float>(C);
#pragma unroll
  for (int it = 0; it < kNarrowIter; ++it) {
    const int64_t r = row0 + it * kRpw;
    const float4 w = make_float4(dv[it].x * g.x, dv[it].y * g.y, dv[it].z * g.z, dv[it].w * g.w);
    float4 h = make_float4((xv[it].x - mean[it]) * rstd[it], (xv[it].y - mean[it]) * rstd[it],
                           (xv[it].z - mean[it]) * rstd[it], (xv[it].w - mean[it]) * rstd[it]);
    if (!active) h = make_float4(0.f, 0.f, 0.f, 0.f);
    const float a = group_sum<LPR>((w.x + w.y) + (w.z + w.w)) * inv_c;
    const float b = group_sum<LPR>((w.x * h.x + w.y * h.y) + (w.z * h.z + w.w * h.w)) * inv_c;
    if (active && r < rows)
      store4(dx + r * C + 4 * sub, make_float4(rstd[it] * (w.x - a - h.x * b), rstd[it] * (w.y - a - h.y * b),
                                              rstd[it] * (w.z - a - h.z * b), rstd[it] * (w.w - a - h.w * b)));
  }
}

// d(gamma), d(beta) partials for narrow rows: the CTA's 256 threads tile (rows x C/4 column vectors) densely — thread t
// owns column vector t % nvec of rows t / nvec, t / nvec + 256 / nvec, ... of its slab — so consecutive threads read
// consecutive addresses whatever C is; the row lanes are then summed through shared memory in a fixed order.
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
ln_bwd_params_narrow_kernel(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean_in,
                            const float* __restrict__ rstd_in, float* __restrict__ partial, int64_t rows, int C,
                            int64_t rows_per_slab) {
  __shared__ float4 sg[256], sb[256];
  const int nvec = C >> 2;
  const int rows_per_iter = 256 / nvec;
  const int t = threadIdx.x, cv = t % nvec, rl = t / nvec;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_slab;
  const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < rows_per_iter) {
    for (int64_t r = r0 + rl; r < r1; r += rows_per_iter) {
      const float mean = mean_in[r], rstd = rstd_in[r];
      const float4 xt = load4(x + r * C + 4 * cv), d = load4(dy + r * C + 4 * cv);
      g.x += d.x * (xt.x - mean) * rstd; g.y += d.y * (xt.y - mean) * rstd;
      g.z += d.z * (xt.z - mean) * rstd; g.w += d.w * (xt.w - mean) * rstd;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
  }
  sg[t] = g;
  sb[t] = b;
  __syncthreads();
  if (t < nvec) {
    for (int k = 1; k < rows_per_iter; ++k) {
      const float4 g2 = sg[t + k * nvec], b2 = sb[t + k * nvec];
      g.x += g2.x; g.y += g2.y; g.z += g2.z; g.w += g2.w;
      b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
    }
    float* out = partial + static_cast<int64_t>(blockIdx.x) * 2 * C;
    store4(out + 4 * t, g);
    store4(out + C + 4 * t, b);
  }
}

// ------------------------------------------------------------------------------------------------
// Column sums of a (rows, C) matrix: the bias gradient of a token-wise Linear (db = sum over tokens of dY), which
// autograd otherwise computes with a generic reduction that costs as much as the LayerNorm kernels above. Same slab
// scheme and the same deterministic finish as d(gamma) / d(beta); the partial buffer's second half stays unused.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ a, float* __restrict__ partial, int64_t rows, int C, int64_t rows_per_slab) {
  __shared__ float4 ss[256];
  const int nvec = C >> 2;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_slab;
  const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // wide rows: thread (tx, ty) = column vector blockIdx.x * 32 + tx, rows ty, ty + 8, ...; narrow rows (C <= 128, one
  // column block): the 256 threads tile (rows x column vectors) densely
  const bool narrow = nvec <= 32;
  const int lanes_per_row = narrow ? nvec : 32;
  const int rows_per_iter = narrow ? 256 / nvec : 8;
  const int t = threadIdx.x;
  const int cv = narrow ? t % nvec : blockIdx.x * 32 + (t & 31);
  const int rl = narrow ? t / nvec : t >> 5;
  if (cv < nvec && rl < rows_per_iter) {
    for (int64_t r = r0 + rl; r < r1; r += rows_per_iter) {
      const float4 v = load4(a + r * C + 4 * cv);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  ss[t] = acc;
  __syncthreads();
  if (rl == 0 && cv < nvec) {
    for (int k = 1; k < rows_per_iter; ++k) {
      const float4 v = ss[t + k * lanes_per_row];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    store4(partial + static_cast<int64_t>(blockIdx.y) * 2 * C + 4 * cv, acc);
  }
}

int param_slabs(int64_t rows, int C) {
  const int col_blocks = (C + 127) / 128;
  int64_t slabs = (4 * 148 + col_blocks - 1) / col_blocks;   // ~4 CTAs per SM in total
  const int64_t max_slabs = (rows + 63) / 64;                // at least 64 rows per slab
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  return static_cast<int>(slabs);
}

template <typename TX, typename TY>
int fwd_typed(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows,
              int C, float eps, cudaStream_t stream) {
  if (C <= 128) {
    auto launch = [&](auto lpr) {
      constexpr int LPR = decltype(lpr)::value;
      const int64_t rows_per_cta = static_cast<int64_t>(kWarpsPerCta) * (32 / LPR) * kNarrowIter;
      const int64_t blocks = (rows + rows_per_cta - 1) / rows_per_cta;
      ln_fwd_narrow_kernel<TX, TY, LPR><<<static_cast<unsigned>(blocks), kWarpsPerCta * 32, 0, stream>>>(
          static_cast<const TX*>(x), gamma, beta, static_cast<TY*>(y), mean, rstd, rows, C, eps);
    };
    if (C <= 32) launch(std::integral_constant<int, 8>{});
    else if (C <= 64) launch(std::integral_constant<int, 16>{});
    else launch(std::integral_constant<int, 32>{});
    return set_cuda_error(cudaGetLastError());
  }
  const int64_t blocks = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  ln_fwd_kernel<TX, TY, TX><<<static_cast<unsigned>(blocks), kWarpsPerCta * 32, 0, stream>>>(
      static_cast<const TX*>(x), nullptr, nullptr, gamma, beta, static_cast<TY*>(y), mean, rstd, rows, C, eps);
  return set_cuda_error(cudaGetLastError());
}

template <typename TX, typename TY, typename TD>
int add_fwd_typed(const void* x, const void* delta, void* xsum, const float* gamma, const float* beta, void* y,
                  float* mean, float* rstd, int64_t rows, int C, float eps, cudaStream_t stream) {
  const int64_t blocks = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  ln_fwd_kernel<TX, TY, TD><<<static_cast<unsigned>(blocks), kWarpsPerCta * 32, 0, stream>>>(
      static_cast<const TX*>(x), static_cast<const TD*>(delta), static_cast<TX*>(xsum), gamma, beta,
      static_cast<TY*>(y), mean, rstd, rows, C, eps);
  return set_cuda_error(cudaGetLastError());
}

template <typename TX, typename TY, typename TD>
int add_bwd_dx_typed(const void* dy, const void* xsum, const float* gamma, const float* mean, const float* rstd,
                     void* dx, const void* dres, void* ddelta, int64_t rows, int C, cudaStream_t stream) {
  const int64_t blocks = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  ln_bwd_dx_kernel<TX, TY, TD><<<static_cast<unsigned>(blocks), kWarpsPerCta * 32, 0, stream>>>(
      static_cast<const TY*>(dy), static_cast<const TX*>(xsum), gamma, mean, rstd, static_cast<TX*>(dx),
      static_cast<const TX*>(dres), static_cast<TD*>(ddelta), rows, C);
  return set_cuda_error(cudaGetLastError());
}

template <typename TX, typename TY>
int bwd_typed(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
              float* dgamma, float* dbeta, float* workspace, int64_t rows, int C, cudaStream_t stream) {
  if (dx != nullptr && C <= 128) {
    auto launch = [&](auto lpr) {
      constexpr int LPR = decltype(lpr)::value;
      const int64_t rows_per_cta = static_cast<int64_t>(kWarpsPerCta) * (32 / LPR) * kNarrowIter;
      const int64_t blocks = (rows + rows_per_cta - 1) / rows_per_cta;
      ln_bwd_dx_narrow_kernel<TX, TY, LPR><<<static_cast<unsigned>(blocks), kWarpsPerCta * 32, 0, stream>>>(
          static_cast<const TY*>(dy), static_cast<const TX*>(x), gamma, mean, rstd, static_cast<TX*>(dx), rows, C);
    };
    if (C <= 32) launch(std::integral_constant<int, 8>{});
    else if (C <= 64) launch(std::integral_constant<int, 16>{});
    else launch(std::integral_constant<int, 32>{});
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  } else if (dx != nullptr) {
    const int64_t blocks = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
    ln_bwd_dx_kernel<TX, TY, TX><<<static_cast<unsigned>(blocks), kWarpsPerCta * 32, 0, stream>>>(
        static_cast<const TY*>(dy), static_cast<const TX*>(x), gamma, mean, rstd, static_cast<TX*>(dx), nullptr, nullptr,
        rows, C);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  if (dgamma != nullptr || dbeta != nullptr) {
    const int slabs = param_slabs(rows, C);
    const int64_t rows_per_slab = (rows + slabs - 1) / slabs;
    if (C <= 128) {
      ln_bwd_params_narrow_kernel<TX, TY><<<slabs, 256, 0, stream>>>(static_cast<const TY*>(dy), static_cast<const TX*>(x),
                                                                     mean, rstd, workspace, rows, C, rows_per_slab);
    } else {
      dim3 grid((C + 127) / 128, slabs);
      ln_bwd_params_kernel<TX, TY><<<grid, 256, 0, stream>>>(static_cast<const TY*>(dy), static_cast<const TX*>(x), mean,
                                                             rstd, workspace, rows, C, rows_per_slab);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
    ln_bwd_params_finish_kernel<<<(2 * C + 7) / 8, 256, 0, stream>>>(workspace, dgamma, dbeta, C, slabs);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  return LCBI_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

size_t layer_norm_bwd_workspace_bytes(int64_t rows, int C) {
  if (rows <= 0 || C <= 0) return 0;
  return static_cast<size_t>(param_slabs(rows, C)) * 2 * C * sizeof(float);
}

// fused residual add + LayerNorm (rows of at least 128 channels; the narrow-row kernels have no fused form)
int add_layer_norm_fwd_launch(const void* x, int x_is_bf16, const void* delta, int delta_is_bf16, void* xsum,
                              const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean, float* rstd,
                              int64_t rows, int C, float eps, cudaStream_t stream) {
  if (rows <= 0 || C <= 0 || !(eps >= 0.f)) return LCBI_ERR_BAD_ARG;
  if (C % 4 != 0 || C < 128) return LCBI_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(delta) || !aligned16(xsum) || !aligned16(y) || (gamma && !aligned16(gamma)) ||
      (beta && !aligned16(beta)))
    return LCBI_ERR_BAD_ARG;
#define LCBI_LN_DISPATCH3(FN, ...)                                                                            \
  (x_is_bf16 ? (y_is_bf16 ? (delta_is_bf16 ? FN<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(__VA_ARGS__)     \
                                           : FN<__nv_bfloat16, __nv_bfloat16, float>(__VA_ARGS__))            \
                          : (delta_is_bf16 ? FN<__nv_bfloat16, float, __nv_bfloat16>(__VA_ARGS__)             \
                                           : FN<__nv_bfloat16, float, float>(__VA_ARGS__)))                   \
             : (y_is_bf16 ? (delta_is_bf16 ? FN<float, __nv_bfloat16, __nv_bfloat16>(__VA_ARGS__)             \
                                           : FN<float, __nv_bfloat16, float>(__VA_ARGS__))                    \
                          : (delta_is_bf16 ? FN<float, float, __nv_bfloat16>(__VA_ARGS__)                     \
                                           : FN<float, float, float>(__VA_ARGS__))))
  return LCBI_LN_DISPATCH3(add_fwd_typed, x, delta, xsum, gamma, beta, y, mean, rstd, rows, C, eps, stream);
}

int add_layer_norm_bwd_launch(const void* dy, int y_is_bf16, const void* dres, const void* xsum, int x_is_bf16,
                              const float* gamma, const float* mean, const float* rstd, void* dx, void* ddelta,
                              int delta_is_bf16, float* dgamma, float* dbeta, float* workspace, size_t workspace_bytes,
                              int64_t rows, int C, cudaStream_t stream) {
  if (rows <= 0 || C <= 0) return LCBI_ERR_BAD_ARG;
  if (C % 4 != 0 || C < 128) return LCBI_ERR_UNSUPPORTED;
  if (!aligned16(dy) || !aligned16(xsum) || !aligned16(dx) || (dres && !aligned16(dres)) ||
      (ddelta && !aligned16(ddelta)) || (gamma && !aligned16(gamma)))
    return LCBI_ERR_BAD_ARG;
  int rc = LCBI_LN_DISPATCH3(add_bwd_dx_typed, dy, xsum, gamma, mean, rstd, dx, dres, ddelta, rows, C, stream);
#undef LCBI_LN_DISPATCH3
  if (rc != LCBI_OK) return rc;
  if (dgamma != nullptr || dbeta != nullptr)   // parameter gradients: the plain path with dx skipped
    return layer_norm_bwd_launch(dy, y_is_bf16, xsum, x_is_bf16, gamma, mean, rstd, nullptr, dgamma, dbeta, workspace,
                                 workspace_bytes, rows, C, stream);
  return LCBI_OK;
}

int bias_grad_launch(const void* dy, int dy_is_bf16, float* dbias, float* workspace, size_t workspace_bytes,
                     int64_t rows, int C, cudaStream_t stream) {
  if (rows <= 0 || C <= 0) return LCBI_ERR_BAD_ARG;
  if (C % 4 != 0) return LCBI_ERR_UNSUPPORTED;
  if (!aligned16(dy) || !aligned16(dbias)) return LCBI_ERR_BAD_ARG;
  if (workspace == nullptr || !aligned16(workspace) || workspace_bytes < layer_norm_bwd_workspace_bytes(rows, C))
    return LCBI_ERR_WORKSPACE;
  const int slabs = param_slabs(rows, C);
  const int64_t rows_per_slab = (rows + slabs - 1) / slabs;
  dim3 grid((C + 127) / 128, slabs);
  if (dy_is_bf16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy), workspace, rows, C, rows_per_slab);
  else
    colsum_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(dy), workspace, rows, C, rows_per_slab);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);
  // only the first C columns of each slab's partial row carry data: sum those into dbias
  ln_bwd_params_finish_kernel<<<(C + 7) / 8, 256, 0, stream>>>(workspace, dbias, nullptr, C, slabs);
  return set_cuda_error(cudaGetLastError());
}

int layer_norm_fwd_launch(const void* x, int x_is_bf16, const float* gamma, const float* beta, void* y, int y_is_bf16,
                          float* mean, float* rstd, int64_t rows, int C, float eps, cudaStream_t stream) {
  if (rows <= 0 || C <= 0 || !(eps >= 0.f)) return LCBI_ERR_BAD_ARG;
  if (C % 4 != 0 || rows >= (int64_t(1) << 31) * kWarpsPerCta) return LCBI_ERR_UNSUPPORTED;
  // rows of bf16 start on 8-byte boundaries when C % 4 == 0; the base pointers must be 16-byte aligned
  if (!aligned16(x) || !aligned16(y) || (gamma && !aligned16(gamma)) || (beta && !aligned16(beta))) return LCBI_ERR_BAD_ARG;
  if (x_is_bf16) {
    return y_is_bf16 ? fwd_typed<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, C, eps, stream)
                     : fwd_typed<__nv_bfloat16, float>(x, gamma, beta, y, mean, rstd, rows, C, eps, stream);
  }
  return y_is_bf16 ? fwd_typed<float, __nv_bfloat16>(x, gamma, beta, y, mean, rstd, rows, C, eps, stream)
                   : fwd_typed<float, float>(x, gamma, beta, y, mean, rstd, rows, C, eps, stream);
}

int layer_norm_bwd_launch(const void* dy, int dy_is_bf16, const void* x, int x_is_bf16, const float* gamma,
                          const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                          float* workspace, size_t workspace_bytes, int64_t rows, int C, cudaStream_t stream) {
  if (rows <= 0 || C <= 0) return LCBI_ERR_BAD_ARG;
  if (C % 4 != 0 || rows >= (int64_t(1) << 31) * kWarpsPerCta) return LCBI_ERR_UNSUPPORTED;
  if (!aligned16(dy) || !aligned16(x) || (dx && !aligned16(dx)) || (gamma && !aligned16(gamma))) return LCBI_ERR_BAD_ARG;
  if ((dgamma != nullptr || dbeta != nullptr) &&
      (workspace == nullptr || !aligned16(workspace) || workspace_bytes < layer_norm_bwd_workspace_bytes(rows, C)))
    return LCBI_ERR_WORKSPACE;
  if (x_is_bf16) {
    return dy_is_bf16 ? bwd_typed<__nv_bfloat16, __nv_bfloat16>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, rows, C, stream)
                      : bwd_typed<__nv_bfloat16, float>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, rows, C, stream);
  }
  return dy_is_bf16 ? bwd_typed<float, __nv_bfloat16>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, rows, C, stream)
                    : bwd_typed<float, float>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, rows, C, stream);
}

}  // namespace lcbi
