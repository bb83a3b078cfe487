// Swin window attention for SMALL windows (<= 64 tokens: 7x7, 8x8, 4x4x4, clamped windows), forward and backward.
//
// Same operation as the kernels of window_attn.cu — the reference's forward_part1 gather / attention core / scatter,
// /root/reference/model/models/backbone_swin.py:339-357, :441-485, :591-628 — organised for the regime where a whole
// (window, head) is one 64 x 64 tile and the gather latency and the per-element bias / mask bookkeeping, not the
// MMAs, dominate:
//   * persistent CTAs, one per (head, slice of the window list), 4 warps = the four 16-row tiles of a window;
//   * the rows and the slot -> token map of the NEXT window are gathered with cp.async into a second buffer while the
//     current window is processed;
//   * forward: the head's relative-position bias tile is held in registers for the whole kernel (x log2 e, -inf in
//     the key columns beyond the window), so a logit costs one FFMA;
//   * the shift-mask test runs only for windows that straddle a shift boundary (flag computed with the slot map);
//   * backward: one fused kernel for dq, dk, dv, d(qkv.bias) and d(bias table) (see below).
#include <type_traits>

#include "lcbi_kernels.h"
#include "window_common.cuh"

namespace lcbi {

namespace {

// =================================================================================================
// Forward for small windows (<= 64 tokens: every 2-D config): persistent CTA = (head, slice of the window list),
// 4 warps = the four 16-row query tiles of a window.
//   * the head's relative-position bias tile lives in REGISTERS (32 values per thread, x log2e, -inf in dead key
//     columns) for the whole kernel: no table gathers and no bounds checks per logit;
//   * the q / k / v rows and the slot -> token map of the NEXT window are gathered with cp.async into a second
//     buffer while the current window is processed (the CTA-per-(window, head) kernel exposed one global-memory
//     round trip per window);
//   * the -100 shift-mask term is evaluated only for windows that straddle a region boundary.
// =================================================================================================
// Resident CTAs per SM the forward is compiled for. The kernel runs at ~0.1 IPC per warp (CTA barriers and dependent
// ldmatrix -> mma -> shuffle chains per window), so more windows in flight is what helps: 4 CTAs (113 registers) 115.6 us,
// 6 CTAs (80 registers, 32 bytes of spills) 97.0 us at cfg2 stage 1. The backward does not follow: capped to 128
// registers for a fourth CTA it slows from 213.6 to 251 us (profiles/r02_small_window_tma_variant.log).
constexpr int kFwdSmallCtas = 6;

template <int D>
struct SmallFwdSmem {
  static constexpr int kStride = Tile<D>::kStride;
  static constexpr int kTileBytes = 64 * kStride;
  static constexpr int kTotal = 2 * 3 * kTileBytes + 2 * 2 * 64 * 4 /*tok, reg x2*/ + 16 /*flags*/;
};

template <int D>
__global__ void __launch_bounds__(128, kFwdSmallCtas)
win_attn_fwd_small_kernel(const WinParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  using L = SmallFwdSmem<D>;
  constexpr int kStride = L::kStride;
  constexpr int kChunks = D / 8;
  const WinGeom& g = p.g;
  const int n = g.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y;
  const int gq = lane >> 2, qq = lane & 3;

  uint8_t* tiles = smem;                                       // [2][3][kTileBytes]: q, k, v
  int* s_meta = reinterpret_cast<int*>(tiles + 2 * 3 * L::kTileBytes);   // [2][2][64]: tok, region
  int* s_flags = s_meta + 2 * 2 * 64;                          // [2]: bit 0 pad tokens present, bit 1 mask needed
  const uint32_t tiles_base = static_cast<uint32_t>(__cvta_generic_to_shared(tiles));

  const int units = p.win_count;
  const int row0 = warp * 16;
  const bool tile_live = row0 < n;
  const int i0 = row0 + gq, i1 = i0 + 8;
  const float mask_log2 = -100.0f * kLog2e;

  // bias[h, i, j] * log2e of this thread's fragment positions; dead key columns read -inf, dead rows 0
  float bias[8][4];
  {
    int rt0 = 0, rt1 = 0, dummy;
    if (i0 < n) relpos_terms(g, i0, rt0, dummy);
    if (i1 < n) relpos_terms(g, i1, rt1, dummy);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + qq * 2 + e;
        float v0 = -INFINITY, v1 = -INFINITY;
        if (j < n) {
          int rtj, ctj;
          relpos_terms(g, j, rtj, ctj);
          v0 = i0 < n ? __ldg(p.table + static_cast<int64_t>(rt0 - ctj) * p.H + h) * kLog2e : 0.f;
          v1 = i1 < n ? __ldg(p.table + static_cast<int64_t>(rt1 - ctj) * p.H + h) * kLog2e : 0.f;
        }
        bias[nt][e] = v0;
        bias[nt][2 + e] = v1;
      }
    }
  }

  auto store_meta = [&](int u, int buf) {                     // threads 0-63: one window slot each
    int b, w, tok = -2, reg = -1, tok0, reg0;
    fdivmod(p.win_begin + u, g.d_nW, b, w);
    if (tid < n) slot_lookup(g, w, tid, tok, reg);
    slot_lookup(g, w, 0, tok0, reg0);
    s_meta[buf * 128 + tid] = tok;
    s_meta[buf * 128 + 64 + tid] = reg;
    const unsigned pads = __ballot_sync(0xffffffffu, tok == -1);
    const unsigned mixed = __ballot_sync(0xffffffffu, tid < n && reg != reg0);
    if (lane == 0 && (pads | mixed)) atomicOr(&s_flags[buf], (pads ? 1 : 0) | (mixed ? 2 : 0));
  };
  auto issue_tiles = [&](int u, int buf) {
    int b, w;
    fdivmod(p.win_begin + u, g.d_nW, b, w);
    const int64_t tok_base = static_cast<int64_t>(b) * g.T;
    const int* tk = s_meta + buf * 128;
#pragma unroll 4
    for (int e = tid; e < 3 * 64 * kChunks; e += 128) {
      const int sel = e / (64 * kChunks);            // 0 q, 1 k, 2 v
      const int r = (e / kChunks) & 63, c = e % kChunks;
      const int t = tk[r];
      const __nv_bfloat16* src = p.qkv;              // any valid address when nothing is read (zero fill)
      if (t >= 0) src = p.qkv + ((tok_base + t) * 3 + sel) * p.C + h * D + c * 8;
      const uint32_t dst = tiles_base + (buf * 3 + sel) * L::kTileBytes + r * kStride + c * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(t >= 0 ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (tid < 2) s_flags[tid] = 0;
  __syncthreads();
  if (static_cast<int>(blockIdx.x) < units) {
    if (tid < 64) store_meta(blockIdx.x, 0);
    __syncthreads();
    issue_tiles(blockIdx.x, 0);
  }

  int it = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    const int u_next = u + gridDim.x;
    int b, w;
    fdivmod(p.win_begin + u, g.d_nW, b, w);
    const int* s_tok = s_meta + cur * 128;
    const int* s_reg = s_tok + 64;
    uint8_t* sQ = tiles + (cur * 3) * L::kTileBytes;
    const uint32_t q_base = tiles_base + (cur * 3) * L::kTileBytes, k_base = q_base + L::kTileBytes;
    const uint32_t v_base = k_base + L::kTileBytes;

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int flags = s_flags[cur];
    if ((flags & 1) && p.qkv_bias != nullptr) {
      // pad tokens enter the reference as zeros before the qkv Linear, so their q / k / v rows are the Linear's bias
      for (int e = tid; e < 3 * 64 * kChunks; e += 128) {
        const int sel = e / (64 * kChunks), r = (e / kChunks) & 63, c = e % kChunks;
        if (s_tok[r] != -1) continue;
        const float* bsrc = p.qkv_bias + sel * p.C + h * D + c * 8;
        uint4 val;
        val.x = pack2_bf16(bsrc[0], bsrc[1]);
        val.y = pack2_bf16(bsrc[2], bsrc[3]);
        val.z = pack2_bf16(bsrc[4], bsrc[5]);
        val.w = pack2_bf16(bsrc[6], bsrc[7]);
        *reinterpret_cast<uint4*>(sQ + sel * L::kTileBytes + r * kStride + c * 16) = val;
      }
    }
    if (tid == 0) s_flags[nxt] = 0;                 // (last read one iteration ago; set again after the barrier)
    __syncthreads();                                // this window's tiles are in; the previous window is fully consumed
    if (u_next < units) {
      if (tid < 64) store_meta(u_next, nxt);
      __syncthreads();
      issue_tiles(u_next, nxt);                     // lands while this window is processed
    }

    if (tile_live) {
      uint32_t aq[D / 16][4];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) load_a_frag<kStride>(aq[kk], q_base, row0, kk * 16, lane);
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t b0, b1;
          load_b_frag_nt<kStride>(b0, b1, k_base, nt * 8, kk * 16, lane);
          mma_bf16_16816(s[nt], aq[kk], b0, b1);
        }
      }
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = fmaf(s[nt][0], p.scale_log2, bias[nt][0]);
        s[nt][1] = fmaf(s[nt][1], p.scale_log2, bias[nt][1]);
        s[nt][2] = fmaf(s[nt][2], p.scale_log2, bias[nt][2]);
        s[nt][3] = fmaf(s[nt][3], p.scale_log2, bias[nt][3]);
      }
      if (flags & 2) {                              // the window straddles a shift-mask region boundary
        const int rg0 = s_reg[i0], rg1 = s_reg[i1];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int2 rj = *reinterpret_cast<const int2*>(s_reg + nt * 8 + qq * 2);
          s[nt][0] += rg0 != rj.x ? mask_log2 : 0.f;
          s[nt][1] += rg0 != rj.y ? mask_log2 : 0.f;
          s[nt][2] += rg1 != rj.x ? mask_log2 : 0.f;
          s[nt][3] += rg1 != rj.y ? mask_log2 : 0.f;
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = ex2f(s[nt][0] - m0); s[nt][1] = ex2f(s[nt][1] - m0);
        s[nt][2] = ex2f(s[nt][2] - m1); s[nt][3] = ex2f(s[nt][3] - m1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      float oacc[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {              // 16 keys per step
        uint32_t ap[4];
        ap[0] = pack2_bf16(s[2 * kb][0], s[2 * kb][1]);
        ap[1] = pack2_bf16(s[2 * kb][2], s[2 * kb][3]);
        ap[2] = pack2_bf16(s[2 * kb + 1][0], s[2 * kb + 1][1]);
        ap[3] = pack2_bf16(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) {
          uint32_t b0, b1;
          load_b_frag_t<kStride>(b0, b1, v_base, kb * 16, nd * 8, lane);
          mma_bf16_16816(oacc[nd], ap, b0, b1);
        }
      }
      const float inv0 = 1.f / l0, inv1 = 1.f / l1;
      // stage the 16 x D output tile in this warp's (already consumed) Q rows, then 16-byte scatter stores
      __syncwarp();
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd) {
        *reinterpret_cast<uint32_t*>(sQ + i0 * kStride + (nd * 8 + qq * 2) * 2) = pack2_bf16(oacc[nd][0] * inv0, oacc[nd][1] * inv0);
        *reinterpret_cast<uint32_t*>(sQ + i1 * kStride + (nd * 8 + qq * 2) * 2) = pack2_bf16(oacc[nd][2] * inv1, oacc[nd][3] * inv1);
      }
      const int t0 = s_tok[i0], t1 = s_tok[i1];
      if (qq == 0) {
        if (t0 >= 0) p.lse2[(static_cast<int64_t>(b) * g.T + t0) * p.H + h] = m0 + log2f(l0);
        if (t1 >= 0) p.lse2[(static_cast<int64_t>(b) * g.T + t1) * p.H + h] = m1 + log2f(l1);
      }
      __syncwarp();
      for (int e = lane; e < 16 * kChunks; e += 32) {
        const int r = row0 + e / kChunks, c = e % kChunks;
        const int t = s_tok[r];
        if (t >= 0)
          *reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * g.T + t) * p.C + h * D + c * 8) =
              *reinterpret_cast<const uint4*>(sQ + r * kStride + c * 16);
      }
    }
  }
}

template <int D>
int launch_small(const WinParams& p, cudaStream_t stream) {
  using L = SmallFwdSmem<D>;
  static unsigned long long configured = 0;   // one bit per device ordinal (one word per head_dim instantiation)
  if (first_launch_on_current_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(win_attn_fwd_small_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) {
      configured = 0;
      return set_cuda_error(e);
    }
  }
  const int units = p.win_count;
  int grid_x = (2 * 148 * kFwdSmallCtas + p.H - 1) / p.H;     // ~2 waves of persistent CTAs per head slice
  if (grid_x > units) grid_x = units;
  if (grid_x < 1) grid_x = 1;
  win_attn_fwd_small_kernel<D><<<dim3(grid_x, p.H), 128, L::kTotal, stream>>>(p);
  return set_cuda_error(cudaGetLastError());
}


// =================================================================================================
// Fused backward for small windows: dq, dk, dv, d(qkv.bias through pad tokens) and d(bias table) in ONE kernel.
//   persistent CTA = (head, slice of the window list), 4 warps.
//   phase 1 (warp = 16-row QUERY tile): S, P = exp2(S*c + bias + mask - lse2), dP = dO V^T, dS = P o (dP - D);
//            dQ = dS K written out; dBias += dS kept in registers ACROSS windows; P, dS -> bf16 smem tiles
//   phase 2 (warp = 16-row KEY tile):  dV = P^T dO, dK = dS^T Q with the A fragments read transposed from smem
//   once per CTA: dBias registers -> atomicAdd into d(relative_position_bias_table).
// No recomputation (5 GEMMs, one exp per logit) and one atomic per (i, j) per CTA instead of per window.
// =================================================================================================
template <int STRIDE_BYTES>
__device__ __forceinline__ void load_a_frag_trans(uint32_t (&a)[4], uint32_t tile_base, int k0, int m0, int lane) {
  // A[m][k] = X[k0 + k][m0 + m] for X stored row-major with k as the row index
  const uint32_t addr = tile_base + (k0 + (lane & 7) + ((lane >> 4) & 1) * 8) * STRIDE_BYTES + (m0 + ((lane >> 3) & 1) * 8) * 2;
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

constexpr int kPStride = 144;   // bytes per row of the bf16 P / dS tiles (64 columns + 16 B pad)

template <int D>
struct SmallBwdSmem {
  static constexpr int kStride = Tile<D>::kStride;
  static constexpr int kTileBytes = 64 * kStride;
  static constexpr int kPBytes = 64 * kPStride;
  // q / k / v / dO tiles and the per-window metadata (token, region, lse, dsum) are double-buffered: the next window
  // is gathered with cp.async while the current one is processed
  static constexpr int kTotal = 2 * 4 * kTileBytes + 2 * kPBytes + 2 * 4 * 64 * 4 /*meta x2*/ + 2 * 64 * 4 /*rt, ct*/ + 16;
};

template <int D>
__global__ void __launch_bounds__(128, 3)
win_attn_bwd_small_kernel(const WinParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  using L = SmallBwdSmem<D>;
  constexpr int kStride = L::kStride;
  constexpr int kChunks = D / 8;
  const WinGeom& g = p.g;
  const int n = g.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.y;

  uint8_t* tiles = smem;                               // [2][4][kTileBytes]: q, k, v, dO of the current / next window
  uint8_t* sP = tiles + 2 * 4 * L::kTileBytes;
  uint8_t* sDS = sP + L::kPBytes;
  float* s_meta = reinterpret_cast<float*>(sDS + L::kPBytes);   // [2][4][64]: lse, dsum, tok, region
  int* s_rt = reinterpret_cast<int*>(s_meta + 2 * 4 * 64);
  int* s_ct = s_rt + 64;
  int* s_pad = s_ct + 64;                              // [2] (+2 unused): does the window hold pad tokens?
  float* tab = reinterpret_cast<float*>(s_pad + 4);    // [tab_rows] bias table column of this head, x log2e

  for (int t = tid; t < g.tab_rows; t += 128) tab[t] = p.table[static_cast<int64_t>(t) * p.H + h] * kLog2e;
  if (tid < 64) {
    int rt = g.tab_rows - 1, ct = 0;
    if (tid < n) relpos_terms(g, tid, rt, ct);
    s_rt[tid] = rt;
    s_ct[tid] = ct;
  }

  const uint32_t tiles_base = static_cast<uint32_t>(__cvta_generic_to_shared(tiles));
  const uint32_t p_base = static_cast<uint32_t>(__cvta_generic_to_shared(sP));
  const uint32_t ds_base = static_cast<uint32_t>(__cvta_generic_to_shared(sDS));
  const int gq = lane >> 2, qq = lane & 3;
  const float mask_log2 = -100.0f * kLog2e;
  const float scale = p.scale_log2 / kLog2e;
  const int units = p.win_count;
  const int row0 = warp * 16;                       // this warp's query tile (phase 1) and key tile (phase 2)
  const bool tile_live = row0 < n;

  float dbias[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) dbias[i][0] = dbias[i][1] = dbias[i][2] = dbias[i][3] = 0.f;
  float pad_dk[D / 8][2], pad_dv[D / 8][2];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) pad_dk[i][0] = pad_dk[i][1] = pad_dv[i][0] = pad_dv[i][1] = 0.f;

  // ---- per-window gather pipeline -----------------------------------------------------------------------------
  // window_meta: slot -> token / region (closed form) and the token's lse / dsum, for one window, by threads 0-63;
  // issue_tiles: cp.async of the window's q / k / v / dO rows (zero-fill for pad and dead rows).
  struct MetaRegs { int tok, reg; float lse, dsum; bool mixed = false; };
  auto window_meta = [&](int u) {
    MetaRegs m{-2, -1, INFINITY, 0.f};              // pad / dead query rows: P = exp2(. - inf) = 0
    int b, w;
    fdivmod(p.win_begin + u, g.d_nW, b, w);
    if (tid < n) slot_lookup(g, w, tid, m.tok, m.reg);
    {
      int tok0, reg0;
      slot_lookup(g, w, 0, tok0, reg0);
      m.mixed = tid < n && m.reg != reg0;           // the window straddles a shift-mask region boundary
    }
    if (m.tok >= 0) {
      const int64_t idx = (static_cast<int64_t>(b) * g.T + m.tok) * p.H + h;
      m.lse = __ldg(p.lse2 + idx);
      m.dsum = __ldg(p.dsum + idx);
    }
    return m;
  };
  auto store_meta = [&](const MetaRegs& m, int buf) {     // threads 0-63
    float* mb = s_meta + buf * 4 * 64;
    mb[tid] = m.lse;
    mb[64 + tid] = m.dsum;
    reinterpret_cast<int*>(mb)[128 + tid] = m.tok;
    reinterpret_cast<int*>(mb)[192 + tid] = m.reg;
    const unsigned pads = __ballot_sync(0xffffffffu, m.tok == -1);
    const unsigned mixed = __ballot_sync(0xffffffffu, m.mixed);
    if (lane == 0 && (pads | mixed)) atomicOr(&s_pad[buf], (pads ? 1 : 0) | (mixed ? 2 : 0));
  };
  auto issue_tiles = [&](int u, int buf) {
    int b, w;
    fdivmod(p.win_begin + u, g.d_nW, b, w);
    const int64_t tok_base = static_cast<int64_t>(b) * g.T;
    const int* tk = reinterpret_cast<const int*>(s_meta + buf * 4 * 64) + 128;
#pragma unroll 4
    for (int e = tid; e < 4 * 64 * kChunks; e += 128) {
      const int sel = e / (64 * kChunks);            // 0 q, 1 k, 2 v, 3 dO
      const int r = (e / kChunks) & 63, c = e % kChunks;
      const int t = tk[r];
      const __nv_bfloat16* src = p.qkv;              // any valid address when nothing is read (zero fill)
      if (t >= 0)
        src = sel < 3 ? p.qkv + ((tok_base + t) * 3 + sel) * p.C + h * D + c * 8
                      : p.d_out + (tok_base + t) * p.C + h * D + c * 8;
      const uint32_t dst = tiles_base + (buf * 4 + sel) * L::kTileBytes + r * kStride + c * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(t >= 0 ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (tid < 4) s_pad[tid] = 0;
  __syncthreads();                                  // tab / terms / flags written
  if (static_cast<int>(blockIdx.x) < units) {
    if (tid < 64) store_meta(window_meta(blockIdx.x), 0);
    __syncthreads();
    issue_tiles(blockIdx.x, 0);
  }

  int it = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    const int u_next = u + gridDim.x;
    int b, w;
    fdivmod(p.win_begin + u, g.d_nW, b, w);
    const float* s_lse = s_meta + cur * 4 * 64;
    const float* s_dsum = s_lse + 64;
    const int* s_tok = reinterpret_cast<const int*>(s_lse) + 128;
    const int* s_reg = s_tok + 64;
    uint8_t* sQ = tiles + (cur * 4 + 0) * L::kTileBytes;
    const uint32_t q_base = tiles_base + (cur * 4 + 0) * L::kTileBytes, k_base = q_base + L::kTileBytes;
    const uint32_t v_base = k_base + L::kTileBytes, do_base = v_base + L::kTileBytes;

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int flags = s_pad[cur];
    if ((flags & 1) && p.qkv_bias != nullptr) {
      // pad tokens enter the reference as zeros before the qkv Linear, so their q / k / v rows are the Linear's bias
      for (int e = tid; e < 3 * 64 * kChunks; e += 128) {
        const int sel = e / (64 * kChunks), r = (e / kChunks) & 63, c = e % kChunks;
        if (s_tok[r] != -1) continue;
        const float* bsrc = p.qkv_bias + sel * p.C + h * D + c * 8;
        uint4 val;
        val.x = pack2_bf16(bsrc[0], bsrc[1]);
        val.y = pack2_bf16(bsrc[2], bsrc[3]);
        val.z = pack2_bf16(bsrc[4], bsrc[5]);
        val.w = pack2_bf16(bsrc[6], bsrc[7]);
        *reinterpret_cast<uint4*>(sQ + sel * L::kTileBytes + r * kStride + c * 16) = val;
      }
    }
    if (tid == 0) s_pad[nxt] = 0;                   // (last read one iteration ago; set again after the barrier)
    __syncthreads();                                // this window's tiles are in; the previous window is fully consumed
    // the next window's metadata: the two dependent global loads stay in flight during phase 1
    MetaRegs next_meta{-2, -1, INFINITY, 0.f};
    if (u_next < units && tid < 64) next_meta = window_meta(u_next);

    // ------------------------------------------------------------ phase 1: this warp's query tile
    if (tile_live) {
      uint32_t aq[D / 16][4], ado[D / 16][4];
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        load_a_frag<kStride>(aq[kk], q_base, row0, kk * 16, lane);
        load_a_frag<kStride>(ado[kk], do_base, row0, kk * 16, lane);
      }
      const int i0 = row0 + gq, i1 = i0 + 8;
      const float lse0 = s_lse[i0], lse1 = s_lse[i1], ds0 = s_dsum[i0], ds1 = s_dsum[i1];
      const int rt0 = s_rt[i0], rt1 = s_rt[i1], rg0 = s_reg[i0], rg1 = s_reg[i1];
      float dq[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {        // 32 keys per step
        float s[4][4], dp[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
          dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            uint32_t b0, b1;
            load_b_frag_nt<kStride>(b0, b1, k_base, half * 32 + nt * 8, kk * 16, lane);
            mma_bf16_16816(s[nt], aq[kk], b0, b1);
            load_b_frag_nt<kStride>(b0, b1, v_base, half * 32 + nt * 8, kk * 16, lane);
            mma_bf16_16816(dp[nt], ado[kk], b0, b1);
          }
        }
        auto probs = [&](auto masked_c) {
          constexpr bool kMasked = decltype(masked_c)::value;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int j = half * 32 + nt * 8 + qq * 2;
            const int2 ct = *reinterpret_cast<const int2*>(s_ct + j);     // dead columns: col_term 0 (valid gather)
            int2 rj = make_int2(0, 0);
            if (kMasked) rj = *reinterpret_cast<const int2*>(s_reg + j);
            float pv[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cte = e ? ct.y : ct.x;
              float l0 = fmaf(s[nt][e], p.scale_log2, tab[rt0 - cte]) - lse0;
              float l1 = fmaf(s[nt][2 + e], p.scale_log2, tab[rt1 - cte]) - lse1;
              if (kMasked) {
                const int rje = e ? rj.y : rj.x;
                l0 += rg0 != rje ? mask_log2 : 0.f;
                l1 += rg1 != rje ? mask_log2 : 0.f;
              }
              const bool live = j + e < n;          // dead key columns must not reach dBias / dK / dV
              const float p0 = live ? ex2f(l0) : 0.f, p1 = live ? ex2f(l1) : 0.f;
              pv[e] = p0;
              pv[2 + e] = p1;
              dp[nt][e] = p0 * (dp[nt][e] - ds0);
              dp[nt][2 + e] = p1 * (dp[nt][2 + e] - ds1);
              dbias[half * 4 + nt][e] += dp[nt][e];
              dbias[half * 4 + nt][2 + e] += dp[nt][2 + e];
            }
            const int col = j * 2;
            *reinterpret_cast<uint32_t*>(sP + i0 * kPStride + col) = pack2_bf16(pv[0], pv[1]);
            *reinterpret_cast<uint32_t*>(sP + i1 * kPStride + col) = pack2_bf16(pv[2], pv[3]);
            *reinterpret_cast<uint32_t*>(sDS + i0 * kPStride + col) = pack2_bf16(dp[nt][0], dp[nt][1]);
            *reinterpret_cast<uint32_t*>(sDS + i1 * kPStride + col) = pack2_bf16(dp[nt][2], dp[nt][3]);
          }
        };
        if (flags & 2) probs(std::true_type{}); else probs(std::false_type{});
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          uint32_t ads[4];
          ads[0] = pack2_bf16(dp[2 * kb][0], dp[2 * kb][1]);
          ads[1] = pack2_bf16(dp[2 * kb][2], dp[2 * kb][3]);
          ads[2] = pack2_bf16(dp[2 * kb + 1][0], dp[2 * kb + 1][1]);
          ads[3] = pack2_bf16(dp[2 * kb + 1][2], dp[2 * kb + 1][3]);
#pragma unroll
          for (int nd = 0; nd < D / 8; ++nd) {
            uint32_t b0, b1;
            load_b_frag_t<kStride>(b0, b1, k_base, half * 32 + kb * 16, nd * 8, lane);
            mma_bf16_16816(dq[nd], ads, b0, b1);
          }
        }
      }
      const int t0 = s_tok[i0], t1 = s_tok[i1];
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd) {
        const int col = h * D + nd * 8 + qq * 2;
        if (t0 >= 0)
          *reinterpret_cast<uint32_t*>(p.dqkv + (static_cast<int64_t>(b) * g.T + t0) * 3 * p.C + col) =
              pack2_bf16(dq[nd][0] * scale, dq[nd][1] * scale);
        if (t1 >= 0)
          *reinterpret_cast<uint32_t*>(p.dqkv + (static_cast<int64_t>(b) * g.T + t1) * 3 * p.C + col) =
              pack2_bf16(dq[nd][2] * scale, dq[nd][3] * scale);
      }
    } else {
      // dead query tile (all slots beyond the window): its P / dS rows must read as zero in phase 2
      for (int e = lane; e < 16 * (kPStride / 4); e += 32) {
        reinterpret_cast<uint32_t*>(sP + row0 * kPStride)[e] = 0u;
        reinterpret_cast<uint32_t*>(sDS + row0 * kPStride)[e] = 0u;
      }
    }
    if (u_next < units && tid < 64) store_meta(next_meta, nxt);
    __syncthreads();                                // P / dS tiles complete; next window's metadata visible
    if (u_next < units) issue_tiles(u_next, nxt);   // lands while phase 2 runs and the loop turns around

    // ------------------------------------------------------------ phase 2: this warp's key tile
    if (tile_live) {
      float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
      }
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {              // 16 queries per k-step
        uint32_t apt[4], adst[4];
        load_a_frag_trans<kPStride>(apt, p_base, kb * 16, row0, lane);
        load_a_frag_trans<kPStride>(adst, ds_base, kb * 16, row0, lane);
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) {
          uint32_t b0, b1;
          load_b_frag_t<kStride>(b0, b1, do_base, kb * 16, nd * 8, lane);
          mma_bf16_16816(dv[nd], apt, b0, b1);
          load_b_frag_t<kStride>(b0, b1, q_base, kb * 16, nd * 8, lane);
          mma_bf16_16816(dk[nd], adst, b0, b1);
        }
      }
      const int j0 = row0 + gq, j1 = j0 + 8;
      const int t0 = s_tok[j0], t1 = s_tok[j1];
#pragma unroll
      for (int nd = 0; nd < D / 8; ++nd) {
        const int col = h * D + nd * 8 + qq * 2;
        if (t0 >= 0) {
          const int64_t base = (static_cast<int64_t>(b) * g.T + t0) * 3 * p.C;
          *reinterpret_cast<uint32_t*>(p.dqkv + base + p.C + col) = pack2_bf16(dk[nd][0] * scale, dk[nd][1] * scale);
          *reinterpret_cast<uint32_t*>(p.dqkv + base + 2 * p.C + col) = pack2_bf16(dv[nd][0], dv[nd][1]);
        } else if (t0 == -1) {
          pad_dk[nd][0] += dk[nd][0] * scale; pad_dk[nd][1] += dk[nd][1] * scale;
          pad_dv[nd][0] += dv[nd][0]; pad_dv[nd][1] += dv[nd][1];
        }
        if (t1 >= 0) {
          const int64_t base = (static_cast<int64_t>(b) * g.T + t1) * 3 * p.C;
          *reinterpret_cast<uint32_t*>(p.dqkv + base + p.C + col) = pack2_bf16(dk[nd][2] * scale, dk[nd][3] * scale);
          *reinterpret_cast<uint32_t*>(p.dqkv + base + 2 * p.C + col) = pack2_bf16(dv[nd][2], dv[nd][3]);
        } else if (t1 == -1) {
          pad_dk[nd][0] += dk[nd][2] * scale; pad_dk[nd][1] += dk[nd][3] * scale;
          pad_dv[nd][0] += dv[nd][2]; pad_dv[nd][1] += dv[nd][3];
        }
      }
    }
  }

  // ---- once per CTA: bias-table gradient and the pad-token share of d(qkv.bias).
  // Every CTA of a head adds into the same (2w-1)^2 table entries: going to global memory per (i, j) pair (2401 adds
  // per CTA onto 169 addresses at 7x7) made the L2 serialise ~4000 same-address atomics per entry - a quarter of the
  // kernel's stall samples at cfg2. The pairs are folded in shared memory first (the bias table and the P tile are dead
  // by now), so that a CTA issues one global add per table entry and per pad-bias column.
  __syncthreads();                                   // all warps are past their last read of tab / sP
  float* s_dt = tab;                                 // [tab_rows]
  float* s_dpad = reinterpret_cast<float*>(sP);      // [2 * D]
  for (int t = tid; t < g.tab_rows; t += 128) s_dt[t] = 0.f;
  if (tid < 2 * D) s_dpad[tid] = 0.f;
  __syncthreads();
  if (p.dtable != nullptr && tile_live) {
    const int i0 = row0 + gq, i1 = i0 + 8;
    const int rt0 = s_rt[i0 < 64 ? i0 : 63], rt1 = s_rt[i1 < 64 ? i1 : 63];
#pragma unroll
    for (int t8 = 0; t8 < 8; ++t8) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = t8 * 8 + qq * 2 + e;
        if (j >= n) continue;
        const int ct = s_ct[j];
        if (i0 < n) atomicAdd(s_dt + (rt0 - ct), dbias[t8][e]);
        if (i1 < n) atomicAdd(s_dt + (rt1 - ct), dbias[t8][2 + e]);
      }
    }
  }
  const bool has_pad_bias = p.dbias_pad != nullptr && g.n * g.nW != g.T;
  if (has_pad_bias) {
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = pad_dk[nd][e], c = pad_dv[nd][e];
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, off);
          c += __shfl_xor_sync(0xffffffffu, c, off);
        }
        if (gq == 0) {
          const int col = nd * 8 + qq * 2 + e;
          if (a != 0.f) atomicAdd(s_dpad + col, a);
          if (c != 0.f) atomicAdd(s_dpad + D + col, c);
        }
      }
    }
  }
  __syncthreads();
  if (p.dtable != nullptr)
    for (int t = tid; t < g.tab_rows; t += 128) {
      const float v = s_dt[t];
      if (v != 0.f) atomicAdd(p.dtable + static_cast<int64_t>(t) * p.H + h, v);
    }
  if (has_pad_bias && tid < 2 * D) {
    const float v = s_dpad[tid];
    if (v != 0.f) atomicAdd(p.dbias_pad + (tid < D ? 1 : 2) * p.C + h * D + (tid < D ? tid : tid - D), v);
  }
}

template <int D>
int launch_small_bwd(const WinParams& p, cudaStream_t stream) {
  using L = SmallBwdSmem<D>;
  const size_t smem = L::kTotal + static_cast<size_t>(round_up(p.g.tab_rows, 4)) * 4;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(win_attn_bwd_small_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_bytes = smem;
  }
  const int units = p.win_count;
  int ctas_per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (ctas_per_sm > 3) ctas_per_sm = 3;             // register-limited (launch bounds)
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  int grid_x = (2 * 148 * ctas_per_sm + p.H - 1) / p.H;   // ~2 waves of persistent CTAs per head slice
  if (grid_x > units) grid_x = units;
  if (grid_x < 1) grid_x = 1;
  win_attn_bwd_small_kernel<D><<<dim3(grid_x, p.H), 128, smem, stream>>>(p);
  return set_cuda_error(cudaGetLastError());
}

}  // namespace

int win_attn_fwd_small_launch(const WinParams& p, int head_dim, cudaStream_t stream) {
  return head_dim == 16 ? launch_small<16>(p, stream) : launch_small<32>(p, stream);
}

int win_attn_bwd_small_launch(const WinParams& p, int head_dim, cudaStream_t stream) {
  return head_dim == 16 ? launch_small_bwd<16>(p, stream) : launch_small_bwd<32>(p, stream);
}

}  // namespace lcbi
