// Swin (shifted-)window attention on tcgen05 / TMEM / TMA for 3-D windows of 128..512 tokens (7x7x7 = 343, 8x8x8 = 512),
// head_dim 16 or 32: forward.
//
// Same contract as window_attn.cu (the generic mma.sync kernels, which keep serving clamped windows, 2-D windows and
// odd geometries): replaces, without any of their copies, SwinTransformerBlock.forward_part1's
// pad -> roll -> window_partition -> attention -> window_reverse -> roll -> crop chain
// (/root/reference/model/models/backbone_swin.py:435-487) and WindowAttention.forward's attention core (:339-357) with
// the relative-position bias (:343-347) and the shift mask of compute_mask (:591-628, never materialised).
//
// Work item = (window, head); persistent CTAs take contiguous ranges of the head-major item list. Per item:
//   * gather (warp 5): the window's q / k / v rows come from the (B, D, H, W, 3C) qkv tensor by TMA. A window that does
//     not cross the cyclic wrap of the shifted frame is a BOX of that tensor: one rank-5 box load for q (all planes) and
//     one per window plane for k and v (UTMALDG; coordinates past the grid - the far-end pad - are zero-filled). The
//     windows that wrap (the last one along each shifted axis) are fetched with TMA tile::gather4 - four arbitrary token
//     rows per instruction, indices from the closed-form window map (UTMALDG.GATHER4). Pad-token rows are then
//     overwritten with qkv.bias (the reference pads after norm1, so a pad token's q/k/v is exactly the bias).
//   * keys / values live PLANE-PADDED in shared memory: plane a' of the window (PH*PW tokens) starts at row 64*a', so a
//     64-key MMA chunk is one plane and the relative-position index of (query i, key (a',b',c')) is
//     row_term(i) - a'*stride0 - (b'*stride1 + c'): a per-chunk base plus a COMPILE-TIME offset - the bias gather is one
//     LDS with an immediate, no index arithmetic per logit (the generic kernels spent ~40 % of their issue slots there).
//   * S = Q K_plane^T on tcgen05 (M128 x N64 x K16/32, operands K-major with 32/64-byte rows: SWIZZLE_32B / _64B),
//     double-buffered in TMEM; softmax warps 0-3 (thread = query row = TMEM lane): logit = s*scale*log2e + bias
//     (+ -100*log2e from a 4-entry per-thread class table when the window straddles a shift-mask region boundary),
//     online softmax with a lazily moved maximum, P (bf16) back into TMEM over S; O += P V_plane (TS MMA, V MN-major).
//   * epilogue: O / l -> bf16 -> the token's row of `out`, log2-domain lse per real token.
// Layout conventions were validated on a B200 by tools/probe_small_swizzle.cu (profiles/r02_probe_small_swizzle.log).
#include <cstdlib>

#include "lcbi_kernels.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"
#include "window_common.cuh"

namespace lcbi {

namespace {

constexpr uint32_t kLayoutSW64 = 4, kLayoutSW32 = 6;
constexpr int kChunk = 64;                 // keys per MMA chunk = one (padded) window plane
constexpr int kTileM = 128;
constexpr float kMaskLog2 = -100.0f * kLog2e;
constexpr float kRescale = 8.0f;           // log2 units the running maximum may lag behind
constexpr int kFwdThreads = 192;
constexpr uint32_t kTmemColsFwd = 256, kTmemOFwd = 128;

__device__ __forceinline__ void tma_gather4(void* dst, const void* tmap, uint64_t* bar, int col, int r0, int r1, int r2,
                                            int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// byte offset of 16-byte chunk c of row r in a dense tile of ROW_BYTES-byte rows under the matching TMA / UMMA swizzle
// (address bits [4, 4+B) ^= bits [7, 7+B); B = 1 for 32-byte rows, 2 for 64-byte rows)
template <int ROW_BYTES>
__device__ __forceinline__ uint32_t swz_off(uint32_t r, uint32_t c) {
  const uint32_t lin = r * ROW_BYTES + c * 16;
  return lin ^ (((lin >> 7) & (ROW_BYTES / 16 - 1)) << 4);
}

struct TcFwdParams {
  WinParams w;
  int n_items;            // win_count * H, head-major: item = head * win_count + (window - win_begin)
  int items_per_cta;
  int mt;                 // 128-row query tiles per window
  int npl;                // planes per window (= window depth) = key chunks
  uint32_t q_bytes, kv_bytes, tab_bytes, buf_bytes;   // per operand buffer (q, k, v, bias table)
};

struct FwdBars {
  uint64_t full[2], ready[2], free_[2];
  uint64_t s_full[2], p_full[2], pv_done, o_full;
  uint32_t tmem_base;
  int buf_head[2];
};

// per-window facts every role derives the same way
struct WinInfo {
  int b, h, w;            // batch, head, window within the image
  int start[3];           // source coordinate of the window's first token along each axis (un-wrapped start < pg)
  bool box;               // no axis wraps: the window is one box of the token grid
  bool has_pad;           // some token lies in the far-end padding
  bool str[3];            // axis straddles a shift-mask region boundary
};

__device__ __forceinline__ WinInfo window_info(const WinParams& p, int item) {
  WinInfo wi;
  const int h = item / p.win_count;
  const int wg = p.win_begin + (item - h * p.win_count);
  wi.h = h;
  wi.b = wg / p.g.nW;
  wi.w = wg - wi.b * p.g.nW;
  int q, w2, w1, w0;
  fdivmod(wi.w, p.g.d_nwin[2], q, w2);
  fdivmod(q, p.g.d_nwin[1], w0, w1);
  const int wc[3] = {w0, w1, w2};
  wi.box = true;
  wi.has_pad = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int s = wc[k] * p.g.win[k] + p.g.shift[k];
    if (s >= p.g.pg[k]) s -= p.g.pg[k];
    wi.start[k] = s;
    if (s + p.g.win[k] > p.g.pg[k]) wi.box = false;
    if (p.g.pg[k] > p.g.grid[k] && s + p.g.win[k] - 1 >= p.g.grid[k]) wi.has_pad = true;
    wi.str[k] = p.g.shift[k] > 0 && wc[k] == p.g.nwin[k] - 1;
  }
  return wi;
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int D, int PH, int PW>
__global__ void __launch_bounds__(kFwdThreads, 2)
win_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_dense, const __grid_constant__ CUtensorMap tm_plane,
                       const __grid_constant__ CUtensorMap tm_rows, const TcFwdParams p) {
  constexpr int RB = D * 2;                                   // bytes per operand row
  constexpr uint32_t kLayout = D == 16 ? kLayoutSW32 : kLayoutSW64;
  constexpr uint32_t kSbo = 8 * RB;
  constexpr int PLANE = PH * PW;
  constexpr int ST1 = 2 * PW - 1, ST0 = ST1 * (2 * PH - 1);
  static_assert(PLANE <= kChunk, "a window plane must fit one 64-key chunk");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // The bias tables are STATIC shared arrays: the compiler then knows the address space and the per-logit gather is an
  // LDS with an immediate offset (through the re-aligned dynamic pointer it would be a generic LD).
  __shared__ float s_tab[2][(2 * PH - 1) * (2 * PH - 1) * (2 * PW - 1)];   // window depth <= PH (win_attn_tc_applicable)
  // dynamic: [buffer 0: q | k | v] [buffer 1: ...] [barriers]
  FwdBars& bars = *reinterpret_cast<FwdBars*>(smem + 2 * p.buf_bytes);
  auto q_of = [&](int buf) { return smem + buf * p.buf_bytes; };
  auto k_of = [&](int buf) { return smem + buf * p.buf_bytes + p.q_bytes; };
  auto v_of = [&](int buf) { return smem + buf * p.buf_bytes + p.q_bytes + p.kv_bytes; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const WinGeom& g = p.w.g;
  const int n = g.n;
  const int item_begin = blockIdx.x * p.items_per_cta;
  const int item_end = min(item_begin + p.items_per_cta, p.n_items);

  // operand buffers start as zeros: rows the loads never write (tile padding, the 64 - PLANE dummy keys of a chunk) must
  // read as zero
  for (uint32_t i = tid; i < 2 * p.buf_bytes / 16; i += kFwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.full[b], 1);
      mbar_init(&bars.ready[b], 1);
      mbar_init(&bars.free_[b], 5);        // the issuer's commit + one arrive per softmax warp
      mbar_init(&bars.s_full[b], 1);
      mbar_init(&bars.p_full[b], 4);
      bars.buf_head[b] = -1;
    }
    mbar_init(&bars.pv_done, 1);
    mbar_init(&bars.o_full, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 4) {
    tmem_alloc(&bars.tmem_base, kTmemColsFwd);
    tmem_relinquish();
  }
  if (warp == 5 && elect_one()) {
    tma_prefetch_desc(&tm_dense);
    tma_prefetch_desc(&tm_plane);
    tma_prefetch_desc(&tm_rows);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 5) {
    // ------------------------------------------------------------------ gather (all 32 lanes)
    constexpr int GP = (PLANE + 3) / 4;                // gather4 groups per plane
    int it = 0;
    for (int item = item_begin; item < item_end; ++item, ++it) {
      const int buf = it & 1, use = it >> 1;
      if (it >= 2) mbar_wait_sleep(&bars.free_[buf], (use - 1) & 1);
      const WinInfo wi = window_info(p.w, item);
      uint8_t *qd = q_of(buf), *kd = k_of(buf), *vd = v_of(buf);
      if (bars.buf_head[buf] != wi.h) {                // this buffer's bias table: the head's column, times log2(e)
        float* tab = s_tab[buf];
        for (int t = lane; t < g.tab_rows; t += 32) tab[t] = p.w.table[static_cast<size_t>(t) * p.w.H + wi.h] * kLog2e;
        __syncwarp();
        if (lane == 0) bars.buf_head[buf] = wi.h;
      }
      const int col_q = wi.h * D, col_k = p.w.C + wi.h * D, col_v = 2 * p.w.C + wi.h * D;
      const int qg = (n + 3) / 4;
      if (wi.box) {
        if (lane == 0) mbar_expect_tx(&bars.full[buf], static_cast<uint32_t>(n + 2 * p.npl * PLANE) * RB);
        __syncwarp();
        if (lane == 0) tma_load_5d(qd, &tm_dense, &bars.full[buf], col_q, wi.start[2], wi.start[1], wi.start[0], wi.b);
        if (lane < p.npl) {
          tma_load_5d(kd + lane * kChunk * RB, &tm_plane, &bars.full[buf], col_k, wi.start[2], wi.start[1],
                      wi.start[0] + lane, wi.b);
          tma_load_5d(vd + lane * kChunk * RB, &tm_plane, &bars.full[buf], col_v, wi.start[2], wi.start[1],
                      wi.start[0] + lane, wi.b);
        }
      } else {
        if (lane == 0) mbar_expect_tx(&bars.full[buf], static_cast<uint32_t>(qg + 2 * p.npl * GP) * 4 * RB);
        __syncwarp();
        const int oob = p.w.B * g.T;                   // a row index past the tensor: zero fill
        const int row0 = wi.b * g.T;
        for (int gq = lane; gq < qg; gq += 32) {       // q: dense slot order
          int r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int s = gq * 4 + u;
            int tok = -1, reg;
            if (s < n) slot_lookup(g, wi.w, s, tok, reg);
            r[u] = tok >= 0 ? row0 + tok : oob;
          }
          tma_gather4(qd + gq * 4 * RB, &tm_rows, &bars.full[buf], col_q, r[0], r[1], r[2], r[3]);
        }
        for (int idx = lane; idx < p.npl * GP; idx += 32) {   // k, v: plane-padded slot order
          const int a = idx / GP, gg = idx - a * GP;
          int r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = gg * 4 + u;
            int tok = -1, reg;
            if (e < PLANE) slot_lookup(g, wi.w, a * PLANE + e, tok, reg);
            r[u] = tok >= 0 ? row0 + tok : oob;
          }
          const uint32_t off = static_cast<uint32_t>(a * kChunk + gg * 4) * RB;
          tma_gather4(kd + off, &tm_rows, &bars.full[buf], col_k, r[0], r[1], r[2], r[3]);
          tma_gather4(vd + off, &tm_rows, &bars.full[buf], col_v, r[0], r[1], r[2], r[3]);
        }
      }
      mbar_wait_sleep(&bars.full[buf], use & 1);
      if (wi.has_pad && p.w.qkv_bias != nullptr) {
        // pad tokens: q / k / v = the qkv Linear's bias (reference backbone_swin.py:441-455 pads the normed tokens
        // with zeros BEFORE the Linear)
        for (int s = lane; s < n; s += 32) {
          int tok, reg;
          slot_lookup(g, wi.w, s, tok, reg);
          if (tok >= 0) continue;
          const int a = s / PLANE, e = s - a * PLANE;
          const uint32_t kr = a * kChunk + e;
#pragma unroll
          for (int c = 0; c < D / 8; ++c) {
            uint4 val[3];
#pragma unroll
            for (int sel = 0; sel < 3; ++sel) {
              const float* bsrc = p.w.qkv_bias + sel * p.w.C + wi.h * D + c * 8;
              val[sel].x = pack2_bf16(bsrc[0], bsrc[1]);
              val[sel].y = pack2_bf16(bsrc[2], bsrc[3]);
              val[sel].z = pack2_bf16(bsrc[4], bsrc[5]);
              val[sel].w = pack2_bf16(bsrc[6], bsrc[7]);
            }
            *reinterpret_cast<uint4*>(qd + swz_off<RB>(s, c)) = val[0];
            *reinterpret_cast<uint4*>(kd + swz_off<RB>(kr, c)) = val[1];
            *reinterpret_cast<uint4*>(vd + swz_off<RB>(kr, c)) = val[2];
          }
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ready[buf]);
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kTileM, kChunk, 0, 0);   // Q K^T: both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(kTileM, D, 0, 1);        // P V: V MN-major
      int g0 = 0;      // chunks issued before this tile (S buffer / barrier phase counter, runs across tiles and items)
      int it = 0;
      for (int item = item_begin; item < item_end; ++item, ++it) {
        const int buf = it & 1;
        mbar_wait_sleep(&bars.ready[buf], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(q_of(buf)), k_addr = smem_u32(k_of(buf)), v_addr = smem_u32(v_of(buf));
        for (int qt = 0; qt < p.mt; ++qt, g0 += p.npl) {
          auto issue_s = [&](int c) {
            const int gc = g0 + c;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
              const uint64_t da = make_smem_desc(q_addr + qt * kTileM * RB + kk * 32, 16, kSbo, kLayout);
              const uint64_t db = make_smem_desc(k_addr + c * kChunk * RB + kk * 32, 16, kSbo, kLayout);
              umma_ss(tmem + (gc & 1) * kChunk, da, db, idesc_s, kk > 0 ? 1u : 0u);
            }
            umma_commit(&bars.s_full[gc & 1]);
          };
          issue_s(0);
          if (p.npl > 1) issue_s(1);
          for (int c = 0; c < p.npl; ++c) {
            const int gc = g0 + c;
            mbar_wait_sleep(&bars.p_full[gc & 1], (gc >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < kChunk / 16; ++kk) {
              const uint64_t db = make_smem_desc(v_addr + (c * kChunk + kk * 16) * RB, 16, kSbo, kLayout);
              umma_ts(tmem + kTmemOFwd, tmem + (gc & 1) * kChunk + kk * 8, db, idesc_o, (c > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&bars.pv_done);
            if (c + 2 < p.npl) issue_s(c + 2);
            else if (c == p.npl - 1) umma_commit(&bars.o_full);
          }
        }
        umma_commit(&bars.free_[buf]);     // every MMA that reads this buffer's operands has retired
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 0-3)
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t t_o = tmem + lane_sel + kTmemOFwd;
    const float cs = p.w.scale_log2;
    int g0 = 0, ot = 0, it = 0;
    for (int item = item_begin; item < item_end; ++item, ++it) {
      const int buf = it & 1;
      const WinInfo wi = window_info(p.w, item);
      mbar_wait(&bars.ready[buf], (it >> 1) & 1);
      const float* tab = s_tab[buf];
      const bool masked = wi.str[0] || wi.str[1] || wi.str[2];    // window-uniform
      for (int qt = 0; qt < p.mt; ++qt, g0 += p.npl, ++ot) {
        const int row = qt * kTileM + tid;
        const int slot = row < n ? row : 0;                        // padded rows compute on slot 0 and are dropped
        const int ai = slot / PLANE, ei = slot - ai * PLANE;
        const int bi = ei / PW, ci = ei - bi * PW;
        // relative-position row term (reference :256-308): idx(i, j) = row_term(i) - col_term(j)
        const float* tab_row = tab + (ai + g.win[0] - 1) * ST0 + (bi + PH - 1) * ST1 + (ci + PW - 1);
        const bool hi_a = ai >= g.win[0] - g.shift[0], hi_b = bi >= PH - PH / 2, hi_c = ci >= PW - PW / 2;
        float m_used = -INFINITY, l = 0.f;

        for (int c = 0; c < p.npl; ++c) {
          const int gc = g0 + c, sb = gc & 1;
          const uint32_t t_s = tmem + lane_sel + sb * kChunk;
          mbar_wait(&bars.s_full[sb], (gc >> 1) & 1);
          tc_fence_after();
          uint32_t sr[64];
          tmem_ld_x32(t_s, sr);
          tmem_ld_x32(t_s + 32, sr + 32);
          tmem_ld_wait();
          const float* trow = tab_row - c * ST0;
          // logit = s * scale*log2e + bias[row_term(i) - col_term(j)]: the key's (b', c') is a compile-time function
          // of e, so the bias gather is one LDS with an immediate offset; pairs of logits share one packed FFMA2
          const float2 cs2 = make_float2(cs, cs);
          if (masked) {
            // shift mask (reference :591-628): -100 when query and key lie in different regions along any axis that
            // straddles a region boundary in this window; the key's region along H / W is a compile-time class of e
            const bool da = wi.str[0] && (hi_a != (c >= g.win[0] - g.shift[0]));
            float mv[2][2];
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
#pragma unroll
              for (int hc = 0; hc < 2; ++hc)
                mv[hb][hc] = (da || (wi.str[1] && hi_b != (hb != 0)) || (wi.str[2] && hi_c != (hc != 0))) ? kMaskLog2 : 0.f;
#pragma unroll
            for (int e = 0; e + 1 < PLANE; e += 2) {
              const int kb0 = e / PW, kc0 = e % PW, kb1 = (e + 1) / PW, kc1 = (e + 1) % PW;
              float2 x = ffma2(make_float2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1])), cs2,
                               make_float2(trow[-(kb0 * ST1 + kc0)], trow[-(kb1 * ST1 + kc1)]));
              x = fadd2(x, make_float2(mv[kb0 >= PH - PH / 2][kc0 >= PW - PW / 2], mv[kb1 >= PH - PH / 2][kc1 >= PW - PW / 2]));
              sr[e] = __float_as_uint(x.x);
              sr[e + 1] = __float_as_uint(x.y);
            }
            if (PLANE & 1) {
              constexpr int e = PLANE - 1, kb = e / PW, kc = e % PW;
              sr[e] = __float_as_uint(fmaf(__uint_as_float(sr[e]), cs, trow[-(kb * ST1 + kc)]) +
                                      mv[kb >= PH - PH / 2][kc >= PW - PW / 2]);
            }
          } else {
#pragma unroll
            for (int e = 0; e + 1 < PLANE; e += 2) {
              const int kb0 = e / PW, kc0 = e % PW, kb1 = (e + 1) / PW, kc1 = (e + 1) % PW;
              const float2 x = ffma2(make_float2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1])), cs2,
                                     make_float2(trow[-(kb0 * ST1 + kc0)], trow[-(kb1 * ST1 + kc1)]));
              sr[e] = __float_as_uint(x.x);
              sr[e + 1] = __float_as_uint(x.y);
            }
            if (PLANE & 1) {
              constexpr int e = PLANE - 1, kb = e / PW, kc = e % PW;
              sr[e] = __float_as_uint(fmaf(__uint_as_float(sr[e]), cs, trow[-(kb * ST1 + kc)]));
            }
          }
          float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]);
#pragma unroll
          for (int e = 2; e + 1 < PLANE; e += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(sr[e]));
            mx1 = fmaxf(mx1, __uint_as_float(sr[e + 1]));
          }
          if (PLANE & 1) mx0 = fmaxf(mx0, __uint_as_float(sr[PLANE - 1]));
          const float m_new = fmaxf(fmaxf(mx0, mx1), m_used);
          const bool need = (m_new - m_used) > kRescale;            // true on the first chunk (m_used = -inf)
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? fast_exp2(m_used - m_new) : 1.0f;
            if (need) m_used = m_new;
            l *= alpha;
            if (c > 0) {
              // O must be quiescent: P V of the previous chunk has completed, and this chunk's is not issued before
              // this warp group signals p_full
              mbar_wait(&bars.pv_done, (gc - 1) & 1);
              tc_fence_after();
              uint32_t o[D];
              if (D == 16) tmem_ld_x16(t_o, o); else tmem_ld_x32(t_o, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < D; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              if (D == 16) tmem_st_x16(t_o, o); else tmem_st_x32(t_o, o);
            }
          }
          float2 sum2 = make_float2(0.f, 0.f);
          const float2 neg_m2 = make_float2(-m_used, -m_used);
          uint32_t pk[32];
#pragma unroll
          for (int e = 0; e < kChunk; e += 2) {
            if (e + 1 < PLANE) {
              const float2 t = fadd2(make_float2(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1])), neg_m2);
              const float e0 = fast_exp2(t.x), e1 = fast_exp2(t.y);
              sum2 = fadd2(sum2, make_float2(e0, e1));
              pk[e / 2] = pack_bf16x2(e0, e1);
            } else if (e < PLANE) {
              const float e0 = fast_exp2(__uint_as_float(sr[e]) - m_used);
              sum2.x += e0;
              pk[e / 2] = pack_bf16x2(e0, 0.f);
            } else {
              pk[e / 2] = 0u;                      // the chunk's dummy keys
            }
          }
          l += sum2.x + sum2.y;
          tmem_st_x32(t_s, pk);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.p_full[sb]);
        }

        // ---- epilogue: O / l -> bf16 -> the token's row of `out`; log2-domain lse
        mbar_wait(&bars.o_full, ot & 1);
        tc_fence_after();
        uint32_t o[D];
        if (D == 16) tmem_ld_x16(t_o, o); else tmem_ld_x32(t_o, o);
        tmem_ld_wait();
        tc_fence_before();
        if (row < n) {
          int tok, reg;
          slot_lookup(g, wi.w, row, tok, reg);
          if (tok >= 0) {
            const float inv_l = 1.0f / l;
            const size_t bt = static_cast<size_t>(wi.b) * g.T + tok;
            uint4* dst = reinterpret_cast<uint4*>(p.w.out + bt * p.w.C + wi.h * D);
#pragma unroll
            for (int c16 = 0; c16 < D / 8; ++c16) {
              uint4 val;
              val.x = pack_bf16x2(__uint_as_float(o[c16 * 8 + 0]) * inv_l, __uint_as_float(o[c16 * 8 + 1]) * inv_l);
              val.y = pack_bf16x2(__uint_as_float(o[c16 * 8 + 2]) * inv_l, __uint_as_float(o[c16 * 8 + 3]) * inv_l);
              val.z = pack_bf16x2(__uint_as_float(o[c16 * 8 + 4]) * inv_l, __uint_as_float(o[c16 * 8 + 5]) * inv_l);
              val.w = pack_bf16x2(__uint_as_float(o[c16 * 8 + 6]) * inv_l, __uint_as_float(o[c16 * 8 + 7]) * inv_l);
              dst[c16] = val;
            }
            p.w.lse2[bt * p.w.H + wi.h] = m_used + log2f(l);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.free_[buf]);   // this warp no longer reads the buffer's bias table
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemColsFwd);
}

// =====================================================================================================
// backward (one kernel, S recomputed once)
// =====================================================================================================
// Transposed formulation like the dense backward (dense_attn_bwd.cu): the exponentiating threads own KEY rows, so P^T and
// dS^T come out in the layout the next GEMMs consume. Per item (window, head), for each 128-key tile kt and each query
// plane a (64-column step, S^T / dP^T double-buffered in TMEM):
//   S^T  = K_kt Q_a^T - lse2/c      A = K tile (smem, K-major)   B = Q plane (smem, K-major)  + one K=16 step carrying the
//   dP^T = V_kt dO_a^T - D          A = V tile                   B = dO plane                    per-query term (hi,mid,lo)
//   compute warps 0-3 (thread = key row, all columns of the plane):
//     P^T = exp2(S^T c + bias [+ mask]);  dS^T = P^T o dP^T;  P^T, dS^T -> bf16 in place in TMEM;  dS^T -> swizzled smem tile
//   table warps 6-9 (thread = key row again, one warp per 32 keys, sharing an SM sub-partition with a compute warp):
//     read their rows of the dS^T smem tile back and do
//     d(bias table)[idx(i, j)] += dS^T into a per-warp fp32 table in shared memory, without atomics: KEYS are held in
//     W-MAJOR order (row = c'*(WD*PH) + a'*PH + b'), so the 32 keys of a warp share one c' (two adjacent ones where a
//     warp straddles a group). Two (key, query) pairs of a warp then hit the same table entry only if the queries' c
//     differ by the keys' c' difference (0 or +-1): queries whose c differ by >= 2 never collide. The read-modify-write
//     runs in rounds - one plane row b, even c then odd c - with a warp barrier between rounds; inside a round the
//     loads, adds and stores of 3-4 entries are independent.
//   dV_kt += P^T dO_a,  dK_kt += dS^T Q_a   (TS MMAs, B MN-major);  every second plane: dQ_pair += dS K_kt (A = the dS^T
//   smem tile read MN-major, M = 128 queries = two planes).  dK/dV leave TMEM per key tile, dQ (all planes) per item.
constexpr int kBwdThreads = 320;          // warps 0-3 compute, 4 issuer, 5 gather, 6-9 d(bias table)
constexpr int kAugBytes = kChunk * 16;       // [64 query rows x 8 bf16], un-swizzled core-matrix layout (row r at r*16)
constexpr uint32_t kTmemS = 0, kTmemDP = 128, kTmemDV = 256;

struct TcBwdParams {
  WinParams w;
  int gs;                 // keys per W-major group = window depth * plane height
  int gsp;                // rows per group in shared memory: gs rounded up to 4 (TMA destinations are 128-byte aligned)
  int n_krows;            // PW * gsp
  int n_items, items_per_cta;
  int kt;                 // 128-row key tiles per window
  int npl;                // query planes per window
  int nbuf;               // operand buffers (2: the next item's loads overlap this item; 1 when shared memory is short)
  uint32_t kv_bytes, q_bytes, aug_bytes, buf_bytes;     // per buffer: k | v | q | dO | lse_aug | d_aug
  uint32_t off_ones, off_zeros, off_ds, off_dtab, off_bars;
  uint32_t dtab_stride;   // bytes between the compute warps' d(bias table) accumulators
};

struct BwdBars {
  uint64_t full[2], ready[2], free_[2];
  uint64_t sdp_full[2], pds_full[2], dq_done, dkv_full, dkv_drained, dq_full, dq_drained;
  uint64_t ds_read[2];      // the table warps have read atom x of the dS^T tile
  uint32_t tmem_base;
  int buf_head[2];
};

__device__ __forceinline__ uint4 split3_bf16_tc(float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  uint4 v;
  v.x = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(mid)) << 16);
  v.y = static_cast<uint32_t>(__bfloat16_as_ushort(lo));
  v.z = 0u;
  v.w = 0u;
  return v;
}

template <int D, int PH, int PW>
__global__ void __launch_bounds__(kBwdThreads, 1)
win_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_colgrp, const __grid_constant__ CUtensorMap tm_plane,
                       const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_do_plane,
                       const __grid_constant__ CUtensorMap tm_do_rows, const TcBwdParams p) {
  constexpr int RB = D * 2;
  constexpr uint32_t kLayout = D == 16 ? kLayoutSW32 : kLayoutSW64;
  constexpr uint32_t kSbo = 8 * RB;
  constexpr int PLANE = PH * PW;
  constexpr int ST1 = 2 * PW - 1, ST0 = ST1 * (2 * PH - 1);
  constexpr uint32_t kTmemDK = kTmemDV + D, kTmemDQ = kTmemDV + 2 * D;
  constexpr int kAtomBytes = kTileM * 128;            // one [128 keys x 64 queries] bf16 SW128 atom of the dS^T tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Static shared arrays keep the address space visible to the compiler: the per-logit bias gather and the d(bias table)
  // read-modify-write are then LDS / STS with immediate offsets (through the re-aligned dynamic pointer they would be
  // generic LD / ST). The four accumulators fit the 48 KB static limit for 7x7 planes; 8x8 planes keep them dynamic.
  constexpr int kTabMax = (2 * PH - 1) * (2 * PH - 1) * (2 * PW - 1);      // window depth <= PH (win_attn_tc_applicable)
  constexpr bool kStaticDtab = (5 * kTabMax * 4 <= 46 * 1024);
  __shared__ float s_tab[kTabMax];                     // the item's head of the bias table, times log2(e)
  __shared__ float s_dtab[kStaticDtab ? 4 : 1][kStaticDtab ? kTabMax : 1];
  __shared__ float s_dpad[3 * D];                      // gradient reaching qkv.bias through the pad tokens (current head)
  BwdBars& bars = *reinterpret_cast<BwdBars*>(smem + p.off_bars);
  auto k_of = [&](int buf) { return smem + buf * p.buf_bytes; };
  auto v_of = [&](int buf) { return smem + buf * p.buf_bytes + p.kv_bytes; };
  auto q_of = [&](int buf) { return smem + buf * p.buf_bytes + 2 * p.kv_bytes; };
  auto do_of = [&](int buf) { return smem + buf * p.buf_bytes + 2 * p.kv_bytes + p.q_bytes; };
  auto lse_of = [&](int buf) { return smem + buf * p.buf_bytes + 2 * p.kv_bytes + 2 * p.q_bytes; };
  auto dsum_of = [&](int buf) { return smem + buf * p.buf_bytes + 2 * p.kv_bytes + 2 * p.q_bytes + p.aug_bytes; };
  uint8_t* const s_ones = smem + p.off_ones;
  uint8_t* const s_zeros = smem + p.off_zeros;
  uint8_t* const s_ds = smem + p.off_ds;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const WinGeom& g = p.w.g;
  const int n = g.n;
  const int item_begin = blockIdx.x * p.items_per_cta;
  const int item_end = min(item_begin + p.items_per_cta, p.n_items);
  const int n_pairs = (p.npl + 1) >> 1;

  // zero everything up to the barriers: operand padding rows, dS^T tile, d(bias table) accumulators, the shared zeros
  for (uint32_t i = tid; i < p.off_bars / 16; i += kBwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (kStaticDtab)
    for (int i = tid; i < 4 * kTabMax; i += kBwdThreads) (&s_dtab[0][0])[i] = 0.f;
  if (tid < 3 * D) s_dpad[tid] = 0.f;
  __syncthreads();
  if (tid < kTileM) *reinterpret_cast<uint4*>(s_ones + tid * 16) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.full[b], 1);
      mbar_init(&bars.ready[b], 1);
      mbar_init(&bars.free_[b], 5);
      mbar_init(&bars.sdp_full[b], 1);
      mbar_init(&bars.pds_full[b], 4);
      bars.buf_head[b] = -1;
    }
    mbar_init(&bars.ds_read[0], 4);
    mbar_init(&bars.ds_read[1], 4);
    mbar_init(&bars.dq_done, 1);
    mbar_init(&bars.dkv_full, 1);
    mbar_init(&bars.dkv_drained, 4);
    mbar_init(&bars.dq_full, 1);
    mbar_init(&bars.dq_drained, 4);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 4) {
    tmem_alloc(&bars.tmem_base, 512);
    tmem_relinquish();
  }
  if (warp == 5 && elect_one()) {
    tma_prefetch_desc(&tm_colgrp);
    tma_prefetch_desc(&tm_plane);
    tma_prefetch_desc(&tm_rows);
    tma_prefetch_desc(&tm_do_plane);
    tma_prefetch_desc(&tm_do_rows);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 5) {
    // ------------------------------------------------------------------ gather (all 32 lanes)
    constexpr int GP = (PLANE + 3) / 4;
    int it = 0;
    for (int item = item_begin; item < item_end; ++item, ++it) {
      const int buf = p.nbuf == 2 ? (it & 1) : 0, use = p.nbuf == 2 ? (it >> 1) : it;
      if (it >= p.nbuf) mbar_wait_sleep(&bars.free_[buf], (use - 1) & 1);
      const WinInfo wi = window_info(p.w, item);
      uint8_t *kd = k_of(buf), *vd = v_of(buf), *qd = q_of(buf), *dod = do_of(buf);
      const int col_q = wi.h * D, col_k = p.w.C + wi.h * D, col_v = 2 * p.w.C + wi.h * D;
      const int kg = p.n_krows / 4;
      if (wi.box) {
        if (lane == 0) mbar_expect_tx(&bars.full[buf], static_cast<uint32_t>(2 * n + 2 * p.npl * PLANE) * RB);
        __syncwarp();
        if (lane < PW) {       // keys / values: one (D, 1, PH, WD) box per window column c', rows land as a'*PH + b'
          tma_load_5d(kd + lane * p.gsp * RB, &tm_colgrp, &bars.full[buf], col_k, wi.start[2] + lane, wi.start[1], wi.start[0], wi.b);
          tma_load_5d(vd + lane * p.gsp * RB, &tm_colgrp, &bars.full[buf], col_v, wi.start[2] + lane, wi.start[1], wi.start[0], wi.b);
        }
        if (lane < p.npl) {
          tma_load_5d(qd + lane * kChunk * RB, &tm_plane, &bars.full[buf], col_q, wi.start[2], wi.start[1],
                      wi.start[0] + lane, wi.b);
          tma_load_5d(dod + lane * kChunk * RB, &tm_do_plane, &bars.full[buf], wi.h * D, wi.start[2], wi.start[1],
                      wi.start[0] + lane, wi.b);
        }
      } else {
        if (lane == 0) mbar_expect_tx(&bars.full[buf], static_cast<uint32_t>(2 * kg + 2 * p.npl * GP) * 4 * RB);
        __syncwarp();
        const int oob = p.w.B * g.T, row0 = wi.b * g.T;
        for (int gq = lane; gq < kg; gq += 32) {       // k, v: W-major key order
          int r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int kr = gq * 4 + u;
            int tok = -1, reg;
            const int cj = kr / p.gsp, rem = kr - cj * p.gsp;        // rem = a' * PH + b' (>= gs: a padding row)
            if (rem < p.gs) slot_lookup(g, wi.w, rem * PW + cj, tok, reg);
            r[u] = tok >= 0 ? row0 + tok : oob;
          }
          tma_gather4(kd + gq * 4 * RB, &tm_rows, &bars.full[buf], col_k, r[0], r[1], r[2], r[3]);
          tma_gather4(vd + gq * 4 * RB, &tm_rows, &bars.full[buf], col_v, r[0], r[1], r[2], r[3]);
        }
        for (int idx = lane; idx < p.npl * GP; idx += 32) {   // q, dO: plane-padded slot order
          const int a = idx / GP, gg = idx - a * GP;
          int r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = gg * 4 + u;
            int tok = -1, reg;
            if (e < PLANE) slot_lookup(g, wi.w, a * PLANE + e, tok, reg);
            r[u] = tok >= 0 ? row0 + tok : oob;
          }
          const uint32_t off = static_cast<uint32_t>(a * kChunk + gg * 4) * RB;
          tma_gather4(qd + off, &tm_rows, &bars.full[buf], col_q, r[0], r[1], r[2], r[3]);
          tma_gather4(dod + off, &tm_do_rows, &bars.full[buf], wi.h * D, r[0], r[1], r[2], r[3]);
        }
      }
      // per-query terms of the extra k-step: -lse2 / c and -D (D = rowsum(dO o O)); pad and dummy queries get a huge
      // negative score offset so that their P is exactly 0
      {
        const float inv_c = 1.0f / p.w.scale_log2;
        uint8_t *la = lse_of(buf), *da = dsum_of(buf);
        for (int idx = lane; idx < p.npl * kChunk; idx += 32) {
          const int a = idx >> 6, e = idx & 63;
          float lv = -1e30f, dv = 0.f;
          if (e < PLANE) {
            int tok, reg;
            slot_lookup(g, wi.w, a * PLANE + e, tok, reg);
            if (tok >= 0) {
              const size_t gi = (static_cast<size_t>(wi.b) * g.T + tok) * p.w.H + wi.h;
              lv = -p.w.lse2[gi] * inv_c;
              dv = -p.w.dsum[gi];
            }
          }
          *reinterpret_cast<uint4*>(la + a * kAugBytes + e * 16) = split3_bf16_tc(lv);
          *reinterpret_cast<uint4*>(da + a * kAugBytes + e * 16) = split3_bf16_tc(dv);
        }
      }
      mbar_wait_sleep(&bars.full[buf], use & 1);
      if (wi.has_pad && p.w.qkv_bias != nullptr) {
        for (int s = lane; s < n; s += 32) {
          int tok, reg;
          slot_lookup(g, wi.w, s, tok, reg);
          if (tok >= 0) continue;
          const int a = s / PLANE, e = s - a * PLANE;
          const uint32_t qr = a * kChunk + e;                        // plane-padded query row
          const uint32_t kr = (e % PW) * p.gsp + s / PW;             // W-major key row: c' * gsp + (a' * PH + b')
#pragma unroll
          for (int c = 0; c < D / 8; ++c) {
            uint4 val[3];
#pragma unroll
            for (int sel = 0; sel < 3; ++sel) {
              const float* bsrc = p.w.qkv_bias + sel * p.w.C + wi.h * D + c * 8;
              val[sel].x = pack2_bf16(bsrc[0], bsrc[1]);
              val[sel].y = pack2_bf16(bsrc[2], bsrc[3]);
              val[sel].z = pack2_bf16(bsrc[4], bsrc[5]);
              val[sel].w = pack2_bf16(bsrc[6], bsrc[7]);
            }
            *reinterpret_cast<uint4*>(qd + swz_off<RB>(qr, c)) = val[0];
            *reinterpret_cast<uint4*>(kd + swz_off<RB>(kr, c)) = val[1];
            *reinterpret_cast<uint4*>(vd + swz_off<RB>(kr, c)) = val[2];
          }
        }
      }
      fence_proxy_async_smem();           // aug tiles and pad rows were written with ordinary stores
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.ready[buf]);
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_nt = make_idesc_bf16(kTileM, kChunk, 0, 0);     // S^T, dP^T: 128 keys x 64 queries
      constexpr uint32_t idesc_kmn = make_idesc_bf16(kTileM, D, 0, 1);         // dV, dK: A K-major (TMEM), B MN-major
      constexpr uint32_t idesc_mnmn = make_idesc_bf16(kTileM, D, 1, 1);        // dQ: A MN-major, B MN-major
      const uint32_t zeros_addr = smem_u32(s_zeros);
      auto aug_desc = [&](const void* tile) {      // k-half 0 in the tile, k-half 1 = the shared zeros (LBO)
        return make_smem_desc(smem_u32(tile), zeros_addr - smem_u32(tile), 128, 0);
      };
      const uint64_t d_ones = aug_desc(s_ones);
      const uint64_t d_ds = make_smem_desc(smem_u32(s_ds), kAtomBytes, 1024, kLayoutSW128);
      int gs0 = 0;      // plane steps issued before this key tile (S^T / dP^T buffer + barrier phase counter)
      int gt = 0;       // key tiles issued (dkv_full / dkv_drained phases)
      int gp = 0;       // dQ MMAs issued (dq_done phase)
      int it = 0;
      for (int item = item_begin; item < item_end; ++item, ++it) {
        const int buf = p.nbuf == 2 ? (it & 1) : 0, use = p.nbuf == 2 ? (it >> 1) : it;
        mbar_wait_sleep(&bars.ready[buf], use & 1);
        tc_fence_after();
        // operand descriptors are built once per item; per MMA only the 14-bit start-address field advances
        const uint64_t d_k = make_smem_desc(smem_u32(k_of(buf)), 16, kSbo, kLayout);
        const uint64_t d_v = make_smem_desc(smem_u32(v_of(buf)), 16, kSbo, kLayout);
        const uint64_t d_q = make_smem_desc(smem_u32(q_of(buf)), 16, kSbo, kLayout);
        const uint64_t d_do = make_smem_desc(smem_u32(do_of(buf)), 16, kSbo, kLayout);
        const uint64_t d_lse = aug_desc(lse_of(buf)), d_dsum = aug_desc(dsum_of(buf));
        auto adv = [](uint64_t d, uint32_t bytes) { return d + (bytes >> 4); };
        // an aug tile further up: the start address grows and the distance to the shared zeros (LBO) shrinks by as much
        auto adv_aug = [](uint64_t d, uint32_t bytes) { const uint64_t st = bytes >> 4; return d + st - (st << 16); };
        if (it > 0) mbar_wait_sleep(&bars.dq_drained, (it - 1) & 1);    // the previous item's dQ left TMEM
        for (int kt = 0; kt < p.kt; ++kt, gs0 += p.npl, ++gt) {
          auto issue_sdp = [&](int a) {
            const int b = (gs0 + a) & 1;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
              umma_ss(tmem + kTmemS + b * kChunk, adv(d_k, kt * kTileM * RB + kk * 32), adv(d_q, a * kChunk * RB + kk * 32),
                      idesc_nt, kk > 0 ? 1u : 0u);
            umma_ss(tmem + kTmemS + b * kChunk, d_ones, adv_aug(d_lse, a * kAugBytes), idesc_nt, 1u);     // - lse2 / c
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
              umma_ss(tmem + kTmemDP + b * kChunk, adv(d_v, kt * kTileM * RB + kk * 32), adv(d_do, a * kChunk * RB + kk * 32),
                      idesc_nt, kk > 0 ? 1u : 0u);
            umma_ss(tmem + kTmemDP + b * kChunk, d_ones, adv_aug(d_dsum, a * kAugBytes), idesc_nt, 1u);   // - D
            umma_commit(&bars.sdp_full[b]);
          };
          issue_sdp(0);
          if (p.npl > 1) issue_sdp(1);
          for (int a = 0; a < p.npl; ++a) {
            const int gs = gs0 + a, b = gs & 1;
            if (a == 0 && gt > 0) mbar_wait_sleep(&bars.dkv_drained, (gt - 1) & 1);   // previous key tile's dV/dK left TMEM
            mbar_wait_sleep(&bars.pds_full[b], (gs >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < kChunk / 16; ++kk)       // dV += P^T dO_a
              umma_ts(tmem + kTmemDV, tmem + kTmemS + b * kChunk + kk * 8, adv(d_do, (a * kChunk + kk * 16) * RB), idesc_kmn,
                      (a > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < kChunk / 16; ++kk)       // dK += dS^T Q_a
              umma_ts(tmem + kTmemDK, tmem + kTmemDP + b * kChunk + kk * 8, adv(d_q, (a * kChunk + kk * 16) * RB), idesc_kmn,
                      (a > 0 || kk > 0) ? 1u : 0u);
            if (a + 2 < p.npl) issue_sdp(a + 2);
            if ((a & 1) || a == p.npl - 1) {
              // dQ_pair += dS K_kt over the 128 queries of planes (a & ~1, a | 1); for an unpaired last plane the second
              // atom holds stale values that only reach dummy rows of the accumulator
#pragma unroll
              for (int kk = 0; kk < kTileM / 16; ++kk)
                umma_ss(tmem + kTmemDQ + (a >> 1) * D, d_ds + ((kk * 2048) >> 4), adv(d_k, (kt * kTileM + kk * 16) * RB),
                        idesc_mnmn, (kt > 0 || kk > 0) ? 1u : 0u);
              umma_commit(&bars.dq_done);
              ++gp;
            }
          }
          umma_commit(&bars.dkv_full);
        }
        umma_commit(&bars.dq_full);
        umma_commit(&bars.free_[buf]);
      }
      (void)gp;
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ compute (thread = key row, W-major key order)
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const float cs = p.w.scale_log2;
    const float sc = p.w.scale_log2 / kLog2e;            // the softmax scale itself (q is scaled before q k^T)
    int cur_head = -1;
    auto flush_head = [&](int head) {       // accumulated pad-token bias gradient -> global, then cleared
      if (head < 0) return;
      named_bar_sync(1, 128);               // every warp's pad-gradient adds of the old head have landed
      if (tid < 3 * D) {
        const float v = s_dpad[tid];
        if (v != 0.f && p.w.dbias_pad != nullptr) atomicAdd(p.w.dbias_pad + (tid / D) * p.w.C + head * D + (tid % D), v);
        s_dpad[tid] = 0.f;
      }
      named_bar_sync(1, 128);
    };
    // a pad token's gradient row: summed over the warp's pad rows with shuffles, one shared-memory add per column
    auto add_pad_rows = [&](const uint32_t* r, bool is_pad, float mul, int sel) {
      if (!__any_sync(0xffffffffu, is_pad)) return;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        float v = is_pad ? __uint_as_float(r[j]) * mul : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(&s_dpad[sel * D + j], v);
      }
    };
    int gs0 = 0, gt = 0, gp_seen = 0, it = 0;
    int atom_writes[2] = {0, 0};
    for (int item = item_begin; item < item_end; ++item, ++it) {
      const int buf = p.nbuf == 2 ? (it & 1) : 0, use = p.nbuf == 2 ? (it >> 1) : it;
      const WinInfo wi = window_info(p.w, item);
      if (wi.h != cur_head) {
        flush_head(cur_head);
        for (int t = tid; t < g.tab_rows; t += 128) s_tab[t] = p.w.table[static_cast<size_t>(t) * p.w.H + wi.h] * kLog2e;
        named_bar_sync(1, 128);
        cur_head = wi.h;
      }
      mbar_wait(&bars.ready[buf], use & 1);
      const bool masked = wi.str[0] || wi.str[1] || wi.str[2];
      const size_t tok_base = static_cast<size_t>(wi.b) * g.T;
      for (int kt = 0; kt < p.kt; ++kt, gs0 += p.npl, ++gt) {
        const int row = kt * kTileM + tid;                 // key row, W-major: c' * gsp + a' * PH + b'
        const bool row_valid = row < p.n_krows && (row % p.gsp) < p.gs;
        const int krow = row_valid ? row : 0;
        const int cj = krow / p.gsp, rem = krow - cj * p.gsp;
        const int aj = rem / PH, bj = rem - aj * PH;
        const int key_slot = rem * PW + cj;                // raster slot of this key in the window
        // idx(i, j) = row_term(i) - col_term(j); this thread's key fixes col_term, the query plane adds a * ST0 and the
        // query's (b, c) is a compile-time offset
        const int t_base = (g.win[0] - 1 - aj) * ST0 + (PH - 1 - bj) * ST1 + (PW - 1 - cj);
        const bool hi_a = aj >= g.win[0] - g.shift[0], hi_b = bj >= PH - PH / 2, hi_c = cj >= PW - PW / 2;

        for (int a = 0; a < p.npl; ++a) {
          const int gs = gs0 + a, b = gs & 1;
          mbar_wait(&bars.sdp_full[b], (gs >> 1) & 1);
          tc_fence_after();
          // only the plane's PLANE columns are loaded (49 = 32 + 16 + 1): registers are short with three warps on two
          // of the SM sub-partitions
          uint32_t sv[PLANE], dpv[PLANE];
          tmem_ld_x32(tmem + lane_sel + kTmemS + b * kChunk, sv);
          tmem_ld_x32(tmem + lane_sel + kTmemDP + b * kChunk, dpv);
          if (PLANE == 64) {
            tmem_ld_x32(tmem + lane_sel + kTmemS + b * kChunk + 32, sv + 32);
            tmem_ld_x32(tmem + lane_sel + kTmemDP + b * kChunk + 32, dpv + 32);
          } else {
            tmem_ld_x16(tmem + lane_sel + kTmemS + b * kChunk + 32, sv + 32);
            tmem_ld_x16(tmem + lane_sel + kTmemDP + b * kChunk + 32, dpv + 32);
            tmem_ld_x1(tmem + lane_sel + kTmemS + b * kChunk + 48, sv + 48);
            tmem_ld_x1(tmem + lane_sel + kTmemDP + b * kChunk + 48, dpv + 48);
          }
          tmem_ld_wait();
          const float* trow = s_tab + t_base + a * ST0;
          float mv[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
          if (masked) {
            const bool da = wi.str[0] && (hi_a != (a >= g.win[0] - g.shift[0]));
#pragma unroll
            for (int hb = 0; hb < 2; ++hb)
#pragma unroll
              for (int hc = 0; hc < 2; ++hc)
                mv[hb][hc] = (da || (wi.str[1] && hi_b != (hb != 0)) || (wi.str[2] && hi_c != (hc != 0))) ? kMaskLog2 : 0.f;
          }
          // P^T = exp2(s c + bias [+ mask]) (the per-query -lse2 arrived through the extra k-step), dS^T = P^T o dP^T
          // (a phase-structured variant with packed FFMA2 / FMUL2 was measured slower: 1552 vs 1459 us per cfg4 block)
          uint32_t pk[32], dsk[32];
#pragma unroll
          for (int e = 0; e < kChunk; e += 2) {
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            if (e < PLANE) {
              const int qb = e / PW, qc = e % PW;
              float x = fmaf(__uint_as_float(sv[e]), cs, trow[qb * ST1 + qc]);
              if (masked) x += mv[qb >= PH - PH / 2][qc >= PW - PW / 2];
              p0 = fast_exp2(x);
              d0 = p0 * __uint_as_float(dpv[e]);
            }
            if (e + 1 < PLANE) {
              const int qb = (e + 1) / PW, qc = (e + 1) % PW;
              float x = fmaf(__uint_as_float(sv[e + 1]), cs, trow[qb * ST1 + qc]);
              if (masked) x += mv[qb >= PH - PH / 2][qc >= PW - PW / 2];
              p1 = fast_exp2(x);
              d1 = p1 * __uint_as_float(dpv[e + 1]);
            }
            pk[e / 2] = pack_bf16x2(p0, p1);
            dsk[e / 2] = pack_bf16x2(d0, d1);
          }
          tmem_st_x32(tmem + lane_sel + kTmemS + b * kChunk, pk);
          tmem_st_x32(tmem + lane_sel + kTmemDP + b * kChunk, dsk);
          // dS^T -> the smem tile dQ reads (atom = plane parity). The tile is single-buffered: the dQ MMA of the previous
          // plane pair must have retired before the first atom is overwritten (it was issued a whole step ago).
          if ((a & 1) == 0 && gp_seen > 0) mbar_wait(&bars.dq_done, (gp_seen - 1) & 1);
          // ... and the table warps must have read this atom's previous contents
          if (atom_writes[a & 1] > 0) mbar_wait(&bars.ds_read[a & 1], (atom_writes[a & 1] - 1) & 1);
          ++atom_writes[a & 1];
          uint8_t* atom = s_ds + (a & 1) * kAtomBytes;
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16)
            *reinterpret_cast<uint4*>(atom + sw128_offset(tid, c16)) =
                make_uint4(dsk[c16 * 4], dsk[c16 * 4 + 1], dsk[c16 * 4 + 2], dsk[c16 * 4 + 3]);
          if ((a & 1) || a == p.npl - 1) ++gp_seen;
          tmem_st_wait();
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.pds_full[b]);
        }

        // ---- key-tile epilogue: dV, dK (scaled) -> the tokens' rows of dqkv; pad keys feed d(qkv.bias)
        mbar_wait(&bars.dkv_full, gt & 1);
        tc_fence_after();
        uint32_t rv[D], rk[D];
        if (D == 16) { tmem_ld_x16(tmem + lane_sel + kTmemDV, rv); tmem_ld_x16(tmem + lane_sel + kTmemDK, rk); }
        else { tmem_ld_x32(tmem + lane_sel + kTmemDV, rv); tmem_ld_x32(tmem + lane_sel + kTmemDK, rk); }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.dkv_drained);
        int tok = -2, reg;
        if (row_valid) slot_lookup(g, wi.w, key_slot, tok, reg);
        if (tok >= 0) {
          __nv_bfloat16* base = p.w.dqkv + (tok_base + tok) * 3 * p.w.C + wi.h * D;
          uint4* dk_dst = reinterpret_cast<uint4*>(base + p.w.C);
          uint4* dv_dst = reinterpret_cast<uint4*>(base + 2 * p.w.C);
#pragma unroll
          for (int c16 = 0; c16 < D / 8; ++c16) {
            uint4 kv, vv;
            kv.x = pack_bf16x2(__uint_as_float(rk[c16 * 8 + 0]) * sc, __uint_as_float(rk[c16 * 8 + 1]) * sc);
            kv.y = pack_bf16x2(__uint_as_float(rk[c16 * 8 + 2]) * sc, __uint_as_float(rk[c16 * 8 + 3]) * sc);
            kv.z = pack_bf16x2(__uint_as_float(rk[c16 * 8 + 4]) * sc, __uint_as_float(rk[c16 * 8 + 5]) * sc);
            kv.w = pack_bf16x2(__uint_as_float(rk[c16 * 8 + 6]) * sc, __uint_as_float(rk[c16 * 8 + 7]) * sc);
            vv.x = pack_bf16x2(__uint_as_float(rv[c16 * 8 + 0]), __uint_as_float(rv[c16 * 8 + 1]));
            vv.y = pack_bf16x2(__uint_as_float(rv[c16 * 8 + 2]), __uint_as_float(rv[c16 * 8 + 3]));
            vv.z = pack_bf16x2(__uint_as_float(rv[c16 * 8 + 4]), __uint_as_float(rv[c16 * 8 + 5]));
            vv.w = pack_bf16x2(__uint_as_float(rv[c16 * 8 + 6]), __uint_as_float(rv[c16 * 8 + 7]));
            dk_dst[c16] = kv;
            dv_dst[c16] = vv;
          }
        }
        if (wi.has_pad) {
          add_pad_rows(rk, tok == -1, sc, 1);
          add_pad_rows(rv, tok == -1, 1.0f, 2);
        }
      }

      // ---- item epilogue: dQ (scaled) of every plane pair
      mbar_wait(&bars.dq_full, it & 1);
      tc_fence_after();
      for (int pr = 0; pr < n_pairs; ++pr) {
        uint32_t rq[D];
        if (D == 16) tmem_ld_x16(tmem + lane_sel + kTmemDQ + pr * D, rq); else tmem_ld_x32(tmem + lane_sel + kTmemDQ + pr * D, rq);
        tmem_ld_wait();
        const int a = 2 * pr + (tid >> 6), e = tid & 63;
        int tok = -2, reg;
        if (a < p.npl && e < PLANE) slot_lookup(g, wi.w, a * PLANE + e, tok, reg);
        if (tok >= 0) {
          uint4* dq_dst = reinterpret_cast<uint4*>(p.w.dqkv + (tok_base + tok) * 3 * p.w.C + wi.h * D);
#pragma unroll
          for (int c16 = 0; c16 < D / 8; ++c16) {
            uint4 qv;
            qv.x = pack_bf16x2(__uint_as_float(rq[c16 * 8 + 0]) * sc, __uint_as_float(rq[c16 * 8 + 1]) * sc);
            qv.y = pack_bf16x2(__uint_as_float(rq[c16 * 8 + 2]) * sc, __uint_as_float(rq[c16 * 8 + 3]) * sc);
            qv.z = pack_bf16x2(__uint_as_float(rq[c16 * 8 + 4]) * sc, __uint_as_float(rq[c16 * 8 + 5]) * sc);
            qv.w = pack_bf16x2(__uint_as_float(rq[c16 * 8 + 6]) * sc, __uint_as_float(rq[c16 * 8 + 7]) * sc);
            dq_dst[c16] = qv;
          }
        }
        if (wi.has_pad) add_pad_rows(rq, tok == -1, sc, 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.dq_drained);
        mbar_arrive(&bars.free_[buf]);
      }
    }
    flush_head(cur_head);
  }

  if (warp >= 6) {
    // ------------------------------------------------------------------ d(relative_position_bias_table) (warps 6-9)
    const int qw = warp - 6;                               // which 32 keys of the tile
    float* const dtab = kStaticDtab ? s_dtab[kStaticDtab ? qw : 0] : reinterpret_cast<float*>(smem + p.off_dtab + qw * p.dtab_stride);
    int cur_head = -1;
    auto flush_dtab = [&](int head) {       // this warp's accumulated d(bias table) -> global, then cleared
      if (head < 0) return;
      for (int t = lane; t < g.tab_rows; t += 32) {
        const float v = dtab[t];
        if (v != 0.f) atomicAdd(p.w.dtable + static_cast<size_t>(t) * p.w.H + head, v);
        dtab[t] = 0.f;
      }
      __syncwarp();
    };
    int gs0 = 0;
    for (int item = item_begin; item < item_end; ++item) {
      const WinInfo wi = window_info(p.w, item);
      if (wi.h != cur_head) {
        flush_dtab(cur_head);
        cur_head = wi.h;
      }
      for (int kt = 0; kt < p.kt; ++kt, gs0 += p.npl) {
        const int trow_i = qw * 32 + lane;                 // row inside the 128-key tile
        const int row = kt * kTileM + trow_i;              // key row, W-major: c' * gsp + a' * PH + b'
        const bool row_valid = row < p.n_krows && (row % p.gsp) < p.gs;
        const int krow = row_valid ? row : 0;
        const int cj = krow / p.gsp, rem = krow - cj * p.gsp;
        const int aj = rem / PH, bj = rem - aj * PH;
        const int t_base = (g.win[0] - 1 - aj) * ST0 + (PH - 1 - bj) * ST1 + (PW - 1 - cj);
        for (int a = 0; a < p.npl; ++a) {
          const int gs = gs0 + a, b = gs & 1;
          mbar_wait(&bars.pds_full[b], (gs >> 1) & 1);     // the compute warps have written this plane's dS^T atom
          const uint8_t* atom = s_ds + (a & 1) * kAtomBytes;
          uint32_t raw[28];                                 // 56 bf16 >= PLANE query columns of this key row
#pragma unroll
          for (int c16 = 0; c16 < 7; ++c16) {
            const uint4 x = *reinterpret_cast<const uint4*>(atom + sw128_offset(trow_i, c16));
            raw[c16 * 4] = x.x; raw[c16 * 4 + 1] = x.y; raw[c16 * 4 + 2] = x.z; raw[c16 * 4 + 3] = x.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.ds_read[a & 1]);
          float* drow = dtab + t_base + a * ST0;
          auto dsv = [&](int e) {                            // element e of the row (bf16 -> fp32)
            return (e & 1) ? __uint_as_float(raw[e >> 1] & 0xFFFF0000u) : __uint_as_float(raw[e >> 1] << 16);
          };
          // hazard-free rounds (see the header): plane row qb, even query columns then odd ones (with 8-wide planes a
          // warp never straddles two key groups, so a whole row is one round); rows past the window do not store
#pragma unroll
          for (int qb = 0; qb < PH; ++qb) {
#pragma unroll
            for (int par = 0; par < ((PW & 1) ? 2 : 1); ++par) {
              float acc[PW];
#pragma unroll
              for (int qc = par; qc < PW; qc += ((PW & 1) ? 2 : 1)) acc[qc] = drow[qb * ST1 + qc];
#pragma unroll
              for (int qc = par; qc < PW; qc += ((PW & 1) ? 2 : 1)) acc[qc] += dsv(qb * PW + qc);
#pragma unroll
              for (int qc = par; qc < PW; qc += ((PW & 1) ? 2 : 1))
                if (row_valid) drow[qb * ST1 + qc] = acc[qc];
              __syncwarp();
            }
          }
        }
      }
    }
    flush_dtab(cur_head);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int make_token_maps(CUtensorMap* dense, CUtensorMap* plane, CUtensorMap* rows, const void* base, const WinGeom& g, int B,
                    int row_elems, int D, CUtensorMap* colgrp = nullptr) {
  // the tensor as (row_elems, W, H, D, B) for box loads and as (row_elems, B*T) for gather4
  const uint64_t dims5[5] = {static_cast<uint64_t>(row_elems), static_cast<uint64_t>(g.grid[2]), static_cast<uint64_t>(g.grid[1]),
                             static_cast<uint64_t>(g.grid[0]), static_cast<uint64_t>(B)};
  const uint64_t rb = static_cast<uint64_t>(row_elems) * 2;
  const uint64_t str5[4] = {rb, rb * g.grid[2], rb * g.grid[2] * g.grid[1], rb * g.T};
  const uint32_t box_dense[5] = {static_cast<uint32_t>(D), static_cast<uint32_t>(g.win[2]), static_cast<uint32_t>(g.win[1]),
                                 static_cast<uint32_t>(g.win[0]), 1};
  const uint32_t box_plane[5] = {static_cast<uint32_t>(D), static_cast<uint32_t>(g.win[2]), static_cast<uint32_t>(g.win[1]), 1, 1};
  const CUtensorMapSwizzle swz = D == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (dense && make_tmap(dense, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims5, str5, box_dense, swz)) return LCBI_ERR_TENSOR_MAP;
  if (plane && make_tmap(plane, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims5, str5, box_plane, swz)) return LCBI_ERR_TENSOR_MAP;
  // one window column c' (all planes, all rows of the plane): rows land as a' * PH + b' (the backward's W-major key order)
  const uint32_t box_col[5] = {static_cast<uint32_t>(D), 1, static_cast<uint32_t>(g.win[1]), static_cast<uint32_t>(g.win[0]), 1};
  if (colgrp && make_tmap(colgrp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims5, str5, box_col, swz)) return LCBI_ERR_TENSOR_MAP;
  const uint64_t dims2[2] = {static_cast<uint64_t>(row_elems), static_cast<uint64_t>(B) * g.T};
  const uint64_t str2[1] = {rb};
  const uint32_t box2[2] = {static_cast<uint32_t>(D), 1};
  if (rows && make_tmap(rows, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims2, str2, box2, swz)) return LCBI_ERR_TENSOR_MAP;
  return LCBI_OK;
}

template <typename K>
int set_smem_tc(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return LCBI_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  return e == cudaSuccess ? LCBI_OK : set_cuda_error(e);
}

uint32_t align_up_u32(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

}  // namespace

// The tcgen05 path serves 3-D windows that are not clamped on any axis (window == constructor window, so the
// relative-position geometry is the window's own), with 7x7 or 8x8 planes, 2..8 planes (128 < n <= 512), and shifts that
// are all zero or the reference's window // 2 (backbone_swin.py:675,687).
bool win_attn_tc_applicable(const WinParams& p, int head_dim) {
  const WinGeom& g = p.g;
  if (head_dim != 16 && head_dim != 32) return false;
  if (g.grid[0] <= 1 || g.n <= 128 || g.n > 512) return false;
  for (int k = 0; k < 3; ++k) {
    if (g.win[k] != g.ctor[k]) return false;
    if (g.shift[k] != 0 && g.shift[k] != g.win[k] / 2) return false;
  }
  const bool any = g.shift[0] || g.shift[1] || g.shift[2], all = g.shift[0] && g.shift[1] && g.shift[2];
  if (any && !all) return false;
  if (!((g.win[1] == 7 && g.win[2] == 7) || (g.win[1] == 8 && g.win[2] == 8))) return false;
  if (g.win[0] < 2 || g.win[0] > g.win[1]) return false;     // the kernels size their bias tables for depth <= plane height
  if (static_cast<int64_t>(p.B) * g.T >= (1ll << 31) - 8) return false;
  return true;
}

int win_attn_fwd_tc_launch(const WinParams& w, int head_dim, cudaStream_t stream) {
  TcFwdParams p;
  p.w = w;
  const WinGeom& g = w.g;
  p.n_items = w.win_count * w.H;
  if (p.n_items <= 0) return LCBI_OK;
  p.mt = (g.n + kTileM - 1) / kTileM;
  p.npl = g.win[0];
  const uint32_t rb = head_dim * 2;
  p.q_bytes = align_up_u32(static_cast<uint32_t>(p.mt) * kTileM * rb, 1024);
  p.kv_bytes = align_up_u32(static_cast<uint32_t>(p.npl) * kChunk * rb, 1024);
  p.tab_bytes = 0;
  p.buf_bytes = p.q_bytes + 2 * p.kv_bytes;
  const size_t smem = 2 * static_cast<size_t>(p.buf_bytes) + sizeof(FwdBars) + 1024;
  const size_t static_smem = 2 * static_cast<size_t>(2 * g.win[1] - 1) * (2 * g.win[1] - 1) * (2 * g.win[2] - 1) * 4;
  CUtensorMap tm_dense, tm_plane, tm_rows;
  int rc = make_token_maps(&tm_dense, &tm_plane, &tm_rows, w.qkv, g, w.B, 3 * w.C, head_dim);
  if (rc) return rc;
  const int num_sms = current_device_sm_count();
  if (num_sms <= 0) return LCBI_ERR_CUDA;
  const int ctas_per_sm = (smem + static_smem + 1024) * 2 <= 227 * 1024 ? 2 : 1;
  int grid = ctas_per_sm * num_sms;
  if (grid > p.n_items) grid = p.n_items;
  p.items_per_cta = (p.n_items + grid - 1) / grid;
  grid = (p.n_items + p.items_per_cta - 1) / p.items_per_cta;
#define LCBI_TC_FWD(DD, PP)                                                                               \
  do {                                                                                                    \
    if ((rc = set_smem_tc(win_attn_fwd_tc_kernel<DD, PP, PP>, smem))) return rc;                          \
    win_attn_fwd_tc_kernel<DD, PP, PP><<<grid, kFwdThreads, smem, stream>>>(tm_dense, tm_plane, tm_rows, p); \
  } while (0)
  if (head_dim == 16) {
    if (g.win[1] == 7) LCBI_TC_FWD(16, 7); else LCBI_TC_FWD(16, 8);
  } else {
    if (g.win[1] == 7) LCBI_TC_FWD(32, 7); else LCBI_TC_FWD(32, 8);
  }
#undef LCBI_TC_FWD
  return set_cuda_error(cudaGetLastError());
}

int win_attn_bwd_tc_launch(const WinParams& w, int head_dim, cudaStream_t stream) {
  TcBwdParams p;
  p.w = w;
  const WinGeom& g = w.g;
  p.n_items = w.win_count * w.H;
  if (p.n_items <= 0) return LCBI_OK;
  p.kt = (g.n + kTileM - 1) / kTileM;
  p.npl = g.win[0];
  p.gs = g.win[0] * g.win[1];
  p.gsp = (p.gs + 3) / 4 * 4;
  p.n_krows = g.win[2] * p.gsp;
  p.kt = (p.n_krows + kTileM - 1) / kTileM;
  const uint32_t rb = head_dim * 2;
  p.kv_bytes = align_up_u32(static_cast<uint32_t>(p.kt) * kTileM * rb, 1024);
  p.q_bytes = align_up_u32(static_cast<uint32_t>(p.npl) * kChunk * rb, 1024);
  p.aug_bytes = static_cast<uint32_t>(p.npl) * kAugBytes;
  p.buf_bytes = 2 * p.kv_bytes + 2 * p.q_bytes + 2 * p.aug_bytes;
  const size_t tab_max = static_cast<size_t>(2 * g.win[1] - 1) * (2 * g.win[1] - 1) * (2 * g.win[2] - 1) * 4;
  const bool static_dtab = 5 * tab_max <= 46 * 1024;           // must mirror kStaticDtab in the kernel
  p.dtab_stride = static_dtab ? 0 : align_up_u32(static_cast<uint32_t>(g.tab_rows) * 4, 16);
  const size_t static_smem = (static_dtab ? 5 : 1) * tab_max + 3 * 32 * 4 + 64;
  size_t smem = 0;
  for (p.nbuf = 2; p.nbuf >= 1; --p.nbuf) {
    p.off_ones = static_cast<uint32_t>(p.nbuf) * p.buf_bytes;
    p.off_zeros = p.off_ones + 2048;
    p.off_ds = align_up_u32(p.off_zeros + 2048, 1024);
    p.off_dtab = p.off_ds + 2 * kTileM * 128;
    p.off_bars = align_up_u32(p.off_dtab + 4 * p.dtab_stride, 1024);
    smem = static_cast<size_t>(p.off_bars) + sizeof(BwdBars) + 1024;
    if (smem + static_smem <= 227 * 1024) break;
  }
  if (p.nbuf < 1) return LCBI_ERR_UNSUPPORTED;       // the caller falls back to the generic kernels
  CUtensorMap tm_colgrp, tm_plane, tm_rows, tm_do_plane, tm_do_rows;
  int rc = make_token_maps(nullptr, &tm_plane, &tm_rows, w.qkv, g, w.B, 3 * w.C, head_dim, &tm_colgrp);
  if (rc) return rc;
  if ((rc = make_token_maps(nullptr, &tm_do_plane, &tm_do_rows, w.d_out, g, w.B, w.C, head_dim))) return rc;
  const int num_sms = current_device_sm_count();
  if (num_sms <= 0) return LCBI_ERR_CUDA;
  int grid = num_sms < p.n_items ? num_sms : p.n_items;
  p.items_per_cta = (p.n_items + grid - 1) / grid;
  grid = (p.n_items + p.items_per_cta - 1) / p.items_per_cta;
#define LCBI_TC_BWD(DD, PP)                                                                                 \
  do {                                                                                                      \
    if ((rc = set_smem_tc(win_attn_bwd_tc_kernel<DD, PP, PP>, smem))) return rc;                            \
    win_attn_bwd_tc_kernel<DD, PP, PP><<<grid, kBwdThreads, smem, stream>>>(tm_colgrp, tm_plane, tm_rows, tm_do_plane, \
                                                                            tm_do_rows, p);                 \
  } while (0)
  if (head_dim == 16) {
    if (g.win[1] == 7) LCBI_TC_BWD(16, 7); else LCBI_TC_BWD(16, 8);
  } else {
    if (g.win[1] == 7) LCBI_TC_BWD(32, 7); else LCBI_TC_BWD(32, 8);
  }
#undef LCBI_TC_BWD
  return set_cuda_error(cudaGetLastError());
}

}  // namespace lcbi
