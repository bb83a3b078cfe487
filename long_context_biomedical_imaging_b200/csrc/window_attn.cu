// Swin (shifted-)window multi-head attention for sm_100a: fused gather -> QK^T + relative-position bias +
// shift mask -> softmax -> PV -> scatter, forward and backward.
//
// Replaces, without any of their copies, the reference's
//   SwinTransformerBlock.forward_part1   /root/reference/model/models/backbone_swin.py:435-487
//     (F.pad -> torch.roll -> window_partition -> attention -> window_reverse -> torch.roll -> crop)
//   WindowAttention.forward attention core   :339-357   (q*scale, q k^T, + bias[index], + mask, softmax, @ v)
//   compute_mask   :591-628   (never materialised: region ids are evaluated per window slot)
// The per-token qkv / proj Linear layers stay outside (they commute with the gather/scatter): the kernels
// read q/k/v straight out of the (B, T, 3, H, d) qkv Linear output through the closed-form window map, take
// `qkv.bias` as the q/k/v rows of the zero-pad tokens (the reference pads AFTER norm1, so a pad token's qkv is
// exactly the bias, :437-445 + :339), and write the attention output back to (B, T, C) at the source token.
//
// These kernels use warp-level mma.sync tiles: head_dim is 16 or 32 and windows hold 49..512 tokens, so the
// work per (window, head) is far below a tcgen05 tile; the bound is HBM traffic / exp throughput, not the
// tensor pipe ("window attention roofline").
#include <cstdlib>
#include <type_traits>

#include "lcbi_kernels.h"
#include "sm100_ptx.cuh"      // packed fp32 pairs (FFMA2 / FADD2 / FMUL2)
#include "window_common.cuh"

#ifndef LCBI_WIN_NT
#define LCBI_WIN_NT 256   /* threads per CTA of the large-window forward / dK-dV kernels (measured: see DESIGN.md) */
#endif

namespace lcbi {

int win_attn_fwd_small_launch(const WinParams& p, int head_dim, cudaStream_t stream);   // window_attn_small.cu
int win_attn_bwd_small_launch(const WinParams& p, int head_dim, cudaStream_t stream);   // window_attn_small.cu
bool win_attn_tc_applicable(const WinParams& p, int head_dim);                           // window_attn_tc.cu
int window_kernel_mode();                                                                // capi.cu: 0 auto, 1 tcgen05, 2 generic
int win_attn_fwd_tc_launch(const WinParams& p, int head_dim, cudaStream_t stream);       // window_attn_tc.cu
int win_attn_bwd_tc_launch(const WinParams& p, int head_dim, cudaStream_t stream);       // window_attn_tc.cu

namespace {

// -------------------------------------------------------------------------------------------------
// per-CTA window metadata in shared memory
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

struct WinMeta {
  int* tok;        // [n_pad] token index, -1 pad token, -2 slot beyond the window
  int* reg;        // [n_pad] region id
  int* row_term;   // [n_pad]
  int* col_term;   // [n_pad]
};

__device__ __forceinline__ void fill_meta(const WinGeom& g, int w, int n_pad, WinMeta m, int tid, int nthreads) {
  for (int s = tid; s < n_pad; s += nthreads) {
    int tok = -2, region = -1, rt = g.tab_rows - 1, ct = 0;   // rt - ct stays a valid table row for dead slots
    if (s < g.n) {
      slot_lookup(g, w, s, tok, region);
      relpos_terms(g, s, rt, ct);
    }
    m.tok[s] = tok;
    m.reg[s] = region;
    m.row_term[s] = rt * 4;     // BYTE offsets into the per-head bias table: the gather address is then
    m.col_term[s] = ct * 4;     // (table base + row_term) - col_term, one integer add per element
  }
}

// true when the window straddles a shift-mask region boundary (only then do the -100 terms exist); warp-uniform.
// Most windows (all of them in an un-shifted block) are region-uniform and take the cheaper logit path.
__device__ __forceinline__ bool window_has_mask(const WinGeom& g, const WinMeta& m, int tid, int nthreads) {
  bool differs = false;
  const int r0 = m.reg[0];
  for (int s = tid; s < g.n; s += nthreads) differs |= m.reg[s] != r0;
  return __syncthreads_or(differs) != 0;
}

// loads rows [row_begin, row_end) of one of q/k/v (or dO when sel == 3) for (batch b, head h) into a smem tile.
// Four 16-byte gathers are issued per thread before any of them is stored, so a thread pays one global-memory
// latency per four chunks instead of one per chunk (the loop has a runtime trip count and is not unrolled otherwise).
template <int D>
__device__ __forceinline__ void load_rows(uint8_t* tile, const WinParams& p, const int* tok, int b, int h, int sel,
                                          int row_begin, int row_end, int tid, int nthreads) {
  constexpr int kChunks = D / 8;   // 16-byte chunks per row
  constexpr int kBatch = 4;
  const int total = (row_end - row_begin) * kChunks;
  const int64_t plane = sel < 3 ? static_cast<int64_t>(3) * p.C : p.C;           // elements per token
  const __nv_bfloat16* base = (sel < 3 ? p.qkv + sel * p.C : p.d_out) + static_cast<int64_t>(b) * p.g.T * plane + h * D;
  for (int e0 = tid; e0 < total; e0 += kBatch * nthreads) {
    uint4 val[kBatch];
    int tk[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = e0 + u * nthreads;
      val[u] = make_uint4(0u, 0u, 0u, 0u);
      tk[u] = -2;
      if (e < total) {
        const int r = row_begin + e / kChunks, c = e % kChunks;
        tk[u] = tok[r];
        if (tk[u] >= 0) val[u] = __ldg(reinterpret_cast<const uint4*>(base + tk[u] * plane + c * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int e = e0 + u * nthreads;
      if (e >= total) continue;
      const int r = row_begin + e / kChunks, c = e % kChunks;
      if (tk[u] == -1 && sel < 3 && p.qkv_bias != nullptr) {
        const float* bsrc = p.qkv_bias + sel * p.C + h * D + c * 8;
        val[u].x = pack2_bf16(bsrc[0], bsrc[1]);
        val[u].y = pack2_bf16(bsrc[2], bsrc[3]);
        val[u].z = pack2_bf16(bsrc[4], bsrc[5]);
        val[u].w = pack2_bf16(bsrc[6], bsrc[7]);
      }
      *reinterpret_cast<uint4*>(tile + r * Tile<D>::kStride + c * 16) = val[u];
    }
  }
}

// =================================================================================================
// forward: CTA = (window, head), 4 warps, each warp owns 16-row query tiles
// =================================================================================================
// kHighOcc: cap registers for 6 (D=32: a few spills) CTAs per SM. Measured on B200: a win when the window is one
// key chunk (n <= 64: the kernel is latency-bound, 235 -> 192 us at cfg2 stage 1) and for d = 16; a loss for
// d = 32 with 343-token windows, where the spills land in the six-iteration key loop.
template <int D, bool kHighOcc, int NT = 128>
__global__ void __launch_bounds__(NT, kHighOcc ? (NT == 128 ? 6 : (NT == 256 ? 3 : 2)) : (NT == 256 ? 2 : 1))
win_attn_fwd_kernel(const WinParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const WinGeom& g = p.g;
  const int n = g.n, n_pad = round_up(n, kKeyChunk);
  constexpr int kStride = Tile<D>::kStride;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + n_pad * kStride;
  uint8_t* sV = sK + n_pad * kStride;
  float* tab = reinterpret_cast<float*>(sV + n_pad * kStride);
  WinMeta meta;
  meta.tok = reinterpret_cast<int*>(tab + round_up(g.tab_rows, 4));
  meta.reg = meta.tok + n_pad;
  meta.row_term = meta.reg + n_pad;
  meta.col_term = meta.row_term + n_pad;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // heads fastest: the H CTAs of one window are scheduled together, so the 32/64-byte head slices of a token row are
  // fetched from DRAM once (window-fastest order read 317 MB for 100 MB of operands at cfg4 stage 1)
  const int h = blockIdx.x % p.H, wg = p.win_begin + blockIdx.x / p.H;
  int b, w;
  fdivmod(wg, g.d_nW, b, w);

  fill_meta(g, w, n_pad, meta, tid, NT);
  for (int t = tid; t < g.tab_rows; t += NT) tab[t] = p.table[t * p.H + h] * kLog2e;
  __syncthreads();
  const bool has_mask = window_has_mask(g, meta, tid, NT);
  load_rows<D>(sQ, p, meta.tok, b, h, 0, 0, n_pad, tid, NT);
  load_rows<D>(sK, p, meta.tok, b, h, 1, 0, n_pad, tid, NT);
  load_rows<D>(sV, p, meta.tok, b, h, 2, 0, n_pad, tid, NT);
  __syncthreads();

  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
  const int gq = lane >> 2, qq = lane & 3;
  const float mask_log2 = -100.0f * kLog2e;
  const int n_qt = (n + 15) / 16;

  for (int qt = warp; qt < n_qt; qt += NT / 32) {
    const int row0 = qt * 16;
    uint32_t aq[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) load_a_frag<kStride>(aq[kk], q_base, row0, kk * 16, lane);
    const int i0 = row0 + gq, i1 = i0 + 8;
    const uint32_t tab_base = static_cast<uint32_t>(__cvta_generic_to_shared(tab));
    const uint32_t tb0 = tab_base + meta.row_term[i0], tb1 = tab_base + meta.row_term[i1];   // table row addresses
    const int rg0 = meta.reg[i0], rg1 = meta.reg[i1];
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float oacc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;

    for (int kc = 0; kc < n_pad; kc += kKeyChunk) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t b0, b1;
          load_b_frag_nt<kStride>(b0, b1, k_base, kc + nt * 8, kk * 16, lane);
          mma_bf16_16816(s[nt], aq[kk], b0, b1);
        }
      }
      // logits (log2 domain) = s*scale*log2e + bias*log2e + mask*log2e ; columns beyond the window -> -inf.
      // Dead columns carry col_term 0, so the table gather stays in range without a bounds check; only the last key
      // chunk can hold them, and only windows that straddle a region boundary need the mask term.
      float cmax0 = -INFINITY, cmax1 = -INFINITY;
      const bool tail = kc + kKeyChunk > n;
      auto logits = [&](auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int j = kc + nt * 8 + qq * 2;
          const int2 ct = *reinterpret_cast<const int2*>(meta.col_term + j);
          int2 rj = make_int2(0, 0);
          if (kMasked) rj = *reinterpret_cast<const int2*>(meta.reg + j);
          float v00 = fmaf(s[nt][0], p.scale_log2, lds_f32(tb0 - ct.x)), v01 = fmaf(s[nt][1], p.scale_log2, lds_f32(tb0 - ct.y));
          float v10 = fmaf(s[nt][2], p.scale_log2, lds_f32(tb1 - ct.x)), v11 = fmaf(s[nt][3], p.scale_log2, lds_f32(tb1 - ct.y));
          if (kMasked) {
            v00 += rg0 != rj.x ? mask_log2 : 0.f; v01 += rg0 != rj.y ? mask_log2 : 0.f;
            v10 += rg1 != rj.x ? mask_log2 : 0.f; v11 += rg1 != rj.y ? mask_log2 : 0.f;
          }
          if (tail) {
            if (j >= n) v00 = v10 = -INFINITY;
            if (j + 1 >= n) v01 = v11 = -INFINITY;
          }
          s[nt][0] = v00; s[nt][1] = v01; s[nt][2] = v10; s[nt][3] = v11;
          cmax0 = fmaxf(cmax0, fmaxf(v00, v01));
          cmax1 = fmaxf(cmax1, fmaxf(v10, v11));
        }
      };
      if (has_mask) logits(std::true_type{}); else logits(std::false_type{});
      cmax0 = fmaxf(cmax0, __shfl_xor_sync(0xffffffffu, cmax0, 1));
      cmax0 = fmaxf(cmax0, __shfl_xor_sync(0xffffffffu, cmax0, 2));
      cmax1 = fmaxf(cmax1, __shfl_xor_sync(0xffffffffu, cmax1, 1));
      cmax1 = fmaxf(cmax1, __shfl_xor_sync(0xffffffffu, cmax1, 2));
      const float mn0 = fmaxf(m0, cmax0), mn1 = fmaxf(m1, cmax1);   // finite: every chunk has >= 1 valid column
      const float a0 = ex2f(m0 - mn0), a1 = ex2f(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        oacc[i][0] *= a0; oacc[i][1] *= a0; oacc[i][2] *= a1; oacc[i][3] *= a1;
      }
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = ex2f(s[nt][0] - m0); s[nt][1] = ex2f(s[nt][1] - m0);
        s[nt][2] = ex2f(s[nt][2] - m1); s[nt][3] = ex2f(s[nt][3] - m1);
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l0 += rs0; l1 += rs1;
      // O += P V
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {   // 16 keys per step
        uint32_t ap[4];
        ap[0] = pack2_bf16(s[2 * kb][0], s[2 * kb][1]);
        ap[1] = pack2_bf16(s[2 * kb][2], s[2 * kb][3]);
        ap[2] = pack2_bf16(s[2 * kb + 1][0], s[2 * kb + 1][1]);
        ap[3] = pack2_bf16(s[2 * kb + 1][2], s[2 * kb + 1][3]);
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) {
          uint32_t b0, b1;
          load_b_frag_t<kStride>(b0, b1, v_base, kc + kb * 16, nd * 8, lane);
          mma_bf16_16816(oacc[nd], ap, b0, b1);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    // stage the 16 x D output tile in this warp's (already consumed) Q rows, then 16-byte scatter stores
    __syncwarp();
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) {
      *reinterpret_cast<uint32_t*>(sQ + i0 * kStride + (nd * 8 + qq * 2) * 2) =
          pack2_bf16(oacc[nd][0] * inv0, oacc[nd][1] * inv0);
      *reinterpret_cast<uint32_t*>(sQ + i1 * kStride + (nd * 8 + qq * 2) * 2) =
          pack2_bf16(oacc[nd][2] * inv1, oacc[nd][3] * inv1);
    }
    const int t0 = meta.tok[i0], t1 = meta.tok[i1];
    if (qq == 0) {
      if (t0 >= 0) p.lse2[(static_cast<int64_t>(b) * g.T + t0) * p.H + h] = m0 + log2f(l0);
      if (t1 >= 0) p.lse2[(static_cast<int64_t>(b) * g.T + t1) * p.H + h] = m1 + log2f(l1);
    }
    __syncwarp();
    constexpr int kChunks = D / 8;
    for (int e = lane; e < 16 * kChunks; e += 32) {
      const int r = row0 + e / kChunks, c = e % kChunks;
      const int t = meta.tok[r];
      if (t >= 0)
        *reinterpret_cast<uint4*>(p.out + (static_cast<int64_t>(b) * g.T + t) * p.C + h * D + c * 8) =
            *reinterpret_cast<const uint4*>(sQ + r * kStride + c * 16);
    }
    __syncwarp();
  }
}

size_t fwd_smem_bytes(const WinGeom& g, int D) {
  const int n_pad = round_up(g.n, kKeyChunk);
  return static_cast<size_t>(3) * n_pad * (D * 2 + 16) + static_cast<size_t>(round_up(g.tab_rows, 4)) * 4 + static_cast<size_t>(n_pad) * 16;
}

// =================================================================================================
// backward prep: dsum[b,t,h] = sum_d dO*O (one warp handles 32/(D/8)... simple: one thread per (b,t,h))
// =================================================================================================
template <int D>
__global__ void win_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                    float* __restrict__ dsum, int64_t total, int H, int C) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int h = static_cast<int>(idx % H);
  const int64_t bt = idx / H;
  const uint4* po = reinterpret_cast<const uint4*>(o + bt * C + h * D);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + bt * C + h * D);
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < D / 8; ++c) {
    const uint4 a = po[c], bq = pd[c];
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bq);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = __bfloat1622float2(a2[i]), y = __bfloat1622float2(b2[i]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  dsum[idx] = acc;
}

// per-row backward scalars in smem: lse2 (+inf for pad/out-of-window rows => P = 0) and dsum
__device__ __forceinline__ void load_row_scalars(const WinParams& p, const int* tok, int b, int h, int row_begin,
                                                 int row_end, float* s_lse, float* s_dsum, int tid, int nthreads) {
  for (int r = row_begin + tid; r < row_end; r += nthreads) {
    const int t = tok[r];
    float l = INFINITY, d = 0.f;
    if (t >= 0) {
      const int64_t idx = (static_cast<int64_t>(b) * p.g.T + t) * p.H + h;
      l = p.lse2[idx];
      d = p.dsum[idx];
    }
    s_lse[r] = l;
    s_dsum[r] = d;
  }
}

// =================================================================================================
// backward dK/dV: CTA = (window, head), 4 warps, each warp owns 16-row KEY tiles and sweeps the queries
//   S^T = K Q^T, P^T = exp2(.), dP^T = V dO^T, dS^T = P^T o (dP^T - D[q]); dV += P^T dO; dK += dS^T Q
// =================================================================================================
template <int D, bool kHighOcc, int NT = 128>
__global__ void __launch_bounds__(NT, kHighOcc ? (NT == 128 ? 5 : (NT == 256 ? 2 : 1)) : 1)
win_attn_bwd_dkdv_kernel(const WinParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const WinGeom& g = p.g;
  const int n = g.n, n_pad = round_up(n, kKeyChunk);
  constexpr int kStride = Tile<D>::kStride;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + n_pad * kStride;
  uint8_t* sV = sK + n_pad * kStride;
  uint8_t* sDO = sV + n_pad * kStride;
  float* tab = reinterpret_cast<float*>(sDO + n_pad * kStride);
  float* s_lse = tab + round_up(g.tab_rows, 4);
  float* s_dsum = s_lse + n_pad;
  WinMeta meta;
  meta.tok = reinterpret_cast<int*>(s_dsum + n_pad);
  meta.reg = meta.tok + n_pad;
  meta.row_term = meta.reg + n_pad;
  meta.col_term = meta.row_term + n_pad;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // heads fastest: the H CTAs of one window are scheduled together, so the 32/64-byte head slices of a token row are
  // fetched from DRAM once (window-fastest order read 317 MB for 100 MB of operands at cfg4 stage 1)
  const int h = blockIdx.x % p.H, wg = p.win_begin + blockIdx.x / p.H;
  int b, w;
  fdivmod(wg, g.d_nW, b, w);

  fill_meta(g, w, n_pad, meta, tid, NT);
  for (int t = tid; t < g.tab_rows; t += NT) tab[t] = p.table[t * p.H + h] * kLog2e;
  __syncthreads();
  load_rows<D>(sQ, p, meta.tok, b, h, 0, 0, n_pad, tid, NT);
  load_rows<D>(sK, p, meta.tok, b, h, 1, 0, n_pad, tid, NT);
  load_rows<D>(sV, p, meta.tok, b, h, 2, 0, n_pad, tid, NT);
  load_rows<D>(sDO, p, meta.tok, b, h, 3, 0, n_pad, tid, NT);
  load_row_scalars(p, meta.tok, b, h, 0, n_pad, s_lse, s_dsum, tid, NT);
  __syncthreads();
  const bool has_mask = window_has_mask(g, meta, tid, NT);
  // per-query column data in one 16-byte record {lse2, dsum, row_term, region}: one LDS.128 per column below
  // (own array at the end of the dynamic smem, see dkdv_smem_bytes)
  float4* qcol = reinterpret_cast<float4*>(meta.col_term + n_pad);
  for (int i = tid; i < n_pad; i += NT)
    qcol[i] = make_float4(s_lse[i], s_dsum[i], __int_as_float(meta.row_term[i]), __int_as_float(meta.reg[i]));
  __syncthreads();

  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
  const uint32_t do_base = static_cast<uint32_t>(__cvta_generic_to_shared(sDO));
  const int gq = lane >> 2, qq = lane & 3;
  const float mask_log2 = -100.0f * kLog2e;
  const int n_kt = (n + 15) / 16;
  const float scale = p.scale_log2 / kLog2e;
  float pad_dk[D / 8][2], pad_dv[D / 8][2];   // gradient reaching qkv.bias through pad-token keys/values
#pragma unroll
  for (int i = 0; i < D / 8; ++i) pad_dk[i][0] = pad_dk[i][1] = pad_dv[i][0] = pad_dv[i][1] = 0.f;

  for (int kt = warp; kt < n_kt; kt += NT / 32) {
    const int key0 = kt * 16;
    uint32_t ak[D / 16][4], av[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      load_a_frag<kStride>(ak[kk], k_base, key0, kk * 16, lane);
      load_a_frag<kStride>(av[kk], v_base, key0, kk * 16, lane);
    }
    const int j0 = key0 + gq, j1 = j0 + 8;           // key rows held by this thread
    const uint32_t tab_base = static_cast<uint32_t>(__cvta_generic_to_shared(tab));
    const uint32_t tc0 = tab_base - meta.col_term[j0], tc1 = tab_base - meta.col_term[j1];
    const int rg0 = meta.reg[j0], rg1 = meta.reg[j1];
    const bool kv0 = j0 < n, kv1 = j1 < n;
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int qc = 0; qc < n_pad; qc += 32) {   // 32 queries per step (4 n-tiles)
      float st[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t b0, b1;
          load_b_frag_nt<kStride>(b0, b1, q_base, qc + nt * 8, kk * 16, lane);
          mma_bf16_16816(st[nt], ak[kk], b0, b1);
          load_b_frag_nt<kStride>(b0, b1, do_base, qc + nt * 8, kk * 16, lane);
          mma_bf16_16816(dp[nt], av[kk], b0, b1);
        }
      }
      auto probs = [&](auto masked_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float4 qc4 = qcol[qc + nt * 8 + qq * 2 + e];    // query (column of the transposed tile): lse2, dsum, rt, region
            const int rt = __float_as_int(qc4.z);
            float l0 = fmaf(st[nt][e], p.scale_log2, lds_f32(tc0 + rt)) - qc4.x;
            float l1 = fmaf(st[nt][2 + e], p.scale_log2, lds_f32(tc1 + rt)) - qc4.x;
            if (kMasked) {
              const int ri = __float_as_int(qc4.w);
              l0 += ri != rg0 ? mask_log2 : 0.f;
              l1 += ri != rg1 ? mask_log2 : 0.f;
            }
            const float p0 = kv0 ? ex2f(l0) : 0.f, p1 = kv1 ? ex2f(l1) : 0.f;
            st[nt][e] = p0;
            st[nt][2 + e] = p1;
            dp[nt][e] = p0 * (dp[nt][e] - qc4.y);
            dp[nt][2 + e] = p1 * (dp[nt][2 + e] - qc4.y);
          }
        }
      };
      if (has_mask) probs(std::true_type{}); else probs(std::false_type{});
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {   // 16 queries per MMA k-step
        uint32_t ap[4], ads[4];
        ap[0] = pack2_bf16(st[2 * kb][0], st[2 * kb][1]);
        ap[1] = pack2_bf16(st[2 * kb][2], st[2 * kb][3]);
        ap[2] = pack2_bf16(st[2 * kb + 1][0], st[2 * kb + 1][1]);
        ap[3] = pack2_bf16(st[2 * kb + 1][2], st[2 * kb + 1][3]);
        ads[0] = pack2_bf16(dp[2 * kb][0], dp[2 * kb][1]);
        ads[1] = pack2_bf16(dp[2 * kb][2], dp[2 * kb][3]);
        ads[2] = pack2_bf16(dp[2 * kb + 1][0], dp[2 * kb + 1][1]);
        ads[3] = pack2_bf16(dp[2 * kb + 1][2], dp[2 * kb + 1][3]);
#pragma unroll
        for (int nd = 0; nd < D / 8; ++nd) {
          uint32_t b0, b1;
          load_b_frag_t<kStride>(b0, b1, do_base, qc + kb * 16, nd * 8, lane);
          mma_bf16_16816(dv[nd], ap, b0, b1);
          load_b_frag_t<kStride>(b0, b1, q_base, qc + kb * 16, nd * 8, lane);
          mma_bf16_16816(dk[nd], ads, b0, b1);
        }
      }
    }
    // write dK (scaled) / dV rows of real tokens; pad-token rows feed the qkv.bias gradient
    const int t0 = meta.tok[j0], t1 = meta.tok[j1];
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
  if (p.dbias_pad != nullptr && g.n * g.nW != g.T) {   // the window grid contains pad tokens
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
          const int col = h * D + nd * 8 + qq * 2 + e;
          if (a != 0.f) atomicAdd(p.dbias_pad + p.C + col, a);
          if (c != 0.f) atomicAdd(p.dbias_pad + 2 * p.C + col, c);
        }
      }
    }
  }
}

size_t dkdv_smem_bytes(const WinGeom& g, int D) {
  const int n_pad = round_up(g.n, kKeyChunk);
  return static_cast<size_t>(4) * n_pad * (D * 2 + 16) + static_cast<size_t>(round_up(g.tab_rows, 4)) * 4 +
         static_cast<size_t>(n_pad) * 8 + static_cast<size_t>(n_pad) * 16 +
         static_cast<size_t>(n_pad) * 16;   // lse/dsum, the four meta arrays, the packed per-query records
}

// =================================================================================================
// backward dQ + d(relative_position_bias_table):
//   CTA = (head, 32-row query slab, window subset); warp (qt, ks) = 16-row query tile x 128-key split.
//   Loops over its windows keeping dBias[32 x n] for the slab in registers; per window dQ = dS K.
// =================================================================================================
constexpr int kKeySplit = 128;

// QT = 16-row query tiles per CTA (slab = 16*QT rows), MAXKS = max number of 128-key splits (block = 32*QT*n_ks)
template <int D, int QT, int MAXKS>
__global__ void __launch_bounds__(32 * QT * MAXKS)
win_attn_bwd_dq_kernel(const WinParams p) {
  constexpr int kSlabRows = 16 * QT;
  extern __shared__ __align__(16) uint8_t smem[];
  const WinGeom& g = p.g;
  const int n = g.n, n_pad = round_up(n, kKeySplit);
  constexpr int kStride = Tile<D>::kStride;
  const int n_ks = n_pad / kKeySplit;
  const int nthreads = blockDim.x;
  // Per-window data is double-buffered: [K, V (n_pad rows), Q, dO (slab rows)] tiles and [tok, reg (n_pad), lse, dsum
  // (slab)] metadata of the NEXT window are gathered with cp.async while the current window is processed.
  const int tile_set_bytes = (2 * n_pad + 2 * kSlabRows) * kStride;
  uint8_t* tiles = smem;                                          // [2][tile_set_bytes]
  float* tab = reinterpret_cast<float*>(tiles + 2 * tile_set_bytes);
  float* s_dq = tab + round_up(g.tab_rows, 4);                    // [n_ks][slab][D] fp32 partial dQ
  int* s_rowterm = reinterpret_cast<int*>(s_dq + n_ks * kSlabRows * D);   // [n_pad] window-independent (bytes)
  int* s_colterm = s_rowterm + n_pad;                             // [n_pad]
  int* s_wmeta = s_colterm + n_pad;                               // [2][2 * n_pad + 2 * slab]: tok, reg, lse, dsum
  const int wmeta_ints = 2 * n_pad + 2 * kSlabRows;
  int* s_flags = s_wmeta + 2 * wmeta_ints;                        // [2]: bit 0 pad tokens present, bit 1 mask needed
  const uint32_t tiles_base = static_cast<uint32_t>(__cvta_generic_to_shared(tiles));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qt = warp % QT, ks = warp / QT;
  const int slab = blockIdx.x, h = blockIdx.y, split = blockIdx.z;
  const int row_base = slab * kSlabRows;          // first window slot of this slab
  const int gq = lane >> 2, qq = lane & 3;
  const float mask_log2 = -100.0f * kLog2e;
  const float scale = p.scale_log2 / kLog2e;

  for (int t = tid; t < g.tab_rows; t += nthreads) tab[t] = p.table[t * p.H + h] * kLog2e;
  for (int s0 = tid; s0 < n_pad; s0 += nthreads) {        // relative-position terms do not depend on the window
    int rt = g.tab_rows - 1, ct = 0;                      // rt - ct stays a valid table row for dead slots
    if (s0 < n) relpos_terms(g, s0, rt, ct);
    s_rowterm[s0] = rt * 4;
    s_colterm[s0] = ct * 4;
  }
  if (tid < 2) s_flags[tid] = 0;

  float dbias[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) dbias[i][0] = dbias[i][1] = dbias[i][2] = dbias[i][3] = 0.f;

  const int total_windows = p.win_count;
  constexpr int kChunks = D / 8;

  // slot -> token / region of window wg into metadata buffer `buf`; lse / dsum of the slab rows by 4-byte cp.async
  auto window_meta = [&](int wg, int buf) {
    int b, w, tok0, reg0;
    fdivmod(wg, g.d_nW, b, w);
    slot_lookup(g, w, 0, tok0, reg0);
    int* wm = s_wmeta + buf * wmeta_ints;
    int fl = 0;
    for (int s0 = tid; s0 < n_pad; s0 += nthreads) {
      int tok = -2, reg = -1;
      if (s0 < n) {
        slot_lookup(g, w, s0, tok, reg);
        if (tok == -1) fl |= 1;
        if (reg != reg0) fl |= 2;
      }
      wm[s0] = tok;
      wm[n_pad + s0] = reg;
      if (s0 >= row_base && s0 < row_base + kSlabRows) {
        const int r = s0 - row_base;
        float* lse_dst = reinterpret_cast<float*>(wm + 2 * n_pad) + r;
        float* ds_dst = lse_dst + kSlabRows;
        if (tok >= 0) {
          const int64_t idx = (static_cast<int64_t>(b) * g.T + tok) * p.H + h;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(lse_dst))),
                       "l"(p.lse2 + idx) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(ds_dst))),
                       "l"(p.dsum + idx) : "memory");
        } else {
          *lse_dst = INFINITY;                      // pad / dead query rows: P = exp2(. - inf) = 0
          *ds_dst = 0.f;
        }
      }
    }
    if (fl) atomicOr(&s_flags[buf], fl);
  };
  // cp.async of the window's K / V rows and the slab's Q / dO rows (zero fill for pad and dead rows)
  auto issue_tiles = [&](int wg, int buf) {
    int b, w;
    fdivmod(wg, g.d_nW, b, w);
    const int64_t tok_base = static_cast<int64_t>(b) * g.T;
    const int* tk = s_wmeta + buf * wmeta_ints;
    const uint32_t set_base = tiles_base + buf * tile_set_bytes;
    const int total = (2 * n_pad + 2 * kSlabRows) * kChunks;
    for (int e = tid; e < total; e += nthreads) {
      const int row = e / kChunks, c = e % kChunks;        // row of the tile set: K rows, V rows, Q slab, dO slab
      int slot, sel;
      if (row < n_pad) { slot = row; sel = 1; }
      else if (row < 2 * n_pad) { slot = row - n_pad; sel = 2; }
      else if (row < 2 * n_pad + kSlabRows) { slot = row_base + row - 2 * n_pad; sel = 0; }
      else { slot = row_base + row - 2 * n_pad - kSlabRows; sel = 3; }
      const int t = tk[slot];
      const __nv_bfloat16* src = p.qkv;                    // any valid address when nothing is read (zero fill)
      if (t >= 0)
        src = sel < 3 ? p.qkv + ((tok_base + t) * 3 + sel) * p.C + h * D + c * 8
                      : p.d_out + (tok_base + t) * p.C + h * D + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(set_base + row * kStride + c * 16), "l"(src),
                   "r"(t >= 0 ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int wg_first = p.win_begin + split;
  const int wg_end = p.win_begin + total_windows;
  __syncthreads();                                  // table, terms and flags written
  if (wg_first < wg_end) {
    window_meta(wg_first, 0);
    __syncthreads();
    issue_tiles(wg_first, 0);
  }

  int it = 0;
  for (int wg = wg_first; wg < wg_end; wg += p.win_splits, ++it) {
    const int cur = it & 1, nxt = cur ^ 1;
    const int wg_next = wg + p.win_splits;
    int b, w;
    fdivmod(wg, g.d_nW, b, w);
    const int* m_tok = s_wmeta + cur * wmeta_ints;
    const int* m_reg = m_tok + n_pad;
    const float* s_lse = reinterpret_cast<const float*>(m_tok + 2 * n_pad);
    const float* s_dsum = s_lse + kSlabRows;
    uint8_t* sK = tiles + cur * tile_set_bytes;
    const uint32_t k_base = tiles_base + cur * tile_set_bytes, v_base = k_base + n_pad * kStride;
    const uint32_t q_base = v_base + n_pad * kStride, do_base = q_base + kSlabRows * kStride;

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int flags = s_flags[cur];
    if ((flags & 1) && p.qkv_bias != nullptr) {
      // pad tokens enter the reference as zeros before the qkv Linear, so their k / v rows are the Linear's bias
      for (int e = tid; e < 2 * n_pad * kChunks; e += nthreads) {
        const int row = e / kChunks, c = e % kChunks;
        const int slot = row < n_pad ? row : row - n_pad, sel = row < n_pad ? 1 : 2;
        if (m_tok[slot] != -1) continue;
        const float* bsrc = p.qkv_bias + sel * p.C + h * D + c * 8;
        uint4 val;
        val.x = pack2_bf16(bsrc[0], bsrc[1]);
        val.y = pack2_bf16(bsrc[2], bsrc[3]);
        val.z = pack2_bf16(bsrc[4], bsrc[5]);
        val.w = pack2_bf16(bsrc[6], bsrc[7]);
        *reinterpret_cast<uint4*>(sK + row * kStride + c * 16) = val;
      }
    }
    if (tid == 0) s_flags[nxt] = 0;                 // (last read one iteration ago; set again after the barrier)
    __syncthreads();                                // this window's data is in; the previous window is fully consumed
    if (wg_next < wg_end) {
      window_meta(wg_next, nxt);
      __syncthreads();
      issue_tiles(wg_next, nxt);                    // lands while this window is processed
    }
    const bool has_mask = (flags & 2) != 0;

    uint32_t aq[D / 16][4], ado[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      load_a_frag<kStride>(aq[kk], q_base, qt * 16, kk * 16, lane);
      load_a_frag<kStride>(ado[kk], do_base, qt * 16, kk * 16, lane);
    }
    const int r0 = qt * 16 + gq, r1 = r0 + 8;                 // rows inside the slab
    const int i0 = row_base + r0, i1 = row_base + r1;         // window slots
    const float lse0 = s_lse[r0], lse1 = s_lse[r1], ds0 = s_dsum[r0], ds1 = s_dsum[r1];
    const uint32_t tab_base = static_cast<uint32_t>(__cvta_generic_to_shared(tab));
    const uint32_t tb0 = tab_base + s_rowterm[i0], tb1 = tab_base + s_rowterm[i1];
    const int rg0 = m_reg[i0], rg1 = m_reg[i1];
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

    const int n_sub = min(4, (n - ks * kKeySplit + 31) / 32);   // 32-key steps of this split that hold real keys
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {            // 32 keys per step inside this warp's 128-key split
      if (sub >= n_sub) break;
      const int key0 = ks * kKeySplit + sub * 32;
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t b0, b1;
          load_b_frag_nt<kStride>(b0, b1, k_base, key0 + nt * 8, kk * 16, lane);
          mma_bf16_16816(s[nt], aq[kk], b0, b1);
          load_b_frag_nt<kStride>(b0, b1, v_base, key0 + nt * 8, kk * 16, lane);
          mma_bf16_16816(dp[nt], ado[kk], b0, b1);
        }
      }
      const bool tail = key0 + 32 > n;           // only then can a key column be a dead slot
      const float2 neg_lse = make_float2(-lse0, -lse1), neg_ds = make_float2(-ds0, -ds1);
      // the dead-slot test of a key column is compiled only into the one 32-key step that holds the window's last keys
      auto grads = [&](auto masked_c, auto tail_c) {
        constexpr bool kMasked = decltype(masked_c)::value;
        constexpr bool kTail = decltype(tail_c)::value;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int j = key0 + nt * 8 + qq * 2;
          const int2 ct = *reinterpret_cast<const int2*>(s_colterm + j);
          int2 rj = make_int2(0, 0);
          if (kMasked) rj = *reinterpret_cast<const int2*>(m_reg + j);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            // rows i0 / i1 of this column as one packed fp32 pair: the per-logit arithmetic around the two exponentials
            // is five FFMA2 / FADD2 / FMUL2 instead of ten scalar instructions (the kernel is issue-bound: ~27
            // instructions per logit, 51 % issue-active at cfg4 stage 1)
            const int cte = e ? ct.y : ct.x;
            float2 l = ffma2(make_float2(s[nt][e], s[nt][2 + e]), make_float2(p.scale_log2, p.scale_log2),
                             make_float2(lds_f32(tb0 - cte), lds_f32(tb1 - cte)));
            l = fadd2(l, neg_lse);
            if (kMasked) {
              const int rje = e ? rj.y : rj.x;
              l = fadd2(l, make_float2(rg0 != rje ? mask_log2 : 0.f, rg1 != rje ? mask_log2 : 0.f));
            }
            float2 d = fmul2(make_float2(ex2f(l.x), ex2f(l.y)), fadd2(make_float2(dp[nt][e], dp[nt][2 + e]), neg_ds));
            if (kTail && j + e >= n) d = make_float2(0.f, 0.f);
            dp[nt][e] = d.x;
            dp[nt][2 + e] = d.y;
            const float2 db = fadd2(make_float2(dbias[sub * 4 + nt][e], dbias[sub * 4 + nt][2 + e]), d);
            dbias[sub * 4 + nt][e] = db.x;
            dbias[sub * 4 + nt][2 + e] = db.y;
          }
        }
      };
      if (tail) {
        if (has_mask) grads(std::true_type{}, std::true_type{}); else grads(std::false_type{}, std::true_type{});
      } else {
        if (has_mask) grads(std::true_type{}, std::false_type{}); else grads(std::false_type{}, std::false_type{});
      }
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
          load_b_frag_t<kStride>(b0, b1, k_base, key0 + kb * 16, nd * 8, lane);
          mma_bf16_16816(dq[nd], ads, b0, b1);
        }
      }
    }
    // reduce the key-split partials of dQ through shared memory, then scatter rows of real tokens
    float* mine = s_dq + (ks * kSlabRows) * D;
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd) {
      *reinterpret_cast<float2*>(mine + r0 * D + nd * 8 + qq * 2) = make_float2(dq[nd][0], dq[nd][1]);
      *reinterpret_cast<float2*>(mine + r1 * D + nd * 8 + qq * 2) = make_float2(dq[nd][2], dq[nd][3]);
    }
    __syncthreads();
    for (int e = tid; e < kSlabRows * (D / 8); e += nthreads) {
      const int r = e / (D / 8), c = e % (D / 8);
      const int t = m_tok[row_base + r];
      if (t < 0) continue;
      float acc[8] = {};
      for (int s2 = 0; s2 < n_ks; ++s2) {
        const float4 x = *reinterpret_cast<const float4*>(s_dq + (s2 * kSlabRows + r) * D + c * 8);
        const float4 y = *reinterpret_cast<const float4*>(s_dq + (s2 * kSlabRows + r) * D + c * 8 + 4);
        acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
        acc[4] += y.x; acc[5] += y.y; acc[6] += y.z; acc[7] += y.w;
      }
      uint4 val;
      val.x = pack2_bf16(acc[0] * scale, acc[1] * scale);
      val.y = pack2_bf16(acc[2] * scale, acc[3] * scale);
      val.z = pack2_bf16(acc[4] * scale, acc[5] * scale);
      val.w = pack2_bf16(acc[6] * scale, acc[7] * scale);
      *reinterpret_cast<uint4*>(p.dqkv + (static_cast<int64_t>(b) * g.T + t) * 3 * p.C + h * D + c * 8) = val;
    }
  }

  // d(relative_position_bias_table)[idx(i,j), h] += dBias[i, j]   (window-independent index).
  // Folded in shared memory first (the bias table copy is dead by now): every CTA of a head targets the same few
  // hundred cache lines, and per-(i, j) global adds serialise in the L2 (see window_attn_small.cu).
  if (p.dtable != nullptr) {
    __syncthreads();                                 // all warps are past their last read of tab
    for (int t = tid; t < g.tab_rows; t += nthreads) tab[t] = 0.f;
    __syncthreads();
    int rt[2], ct_dummy;
    const int i0 = row_base + qt * 16 + gq, i1 = i0 + 8;
    relpos_terms(g, i0 < n ? i0 : 0, rt[0], ct_dummy);
    relpos_terms(g, i1 < n ? i1 : 0, rt[1], ct_dummy);
#pragma unroll
    for (int t16 = 0; t16 < 16; ++t16) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = ks * kKeySplit + t16 * 8 + qq * 2 + e;
        if (j >= n) continue;
        int rtj, ctj;
        relpos_terms(g, j, rtj, ctj);
        if (i0 < n) atomicAdd(tab + (rt[0] - ctj), dbias[t16][e]);
        if (i1 < n) atomicAdd(tab + (rt[1] - ctj), dbias[t16][2 + e]);
      }
    }
    __syncthreads();
    for (int t = tid; t < g.tab_rows; t += nthreads) {
      const float v = tab[t];
      if (v != 0.f) atomicAdd(p.dtable + static_cast<int64_t>(t) * p.H + h, v);
    }
  }
}

size_t dq_smem_bytes(const WinGeom& g, int D, int slab_rows) {
  const int n_pad = round_up(g.n, kKeySplit);
  const int n_ks = n_pad / kKeySplit;
  const size_t tile_set = static_cast<size_t>(2 * n_pad + 2 * slab_rows) * (D * 2 + 16);
  return 2 * tile_set + static_cast<size_t>(round_up(g.tab_rows, 4)) * 4 + static_cast<size_t>(n_ks) * slab_rows * D * 4 +
         static_cast<size_t>(2) * n_pad * 4 + static_cast<size_t>(2) * (2 * n_pad + 2 * slab_rows) * 4 + 16;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 227 * 1024) return LCBI_ERR_UNSUPPORTED;
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  return LCBI_OK;
}

int fill_params(WinParams& p, const WinAttnArgs& a) {
  if (a.B <= 0 || a.H <= 0) return LCBI_ERR_BAD_ARG;
  if (a.head_dim != 16 && a.head_dim != 32) return LCBI_ERR_UNSUPPORTED;
  if (fill_win_geom(p.g, a.ndim, a.grid, a.window, a.shift)) return LCBI_ERR_BAD_ARG;
  if (p.g.n > 512) return LCBI_ERR_UNSUPPORTED;
  p.B = a.B; p.H = a.H; p.C = a.H * a.head_dim;
  p.scale_log2 = a.scale * kLog2e;
  p.qkv = static_cast<const __nv_bfloat16*>(a.qkv);
  p.qkv_bias = a.qkv_bias;
  p.table = a.table;
  p.out = static_cast<__nv_bfloat16*>(a.out);
  p.lse2 = a.lse2;
  p.d_out = static_cast<const __nv_bfloat16*>(a.d_out);
  p.dsum = a.dsum;
  p.dqkv = static_cast<__nv_bfloat16*>(a.dqkv);
  p.dbias_pad = a.dbias_pad;
  p.dtable = a.dtable;
  p.win_splits = 1;
  const int all_windows = a.B * p.g.nW;
  p.win_begin = a.win_count < 0 ? 0 : a.win_begin;
  p.win_count = a.win_count < 0 ? all_windows : a.win_count;
  if (p.win_begin < 0 || p.win_count < 0 || p.win_begin + p.win_count > all_windows) return LCBI_ERR_BAD_ARG;
  return LCBI_OK;
}

}  // namespace

int win_attn_fwd_launch(const WinAttnArgs& a, cudaStream_t stream) {
  WinParams p;
  int rc = fill_params(p, a);
  if (rc) return rc;
  // <= 64-token windows with enough of them to stream: persistent kernel with the bias tile in registers and cp.async
  // double buffering (window_attn_small.cu; cfg2 stage 1: 161 -> 119 us). With few windows (late stages) the
  // CTA-per-(window, head) kernel below fills the machine better (measured).
  static const bool legacy_small_fwd = std::getenv("LCBI_WIN_LEGACY_FWD") != nullptr;
  if (p.g.n <= 64 && p.win_count >= 1024 && !legacy_small_fwd) return win_attn_fwd_small_launch(p, a.head_dim, stream);
  // 3-D windows of 128..512 tokens (7^3, 8^3): tcgen05 / TMEM kernel fed by TMA box / gather4 loads (window_attn_tc.cu).
  // Measured on B200 (profiles/r02_swin_tc_vs_generic.log): its persistent CTAs are on par with the generic kernel when
  // there are many (window, head) items per SM (cfg4 stage 1: 256 vs 251 us) and lose when a CTA gets one or two items
  // (stages 2-4) or with head_dim 32, so that is where the line is drawn; lcbi_set_window_kernel_mode overrides it.
  if (win_attn_tc_applicable(p, a.head_dim)) {
    const int mode = window_kernel_mode();
    const bool use_tc = mode == 1 || (mode == 0 && a.head_dim == 16 && p.win_count * p.H >= 1536);
    if (use_tc) {
      rc = win_attn_fwd_tc_launch(p, a.head_dim, stream);
      if (rc != LCBI_ERR_UNSUPPORTED) return rc;
    }
  }
  const size_t smem = fwd_smem_bytes(p.g, a.head_dim);
  dim3 grid(static_cast<unsigned>(p.win_count) * p.H);
  if (a.head_dim == 16) {
    if (p.g.n > 128) {      // eight warps share the window's tiles: twice the warps per SM for the same smem
      if ((rc = set_smem(win_attn_fwd_kernel<16, true, LCBI_WIN_NT>, smem))) return rc;
      win_attn_fwd_kernel<16, true, LCBI_WIN_NT><<<grid, LCBI_WIN_NT, smem, stream>>>(p);
    } else {
      if ((rc = set_smem(win_attn_fwd_kernel<16, true>, smem))) return rc;
      win_attn_fwd_kernel<16, true><<<grid, 128, smem, stream>>>(p);
    }
  } else if (p.g.n <= 64) {
    if ((rc = set_smem(win_attn_fwd_kernel<32, true>, smem))) return rc;
    win_attn_fwd_kernel<32, true><<<grid, 128, smem, stream>>>(p);
  } else {
    if ((rc = set_smem(win_attn_fwd_kernel<32, false, LCBI_WIN_NT>, smem))) return rc;
    win_attn_fwd_kernel<32, false, LCBI_WIN_NT><<<grid, LCBI_WIN_NT, smem, stream>>>(p);
  }
  return set_cuda_error(cudaGetLastError());
}

int win_attn_bwd_launch(const WinAttnArgs& a, const void* o, cudaStream_t stream) {
  WinParams p;
  int rc = fill_params(p, a);
  if (rc) return rc;
  const int D = a.head_dim;
  // 1. dsum = rowsum(dO o O)
  {
    const int64_t total = static_cast<int64_t>(p.B) * p.g.T * p.H;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    if (D == 16)
      win_bwd_prep_kernel<16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(o), p.d_out, a.dsum, total, p.H, p.C);
    else
      win_bwd_prep_kernel<32><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(o), p.d_out, a.dsum, total, p.H, p.C);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  // small windows (<= 64 tokens): one fused kernel produces dq, dk, dv, the pad-token bias gradient and d(table)
  static const bool legacy_small_bwd = std::getenv("LCBI_WIN_LEGACY_BWD") != nullptr;
  if (p.g.n <= 64 && !legacy_small_bwd) return win_attn_bwd_small_launch(p, D, stream);
  // 3-D windows of 128..512 tokens: ONE tcgen05 kernel produces dq, dk, dv, the pad-token bias gradient and d(table) with
  // a single recompute of S (window_attn_tc.cu); shapes whose operands do not fit its shared memory fall through
  // Measured (same log): faster than the two generic kernels for head_dim 16 up to ~1500 items (cfg4 stages 2-4: 385 vs
  // 461, 224 vs 293, 173 vs 195 us), 10 % slower at stage 1 (3000 items) and slower for head_dim 32 (no room to prefetch).
  if (win_attn_tc_applicable(p, D)) {
    const int mode = window_kernel_mode();
    const bool use_tc = mode == 1 || (mode == 0 && D == 16 && p.win_count * p.H < 1536);
    if (use_tc) {
      rc = win_attn_bwd_tc_launch(p, D, stream);
      if (rc != LCBI_ERR_UNSUPPORTED) return rc;
    }
  }
  // 2. dK, dV (+ pad-token bias gradient)
  {
    const size_t smem = dkdv_smem_bytes(p.g, D);
    dim3 grid(static_cast<unsigned>(p.win_count) * p.H);
    if (D == 16) {
      if (p.g.n > 128) {
        if ((rc = set_smem(win_attn_bwd_dkdv_kernel<16, true, LCBI_WIN_NT>, smem))) return rc;
        win_attn_bwd_dkdv_kernel<16, true, LCBI_WIN_NT><<<grid, LCBI_WIN_NT, smem, stream>>>(p);
      } else {
        if ((rc = set_smem(win_attn_bwd_dkdv_kernel<16, true>, smem))) return rc;
        win_attn_bwd_dkdv_kernel<16, true><<<grid, 128, smem, stream>>>(p);
      }
    } else if (p.g.n <= 64) {
      if ((rc = set_smem(win_attn_bwd_dkdv_kernel<32, true>, smem))) return rc;
      win_attn_bwd_dkdv_kernel<32, true><<<grid, 128, smem, stream>>>(p);
    } else {
      if ((rc = set_smem(win_attn_bwd_dkdv_kernel<32, false, LCBI_WIN_NT>, smem))) return rc;
      win_attn_bwd_dkdv_kernel<32, false, LCBI_WIN_NT><<<grid, LCBI_WIN_NT, smem, stream>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  // 3. dQ (+ bias-table gradient): 64-row slabs (4 query tiles) up to 384-token windows, 32-row slabs above
  {
    const int n_ks = round_up(p.g.n, kKeySplit) / kKeySplit;
    // 64-row slabs (one 384-thread CTA per SM) win when there are many windows to stream through a CTA; with few
    // windows (late stages: 27 and 8 per image) twice as many 32-row CTAs fill the machine better (measured)
    const int qt = (n_ks <= 3 && p.win_count > 64) ? 4 : 2;
    const int slab_rows = 16 * qt;
    const size_t smem = dq_smem_bytes(p.g, D, slab_rows);
    const int n_slabs = (p.g.n + slab_rows - 1) / slab_rows;
    const int total_windows = p.win_count;
    int splits = (6 * 148 + n_slabs * p.H - 1) / (n_slabs * p.H);   // aim at ~6 CTAs' worth of work per SM
    if (splits > total_windows) splits = total_windows;
    if (splits < 1) splits = 1;
    p.win_splits = splits;
    dim3 grid(n_slabs, p.H, splits);
    const int threads = 32 * qt * n_ks;
#define LCBI_DQ_LAUNCH(DD, QQ, KK)                                                       \
  do {                                                                                   \
    if ((rc = set_smem(win_attn_bwd_dq_kernel<DD, QQ, KK>, smem))) return rc;            \
    win_attn_bwd_dq_kernel<DD, QQ, KK><<<grid, threads, smem, stream>>>(p);              \
  } while (0)
    if (D == 16) {
      if (qt == 4) LCBI_DQ_LAUNCH(16, 4, 3); else LCBI_DQ_LAUNCH(16, 2, 4);
    } else {
      if (qt == 4) LCBI_DQ_LAUNCH(32, 4, 3); else LCBI_DQ_LAUNCH(32, 2, 4);
    }
#undef LCBI_DQ_LAUNCH
  }
  return set_cuda_error(cudaGetLastError());
}

// -------------------------------------------------------------------------------------------------
// index maps as tensors (used by the bit-exactness tests): gather map, region ids, relative-position index
// -------------------------------------------------------------------------------------------------
namespace {
__global__ void window_maps_kernel(WinGeom g, int* gather, int* region, int* relidx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < g.nW * g.n) {
    int tok, reg;
    slot_lookup(g, idx / g.n, idx % g.n, tok, reg);
    if (gather) gather[idx] = tok;
    if (region) region[idx] = reg;
  }
  if (relidx && idx < g.n * g.n) {
    int rt, ct, rt2, ct2;
    relpos_terms(g, idx / g.n, rt, ct);
    relpos_terms(g, idx % g.n, rt2, ct2);
    relidx[idx] = rt - ct2;
  }
}
}  // namespace

int window_maps_launch(int ndim, const int* grid, const int* window, const int* shift, int* gather, int* region,
                       int* relidx, int* n_out, int* nw_out, cudaStream_t stream) {
  WinGeom g;
  if (fill_win_geom(g, ndim, grid, window, shift)) return LCBI_ERR_BAD_ARG;
  if (n_out) *n_out = g.n;
  if (nw_out) *nw_out = g.nW;
  if (gather || region || relidx) {
    const int total = max(g.nW * g.n, g.n * g.n);
    window_maps_kernel<<<(total + 255) / 256, 256, 0, stream>>>(g, gather, region, relidx);
    return set_cuda_error(cudaGetLastError());
  }
  return LCBI_OK;
}

}  // namespace lcbi
