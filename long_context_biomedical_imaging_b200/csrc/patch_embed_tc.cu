// Patch embedding on tcgen05 / TMEM for reduction lengths K = Cin * prod(patch) that are multiples of 64 (cfg1: 16x16
// patches, K = 256; cfg3: 8x8x8 patches, K = 512), fp32 image and weights, at fp32 accuracy.
//
// Same contract as patch_embed.cu / patch_embed_mma.cu (the reference's MONAI PatchEmbeddingBlock / PatchEmbed call
// sites, /root/reference/model/models/backbone_vit.py:351-361,383 and backbone_swin.py:800-806,885):
//   out[m, n] = sum_k patch(m)[k] * W[n, k] + bias[n] (+ pos[m mod Np, n]),   m = (batch, gz, gy, gx) raster.
//
// fp32 accuracy on bf16 tensor cores: both operands are split x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) and
//   x * w ~= hi_x hi_w + hi_x lo_w + lo_x hi_w        (dropped lo*lo ~ 2^-18; measured max-rel ~ 2e-5 < 1e-4)
// is accumulated in TMEM by three tcgen05.mma per k-step.
//
// Two launches over a caller-owned workspace (lcbi_patch_embed_workspace_bytes):
//   1. pe_split_kernel   HBM-bound pre-pass: image -> patch-major A_hi, A_lo (M, K) bf16 (the im2col and the split in one
//                        coalesced sweep: 4 bytes read + 4 bytes written per pixel), and W -> W_hi, W_lo.
//   2. pe_gemm_kernel    persistent TMA-fed GEMM, one CTA per SM, tile = 128 patches x 128 features, K in blocks of 64
//                        through a 3-stage ring of (A_hi, A_lo, W_hi, W_lo) SWIZZLE_128B tiles:
//        warp 0       TMA producer          warp 1   UMMA issuer (4 k-steps x 3 products per k-block, two TMEM accumulators)
//        warps 4-11   epilogue: TMEM -> 32 x 16 transposes through padded shared memory -> + bias (+ position embedding,
//                     prefetched before the accumulator wait) -> 64-byte row segments of `out`, overlapping the next tile
//   Backward (dW, dbias): pe_split_kernel again, pe_split_dout_kernel (dOut -> G_hi, G_lo and the column sums), and
//   pe_bwd_w_gemm_kernel (both operands MN-major, M split over the CTAs, fp32 reductions into dW) - see below.
// Two fused single-kernel versions were measured first at cfg3 (B = 16): one TMA box per patch and k-block straight from
// the image with converter warps in between took 368 us (the TMA unit paced by 256-byte boxes of 32-byte runs), one box
// per patch ROW 245 us (2-stage TMA -> convert -> MMA chain, latency-bound, and every feature tile re-converts its
// patches: 6x at N = 768). Splitting once costs one extra round trip of the image through HBM (27 us) and wins: 27 + 62 us.
// The GEMM is bound by shared-memory bandwidth, not by the tensor pipe's arithmetic: a 128x128x16 SS MMA reads 8 KB of
// operands in its 64 cycles (the full 128 B/clk) while TMA refills 64 KB per k-block - the hmma sub-pipe counter reads as
// occupied for the whole kernel although the MMAs need half of it (profiles/r02_ncu_pe_gemm_summary.txt). The next step is
// a 128 x 256 or cta_group::2 256 x 256 tile (12 KB of operand reads per 128-cycle MMA: 96 B/clk).
#include "lcbi_kernels.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace lcbi {

namespace {

constexpr int TM = 128, TN = 128, TK = 64;
constexpr int kStages = 3;
constexpr int kPeThreads = 384;
constexpr int kEpiWarps = 8;                  // warps 4-11: TMEM lane quarter = warp % 4, feature half = (warp - 4) / 4
constexpr int kEpiCols = 16, kEpiPitch = 20;  // staging chunk: 32 rows x 16 features, row pitch 20 floats (conflict-free)
constexpr int kOpBytes = TM * TK * 2;         // 16 KB per bf16 operand tile

struct PeTcParams {
  int B, Cin, D, H, W, Pd, Ph, Pw, Gd, Gh, Gw, N, K;
  int64_t M;
  int Np;                 // patches per image
  int n_nt, n_tiles, n_kb;
  const float* bias;
  const float* pos;       // (Np, N) or nullptr
  void* out;
  int out_is_bf16;
};

struct __align__(1024) PeTcSmem {
  uint8_t a_hi[kStages][kOpBytes], a_lo[kStages][kOpBytes];
  uint8_t b_hi[kStages][kOpBytes], b_lo[kStages][kOpBytes];
  float stage[kEpiWarps][32 * kEpiPitch];
  uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    l[i] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Pre-pass. One thread = 16 consecutive k of one patch (two runs of 8 pixels: the same image row when Pw >= 16, two
// rows when Pw = 8); consecutive threads = consecutive patches of a patch row, so a warp reads whole stretches of an
// image row and writes one 32-byte sector per thread and output. The tail of the grid splits the weights.
__global__ void __launch_bounds__(256)
pe_split_kernel(const float* __restrict__ img, const float* __restrict__ w, __nv_bfloat16* __restrict__ ahi,
                __nv_bfloat16* __restrict__ alo, __nv_bfloat16* __restrict__ whi, __nv_bfloat16* __restrict__ wlo,
                const PeTcParams p, int64_t n_img_units, int64_t n_w_units) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_img_units) {
    const int64_t wi = idx - n_img_units;
    if (wi >= n_w_units) return;
    const float4* src = reinterpret_cast<const float4*>(w + wi * 16);
    uint4 h0, l0, h1, l1;
    split8(src[0], src[1], h0, l0);
    split8(src[2], src[3], h1, l1);
    uint4* dh = reinterpret_cast<uint4*>(whi + wi * 16);
    uint4* dl = reinterpret_cast<uint4*>(wlo + wi * 16);
    dh[0] = h0; dh[1] = h1;
    dl[0] = l0; dl[1] = l1;
    return;
  }
  const int n_k16 = p.K / 16;
  const int gx = static_cast<int>(idx % p.Gw);
  int64_t t = idx / p.Gw;
  const int k16 = static_cast<int>(t % n_k16);
  const int64_t prow = t / n_k16;                       // patch row: (batch, gz, gy)
  const int gy = static_cast<int>(prow % p.Gh);
  const int gz = static_cast<int>((prow / p.Gh) % p.Gd);
  const int bb = static_cast<int>(prow / (static_cast<int64_t>(p.Gh) * p.Gd));
  const int64_t m = prow * p.Gw + gx;
  const int slice = p.Pw * p.Ph;
  uint4 hi[2], lo[2];
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    const int k = k16 * 16 + hlf * 8;
    const int cin = k / (slice * p.Pd), rem = k - cin * slice * p.Pd;
    const int kz = rem / slice, rem2 = rem - kz * slice;
    const int ky = rem2 / p.Pw, kx = rem2 - ky * p.Pw;
    const int z = gz * p.Pd + kz, y = gy * p.Ph + ky, x = gx * p.Pw + kx;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (z < p.D && y < p.H) {                            // trailing zero pad of the Swin PatchEmbed
      const float* src = img + (((static_cast<int64_t>(bb) * p.Cin + cin) * p.D + z) * p.H + y) * p.W + x;
      if (x + 4 <= p.W) a = __ldg(reinterpret_cast<const float4*>(src));
      if (x + 8 <= p.W) b = __ldg(reinterpret_cast<const float4*>(src + 4));
    }
    split8(a, b, hi[hlf], lo[hlf]);
  }
  uint4* dh = reinterpret_cast<uint4*>(ahi + m * p.K + k16 * 16);
  uint4* dl = reinterpret_cast<uint4*>(alo + m * p.K + k16 * 16);
  dh[0] = hi[0]; dh[1] = hi[1];
  dl[0] = lo[0]; dl[1] = lo[1];
}

__global__ void __launch_bounds__(kPeThreads, 1)
pe_gemm_kernel(const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
               const __grid_constant__ CUtensorMap tm_whi, const __grid_constant__ CUtensorMap tm_wlo, const PeTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  PeTcSmem& sm = *reinterpret_cast<PeTcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);         // the issuer's commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sm.acc_full[s], 1);
      mbar_init(&sm.acc_empty[s], kEpiWarps);     // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, 256);
    tmem_relinquish();
  }
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_ahi);
    tma_prefetch_desc(&tm_alo);
    tma_prefetch_desc(&tm_whi);
    tma_prefetch_desc(&tm_wlo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int gkb = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_nt, nt = tile - mt * p.n_nt;
        for (int kb = 0; kb < p.n_kb; ++kb, ++gkb) {
          const int s = gkb % kStages;
          const uint32_t ph = (gkb / kStages) & 1;
          mbar_wait(&sm.empty[s], ph ^ 1);
          mbar_expect_tx(&sm.full[s], 4 * kOpBytes);
          tma_load_2d(sm.a_hi[s], &tm_ahi, &sm.full[s], kb * TK, mt * TM);
          tma_load_2d(sm.a_lo[s], &tm_alo, &sm.full[s], kb * TK, mt * TM);
          tma_load_2d(sm.b_hi[s], &tm_whi, &sm.full[s], kb * TK, nt * TN);
          tma_load_2d(sm.b_lo[s], &tm_wlo, &sm.full[s], kb * TK, nt * TN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(TM, TN, 0, 0);
      int gkb = 0, gt = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++gt) {
        const int buf = gt & 1;
        mbar_wait(&sm.acc_empty[buf], ((gt >> 1) & 1) ^ 1);     // the epilogue drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem + buf * TN;
        for (int kb = 0; kb < p.n_kb; ++kb, ++gkb) {
          const int s = gkb % kStages;
          const uint32_t ph = (gkb / kStages) & 1;
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          const uint32_t ah = smem_u32(sm.a_hi[s]), al = smem_u32(sm.a_lo[s]), bh = smem_u32(sm.b_hi[s]), bl = smem_u32(sm.b_lo[s]);
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            const uint64_t d_ah = make_smem_desc(ah + kk * 32, 16, 1024, kLayoutSW128);
            const uint64_t d_al = make_smem_desc(al + kk * 32, 16, 1024, kLayoutSW128);
            const uint64_t d_bh = make_smem_desc(bh + kk * 32, 16, 1024, kLayoutSW128);
            const uint64_t d_bl = make_smem_desc(bl + kk * 32, 16, 1024, kLayoutSW128);
            umma_ss(acc, d_ah, d_bh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_ss(acc, d_ah, d_bl, idesc, 1u);
            umma_ss(acc, d_al, d_bh, idesc, 1u);
          }
          umma_commit(&sm.empty[s]);
        }
        umma_commit(&sm.acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: + bias (+ pos) -> out
    // TMEM hands every lane one ROW of the tile; writing rows from lanes would scatter 16-byte pieces over 32 rows per
    // instruction (measured: the epilogue, not the MMAs, set the pace). Each warp therefore transposes 32 x 16 chunks
    // through shared memory: 4 lanes cover 64 contiguous bytes of a row, 8 rows per instruction, for the position
    // embedding read and the output write alike.
    const int ew = warp - 4;
    const int quarter = warp & 3, half = ew >> 2;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    float* stg = sm.stage[ew];
    const int rr = lane >> 2, cc = (lane & 3) * 4;
    int gt = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++gt) {
      const int mt = tile / p.n_nt, nt = tile - mt * p.n_nt;
      const int buf = gt & 1;
      const int64_t m_base = static_cast<int64_t>(mt) * TM + quarter * 32;
      const int n_base = nt * TN + half * (TN / 2);
      // bias and position-embedding pieces of the whole tile are fetched BEFORE waiting for the accumulator, so that their
      // latency hides behind the MMAs (fetching them per chunk serialised ~4 L2 round trips per tile)
      constexpr int kChunks = (TN / 2) / kEpiCols;
      float4 b4[kChunks], q4[kChunks][4];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int n = n_base + c * kEpiCols + cc;
        b4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.N) b4[c] = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t m = m_base + i * 8 + rr;
          q4[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.pos != nullptr && m < p.M && n < p.N)
            q4[c][i] = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(m % p.Np) * p.N + n));
        }
      }
      mbar_wait(&sm.acc_full[buf], (gt >> 1) & 1);
      tc_fence_after();
      uint32_t r[2][16];
      tmem_ld_x16(tmem + lane_sel + buf * TN + half * (TN / 2), r[0]);
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        tmem_ld_wait();
        if (c + 1 < kChunks) tmem_ld_x16(tmem + lane_sel + buf * TN + half * (TN / 2) + (c + 1) * kEpiCols, r[(c + 1) & 1]);
        __syncwarp();                                   // the previous chunk's reads of the staging buffer are done
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<uint4*>(stg + lane * kEpiPitch + j) =
              make_uint4(r[c & 1][j], r[c & 1][j + 1], r[c & 1][j + 2], r[c & 1][j + 3]);
        __syncwarp();
        const int n = n_base + c * kEpiCols + cc;
        if (n < p.N) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int64_t m = m_base + i * 8 + rr;
            const float4 a4 = *reinterpret_cast<const float4*>(stg + (i * 8 + rr) * kEpiPitch + cc);
            const float4 v = make_float4(a4.x + b4[c].x + q4[c][i].x, a4.y + b4[c].y + q4[c][i].y, a4.z + b4[c].z + q4[c][i].z,
                                         a4.w + b4[c].w + q4[c][i].w);
            if (m >= p.M) continue;
            if (p.out_is_bf16) {
              __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(m) * p.N + n;
              *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
            } else {
              *reinterpret_cast<float4*>(static_cast<float*>(p.out) + static_cast<size_t>(m) * p.N + n) = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.acc_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}


// ------------------------------------------------------------------------------------------------------------------
// Weight gradient: dW[n, k] = sum_m dOut[m, n] * patch(m)[k], and dbias[n] = sum_m dOut[m, n].
// The contraction runs over the M patches, the slow dimension of both (M, N) and (M, K) arrays, so both UMMA operands are
// MN-major: a stage holds [64 patches x 128] tiles of G_hi, G_lo (dOut split) and A_hi, A_lo (patches split), each as two
// [64 x 64] SWIZZLE_128B atoms 8 KB apart (the descriptor's leading-dimension offset). There are only
// ceil(N/128) * ceil(K/128) output tiles (24 at cfg3), so the M range is split over as many CTAs as fill the GPU and the
// partial tiles are added into the zeroed dW with fp32 reductions.
// ------------------------------------------------------------------------------------------------------------------
struct __align__(1024) PeBwSmem {
  uint8_t g_hi[kStages][kOpBytes], g_lo[kStages][kOpBytes];
  uint8_t a_hi[kStages][kOpBytes], a_lo[kStages][kOpBytes];
  uint64_t full[kStages], empty[kStages], acc_full;
  uint32_t tmem_base;
};

// dOut (M, N) fp32 -> G_hi, G_lo bf16, and the column sums (dbias) on the way: 8 features per thread.
__global__ void __launch_bounds__(256)
pe_split_dout_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ ghi, __nv_bfloat16* __restrict__ glo,
                     float* __restrict__ dbias, int64_t M, int N, int rows_per_block) {
  extern __shared__ float s_sum[];                   // N floats
  const int ncg = N / 8, rpp = 256 / ncg;
  const int cg = threadIdx.x % ncg, ro = threadIdx.x / ncg;
  for (int i = threadIdx.x; i < N; i += 256) s_sum[i] = 0.f;
  __syncthreads();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  if (ro < rpp) {
#pragma unroll 4
    for (int64_t r = r0 + ro; r < r1; r += rpp) {
      const float4* src = reinterpret_cast<const float4*>(dout + r * N + cg * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      uint4 hi, lo;
      split8(a, b, hi, lo);
      *reinterpret_cast<uint4*>(ghi + r * N + cg * 8) = hi;
      *reinterpret_cast<uint4*>(glo + r * N + cg * 8) = lo;
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    if (dbias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_sum[cg * 8 + j], acc[j]);
    }
  }
  if (dbias == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += 256) atomicAdd(dbias + i, s_sum[i]);
}

__global__ void __launch_bounds__(256, 1)
pe_bwd_w_gemm_kernel(const __grid_constant__ CUtensorMap tm_ghi, const __grid_constant__ CUtensorMap tm_glo,
                     const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                     float* __restrict__ dw, int N, int K, int n_kt, int n_mb, int splits, int mb_per_split) {
  extern __shared__ uint8_t smem_raw[];
  PeBwSmem& sm = *reinterpret_cast<PeBwSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / splits, split = blockIdx.x - tile * splits;
  const int nt = tile / n_kt, kt = tile - nt * n_kt;
  const int mb0 = split * mb_per_split;
  const int mb1 = mb0 + mb_per_split < n_mb ? mb0 + mb_per_split : n_mb;      // host guarantees mb0 < n_mb

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(&sm.acc_full, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) {
    tmem_alloc(&sm.tmem_base, 128);
    tmem_relinquish();
  }
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_ghi);
    tma_prefetch_desc(&tm_glo);
    tma_prefetch_desc(&tm_ahi);
    tma_prefetch_desc(&tm_alo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  constexpr int kAtom = kOpBytes / 2;      // [64 patches x 64] bf16, 8 KB

  if (warp == 0) {
    if (elect_one()) {
      for (int mb = mb0, g = 0; mb < mb1; ++mb, ++g) {
        const int s = g % kStages;
        const uint32_t ph = (g / kStages) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        mbar_expect_tx(&sm.full[s], 4 * kOpBytes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tma_load_2d(sm.g_hi[s] + h * kAtom, &tm_ghi, &sm.full[s], nt * TM + h * 64, mb * 64);
          tma_load_2d(sm.g_lo[s] + h * kAtom, &tm_glo, &sm.full[s], nt * TM + h * 64, mb * 64);
          tma_load_2d(sm.a_hi[s] + h * kAtom, &tm_ahi, &sm.full[s], kt * TN + h * 64, mb * 64);
          tma_load_2d(sm.a_lo[s] + h * kAtom, &tm_alo, &sm.full[s], kt * TN + h * 64, mb * 64);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(TM, TN, 1, 1);       // both operands MN-major
      for (int mb = mb0, g = 0; mb < mb1; ++mb, ++g) {
        const int s = g % kStages;
        const uint32_t ph = (g / kStages) & 1;
        mbar_wait(&sm.full[s], ph);
        tc_fence_after();
        const uint32_t gh = smem_u32(sm.g_hi[s]), gl = smem_u32(sm.g_lo[s]), ah = smem_u32(sm.a_hi[s]), al = smem_u32(sm.a_lo[s]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {                                // 16 patches per k-step: 16 rows of 128 bytes
          const uint64_t d_gh = make_smem_desc(gh + kk * 2048, kAtom, 1024, kLayoutSW128);
          const uint64_t d_gl = make_smem_desc(gl + kk * 2048, kAtom, 1024, kLayoutSW128);
          const uint64_t d_ah = make_smem_desc(ah + kk * 2048, kAtom, 1024, kLayoutSW128);
          const uint64_t d_al = make_smem_desc(al + kk * 2048, kAtom, 1024, kLayoutSW128);
          umma_ss(tmem, d_gh, d_ah, idesc, (g > 0 || kk > 0) ? 1u : 0u);
          umma_ss(tmem, d_gh, d_al, idesc, 1u);
          umma_ss(tmem, d_gl, d_ah, idesc, 1u);
        }
        umma_commit(&sm.empty[s]);
      }
      umma_commit(&sm.acc_full);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int n = nt * TM + ew * 32 + lane;
    mbar_wait(&sm.acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < TN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_x32(tmem + (static_cast<uint32_t>(ew * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int k0 = kt * TN + c * 32;
      if (n < N && k0 < K) {
        float* dst = dw + static_cast<size_t>(n) * K + k0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + j < K) atomicAdd(dst + j, __uint_as_float(r[j]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

}  // namespace

// fp32 image, K a multiple of 64, patch rows that are multiples of 8 pixels, 16-byte aligned image rows, N % 8 == 0.
// Trailing zero padding (grid * patch > image) is handled by the pre-pass.
bool patch_embed_tc_applicable(int img_is_bf16, int Cin, const int* img_dims, const int* patch, const int* grid, int N) {
  const int K = Cin * patch[0] * patch[1] * patch[2];
  (void)grid;
  if (img_is_bf16 || K < 64 || K % TK != 0 || N % 8 != 0) return false;
  if (patch[2] % 8 != 0 || img_dims[2] % 4 != 0 || N > 2048) return false;
  return true;
}

// forward: A_hi, A_lo (M, K) + W_hi, W_lo (N padded to 128, K); backward: A_hi, A_lo (M, K) + G_hi, G_lo (M, N); all bf16
size_t patch_embed_tc_workspace_bytes(int64_t M, int N, int K) {
  const size_t n_pad = (static_cast<size_t>(N) + TN - 1) / TN * TN;
  const size_t m_pad = (static_cast<size_t>(M) + TM - 1) / TM * TM;
  const size_t fwd = 2 * (n_pad + m_pad) * K * 2, bwd = 2 * m_pad * (static_cast<size_t>(K) + N) * 2;
  return (fwd > bwd ? fwd : bwd) + 256;
}

static PeTcParams pe_tc_geometry(int B, int Cin, const int* img_dims, const int* patch, const int* grid, int N) {
  PeTcParams p{};
  p.B = B; p.Cin = Cin; p.D = img_dims[0]; p.H = img_dims[1]; p.W = img_dims[2];
  p.Pd = patch[0]; p.Ph = patch[1]; p.Pw = patch[2];
  p.Gd = grid[0]; p.Gh = grid[1]; p.Gw = grid[2];
  p.N = N; p.K = Cin * patch[0] * patch[1] * patch[2];
  p.Np = grid[0] * grid[1] * grid[2];
  p.M = static_cast<int64_t>(B) * p.Np;
  p.n_nt = (N + TN - 1) / TN;
  p.n_tiles = static_cast<int>((p.M + TM - 1) / TM) * p.n_nt;
  p.n_kb = p.K / TK;
  return p;
}

// dW and dbias (both zeroed by the caller) of the tcgen05 path; fp32 dOut, N <= 2048.
int patch_embed_tc_bwd_w_launch(const void* img, const float* dout, float* dw, float* dbias, int B, int Cin, const int* img_dims,
                                const int* patch, const int* grid, int N, void* workspace, size_t workspace_bytes,
                                cudaStream_t stream) {
  PeTcParams p = pe_tc_geometry(B, Cin, img_dims, patch, grid, N);
  if (workspace == nullptr || workspace_bytes < patch_embed_tc_workspace_bytes(p.M, N, p.K) ||
      (reinterpret_cast<uintptr_t>(workspace) & 127) != 0)
    return LCBI_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(img) & 15) != 0 || (reinterpret_cast<uintptr_t>(dout) & 15) != 0) return LCBI_ERR_BAD_ARG;
  const size_t m_pad = static_cast<size_t>((p.M + TM - 1) / TM) * TM;
  __nv_bfloat16* ahi = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* alo = ahi + m_pad * p.K;
  __nv_bfloat16* ghi = alo + m_pad * p.K;
  __nv_bfloat16* glo = ghi + m_pad * N;
  {
    const int64_t n_img_units = p.M * (p.K / 16);
    pe_split_kernel<<<static_cast<unsigned>((n_img_units + 255) / 256), 256, 0, stream>>>(
        static_cast<const float*>(img), nullptr, ahi, alo, nullptr, nullptr, p, n_img_units, 0);
    const int rpp = 256 / (N / 8);
    int64_t rows = (p.M + 4 * 148 - 1) / (4 * 148);
    rows = (rows + rpp - 1) / rpp * rpp;
    pe_split_dout_kernel<<<static_cast<unsigned>((p.M + rows - 1) / rows), 256, N * sizeof(float), stream>>>(
        dout, ghi, glo, dbias, p.M, N, static_cast<int>(rows));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  CUtensorMap tm_ghi, tm_glo, tm_ahi, tm_alo;
  {
    const uint64_t gd[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(p.M)};
    const uint64_t ad[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
    const uint64_t gs[1] = {static_cast<uint64_t>(N) * 2}, as[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t box[2] = {64, 64};
    if (make_tmap(&tm_ghi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ghi, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap(&tm_glo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, glo, gd, gs, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap(&tm_ahi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ahi, ad, as, box, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap(&tm_alo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, alo, ad, as, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return LCBI_ERR_TENSOR_MAP;
  }
  static unsigned long long configured = 0;
  const int smem = static_cast<int>(sizeof(PeBwSmem)) + 1024;
  if (first_launch_on_current_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(pe_bwd_w_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      configured = 0;
      return set_cuda_error(e);
    }
  }
  const int num_sms = current_device_sm_count();
  if (num_sms <= 0) return LCBI_ERR_CUDA;
  const int n_kt = (p.K + TN - 1) / TN, n_out = p.n_nt * n_kt;
  const int n_mb = static_cast<int>((p.M + 63) / 64);
  int splits = num_sms / n_out;
  if (splits < 1) splits = 1;
  if (splits > n_mb) splits = n_mb;
  const int mb_per_split = (n_mb + splits - 1) / splits;
  splits = (n_mb + mb_per_split - 1) / mb_per_split;              // no empty split
  pe_bwd_w_gemm_kernel<<<n_out * splits, 256, smem, stream>>>(tm_ghi, tm_glo, tm_ahi, tm_alo, dw, N, p.K, n_kt, n_mb, splits,
                                                             mb_per_split);
  return set_cuda_error(cudaGetLastError());
}

int patch_embed_tc_fwd_launch(const void* img, const float* w, const float* bias, const float* pos, void* out, int out_is_bf16,
                              int B, int Cin, const int* img_dims, const int* patch, const int* grid, int N, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream) {
  PeTcParams p = pe_tc_geometry(B, Cin, img_dims, patch, grid, N);
  p.bias = bias; p.pos = pos; p.out = out; p.out_is_bf16 = out_is_bf16;
  if (workspace == nullptr || workspace_bytes < patch_embed_tc_workspace_bytes(p.M, N, p.K) ||
      (reinterpret_cast<uintptr_t>(workspace) & 127) != 0)
    return LCBI_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(img) & 15) != 0 || (reinterpret_cast<uintptr_t>(w) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(bias) & 15) != 0 || (pos != nullptr && (reinterpret_cast<uintptr_t>(pos) & 15) != 0) ||
      (reinterpret_cast<uintptr_t>(out) & 15) != 0)
    return LCBI_ERR_BAD_ARG;

  const size_t n_pad = static_cast<size_t>(p.n_nt) * TN;
  const size_t m_pad = static_cast<size_t>((p.M + TM - 1) / TM) * TM;
  __nv_bfloat16* ahi = static_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* alo = ahi + m_pad * p.K;
  __nv_bfloat16* whi = alo + m_pad * p.K;
  __nv_bfloat16* wlo = whi + n_pad * p.K;
  {
    const int64_t n_img_units = p.M * (p.K / 16), n_w_units = static_cast<int64_t>(N) * p.K / 16;
    const int64_t total = n_img_units + n_w_units;
    pe_split_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(static_cast<const float*>(img), w, ahi, alo,
                                                                                    whi, wlo, p, n_img_units, n_w_units);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }
  CUtensorMap tm_ahi, tm_alo, tm_whi, tm_wlo;
  {
    const uint64_t ad[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
    const uint64_t wd[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(N)};
    const uint64_t st[1] = {static_cast<uint64_t>(p.K) * 2};
    const uint32_t ab[2] = {TK, TM}, wb[2] = {TK, TN};
    if (make_tmap(&tm_ahi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ahi, ad, st, ab, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap(&tm_alo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, alo, ad, st, ab, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap(&tm_whi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, whi, wd, st, wb, CU_TENSOR_MAP_SWIZZLE_128B) ||
        make_tmap(&tm_wlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wlo, wd, st, wb, CU_TENSOR_MAP_SWIZZLE_128B))
      return LCBI_ERR_TENSOR_MAP;
  }
  static unsigned long long configured = 0;
  const int smem = static_cast<int>(sizeof(PeTcSmem)) + 1024;
  if (first_launch_on_current_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(pe_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      configured = 0;
      return set_cuda_error(e);
    }
  }
  const int num_sms = current_device_sm_count();
  if (num_sms <= 0) return LCBI_ERR_CUDA;
  const int grid_x = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  pe_gemm_kernel<<<grid_x, kPeThreads, smem, stream>>>(tm_ahi, tm_alo, tm_whi, tm_wlo, p);
  return set_cuda_error(cudaGetLastError());
}

}  // namespace lcbi
