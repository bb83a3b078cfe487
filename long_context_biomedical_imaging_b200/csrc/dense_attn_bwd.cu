// Dense (ViT global) attention backward for sm_100a, head_dim 64, bf16 in / fp32 accumulate.
//
// Gradient of the reference's SABlock attention core (/root/reference/model/models/backbone_vit.py:191-201,
// what autograd derives from einsum -> *scale -> softmax -> einsum) computed flash-style: S and P are
// recomputed per tile from q, k and the forward's log-sum-exp, nothing N x N is stored.
//
// Three launches:
//   bwd_prep      per query row: a = -lse/scale and b = -D, D = sum_d dO*O, each split into three bf16 parts (hi, mid,
//                 lo) and laid out as [64 x 16] UMMA operand tiles; zeroes the fp32 dQ accumulator
//   bwd_main      persistent; work item = one 128-key tile of one (batch, head), looping over 64-query steps.
//                 dK, dV accumulate in TMEM over an item; each 128-query tile's dQ partial is reduced into an fp32
//                 accumulator with TMA reduce-adds (cp.reduce.async.bulk.tensor .add.f32).
//   bwd_finish    dq = bf16(dq_accum)
//
// bwd_main: all five GEMMs run on tcgen05 with accumulators in TMEM, in a transposed formulation so that the
// exponentiating threads own KEY rows and P^T / dS^T come out in the layout the next GEMM wants. K and V sit in TMEM
// as A operands; S^T and dP^T are double-buffered in TMEM, so the tensor core computes step s+2's S^T / dP^T while
// the compute warps work on step s+1:
//   S^T(s)  = K Q_s^T  - lse/scale    A = K  (TMEM)   B = Q_s  (smem, K-major)  + one K=16 step: A = (1,1,1,0..) tile,
//   dP^T(s) = V dO_s^T - D            A = V  (TMEM)   B = dO_s (smem, K-major)    B = the step's (hi, mid, lo) tile
//   one fused pass per step (warps 0-7, thread = key row x 32 query columns), no per-query loads needed:
//       P^T  = exp2(S^T * scale*log2e)   -> bf16 -> in place over the first half of the S^T columns it came from
//       dS^T = P^T o dP^T                -> bf16 -> in place over dP^T, and into the swizzled dS^T smem tile
//   dV  += P^T(s)  dO_s    A = P^T  (TMEM)   B = dO_s (smem, MN-major)   -> TMEM [256,320)
//   dK  += dS^T(s) Q_s     A = dS^T (TMEM)   B = Q_s  (smem, MN-major)   -> TMEM [320,384)
//   every second step:  dQ_i = dS(i) K   A = both dS^T smem atoms read MN-major, B = K (smem, MN-major) -> [384,448)
// Warp roles (512 threads): warps 0-7 compute, warps 8-11 drain dQ_i (TMEM -> x scale -> swizzled fp32 smem -> TMA
// reduce-add), warp 12 TMA producer (Q / dO / row-term tiles through 4-stage rings that run
// across items, K double-buffered across items, V), warp 13 UMMA issuer (operand descriptors are built once; only
// the start-address field advances), warps 14-15 idle.
// Item transitions overlap: K(next) is prefetched during the item, V(next) follows once V has been copied to TMEM,
// the compute warps move K/V(next) into TMEM right after their last step, so S^T / dP^T of the next item run while
// dV / dK of this one are drained (TMEM -> bf16 -> the idle dS^T smem buffer -> TMA store).
#include "lcbi_kernels.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace lcbi {

namespace {

constexpr int kTile = 128;
constexpr int kHeadDim = 64;
constexpr int kTileBytes = kTile * kHeadDim * 2;  // 16 KB
constexpr int kNumThreads = 512;
constexpr float kLog2e = 1.4426950408889634f;

constexpr int kStep = 64;                          // queries per pipeline step
constexpr int kStepBytes = kStep * kHeadDim * 2;   // 8 KB
#ifndef LCBI_BWD_QSTAGES
#define LCBI_BWD_QSTAGES 4
#endif
constexpr int kQStages = LCBI_BWD_QSTAGES;   // ring depth of the Q / dO / row-term step tiles
// Per-query row terms ride along as one extra K=16 step of the S^T / dP^T GEMMs: the fp32 value is split into three
// bf16 parts (hi, mid, lo) stored in columns 0-2 of a [64 queries x 16] K-major tile, multiplied by a constant
// [128 keys x 16] tile holding (1, 1, 1, 0, ...). Tiles use the un-swizzled canonical layout: 8-row x 16-byte core
// matrices 128 bytes apart (SBO). Only the first 16-byte k-half of a row carries data; the second k-half of EVERY
// tile is one shared block of zeros, reached through the descriptor's leading byte offset (LBO = zeros - tile).
constexpr int kAugBytes = kStep * 16;              // 1 KB: [64 rows x 8 bf16]
constexpr uint32_t kAugSbo = 128;
__host__ __device__ constexpr int aug_chunk_offset(int row) {   // byte offset of a row's 16-byte chunk
  return (row >> 3) * static_cast<int>(kAugSbo) + (row & 7) * 16;
}

// TMEM columns: S^T and dP^T are double-buffered per 64-query step; the bf16 P^T / dS^T of a step overwrite, in
// place, the first half of the columns their owner thread read (thread (row, hh) owns columns [32hh, 32hh+32) of a
// buffer and writes 16 packed columns at [32hh, 32hh+16)); K and V sit in TMEM as the A operands of S^T / dP^T.
constexpr uint32_t kTmemS = 0, kTmemDP = 128, kTmemDV = 256, kTmemDK = 320, kTmemDQ = 384, kTmemK = 448, kTmemV = 480;

struct __align__(1024) BwdSmem {
  uint8_t k[2][kTileBytes];             // K of item it in k[it & 1]: the next item's K is prefetched during this one
  uint8_t v[kTileBytes];                // only read by the copy into TMEM, so the next item's V can follow early too
  uint8_t q[kQStages][kStepBytes];      // 64-query tiles; together also the dK staging area of the epilogue
  uint8_t dout[kQStages][kStepBytes];   // likewise dV staging
  uint8_t ds[2][2 * kTileBytes];        // dS^T per 128-query tile (double-buffered): two [128 keys x 64 queries] atoms
  uint8_t dq_stage[2 * kTileBytes];               // two [128 queries x 32 fp32] SW128 tiles (TMA-reduce drain)
  uint8_t lse_aug[kQStages][kAugBytes];   // per-query -lse/scale as the B operand of one extra k-step of S^T
  uint8_t d_aug[kQStages][kAugBytes];     // per-query -D likewise for dP^T
  uint8_t ones[2 * kAugBytes];            // [128 keys x 8] constant A operand of those k-steps: (1, 1, 1, 0, ...)
  uint8_t aug_zeros[2 * kAugBytes];       // second k-half of every row-term tile (placed after them: LBO > 0)
  uint64_t k_full[2], v_full;
  uint64_t q_full[kQStages], q_empty[kQStages], do_full[kQStages], do_empty[kQStages];
  uint64_t sdp_full[2], pds_full[2], kvt_full, dq_full, dq_empty, dkv_full, dkv_drained;
  uint32_t tmem_base;
};

#ifdef LCBI_TRACE
__device__ long long* g_bwd_trace = nullptr;
__device__ long long* g_bwd_item_times = nullptr;   // [cta][item (<= 32)][8] globaltimer stamps of compute thread 0
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define LCBI_ITEM_T(slot)                                                                              \
  do {                                                                                                 \
    if (g_bwd_item_times != nullptr && tid == 0 && it < 32)                                            \
      g_bwd_item_times[(static_cast<size_t>(blockIdx.x) * 32 + it) * 8 + (slot)] = global_ns();       \
  } while (0)
#define LCBI_TR_INIT() \
  long long* const lcbi_tr = (blockIdx.x == 0) ? g_bwd_trace : nullptr
#define LCBI_TR(role, step, ev)                                                            \
  do {                                                                                     \
    if (lcbi_tr != nullptr && (step) < 16) lcbi_tr[((role) * 16 + (step)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define LCBI_TR_INIT() do { } while (0)
#define LCBI_TR(role, step, ev) do { } while (0)
#define LCBI_ITEM_T(slot) do { } while (0)
#endif

static_assert(sizeof(BwdSmem) + 1024 <= 232448, "BwdSmem exceeds the 227 KB dynamic shared memory limit");

struct BwdParams {
  int B, H, Nq, Nk, Nq_pad;
  int n_kv_tiles, n_items;
  float scale, scale_log2;
  const __nv_bfloat16* lse_aug;   // (B,H,Nq_pad/64) tiles of kAugBytes: -lse/scale split into bf16 (hi, mid, lo)
  const __nv_bfloat16* d_aug;     // likewise -D, D = rowsum(dO o O)
  float* dq_acc;       // fp32 (B,Nq,H,64) accumulator
  int accumulate_dkv;
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// prep: per query row the two terms the main kernel adds through its extra k-step, a = -lse/scale and b = -D with
// D = rowsum(dO o O), each split into three bf16 parts and written in the operand tile layout; 8 threads per row.
// Also zeroes the fp32 dQ accumulator (unless the caller accumulates into its own), saving a memset launch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 split3_bf16(float x) {
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

__global__ void bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                const float* __restrict__ lse, uint8_t* __restrict__ lse_aug,
                                uint8_t* __restrict__ d_aug, float* __restrict__ dq_acc_to_zero, float inv_scale, int B,
                                int H, int Nq, int Nq_pad,
                                int64_t o_sb, int64_t o_sr, int64_t o_sh, int64_t do_sb, int64_t do_sr, int64_t do_sh) {
  // 8 threads per (b, h, row): each loads 16 bytes of O and dO; rows ordered (b, row, h) so that a warp's four
  // rows are adjacent heads of one token (contiguous 512 bytes in the usual (B,N,H,d) layout)
  const int sub = threadIdx.x & 7;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int64_t total = static_cast<int64_t>(B) * H * Nq_pad;
  if (r >= total) return;
  const int h = static_cast<int>(r % H);
  const int row = static_cast<int>((r / H) % Nq_pad);
  const int b = static_cast<int>(r / (static_cast<int64_t>(H) * Nq_pad));
  const int64_t tile = (static_cast<int64_t>(b) * H + h) * (Nq_pad / kStep) + row / kStep;
  const int64_t chunk0 = tile * kAugBytes + aug_chunk_offset(row % kStep);
  float acc = 0.f;
  if (row < Nq) {
    if (dq_acc_to_zero != nullptr) {   // the fp32 dQ accumulator row of (b, row, h): 256 bytes, 32 per thread
      float4* z = reinterpret_cast<float4*>(dq_acc_to_zero + ((static_cast<int64_t>(b) * Nq + row) * H + h) * kHeadDim) + sub * 2;
      z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint4 ov = *reinterpret_cast<const uint4*>(o + b * o_sb + row * o_sr + h * o_sh + sub * 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(d_o + b * do_sb + row * do_sr + h * do_sh + sub * 8);
    const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
    const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = __bfloat1622float2(o2[i]), y = __bfloat1622float2(d2[i]);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0) {
    // padded query rows get a huge negative score offset: exp2 of it is exactly 0, so they contribute nothing
    const float a = row < Nq ? -lse[(static_cast<int64_t>(b) * H + h) * Nq + row] * inv_scale : -1e30f;
    *reinterpret_cast<uint4*>(lse_aug + chunk0) = split3_bf16(a);
    *reinterpret_cast<uint4*>(d_aug + chunk0) = split3_bf16(-acc);
  }
}

// dq = bf16(dq_accum); one thread = 8 consecutive elements of one (b, row, h)
__global__ void bwd_finish_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int B, int H, int Nq,
                                  int64_t sb, int64_t sr, int64_t sh) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * Nq * H * (kHeadDim / 8);
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % (kHeadDim / 8));
  const int h = static_cast<int>((idx / (kHeadDim / 8)) % H);
  const int row = static_cast<int>((idx / (kHeadDim / 8) / H) % Nq);
  const int b = static_cast<int>(idx / (kHeadDim / 8) / H / Nq);
  const float4 a0 = *reinterpret_cast<const float4*>(acc + idx * 8);
  const float4 a1 = *reinterpret_cast<const float4*>(acc + idx * 8 + 4);
  uint4 out;
  out.x = pack_bf16x2(a0.x, a0.y);
  out.y = pack_bf16x2(a0.z, a0.w);
  out.z = pack_bf16x2(a1.x, a1.y);
  out.w = pack_bf16x2(a1.z, a1.w);
  *reinterpret_cast<uint4*>(dq + b * sb + row * sr + h * sh + c8 * 8) = out;
}

// ------------------------------------------------------------------------------------------------
// main
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

__global__ void __launch_bounds__(kNumThreads, 1)
dense_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_dqacc, const __grid_constant__ CUtensorMap tm_dk,
                      const __grid_constant__ CUtensorMap tm_dv, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_steps = p.Nq_pad / kStep;      // 64-query steps per item; even because Nq_pad is a multiple of 128
  const int n_tiles = n_steps >> 1;          // 128-query tiles per item
  const int first_item = blockIdx.x, item_stride = gridDim.x;
  LCBI_TR_INIT();

  // Work item = one 128-key tile of one (batch, head); the kernel is persistent (CTA c takes items c, c + grid, ...).
  // Adjacent items are adjacent key tiles of one (batch, head), so the CTAs running at any moment stream the same
  // Q / dO rows (L2 hits). Every ring and barrier phase below derives from counters that run across items:
  //   gs = it * n_steps + s   (64-query step)       gi = it * n_tiles + i   (128-query tile)
  auto decode = [&](int item, int& kv_base, int& head, int& batch) {
    const int kv_tile = item % p.n_kv_tiles, bh = item / p.n_kv_tiles;
    head = bh % p.H;
    batch = bh / p.H;
    kv_base = kv_tile * kTile;
  };
  auto load_step = [&](int gs, int s, int head, int batch) {   // Q_s with its row terms, and dO_s, into ring stage gs % 4
    const int st = gs % kQStages;
    const size_t aug = ((static_cast<size_t>(batch) * p.H + head) * n_steps + s) * (kAugBytes / 2);
    mbar_expect_tx(&sm.q_full[st], kStepBytes + 2 * kAugBytes);
    tma_load_4d(sm.q[st], &tm_q, &sm.q_full[st], 0, head, s * kStep, batch);
    bulk_load_1d(sm.lse_aug[st], p.lse_aug + aug, kAugBytes, &sm.q_full[st]);
    bulk_load_1d(sm.d_aug[st], p.d_aug + aug, kAugBytes, &sm.q_full[st]);
    mbar_expect_tx(&sm.do_full[st], kStepBytes);
    tma_load_4d(sm.dout[st], &tm_do, &sm.do_full[st], 0, head, s * kStep, batch);
  };
  auto load_k = [&](int buf, int kv_base, int head, int batch) {
    mbar_expect_tx(&sm.k_full[buf], kTileBytes);
    tma_load_4d(sm.k[buf], &tm_k, &sm.k_full[buf], 0, head, kv_base, batch);
  };
  auto load_v = [&](int kv_base, int head, int batch) {
    mbar_expect_tx(&sm.v_full, kTileBytes);
    tma_load_4d(sm.v, &tm_v, &sm.v_full, 0, head, kv_base, batch);
  };
  const int n_prefill = n_steps < kQStages ? n_steps : kQStages;

  // The TMA warp initialises the barriers and starts the first item's loads right away, before the CTA-wide sync
  // below (TMEM allocation, constant tile).
  if (warp == 12 && elect_one()) {
    mbar_init(&sm.k_full[0], 1);
    mbar_init(&sm.k_full[1], 1);
    mbar_init(&sm.v_full, 1);
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(&sm.q_full[s], 1);
      mbar_init(&sm.q_empty[s], 1);
      mbar_init(&sm.do_full[s], 1);
      mbar_init(&sm.do_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.sdp_full[b], 1);
      mbar_init(&sm.pds_full[b], 8);       // one arrive per compute warp
    }
    mbar_init(&sm.kvt_full, 8);
    mbar_init(&sm.dq_full, 1);
    mbar_init(&sm.dq_empty, 4);        // one arrive per drain warp
    mbar_init(&sm.dkv_full, 1);
    mbar_init(&sm.dkv_drained, 8);     // one arrive per compute warp
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    if (first_item < p.n_items) {
      int kv_base, head, batch;
      decode(first_item, kv_base, head, batch);
      load_k(0, kv_base, head, batch);
      load_v(kv_base, head, batch);
      for (int s = 0; s < n_prefill; ++s) load_step(s, s, head, batch);
    }
  }
  if (tid < 2 * kAugBytes / 16) {     // constant A tile of the extra k-step (columns 0-2 = 1.0) and the shared zeros
    *reinterpret_cast<uint4*>(sm.ones + tid * 16) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
    *reinterpret_cast<uint4*>(sm.aug_zeros + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 13) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int it = 0;
      for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
        int kv_base, head, batch;
        const int next = item + item_stride;
        if (next < p.n_items) {
          // K of the next item goes into the other K buffer right away: its previous tenant (item it-1) retired its
          // last dQ GEMM before this item started
          decode(next, kv_base, head, batch);
          if (it >= 1) mbar_wait(&sm.dkv_full, (it - 1) & 1);
          load_k((it + 1) & 1, kv_base, head, batch);
        }
        decode(item, kv_base, head, batch);
        for (int s = n_prefill; s < n_steps; ++s) {
          const int gs = it * n_steps + s, st = gs % kQStages;
          const uint32_t ph = (gs / kQStages) & 1;
          mbar_wait(&sm.q_empty[st], ph ^ 1);
          mbar_wait(&sm.do_empty[st], ph ^ 1);
          load_step(gs, s, head, batch);
        }
        if (next < p.n_items) {
          decode(next, kv_base, head, batch);
          mbar_wait(&sm.kvt_full, it & 1);     // V of this item has been copied into TMEM: its smem tile is free
          load_v(kv_base, head, batch);
          // the next item's first steps: their ring stages free up while this item's last steps retire
          for (int s = 0; s < n_prefill; ++s) {
            const int gs = (it + 1) * n_steps + s, st = gs % kQStages;
            const uint32_t ph = (gs / kQStages) & 1;
            mbar_wait(&sm.q_empty[st], ph ^ 1);
            mbar_wait(&sm.do_empty[st], ph ^ 1);
            load_step(gs, s, head, batch);
          }
        }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_nt = make_idesc_bf16(kTile, kStep, 0, 0);       // S^T, dP^T : 128 x 64
      constexpr uint32_t idesc_kmn = make_idesc_bf16(kTile, kHeadDim, 0, 1);   // dV, dK: A K-major, B MN-major
      constexpr uint32_t idesc_mnmn = make_idesc_bf16(kTile, kHeadDim, 1, 1);  // dQ: A MN-major, B MN-major
      // loop-invariant operand descriptors (per k-step only the 14-bit start address field advances)
      const uint64_t d_k0 = make_smem_desc(smem_u32(sm.k[0]), 16, 1024, kLayoutSW128);        // the two K buffers are contiguous
      const uint64_t d_q0 = make_smem_desc(smem_u32(sm.q[0]), 16, 1024, kLayoutSW128);      // stages are contiguous
      const uint64_t d_do0 = make_smem_desc(smem_u32(sm.dout[0]), 16, 1024, kLayoutSW128);
      const uint64_t d_ds_mn0 = make_smem_desc(smem_u32(sm.ds[0]), kTileBytes, 1024, kLayoutSW128);
      const uint32_t zeros_addr = smem_u32(sm.aug_zeros);
      auto aug_desc = [&](const void* tile) {     // k-half 0 in the tile, k-half 1 = the shared zeros
        return make_smem_desc(smem_u32(tile), zeros_addr - smem_u32(tile), kAugSbo, 0);
      };
      const uint64_t d_ones = aug_desc(sm.ones);
      const uint64_t d_lse0 = aug_desc(sm.lse_aug[0]), d_d0 = aug_desc(sm.d_aug[0]);   // stages are contiguous
      // stage st: the start address moves up by st KB and the distance to the zeros shrinks by as much
      auto aug_stage = [](uint64_t d0, int st) {
        const uint64_t step = static_cast<uint64_t>(st) * (kAugBytes >> 4);
        return d0 + step - (step << 16);
      };

      auto wait_sdp_operands = [&](int gs) {   // Q (+ row terms) and dO of step gs have landed
        mbar_wait(&sm.q_full[gs % kQStages], (gs / kQStages) & 1);
        mbar_wait(&sm.do_full[gs % kQStages], (gs / kQStages) & 1);
      };
      auto issue_sdp = [&](int gs) {           // S^T = K Q^T, dP^T = V dO^T of step gs into buffer gs & 1
        const int st = gs % kQStages, b = gs & 1;
        const uint64_t dq_s = desc_advance(d_q0, st * kStepBytes), ddo_s = desc_advance(d_do0, st * kStepBytes);
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk)   // A = K from TMEM (8 packed columns per 16-wide k-step)
          umma_ts(tmem + kTmemS + b * kStep, tmem + kTmemK + kk * 8, desc_advance(dq_s, kk * 32), idesc_nt,
                  kk > 0 ? 1u : 0u);
        umma_ss(tmem + kTmemS + b * kStep, d_ones, aug_stage(d_lse0, st), idesc_nt, 1u);   // - lse/scale
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk)   // A = V from TMEM
          umma_ts(tmem + kTmemDP + b * kStep, tmem + kTmemV + kk * 8, desc_advance(ddo_s, kk * 32), idesc_nt,
                  kk > 0 ? 1u : 0u);
        umma_ss(tmem + kTmemDP + b * kStep, d_ones, aug_stage(d_d0, st), idesc_nt, 1u);    // - D
        umma_commit(&sm.sdp_full[b]);
      };

      int it = 0;
      for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
        const int gs0 = it * n_steps, gi0 = it * n_tiles;
        const uint64_t d_k = desc_advance(d_k0, (it & 1) * kTileBytes);
        mbar_wait(&sm.kvt_full, it & 1);     // K, V of this item copied into TMEM by the compute warps
        wait_sdp_operands(gs0);
        tc_fence_after();
        issue_sdp(gs0);
        if (n_steps > 1) {
          wait_sdp_operands(gs0 + 1);
          tc_fence_after();
          issue_sdp(gs0 + 1);
        }

        for (int s = 0; s < n_steps; ++s) {
          const int gs = gs0 + s, st = gs % kQStages, b = gs & 1, gi = gi0 + (s >> 1);
          // every wait that does not depend on the compute warps comes first, so its latency hides behind their work
          // and the GEMMs below go out back to back once P^T / dS^T of the step are ready
          if (s + 2 < n_steps) wait_sdp_operands(gs + 2);
          if (s & 1) mbar_wait(&sm.dq_empty, (gi & 1) ^ 1);
          if (s == 0 && it > 0) mbar_wait(&sm.dkv_drained, (it - 1) & 1);   // previous item's dV/dK left TMEM
          mbar_wait(&sm.pds_full[b], (gs >> 1) & 1);
          if (it == 0) LCBI_TR(2, s, 0);
          tc_fence_after();
          const uint64_t dq_s = desc_advance(d_q0, st * kStepBytes), ddo_s = desc_advance(d_do0, st * kStepBytes);
          // dV += P^T(s) dO_s   (A = bf16 P^T, in place in the S buffer: 16 queries = 8 columns per k-step, the
          //                      two 32-query halves start at columns 0 and 32)
#pragma unroll
          for (int kk = 0; kk < kStep / 16; ++kk)
            umma_ts(tmem + kTmemDV, tmem + kTmemS + b * kStep + (kk >> 1) * 32 + (kk & 1) * 8,
                    desc_advance(ddo_s, kk * 2048), idesc_kmn, (s > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&sm.do_empty[st]);
          // dK += dS^T(s) Q_s   (A = bf16 dS^T, in place in the dP buffer)
#pragma unroll
          for (int kk = 0; kk < kStep / 16; ++kk)
            umma_ts(tmem + kTmemDK, tmem + kTmemDP + b * kStep + (kk >> 1) * 32 + (kk & 1) * 8,
                    desc_advance(dq_s, kk * 2048), idesc_kmn, (s > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&sm.q_empty[st]);
          if (it == 0) LCBI_TR(2, s, 1);
          if (s + 2 < n_steps) issue_sdp(gs + 2);
          if (it == 0) LCBI_TR(2, s, 2);
          if (s & 1) {
            // dQ_i = dS(i) K over the whole 128-query tile (A = dS^T read MN-major: both atoms)
            const uint64_t dds_mn = desc_advance(d_ds_mn0, (gi & 1) * 2 * kTileBytes);
#pragma unroll
            for (int kk = 0; kk < kTile / 16; ++kk)
              umma_ss(tmem + kTmemDQ, desc_advance(dds_mn, kk * 2048), desc_advance(d_k, kk * 2048), idesc_mnmn,
                      kk > 0 ? 1u : 0u);
            umma_commit(&sm.dq_full);
            if (it == 0) LCBI_TR(2, s, 3);
          }
        }
        umma_commit(&sm.dkv_full);
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ------------------------------------------------------------------ dQ drain (warps 8-11)
    const int row = (warp & 3) * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_dq = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kTmemDQ;
    const bool issuer = (tid == 8 * 32);
    int it = 0;
    for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
      int kv_base, head, batch;
      decode(item, kv_base, head, batch);
      for (int i = 0; i < n_tiles; ++i) {
        const int gi = it * n_tiles + i;
        mbar_wait(&sm.dq_full, gi & 1);
        if (issuer && it == 0) LCBI_TR(3, i, 0);
        tc_fence_after();
        uint32_t r[64];
        tmem_ld_x32(t_dq, r);
        tmem_ld_x32(t_dq + 32, r + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.dq_empty);
        if (issuer) tma_store_wait_read<0>();   // the previous reduce has finished reading the staging tiles
        named_bar_sync(3, 128);
#pragma unroll
        for (int c = 0; c < 16; ++c) {          // 16-byte chunk c of the 256-byte fp32 row (softmax scale folded in)
          uint4 val = make_uint4(__float_as_uint(__uint_as_float(r[c * 4]) * p.scale),
                                 __float_as_uint(__uint_as_float(r[c * 4 + 1]) * p.scale),
                                 __float_as_uint(__uint_as_float(r[c * 4 + 2]) * p.scale),
                                 __float_as_uint(__uint_as_float(r[c * 4 + 3]) * p.scale));
          *reinterpret_cast<uint4*>(sm.dq_stage + (c >> 3) * kTileBytes + sw128_offset(row, c & 7)) = val;
        }
        fence_proxy_async_smem();
        named_bar_sync(4, 128);
        if (issuer) {
          tma_reduce_add_4d(&tm_dqacc, sm.dq_stage, 0, head, i * kTile, batch);
          tma_reduce_add_4d(&tm_dqacc, sm.dq_stage + kTileBytes, 32, head, i * kTile, batch);
          tma_store_commit();
        }
        if (issuer && it == 0) LCBI_TR(3, i, 1);
      }
    }
    if (issuer) tma_store_wait_read<0>();
    (void)row;
  } else if (warp < 8) {
    // ------------------------------------------------------------------ P^T and dS^T in one pass (warps 0-7)
    const int hh = warp >> 2;                   // which 32-query half of the 64-query step this thread handles
    const int row = (warp & 3) * 32 + lane;     // key row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float c = p.scale_log2;
    const bool store_issuer = (tid & 127) == 0;
    bool store_pending = false;                 // the previous item's dV/dK store may still be reading its staging tiles
    auto finish_store = [&]() {
      if (store_issuer) tma_store_wait_read<0>();
      named_bar_sync(7, 256);
      store_pending = false;
    };

    // K (warps 0-3) and V (warps 4-7) rows of item `it_kv` -> TMEM: the A operands of every S^T / dP^T GEMM of that
    // item. Callers guarantee that the previous item's S^T / dP^T GEMMs have all retired.
    auto copy_kv_to_tmem = [&](int it_kv) {
      if (hh) mbar_wait(&sm.v_full, it_kv & 1);
      else mbar_wait(&sm.k_full[it_kv & 1], (it_kv >> 1) & 1);
      const uint8_t* src = hh ? sm.v : sm.k[it_kv & 1];
      uint32_t kv[32];
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint4 x = *reinterpret_cast<const uint4*>(src + sw128_offset(row, c16));
        kv[c16 * 4 + 0] = x.x; kv[c16 * 4 + 1] = x.y; kv[c16 * 4 + 2] = x.z; kv[c16 * 4 + 3] = x.w;
      }
      tmem_st_x32(tmem + lane_sel + (hh ? kTmemV : kTmemK), kv);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.kvt_full);
    };

    int it = 0;
    for (int item = first_item; item < p.n_items; item += item_stride, ++it) {
      int kv_base, head, batch;
      decode(item, kv_base, head, batch);
      const int gs0 = it * n_steps, gi0 = it * n_tiles;

      LCBI_ITEM_T(0);
      if (it == 0) copy_kv_to_tmem(0);           // later items: done at the end of the previous item's last step

      for (int s = 0; s < n_steps; ++s) {
        const int gs = gs0 + s, b = gs & 1, gi = gi0 + (s >> 1);
        // the staging tiles of the previous item's dV/dK store live in the dS^T buffer of this item's SECOND tile
        // (both buffers when the store was an fp32 accumulate): make sure they were read before overwriting them
        if (s == 1) LCBI_ITEM_T(1);
        if (s == 2) LCBI_ITEM_T(2);
        if (store_pending && (s == 2 || p.accumulate_dkv)) finish_store();
        if (lane == 0 && it == 0) LCBI_TR(hh, s, 0);
        // One wait per step: the commit also covers dQ of two tiles ago, the last reader of this dS^T smem buffer.
        mbar_wait(&sm.sdp_full[b], (gs >> 1) & 1);
        if (lane == 0 && it == 0) LCBI_TR(hh, s, 1);
        tc_fence_after();
        uint8_t* ds_atom = sm.ds[gi & 1] + (s & 1) * kTileBytes;
        uint32_t sv[32], dpv[32];
        tmem_ld_x32(tmem + lane_sel + kTmemS + b * kStep + hh * 32, sv);
        tmem_ld_x32(tmem + lane_sel + kTmemDP + b * kStep + hh * 32, dpv);
        tmem_ld_wait();
        if (lane == 0 && it == 0) LCBI_TR(hh, s, 2);
        uint32_t pk[16], dsk[16];
        // the per-query terms were already added by the tensor core: sv = q.k - lse/scale, dpv = dO.v - D
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = fast_exp2(__uint_as_float(sv[e]) * c), p1 = fast_exp2(__uint_as_float(sv[e + 1]) * c);
          pk[e >> 1] = pack_bf16x2(p0, p1);
          dsk[e >> 1] = pack_bf16x2(p0 * __uint_as_float(dpv[e]), p1 * __uint_as_float(dpv[e + 1]));
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(ds_atom + sw128_offset(row, hh * 4 + g))),
                       "r"(dsk[g * 4]), "r"(dsk[g * 4 + 1]), "r"(dsk[g * 4 + 2]), "r"(dsk[g * 4 + 3]) : "memory");
        // in place: the packed results overwrite the first 16 of the 32 columns this thread just read
        tmem_st_x16(tmem + lane_sel + kTmemS + b * kStep + hh * 32, pk);
        tmem_st_x16(tmem + lane_sel + kTmemDP + b * kStep + hh * 32, dsk);
        if (lane == 0 && it == 0) LCBI_TR(hh, s, 3);
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.pds_full[b]);
        if (lane == 0 && it == 0) LCBI_TR(hh, s, 4);
      }

      // The S^T / dP^T GEMMs of this item have all retired (this thread consumed the last of them), so K and V of the
      // next item can take their place in TMEM now: its first GEMMs then run while this item's epilogue drains dV / dK.
      LCBI_ITEM_T(3);
      if (item + item_stride < p.n_items) copy_kv_to_tmem(it + 1);
      LCBI_ITEM_T(4);

      // ---- item epilogue: warps 0-3 drain dV, warps 4-7 drain dK (scaled)
      mbar_wait(&sm.dkv_full, it & 1);          // every GEMM of the item has retired
      tc_fence_after();
      uint32_t r[64];
      const uint32_t t_src = tmem + lane_sel + (hh ? kTmemDK : kTmemDV);
      tmem_ld_x32(t_src, r);
      tmem_ld_x32(t_src + 32, r + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      LCBI_ITEM_T(5);
      if (lane == 0) mbar_arrive(&sm.dkv_drained);   // the next item's first dV / dK GEMM may overwrite TMEM
      if (store_pending) finish_store();             // (items with fewer than three steps)
      const float mul = hh ? p.scale : 1.0f;
      const CUtensorMap* tm = hh ? &tm_dk : &tm_dv;
      if (!p.accumulate_dkv) {
        // staging: the dS^T buffer that the NEXT item's first tile does not use (free now, every dQ GEMM has retired)
        uint8_t* stage = sm.ds[(gi0 + n_tiles + 1) & 1] + hh * kTileBytes;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          uint4 val;
          val.x = pack_bf16x2(__uint_as_float(r[c16 * 8 + 0]) * mul, __uint_as_float(r[c16 * 8 + 1]) * mul);
          val.y = pack_bf16x2(__uint_as_float(r[c16 * 8 + 2]) * mul, __uint_as_float(r[c16 * 8 + 3]) * mul);
          val.z = pack_bf16x2(__uint_as_float(r[c16 * 8 + 4]) * mul, __uint_as_float(r[c16 * 8 + 5]) * mul);
          val.w = pack_bf16x2(__uint_as_float(r[c16 * 8 + 6]) * mul, __uint_as_float(r[c16 * 8 + 7]) * mul);
          *reinterpret_cast<uint4*>(stage + sw128_offset(row, c16)) = val;
        }
        fence_proxy_async_smem();
        named_bar_sync(5 + hh, 128);
        if (store_issuer) {
          tma_store_4d(tm, stage, 0, head, kv_base, batch);
          tma_store_commit();
        }
      } else {
        uint8_t* stage = sm.ds[hh];              // fp32: two [128 x 32] tiles per matrix
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {
          uint4 val = make_uint4(__float_as_uint(__uint_as_float(r[cc * 4]) * mul),
                                 __float_as_uint(__uint_as_float(r[cc * 4 + 1]) * mul),
                                 __float_as_uint(__uint_as_float(r[cc * 4 + 2]) * mul),
                                 __float_as_uint(__uint_as_float(r[cc * 4 + 3]) * mul));
          *reinterpret_cast<uint4*>(stage + (cc >> 3) * kTileBytes + sw128_offset(row, cc & 7)) = val;
        }
        fence_proxy_async_smem();
        named_bar_sync(5 + hh, 128);
        if (store_issuer) {
          tma_reduce_add_4d(tm, stage, 0, head, kv_base, batch);
          tma_reduce_add_4d(tm, stage + kTileBytes, 32, head, kv_base, batch);
          tma_store_commit();
        }
      }
      store_pending = true;
      LCBI_ITEM_T(6);
    }
    if (store_pending && store_issuer) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) tmem_dealloc(tmem, 512);
}

int make_bf16_map(CUtensorMap* m, const void* base, int B, int H, int N, const int64_t* st, int box_rows = kTile) {
  const uint64_t dims[4] = {static_cast<uint64_t>(kHeadDim), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(st[2]) * 2, static_cast<uint64_t>(st[1]) * 2,
                               static_cast<uint64_t>(st[0]) * 2};
  const uint32_t box[4] = {kHeadDim, 1, static_cast<uint32_t>(box_rows), 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// fp32 (B, N, H, 64) contiguous accumulator, box = 32 floats (128 B) x 128 rows
int make_f32_acc_map(CUtensorMap* m, const void* base, int B, int H, int N) {
  const uint64_t dims[4] = {static_cast<uint64_t>(kHeadDim), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(kHeadDim) * 4, static_cast<uint64_t>(H) * kHeadDim * 4,
                               static_cast<uint64_t>(N) * H * kHeadDim * 4};
  const uint32_t box[4] = {32, 1, kTile, 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

#ifdef LCBI_TRACE
extern "C" int lcbi_debug_set_bwd_trace(long long* ptr) {
  return static_cast<int>(cudaMemcpyToSymbol(g_bwd_trace, &ptr, sizeof(ptr)));
}
extern "C" int lcbi_debug_set_bwd_item_times(long long* ptr) {
  return static_cast<int>(cudaMemcpyToSymbol(g_bwd_item_times, &ptr, sizeof(ptr)));
}
#endif

// Workspace = [fp32 dQ accumulator (B,Nq,H,64), absent when the caller accumulates into its own] [-lse/scale tiles] [-D tiles]
size_t dense_attn_bwd_workspace_bytes(int B, int H, int Nq, int head_dim, int accumulate_dq) {
  const size_t nq_pad = align_up(static_cast<size_t>(Nq), kTile);
  const size_t acc = accumulate_dq ? 0 : align_up(static_cast<size_t>(B) * Nq * H * head_dim * 4, 128);
  const size_t vec = align_up(static_cast<size_t>(B) * H * (nq_pad / kStep) * kAugBytes, 128);
  return acc + 2 * vec;
}

int dense_attn_bwd_launch(const DenseAttnBwdArgs& a, cudaStream_t stream) {
  if (a.head_dim != kHeadDim) return LCBI_ERR_UNSUPPORTED;
  if (a.B <= 0 || a.H <= 0 || a.Nq <= 0 || a.Nk <= 0) return LCBI_ERR_BAD_ARG;
  const int64_t* all_strides[] = {a.q_strides, a.k_strides, a.v_strides, a.o_strides,
                                  a.do_strides, a.dq_strides, a.dk_strides, a.dv_strides};
  for (const int64_t* s : all_strides)
    for (int i = 0; i < 3; ++i)
      if (s[i] % 8 != 0) return LCBI_ERR_BAD_ARG;
  const void* ptrs[] = {a.q, a.k, a.v, a.o, a.d_o, a.dq, a.dk, a.dv};
  for (const void* ptr : ptrs)
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return LCBI_ERR_BAD_ARG;
  if (a.workspace_bytes < dense_attn_bwd_workspace_bytes(a.B, a.H, a.Nq, a.head_dim, a.accumulate_dq) ||
      (reinterpret_cast<uintptr_t>(a.workspace) & 127))
    return LCBI_ERR_WORKSPACE;

  const int nq_pad = static_cast<int>(align_up(static_cast<size_t>(a.Nq), kTile));
  const size_t acc_bytes = a.accumulate_dq ? 0 : align_up(static_cast<size_t>(a.B) * a.Nq * a.H * kHeadDim * 4, 128);
  const size_t vec_bytes = align_up(static_cast<size_t>(a.B) * a.H * (nq_pad / kStep) * kAugBytes, 128);
  float* dq_acc = a.accumulate_dq ? reinterpret_cast<float*>(a.dq) : reinterpret_cast<float*>(a.workspace);
  uint8_t* lse_aug = reinterpret_cast<uint8_t*>(a.workspace) + acc_bytes;
  uint8_t* d_aug = reinterpret_cast<uint8_t*>(a.workspace) + acc_bytes + vec_bytes;
  if (!(a.scale > 0.f)) return LCBI_ERR_BAD_ARG;

  CUtensorMap tq, tk, tv, tdo, tacc, tdk, tdv;
  if (make_bf16_map(&tq, a.q, a.B, a.H, a.Nq, a.q_strides, kStep) || make_bf16_map(&tk, a.k, a.B, a.H, a.Nk, a.k_strides) ||
      make_bf16_map(&tv, a.v, a.B, a.H, a.Nk, a.v_strides) || make_bf16_map(&tdo, a.d_o, a.B, a.H, a.Nq, a.do_strides, kStep) ||
      make_f32_acc_map(&tacc, dq_acc, a.B, a.H, a.Nq))
    return LCBI_ERR_TENSOR_MAP;
  if (a.accumulate_dkv) {
    if (make_f32_acc_map(&tdk, a.dk, a.B, a.H, a.Nk) || make_f32_acc_map(&tdv, a.dv, a.B, a.H, a.Nk))
      return LCBI_ERR_TENSOR_MAP;
  } else {
    if (make_bf16_map(&tdk, a.dk, a.B, a.H, a.Nk, a.dk_strides) || make_bf16_map(&tdv, a.dv, a.B, a.H, a.Nk, a.dv_strides))
      return LCBI_ERR_TENSOR_MAP;
  }

  cudaError_t e = cudaSuccess;
  {
    const int64_t rows = static_cast<int64_t>(a.B) * a.H * nq_pad;
    const int threads = 256;
    const int64_t blocks = (rows * 8 + threads - 1) / threads;
    bwd_prep_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(a.o), reinterpret_cast<const __nv_bfloat16*>(a.d_o), a.lse, lse_aug, d_aug,
        a.accumulate_dq ? nullptr : dq_acc, 1.0f / a.scale, a.B, a.H, a.Nq, nq_pad, a.o_strides[0], a.o_strides[1], a.o_strides[2], a.do_strides[0], a.do_strides[1],
        a.do_strides[2]);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e);
  }

  static unsigned long long configured = 0;   // one bit per device ordinal
  const int smem_bytes = static_cast<int>(sizeof(BwdSmem)) + 1024;
  if (first_launch_on_current_device(&configured)) {
    e = cudaFuncSetAttribute(dense_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      configured = 0;
      return set_cuda_error(e);
    }
  }
  BwdParams p;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.Nq_pad = nq_pad;
  p.scale = a.scale;
  p.scale_log2 = a.scale * kLog2e;
  p.lse_aug = reinterpret_cast<const __nv_bfloat16*>(lse_aug);
  p.d_aug = reinterpret_cast<const __nv_bfloat16*>(d_aug);
  p.dq_acc = dq_acc;
  p.accumulate_dkv = a.accumulate_dkv;
  const int num_sms = current_device_sm_count();
  if (num_sms <= 0) return LCBI_ERR_CUDA;
  p.n_kv_tiles = (a.Nk + kTile - 1) / kTile;
  p.n_items = p.n_kv_tiles * a.H * a.B;
  const int slots = num_sms - reserved_sms() > 1 ? num_sms - reserved_sms() : 1;
  dim3 grid(p.n_items < slots ? p.n_items : slots);
  dense_attn_bwd_kernel<<<grid, kNumThreads, smem_bytes, stream>>>(tq, tk, tv, tdo, tacc, tdk, tdv, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);

  if (!a.accumulate_dq) {
    const int64_t total = static_cast<int64_t>(a.B) * a.Nq * a.H * (kHeadDim / 8);
    const int threads = 256;
    bwd_finish_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
        dq_acc, reinterpret_cast<__nv_bfloat16*>(a.dq), a.B, a.H, a.Nq, a.dq_strides[0], a.dq_strides[1],
        a.dq_strides[2]);
    e = cudaGetLastError();
  }
  return set_cuda_error(e);
}

}  // namespace lcbi
