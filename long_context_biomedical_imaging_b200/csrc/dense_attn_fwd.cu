// Dense (ViT global) multi-head self-attention forward for sm_100a, head_dim 64, bf16 in / fp32 accumulate.
//
// Replaces the attention core of the reference's SABlock.forward
// (/root/reference/model/models/backbone_vit.py:191-201): softmax(scale * q k^T) v, without ever
// materialising the B*H*N*N matrix and reading q/k/v straight out of the (B,N,3,H,d) qkv tensor.
//
// Work item = one 128-row query tile of one (batch, head). The kernel is PERSISTENT: CTA c handles items c, c + grid,
// c + 2 grid, ... and two CTAs are resident per SM (192 threads, 256 TMEM columns, ~97 KB smem each), so that
//   * the two tiles on an SM run out of phase and share the MUFU and tensor pipes instead of convoying,
//   * an item's prologue (Q load, first S) and epilogue (O scale + store) overlap the other CTA's main loop, and the
//     next item's Q tile and K/V tiles are prefetched while the current item finishes (measured: a non-persistent
//     one-CTA-per-SM version spent ~5 us of every ~28 us item in launch / prologue / epilogue).
// Warp roles:
//   warp 5      TMA producer   (Q per item, double-buffered; K_j / V_j through 4-stage rings that run across items)
//   warp 4      UMMA issuer    (S = Q K_j^T -> TMEM;  O += P V_j with P read from TMEM)
//   warps 0-3   softmax + epilogue (one thread = one query row = one TMEM lane)
// TMEM: S is double-buffered, S_a [0,64) S_b [64,128), so the tensor core computes S(j+1) while the softmax warps
// exponentiate S(j); O [128,192); P (bf16) overwrites the first 32 columns of the S buffer it came from.
// Online softmax uses a lazily updated running max (only moved when the new max exceeds it by 2^8) so the O rescale
// in TMEM is rare. All mbarrier phases are derived from counters that keep running across items.
#include "lcbi_kernels.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace lcbi {

namespace {

constexpr int kBlockM = 128;      // rows per query tile
constexpr int kBlockN = 64;       // keys per KV tile
constexpr int kHeadDim = 64;
constexpr int kStages = 4;        // K and V ring depth
constexpr int kQTileBytes = kBlockM * kHeadDim * 2;   // 16 KB
constexpr int kKVTileBytes = kBlockN * kHeadDim * 2;  // 8 KB
constexpr int kSoftmaxWarps = 4, kMmaWarp = 4, kTmaWarp = 5;
constexpr int kNumThreads = 192;
constexpr int kCtasPerSm = 2;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemO = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

__device__ __forceinline__ constexpr uint32_t tmem_s(int buf) { return buf * 64; }

struct __align__(1024) FwdSmem {
  uint8_t q[2][kQTileBytes];        // Q of item it in q[it & 1]; reused as the O staging tile of that item
  uint8_t k[kStages][kKVTileBytes];
  uint8_t v[kStages][kKVTileBytes];
  uint64_t q_full[2], q_free[2];
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], p_full[2], pv_done, o_full;
  uint32_t tmem_base;
};

#ifdef LCBI_TRACE
// debug build only (tools/trace_dense.py): per-role clock64 timestamps of the first item of CTA 0
__device__ long long* g_fwd_trace = nullptr;
#define LCBI_TR_INIT() long long* const lcbi_tr = (blockIdx.x == 0) ? g_fwd_trace : nullptr
#define LCBI_TR(role, step, ev)                                                                       \
  do {                                                                                                \
    if (lcbi_tr != nullptr && it == 0 && (step) < 32) lcbi_tr[((role) * 32 + (step)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define LCBI_TR_INIT() do { } while (0)
#define LCBI_TR(role, step, ev) do { } while (0)
#endif

struct FwdParams {
  int B, H, Nq, Nk;
  int n_q_tiles, n_items;
  float scale_log2;   // scale * log2(e)
  float* lse;         // (B, H, Nq) natural log
  // Ring steps (lcbi_dense_attn_fwd_state): the online-softmax state of every query row is carried across launches,
  // one launch per visiting K/V shard, so the shards fold into ONE running result without a merge pass.
  float* st_o;        // fp32 (B, Nq, H, 64) un-normalised running output, or nullptr (plain call)
  float* st_m;        // fp32 (B, H, Nq) running max actually subtracted, raw score units
  float* st_l;        // fp32 (B, H, Nq) running row sum
  int carry_in;       // 1: start every row from (st_o, st_m, st_l) instead of (0, -inf, 0)
  int carry_out;      // 1: write the state back instead of the normalised bf16 output and lse
};

// Timing-only ablations (results become wrong; every barrier still fires): bit 0 no exp2, bit 1 no row-max pass,
// bit 2 softmax warps only wait and arrive, bit 3 bf16 packing by truncation (one PRMT) instead of F2FP.
#ifndef LCBI_FWD_ABLATE
#define LCBI_FWD_ABLATE 0
#endif
constexpr int kAblate = LCBI_FWD_ABLATE;

// kCarry: the ring-step variant (state carried in / out); the plain kernel compiles without any of that code
template <bool kCarry>
__global__ void __launch_bounds__(kNumThreads, kCtasPerSm)
dense_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int n_kv = (p.Nk + kBlockN - 1) / kBlockN;
  LCBI_TR_INIT();

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sm.q_full[b], 1);
      mbar_init(&sm.q_free[b], 1);
      mbar_init(&sm.s_full[b], 1);
      mbar_init(&sm.p_full[b], kSoftmaxWarps);    // one arrive per softmax warp
    }
    mbar_init(&sm.pv_done, 1);
    mbar_init(&sm.o_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.k_full[s], 1);
      mbar_init(&sm.k_empty[s], 1);
      mbar_init(&sm.v_full[s], 1);
      mbar_init(&sm.v_empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == kMmaWarp) {
    tmem_alloc(&sm.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (warp == kTmaWarp && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  // item -> (batch, head, first query row): adjacent items are adjacent query tiles of one (batch, head), so the CTAs
  // running at any moment read a small set of K/V tensors (L2-resident)
  auto decode = [&](int item, int& batch, int& head, int& q_base) {
    const int tile = item % p.n_q_tiles, bh = item / p.n_q_tiles;
    head = bh % p.H;
    batch = bh / p.H;
    q_base = tile * kBlockM;
  };

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int g = 0;   // K/V tiles issued so far (ring position), runs across items
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        int batch, head, q_base;
        decode(item, batch, head, q_base);
        const int qb = it & 1;
        if (it >= 2) mbar_wait(&sm.q_free[qb], ((it >> 1) - 1) & 1);   // O store of item it-2 has left the buffer
        mbar_expect_tx(&sm.q_full[qb], kQTileBytes);
        tma_load_4d(sm.q[qb], &tm_q, &sm.q_full[qb], 0, head, q_base, batch);
        for (int j = 0; j < n_kv; ++j, ++g) {
          const int s = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          mbar_wait(&sm.k_empty[s], ph ^ 1);
          mbar_expect_tx(&sm.k_full[s], kKVTileBytes);
          tma_load_4d(sm.k[s], &tm_k, &sm.k_full[s], 0, head, j * kBlockN, batch);
          mbar_wait(&sm.v_empty[s], ph ^ 1);
          mbar_expect_tx(&sm.v_full[s], kKVTileBytes);
          tma_load_4d(sm.v[s], &tm_v, &sm.v_full[s], 0, head, j * kBlockN, batch);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);   // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);  // P V   : V is MN-major
      int g0 = 0;  // global index of this item's first K/V tile
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it, g0 += n_kv) {
        const uint32_t q_addr = smem_u32(sm.q[it & 1]);
        auto issue_s = [&](int g) {               // S(g) -> buffer g & 1
          const uint32_t k_addr = smem_u32(sm.k[g % kStages]);
#pragma unroll
          for (int kk = 0; kk < kHeadDim / 16; ++kk) {
            const uint64_t da = make_smem_desc(q_addr + kk * 32, 16, 1024, kLayoutSW128);
            const uint64_t db = make_smem_desc(k_addr + kk * 32, 16, 1024, kLayoutSW128);
            umma_ss(tmem + tmem_s(g & 1), da, db, idesc_s, kk > 0 ? 1u : 0u);
          }
          umma_commit(&sm.s_full[g & 1]);
        };
        auto issue_pv = [&](int g, bool first) {  // O (+)= P(g) V(g); with a carried-in state O is never overwritten
          const uint32_t v_addr = smem_u32(sm.v[g % kStages]);
#pragma unroll
          for (int kk = 0; kk < kBlockN / 16; ++kk) {
            const uint64_t db = make_smem_desc(v_addr + kk * 2048, 16, 1024, kLayoutSW128);
            umma_ts(tmem + kTmemO, tmem + tmem_s(g & 1) + kk * 8, db, idesc_o, (!first || kk > 0) ? 1u : 0u);
          }
          umma_commit(&sm.pv_done);
        };

        mbar_wait(&sm.q_full[it & 1], (it >> 1) & 1);
        for (int j0 = 0; j0 < 2 && j0 < n_kv; ++j0) {
          const int g = g0 + j0;
          mbar_wait(&sm.k_full[g % kStages], (g / kStages) & 1);
          tc_fence_after();
          issue_s(g);
          umma_commit(&sm.k_empty[g % kStages]);
        }
        for (int j = 0; j < n_kv; ++j) {
          const int g = g0 + j;
          const bool has_next = (j + 2) < n_kv;
          // the waits that do not depend on the softmax warps come first, so their latency hides behind that work
          mbar_wait(&sm.v_full[g % kStages], (g / kStages) & 1);
          if (has_next) mbar_wait(&sm.k_full[(g + 2) % kStages], ((g + 2) / kStages) & 1);
          mbar_wait(&sm.p_full[g & 1], (g >> 1) & 1);
          LCBI_TR(1, j, 0);
          tc_fence_after();
          issue_pv(g, j == 0 && !(kCarry && p.carry_in));
          if (has_next) issue_s(g + 2);
          else if (j == n_kv - 1) umma_commit(&sm.o_full);
          umma_commit(&sm.v_empty[g % kStages]);
          if (has_next) umma_commit(&sm.k_empty[(g + 2) % kStages]);
          LCBI_TR(1, j, 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int row = tid;                       // row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t t_o = tmem + lane_sel + kTmemO;
    const float c = p.scale_log2;
    int g0 = 0;
    int it = 0;
    int store_pending = -1;                    // Q/O buffer whose TMA store has not been waited for yet (thread 0)
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it, g0 += n_kv) {
      int batch, head, q_base;
      decode(item, batch, head, q_base);
      float m_used = -INFINITY;  // running max actually subtracted (raw score units)
      float l = 0.f;
      const int q_row_st = q_base + row;
      const bool row_valid = q_row_st < p.Nq;
      if (kCarry && p.carry_in) {
        // the state of the previous ring steps: O goes back into TMEM before the first P V of this item is issued (the
        // issuer waits for p_full of the first K/V tile, which this thread signals only after these stores)
        const size_t ml_idx = (static_cast<size_t>(batch) * p.H + head) * p.Nq + q_row_st;
        if (row_valid) {
          m_used = p.st_m[ml_idx];
          l = p.st_l[ml_idx];
        }
        const float4* src = reinterpret_cast<const float4*>(
            p.st_o + ((static_cast<size_t>(batch) * p.Nq + (row_valid ? q_row_st : 0)) * p.H + head) * kHeadDim);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = row_valid ? src[half * 8 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
            o[4 * i] = __float_as_uint(x.x); o[4 * i + 1] = __float_as_uint(x.y);
            o[4 * i + 2] = __float_as_uint(x.z); o[4 * i + 3] = __float_as_uint(x.w);
          }
          tmem_st_x32(t_o + half * 32, o);
        }
        tmem_st_wait();
      }

      for (int j = 0; j < n_kv; ++j) {
        const int g = g0 + j;
        const int buf = g & 1;
        const uint32_t t_s = tmem + lane_sel + tmem_s(buf);
        if (row == 0) LCBI_TR(0, j, 0);
        mbar_wait(&sm.s_full[buf], (g >> 1) & 1);
        if (row == 0) LCBI_TR(0, j, 1);
        tc_fence_after();
        if (kAblate & 4) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.p_full[buf]);
          l = 1.f;
          m_used = 0.f;
          continue;
        }
        uint32_t sr[64];
        tmem_ld_x32(t_s, sr);
        tmem_ld_x32(t_s + 32, sr + 32);
        tmem_ld_wait();
        if (row == 0) LCBI_TR(0, j, 2);

        const int valid = p.Nk - j * kBlockN;  // >= 1
        if (valid < kBlockN) {
#pragma unroll
          for (int i = 0; i < kBlockN; ++i)
            if (i >= valid) sr[i] = __float_as_uint(-INFINITY);
        }
        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]),
              mx3 = __uint_as_float(sr[3]);
#pragma unroll
        for (int i = 4; i < ((kAblate & 2) ? 4 : kBlockN); i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
        }
        const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);

        const bool need = (m_new - m_used) * c > kRescaleThreshold;  // true on the first tile (m_used = -inf)
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? fast_exp2((m_used - m_new) * c) : 1.0f;
          if (need) m_used = m_new;
          l *= alpha;
          if (j > 0 || (kCarry && p.carry_in)) {
            // O must be quiescent: P(g-1) V has completed (the S buffers are double-buffered, so s_full alone
            // does not imply it) and P(g) V is not issued before this warp group signals p_full. (j == 0 with a
            // carried-in state: O was just written by this very thread and no P V of the item exists yet.)
            if (j > 0) mbar_wait(&sm.pv_done, (g - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t o[32];
              tmem_ld_x32(t_o + half * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x32(t_o + half * 32, o);
            }
          }
        }

        if (row == 0) LCBI_TR(0, j, 3);
        const float mc = m_used * c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < kBlockN; i += 4) {
          auto ex = [](float x) { return (kAblate & 1) ? x : fast_exp2(x); };
          const float e0 = ex(fmaf(__uint_as_float(sr[i]), c, -mc));
          const float e1 = ex(fmaf(__uint_as_float(sr[i + 1]), c, -mc));
          const float e2 = ex(fmaf(__uint_as_float(sr[i + 2]), c, -mc));
          const float e3 = ex(fmaf(__uint_as_float(sr[i + 3]), c, -mc));
          s0 += e0; s1 += e1; s2 += e2; s3 += e3;
          pk[i / 2] = (kAblate & 8) ? __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632) : pack_bf16x2(e0, e1);
          pk[i / 2 + 1] = (kAblate & 8) ? __byte_perm(__float_as_uint(e2), __float_as_uint(e3), 0x7632) : pack_bf16x2(e2, e3);
        }
        l += (s0 + s1) + (s2 + s3);
        if (row == 0) LCBI_TR(0, j, 4);
        tmem_st_x32(t_s, pk);
        tmem_st_wait();
        if (row == 0) LCBI_TR(0, j, 5);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.p_full[buf]);
        if (row == 0) LCBI_TR(0, j, 6);

        if (j == 1 && tid == 0 && store_pending >= 0) {
          // the previous item's O store has certainly been read out of its staging tile by now: hand the buffer back
          // to the producer (it becomes the Q tile of the item after this one)
          tma_store_wait_read<0>();
          mbar_arrive(&sm.q_free[store_pending]);
          store_pending = -1;
        }
      }

      // ---- epilogue: O / l -> bf16 -> swizzled smem tile (the item's Q tile, dead by now) -> TMA store
      mbar_wait(&sm.o_full, it & 1);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_x32(t_o, o);
      tmem_ld_x32(t_o + 32, o + 32);
      tmem_ld_wait();
      tc_fence_before();
      if (tid == 0 && store_pending >= 0) {     // only when an item has fewer than two K/V tiles
        tma_store_wait_read<0>();
        mbar_arrive(&sm.q_free[store_pending]);
        store_pending = -1;
      }
      if (kCarry && p.carry_out) {
        // ring step that is not the last: the un-normalised state goes back to HBM, nothing is staged or TMA-stored,
        // and the item's Q tile (dead since o_full) returns to the producer right away
        if (tid == 0) mbar_arrive(&sm.q_free[it & 1]);
        if (row_valid) {
          const size_t ml_idx = (static_cast<size_t>(batch) * p.H + head) * p.Nq + q_row_st;
          p.st_m[ml_idx] = m_used;
          p.st_l[ml_idx] = l;
          float4* dst = reinterpret_cast<float4*>(
              p.st_o + ((static_cast<size_t>(batch) * p.Nq + q_row_st) * p.H + head) * kHeadDim);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            dst[i] = make_float4(__uint_as_float(o[4 * i]), __uint_as_float(o[4 * i + 1]), __uint_as_float(o[4 * i + 2]),
                                 __uint_as_float(o[4 * i + 3]));
        }
        continue;
      }
      const float inv_l = 1.0f / l;
      uint8_t* stage = sm.q[it & 1];
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        uint4 val;
        val.x = pack_bf16x2(__uint_as_float(o[c16 * 8 + 0]) * inv_l, __uint_as_float(o[c16 * 8 + 1]) * inv_l);
        val.y = pack_bf16x2(__uint_as_float(o[c16 * 8 + 2]) * inv_l, __uint_as_float(o[c16 * 8 + 3]) * inv_l);
        val.z = pack_bf16x2(__uint_as_float(o[c16 * 8 + 4]) * inv_l, __uint_as_float(o[c16 * 8 + 5]) * inv_l);
        val.w = pack_bf16x2(__uint_as_float(o[c16 * 8 + 6]) * inv_l, __uint_as_float(o[c16 * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stage + sw128_offset(row, c16)) = val;
      }
      const int q_row = q_base + row;
      if (q_row < p.Nq)
        p.lse[(static_cast<size_t>(batch) * p.H + head) * p.Nq + q_row] = (m_used * c + log2f(l)) * kLn2;
      fence_proxy_async_smem();
      named_bar_sync(1, kSoftmaxWarps * 32);
      if (tid == 0) {
        tma_store_4d(&tm_o, stage, 0, head, q_base, batch);
        tma_store_commit();
        store_pending = it & 1;
      }
    }
    if (tid == 0 && store_pending >= 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

// Builds the rank-4 (d, H, N, B) bf16 tensor map of a (B, N, H, d) strided view, box = (64, 1, 128, 1).
static int make_bhnd_map(CUtensorMap* m, const void* base, int B, int H, int N, int64_t batch_stride,
                         int64_t row_stride, int64_t head_stride, int box_rows) {
  const uint64_t dims[4] = {static_cast<uint64_t>(kHeadDim), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(head_stride) * 2, static_cast<uint64_t>(row_stride) * 2,
                               static_cast<uint64_t>(batch_stride) * 2};
  const uint32_t box[4] = {kHeadDim, 1, static_cast<uint32_t>(box_rows), 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int dense_attn_fwd_launch(const DenseAttnArgs& a, cudaStream_t stream) {
  if (a.head_dim != kHeadDim) return LCBI_ERR_UNSUPPORTED;
  if (a.B <= 0 || a.H <= 0 || a.Nq <= 0 || a.Nk <= 0) return LCBI_ERR_BAD_ARG;
  const int64_t strides_all[] = {a.q_strides[0], a.q_strides[1], a.q_strides[2], a.k_strides[0], a.k_strides[1],
                                 a.k_strides[2], a.v_strides[0], a.v_strides[1], a.v_strides[2], a.o_strides[0],
                                 a.o_strides[1], a.o_strides[2]};
  for (int64_t s : strides_all)
    if (s % 8 != 0) return LCBI_ERR_BAD_ARG;  // TMA needs 16-byte aligned strides
  if ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.k) | reinterpret_cast<uintptr_t>(a.v) |
       reinterpret_cast<uintptr_t>(a.o)) & 15)
    return LCBI_ERR_BAD_ARG;

  CUtensorMap tq, tk, tv, to;
  if (make_bhnd_map(&tq, a.q, a.B, a.H, a.Nq, a.q_strides[0], a.q_strides[1], a.q_strides[2], kBlockM) ||
      make_bhnd_map(&tk, a.k, a.B, a.H, a.Nk, a.k_strides[0], a.k_strides[1], a.k_strides[2], kBlockN) ||
      make_bhnd_map(&tv, a.v, a.B, a.H, a.Nk, a.v_strides[0], a.v_strides[1], a.v_strides[2], kBlockN) ||
      make_bhnd_map(&to, a.o, a.B, a.H, a.Nq, a.o_strides[0], a.o_strides[1], a.o_strides[2], kBlockM))
    return LCBI_ERR_TENSOR_MAP;

  static unsigned long long configured = 0;   // one bit per device ordinal
  const int smem_bytes = static_cast<int>(sizeof(FwdSmem)) + 1024;
  if (first_launch_on_current_device(&configured)) {
    cudaError_t e = cudaFuncSetAttribute(dense_attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      configured = 0;
      return set_cuda_error(e);
    }
  }
  const int num_sms = current_device_sm_count();
  if (num_sms <= 0) return LCBI_ERR_CUDA;
  FwdParams p;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk;
  p.n_q_tiles = (a.Nq + kBlockM - 1) / kBlockM;
  p.n_items = p.n_q_tiles * a.H * a.B;
  p.scale_log2 = a.scale * kLog2e;
  p.lse = a.lse;
  p.st_o = a.state_o; p.st_m = a.state_m; p.st_l = a.state_l;
  p.carry_in = (a.state_o != nullptr && !a.state_first) ? 1 : 0;
  p.carry_out = (a.state_o != nullptr && !a.state_last) ? 1 : 0;
  if (a.state_o != nullptr && (a.state_m == nullptr || a.state_l == nullptr)) return LCBI_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(a.state_o) & 15) return LCBI_ERR_BAD_ARG;
  const int slots = kCtasPerSm * (num_sms - reserved_sms() > 1 ? num_sms - reserved_sms() : 1);
  dim3 grid(p.n_items < slots ? p.n_items : slots);
  if (a.state_o != nullptr && !(a.state_first && a.state_last))
    dense_attn_fwd_kernel<true><<<grid, kNumThreads, smem_bytes, stream>>>(tq, tk, tv, to, p);
  else
    dense_attn_fwd_kernel<false><<<grid, kNumThreads, smem_bytes, stream>>>(tq, tk, tv, to, p);
  return set_cuda_error(cudaGetLastError());
}

#ifdef LCBI_TRACE
extern "C" int lcbi_debug_set_fwd_trace(long long* ptr) {
  return static_cast<int>(cudaMemcpyToSymbol(g_fwd_trace, &ptr, sizeof(ptr)));
}
#endif

}  // namespace lcbi
