// Dense (ViT global) multi-head self-attention forward for sm_100a, head_dim 64, bf16 in / fp32 accumulate.
//
// Replaces the attention core of the reference's SABlock.forward
// (/root/reference/model/models/backbone_vit.py:191-201): softmax(scale * q k^T) v, without ever
// materialising the B*H*N*N matrix and reading q/k/v straight out of the (B,N,3,H,d) qkv tensor.
//
// One CTA = two 128-row query tiles of one (batch, head) sharing every 64-key K/V tile:
//   warp 9      TMA producer   (Q0,Q1 once; K_j / V_j through 4-stage rings)
//   warp 8      UMMA issuer    (S_t = Q_t K_j^T -> TMEM;  O_t += P_t V_j with P_t read from TMEM)
//   warps 0-3   softmax for query tile 0 (one thread = one query row = one TMEM lane)
//   warps 4-7   softmax for query tile 1
// TMEM (512 columns): S_t is DOUBLE-BUFFERED per query tile — S0a [0,64) S0b [64,128) S1a [128,192) S1b [192,256) —
// so the tensor core computes S_t(j+1) while the softmax group exponentiates S_t(j); O0 [256,320) O1 [320,384);
// P_t (bf16) overwrites the first 32 columns of the S buffer it came from. The issuer polls the two groups'
// "P ready" barriers and serves whichever arrives first, so the groups drift out of phase and share the MUFU
// pipe instead of convoying. Online softmax uses a lazily updated running max (only moved when the new max
// exceeds it by 2^8) so the O rescale in TMEM is rare.
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
constexpr int kNumThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr bool kExpPingPong = false;
#ifndef LCBI_FWD_STAGGER_NS
#define LCBI_FWD_STAGGER_NS 0   /* measured: 350 ns stagger gives 0.280 ms vs 0.276 ms in lock-step: no gain */
#endif
constexpr unsigned kStaggerNs = LCBI_FWD_STAGGER_NS;

__device__ __forceinline__ constexpr uint32_t tmem_s(int t, int buf) { return t * 128 + buf * 64; }
__device__ __forceinline__ constexpr uint32_t tmem_o(int t) { return 256 + t * 64; }

struct __align__(1024) FwdSmem {
  uint8_t q[2][kQTileBytes];        // also the O staging tiles for the TMA store
  uint8_t k[kStages][kKVTileBytes];
  uint8_t v[kStages][kKVTileBytes];
  uint64_t q_full[2];
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2][2], p_full[2][2], pv_done[2], o_full[2];
  uint32_t tmem_base;
};

#ifdef LCBI_TRACE
// debug build only (tools/trace_dense.py): per-role clock64 timestamps of CTA (0,0,0)
__device__ long long* g_fwd_trace = nullptr;
#define LCBI_TR_INIT() \
  long long* const lcbi_tr = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_fwd_trace : nullptr
#define LCBI_TR(role, step, ev)                                                            \
  do {                                                                                     \
    if (lcbi_tr != nullptr && (step) < 32) lcbi_tr[((role) * 32 + (step)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define LCBI_TR_INIT() do { } while (0)
#define LCBI_TR(role, step, ev) do { } while (0)
#endif

struct FwdParams {
  int B, H, Nq, Nk;
  float scale_log2;   // scale * log2(e)
  float* lse;         // (B, H, Nq) natural log
};

__global__ void __launch_bounds__(kNumThreads, 1)
dense_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  // 1-D grid, heavy CTAs first: every (batch, head) has n_full query-tile pairs with two full tiles and possibly one
  // lighter trailing pair; scheduling the light ones last shortens the final partial wave.
  const int n_full = p.Nq / (2 * kBlockM);
  const int bh_total = p.B * p.H;
  int pair, bh;
  if (static_cast<int>(blockIdx.x) < n_full * bh_total) {
    pair = blockIdx.x % n_full;
    bh = blockIdx.x / n_full;
  } else {
    pair = n_full;
    bh = blockIdx.x - n_full * bh_total;
  }
  const int head = bh % p.H, batch = bh / p.H;
  const int q_base = pair * (2 * kBlockM);
  const bool tile1_active = (q_base + kBlockM) < p.Nq;
  const int n_kv = (p.Nk + kBlockN - 1) / kBlockN;
  LCBI_TR_INIT();

  if (tid == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sm.q_full[t], 1);
      mbar_init(&sm.pv_done[t], 1);
      mbar_init(&sm.o_full[t], 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&sm.s_full[t][b], 1);
        mbar_init(&sm.p_full[t][b], 4);    // one arrive per softmax warp
      }
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.k_full[s], 1);
      mbar_init(&sm.k_empty[s], 1);
      mbar_init(&sm.v_full[s], 1);
      mbar_init(&sm.v_empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 8) {
    tmem_alloc(&sm.tmem_base, 512);
    tmem_relinquish();
  }
  if (warp == 9 && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(&sm.q_full[0], kQTileBytes);
      tma_load_4d(sm.q[0], &tm_q, &sm.q_full[0], 0, head, q_base, batch);
      if (tile1_active) {
        mbar_expect_tx(&sm.q_full[1], kQTileBytes);
        tma_load_4d(sm.q[1], &tm_q, &sm.q_full[1], 0, head, q_base + kBlockM, batch);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        mbar_wait(&sm.k_empty[s], ph ^ 1);
        mbar_expect_tx(&sm.k_full[s], kKVTileBytes);
        tma_load_4d(sm.k[s], &tm_k, &sm.k_full[s], 0, head, j * kBlockN, batch);
        mbar_wait(&sm.v_empty[s], ph ^ 1);
        mbar_expect_tx(&sm.v_full[s], kKVTileBytes);
        tma_load_4d(sm.v[s], &tm_v, &sm.v_full[s], 0, head, j * kBlockN, batch);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);   // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);  // P V   : V is MN-major
      const uint32_t q_addr[2] = {smem_u32(sm.q[0]), smem_u32(sm.q[1])};
      const int n_tiles = tile1_active ? 2 : 1;

      auto issue_s = [&](int t, int j) {        // S_t(j) -> buffer j & 1
        const uint32_t k_addr = smem_u32(sm.k[j % kStages]);
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) {
          const uint64_t da = make_smem_desc(q_addr[t] + kk * 32, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(k_addr + kk * 32, 16, 1024, kLayoutSW128);
          umma_ss(tmem + tmem_s(t, j & 1), da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&sm.s_full[t][j & 1]);
      };
      auto issue_pv = [&](int t, int j) {       // O_t += P_t(j) V_j
        const uint32_t v_addr = smem_u32(sm.v[j % kStages]);
#pragma unroll
        for (int kk = 0; kk < kBlockN / 16; ++kk) {
          const uint64_t db = make_smem_desc(v_addr + kk * 2048, 16, 1024, kLayoutSW128);
          umma_ts(tmem + tmem_o(t), tmem + tmem_s(t, j & 1) + kk * 8, db, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&sm.pv_done[t]);
      };

      mbar_wait(&sm.q_full[0], 0);
      if (tile1_active) mbar_wait(&sm.q_full[1], 0);
      for (int j0 = 0; j0 < 2 && j0 < n_kv; ++j0) {
        mbar_wait(&sm.k_full[j0 % kStages], 0);
        tc_fence_after();
        for (int t = 0; t < n_tiles; ++t) issue_s(t, j0);
        umma_commit(&sm.k_empty[j0 % kStages]);
      }

      for (int j = 0; j < n_kv; ++j) {
        const bool has_next = (j + 2) < n_kv;
        mbar_wait(&sm.v_full[j % kStages], (j / kStages) & 1);
        if (has_next) mbar_wait(&sm.k_full[(j + 2) % kStages], ((j + 2) / kStages) & 1);
        uint32_t pending = (1u << n_tiles) - 1u;
        const long long t_start = clock64();
        while (pending) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if ((pending >> t) & 1u) {
              if (mbar_try_wait(&sm.p_full[t][j & 1], (j >> 1) & 1)) {
                LCBI_TR(2, j, t * 3);
                tc_fence_after();
                issue_pv(t, j);
                if (has_next) issue_s(t, j + 2);
                else if (j == n_kv - 1) umma_commit(&sm.o_full[t]);
                LCBI_TR(2, j, t * 3 + 1);
                pending &= ~(1u << t);
              }
            }
          }
          if (clock64() - t_start > LCBI_WATCHDOG_CYCLES) {
            printf("lcbi watchdog: fwd issuer stuck (block %d,%d,%d step %d pending %u)\n", blockIdx.x, blockIdx.y,
                   blockIdx.z, j, pending);
            __trap();
          }
        }
        umma_commit(&sm.v_empty[j % kStages]);
        if (has_next) umma_commit(&sm.k_empty[(j + 2) % kStages]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int t = warp >> 2;                   // query tile handled by this warpgroup
    const int row = tid & 127;                 // row inside the tile == TMEM lane
    if (t == 0 || tile1_active) {
      const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t t_o = tmem + lane_sel + tmem_o(t);
      const float c = p.scale_log2;
      float m_used = -INFINITY;  // running max actually subtracted (raw score units)
      float l = 0.f;
      // Optional strict alternation of the two groups' exp phases (named barriers 3 / 4 hand a token back and
      // forth). Measured on B200 (profiles/r01_fwd_trace_pingpong.log): one warp per scheduler cannot saturate the
      // MUFU pipe (725 cycles per 64 exps instead of 512), so letting both groups exponentiate concurrently is
      // faster (0.276 ms vs 0.300 ms at cfg3 B=16). Kept for experiments, off by default.
      const bool pingpong = kExpPingPong && tile1_active;
      if (pingpong && t == 1) named_bar_arrive(3, 256);

      for (int j = 0; j < n_kv; ++j) {
        const int buf = j & 1;
        const uint32_t t_s = tmem + lane_sel + tmem_s(t, buf);
        if (row == 0) LCBI_TR(t, j, 0);
        mbar_wait(&sm.s_full[t][buf], (j >> 1) & 1);
        if (kStaggerNs > 0 && t == 1 && j == 0) __nanosleep(kStaggerNs);   // one-off phase offset between the groups
        if (row == 0) LCBI_TR(t, j, 1);
        tc_fence_after();
        uint32_t sr[64];
        tmem_ld_x32(t_s, sr);
        tmem_ld_x32(t_s + 32, sr + 32);
        tmem_ld_wait();
        if (row == 0) LCBI_TR(t, j, 2);

        const int valid = p.Nk - j * kBlockN;  // >= 1
        if (valid < kBlockN) {
#pragma unroll
          for (int i = 0; i < kBlockN; ++i)
            if (i >= valid) sr[i] = __float_as_uint(-INFINITY);
        }
        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]),
              mx3 = __uint_as_float(sr[3]);
#pragma unroll
        for (int i = 4; i < kBlockN; i += 4) {
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
          if (j > 0) {
            // O_t must be quiescent: P_t(j-1) V has completed (the S buffers are double-buffered, so s_full
            // alone does not imply it) and P_t(j) V is not issued before this group signals p_full.
            mbar_wait(&sm.pv_done[t], (j - 1) & 1);
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

        if (pingpong) named_bar_sync(3 + t, 256);
        if (row == 0) LCBI_TR(t, j, 3);
        const float mc = m_used * c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < kBlockN; i += 4) {
          const float e0 = fast_exp2(fmaf(__uint_as_float(sr[i]), c, -mc));
          const float e1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), c, -mc));
          const float e2 = fast_exp2(fmaf(__uint_as_float(sr[i + 2]), c, -mc));
          const float e3 = fast_exp2(fmaf(__uint_as_float(sr[i + 3]), c, -mc));
          s0 += e0; s1 += e1; s2 += e2; s3 += e3;
          pk[i / 2] = pack_bf16x2(e0, e1);
          pk[i / 2 + 1] = pack_bf16x2(e2, e3);
        }
        l += (s0 + s1) + (s2 + s3);
        if (pingpong) named_bar_arrive(4 - t, 256);
        if (row == 0) LCBI_TR(t, j, 4);
        tmem_st_x32(t_s, pk);
        tmem_st_wait();
        if (row == 0) LCBI_TR(t, j, 5);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.p_full[t][buf]);
        if (row == 0) LCBI_TR(t, j, 6);
      }

      if (pingpong && t == 0) named_bar_sync(3, 256);   // absorb the partner's last token
      // ---- epilogue: O / l -> bf16 -> swizzled smem tile (reuses the Q tile) -> TMA store
      mbar_wait(&sm.o_full[t], 0);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_x32(t_o, o);
      tmem_ld_x32(t_o + 32, o + 32);
      tmem_ld_wait();
      const float inv_l = 1.0f / l;
      uint8_t* stage = sm.q[t];
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        uint4 val;
        val.x = pack_bf16x2(__uint_as_float(o[c16 * 8 + 0]) * inv_l, __uint_as_float(o[c16 * 8 + 1]) * inv_l);
        val.y = pack_bf16x2(__uint_as_float(o[c16 * 8 + 2]) * inv_l, __uint_as_float(o[c16 * 8 + 3]) * inv_l);
        val.z = pack_bf16x2(__uint_as_float(o[c16 * 8 + 4]) * inv_l, __uint_as_float(o[c16 * 8 + 5]) * inv_l);
        val.w = pack_bf16x2(__uint_as_float(o[c16 * 8 + 6]) * inv_l, __uint_as_float(o[c16 * 8 + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stage + sw128_offset(row, c16)) = val;
      }
      const int q_row = q_base + t * kBlockM + row;
      if (q_row < p.Nq)
        p.lse[(static_cast<size_t>(batch) * p.H + head) * p.Nq + q_row] = (m_used * c + log2f(l)) * kLn2;
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if ((tid & 127) == 0) {
        tma_store_4d(&tm_o, stage, 0, head, q_base + t * kBlockM, batch);
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
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

  static bool attr_set = false;
  const int smem_bytes = static_cast<int>(sizeof(FwdSmem)) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dense_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  FwdParams p;
  p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk;
  p.scale_log2 = a.scale * kLog2e;
  p.lse = a.lse;
  dim3 grid(((a.Nq + 2 * kBlockM - 1) / (2 * kBlockM)) * a.H * a.B);
  dense_attn_fwd_kernel<<<grid, kNumThreads, smem_bytes, stream>>>(tq, tk, tv, to, p);
  return set_cuda_error(cudaGetLastError());
}

#ifdef LCBI_TRACE
extern "C" int lcbi_debug_set_fwd_trace(long long* ptr) {
  return static_cast<int>(cudaMemcpyToSymbol(g_fwd_trace, &ptr, sizeof(ptr)));
}
#endif

}  // namespace lcbi
