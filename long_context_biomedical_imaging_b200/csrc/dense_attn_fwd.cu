// Dense (ViT global) multi-head self-attention forward for sm_100a, head_dim 64, bf16 in / fp32 accumulate.
//
// Replaces the attention core of the reference's SABlock.forward
// (/root/reference/model/models/backbone_vit.py:191-201): softmax(scale * q k^T) v, without ever
// materialising the B*H*N*N matrix and reading q/k/v straight out of the (B,N,3,H,d) qkv tensor.
//
// One CTA = two 128-row query tiles of one (batch, head) sharing every K/V tile:
//   warp 9      TMA producer   (Q0,Q1 once; K_j / V_j through a 3-stage ring)
//   warp 8      UMMA issuer    (S_t = Q_t K_j^T  -> TMEM;  O_t += P_t V_j with P_t read from TMEM)
//   warps 0-3   softmax for query tile 0 (one thread = one query row = one TMEM lane)
//   warps 4-7   softmax for query tile 1
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384); P_t (bf16) aliases S_t[0,64).
// The two softmax groups ping-pong against the tensor pipe: while group 0 exponentiates S0^j the
// tensor core computes S1^j / P1 V.  Online softmax uses a lazily updated running max (only moved when
// the new max exceeds it by 2^8) so the O rescale in TMEM is rare.
#include "lcbi_kernels.h"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace lcbi {

namespace {

constexpr int kBlockM = 128;      // rows per query tile
constexpr int kBlockN = 128;      // keys per KV tile
constexpr int kHeadDim = 64;
constexpr int kStages = 3;        // K and V ring depth
constexpr int kTileBytes = kBlockM * kHeadDim * 2;  // 16 KB
constexpr int kNumThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

constexpr uint32_t kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 320;

struct __align__(1024) FwdSmem {
  uint8_t q[2][kTileBytes];        // also the O staging tiles for the TMA store
  uint8_t k[kStages][kTileBytes];
  uint8_t v[kStages][kTileBytes];
  uint64_t q_full[2];
  uint64_t k_full[kStages], k_empty[kStages];
  uint64_t v_full[kStages], v_empty[kStages];
  uint64_t s_full[2], p_full[2], o_full[2];
  uint32_t tmem_base;
};

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
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q_base = blockIdx.x * (2 * kBlockM);
  const bool tile1_active = (q_base + kBlockM) < p.Nq;
  const int n_kv = (p.Nk + kBlockN - 1) / kBlockN;

  if (tid == 0) {
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sm.q_full[t], 1);
      mbar_init(&sm.s_full[t], 1);
      mbar_init(&sm.p_full[t], 128);
      mbar_init(&sm.o_full[t], 1);
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
      mbar_expect_tx(&sm.q_full[0], kTileBytes);
      tma_load_4d(sm.q[0], &tm_q, &sm.q_full[0], 0, head, q_base, batch);
      if (tile1_active) {
        mbar_expect_tx(&sm.q_full[1], kTileBytes);
        tma_load_4d(sm.q[1], &tm_q, &sm.q_full[1], 0, head, q_base + kBlockM, batch);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        mbar_wait(&sm.k_empty[s], ph ^ 1);
        mbar_expect_tx(&sm.k_full[s], kTileBytes);
        tma_load_4d(sm.k[s], &tm_k, &sm.k_full[s], 0, head, j * kBlockN, batch);
        mbar_wait(&sm.v_empty[s], ph ^ 1);
        mbar_expect_tx(&sm.v_full[s], kTileBytes);
        tma_load_4d(sm.v[s], &tm_v, &sm.v_full[s], 0, head, j * kBlockN, batch);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBlockM, kBlockN, 0, 0);   // Q K^T : both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(kBlockM, kHeadDim, 0, 1);  // P V   : V is MN-major
      const uint32_t q_addr[2] = {smem_u32(sm.q[0]), smem_u32(sm.q[1])};

      auto issue_s = [&](int t, int stage) {
        const uint32_t k_addr = smem_u32(sm.k[stage]);
#pragma unroll
        for (int kk = 0; kk < kHeadDim / 16; ++kk) {
          const uint64_t da = make_smem_desc(q_addr[t] + kk * 32, 16, 1024, kLayoutSW128);
          const uint64_t db = make_smem_desc(k_addr + kk * 32, 16, 1024, kLayoutSW128);
          umma_ss(tmem + (t ? kTmemS1 : kTmemS0), da, db, idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&sm.s_full[t]);
      };
      auto issue_pv = [&](int t, int stage, bool accumulate) {
        const uint32_t v_addr = smem_u32(sm.v[stage]);
#pragma unroll
        for (int kk = 0; kk < kBlockN / 16; ++kk) {
          const uint64_t db = make_smem_desc(v_addr + kk * 2048, 16, 1024, kLayoutSW128);
          umma_ts(tmem + (t ? kTmemO1 : kTmemO0), tmem + (t ? kTmemS1 : kTmemS0) + kk * 8, db, idesc_o,
                  (accumulate || kk > 0) ? 1u : 0u);
        }
      };

      mbar_wait(&sm.q_full[0], 0);
      mbar_wait(&sm.k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      if (tile1_active) {
        mbar_wait(&sm.q_full[1], 0);
        tc_fence_after();
        issue_s(1, 0);
      }
      umma_commit(&sm.k_empty[0]);

      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % kStages;
        const uint32_t vph = (j / kStages) & 1;
        const int ks = (j + 1) % kStages;
        const uint32_t kph = ((j + 1) / kStages) & 1;
        const bool more = (j + 1) < n_kv;

        mbar_wait(&sm.v_full[vs], vph);
        mbar_wait(&sm.p_full[0], j & 1);
        tc_fence_after();
        issue_pv(0, vs, j > 0);
        if (more) {
          mbar_wait(&sm.k_full[ks], kph);
          tc_fence_after();
          issue_s(0, ks);
        } else {
          umma_commit(&sm.o_full[0]);
        }
        if (tile1_active) {
          mbar_wait(&sm.p_full[1], j & 1);
          tc_fence_after();
          issue_pv(1, vs, j > 0);
        }
        umma_commit(&sm.v_empty[vs]);
        if (more) {
          if (tile1_active) issue_s(1, ks);
          umma_commit(&sm.k_empty[ks]);
        } else if (tile1_active) {
          umma_commit(&sm.o_full[1]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int t = warp >> 2;                   // query tile handled by this warpgroup
    const int row = tid & 127;                 // row inside the tile == TMEM lane
    if (t == 0 || tile1_active) {
      const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t t_s = tmem + lane_sel + (t ? kTmemS1 : kTmemS0);
      const uint32_t t_o = tmem + lane_sel + (t ? kTmemO1 : kTmemO0);
      const float c = p.scale_log2;
      float m_used = -INFINITY;  // running max actually subtracted (raw score units)
      float l = 0.f;

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&sm.s_full[t], j & 1);
        tc_fence_after();
        uint32_t sr[128];
        tmem_ld_x32(t_s + 0, sr + 0);
        tmem_ld_x32(t_s + 32, sr + 32);
        tmem_ld_x32(t_s + 64, sr + 64);
        tmem_ld_x32(t_s + 96, sr + 96);
        tmem_ld_wait();

        const int valid = p.Nk - j * kBlockN;  // >= 1
        if (valid < kBlockN) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) sr[i] = __float_as_uint(-INFINITY);
        }
        float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]),
              mx3 = __uint_as_float(sr[3]);
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
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
            uint32_t o[64];
            tmem_ld_x32(t_o, o);
            tmem_ld_x32(t_o + 32, o + 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(t_o, o);
            tmem_st_x32(t_o + 32, o + 32);
          }
        }

        const float mc = m_used * c;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        uint32_t pk[64];
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          const float e0 = fast_exp2(fmaf(__uint_as_float(sr[i]), c, -mc));
          const float e1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), c, -mc));
          const float e2 = fast_exp2(fmaf(__uint_as_float(sr[i + 2]), c, -mc));
          const float e3 = fast_exp2(fmaf(__uint_as_float(sr[i + 3]), c, -mc));
          s0 += e0; s1 += e1; s2 += e2; s3 += e3;
          pk[i / 2] = pack_bf16x2(e0, e1);
          pk[i / 2 + 1] = pack_bf16x2(e2, e3);
        }
        l += (s0 + s1) + (s2 + s3);
        tmem_st_x32(t_s, pk);
        tmem_st_x32(t_s + 32, pk + 32);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&sm.p_full[t]);
      }

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
                         int64_t row_stride, int64_t head_stride) {
  const uint64_t dims[4] = {static_cast<uint64_t>(kHeadDim), static_cast<uint64_t>(H), static_cast<uint64_t>(N),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(head_stride) * 2, static_cast<uint64_t>(row_stride) * 2,
                               static_cast<uint64_t>(batch_stride) * 2};
  const uint32_t box[4] = {kHeadDim, 1, kBlockM, 1};
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
  if (make_bhnd_map(&tq, a.q, a.B, a.H, a.Nq, a.q_strides[0], a.q_strides[1], a.q_strides[2]) ||
      make_bhnd_map(&tk, a.k, a.B, a.H, a.Nk, a.k_strides[0], a.k_strides[1], a.k_strides[2]) ||
      make_bhnd_map(&tv, a.v, a.B, a.H, a.Nk, a.v_strides[0], a.v_strides[1], a.v_strides[2]) ||
      make_bhnd_map(&to, a.o, a.B, a.H, a.Nq, a.o_strides[0], a.o_strides[1], a.o_strides[2]))
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
  dim3 grid((a.Nq + 2 * kBlockM - 1) / (2 * kBlockM), a.H, a.B);
  dense_attn_fwd_kernel<<<grid, kNumThreads, smem_bytes, stream>>>(tq, tk, tv, to, p);
  return set_cuda_error(cudaGetLastError());
}

}  // namespace lcbi
