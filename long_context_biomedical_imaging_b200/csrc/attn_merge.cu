// Log-sum-exp merge of partial attention results — the combine step of ring (sequence-parallel) attention.
//
// New functionality relative to the reference (which has no sequence parallelism: SURVEY §5, §8e): each ring step
// produces, for the local query shard, a normalised partial output o_s (bf16) and its log-sum-exp lse_s over the
// visiting K/V shard. The running result is kept in fp32:
//     lse' = log(exp(lse) + exp(lse_s));   acc' = exp(lse - lse') * acc + exp(lse_s - lse') * o_s
// One warp per (batch, row, head): 32 lanes x 2 elements = head_dim 64. HBM-bound elementwise kernel.
#include <cuda_bf16.h>

#include "lcbi_kernels.h"

namespace lcbi {

namespace {

__global__ void attn_merge_kernel(float* __restrict__ acc, float* __restrict__ lse_acc,
                                  const __nv_bfloat16* __restrict__ o_s, const float* __restrict__ lse_s,
                                  __nv_bfloat16* __restrict__ out_bf16, int B, int N, int H, int first) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;   // (b, n, h) flattened
  const int64_t total = static_cast<int64_t>(B) * N * H;
  if (row >= total) return;
  const int h = static_cast<int>(row % H);
  const int64_t bn = row / H;
  const int n = static_cast<int>(bn % N);
  const int b = static_cast<int>(bn / N);
  const int64_t lse_idx = (static_cast<int64_t>(b) * H + h) * N + n;                        // lse is (B, H, N)
  const float ls = lse_s[lse_idx];
  // lane 0 alone reads the running lse and broadcasts it: the same lane overwrites that address below, so no other
  // lane may still have a load of it in flight (no reliance on the warp staying converged)
  float la = 0.f;
  if (!first && lane == 0) la = lse_acc[lse_idx];
  la = __shfl_sync(0xffffffffu, la, 0);
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(o_s + row * 64 + lane * 2);
  float2 x = __bfloat1622float2(v);
  float2 r;
  float lnew;
  if (first) {
    r = x;
    lnew = ls;
  } else {
    const float m = fmaxf(la, ls);
    const float wa = __expf(la - m), ws = __expf(ls - m);
    const float inv = 1.f / (wa + ws);
    lnew = m + __logf(wa + ws);
    const float2 a = *reinterpret_cast<const float2*>(acc + row * 64 + lane * 2);
    r.x = (a.x * wa + x.x * ws) * inv;
    r.y = (a.y * wa + x.y * ws) * inv;
  }
  *reinterpret_cast<float2*>(acc + row * 64 + lane * 2) = r;
  if (out_bf16 != nullptr)
    *reinterpret_cast<__nv_bfloat162*>(out_bf16 + row * 64 + lane * 2) = __floats2bfloat162_rn(r.x, r.y);
  if (lane == 0) lse_acc[lse_idx] = lnew;
}

}  // namespace

int attn_merge_launch(float* acc, float* lse_acc, const void* o_s, const float* lse_s, void* out_bf16, int B, int N,
                      int H, int head_dim, int first, cudaStream_t stream) {
  if (head_dim != 64) return LCBI_ERR_UNSUPPORTED;
  if (B <= 0 || N <= 0 || H <= 0) return LCBI_ERR_BAD_ARG;
  const int64_t rows = static_cast<int64_t>(B) * N * H;
  const int threads = 256;
  const int64_t blocks = (rows * 32 + threads - 1) / threads;
  attn_merge_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      acc, lse_acc, static_cast<const __nv_bfloat16*>(o_s), lse_s, static_cast<__nv_bfloat16*>(out_bf16), B, N, H, first);
  return set_cuda_error(cudaGetLastError());
}

}  // namespace lcbi
