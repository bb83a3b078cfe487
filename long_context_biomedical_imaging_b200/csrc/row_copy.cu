// Row gather / scatter for the window-sharded Swin exchange (window_parallel.py; SURVEY 8e "Swin windows", option (a)).
//
// Each rank owns the rows of its windows' tokens. Before the all-gather they are packed window-major
// (gather: dst[i] = src[ids[i]]), afterwards the ranks' rows are put back in token order (scatter: dst[ids[i]] = src[i]).
// torch's index_select / index_copy_ did this at ~0.2 TB/s (per-element index kernels: 250 us per 75 MB dqkv at cfg4
// stage 1, a quarter of the sharded step); rows are 96 .. 576 contiguous bytes, so one 16-byte vector per thread with
// the row id read once per vector is a plain streaming copy.
#include "lcbi_kernels.h"

namespace lcbi {

namespace {

template <bool kScatter>
__global__ void __launch_bounds__(256)
row_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int64_t* __restrict__ ids, int64_t n_rows,
                int row_vecs) {
  const int64_t total = n_rows * row_vecs;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / row_vecs;
    const int v = static_cast<int>(idx - row * row_vecs);
    const int64_t other = __ldg(ids + row);
    if (kScatter) dst[other * row_vecs + v] = src[idx];
    else dst[idx] = __ldg(src + other * row_vecs + v);
  }
}

}  // namespace

int row_copy_launch(const void* src, const int64_t* ids, void* dst, int64_t n_rows, int row_bytes, int scatter,
                    cudaStream_t stream) {
  if (n_rows < 0 || row_bytes <= 0 || row_bytes % 16 != 0) return LCBI_ERR_BAD_ARG;
  if (n_rows == 0) return LCBI_OK;
  if ((reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0) return LCBI_ERR_BAD_ARG;
  const int row_vecs = row_bytes / 16;
  const int64_t total = n_rows * row_vecs;
  const int sms = current_device_sm_count();
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sms > 0 ? sms : 148) * 16;
  if (blocks > cap) blocks = cap;
  if (scatter)
    row_copy_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint4*>(src),
                                                                             static_cast<uint4*>(dst), ids, n_rows, row_vecs);
  else
    row_copy_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint4*>(src),
                                                                              static_cast<uint4*>(dst), ids, n_rows, row_vecs);
  return set_cuda_error(cudaGetLastError());
}

}  // namespace lcbi
