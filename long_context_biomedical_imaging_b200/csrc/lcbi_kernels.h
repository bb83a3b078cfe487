// Internal launch interface shared by the kernel translation units and capi.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/lcbi_b200.h"

namespace lcbi {

// records the CUDA error text for lcbi_last_error(); returns LCBI_OK or LCBI_ERR_CUDA
int set_cuda_error(cudaError_t e);
void set_error_text(const char* msg);

struct DenseAttnArgs {
  const void* q; const void* k; const void* v;   // bf16, (B, N, H, d) strided views, d contiguous
  void* o;                                       // bf16, (B, Nq, H, d) strided view
  float* lse;                                    // fp32 (B, H, Nq), natural log of the row sums
  int B, H, Nq, Nk, head_dim;
  int64_t q_strides[3], k_strides[3], v_strides[3], o_strides[3];  // (batch, row, head) in elements
  float scale;
  // carried online-softmax state (ring steps); all nullptr / 0 for a plain call
  float* state_o = nullptr;   // fp32 (B, Nq, H, d) contiguous, un-normalised running output
  float* state_m = nullptr;   // fp32 (B, H, Nq) running max (raw score units)
  float* state_l = nullptr;   // fp32 (B, H, Nq) running sum
  int state_first = 0;        // 1: the incoming state is ignored (first K/V shard)
  int state_last = 0;         // 1: normalise and write o (bf16) + lse; otherwise only the state is written back
};
int dense_attn_fwd_launch(const DenseAttnArgs& a, cudaStream_t stream);

struct DenseAttnBwdArgs {
  const void* q; const void* k; const void* v; const void* o; const void* d_o;  // bf16
  const float* lse;                                                             // (B, H, Nq)
  void* dq; void* dk; void* dv;                                                 // bf16 outputs
  void* workspace; size_t workspace_bytes;
  int B, H, Nq, Nk, head_dim;
  int64_t q_strides[3], k_strides[3], v_strides[3], o_strides[3], do_strides[3];
  int64_t dq_strides[3], dk_strides[3], dv_strides[3];
  float scale;
  int accumulate_dkv;   // 0: write bf16 dk/dv; 1: dk/dv are fp32 (B,Nk,H,d) contiguous accumulators (+=)
  int accumulate_dq;    // 0: write bf16 dq;    1: dq is an fp32 (B,Nq,H,d) contiguous accumulator (+=)
};
size_t dense_attn_bwd_workspace_bytes(int B, int H, int Nq, int head_dim, int accumulate_dq);
int dense_attn_bwd_launch(const DenseAttnBwdArgs& a, cudaStream_t stream);

// SMs the persistent dense kernels leave free (for communication kernels running beside them); capi.cu
int reserved_sms();

// Per-device launch state (capi.cu). cudaFuncSetAttribute applies to the CURRENT device only, so "already configured"
// is remembered per device ordinal, not per process: one bit per device in a word owned by the kernel's launcher.
// Returns true when the current device's bit was not set yet (and sets it).
bool first_launch_on_current_device(unsigned long long* seen_mask);
// SM count of the current device (cached per ordinal); <= 0 on error (text recorded for lcbi_last_error()).
int current_device_sm_count();

// patch embedding (patch_embed.cu). img_dims / patch / grid are (D, H, W)-ordered triples (D = 1 for 2-D).
int patch_embed_fwd_launch(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                           void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                           const int* grid, int N, cudaStream_t stream);
// workspace (optional): with it, shapes patch_embed_tc_applicable accepts take the tcgen05 weight-gradient path
int patch_embed_bwd_launch(const void* img, int img_is_bf16, const float* w, const void* dout, int dout_is_bf16,
                           float* dw, float* dbias, float* dpos, float* dimg, int B, int Cin, const int* img_dims,
                           const int* patch, const int* grid, int N, cudaStream_t stream, void* workspace = nullptr,
                           size_t workspace_bytes = 0);

// tensor-core variants for K >= 64 at fp32 accuracy (patch_embed_mma.cu); `applicable` is a host-side shape test
bool patch_embed_mma_applicable(int Cin, const int* img_dims, const int* patch, const int* grid, int N);
int patch_embed_fwd_mma_launch(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                               void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                               const int* grid, int N, cudaStream_t stream);

// tcgen05 / TMEM / TMA implicit GEMM for K % 64 == 0, fp32 image (patch_embed_tc.cu); needs a workspace for the split operands
bool patch_embed_tc_applicable(int img_is_bf16, int Cin, const int* img_dims, const int* patch, const int* grid, int N);
size_t patch_embed_tc_workspace_bytes(int64_t M, int N, int K);
int patch_embed_tc_fwd_launch(const void* img, const float* w, const float* bias, const float* pos, void* out, int out_is_bf16,
                              int B, int Cin, const int* img_dims, const int* patch, const int* grid, int N, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream);
int patch_embed_tc_bwd_w_launch(const void* img, const float* dout, float* dw, float* dbias, int B, int Cin, const int* img_dims,
                                const int* patch, const int* grid, int N, void* workspace, size_t workspace_bytes,
                                cudaStream_t stream);

int patch_embed_bwd_w_mma_launch(const void* img, int img_is_bf16, const void* dout, int dout_is_bf16, float* dw,
                                 float* dbias, int B, int Cin, const int* img_dims, const int* patch, const int* grid,
                                 int N, cudaStream_t stream);

// Swin window attention (window_attn.cu)
struct WinAttnArgs {
  int ndim;                       // 2 or 3
  const int* grid;                // token grid, ndim entries
  const int* window;              // constructor window size
  const int* shift;               // constructor shift size (all zeros for W-MSA blocks)
  int B, H, head_dim;
  int win_begin, win_count;       // range of the flattened (batch, window) list; count < 0: all windows
  float scale;
  const void* qkv;                // bf16 (B, T, 3, H, d)
  const float* qkv_bias;          // fp32 (3*H*d) or null
  const float* table;             // fp32 (prod(2*window-1), H)
  void* out;                      // bf16 (B, T, H*d)
  float* lse2;                    // fp32 (B, T, H)
  const void* d_out;              // bf16 (B, T, H*d)            [bwd]
  float* dsum;                    // fp32 (B, T, H) scratch      [bwd]
  void* dqkv;                     // bf16 (B, T, 3, H, d)        [bwd]
  float* dbias_pad;               // fp32 (3*H*d), +=            [bwd]
  float* dtable;                  // fp32 like table, +=         [bwd]
};
int win_attn_fwd_launch(const WinAttnArgs& a, cudaStream_t stream);
int win_attn_bwd_launch(const WinAttnArgs& a, const void* o, cudaStream_t stream);
int window_maps_launch(int ndim, const int* grid, const int* window, const int* shift, int* gather, int* region,
                       int* relidx, int* n_out, int* nw_out, cudaStream_t stream);

// ring-attention combine step (attn_merge.cu)
int attn_merge_launch(float* acc, float* lse_acc, const void* o_s, const float* lse_s, void* out_bf16, int B, int N,
                      int H, int head_dim, int first, cudaStream_t stream);

// LayerNorm over the channel axis (layer_norm.cu)
int layer_norm_fwd_launch(const void* x, int x_is_bf16, const float* gamma, const float* beta, void* y, int y_is_bf16,
                          float* mean, float* rstd, int64_t rows, int C, float eps, cudaStream_t stream);
size_t layer_norm_bwd_workspace_bytes(int64_t rows, int C);
int layer_norm_bwd_launch(const void* dy, int dy_is_bf16, const void* x, int x_is_bf16, const float* gamma,
                          const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                          float* workspace, size_t workspace_bytes, int64_t rows, int C, cudaStream_t stream);

int add_layer_norm_fwd_launch(const void* x, int x_is_bf16, const void* delta, int delta_is_bf16, void* xsum,
                              const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean, float* rstd,
                              int64_t rows, int C, float eps, cudaStream_t stream);
int add_layer_norm_bwd_launch(const void* dy, int y_is_bf16, const void* dres, const void* xsum, int x_is_bf16,
                              const float* gamma, const float* mean, const float* rstd, void* dx, void* ddelta,
                              int delta_is_bf16, float* dgamma, float* dbeta, float* workspace, size_t workspace_bytes,
                              int64_t rows, int C, cudaStream_t stream);
int bias_grad_launch(const void* dy, int dy_is_bf16, float* dbias, float* workspace, size_t workspace_bytes,
                     int64_t rows, int C, cudaStream_t stream);

// window-sharded exchange: dst[i] = src[ids[i]] (scatter == 0) or dst[ids[i]] = src[i] (scatter != 0); rows of row_bytes
// (a multiple of 16) contiguous bytes (row_copy.cu)
int row_copy_launch(const void* src, const int64_t* ids, void* dst, int64_t n_rows, int row_bytes, int scatter,
                    cudaStream_t stream);

}  // namespace lcbi
