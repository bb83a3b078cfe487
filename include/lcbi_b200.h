/* lcbi_b200 — C ABI of the B200-native token-mixing hot path of NHLBI/long_context_biomedical_imaging.
 *
 * The reference is pure Python: its "plugin boundary" for this path is the nn.Module seam
 * (SABlock.forward, WindowAttention.forward / SwinTransformerBlock.forward_part1, MONAI patch embedding).
 * This header is what a Python (ctypes / torch extension) binding of those seams calls; every entry point
 * cites the reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's caching allocator); the library
 *     borrows them for the duration of the enqueue and allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no internal synchronisation;
 *   - return value: LCBI_OK (0) or a negative lcbi_status; lcbi_last_error() gives the text
 *     (thread-local). Functions never throw and never fall back to a CPU path;
 *   - strides are in ELEMENTS, ordered (batch, row/token, head); the innermost head_dim is contiguous.
 */
#ifndef LCBI_B200_H_
#define LCBI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum lcbi_status {
  LCBI_OK = 0,
  LCBI_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, misaligned pointer/stride */
  LCBI_ERR_UNSUPPORTED = -2,  /* shape outside what the sm_100a kernels implement (e.g. head_dim) */
  LCBI_ERR_CUDA = -3,         /* CUDA runtime error (text in lcbi_last_error) */
  LCBI_ERR_TENSOR_MAP = -4,   /* cuTensorMapEncodeTiled rejected the view */
  LCBI_ERR_WORKSPACE = -5     /* workspace too small */
} lcbi_status;

#define LCBI_B200_VERSION 100 /* 0.1.0 */

int lcbi_version(void);

/* The dense attention kernels are persistent (one or two CTAs per SM for the whole launch). When a communication
 * kernel must run BESIDE them (ring K/V exchange over NCCL on a side stream), it needs SMs of its own: the next
 * launches ON THE CURRENT DEVICE leave `n` SMs unused (0..64, default 0; one setting per device ordinal).
 * lcbi_get_reserved_sms() returns the current device's setting so a caller can restore it afterwards. */
int lcbi_set_reserved_sms(int n);
int lcbi_get_reserved_sms(void);
const char* lcbi_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Dense (ViT global) attention core.  Replaces model/models/backbone_vit.py:191-201
 * (einsum q k^T * scale -> softmax -> einsum with v), reading q/k/v in place from the qkv Linear output
 * laid out (B, N, 3, H, d) (the "b h (qkv l d)" rearrange at backbone_vit.py:168) and writing
 * (B, N, H*d) (the "b h l d -> b l (h d)" rearrange at :169,201).
 *   q: (B,Nq,H,d) view, k/v: (B,Nk,H,d) views, o: (B,Nq,H,d) view — bf16; lse: (B,H,Nq) fp32, natural log
 *   of sum_j exp(scale*q.k_j) per row (needed by the backward and by ring-attention merging).
 *   head_dim must be 64 (every ViT preset of the reference: backbone_vit.py:56-76).
 * ---------------------------------------------------------------------------------------------- */
int lcbi_dense_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq,
                        int Nk, int head_dim, const int64_t* q_strides, const int64_t* k_strides,
                        const int64_t* v_strides, const int64_t* o_strides, float scale, void* stream);

/* Ring (sequence-parallel) step of the same forward — functionality the reference does not have ("optional
 * running (m, l, O) in/out for ring steps"). One call per visiting K/V shard; the online-softmax state of every local
 * query row is carried from call to call instead of being merged afterwards:
 *   state_o: fp32 (B,Nq,H,d) contiguous un-normalised running output; state_m / state_l: fp32 (B,H,Nq) running max
 *   (raw score units) and running row sum. first != 0: the incoming state is ignored (first shard). last != 0: the
 *   result is normalised and written to o (bf16) and lse exactly as lcbi_dense_attn_fwd does, and the state is left
 *   untouched; otherwise only the state is written back (o, lse, o_strides may be NULL). first && last == a plain call. */
int lcbi_dense_attn_fwd_state(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq,
                              int Nk, int head_dim, const int64_t* q_strides, const int64_t* k_strides,
                              const int64_t* v_strides, const int64_t* o_strides, float scale, float* state_o,
                              float* state_m, float* state_l, int first, int last, void* stream);

/* Backward of the above (what autograd derives for backbone_vit.py:191-201).
 *   d_o, o: (B,Nq,H,d) bf16 views; dq: (B,Nq,H,d), dk/dv: (B,Nk,H,d) bf16 views.
 *   accumulate_dkv != 0: dk/dv are instead fp32 (B,Nk,H,d) CONTIGUOUS buffers that are accumulated into
 *   (ring sequence-parallel steps); their stride arguments are ignored.
 *   accumulate_dq != 0: likewise dq is an fp32 (B,Nq,H,d) contiguous buffer that is accumulated into.
 *   workspace: at least lcbi_dense_attn_bwd_workspace_bytes() bytes, 128-byte aligned; with accumulate_dq the fp32 dQ
 *   accumulator inside it is not needed and lcbi_dense_attn_bwd_workspace_bytes_for(..., 1) (row-term tiles only) suffices. */
size_t lcbi_dense_attn_bwd_workspace_bytes(int B, int H, int Nq, int head_dim);
size_t lcbi_dense_attn_bwd_workspace_bytes_for(int B, int H, int Nq, int head_dim, int accumulate_dq);
int lcbi_dense_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                        const float* lse, void* dq, void* dk, void* dv, int B, int H, int Nq, int Nk, int head_dim,
                        const int64_t* q_strides, const int64_t* k_strides, const int64_t* v_strides,
                        const int64_t* o_strides, const int64_t* do_strides, const int64_t* dq_strides,
                        const int64_t* dk_strides, const int64_t* dv_strides, float scale, int accumulate_dkv,
                        int accumulate_dq, void* workspace, size_t workspace_bytes, void* stream);

/* Combine step of ring (sequence-parallel) attention — functionality the reference does not have (it cannot run
 * global attention beyond a few thousand tokens: the N x N matrix of backbone_vit.py:193 is materialised).
 * Merges a partial result (o_s bf16 (B,N,H,64) contiguous, lse_s fp32 (B,H,N)) into the running fp32 accumulator
 * (acc (B,N,H,64), lse_acc (B,H,N)); first != 0 initialises the accumulator instead. If out_bf16 is not NULL the
 * merged output is also written there as bf16 (B,N,H,64). */
int lcbi_attn_merge(float* acc, float* lse_acc, const void* o_s, const float* lse_s, void* out_bf16, int B, int N, int H,
                    int head_dim, int first, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Patch-embedding projection. Replaces MONAI 1.3.0 PatchEmbeddingBlock (proj_type="conv") as called at
 * model/models/backbone_vit.py:351-361,383 and MONAI PatchEmbed as called at backbone_swin.py:800-806,885:
 * a strided convolution with kernel == stride == patch, i.e. an implicit GEMM over non-overlapping patches.
 *   img : (B, Cin, D, H, W) contiguous, fp32 or bf16 (D = 1 for 2-D)
 *   w   : (N, Cin*Pd*Ph*Pw) fp32 — the conv weight (hidden, Cin, *patch) viewed flat;  bias: (N) fp32
 *   pos : (Gd*Gh*Gw, N) fp32 position embedding added per token, or NULL (Swin)
 *   out : (B, Gd*Gh*Gw, N) token-major fp32 or bf16 (tokens in raster order of the patch grid)
 *   img_dims/patch/grid: int[3] in (D,H,W) order. grid = floor(img/patch) reproduces PatchEmbeddingBlock,
 *   grid = ceil(img/patch) reproduces PatchEmbed's trailing zero padding.
 * The backward writes dw (N,K), dbias (N), and optionally dpos (Gd*Gh*Gw, N) and dimg (B,Cin,D,H,W), all fp32;
 * pass NULL for dpos / dimg / dbias to skip them.
 * ---------------------------------------------------------------------------------------------- */
int lcbi_patch_embed_fwd(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                         void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                         const int* grid, int N, void* stream);
int lcbi_patch_embed_bwd(const void* img, int img_is_bf16, const float* w, const void* dout, int dout_is_bf16,
                         float* dw, float* dbias, float* dpos, float* dimg, int B, int Cin, const int* img_dims,
                         const int* patch, const int* grid, int N, void* stream);
/* The same forward with a caller-owned workspace of lcbi_patch_embed_workspace_bytes(B, Cin, patch, grid, N) bytes
 * (128-byte aligned). With it, reduction lengths K = Cin*prod(patch) that are multiples of 64 (cfg1 K = 256, cfg3
 * K = 512; fp32 image, patch width a multiple of 8) run on tcgen05 / TMEM: a pre-pass writes the patches and the
 * weights as bf16 (hi, lo) pairs into the workspace, and a TMA-fed GEMM with three tensor-core products per k-step
 * keeps fp32 accuracy (max-rel ~2e-5). Other shapes, or workspace == NULL, take the kernels of lcbi_patch_embed_fwd;
 * lcbi_patch_embed_workspace_bytes returns 0 when no shape with this K uses a workspace. */
size_t lcbi_patch_embed_workspace_bytes(int B, int Cin, const int* patch, const int* grid, int N);
int lcbi_patch_embed_fwd_ws(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                            void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                            const int* grid, int N, void* workspace, size_t workspace_bytes, void* stream);
/* The backward with the same workspace: for the shapes above and fp32 dout, dw = dout^T * patches runs on tcgen05 with
 * both operands split into bf16 (hi, lo) pairs in the workspace (the M patches are the contraction; partial tiles of
 * the M-split are added into dw), and dbias rides on the split of dout. dpos / dimg as in lcbi_patch_embed_bwd. */
int lcbi_patch_embed_bwd_ws(const void* img, int img_is_bf16, const float* w, const void* dout, int dout_is_bf16,
                            float* dw, float* dbias, float* dpos, float* dimg, int B, int Cin, const int* img_dims,
                            const int* patch, const int* grid, int N, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Swin (shifted-)window attention. Replaces, as ONE gather -> attention -> scatter kernel, the body of
 * SwinTransformerBlock.forward_part1 between the qkv and proj Linear layers
 * (model/models/backbone_swin.py:441-485: F.pad, torch.roll, window_partition, window_reverse, torch.roll, crop)
 * and the attention core of WindowAttention.forward (:339-357: q*scale, q k^T, + relative_position_bias_table[
 * relative_position_index[:n,:n]], + compute_mask(...) (:591-628), softmax, @ v).
 *   ndim 2|3; grid/window/shift: int[ndim] token grid and the CONSTRUCTOR window / shift sizes (the per-axis
 *   clamp of get_window_size, :200-224, is applied inside).
 *   qkv      : bf16 (B, T, 3, H, d) — output of the qkv Linear on the UN-padded, un-shifted token grid
 *   qkv_bias : fp32 (3*H*d) — q/k/v of the zero-pad tokens (the reference pads after norm1), or NULL
 *   table    : fp32 (prod(2*window-1), H) relative_position_bias_table
 *   out      : bf16 (B, T, H*d) attention output at the source token position (input of the proj Linear)
 *   lse2     : fp32 (B, T, H) log2-domain log-sum-exp per real token (consumed by the backward)
 * head_dim must be 16 or 32 (all Swin presets: backbone_swin.py:56-94); prod(clamped window) <= 512.
 *
 * Backward: d_out bf16 (B,T,H*d), o = forward output; writes dqkv bf16 (B,T,3,H,d); ACCUMULATES (+=, caller
 * zero-initialises) dbias_pad fp32 (3*H*d) = gradient reaching qkv.bias through pad tokens, and dtable fp32
 * (same shape as table). dsum: fp32 (B,T,H) scratch.
 * ---------------------------------------------------------------------------------------------- */
int lcbi_win_attn_fwd(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                      float scale, const void* qkv, const float* qkv_bias, const float* table, void* out, float* lse2,
                      void* stream);
int lcbi_win_attn_bwd(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                      float scale, const void* qkv, const float* qkv_bias, const float* table, const void* o,
                      const float* lse2, const void* d_out, float* dsum, void* dqkv, float* dbias_pad, float* dtable,
                      void* stream);

/* Which kernels serve 3-D windows of 128..512 tokens: 0 = automatic (by measured crossover: the tcgen05 / TMEM / TMA
 * kernels of window_attn_tc.cu where they are at least as fast as the generic mma.sync kernels), 1 = tcgen05 wherever
 * applicable, 2 = generic only. Process-wide; for tests and A/B timing. */
int lcbi_set_window_kernel_mode(int mode);

/* Window-sharded variants (the windows of one block are independent, so the flattened (batch, window) list
 * can be split across GPUs): only windows [win_begin, win_begin + win_count) are processed (win_count < 0: all).
 * Forward writes `out` / `lse2` rows of the tokens inside those windows only; backward writes the matching `dqkv` rows
 * and ADDS this range's contribution to dbias_pad / dtable. Every token belongs to exactly one window, so the
 * per-range results of a partition are disjoint in out / dqkv and sum to the full-block result elsewhere. */
int lcbi_win_attn_fwd_range(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                            float scale, const void* qkv, const float* qkv_bias, const float* table, void* out,
                            float* lse2, int win_begin, int win_count, void* stream);
int lcbi_win_attn_bwd_range(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                            float scale, const void* qkv, const float* qkv_bias, const float* table, const void* o,
                            const float* lse2, const void* d_out, float* dsum, void* dqkv, float* dbias_pad,
                            float* dtable, int win_begin, int win_count, void* stream);

/* Index maps as tensors, for bit-exactness checks against window_partition(roll(pad(.))) (:135-165,:459-468),
 * compute_mask (:591-628) and relative_position_index (:256-308):
 *   gather (nW*n) int32: source token of each window slot, -1 for pad tokens; region (nW*n) int32: shift-mask
 *   region id (mask[w,i,j] = region[w,i] == region[w,j] ? 0 : -100); relidx (n*n) int32: the [:n,:n] slice.
 * Any of the three may be NULL. n_out / nw_out (host ints) receive tokens per window / windows per image. */
int lcbi_window_maps(int ndim, const int* grid, const int* window, const int* shift, int* gather, int* region,
                     int* relidx, int* n_out, int* nw_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the channel axis of token rows: the norm1 / norm2 that feed the qkv projection and the MLP in the
 * encoder blocks (torch.nn.LayerNorm(hidden), model/models/backbone_vit.py:260-263 with the modules built at :249-258;
 * backbone_swin.py:437,489-490). Statistics and arithmetic in fp32 (layer_norm is an autocast-to-fp32 op in the
 * reference); the output is written directly in the dtype the following Linear consumes.
 *   x  : (rows, C) contiguous fp32 or bf16, C % 4 == 0;  gamma, beta: (C) fp32 or NULL (no affine: the
 *        F.layer_norm of backbone_swin.py:866-879);  y: (rows, C) fp32 or bf16;  mean, rstd: (rows) fp32 (saved
 *        for the backward).
 * Backward: dy like y, dx like x (NULL to skip), dgamma / dbeta (C) fp32 WRITTEN, not accumulated (NULL to skip
 * both or either); the column reduction is deterministic (slab partials in `workspace`, summed in order).
 * ---------------------------------------------------------------------------------------------- */
int lcbi_layer_norm_fwd(const void* x, int x_is_bf16, const float* gamma, const float* beta, void* y, int y_is_bf16,
                        float* mean, float* rstd, int64_t rows, int C, float eps, void* stream);
size_t lcbi_layer_norm_bwd_workspace_bytes(int64_t rows, int C);
int lcbi_layer_norm_bwd(const void* dy, int dy_is_bf16, const void* x, int x_is_bf16, const float* gamma,
                        const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, void* workspace,
                        size_t workspace_bytes, int64_t rows, int C, void* stream);

/* The same LayerNorm fused with the residual add in front of it: the blocks compute `x = x + branch(...)` and then
 * normalise the sum (backbone_vit.py:261-262 `x = x + self.attn(self.norm1(x)); x = x + self.mlp(self.norm2(x))`, the
 * next block's norm1 or the encoder's final norm :395 following). Forward: xsum = x + delta (written, in x's dtype: the
 * new residual stream and hidden state), y = LayerNorm(xsum). Backward: dx = dxsum + dLayerNorm(dy) where dxsum (may be
 * NULL) is the gradient arriving at xsum from everything else that reads it; the same total is written as ddelta in
 * delta's dtype (NULL to skip). Rows of at least 128 channels (C % 4 == 0); other arguments as above. */
int lcbi_add_layer_norm_fwd(const void* x, int x_is_bf16, const void* delta, int delta_is_bf16, void* xsum,
                            const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean, float* rstd,
                            int64_t rows, int C, float eps, void* stream);
int lcbi_add_layer_norm_bwd(const void* dy, int dy_is_bf16, const void* dxsum, const void* xsum, int x_is_bf16,
                            const float* gamma, const float* mean, const float* rstd, void* dx, void* ddelta,
                            int delta_is_bf16, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                            int64_t rows, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row gather / scatter of the window-sharded Swin exchange (no reference counterpart: the reference has no window
 * parallelism). Rows are row_bytes (a multiple of 16) contiguous bytes, ids are int64 on the device.
 *   lcbi_gather_rows : dst[i, :] = src[ids[i], :]      i < n_rows   (pack the rank's window rows before the all-gather)
 *   lcbi_scatter_rows: dst[ids[i], :] = src[i, :]      i < n_rows   (put every rank's rows back in token order)
 * ---------------------------------------------------------------------------------------------- */
int lcbi_gather_rows(const void* src, const int64_t* ids, void* dst, int64_t n_rows, int row_bytes, void* stream);
int lcbi_scatter_rows(const void* src, const int64_t* ids, void* dst, int64_t n_rows, int row_bytes, void* stream);

/* Bias gradient of the token-wise Linear layers that bracket the attention kernels (qkv with bias at
 * backbone_swin.py:309, out_proj / proj at backbone_vit.py:167,202 and backbone_swin.py:311,358, the MLP's
 * linear1 / linear2): dbias (C) fp32 = column sums of dy (rows, C) fp32 or bf16, C % 4 == 0, WRITTEN not accumulated.
 * Deterministic (slab partials summed in order); `workspace` as for lcbi_layer_norm_bwd
 * (lcbi_layer_norm_bwd_workspace_bytes(rows, C)). The GEMMs of those layers stay cuBLAS. */
int lcbi_bias_grad(const void* dy, int dy_is_bf16, float* dbias, void* workspace, size_t workspace_bytes, int64_t rows,
                   int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCBI_B200_H_ */
