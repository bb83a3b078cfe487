// extern "C" entry points declared in include/lcbi_b200.h: argument checking + dispatch to the launchers.
#include <cstdio>
#include <cstring>

#include "lcbi_kernels.h"
#include "tma_host.h"

namespace lcbi {

static thread_local char g_err[512] = "";

void set_error_text(const char* msg) {
  std::strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int set_cuda_error(cudaError_t e) {
  if (e == cudaSuccess) return LCBI_OK;
  std::snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", static_cast<int>(e), cudaGetErrorString(e));
  return LCBI_ERR_CUDA;
}

static int fail(int code, const char* msg) {
  if (code == LCBI_ERR_TENSOR_MAP) {       // name the rejected view
    std::snprintf(g_err, sizeof(g_err), "%s [%s]", msg, tmap_error_text());
    return code;
  }
  set_error_text(msg);
  return code;
}

static void copy3(int64_t* dst, const int64_t* src) {
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = src[2];
}

}  // namespace lcbi

using namespace lcbi;

// SMs the persistent dense kernels leave free, per device ordinal (a process may drive several devices)
static int g_reserved_sms[64] = {0};
static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}
static int g_window_kernel_mode = 0;
namespace lcbi {
int window_kernel_mode() { return __atomic_load_n(&g_window_kernel_mode, __ATOMIC_RELAXED); }
int reserved_sms() { return __atomic_load_n(&g_reserved_sms[current_device_slot()], __ATOMIC_RELAXED); }

bool first_launch_on_current_device(unsigned long long* seen_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;   // unknown: configure every time
  const unsigned long long bit = 1ull << dev;
  const unsigned long long old = __atomic_fetch_or(seen_mask, bit, __ATOMIC_RELAXED);
  return (old & bit) == 0;
}

int current_device_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return 0;
  }
  if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return 0;
  }
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}
}  // namespace lcbi

extern "C" {

int lcbi_version(void) { return LCBI_B200_VERSION; }

int lcbi_set_reserved_sms(int n) {
  if (n < 0 || n > 64) return fail(LCBI_ERR_BAD_ARG, "lcbi_set_reserved_sms: expected 0..64");
  __atomic_store_n(&g_reserved_sms[current_device_slot()], n, __ATOMIC_RELAXED);
  return LCBI_OK;
}

int lcbi_get_reserved_sms(void) { return lcbi::reserved_sms(); }

const char* lcbi_last_error(void) { return g_err; }

int lcbi_set_window_kernel_mode(int mode) {
  if (mode < 0 || mode > 2) return fail(LCBI_ERR_BAD_ARG, "lcbi_set_window_kernel_mode: expected 0 (auto), 1 (tcgen05) or 2 (generic)");
  __atomic_store_n(&g_window_kernel_mode, mode, __ATOMIC_RELAXED);
  return LCBI_OK;
}

int lcbi_dense_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq,
                        int Nk, int head_dim, const int64_t* q_strides, const int64_t* k_strides,
                        const int64_t* v_strides, const int64_t* o_strides, float scale, void* stream) {
  if (!q || !k || !v || !o || !lse || !q_strides || !k_strides || !v_strides || !o_strides)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_dense_attn_fwd: null pointer argument");
  DenseAttnArgs a;
  a.q = q; a.k = k; a.v = v; a.o = o; a.lse = lse;
  a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.head_dim = head_dim;
  copy3(a.q_strides, q_strides); copy3(a.k_strides, k_strides);
  copy3(a.v_strides, v_strides); copy3(a.o_strides, o_strides);
  a.scale = scale;
  int rc = dense_attn_fwd_launch(a, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_UNSUPPORTED) return fail(rc, "lcbi_dense_attn_fwd: only head_dim == 64 is implemented");
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_dense_attn_fwd: bad size, or pointer/stride not 16-byte aligned");
  if (rc == LCBI_ERR_TENSOR_MAP) return fail(rc, "lcbi_dense_attn_fwd: TMA tensor map encode failed");
  return rc;
}

int lcbi_dense_attn_fwd_state(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int Nq,
                              int Nk, int head_dim, const int64_t* q_strides, const int64_t* k_strides,
                              const int64_t* v_strides, const int64_t* o_strides, float scale, float* state_o,
                              float* state_m, float* state_l, int first, int last, void* stream) {
  if (!q || !k || !v || !q_strides || !k_strides || !v_strides || !state_o || !state_m || !state_l)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_dense_attn_fwd_state: null pointer argument");
  if (last && (!o || !lse || !o_strides))
    return fail(LCBI_ERR_BAD_ARG, "lcbi_dense_attn_fwd_state: the last step needs o, lse and o_strides");
  DenseAttnArgs a;
  a.q = q; a.k = k; a.v = v; a.lse = lse;
  a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.head_dim = head_dim;
  copy3(a.q_strides, q_strides); copy3(a.k_strides, k_strides); copy3(a.v_strides, v_strides);
  if (last) {
    a.o = o;
    copy3(a.o_strides, o_strides);
  } else {                       // nothing is stored through the output map on a non-final step: alias it to q
    a.o = const_cast<void*>(q);
    copy3(a.o_strides, q_strides);
  }
  a.scale = scale;
  a.state_o = state_o; a.state_m = state_m; a.state_l = state_l;
  a.state_first = first ? 1 : 0;
  a.state_last = last ? 1 : 0;
  int rc = dense_attn_fwd_launch(a, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_UNSUPPORTED) return fail(rc, "lcbi_dense_attn_fwd_state: only head_dim == 64 is implemented");
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_dense_attn_fwd_state: bad size, or pointer/stride not 16-byte aligned");
  if (rc == LCBI_ERR_TENSOR_MAP) return fail(rc, "lcbi_dense_attn_fwd_state: TMA tensor map encode failed");
  return rc;
}

size_t lcbi_dense_attn_bwd_workspace_bytes(int B, int H, int Nq, int head_dim) {
  return dense_attn_bwd_workspace_bytes(B, H, Nq, head_dim, 0);
}

size_t lcbi_dense_attn_bwd_workspace_bytes_for(int B, int H, int Nq, int head_dim, int accumulate_dq) {
  return dense_attn_bwd_workspace_bytes(B, H, Nq, head_dim, accumulate_dq);
}

int lcbi_dense_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                        const float* lse, void* dq, void* dk, void* dv, int B, int H, int Nq, int Nk, int head_dim,
                        const int64_t* q_strides, const int64_t* k_strides, const int64_t* v_strides,
                        const int64_t* o_strides, const int64_t* do_strides, const int64_t* dq_strides,
                        const int64_t* dk_strides, const int64_t* dv_strides, float scale, int accumulate_dkv,
                        int accumulate_dq, void* workspace, size_t workspace_bytes, void* stream) {
  if (!q || !k || !v || !o || !d_o || !lse || !dq || !dk || !dv || !workspace || !q_strides || !k_strides ||
      !v_strides || !o_strides || !do_strides || (!accumulate_dq && !dq_strides))
    return fail(LCBI_ERR_BAD_ARG, "lcbi_dense_attn_bwd: null pointer argument");
  if (!accumulate_dkv && (!dk_strides || !dv_strides))
    return fail(LCBI_ERR_BAD_ARG, "lcbi_dense_attn_bwd: dk/dv strides required unless accumulate_dkv");
  DenseAttnBwdArgs a;
  a.q = q; a.k = k; a.v = v; a.o = o; a.d_o = d_o; a.lse = lse;
  a.dq = dq; a.dk = dk; a.dv = dv;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.head_dim = head_dim;
  copy3(a.q_strides, q_strides); copy3(a.k_strides, k_strides); copy3(a.v_strides, v_strides);
  copy3(a.o_strides, o_strides); copy3(a.do_strides, do_strides);
  const int64_t contiguous_q[3] = {static_cast<int64_t>(Nq) * H * head_dim, static_cast<int64_t>(H) * head_dim, head_dim};
  copy3(a.dq_strides, accumulate_dq ? contiguous_q : dq_strides);
  const int64_t contiguous[3] = {static_cast<int64_t>(Nk) * H * head_dim, static_cast<int64_t>(H) * head_dim, head_dim};
  copy3(a.dk_strides, accumulate_dkv ? contiguous : dk_strides);
  copy3(a.dv_strides, accumulate_dkv ? contiguous : dv_strides);
  a.scale = scale;
  a.accumulate_dkv = accumulate_dkv;
  a.accumulate_dq = accumulate_dq;
  int rc = dense_attn_bwd_launch(a, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_UNSUPPORTED) return fail(rc, "lcbi_dense_attn_bwd: only head_dim == 64 is implemented");
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_dense_attn_bwd: bad size, or pointer/stride not 16-byte aligned");
  if (rc == LCBI_ERR_TENSOR_MAP) return fail(rc, "lcbi_dense_attn_bwd: TMA tensor map encode failed");
  if (rc == LCBI_ERR_WORKSPACE) return fail(rc, "lcbi_dense_attn_bwd: workspace too small or misaligned");
  return rc;
}

int lcbi_patch_embed_fwd(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                         void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                         const int* grid, int N, void* stream) {
  if (!img || !w || !bias || !out || !img_dims || !patch || !grid)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_patch_embed_fwd: null pointer argument");
  int rc = patch_embed_fwd_launch(img, img_is_bf16, w, bias, pos, out, out_is_bf16, B, Cin, img_dims, patch, grid, N,
                                  static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_patch_embed_fwd: non-positive size");
  return rc;
}

size_t lcbi_patch_embed_workspace_bytes(int B, int Cin, const int* patch, const int* grid, int N) {
  if (!patch || !grid || B <= 0 || Cin <= 0 || N <= 0) return 0;
  const int K = Cin * patch[0] * patch[1] * patch[2];
  if (K < 64 || K % 64 != 0) return 0;          // no shape with this K takes the workspace path
  return patch_embed_tc_workspace_bytes(static_cast<int64_t>(B) * grid[0] * grid[1] * grid[2], N, K);
}

int lcbi_patch_embed_fwd_ws(const void* img, int img_is_bf16, const float* w, const float* bias, const float* pos,
                            void* out, int out_is_bf16, int B, int Cin, const int* img_dims, const int* patch,
                            const int* grid, int N, void* workspace, size_t workspace_bytes, void* stream) {
  if (!img || !w || !bias || !out || !img_dims || !patch || !grid)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_patch_embed_fwd_ws: null pointer argument");
  if (B <= 0 || Cin <= 0 || N <= 0) return fail(LCBI_ERR_BAD_ARG, "lcbi_patch_embed_fwd_ws: non-positive size");
  if (workspace != nullptr && patch_embed_tc_applicable(img_is_bf16, Cin, img_dims, patch, grid, N)) {
    int rc = patch_embed_tc_fwd_launch(img, w, bias, pos, out, out_is_bf16, B, Cin, img_dims, patch, grid, N, workspace,
                                       workspace_bytes, static_cast<cudaStream_t>(stream));
    if (rc == LCBI_ERR_WORKSPACE) return fail(rc, "lcbi_patch_embed_fwd_ws: workspace too small or misaligned");
    if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_patch_embed_fwd_ws: pointers must be 16-byte aligned");
    if (rc == LCBI_ERR_TENSOR_MAP) return fail(rc, "lcbi_patch_embed_fwd_ws: TMA tensor map encode failed");
    return rc;
  }
  return lcbi_patch_embed_fwd(img, img_is_bf16, w, bias, pos, out, out_is_bf16, B, Cin, img_dims, patch, grid, N, stream);
}

int lcbi_patch_embed_bwd(const void* img, int img_is_bf16, const float* w, const void* dout, int dout_is_bf16,
                         float* dw, float* dbias, float* dpos, float* dimg, int B, int Cin, const int* img_dims,
                         const int* patch, const int* grid, int N, void* stream) {
  return lcbi_patch_embed_bwd_ws(img, img_is_bf16, w, dout, dout_is_bf16, dw, dbias, dpos, dimg, B, Cin, img_dims, patch, grid,
                                 N, nullptr, 0, stream);
}

int lcbi_patch_embed_bwd_ws(const void* img, int img_is_bf16, const float* w, const void* dout, int dout_is_bf16,
                            float* dw, float* dbias, float* dpos, float* dimg, int B, int Cin, const int* img_dims,
                            const int* patch, const int* grid, int N, void* workspace, size_t workspace_bytes, void* stream) {
  if (!img || !w || !dout || !dw || !img_dims || !patch || !grid)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_patch_embed_bwd: null pointer argument");
  int rc = patch_embed_bwd_launch(img, img_is_bf16, w, dout, dout_is_bf16, dw, dbias, dpos, dimg, B, Cin, img_dims,
                                  patch, grid, N, static_cast<cudaStream_t>(stream), workspace, workspace_bytes);
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_patch_embed_bwd: non-positive size or misaligned pointer");
  if (rc == LCBI_ERR_WORKSPACE) return fail(rc, "lcbi_patch_embed_bwd_ws: workspace too small or misaligned");
  if (rc == LCBI_ERR_TENSOR_MAP) return fail(rc, "lcbi_patch_embed_bwd_ws: TMA tensor map encode failed");
  return rc;
}

int lcbi_attn_merge(float* acc, float* lse_acc, const void* o_s, const float* lse_s, void* out_bf16, int B, int N, int H,
                    int head_dim, int first, void* stream) {
  if (!acc || !lse_acc || !o_s || !lse_s) return fail(LCBI_ERR_BAD_ARG, "lcbi_attn_merge: null pointer argument");
  int rc = attn_merge_launch(acc, lse_acc, o_s, lse_s, out_bf16, B, N, H, head_dim, first, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_UNSUPPORTED) return fail(rc, "lcbi_attn_merge: only head_dim == 64 is implemented");
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_attn_merge: non-positive size");
  return rc;
}

static int win_status(int rc, const char* who) {
  static thread_local char buf[256];
  const char* why = nullptr;
  if (rc == LCBI_ERR_UNSUPPORTED) why = "head_dim must be 16 or 32 and the window must hold <= 512 tokens";
  if (rc == LCBI_ERR_BAD_ARG) why = "bad ndim / grid / window / shift / batch / heads";
  if (!why) return rc;
  std::snprintf(buf, sizeof(buf), "%s: %s", who, why);
  return fail(rc, buf);
}

int lcbi_win_attn_fwd(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                      float scale, const void* qkv, const float* qkv_bias, const float* table, void* out, float* lse2,
                      void* stream) {
  return lcbi_win_attn_fwd_range(ndim, grid, window, shift, B, H, head_dim, scale, qkv, qkv_bias, table, out, lse2, 0, -1,
                                 stream);
}

int lcbi_win_attn_fwd_range(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                            float scale, const void* qkv, const float* qkv_bias, const float* table, void* out,
                            float* lse2, int win_begin, int win_count, void* stream) {
  if (!grid || !window || !shift || !qkv || !table || !out || !lse2)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_win_attn_fwd: null pointer argument");
  if (win_count == 0) return LCBI_OK;
  WinAttnArgs a{};
  a.ndim = ndim; a.grid = grid; a.window = window; a.shift = shift;
  a.B = B; a.H = H; a.head_dim = head_dim; a.scale = scale; a.win_begin = win_begin; a.win_count = win_count;
  a.qkv = qkv; a.qkv_bias = qkv_bias; a.table = table; a.out = out; a.lse2 = lse2;
  return win_status(win_attn_fwd_launch(a, static_cast<cudaStream_t>(stream)), "lcbi_win_attn_fwd");
}

int lcbi_win_attn_bwd(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                      float scale, const void* qkv, const float* qkv_bias, const float* table, const void* o,
                      const float* lse2, const void* d_out, float* dsum, void* dqkv, float* dbias_pad, float* dtable,
                      void* stream) {
  return lcbi_win_attn_bwd_range(ndim, grid, window, shift, B, H, head_dim, scale, qkv, qkv_bias, table, o, lse2, d_out,
                                 dsum, dqkv, dbias_pad, dtable, 0, -1, stream);
}

int lcbi_win_attn_bwd_range(int ndim, const int* grid, const int* window, const int* shift, int B, int H, int head_dim,
                            float scale, const void* qkv, const float* qkv_bias, const float* table, const void* o,
                            const float* lse2, const void* d_out, float* dsum, void* dqkv, float* dbias_pad,
                            float* dtable, int win_begin, int win_count, void* stream) {
  if (!grid || !window || !shift || !qkv || !table || !o || !lse2 || !d_out || !dsum || !dqkv)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_win_attn_bwd: null pointer argument");
  if (win_count == 0) return LCBI_OK;
  WinAttnArgs a{};
  a.ndim = ndim; a.grid = grid; a.window = window; a.shift = shift;
  a.B = B; a.H = H; a.head_dim = head_dim; a.scale = scale; a.win_begin = win_begin; a.win_count = win_count;
  a.qkv = qkv; a.qkv_bias = qkv_bias; a.table = table; a.lse2 = const_cast<float*>(lse2);
  a.d_out = d_out; a.dsum = dsum; a.dqkv = dqkv; a.dbias_pad = dbias_pad; a.dtable = dtable;
  return win_status(win_attn_bwd_launch(a, o, static_cast<cudaStream_t>(stream)), "lcbi_win_attn_bwd");
}

int lcbi_window_maps(int ndim, const int* grid, const int* window, const int* shift, int* gather, int* region,
                     int* relidx, int* n_out, int* nw_out, void* stream) {
  if (!grid || !window || !shift) return fail(LCBI_ERR_BAD_ARG, "lcbi_window_maps: null pointer argument");
  return win_status(window_maps_launch(ndim, grid, window, shift, gather, region, relidx, n_out, nw_out,
                                       static_cast<cudaStream_t>(stream)), "lcbi_window_maps");
}

static int ln_status(int rc, const char* who) {
  static thread_local char buf[256];
  const char* why = nullptr;
  if (rc == LCBI_ERR_UNSUPPORTED) why = "the channel count must be a multiple of 4";
  if (rc == LCBI_ERR_BAD_ARG) why = "non-positive rows / channels, negative eps, or a pointer that is not 16-byte aligned";
  if (rc == LCBI_ERR_WORKSPACE) why = "workspace missing, misaligned or smaller than lcbi_layer_norm_bwd_workspace_bytes";
  if (!why) return rc;
  std::snprintf(buf, sizeof(buf), "%s: %s", who, why);
  return fail(rc, buf);
}

int lcbi_layer_norm_fwd(const void* x, int x_is_bf16, const float* gamma, const float* beta, void* y, int y_is_bf16,
                        float* mean, float* rstd, int64_t rows, int C, float eps, void* stream) {
  if (!x || !y || !mean || !rstd) return fail(LCBI_ERR_BAD_ARG, "lcbi_layer_norm_fwd: null pointer argument");
  return ln_status(layer_norm_fwd_launch(x, x_is_bf16, gamma, beta, y, y_is_bf16, mean, rstd, rows, C, eps,
                                         static_cast<cudaStream_t>(stream)), "lcbi_layer_norm_fwd");
}

size_t lcbi_layer_norm_bwd_workspace_bytes(int64_t rows, int C) { return layer_norm_bwd_workspace_bytes(rows, C); }

int lcbi_layer_norm_bwd(const void* dy, int dy_is_bf16, const void* x, int x_is_bf16, const float* gamma,
                        const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta, void* workspace,
                        size_t workspace_bytes, int64_t rows, int C, void* stream) {
  if (!dy || !x || !mean || !rstd) return fail(LCBI_ERR_BAD_ARG, "lcbi_layer_norm_bwd: null pointer argument");
  return ln_status(layer_norm_bwd_launch(dy, dy_is_bf16, x, x_is_bf16, gamma, mean, rstd, dx, dgamma, dbeta,
                                          static_cast<float*>(workspace), workspace_bytes, rows, C,
                                          static_cast<cudaStream_t>(stream)), "lcbi_layer_norm_bwd");
}

int lcbi_add_layer_norm_fwd(const void* x, int x_is_bf16, const void* delta, int delta_is_bf16, void* xsum,
                            const float* gamma, const float* beta, void* y, int y_is_bf16, float* mean, float* rstd,
                            int64_t rows, int C, float eps, void* stream) {
  if (!x || !delta || !xsum || !y || !mean || !rstd)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_add_layer_norm_fwd: null pointer argument");
  const int rc = add_layer_norm_fwd_launch(x, x_is_bf16, delta, delta_is_bf16, xsum, gamma, beta, y, y_is_bf16, mean,
                                           rstd, rows, C, eps, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_UNSUPPORTED)
    return fail(rc, "lcbi_add_layer_norm_fwd: the channel count must be a multiple of 4 and at least 128");
  return ln_status(rc, "lcbi_add_layer_norm_fwd");
}

int lcbi_add_layer_norm_bwd(const void* dy, int dy_is_bf16, const void* dxsum, const void* xsum, int x_is_bf16,
                            const float* gamma, const float* mean, const float* rstd, void* dx, void* ddelta,
                            int delta_is_bf16, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                            int64_t rows, int C, void* stream) {
  if (!dy || !xsum || !mean || !rstd || !dx)
    return fail(LCBI_ERR_BAD_ARG, "lcbi_add_layer_norm_bwd: null pointer argument");
  const int rc = add_layer_norm_bwd_launch(dy, dy_is_bf16, dxsum, xsum, x_is_bf16, gamma, mean, rstd, dx, ddelta,
                                           delta_is_bf16, dgamma, dbeta, static_cast<float*>(workspace),
                                           workspace_bytes, rows, C, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_UNSUPPORTED)
    return fail(rc, "lcbi_add_layer_norm_bwd: the channel count must be a multiple of 4 and at least 128");
  return ln_status(rc, "lcbi_add_layer_norm_bwd");
}

int lcbi_gather_rows(const void* src, const int64_t* ids, void* dst, int64_t n_rows, int row_bytes, void* stream) {
  if (n_rows > 0 && (!src || !ids || !dst)) return fail(LCBI_ERR_BAD_ARG, "lcbi_gather_rows: null pointer argument");
  int rc = row_copy_launch(src, ids, dst, n_rows, row_bytes, 0, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_gather_rows: rows must be a positive multiple of 16 bytes, 16-byte aligned");
  return rc;
}

int lcbi_scatter_rows(const void* src, const int64_t* ids, void* dst, int64_t n_rows, int row_bytes, void* stream) {
  if (n_rows > 0 && (!src || !ids || !dst)) return fail(LCBI_ERR_BAD_ARG, "lcbi_scatter_rows: null pointer argument");
  int rc = row_copy_launch(src, ids, dst, n_rows, row_bytes, 1, static_cast<cudaStream_t>(stream));
  if (rc == LCBI_ERR_BAD_ARG) return fail(rc, "lcbi_scatter_rows: rows must be a positive multiple of 16 bytes, 16-byte aligned");
  return rc;
}

int lcbi_bias_grad(const void* dy, int dy_is_bf16, float* dbias, void* workspace, size_t workspace_bytes, int64_t rows,
                   int C, void* stream) {
  if (!dy || !dbias) return fail(LCBI_ERR_BAD_ARG, "lcbi_bias_grad: null pointer argument");
  return ln_status(bias_grad_launch(dy, dy_is_bf16, dbias, static_cast<float*>(workspace), workspace_bytes, rows, C,
                                    static_cast<cudaStream_t>(stream)), "lcbi_bias_grad");
}

}  // extern "C"
