// Shared pieces of the Swin window-attention kernels: window geometry (closed-form index maps) and
// warp-level bf16 MMA fragments.
//
// Geometry restates, as integer arithmetic evaluated inside the kernels, what the reference builds with
// pad / roll / view / permute copies (/root/reference/model/models/backbone_swin.py):
//   get_window_size :200-224, forward_part1 pad+roll+window_partition :435-468, window_reverse+roll+crop :470-485,
//   compute_mask :591-628, relative_position_index :256-308 and its [:n,:n] slice :343-345.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace lcbi {

// Division by a runtime-constant divisor as multiply-high + shift (dividend < 2^31): the per-slot index maps below
// unravel slot and window numbers along three axes, and plain runtime integer division made that ~20 % of all
// instructions of the window kernels (ncu source view, profiles/r01_ncu_swin4_bwd_hotlines.txt).
struct FastDiv {
  uint32_t mul, shr, div;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.div = static_cast<uint32_t>(d);
  if (d <= 1) {
    f.mul = 0; f.shr = 0;
    return f;
  }
  uint32_t lg = 0;
  while ((1u << lg) < static_cast<uint32_t>(d)) ++lg;      // ceil(log2 d)
  const uint32_t p = 31 + lg;
  f.mul = static_cast<uint32_t>(((1ull << p) + static_cast<uint32_t>(d) - 1) / static_cast<uint32_t>(d));
  f.shr = p - 32;
  return f;
}
__device__ __forceinline__ int fdiv(int x, const FastDiv& f) {
  return f.div == 1 ? x : static_cast<int>(__umulhi(static_cast<uint32_t>(x), f.mul) >> f.shr);
}
__device__ __forceinline__ void fdivmod(int x, const FastDiv& f, int& q, int& r) {
  q = fdiv(x, f);
  r = x - q * static_cast<int>(f.div);
}

struct WinGeom {
  int grid[3];      // un-padded token grid (D,H,W); D == 1 for 2-D
  int win[3];       // window actually used (clamped to the grid)
  int shift[3];     // shift actually used (0 on clamped axes)
  int pg[3];        // padded grid (multiple of win)
  int nwin[3];      // windows per axis
  int ctor[3];      // constructor window size: defines the relative-position index geometry
  int n;            // tokens per window
  int nW;           // windows per image
  int T;            // tokens per image
  int tab_rows;     // prod(2*ctor-1)
  FastDiv d_win[3], d_ctor[3], d_nwin[3], d_nW;   // divisors of the index maps
};

// host: fills g from (grid, constructor window, constructor shift); ndim 2 or 3 (2-D gets a leading 1)
inline int fill_win_geom(WinGeom& g, int ndim, const int* grid, const int* window, const int* shift) {
  if (ndim != 2 && ndim != 3) return -1;
  const int off = 3 - ndim;
  g.grid[0] = 1; g.ctor[0] = 1; g.win[0] = 1; g.shift[0] = 0;
  for (int i = 0; i < ndim; ++i) {
    if (grid[i] <= 0 || window[i] <= 0 || shift[i] < 0) return -1;
    g.grid[off + i] = grid[i];
    g.ctor[off + i] = window[i];
    if (grid[i] <= window[i]) {        // reference :215-219
      g.win[off + i] = grid[i];
      g.shift[off + i] = 0;
    } else {
      g.win[off + i] = window[i];
      g.shift[off + i] = shift[i];
    }
  }
  g.n = 1; g.nW = 1; g.T = 1; g.tab_rows = 1;
  for (int k = 0; k < 3; ++k) {
    g.pg[k] = (g.grid[k] + g.win[k] - 1) / g.win[k] * g.win[k];
    g.nwin[k] = g.pg[k] / g.win[k];
    g.n *= g.win[k];
    g.nW *= g.nwin[k];
    g.T *= g.grid[k];
    g.tab_rows *= 2 * g.ctor[k] - 1;
  }
  for (int k = 0; k < 3; ++k) {
    g.d_win[k] = make_fastdiv(g.win[k]);
    g.d_ctor[k] = make_fastdiv(g.ctor[k]);
    g.d_nwin[k] = make_fastdiv(g.nwin[k]);
  }
  g.d_nW = make_fastdiv(g.nW);
  return 0;
}

// token index feeding (window w, slot s), -1 for a zero-pad token; also the shift-mask region id of the slot
__device__ __forceinline__ void slot_lookup(const WinGeom& g, int w, int s, int& tok, int& region) {
  int w0, w1, w2, t0, t1, t2, q;
  fdivmod(w, g.d_nwin[2], q, w2);
  fdivmod(q, g.d_nwin[1], w0, w1);
  fdivmod(s, g.d_win[2], q, t2);
  fdivmod(q, g.d_win[1], t0, t1);
  const int wc[3] = {w0, w1, w2}, tc[3] = {t0, t1, t2};
  int src[3];
  bool inside = true;
  region = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int p = wc[k] * g.win[k] + tc[k];               // position in the rolled (shifted) padded frame
    int s_k = p + g.shift[k];
    if (s_k >= g.pg[k]) s_k -= g.pg[k];                    // roll by -shift (reference :461-463)
    src[k] = s_k;
    inside = inside && (s_k < g.grid[k]);
    const int reg = (g.shift[k] == 0) ? 2 : (p < g.pg[k] - g.win[k] ? 0 : (p < g.pg[k] - g.shift[k] ? 1 : 2));
    region = region * 3 + reg;                              // reference :609-613 counter order
  }
  tok = inside ? (src[0] * g.grid[1] + src[1]) * g.grid[2] + src[2] : -1;
}

// relative-position index = row_term(i) - col_term(j), slots unravelled in the CONSTRUCTOR window
__device__ __forceinline__ void relpos_terms(const WinGeom& g, int s, int& row_term, int& col_term) {
  int c0, c1, c2, q;
  fdivmod(s, g.d_ctor[2], q, c2);
  fdivmod(q, g.d_ctor[1], c0, c1);
  const int st1 = 2 * g.ctor[2] - 1, st0 = st1 * (2 * g.ctor[1] - 1);
  col_term = c0 * st0 + c1 * st1 + c2;
  row_term = col_term + (g.ctor[0] - 1) * st0 + (g.ctor[1] - 1) * st1 + (g.ctor[2] - 1);
}

// ---------------------------------------------------------------------------------------------
// kernel parameters shared by window_attn.cu and window_attn_small.cu
// ---------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kKeyChunk = 64;

struct WinParams {
  WinGeom g;
  int B, H, C;
  float scale_log2;             // head_dim^-0.5 * log2(e)
  const __nv_bfloat16* qkv;     // (B, T, 3, H, D)
  const float* qkv_bias;        // (3*C) or nullptr
  const float* table;           // (tab_rows, H)
  __nv_bfloat16* out;           // (B, T, C)                      [fwd]
  float* lse2;                  // (B, T, H) log2-domain logsumexp [fwd out / bwd in]
  // backward
  const __nv_bfloat16* d_out;   // (B, T, C)
  const float* dsum;            // (B, T, H) rowsum(dO o O)
  __nv_bfloat16* dqkv;          // (B, T, 3, H, D)
  float* dbias_pad;             // (3*C) fp32, += gradient reaching qkv.bias through the pad tokens
  float* dtable;                // (tab_rows, H) fp32, +=
  int win_splits;               // dq kernel: number of window subsets
  int win_begin, win_count;     // range of the flattened (batch, window) list this launch processes
};

template <int D>
struct Tile {
  static constexpr int kStride = D * 2 + 16;   // bytes per smem row: +16 keeps ldmatrix conflict-free
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }


// ---------------------------------------------------------------------------------------------
// warp-level MMA helpers (m16n8k16, bf16 x bf16 -> fp32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A fragment (16 rows x 16 k) of a row-major smem tile: row0 = first row, k0 = first column
template <int STRIDE_BYTES>
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], uint32_t tile_base, int row0, int k0, int lane) {
  const uint32_t addr = tile_base + (row0 + (lane & 15)) * STRIDE_BYTES + (k0 + (lane >> 4) * 8) * 2;
  ldsm_x4(a, addr);
}
// B fragment for X^T as B (B[k][n] = X[n0+n][k0+k]): 8 rows of X, 16 columns -> (b0, b1)
template <int STRIDE_BYTES>
__device__ __forceinline__ void load_b_frag_nt(uint32_t& b0, uint32_t& b1, uint32_t tile_base, int n0, int k0, int lane) {
  const uint32_t addr = tile_base + (n0 + (lane & 7)) * STRIDE_BYTES + (k0 + ((lane >> 3) & 1) * 8) * 2;
  ldsm_x2(b0, b1, addr);
}
// B fragment for X as B (B[k][n] = X[k0+k][n0+n]): 16 rows of X, 8 columns -> (b0, b1) via transposing load
template <int STRIDE_BYTES>
__device__ __forceinline__ void load_b_frag_t(uint32_t& b0, uint32_t& b1, uint32_t tile_base, int k0, int n0, int lane) {
  const uint32_t addr = tile_base + (k0 + (lane & 15)) * STRIDE_BYTES + n0 * 2;
  ldsm_x2_trans(b0, b1, addr);
}

}  // namespace lcbi
