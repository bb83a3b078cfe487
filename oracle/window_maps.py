"""CPU oracle (numpy, integer arithmetic) for the Swin window / shift / bias index maps.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
The product package never imports this module.

Closed-form restatement of the index movement the reference performs with pad/roll/view/permute copies
(all citations are into /root/reference/model/models/backbone_swin.py):

  * get_window_size                  :200-224   -> resolve_window
  * forward_part1 pad + roll + window_partition (+ reverse/roll/crop)  :435-487, :135-197  -> gather_map
  * compute_mask                     :591-628   -> region_ids / shift_mask
  * WindowAttention.__init__ relative_position_index  :256-308 (+ the [:n,:n] slice at :343-345)
                                                 -> rel_pos_index / rel_pos_index_used

Pinned bit-exactly against the reference's own functions by oracle/make_golden.py ->
tests/golden/window_maps.npz (see tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np


def resolve_window(grid, window, shift):
    """Per-axis clamp (reference :200-224): grid_k <= window_k -> window_k = grid_k and shift_k = 0."""
    win, sh = [], []
    for g, w, s in zip(grid, window, shift):
        if g <= w:
            win.append(int(g))
            sh.append(0)
        else:
            win.append(int(w))
            sh.append(int(s))
    return tuple(win), tuple(sh)


def padded_grid(grid, win):
    """Far-end zero padding to a window multiple (reference :441-445, :452-455)."""
    return tuple(int(-(-g // w) * w) for g, w in zip(grid, win))


def _window_token_coords(grid, window, shift):
    """Coordinates, in the UNSHIFTED padded frame, of every (window, token) slot.

    Returns (coords [nW, n, K] int64, win, sh, padded). Window order is raster over the window grid,
    token order raster inside the window (reference permute at :159 / :164); the shifted frame position
    p maps to source position (p + s) mod Lp because the reference rolls by -s (:461-463).
    """
    win, sh = resolve_window(grid, window, shift)
    pg = padded_grid(grid, win)
    K = len(grid)
    nwin = [p // w for p, w in zip(pg, win)]
    wgrid = np.stack(np.meshgrid(*[np.arange(c) for c in nwin], indexing="ij"), -1).reshape(-1, K)
    tgrid = np.stack(np.meshgrid(*[np.arange(w) for w in win], indexing="ij"), -1).reshape(-1, K)
    shifted = wgrid[:, None, :] * np.asarray(win)[None, None, :] + tgrid[None, :, :]  # position in rolled frame
    src = (shifted + np.asarray(sh)[None, None, :]) % np.asarray(pg)[None, None, :]
    return src.astype(np.int64), shifted.astype(np.int64), win, sh, pg


def gather_map(grid, window, shift):
    """[nW, n] int64: flat row-major index into the un-padded token grid feeding each window slot,
    -1 where the slot is a zero pad token. The attention output of a slot is written back to the same
    source index (reverse + roll(+s) + crop, reference :470-485), pad slots are dropped."""
    src, _, win, sh, pg = _window_token_coords(grid, window, shift)
    g = np.asarray(grid)
    inside = (src < g[None, None, :]).all(-1)
    strides = np.ones(len(grid), dtype=np.int64)
    for k in range(len(grid) - 2, -1, -1):
        strides[k] = strides[k + 1] * grid[k + 1]
    flat = (src * strides[None, None, :]).sum(-1)
    return np.where(inside, flat, -1).astype(np.int64)


def region_ids(grid, window, shift):
    """[nW, n] int32 region id of every window slot (reference compute_mask :604-624).

    Along axis k (padded length L, window W, shift s, measured in the SHIFTED frame):
    reg = 0 if p < L-W, 1 if p < L-s, else 2; with s == 0 the slices `slice(-W, -0)` is empty and
    `slice(-0, None)` covers everything, so every position gets the LAST loop value, i.e. reg = 2 — a
    constant, so equality between positions is unaffected. id = sum_k reg_k * 3^(K-1-k) (counter order of
    the nested loops)."""
    _, shifted, win, sh, pg = _window_token_coords(grid, window, shift)
    K = len(grid)
    ids = np.zeros(shifted.shape[:2], dtype=np.int64)
    for k in range(K):
        p = shifted[..., k]
        L, W, s = pg[k], win[k], sh[k]
        if s == 0:
            reg = np.full_like(p, 2)
        else:
            reg = np.where(p < L - W, 0, np.where(p < L - s, 1, 2))
        ids = ids * 3 + reg
    return ids.astype(np.int32)


def shift_mask(grid, window, shift):
    """[nW, n, n] float32: 0 where the two slots share a region, -100 otherwise (reference :625-626).
    mask[w, i, j] = f(id[w, j] - id[w, i]) — symmetric."""
    ids = region_ids(grid, window, shift)
    neq = ids[:, None, :] != ids[:, :, None]
    return np.where(neq, np.float32(-100.0), np.float32(0.0)).astype(np.float32)


def rel_pos_index(window_ctor):
    """[n, n] int64 relative-position index built from the CONSTRUCTOR window size (reference :268-307):
    idx[i, j] = sum_k (c_k(i) - c_k(j) + W_k - 1) * prod_{m>k} (2 W_m - 1)."""
    W = [int(w) for w in window_ctor]
    K = len(W)
    coords = np.stack(np.meshgrid(*[np.arange(w) for w in W], indexing="ij"), 0).reshape(K, -1)
    rel = coords[:, :, None] - coords[:, None, :]
    idx = np.zeros(rel.shape[1:], dtype=np.int64)
    for k in range(K):
        stride = 1
        for m in range(k + 1, K):
            stride *= 2 * W[m] - 1
        idx += (rel[k] + W[k] - 1) * stride
    return idx


def rel_pos_index_used(window_ctor, n):
    """The `[:n, :n]` slice the reference applies at forward time (:343-345) when the window was clamped:
    the first n flattened positions of the UN-clamped geometry."""
    return rel_pos_index(window_ctor)[:n, :n]


def bias_table_rows(window_ctor):
    out = 1
    for w in window_ctor:
        out *= 2 * int(w) - 1
    return out
