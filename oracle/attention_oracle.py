"""CPU oracle (plain PyTorch, fp32/fp64, differentiable) for the token-mixing hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs. The product package never imports this module and has no CPU fallback.

Restates, from the reference's algorithm (citations into /root/reference/model/models/):
  * SABlock.forward attention core         backbone_vit.py:191-201   -> dense_attention / sablock
  * WindowAttention.forward                backbone_swin.py:335-358  -> window_attention
  * SwinTransformerBlock.forward_part1     backbone_swin.py:435-487  -> swin_part1 (gather/scatter form)
  * MONAI 1.3.0 PatchEmbeddingBlock / PatchEmbed (absent third-party dependency, requirements.txt:5;
    call sites backbone_vit.py:351-361,383 and backbone_swin.py:800-806,885) -> patch_embed_vit / _swin
  * torch.nn.LayerNorm as applied by the blocks   backbone_vit.py:260-263, backbone_swin.py:437,489 -> layer_norm_rows

Pinned against outputs of the unmodified reference run in the build container
(oracle/make_golden.py -> tests/golden/*.npz; checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import window_maps as wm


# ------------------------------------------------------------------------------------------------
# dense (ViT) attention
# ------------------------------------------------------------------------------------------------
def dense_attention(q, k, v, scale):
    """q,k,v: (B,H,N,d). softmax(scale * q k^T) v; the scale multiplies AFTER the product
    (backbone_vit.py:193)."""
    s = torch.einsum("bhxd,bhyd->bhxy", q, k) * scale
    p = s.softmax(dim=-1)
    return torch.einsum("bhxy,bhyd->bhxd", p, v)


def split_qkv_vit(qkv, num_heads):
    """(B,N,3C) -> q,k,v (B,H,N,d); feature index = s*C + h*d + j ("b h (qkv l d) -> qkv b l h d",
    backbone_vit.py:168)."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // num_heads
    t = qkv.reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    return t[0], t[1], t[2]


def sablock(x, w_qkv, w_out, b_out, num_heads):
    """backbone_vit.py:189-203 with dropout p=0: qkv Linear (no bias) -> attention -> out_proj."""
    B, N, C = x.shape
    q, k, v = split_qkv_vit(F.linear(x, w_qkv), num_heads)
    o = dense_attention(q, k, v, (C // num_heads) ** -0.5)
    o = o.permute(0, 2, 1, 3).reshape(B, N, C)
    return F.linear(o, w_out, b_out)


def dense_attention_rows(q_rows, k, v, scale, chunk=8192):
    """Chunked fp64 attention for a subset of query rows (used to spot-check very long sequences where
    the N x N matrix cannot be materialised). q_rows: (R,d), k,v: (N,d)."""
    q_rows, k, v = q_rows.double(), k.double(), v.double()
    m = torch.full((q_rows.shape[0],), -float("inf"), dtype=torch.float64)
    l = torch.zeros_like(m)
    acc = torch.zeros(q_rows.shape[0], v.shape[1], dtype=torch.float64)
    for s in range(0, k.shape[0], chunk):
        sc = (q_rows @ k[s:s + chunk].T) * scale
        m_new = torch.maximum(m, sc.max(-1).values)
        a = torch.exp(m - m_new)
        p = torch.exp(sc - m_new[:, None])
        l = l * a + p.sum(-1)
        acc = acc * a[:, None] + p @ v[s:s + chunk]
        m = m_new
    return acc / l[:, None], m + torch.log(l)


# ------------------------------------------------------------------------------------------------
# window (Swin) attention
# ------------------------------------------------------------------------------------------------
def window_attention(xw, w_qkv, b_qkv, table, index_nn, mask, w_proj, b_proj, num_heads):
    """backbone_swin.py:335-358. xw: (B*nW, n, C); index_nn: (n,n) long (already `[:n,:n]`-sliced);
    mask: (nW,n,n) or None. q is scaled BEFORE q k^T (:341)."""
    b, n, c = xw.shape
    d = c // num_heads
    qkv = F.linear(xw, w_qkv, b_qkv).reshape(b, n, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
    s = q @ k.transpose(-2, -1)
    bias = table[index_nn.reshape(-1)].reshape(n, n, num_heads).permute(2, 0, 1)
    s = s + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        s = (s.view(b // nw, nw, num_heads, n, n) + mask[None, :, None]).view(b, num_heads, n, n)
    p = s.softmax(dim=-1)
    o = (p @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(o, w_proj, b_proj)


def swin_part1(x, window_ctor, shift_ctor, w_qkv, b_qkv, table, w_proj, b_proj, num_heads,
               norm_w=None, norm_b=None):
    """SwinTransformerBlock.forward_part1 (backbone_swin.py:435-487) written as ONE gather and ONE
    scatter through oracle.window_maps.gather_map instead of pad/roll/partition/reverse/roll/crop.
    x: (B, *grid, C) channel-last. norm_w/norm_b: norm1 affine (LayerNorm over C) or None to skip."""
    B, C = x.shape[0], x.shape[-1]
    grid = tuple(x.shape[1:-1])
    win, sh = wm.resolve_window(grid, window_ctor, shift_ctor)
    n = int(np.prod(win))
    gmap = torch.from_numpy(wm.gather_map(grid, window_ctor, shift_ctor)).to(x.device)  # (nW, n)
    nW = gmap.shape[0]
    xn = F.layer_norm(x, (C,), norm_w, norm_b) if norm_w is not None else x
    flat = xn.reshape(B, -1, C)
    valid = gmap >= 0
    idx = gmap.clamp(min=0)
    xw = flat[:, idx.reshape(-1)].reshape(B, nW, n, C) * valid[None, :, :, None].to(x.dtype)  # pad rows = 0
    index_nn = torch.from_numpy(wm.rel_pos_index_used(window_ctor, n)).to(x.device)
    mask = (torch.from_numpy(wm.shift_mask(grid, window_ctor, shift_ctor)).to(device=x.device, dtype=x.dtype)
            if any(s > 0 for s in sh) else None)
    yw = window_attention(xw.reshape(B * nW, n, C), w_qkv, b_qkv, table, index_nn, mask, w_proj, b_proj,
                          num_heads).reshape(B, nW * n, C)
    out = torch.zeros_like(flat)
    sel = valid.reshape(-1)
    out[:, idx.reshape(-1)[sel]] = yw[:, sel]
    return out.reshape(x.shape)


def window_attention_core(qkv, qkv_bias, table, grid, window_ctor, shift_ctor, num_heads, scale=None, win_range=None):
    """The part of forward_part1 between the qkv Linear and the proj Linear (backbone_swin.py:339-357 inside
    :441-485), on the un-padded token grid: qkv (B, *grid, 3C) with feature index s*C + h*d + j (:339). Pad tokens
    enter the reference as zeros BEFORE the Linear (:441-455), so their q/k/v rows are the Linear's bias.
    win_range = (begin, count) keeps only the windows [begin, begin+count) of the flattened (batch, window) list and
    returns zeros for the tokens of all other windows (window-sharded execution)."""
    B, C = qkv.shape[0], qkv.shape[-1] // 3
    d = C // num_heads
    scale = d ** -0.5 if scale is None else scale
    grid = tuple(int(g) for g in grid)
    win, sh = wm.resolve_window(grid, window_ctor, shift_ctor)
    n = int(np.prod(win))
    gmap = torch.from_numpy(wm.gather_map(grid, window_ctor, shift_ctor)).to(qkv.device)  # (nW, n)
    nW = gmap.shape[0]
    flat = qkv.reshape(B, -1, 3 * C)
    valid = gmap >= 0
    idx = gmap.clamp(min=0)
    rows = flat[:, idx.reshape(-1)].reshape(B, nW, n, 3 * C)
    pad_row = qkv_bias if qkv_bias is not None else torch.zeros(3 * C, dtype=qkv.dtype, device=qkv.device)
    rows = torch.where(valid[None, :, :, None], rows, pad_row.to(rows.dtype).expand_as(rows))
    t = rows.reshape(B * nW, n, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = t[0] * scale, t[1], t[2]
    s = q @ k.transpose(-2, -1)
    index_nn = torch.from_numpy(wm.rel_pos_index_used(window_ctor, n)).to(qkv.device)
    s = s + table[index_nn.reshape(-1)].reshape(n, n, num_heads).permute(2, 0, 1).unsqueeze(0)
    if any(x > 0 for x in sh):
        mask = torch.from_numpy(wm.shift_mask(grid, window_ctor, shift_ctor)).to(device=qkv.device, dtype=s.dtype)
        s = (s.view(B, nW, num_heads, n, n) + mask[None, :, None]).view(B * nW, num_heads, n, n)
    o = (s.softmax(dim=-1) @ v).transpose(1, 2).reshape(B, nW, n, C)
    if win_range is not None:
        unit = torch.arange(B * nW, device=qkv.device).reshape(B, nW)
        keep = (unit >= win_range[0]) & (unit < win_range[0] + win_range[1])
        o = o * keep[:, :, None, None].to(o.dtype)
    out = torch.zeros(B, flat.shape[1], C, dtype=o.dtype, device=qkv.device)
    sel = valid.reshape(-1)
    out[:, idx.reshape(-1)[sel]] = o.reshape(B, nW * n, C)[:, sel]
    return out.reshape(B, *grid, C)


# ------------------------------------------------------------------------------------------------
# patch embedding (MONAI 1.3.0 semantics)
# ------------------------------------------------------------------------------------------------
def patch_embed_vit(img, weight, bias, pos):
    """conv(kernel=stride=patch) -> flatten(2).transpose -> + position_embeddings. img (B,Cin,*sp)."""
    patch = tuple(weight.shape[2:])
    conv = F.conv3d if len(patch) == 3 else F.conv2d
    y = conv(img, weight, bias, stride=patch)
    return y.flatten(2).transpose(-1, -2) + pos


def patch_embed_swin(img, weight, bias):
    """trailing zero-pad to a patch multiple -> conv(kernel=stride=patch); channel-first output."""
    patch = tuple(weight.shape[2:])
    pads = []
    for size, p in zip(reversed(img.shape[2:]), reversed(patch)):
        pads += [0, (p - size % p) % p]
    if any(pads):
        img = F.pad(img, pads)
    conv = F.conv3d if len(patch) == 3 else F.conv2d
    return conv(img, weight, bias, stride=patch)


# ------------------------------------------------------------------------------------------------
# deterministic parameter fill shared by the golden generator and the tests
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# LayerNorm in front of the qkv projection / the MLP
# ------------------------------------------------------------------------------------------------
def layer_norm_rows(x, weight=None, bias=None, eps=1e-5):
    """torch.nn.LayerNorm(C) as the encoder blocks apply it to token rows (norm1 / norm2 at backbone_vit.py:260-263,
    backbone_swin.py:437,489; modules built at backbone_vit.py:249-258): per row, mean and BIASED variance over the
    channel axis, (x - mean) / sqrt(var + eps), then the affine. Written out (not F.layer_norm) so that the test
    that pins it against torch.nn.LayerNorm means something."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return y


def fill_parameters_(module, seed, std=0.05):
    """Overwrites every floating-point parameter with seeded N(0, std) values (LayerNorm weights get
    1 + N(0, std)), in state_dict order. Integer buffers are left alone. Makes golden fixtures
    reproducible without storing state_dicts; relative_position_bias_table gets std 0.5 so that the
    bias path is numerically visible."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            s = 0.5 if name.endswith("relative_position_bias_table") else std
            vals = torch.randn(p.shape, generator=g, dtype=torch.float32) * s
            if "norm" in name and name.endswith("weight"):
                vals = vals + 1.0
            p.copy_(vals.to(p.dtype))
    return module
