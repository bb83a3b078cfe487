"""Swin encoder: host-side mirror of the reference interface, B200-native shifted-window attention.

Mirrors /root/reference/model/models/backbone_swin.py so `EncoderDecoderModel` (reference
model/model_base.py:44-47) can use it unchanged:

  custom_Swin(config, input_feature_channels) -> (model, [C, 2C, 4C, 8C, 16C])      reference :44-129
  window_partition / window_reverse / get_window_size / compute_mask                 reference :135-224, :591-628
  WindowAttention / SwinTransformerBlock / PatchMergingV2 / BasicLayer /
  SwinTransformer_with_alt_ops                                                       reference :227-911

Same constructor arguments, same state_dict keys and shapes (patch_embed.proj.*, layers{1..4}.0.blocks.{j}.
{norm1,norm2}.*, .attn.relative_position_bias_table, .attn.relative_position_index (int64 buffer),
.attn.qkv.*, .attn.proj.*, .mlp.linear{1,2}.*, layers{s}.0.downsample.{reduction.weight,norm.*}), same
6-element hidden-state list (channel-first, layer-normed copies; unsqueeze(2) for 2-D).

What changes: `forward_part1` is norm1 -> qkv Linear -> ONE fused kernel (window gather with cyclic shift,
QK^T + relative-position bias + shift mask, softmax, PV, scatter back) -> proj Linear. There is no F.pad,
torch.roll, window_partition, window_reverse, mask tensor or per-forward bias gather; tokens stay channel-last
between stages and channel-first tensors are exposed as views exactly where the reference returns them.

Window parallelism (new relative to the reference): `window_group=<process group>` (or `config.Swin.window_group`)
splits the (batch, window) list of every attention block over the ranks of the group (`window_parallel.py`);
everything token-wise stays replicated, so every rank still returns the full hidden states.
"""
from __future__ import annotations

import itertools
from collections.abc import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint
from torch.nn import LayerNorm

from . import ops, window_parallel
from .blocks import MLPBlock as Mlp
from .blocks import PatchEmbed, apply_layer_norm

_ALT_MIXER_MSG = ("use_hyena/use_mamba route to the reference's HyenaOperator / MambaVisionMixer "
                  "(model/models/hyena.py, mamba.py), which are outside the B200 attention hot path; build the "
                  "reference encoder for those configurations")

_PRESETS = {
    "unetr": (48, [2, 2, 2, 2], [3, 6, 12, 24]),
    "tiny": (96, [2, 2, 6, 2], [3, 6, 12, 24]),
    "small": (96, [2, 2, 18, 2], [3, 6, 12, 24]),
    "base": (128, [2, 2, 18, 2], [4, 8, 16, 32]),
    "large": (192, [2, 2, 18, 2], [6, 12, 24, 48]),
}


def custom_Swin(config, input_feature_channels):
    """Factory with the reference's contract (backbone_swin.py:44-129)."""
    size = config.Swin.size
    if size in _PRESETS:
        embed_dim, depths, num_heads = _PRESETS[size]
        depths, num_heads = list(depths), list(num_heads)
        config.Swin.embed_dim = embed_dim
        config.Swin.depths = depths
        config.Swin.num_heads = num_heads
    elif size == "custom":
        embed_dim, depths, num_heads = config.Swin.embed_dim, config.Swin.depths, config.Swin.num_heads
    else:
        raise ValueError(f"Unknown model size {config.Swin.size} specified in config.")

    if config.time == 1:
        spatial_dims = 2
        if len(config.Swin.patch_size) == 3:
            patch, window = config.Swin.patch_size[1:], config.Swin.window_size[1:]
        else:
            patch, window = config.Swin.patch_size, config.Swin.window_size
    else:
        spatial_dims = 3
        patch, window = config.Swin.patch_size, config.Swin.window_size

    model = SwinTransformer_with_alt_ops(use_hyena=config.Swin.use_hyena, use_mamba=config.Swin.use_mamba,
                                         in_chans=input_feature_channels, embed_dim=embed_dim, window_size=window,
                                         patch_size=patch, depths=depths, num_heads=num_heads, spatial_dims=spatial_dims)
    if getattr(config.Swin, "window_group", None) is not None:
        model.set_window_group(config.Swin.window_group)
    n = len(depths)
    return model, [embed_dim * 2 ** i for i in range(n)] + [embed_dim * 2 ** n]


# --------------------------------------------------------------------------------------------------
# Index helpers with the reference's signatures (not used by the fused path; kept for callers and tests)
# --------------------------------------------------------------------------------------------------
def window_partition(x, window_size):
    """(B, [D,] H, W, C) -> (B*nW, n, C); windows raster-ordered, tokens raster-ordered inside a window."""
    if x.dim() == 5:
        b, d, h, w, c = x.shape
        wd, wh, ww = window_size
        x = x.reshape(b, d // wd, wd, h // wh, wh, w // ww, ww, c)
        return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, wd * wh * ww, c)
    b, h, w, c = x.shape
    wh, ww = window_size
    x = x.reshape(b, h // wh, wh, w // ww, ww, c)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, wh * ww, c)


def window_reverse(windows, window_size, dims):
    """Inverse of window_partition; dims = (B, [D,] H, W)."""
    if len(dims) == 4:
        b, d, h, w = dims
        wd, wh, ww = window_size
        x = windows.reshape(b, d // wd, h // wh, w // ww, wd, wh, ww, -1)
        return x.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(b, d, h, w, -1)
    b, h, w = dims
    wh, ww = window_size
    x = windows.reshape(b, h // wh, w // ww, wh, ww, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, -1)


def get_window_size(x_size, window_size, shift_size=None):
    """Per-axis clamp: an axis not larger than the window uses the whole axis and no shift."""
    use_window = [int(g) if g <= w else int(w) for g, w in zip(x_size, window_size)]
    if shift_size is None:
        return tuple(use_window)
    use_shift = [0 if g <= w else int(s) for g, w, s in zip(x_size, window_size, shift_size)]
    return tuple(use_window), tuple(use_shift)


def compute_mask(dims, window_size, shift_size, device):
    """(nW, n, n) additive shift mask (0 / -100) for an already padded grid `dims`, from the same region ids the
    kernels evaluate per slot. The fused path never materialises it."""
    dims = [int(d) for d in dims]
    coords = []
    for L, W, s in zip(dims, window_size, shift_size):
        p = torch.arange(L, device=device)
        coords.append(torch.full_like(p, 2) if s == 0 else torch.where(p < L - W, 0, torch.where(p < L - s, 1, 2)))
    ids = torch.zeros(dims, dtype=torch.long, device=device)
    for k, r in enumerate(coords):
        shape = [1] * len(dims)
        shape[k] = dims[k]
        ids = ids * 3 + r.reshape(shape)
    win = window_partition(ids.reshape(1, *dims, 1).float(), window_size).squeeze(-1)
    diff = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _relative_position_index(window_size):
    coords = torch.stack(torch.meshgrid(*[torch.arange(w) for w in window_size], indexing="ij")).flatten(1)
    rel = coords[:, :, None] - coords[:, None, :]
    idx = torch.zeros(rel.shape[1:], dtype=torch.long)
    for k, w in enumerate(window_size):
        stride = 1
        for m in window_size[k + 1:]:
            stride *= 2 * m - 1
        idx += (rel[k] + w - 1) * stride
    return idx


class WindowAttention(nn.Module):
    """Window attention parameters + fused kernel entry (reference backbone_swin.py:227-367)."""

    def __init__(self, use_hyena: bool, use_mamba: bool, dim: int, num_heads: int, window_size: Sequence[int],
                 qkv_bias: bool = False, attn_drop: float = 0.0, proj_drop: float = 0.0) -> None:
        super().__init__()
        if use_hyena or use_mamba:
            raise NotImplementedError(_ALT_MIXER_MSG)
        if attn_drop != 0.0 or proj_drop != 0.0:
            raise NotImplementedError("attention dropout is always 0 on the reference's reachable paths")
        if dim % num_heads != 0:
            raise ValueError("dim should be divisible by num_heads.")
        self.dim = dim
        self.use_hyena = use_hyena
        self.use_mamba = use_mamba
        self.window_size = tuple(int(w) for w in window_size)
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        rows = 1
        for w in self.window_size:
            rows *= 2 * w - 1
        self.relative_position_bias_table = nn.Parameter(torch.zeros(rows, num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        self.window_group = None     # process group the block's windows are sharded over (set_window_group)

    def forward_grid(self, x, shift_size):
        """x: (B, *grid, C) normed tokens on the un-padded grid -> (B, *grid, C). The whole of the reference's
        pad/roll/partition -> attention -> reverse/roll/crop chain."""
        qkv = ops.linear(x, self.qkv.weight, self.qkv.bias)
        grid = tuple(x.shape[1:-1])
        group = self.window_group
        if group is not None:
            import torch.distributed as dist

            pg = None if group == "world" else group
            # blocks with fewer windows than ranks (late stages: 27 / 8 windows) stay replicated
            if (dist.is_available() and dist.is_initialized() and dist.get_world_size(pg) > 1 and
                    x.shape[0] * window_parallel.count_windows(grid, self.window_size) >= dist.get_world_size(pg)):
                o = window_parallel.window_attention_sharded(qkv, self.qkv.bias, self.relative_position_bias_table,
                                                             grid, self.window_size, shift_size, self.num_heads,
                                                             self.scale, group=pg)
                return ops.linear(o, self.proj.weight, self.proj.bias)
        o = ops.window_attention(qkv, self.qkv.bias, self.relative_position_bias_table, grid,
                                 self.window_size, shift_size, self.num_heads, self.scale)
        return ops.linear(o, self.proj.weight, self.proj.bias)

    def forward(self, x, mask):
        """Reference seam on pre-partitioned windows (B*nW, n, C). Only mask=None is expressible without the token
        grid (each window is treated as its own un-shifted grid); shifted windows go through forward_grid."""
        if mask is not None:
            raise NotImplementedError("the fused kernel derives the shift mask from the token grid: call "
                                      "SwinTransformerBlock.forward_part1 / WindowAttention.forward_grid instead of "
                                      "passing pre-partitioned windows with a mask tensor")
        b, n, c = x.shape
        vol = 1
        for w in self.window_size:
            vol *= w
        if n != vol:
            raise NotImplementedError("pre-partitioned clamped windows need the token grid: use forward_grid")
        x = x.reshape(b, *self.window_size, c)
        return self.forward_grid(x, tuple(0 for _ in self.window_size)).reshape(b, n, c)


class SwinTransformerBlock(nn.Module):
    """Reference backbone_swin.py:370-537."""

    def __init__(self, use_hyena: bool, use_mamba: bool, dim: int, num_heads: int, window_size: Sequence[int],
                 shift_size: Sequence[int], mlp_ratio: float = 4.0, qkv_bias: bool = True, drop: float = 0.0,
                 attn_drop: float = 0.0, drop_path: float = 0.0, act_layer: str = "GELU",
                 norm_layer: type[LayerNorm] = nn.LayerNorm, use_checkpoint: bool = False) -> None:
        super().__init__()
        if drop_path != 0.0:
            raise NotImplementedError("drop_path is always 0 on the reference's reachable paths (backbone_swin.py:758)")
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = tuple(int(w) for w in window_size)
        self.shift_size = tuple(int(s) for s in shift_size)
        self.mlp_ratio = mlp_ratio
        self.use_checkpoint = use_checkpoint
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(use_hyena, use_mamba, dim, window_size=self.window_size, num_heads=num_heads,
                                    qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(hidden_size=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer, dropout_rate=drop,
                       dropout_mode="swin")

    def forward_part1(self, x, mask_matrix=None):
        """x: (B, [D,] H, W, C) channel-last. `mask_matrix` is accepted for signature parity and ignored: the
        kernel evaluates the shift-mask regions itself."""
        return self.attn.forward_grid(apply_layer_norm(self.norm1, x), self.shift_size)

    def forward_part2(self, x):
        return self.drop_path(self.mlp(apply_layer_norm(self.norm2, x)))

    def forward(self, x, mask_matrix=None):
        shortcut = x
        if self.use_checkpoint:
            x = checkpoint.checkpoint(self.forward_part1, x, mask_matrix, use_reentrant=False)
        else:
            x = self.forward_part1(x, mask_matrix)
        x = shortcut + self.drop_path(x)
        if self.use_checkpoint:
            x = x + checkpoint.checkpoint(self.forward_part2, x, use_reentrant=False)
        else:
            x = x + self.forward_part2(x)
        return x


class PatchMergingV2(nn.Module):
    """2x downsampling: concat the 2^k strided sub-grids, LayerNorm, Linear (reference :540-585). Plain torch
    (adjacent component)."""

    def __init__(self, dim: int, norm_layer: type[LayerNorm] = nn.LayerNorm, spatial_dims: int = 3) -> None:
        super().__init__()
        self.dim = dim
        k = 8 if spatial_dims == 3 else 4
        self.reduction = nn.Linear(k * dim, 2 * dim, bias=False)
        self.norm = norm_layer(k * dim)

    def forward(self, x):
        if x.dim() == 5:
            b, d, h, w, c = x.shape
            if (h % 2) or (w % 2) or (d % 2):
                x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2, 0, d % 2))
            x = torch.cat([x[:, i::2, j::2, k::2, :] for i, j, k in itertools.product(range(2), range(2), range(2))], -1)
        else:
            b, h, w, c = x.shape
            if (h % 2) or (w % 2):
                x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2))
            x = torch.cat([x[:, j::2, i::2, :] for i, j in itertools.product(range(2), range(2))], -1)
        return self.reduction(apply_layer_norm(self.norm, x))


MERGING_MODE = {"mergingv2": PatchMergingV2}


class BasicLayer(nn.Module):
    """One Swin stage: alternating W-MSA / SW-MSA blocks + patch merging (reference :631-733)."""

    def __init__(self, use_hyena: bool, use_mamba: bool, dim: int, depth: int, num_heads: int,
                 window_size: Sequence[int], drop_path: list, mlp_ratio: float = 4.0, qkv_bias: bool = False,
                 drop: float = 0.0, attn_drop: float = 0.0, norm_layer: type[LayerNorm] = nn.LayerNorm,
                 downsample: nn.Module | None = None, use_checkpoint: bool = False) -> None:
        super().__init__()
        if use_hyena or use_mamba:
            raise NotImplementedError(_ALT_MIXER_MSG)
        self.window_size = tuple(int(w) for w in window_size)
        self.shift_size = tuple(w // 2 for w in self.window_size)
        self.no_shift = tuple(0 for _ in self.window_size)
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(use_hyena=use_hyena, use_mamba=use_mamba, dim=dim, num_heads=num_heads,
                                 window_size=self.window_size,
                                 shift_size=self.no_shift if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio,
                                 qkv_bias=qkv_bias, drop=drop, attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer, use_checkpoint=use_checkpoint)
            for i in range(depth)])
        self.downsample = downsample
        if callable(self.downsample):
            self.downsample = downsample(dim=dim, norm_layer=norm_layer, spatial_dims=len(self.window_size))

    def forward_tokens(self, x):
        """Channel-last in, channel-last out: (B, [D,] H, W, C) -> (B, [D/2,] H/2, W/2, 2C)."""
        for blk in self.blocks:
            x = blk(x, None)
        if self.downsample is not None:
            x = self.downsample(x)
        return x

    def forward(self, x):
        """Reference signature: channel-first in / out."""
        perm_in = (0, 2, 3, 4, 1) if x.dim() == 5 else (0, 2, 3, 1)
        perm_out = (0, 4, 1, 2, 3) if x.dim() == 5 else (0, 3, 1, 2)
        return self.forward_tokens(x.permute(*perm_in)).permute(*perm_out)


class SwinTransformer_with_alt_ops(nn.Module):
    """Swin encoder returning [input, x0_out, x1_out, x2_out, x3_out, x4_out] (reference :736-911)."""

    def __init__(self, use_hyena: bool, use_mamba: bool, in_chans: int, embed_dim: int, window_size: Sequence[int],
                 patch_size: Sequence[int], depths: Sequence[int], num_heads: Sequence[int], mlp_ratio: float = 4.0,
                 qkv_bias: bool = True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, norm_layer: type[LayerNorm] = nn.LayerNorm, patch_norm: bool = False,
                 use_checkpoint: bool = False, spatial_dims: int = 3, downsample="mergingv2", use_v2=False) -> None:
        super().__init__()
        if use_hyena or use_mamba:
            raise NotImplementedError(_ALT_MIXER_MSG)
        if use_v2:
            raise NotImplementedError("use_v2 residual conv blocks are never enabled by custom_Swin (:117-125)")
        if drop_rate != 0.0 or attn_drop_rate != 0.0 or drop_path_rate != 0.0:
            raise NotImplementedError("all dropout rates are 0 on the reference's reachable paths (:756-758)")
        if len(depths) != 4 or len(num_heads) != 4:
            raise ValueError("the reference encoder has exactly four stages (layers1..layers4)")
        self.use_hyena = use_hyena
        self.use_mamba = use_mamba
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.window_size = tuple(int(w) for w in window_size)
        self.patch_size = tuple(int(p) for p in patch_size)
        self.spatial_dims = spatial_dims
        self.patch_embed = PatchEmbed(patch_size=self.patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None, spatial_dims=spatial_dims)
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.use_v2 = use_v2
        self.layers1 = nn.ModuleList()
        self.layers2 = nn.ModuleList()
        self.layers3 = nn.ModuleList()
        self.layers4 = nn.ModuleList()
        down = MERGING_MODE[downsample] if isinstance(downsample, str) else downsample
        for i_layer, holder in enumerate([self.layers1, self.layers2, self.layers3, self.layers4]):
            holder.append(BasicLayer(use_hyena=use_hyena, use_mamba=use_mamba, dim=int(embed_dim * 2 ** i_layer),
                                     depth=depths[i_layer], num_heads=num_heads[i_layer], window_size=self.window_size,
                                     drop_path=[0.0] * depths[i_layer], mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     drop=drop_rate, attn_drop=attn_drop_rate, norm_layer=norm_layer, downsample=down,
                                     use_checkpoint=use_checkpoint))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))

    def set_window_group(self, group):
        """Shard the windows of every attention block over `group` ("world" = the default group; None = off)."""
        for m in self.modules():
            if isinstance(m, WindowAttention):
                m.window_group = group

    @staticmethod
    def _channel_first(x_cl):
        return x_cl.permute(0, 4, 1, 2, 3) if x_cl.dim() == 5 else x_cl.permute(0, 3, 1, 2)

    def proj_out(self, x, normalize=False):
        """Reference signature (channel-first in/out): affine-free LayerNorm over channels."""
        if not normalize:
            return x
        perm_in = (0, 2, 3, 4, 1) if x.dim() == 5 else (0, 2, 3, 1)
        return self._channel_first(F.layer_norm(x.permute(*perm_in), [x.shape[1]]))

    def _out(self, x_cl, normalize):
        if not normalize:
            return self._channel_first(x_cl)
        if x_cl.shape[-1] % 4 == 0:   # affine-free; fp32 out like the reference's autocast layer_norm
            return self._channel_first(ops.layer_norm(x_cl, None, None, 1e-5, out_dtype=torch.float32))
        return self._channel_first(F.layer_norm(x_cl, [x_cl.shape[-1]]))

    def forward(self, x, normalize=True):
        if self.spatial_dims == 2:
            x = x.squeeze(2)
        hidden_states_out = [x]
        out_dtype = torch.bfloat16 if (torch.is_autocast_enabled("cuda") and
                                       torch.get_autocast_dtype("cuda") == torch.bfloat16) else torch.float32
        t = self.pos_drop(self.patch_embed.forward_tokens(x, out_dtype))     # channel-last token grid
        hidden_states_out.append(self._out(t, normalize))
        for stage in (self.layers1, self.layers2, self.layers3, self.layers4):
            t = stage[0].forward_tokens(t)
            hidden_states_out.append(self._out(t, normalize))
        if self.spatial_dims == 2:
            hidden_states_out = [h.unsqueeze(2) for h in hidden_states_out]
        return hidden_states_out
