"""Building blocks the reference takes from MONAI 1.3.0 (not vendored in the reference; requirements.txt:5),
re-implemented with identical parameter names / shapes so state_dicts interchange.

  PatchEmbeddingBlock  (reference call site backbone_vit.py:351-361,383)   keys: patch_embeddings.{weight,bias},
                                                                                 position_embeddings
  PatchEmbed           (reference call site backbone_swin.py:800-806,885)  keys: proj.{weight,bias}
  MLPBlock             (backbone_vit.py:249, backbone_swin.py:433)         keys: linear1.*, linear2.*

The two patch-embedding modules run the fused CUDA kernel (ops.patch_embed); MLPBlock is plain torch
(adjacent component). apply_layer_norm routes the blocks' nn.LayerNorm modules through the
LayerNorm kernels (the normalisation in front of the qkv projection).
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
import torch.nn as nn

from . import ops


def ensure_tuple_rep(x, n):
    if isinstance(x, (list, tuple)):
        if len(x) == n:
            return tuple(int(v) for v in x)
        if len(x) == 1:
            return tuple(int(x[0]) for _ in range(n))
        raise ValueError(f"sequence must have length {n}, got {len(x)}")
    return tuple(int(x) for _ in range(n))


def apply_layer_norm(norm, x, out_dtype=None):
    """`norm(x)` for the encoder blocks' nn.LayerNorm modules through the fused kernels (ops.layer_norm: fp32
    statistics, output written in the dtype the next Linear consumes). The module keeps its parameters (state_dict
    keys norm1 / norm2 / norm as in the reference: backbone_vit.py:249-258, backbone_swin.py:419-433); any other
    norm_layer, or a channel count the kernel does not take, runs as the module itself."""
    ops._require_cuda(x)      # like every operator of this package: no CPU path
    if (type(norm) is nn.LayerNorm and len(norm.normalized_shape) == 1 and
            norm.normalized_shape[0] == x.shape[-1] and x.shape[-1] % 4 == 0):
        return ops.layer_norm(x, norm.weight, norm.bias, norm.eps, out_dtype)
    y = norm(x)               # another norm_layer class or an odd channel count: the (CUDA) module itself
    return y if out_dtype is None else y.to(out_dtype)


def apply_add_layer_norm(norm, x, delta, out_dtype=None):
    """(x + delta, norm(x + delta)): the residual add of a pre-norm block together with the LayerNorm that consumes the
    sum (reference backbone_vit.py:261-262). `delta=None` means there is nothing to add yet (first block)."""
    if delta is None:
        return x, apply_layer_norm(norm, x, out_dtype)
    ops._require_cuda(x, delta)
    if (type(norm) is nn.LayerNorm and len(norm.normalized_shape) == 1 and
            norm.normalized_shape[0] == x.shape[-1]):
        return ops.add_layer_norm(x, delta, norm.weight, norm.bias, norm.eps, out_dtype)
    x = x + delta
    return x, apply_layer_norm(norm, x, out_dtype)


class MLPBlock(nn.Module):
    """linear1 -> GELU -> linear2 (all dropouts are 0 on the reference's reachable paths)."""

    def __init__(self, hidden_size: int, mlp_dim: int, dropout_rate: float = 0.0, act: str = "GELU",
                 dropout_mode: str = "vit") -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if act != "GELU":
            raise ValueError("only GELU is used by the reference encoders")
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        h = self.fn(ops.linear(x, self.linear1.weight, self.linear1.bias))
        return self.drop2(ops.linear(self.drop1(h), self.linear2.weight, self.linear2.bias))


def _conv_like_init_(weight, bias):
    """torch.nn.ConvNd default initialisation (what MONAI's conv projection gets)."""
    nn.init.kaiming_uniform_(weight, a=5 ** 0.5)
    fan_in = weight[0].numel()
    bound = 1 / fan_in ** 0.5 if fan_in > 0 else 0
    nn.init.uniform_(bias, -bound, bound)


class _ConvParams(nn.Module):
    """Holds `weight (out, in, *kernel)` / `bias (out)` under the same names a torch ConvNd would."""

    def __init__(self, in_channels, out_channels, kernel):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *kernel))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.kernel_size = tuple(kernel)
        self.stride = tuple(kernel)
        _conv_like_init_(self.weight, self.bias)


class PatchEmbeddingBlock(nn.Module):
    """ViT patch embedding: strided-conv projection + learnable position embedding -> (B, N, hidden) fp32."""

    def __init__(self, in_channels: int, img_size: Sequence[int] | int, patch_size: Sequence[int] | int,
                 hidden_size: int, num_heads: int, proj_type: str = "conv", pos_embed_type: str = "learnable",
                 dropout_rate: float = 0.0, spatial_dims: int = 3) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden size should be divisible by num_heads.")
        if proj_type != "conv":
            raise ValueError("only proj_type='conv' is used by the reference encoder (backbone_vit.py:288,357)")
        img_size = ensure_tuple_rep(img_size, spatial_dims)
        patch_size = ensure_tuple_rep(patch_size, spatial_dims)
        for m, p in zip(img_size, patch_size):
            if m < p:
                raise ValueError("patch_size should be smaller than img_size.")
        self.img_size, self.patch_size = img_size, patch_size
        self.grid = tuple(m // p for m, p in zip(img_size, patch_size))
        self.n_patches = 1
        for g in self.grid:
            self.n_patches *= g
        self.patch_embeddings = _ConvParams(in_channels, hidden_size, patch_size)
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden_size))
        self.dropout = nn.Dropout(dropout_rate)
        if pos_embed_type == "learnable":
            nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        elif pos_embed_type != "none":
            raise ValueError(f"pos_embed_type {pos_embed_type} not supported")

    def forward(self, x, shard=None):
        """shard = (rank, world): x is this rank's slab of the image along the first spatial axis and the tokens it
        yields are rows [rank*N/world, (rank+1)*N/world) of the sequence (sequence-parallel encoders)."""
        grid = tuple(s // p for s, p in zip(x.shape[2:], self.patch_size))   # conv floors
        n = 1
        for g in grid:
            n *= g
        pos = self.position_embeddings
        if shard is not None:
            rank, world = shard
            if n * world != self.n_patches:
                raise RuntimeError(f"shard yields {n} patches, expected {self.n_patches} / {world}")
            pos = pos[:, rank * n:(rank + 1) * n]
        elif n != self.n_patches:
            raise RuntimeError(f"input yields {n} patches, position embedding has {self.n_patches}")
        return ops.patch_embed(x, self.patch_embeddings.weight, self.patch_embeddings.bias, pos, grid, torch.float32)


class PatchEmbed(nn.Module):
    """Swin patch embedding: trailing zero-pad to a patch multiple + strided-conv projection.

    `forward` returns what MONAI's PatchEmbed returns at the reference's `self.patch_embed(x)` seam
    (backbone_swin.py:885): the CHANNEL-FIRST grid (B, C, *grid) - as a permuted view of the channel-last tokens the
    kernel writes, so nothing is copied. The encoder itself calls `forward_tokens` and stays channel-last."""

    def __init__(self, patch_size: Sequence[int] | int = 2, in_chans: int = 1, embed_dim: int = 48, norm_layer=None,
                 spatial_dims: int = 3) -> None:
        super().__init__()
        if spatial_dims not in (2, 3):
            raise ValueError("spatial dimension should be 2 or 3.")
        if norm_layer is not None:
            raise NotImplementedError("patch_norm is always False in the reference (backbone_swin.py:760,804)")
        self.patch_size = ensure_tuple_rep(patch_size, spatial_dims)
        self.embed_dim = embed_dim
        self.proj = _ConvParams(in_chans, embed_dim, self.patch_size)
        self.norm = None

    def forward_tokens(self, x, out_dtype=torch.float32):
        """Channel-last token grid (B, *grid, C)."""
        grid = tuple(-(-s // p) for s, p in zip(x.shape[2:], self.patch_size))   # ceil: trailing zero pad
        tokens = ops.patch_embed(x, self.proj.weight, self.proj.bias, None, grid, out_dtype)
        return tokens.view(x.shape[0], *grid, self.embed_dim)

    def forward(self, x, out_dtype=torch.float32):
        t = self.forward_tokens(x, out_dtype)
        return t.permute(0, t.dim() - 1, *range(1, t.dim() - 1))
