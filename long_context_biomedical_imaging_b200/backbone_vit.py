"""ViT encoder: host-side mirror of the reference interface, B200-native token mixing.

Mirrors /root/reference/model/models/backbone_vit.py so that `EncoderDecoderModel` (reference
model/model_base.py:44-47) can use it unchanged:

  custom_ViT(config, input_feature_channels) -> (model, [hidden]*13)        reference :45-116
  SABlock / TransformerBlock / ViT_with_alt_ops                              reference :120-397

Same constructor arguments, same config mutation, same state_dict keys and shapes
(patch_embedding.position_embeddings, patch_embedding.patch_embeddings.{weight,bias}, [cls_token],
blocks.{i}.{norm1,norm2}.*, blocks.{i}.attn.qkv.weight, blocks.{i}.attn.out_proj.*,
blocks.{i}.mlp.linear{1,2}.*, norm.*), same 14-element hidden-state list. What changes is HOW the hot
path runs: the attention core is one fused tcgen05 kernel reading q/k/v in place from the qkv Linear
output (no rearrange copies, no N x N matrix), and the patch embedding is a fused CUDA kernel
(conv-as-GEMM + bias + position embedding). LayerNorm / MLP / qkv / out_proj stay torch ops.

Sequence parallelism (new relative to the reference, which has DDP only: trainer/trainer_base.py:94-98): pass
`sequence_group=<process group>` (or set `config.ViT.sequence_group`) and every rank runs the encoder on its
contiguous slice of the token sequence — its rows of the image go through the patch embedding with its rows of
`position_embeddings`, LayerNorm / MLP / qkv / out_proj are token-local, and the attention core becomes ring attention
over the group (`ring.py`). Hidden states are returned as local slices, or gathered with `gather_outputs=True`.

use_hyena / use_mamba select the reference's alternative mixers, which are outside this package's scope:
requesting them raises NotImplementedError naming the reference module to use instead.
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
import torch.nn as nn

from . import ops, ring
from .blocks import MLPBlock, PatchEmbeddingBlock, apply_add_layer_norm, apply_layer_norm

_ALT_MIXER_MSG = ("use_hyena/use_mamba route to the reference's HyenaOperator / MambaVisionMixer "
                  "(model/models/hyena.py, mamba.py), which are outside the B200 attention hot path; build the "
                  "reference encoder for those configurations")


def custom_ViT(config, input_feature_channels):
    """Factory with the reference's contract (backbone_vit.py:45-116): resolves the size preset, writes the
    resolved sizes back into `config.ViT`, returns (model, [hidden_size] * 13)."""
    size = config.ViT.size
    if size == "small":
        hidden_size, mlp_dim, num_layers, num_heads = 384, 1536, 12, 6
    elif size == "base":
        hidden_size, mlp_dim, num_layers, num_heads = 768, 3072, 12, 12
    elif size == "custom":
        hidden_size, mlp_dim = config.ViT.hidden_size, config.ViT.mlp_dim
        num_layers, num_heads = config.ViT.num_layers, config.ViT.num_heads
    else:
        raise ValueError(f"Unknown model size {config.ViT.size} specified in config.")
    config.ViT.hidden_size = hidden_size
    config.ViT.mlp_dim = mlp_dim
    config.ViT.num_layers = num_layers
    config.ViT.num_heads = num_heads

    if config.time == 1:
        spatial_dims = 2
        input_size = [config.height, config.width]
        patch = config.ViT.patch_size[1:] if len(config.ViT.patch_size) == 3 else config.ViT.patch_size
    else:
        spatial_dims = 3
        input_size = [config.time, config.height, config.width]
        patch = config.ViT.patch_size

    model = ViT_with_alt_ops(use_hyena=config.ViT.use_hyena, use_mamba=config.ViT.use_mamba,
                             in_channels=input_feature_channels, img_size=input_size, patch_size=patch,
                             hidden_size=hidden_size, mlp_dim=mlp_dim, num_layers=num_layers, num_heads=num_heads,
                             dropout_rate=0.0, spatial_dims=spatial_dims, classification=config.task_type == "class",
                             sequence_group=getattr(config.ViT, "sequence_group", None),
                             gather_outputs=getattr(config.ViT, "gather_outputs", True))
    return model, [hidden_size] * 13


class SABlock(nn.Module):
    """Global multi-head self-attention (reference backbone_vit.py:120-211).

    forward(x: (B,N,C)) -> (B,N,C): qkv Linear (no bias) -> fused flash attention kernel -> out_proj."""

    def __init__(self, use_hyena: bool, use_mamba: bool, hidden_size: int, num_heads: int, dropout_rate: float = 0.0,
                 qkv_bias: bool = False, save_attn: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden size should be divisible by num_heads.")
        if use_hyena or use_mamba:
            raise NotImplementedError(_ALT_MIXER_MSG)
        if save_attn:
            raise NotImplementedError("save_attn materialises the N x N attention matrix, which the fused kernel "
                                      "never forms (custom_ViT never enables it: reference backbone_vit.py:101-112)")
        if dropout_rate != 0.0:
            raise NotImplementedError("dropout inside attention is always 0 on the reference's reachable paths "
                                      "(backbone_vit.py:110)")
        self.num_heads = num_heads
        self.use_hyena = use_hyena
        self.use_mamba = use_mamba
        self.head_dim = hidden_size // num_heads
        self.scale = self.head_dim ** -0.5
        self.save_attn = save_attn
        self.att_mat = torch.Tensor()
        self.qkv = nn.Linear(hidden_size, hidden_size * 3, bias=qkv_bias)
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.ring_comm = None      # set by ViT_with_alt_ops when the sequence is sharded over a process group

    def forward(self, x):
        qkv = self.qkv(x)                                        # (B, N, 3C), feature = s*C + h*d + j
        if self.ring_comm is not None:                           # x is this rank's token slice: ring K/V exchange
            o = ring.ring_attention_qkv(qkv, self.num_heads, self.ring_comm, self.scale)
        else:
            o = ops.dense_attention_qkv(qkv, self.num_heads, self.scale)   # (B, N, C), feature = h*d + j
        return ops.linear(o, self.out_proj.weight, self.out_proj.bias)   # cuBLAS GEMMs, fused-kernel bias gradient


class TransformerBlock(nn.Module):
    """Pre-norm residual block (reference backbone_vit.py:213-263)."""

    def __init__(self, use_hyena: bool, use_mamba: bool, hidden_size: int, mlp_dim: int, num_heads: int,
                 dropout_rate: float = 0.0, qkv_bias: bool = False, save_attn: bool = False) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.mlp = MLPBlock(hidden_size, mlp_dim, dropout_rate)
        self.norm1 = nn.LayerNorm(hidden_size)
        self.use_hyena = use_hyena
        self.use_mamba = use_mamba
        self.attn = SABlock(use_hyena, use_mamba, hidden_size, num_heads, dropout_rate, qkv_bias, save_attn)
        self.norm2 = nn.LayerNorm(hidden_size)

    def forward(self, x):
        x, y = apply_add_layer_norm(self.norm2, x, self.attn(apply_layer_norm(self.norm1, x)))   # x + attn, norm2 of it
        return x + self.mlp(y)


class ViT_with_alt_ops(nn.Module):
    """ViT encoder returning [input, block_1 .. block_L, final_norm] (reference backbone_vit.py:265-397)."""

    def __init__(self, use_hyena: bool, use_mamba: bool, in_channels: int, img_size: Sequence[int] | int,
                 patch_size: Sequence[int] | int, hidden_size: int = 768, mlp_dim: int = 3072, num_layers: int = 12,
                 num_heads: int = 12, pos_embed: str = "conv", proj_type: str = "conv",
                 pos_embed_type: str = "learnable", classification: bool = False, num_classes: int = 2,
                 dropout_rate: float = 0.0, spatial_dims: int = 3, post_activation="Tanh", qkv_bias: bool = False,
                 save_attn: bool = False, sequence_group=None, gather_outputs: bool = True) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        if use_hyena or use_mamba:
            raise NotImplementedError(_ALT_MIXER_MSG)
        self.classification = classification
        self.spatial_dims = spatial_dims
        self.patch_embedding = PatchEmbeddingBlock(in_channels=in_channels, img_size=img_size, patch_size=patch_size,
                                                   hidden_size=hidden_size, num_heads=num_heads, proj_type=proj_type,
                                                   pos_embed_type=pos_embed_type, dropout_rate=dropout_rate,
                                                   spatial_dims=spatial_dims)
        self.blocks = nn.ModuleList([
            TransformerBlock(use_hyena, use_mamba, hidden_size, mlp_dim, num_heads, dropout_rate, qkv_bias, save_attn)
            for _ in range(num_layers)])
        self.norm = nn.LayerNorm(hidden_size)
        if self.classification:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden_size))
        self.sequence_group = None
        self.gather_outputs = gather_outputs
        self._ring_comm = None
        if sequence_group is not None:
            self.set_sequence_group(sequence_group)

    def set_sequence_group(self, group, gather_outputs=None):
        """Shard the token sequence over `group` (None: back to one device per replica). The communicator is created
        at the first forward (torch.distributed must be initialised by then)."""
        if group is not None and self.classification:
            raise NotImplementedError("sequence parallelism with a cls token (task_type 'class') would make the shards "
                                      "uneven; the long-sequence configurations are dense-prediction encoders")
        self.sequence_group = group
        self._ring_comm = None
        if gather_outputs is not None:
            self.gather_outputs = gather_outputs
        if group is None:
            for blk in self.blocks:
                blk.attn.ring_comm = None

    def _sequence_comm(self):
        import torch.distributed as dist

        if self.sequence_group is None or not (dist.is_available() and dist.is_initialized()):
            return None
        group = None if self.sequence_group == "world" else self.sequence_group
        if dist.get_world_size(group) == 1:
            return None
        if self._ring_comm is None:
            self._ring_comm = ring.RingComm(group)
            for blk in self.blocks:
                blk.attn.ring_comm = self._ring_comm
        return self._ring_comm

    def forward(self, x):
        if self.spatial_dims == 2:
            x = x.squeeze(2)
        hidden_states_out = [x]
        comm = self._sequence_comm()
        if comm is not None:
            # rank r owns a contiguous block of rows of the patch grid = a slab of the image along its first spatial
            # axis (tokens are raster-ordered, reference backbone_vit.py:383): embed only that slab, with its rows of
            # position_embeddings
            g0 = self.patch_embedding.grid[0]
            if g0 % comm.world != 0:
                raise ValueError(f"sequence parallelism needs the first patch-grid axis ({g0}) divisible by the group "
                                 f"size ({comm.world})")
            rows = g0 // comm.world * self.patch_embedding.patch_size[0]
            x = self.patch_embedding(x.narrow(2, comm.rank * rows, rows), shard=(comm.rank, comm.world))
        else:
            x = self.patch_embedding(x)
        if hasattr(self, "cls_token"):
            x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype), x), dim=1)
        # The blocks' arithmetic (reference :260-263) with every residual add folded into the LayerNorm that reads the
        # sum: a block's MLP output stays pending until the next block's norm1 (or the final norm) adds it, and that
        # fused pass also produces the block's hidden state. blk(x) itself remains the per-block seam.
        pending = None
        for blk in self.blocks:
            x, y = apply_add_layer_norm(blk.norm1, x, pending)
            if pending is not None:
                hidden_states_out.append(x)
            x, y = apply_add_layer_norm(blk.norm2, x, blk.attn(y))
            pending = blk.mlp(y)
        x, y = apply_add_layer_norm(self.norm, x, pending, out_dtype=x.dtype)   # final norm: a hidden state, keeps x's dtype
        if pending is not None:
            hidden_states_out.append(x)
        hidden_states_out.append(y)
        if comm is not None and self.gather_outputs:      # the decoders are replicated: hand them whole sequences
            hidden_states_out = hidden_states_out[:1] + [ring.gather_sequence(h, comm.group) for h in hidden_states_out[1:]]
        return hidden_states_out
