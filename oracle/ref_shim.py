"""Import the UNMODIFIED reference encoders from /root/reference for oracle validation.

TEST INFRASTRUCTURE ONLY. Nothing in the product package imports this module.

The reference files (`model/models/backbone_vit.py:19-27`, `backbone_swin.py:19-27`) import MONAI 1.3.0,
timm and mamba_ssm, none of which are installed here. MONAI's PatchEmbeddingBlock / PatchEmbed / MLPBlock
are third-party code absent from /root/reference (pinned `monai==1.3.0`, reference `requirements.txt:5`);
their published semantics are restated below as stubs (strided conv + flatten/transpose + learnable
position embedding; trailing zero-pad + strided conv; linear1 -> GELU -> linear2), which is all the
reference's call sites (`backbone_vit.py:351-361,383`, `backbone_swin.py:800-806,885,433`) rely on.

This module only works where the reference's files exist: /root/reference in the build container, or the copy of
the four model files (`model/models/backbone_{vit,swin}.py`, `hyena.py`, `mamba.py`) that `__graft_entry__.build()`
stages under the git-ignored `baseline/_ref/` so that `bench.py --impl reference` can time the reference's OWN
attention code on the GPU box's host cores. It is used by `oracle/make_golden.py` to generate the committed fixtures
under tests/golden/ and by CPU tests that are skipped when no reference tree is present.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO_ROOT, "baseline", "_ref")
STAGED_FILES = ("backbone_vit.py", "backbone_swin.py", "hyena.py", "mamba.py")


def _pick_root():
    env = os.environ.get("LCBI_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", STAGED_ROOT):
        if os.path.isfile(os.path.join(root, "model", "models", "backbone_swin.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def stage_reference_files(src_root="/root/reference"):
    """Copies the four reference model files into baseline/_ref/ (git-ignored, travels to the GPU box). Returns the
    staged root, or None when the source tree is absent. Called by __graft_entry__.build() in the build container."""
    import shutil

    src_dir = os.path.join(src_root, "model", "models")
    if not all(os.path.isfile(os.path.join(src_dir, f)) for f in STAGED_FILES):
        return None
    dst_dir = os.path.join(STAGED_ROOT, "model", "models")
    os.makedirs(dst_dir, exist_ok=True)
    for f in STAGED_FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(dst_dir, f))
    return STAGED_ROOT


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "models", "backbone_swin.py"))


def _ensure_tuple_rep(x, n):
    if isinstance(x, (list, tuple)):
        if len(x) == n:
            return tuple(x)
        if len(x) == 1:
            return tuple(x) * n
        raise ValueError(f"sequence must have length {n}, got {len(x)}")
    return (x,) * n


class _PatchEmbeddingBlock(nn.Module):
    """MONAI 1.3.0 `PatchEmbeddingBlock` with proj_type='conv' (restated from its published behaviour)."""

    def __init__(self, in_channels, img_size, patch_size, hidden_size, num_heads, proj_type="conv",
                 pos_embed_type="learnable", dropout_rate=0.0, spatial_dims=3, **_):
        super().__init__()
        if proj_type != "conv":
            raise NotImplementedError("shim only restates proj_type='conv'")
        img_size = _ensure_tuple_rep(img_size, spatial_dims)
        patch_size = _ensure_tuple_rep(patch_size, spatial_dims)
        conv = nn.Conv3d if spatial_dims == 3 else nn.Conv2d
        self.patch_embeddings = conv(in_channels, hidden_size, kernel_size=patch_size, stride=patch_size)
        self.n_patches = 1
        for i, p in zip(img_size, patch_size):
            self.n_patches *= i // p
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden_size))
        self.dropout = nn.Dropout(dropout_rate)
        if pos_embed_type == "learnable":
            nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        elif pos_embed_type != "none":
            raise NotImplementedError(pos_embed_type)

    def forward(self, x):
        x = self.patch_embeddings(x)
        x = x.flatten(2).transpose(-1, -2)
        return self.dropout(x + self.position_embeddings)


class _PatchEmbed(nn.Module):
    """MONAI 1.3.0 `PatchEmbed` (Swin): trailing zero-pad to a patch multiple, strided conv, channel-first."""

    def __init__(self, patch_size=2, in_chans=1, embed_dim=48, norm_layer=None, spatial_dims=3):
        super().__init__()
        patch_size = _ensure_tuple_rep(patch_size, spatial_dims)
        self.patch_size = patch_size
        self.embed_dim = embed_dim
        conv = nn.Conv3d if spatial_dims == 3 else nn.Conv2d
        self.proj = conv(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        sp = x.shape[2:]
        pads = []
        for size, p in zip(reversed(sp), reversed(self.patch_size)):
            pads += [0, (p - size % p) % p]
        if any(pads):
            x = F.pad(x, pads)
        x = self.proj(x)
        if self.norm is not None:
            shp = x.shape
            x = x.flatten(2).transpose(1, 2)
            x = self.norm(x)
            x = x.transpose(1, 2).view(shp)
        return x


class _MLPBlock(nn.Module):
    def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, act="GELU", dropout_mode="vit"):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        return self.drop2(self.linear2(self.drop1(self.fn(self.linear1(x)))))


class _DropPath(nn.Module):
    def __init__(self, p=0.0):
        super().__init__()
        if p != 0.0:
            raise NotImplementedError("drop_path is always 0 on the reference's reachable paths")

    def forward(self, x):
        return x


def _deprecated_arg(*_a, **_k):
    def deco(fn):
        return fn
    return deco


def _optional_import(module, name=""):
    try:
        m = importlib.import_module(module)
        return (getattr(m, name) if name else m), True
    except Exception:  # pragma: no cover
        return None, False


def _look_up_option(key, options):
    if isinstance(options, dict):
        return options[key]
    if key in options:
        return key
    raise ValueError(key)


def _unavailable(*_a, **_k):
    raise NotImplementedError("out of scope for the attention hot path")


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    placeholder = type("Placeholder", (nn.Module,), {"__init__": _unavailable})
    mod("monai")
    mod("monai.networks")
    mod("monai.networks.blocks", PatchEmbed=_PatchEmbed, MLPBlock=_MLPBlock, UnetOutBlock=placeholder,
        UnetrBasicBlock=placeholder, UnetrUpBlock=placeholder, UnetrPrUpBlock=placeholder)
    mod("monai.networks.blocks.patchembedding", PatchEmbeddingBlock=_PatchEmbeddingBlock)
    mod("monai.networks.blocks.mlp", MLPBlock=_MLPBlock)
    mod("monai.networks.layers", DropPath=_DropPath,
        trunc_normal_=lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: nn.init.trunc_normal_(t, mean, std, a, b))
    mod("monai.utils", deprecated_arg=_deprecated_arg, optional_import=_optional_import,
        ensure_tuple_rep=_ensure_tuple_rep, look_up_option=_look_up_option)
    mod("monai.utils.deprecate_utils", deprecated_arg=_deprecated_arg)
    mod("timm")
    mod("timm.models")
    mod("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, DropPath=_DropPath, LayerNorm2d=placeholder)
    mod("timm.models.vision_transformer", Mlp=placeholder, PatchEmbed=placeholder)
    mod("timm.models._builder", resolve_pretrained_cfg=_unavailable)
    mod("timm.models._manipulate", named_apply=_unavailable)
    mod("mamba_ssm")
    mod("mamba_ssm.ops")
    mod("mamba_ssm.ops.selective_scan_interface", selective_scan_fn=_unavailable)


_loaded = {}


def load_reference():
    """Returns (backbone_vit, backbone_swin) modules of the unmodified reference."""
    if _loaded:
        return _loaded["vit"], _loaded["swin"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stubs()
    models_dir = os.path.join(REFERENCE_ROOT, "model", "models")
    if models_dir not in sys.path:
        sys.path.insert(0, models_dir)
    # names are suffixed so they cannot shadow the product package's modules of the same name
    import importlib.util

    def _imp(alias, fname):
        spec = importlib.util.spec_from_file_location(alias, os.path.join(models_dir, fname))
        m = importlib.util.module_from_spec(spec)
        sys.modules[alias] = m
        spec.loader.exec_module(m)
        return m

    _loaded["vit"] = _imp("lcbi_reference_backbone_vit", "backbone_vit.py")
    _loaded["swin"] = _imp("lcbi_reference_backbone_swin", "backbone_swin.py")
    return _loaded["vit"], _loaded["swin"]
