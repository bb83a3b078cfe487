"""CPU: the C-ABI library builds, loads, and exports every symbol include/lcbi_b200.h declares; host logic of
the module mirrors (factories, state_dict layout, error behaviour) that needs no GPU."""
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lcbi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lcbi_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from long_context_biomedical_imaging_b200 import _lib, build

    build.build_library()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from long_context_biomedical_imaging_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 6
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lcbi_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.lcbi_version() == 100


def test_null_arguments_are_rejected_without_touching_the_gpu(lib):
    rc = lib.lcbi_dense_attn_fwd(None, None, None, None, None, 1, 1, 1, 1, 64, None, None, None, None, 1.0, None)
    assert rc == -1
    assert b"null pointer" in lib.lcbi_last_error()
    assert lib.lcbi_dense_attn_bwd_workspace_bytes(2, 3, 197, 64) >= 2 * 197 * 3 * 64 * 4 + 2 * 2 * 3 * 256 * 4
    # patch embedding: a workspace only for reduction lengths the tcgen05 path takes (K % 64 == 0); host-side sizes
    from long_context_biomedical_imaging_b200 import _lib

    cfg3 = lib.lcbi_patch_embed_workspace_bytes(16, 1, _lib.int3((8, 8, 8)), _lib.int3((12, 12, 12)), 768)
    m, k, n = 16 * 12 ** 3, 512, 768
    assert cfg3 >= 2 * m * (k + n) * 2                       # split patches + split dOut, bf16 (hi, lo) pairs
    assert lib.lcbi_patch_embed_workspace_bytes(1, 1, _lib.int3((1, 2, 2)), _lib.int3((1, 512, 512)), 768) == 0   # cfg5, K = 4
    assert lib.lcbi_patch_embed_workspace_bytes(1, 1, None, None, 768) == 0
    rc = lib.lcbi_patch_embed_fwd_ws(None, 0, None, None, None, None, 0, 1, 1, None, None, None, 8, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.lcbi_last_error()
    rc = lib.lcbi_patch_embed_bwd_ws(None, 0, None, None, 0, None, None, None, None, 1, 1, None, None, None, 8, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.lcbi_last_error()


def test_ops_refuse_cpu_tensors():
    from long_context_biomedical_imaging_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dense_attention_qkv(torch.randn(1, 8, 192), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.patch_embed(torch.randn(1, 1, 8, 8), torch.randn(4, 1, 2, 2), torch.randn(4), None, (4, 4))
    from long_context_biomedical_imaging_b200.blocks import apply_add_layer_norm, apply_layer_norm

    ln = torch.nn.LayerNorm(8)
    for call in (lambda: ops.layer_norm(torch.randn(4, 8), None, None),
                 lambda: ops.add_layer_norm(torch.randn(4, 256), torch.randn(4, 256), None, None),
                 lambda: ops.linear(torch.randn(4, 8), torch.randn(8, 8), torch.randn(8)),
                 lambda: ops.bias_grad(torch.randn(4, 8)),
                 lambda: apply_layer_norm(ln, torch.randn(4, 8)),
                 lambda: apply_add_layer_norm(ln, torch.randn(4, 8), torch.randn(4, 8))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def _vit_cfg(**kw):
    base = dict(size="base", patch_size=[8, 8, 8], use_hyena=False, use_mamba=False)
    base.update(kw)
    return types.SimpleNamespace(ViT=types.SimpleNamespace(**base), time=96, height=96, width=96, task_type="seg")


def test_custom_vit_factory_contract():
    from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT

    cfg = _vit_cfg()
    model, chans = custom_ViT(cfg, 1)
    assert chans == [768] * 13
    assert (cfg.ViT.hidden_size, cfg.ViT.mlp_dim, cfg.ViT.num_layers, cfg.ViT.num_heads) == (768, 3072, 12, 12)
    sd = model.state_dict()
    assert sd["patch_embedding.position_embeddings"].shape == (1, 1728, 768)
    assert sd["patch_embedding.patch_embeddings.weight"].shape == (768, 1, 8, 8, 8)
    assert sd["blocks.0.attn.qkv.weight"].shape == (2304, 768)
    assert "blocks.0.attn.qkv.bias" not in sd
    assert sd["blocks.11.mlp.linear2.bias"].shape == (768,)
    cfg = _vit_cfg(size="small")
    cfg.time, cfg.task_type = 1, "class"
    model, chans = custom_ViT(cfg, 3)
    assert chans == [384] * 13 and "cls_token" in model.state_dict()
    assert model.state_dict()["patch_embedding.patch_embeddings.weight"].shape == (384, 3, 8, 8)
    with pytest.raises(ValueError):
        custom_ViT(_vit_cfg(size="giant"), 1)
    with pytest.raises(NotImplementedError):
        custom_ViT(_vit_cfg(use_hyena=True), 1)
