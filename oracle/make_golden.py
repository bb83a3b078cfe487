"""Generates tests/golden/*.npz by running the UNMODIFIED reference (through oracle/ref_shim.py).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The fixtures are small, committed, and are what the GPU box checks against (the reference tree does
not exist there). TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
import os
import types

import numpy as np
import torch
import torch.nn.functional as F

from .attention_oracle import fill_parameters_
from .ref_shim import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (grid, ctor window, ctor shift) — cfg2 / cfg4 stage geometries, the shipped scripts' window 4 / 8,
# clamped and mixed-clamp cases.
MAP_CASES = [
    ((128, 128), (7, 7), (3, 3)), ((64, 64), (7, 7), (3, 3)), ((32, 32), (7, 7), (3, 3)), ((16, 16), (7, 7), (3, 3)),
    ((128, 128), (7, 7), (0, 0)), ((10, 9), (7, 7), (3, 3)), ((4, 4), (7, 7), (3, 3)), ((7, 7), (7, 7), (3, 3)),
    ((16, 16), (4, 4), (2, 2)), ((24, 24), (8, 8), (4, 4)), ((5, 20), (7, 7), (3, 3)),
    ((64, 64, 64), (7, 7, 7), (3, 3, 3)), ((32, 32, 32), (7, 7, 7), (3, 3, 3)), ((16, 16, 16), (7, 7, 7), (3, 3, 3)),
    ((8, 8, 8), (7, 7, 7), (3, 3, 3)), ((8, 8, 8), (7, 7, 7), (0, 0, 0)), ((4, 4, 4), (7, 7, 7), (3, 3, 3)),
    ((4, 16, 16), (7, 7, 7), (3, 3, 3)), ((5, 7, 4), (3, 3, 3), (1, 1, 1)), ((2, 8, 8), (4, 4, 4), (2, 2, 2)),
    ((16, 16, 16), (4, 4, 4), (2, 2, 2)), ((9, 6, 11), (4, 4, 4), (2, 2, 2)),
]
FULL_LIMIT = 20000  # store the full array when it has at most this many elements, else its sha256


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _store(d, key, arr):
    arr = np.ascontiguousarray(arr)
    d[key + "/sha256"] = np.array(_sha(arr))
    d[key + "/shape"] = np.array(arr.shape, dtype=np.int64)
    if arr.size <= FULL_LIMIT:
        d[key + "/full"] = arr


def reference_maps(swin, grid, window, shift):
    """Drive the reference's own pad/roll/partition/compute_mask code on an arange tensor."""
    K = len(grid)
    win, sh = swin.get_window_size(grid, window, shift)
    n_tok = int(np.prod(grid))
    x = torch.arange(1, n_tok + 1, dtype=torch.float64).reshape(1, *grid, 1)  # 0 is reserved for padding
    pads = [0, 0]
    for g, w in zip(reversed(grid), reversed(win)):
        pads += [0, (w - g % w) % w]
    xp = F.pad(x, pads)
    if any(s > 0 for s in sh):
        xp = torch.roll(xp, shifts=tuple(-s for s in sh), dims=tuple(range(1, K + 1)))
    gmap = swin.window_partition(xp, win).squeeze(-1).to(torch.int64).numpy() - 1  # pad -> -1
    pg = [int(np.ceil(g / w)) * w for g, w in zip(grid, win)]
    mask = swin.compute_mask(pg, win, sh, "cpu").numpy().astype(np.float32)
    return gmap, mask, win, sh


def make_window_maps(swin):
    d = {}
    names = []
    for ci, (grid, window, shift) in enumerate(MAP_CASES):
        gmap, mask, win, sh = reference_maps(swin, grid, window, shift)
        key = f"case{ci}"
        names.append(key)
        d[key + "/grid"] = np.array(grid)
        d[key + "/window"] = np.array(window)
        d[key + "/shift"] = np.array(shift)
        d[key + "/win_used"] = np.array(win)
        d[key + "/shift_used"] = np.array(sh)
        _store(d, key + "/gather_map", gmap.astype(np.int64))
        _store(d, key + "/mask", mask)
    for window in [(7, 7), (4, 4), (8, 8), (7, 7, 7), (4, 4, 4), (3, 3, 3), (2, 4, 4)]:
        attn = swin.WindowAttention(False, False, dim=4, num_heads=1, window_size=window, qkv_bias=True)
        _store(d, "relidx/" + "x".join(map(str, window)), attn.relative_position_index.numpy().astype(np.int64))
    d["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "window_maps.npz"), **d)


def _grads(out, leaves):
    g = torch.Generator().manual_seed(1234)
    dout = torch.randn(out.shape, generator=g)
    grads = torch.autograd.grad(out, leaves, dout)
    return dout, grads


def make_sablock(vit):
    d = {}
    for name, (B, N, C, H) in {"a": (2, 50, 128, 2), "b": (1, 197, 192, 3)}.items():
        torch.manual_seed(7)
        blk = vit.SABlock(False, False, C, H)
        fill_parameters_(blk, 11)
        x = torch.randn(B, N, C, requires_grad=True)
        y = blk(x)
        leaves = [x, blk.qkv.weight, blk.out_proj.weight, blk.out_proj.bias]
        dout, gr = _grads(y, leaves)
        for k, v in dict(x=x, w_qkv=blk.qkv.weight, w_out=blk.out_proj.weight, b_out=blk.out_proj.bias, y=y,
                         dout=dout, dx=gr[0], dw_qkv=gr[1], dw_out=gr[2], db_out=gr[3]).items():
            d[f"{name}/{k}"] = v.detach().numpy()
        d[f"{name}/heads"] = np.array(H)
    np.savez_compressed(os.path.join(OUT, "sablock.npz"), **d)


SWIN_BLOCK_CASES = {
    # name: (B, grid, C, heads, ctor window, ctor shift)
    "2d_shift": (2, (10, 9), 32, 2, (7, 7), (3, 3)),
    "2d_noshift": (2, (10, 9), 32, 2, (7, 7), (0, 0)),
    "2d_w4": (1, (8, 8), 64, 2, (4, 4), (2, 2)),
    "3d_shift": (1, (5, 7, 4), 32, 2, (3, 3, 3), (1, 1, 1)),
    "3d_clamp": (2, (2, 8, 8), 48, 3, (4, 4, 4), (2, 2, 2)),
    "3d_allclamp": (1, (3, 3, 3), 32, 2, (7, 7, 7), (3, 3, 3)),
}


def make_swin_part1(swin):
    d = {}
    for name, (B, grid, C, H, window, shift) in SWIN_BLOCK_CASES.items():
        blk = swin.SwinTransformerBlock(False, False, dim=C, num_heads=H, window_size=window, shift_size=shift)
        fill_parameters_(blk, 5)
        g = torch.Generator().manual_seed(3)
        x = torch.randn(B, *grid, C, generator=g, requires_grad=True)
        win, sh = swin.get_window_size(grid, window, shift)
        pg = [int(np.ceil(a / w)) * w for a, w in zip(grid, win)]
        mask = swin.compute_mask(pg, win, sh, "cpu")
        y = blk.forward_part1(x, mask)
        a = blk.attn
        leaves = [x, blk.norm1.weight, blk.norm1.bias, a.qkv.weight, a.qkv.bias, a.relative_position_bias_table,
                  a.proj.weight, a.proj.bias]
        dout, gr = _grads(y, leaves)
        vals = dict(x=x, norm_w=blk.norm1.weight, norm_b=blk.norm1.bias, w_qkv=a.qkv.weight, b_qkv=a.qkv.bias,
                    table=a.relative_position_bias_table, w_proj=a.proj.weight, b_proj=a.proj.bias, y=y, dout=dout,
                    dx=gr[0], dnorm_w=gr[1], dnorm_b=gr[2], dw_qkv=gr[3], db_qkv=gr[4], dtable=gr[5], dw_proj=gr[6],
                    db_proj=gr[7])
        for k, v in vals.items():
            d[f"{name}/{k}"] = v.detach().numpy()
        d[f"{name}/meta"] = np.array([B, C, H, len(grid)] + list(grid) + list(window) + list(shift))
    d["cases"] = np.array(list(SWIN_BLOCK_CASES))
    np.savez_compressed(os.path.join(OUT, "swin_part1.npz"), **d)


def make_patch_embed():
    from .ref_shim import _PatchEmbed, _PatchEmbeddingBlock  # MONAI semantics as the reference calls them

    d = {}
    cases = {"vit2d": (2, 1, (32, 48), (16, 16), 24), "vit3d": (1, 2, (8, 16, 16), (4, 8, 8), 32),
             "vit2d_p2": (1, 1, (16, 16), (2, 2), 64)}
    for name, (B, Cin, img, patch, hid) in cases.items():
        m = _PatchEmbeddingBlock(Cin, img, patch, hid, 1, spatial_dims=len(img))
        fill_parameters_(m, 21, std=0.2)
        g = torch.Generator().manual_seed(2)
        x = torch.randn(B, Cin, *img, generator=g, requires_grad=True)
        y = m(x)
        leaves = [x, m.patch_embeddings.weight, m.patch_embeddings.bias, m.position_embeddings]
        dout, gr = _grads(y, leaves)
        for k, v in dict(x=x, w=m.patch_embeddings.weight, b=m.patch_embeddings.bias, pos=m.position_embeddings, y=y,
                         dout=dout, dx=gr[0], dw=gr[1], db=gr[2], dpos=gr[3]).items():
            d[f"{name}/{k}"] = v.detach().numpy()
    cases = {"swin2d": (2, 1, (18, 21), (4, 4), 24), "swin3d": (1, 1, (9, 10, 11), (2, 2, 2), 48)}
    for name, (B, Cin, img, patch, emb) in cases.items():
        m = _PatchEmbed(patch, Cin, emb, None, spatial_dims=len(img))
        fill_parameters_(m, 22, std=0.2)
        g = torch.Generator().manual_seed(2)
        x = torch.randn(B, Cin, *img, generator=g, requires_grad=True)
        y = m(x)
        dout, gr = _grads(y, [x, m.proj.weight, m.proj.bias])
        for k, v in dict(x=x, w=m.proj.weight, b=m.proj.bias, y=y, dout=dout, dx=gr[0], dw=gr[1], db=gr[2]).items():
            d[f"{name}/{k}"] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "patch_embed.npz"), **d)


def vit_config(size="custom", hidden=64, mlp=128, layers=2, heads=1, patch=(1, 8, 8), t=1, h=32, w=32, task="seg"):
    return types.SimpleNamespace(ViT=types.SimpleNamespace(size=size, hidden_size=hidden, mlp_dim=mlp, num_layers=layers,
                                                           num_heads=heads, patch_size=list(patch), use_hyena=False,
                                                           use_mamba=False), time=t, height=h, width=w, task_type=task)


def swin_config(size="custom", embed=16, depths=(2, 2, 2, 2), heads=(1, 2, 4, 8), patch=(1, 2, 2), window=(1, 4, 4), t=1,
                h=48, w=40, task="seg"):
    return types.SimpleNamespace(Swin=types.SimpleNamespace(size=size, embed_dim=embed, depths=list(depths),
                                                            num_heads=list(heads), patch_size=list(patch),
                                                            window_size=list(window), use_hyena=False, use_mamba=False),
                                 time=t, height=h, width=w, task_type=task)


ENCODER_CASES = {
    "vit2d_seg": ("vit", dict(hidden=128, mlp=256, layers=2, heads=2, patch=(1, 8, 8), t=1, h=32, w=48), (2, 1, 1, 32, 48)),
    "vit2d_cls": ("vit", dict(hidden=64, mlp=128, layers=2, heads=1, patch=(1, 8, 8), t=1, h=32, w=32, task="class"), (2, 3, 1, 32, 32)),
    "vit3d_seg": ("vit", dict(hidden=64, mlp=128, layers=2, heads=1, patch=(4, 8, 8), t=8, h=16, w=16), (1, 1, 8, 16, 16)),
    "swin2d": ("swin", dict(embed=16, heads=(1, 2, 4, 8), patch=(1, 2, 2), window=(1, 4, 4), t=1, h=48, w=40), (2, 1, 1, 48, 40)),
    "swin3d": ("swin", dict(embed=32, heads=(2, 4, 8, 16), patch=(2, 2, 2), window=(3, 3, 3), t=16, h=24, w=20), (1, 1, 16, 24, 20)),
}


def make_encoders(vit, swin):
    d = {}
    for name, (kind, kw, in_shape) in ENCODER_CASES.items():
        if kind == "vit":
            model, _ = vit.custom_ViT(vit_config(**kw), in_shape[1])
        else:
            model, _ = swin.custom_Swin(swin_config(**kw), in_shape[1])
        fill_parameters_(model, 31)
        g = torch.Generator().manual_seed(9)
        x = torch.randn(*in_shape, generator=g)
        outs = model(x)
        loss = sum((o.float() * torch.linspace(-1, 1, o.numel()).reshape(o.shape)).sum() for o in outs[1:])
        loss.backward()
        d[f"{name}/n_out"] = np.array(len(outs))
        for i, o in enumerate(outs):
            d[f"{name}/out{i}"] = o.detach().numpy()
        keys = []
        for pname, p in model.named_parameters():
            keys.append(pname)
            gnp = p.grad.detach().numpy()
            d[f"{name}/gradnorm/{pname}"] = np.array(np.linalg.norm(gnp.astype(np.float64)))
            if gnp.size <= 4096:
                d[f"{name}/grad/{pname}"] = gnp
        d[f"{name}/param_names"] = np.array(keys)
        d[f"{name}/state_keys"] = np.array(list(model.state_dict().keys()))
        d[f"{name}/state_shapes"] = np.array([",".join(map(str, v.shape)) for v in model.state_dict().values()])
    np.savez_compressed(os.path.join(OUT, "encoders.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    vit, swin = load_reference()
    make_window_maps(swin)
    make_sablock(vit)
    make_swin_part1(swin)
    make_patch_embed()
    make_encoders(vit, swin)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
