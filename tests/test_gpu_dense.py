"""GPU parity tests (through the C ABI) for the dense/ViT path: attention core, SABlock, patch embedding,
whole ViT encoder. Tolerances from BASELINE.json north_star: bf16 path max-rel <= 2e-2 against the fp32
reference on the same inputs; fp32 patch-embed path <= 1e-4."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden, max_rel
from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
FP32_TOL = 1e-4


def _t(a, dev="cuda"):
    return torch.from_numpy(np.asarray(a)).to(dev)


@pytest.mark.parametrize("B,H,N", [(2, 3, 196), (2, 3, 197), (1, 2, 128), (1, 1, 1), (1, 2, 129), (1, 12, 1728),
                                   (3, 2, 640), (4, 12, 1728), (1, 12, 8192)])   # SURVEY Appendix C dense matrix
def test_dense_attention_core_vs_oracle(B, H, N):
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(B * 1000 + N)
    d = 64
    qkv = torch.randn(B, N, 3, H, d, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.dense_attn_fwd(q, k, v, d ** -0.5)
    d_o = torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16)
    dq, dk, dv = ops.dense_attn_bwd(q, k, v, o, d_o, lse, d ** -0.5)

    ref_qkv = qkv.float().requires_grad_(True)
    rq, rk, rv = [ref_qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3)]
    ref = ao.dense_attention(rq, rk, rv, d ** -0.5).permute(0, 2, 1, 3)
    (g,) = torch.autograd.grad(ref, ref_qkv, d_o.float())
    assert max_rel(o.float().cpu(), ref.detach().cpu()) < BF16_TOL
    ref_lse = torch.logsumexp(torch.einsum("bhxd,bhyd->bhxy", rq, rk).detach() * d ** -0.5, -1)
    assert max_rel(lse.cpu(), ref_lse.cpu()) < 1e-4
    for got, want, name in ((dq, g[:, :, 0], "dq"), (dk, g[:, :, 1], "dk"), (dv, g[:, :, 2], "dv")):
        assert max_rel(got.float().cpu(), want.cpu()) < BF16_TOL, name


def test_dense_attention_cross_lengths():
    """Nq != Nk (what a ring step sees)."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(5)
    B, H, Nq, Nk, d = 2, 2, 200, 333, 64
    q = torch.randn(B, Nq, H, d, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Nk, H, d, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Nk, H, d, device="cuda").to(torch.bfloat16)
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    d_o = torch.randn_like(o)
    dq, dk, dv = ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
    qf, kf, vf = [t.float().requires_grad_(True) for t in (q, k, v)]
    ref = ao.dense_attention(qf.permute(0, 2, 1, 3), kf.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), 0.125)
    ref = ref.permute(0, 2, 1, 3)
    gq, gk, gv = torch.autograd.grad(ref, [qf, kf, vf], d_o.float())
    assert max_rel(o.float().cpu(), ref.detach().cpu()) < BF16_TOL
    assert max_rel(dq.float().cpu(), gq.cpu()) < BF16_TOL
    assert max_rel(dk.float().cpu(), gk.cpu()) < BF16_TOL
    assert max_rel(dv.float().cpu(), gv.cpu()) < BF16_TOL


def test_dense_backward_is_linear_in_grad_out():
    """GradScaler multiplies the loss: the backward must be linear in grad_out. (Not bit-exact: the dQ partials
    are summed with fp32 TMA reduce-adds whose order varies, so results may differ by one bf16 ulp.)"""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(1)
    q, k, v = [torch.randn(1, 300, 2, 64, device="cuda").to(torch.bfloat16) for _ in range(3)]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    d_o = torch.randn_like(o)
    a = ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
    b = ops.dense_attn_bwd(q, k, v, o, d_o * 1024, lse, 0.125)
    for x, y in zip(a, b):
        assert max_rel(y.float().cpu(), (x.float() * 1024).cpu()) < 1e-2


@pytest.mark.parametrize("case", ["a", "b"])
@pytest.mark.parametrize("autocast", [True, False])
def test_sablock_module_vs_golden(case, autocast):
    from long_context_biomedical_imaging_b200.backbone_vit import SABlock

    g = load_golden("sablock.npz")
    H = int(g[f"{case}/heads"])
    C = g[f"{case}/x"].shape[-1]
    blk = SABlock(False, False, C, H).cuda()
    with torch.no_grad():
        blk.qkv.weight.copy_(_t(g[f"{case}/w_qkv"]))
        blk.out_proj.weight.copy_(_t(g[f"{case}/w_out"]))
        blk.out_proj.bias.copy_(_t(g[f"{case}/b_out"]))
    x = _t(g[f"{case}/x"]).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = blk(x)
    y.float().backward(_t(g[f"{case}/dout"]))
    assert max_rel(y.detach().float().cpu(), g[f"{case}/y"]) < BF16_TOL
    assert max_rel(x.grad.cpu(), g[f"{case}/dx"]) < BF16_TOL
    assert max_rel(blk.qkv.weight.grad.cpu(), g[f"{case}/dw_qkv"]) < BF16_TOL
    assert max_rel(blk.out_proj.weight.grad.cpu(), g[f"{case}/dw_out"]) < BF16_TOL
    assert max_rel(blk.out_proj.bias.grad.cpu(), g[f"{case}/db_out"]) < BF16_TOL


def test_patch_embed_vs_golden():
    from long_context_biomedical_imaging_b200 import ops

    g = load_golden("patch_embed.npz")
    for name in ["vit2d", "vit3d", "vit2d_p2", "swin2d", "swin3d"]:
        x = _t(g[f"{name}/x"]).requires_grad_(True)
        w = _t(g[f"{name}/w"]).requires_grad_(True)
        b = _t(g[f"{name}/b"]).requires_grad_(True)
        is_vit = name.startswith("vit")
        pos = _t(g[f"{name}/pos"]).requires_grad_(True) if is_vit else None
        patch = w.shape[2:]
        if is_vit:
            grid = [s // p for s, p in zip(x.shape[2:], patch)]
        else:
            grid = [-(-s // p) for s, p in zip(x.shape[2:], patch)]
        y = ops.patch_embed(x, w, b, pos, grid, torch.float32)
        want_y = g[f"{name}/y"]
        dout = _t(g[f"{name}/dout"])
        if not is_vit:  # golden is channel-first (B,C,*grid); ours is token-major
            want_y = np.moveaxis(want_y, 1, -1).reshape(want_y.shape[0], -1, want_y.shape[1])
            dout = dout.movedim(1, -1).reshape(dout.shape[0], -1, dout.shape[1]).contiguous()
        assert max_rel(y.detach().cpu(), want_y) < FP32_TOL, name
        y.backward(dout)
        assert max_rel(x.grad.cpu(), g[f"{name}/dx"]) < FP32_TOL, name
        assert max_rel(w.grad.cpu(), g[f"{name}/dw"]) < FP32_TOL, name
        assert max_rel(b.grad.cpu(), g[f"{name}/db"]) < FP32_TOL, name
        if is_vit:
            assert max_rel(pos.grad.cpu(), g[f"{name}/dpos"]) < FP32_TOL, name


@pytest.mark.parametrize("shape", [
    # (B, Cin, image dims, patch, hidden): K = Cin * prod(patch) >= 64 and % 32 == 0 -> split-bf16 tensor-core kernels
    (2, 1, (1, 48, 80), (1, 16, 16), 192),      # cfg1-like, K = 256, M = 30 (ragged tile)
    (3, 1, (16, 24, 40), (8, 8, 8), 200),       # cfg3-like, K = 512, N not a multiple of 128
    (2, 3, (1, 20, 24), (1, 4, 8), 72),         # K = 96: three 32-chunks, partial 128-wide k tile in the backward
    (2, 1, (1, 224, 224), (1, 16, 16), 192),    # cfg1 in full: M = 392 (3 tiles + tail), N = 192 (1.5 feature tiles)
    (2, 1, (32, 48, 96), (8, 8, 8), 768),       # cfg3 geometry per image row, M = 576, all 6 feature tiles, 4 dW k tiles
    (1, 2, (1, 40, 64), (1, 8, 8), 136),        # K = 128 over two channels, N % 8 == 0 but not % 16
])
def test_patch_embed_large_k_tensor_core_path_fp32_accuracy(shape):
    """The K >= 64 path multiplies hi/lo bf16 splits on the tensor cores; it must still meet the fp32 tolerance
    against torch's fp32 strided convolution (the MONAI PatchEmbeddingBlock semantics)."""
    import torch.nn.functional as F
    from long_context_biomedical_imaging_b200 import ops

    B, cin, dims, patch, hidden = shape
    torch.manual_seed(5)
    x = torch.randn(B, cin, *dims, device="cuda", requires_grad=True)
    w = (torch.randn(hidden, cin, *patch, device="cuda") * 0.05).requires_grad_(True)
    b = torch.randn(hidden, device="cuda", requires_grad=True)
    grid = [s // p for s, p in zip(dims, patch)]
    n_tok = grid[0] * grid[1] * grid[2]
    pos = torch.randn(1, n_tok, hidden, device="cuda", requires_grad=True)
    y = ops.patch_embed(x, w, b, pos, grid, torch.float32)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv3d(x.double(), w.double(), b.double(), stride=patch).flatten(2).transpose(1, 2) + pos.double()
        dout = torch.randn_like(y)
        gx, gw, gb, gp = torch.autograd.grad(ref, (x, w, b, pos), dout.double())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert max_rel(y.detach().cpu(), ref.detach().float().cpu()) < FP32_TOL
    y.backward(dout)
    assert max_rel(w.grad.cpu(), gw.float().cpu()) < FP32_TOL
    assert max_rel(b.grad.cpu(), gb.float().cpu()) < FP32_TOL
    assert max_rel(pos.grad.cpu(), gp.float().cpu()) < FP32_TOL
    assert max_rel(x.grad.cpu(), gx.float().cpu()) < FP32_TOL


@pytest.mark.parametrize("name,B,cin,dims,patch,hidden,vit", [
    ("cfg2", 4, 1, (1, 512, 512), (1, 4, 4), 96, False),        # Swin-T 2D patch 4: K = 16, bf16 tokens
    ("cfg4", 1, 1, (128, 128, 128), (2, 2, 2), 48, False),      # Swin 3D 'unetr' patch 2: K = 8
    ("cfg5", 1, 1, (1, 1024, 1024), (1, 2, 2), 768, True),      # ViT-B 2D patch 2: K = 4, 262,144 tokens, fp32 + pos
    ("pad", 2, 1, (1, 70, 90), (1, 4, 4), 32, False),           # trailing zero pad (PatchEmbed), ragged last chunk
    ("cin3", 2, 2, (4, 12, 136), (2, 2, 2), 40, True),          # K = 16 over two channels, 68 patches per row (2 chunks)
])
def test_patch_embed_small_k_streaming_kernel_full_size(name, B, cin, dims, patch, hidden, vit):
    """Full-size parity of the K in {4, 8, 16} streaming forward (and the small-K backward) against torch's fp32 strided
    convolution: MONAI PatchEmbeddingBlock (ViT: + position embedding, fp32) and PatchEmbed (Swin: zero pad, bf16)."""
    import torch.nn.functional as F
    from long_context_biomedical_imaging_b200 import ops

    def _mr(a, b):                                     # conftest.max_rel on the device (the cfg5 tensors hold 2e8 values)
        return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()

    torch.manual_seed(11)
    x = torch.randn(B, cin, *dims, device="cuda")
    w = (torch.randn(hidden, cin, *patch, device="cuda") * 0.2).requires_grad_(True)
    b = torch.randn(hidden, device="cuda", requires_grad=True)
    grid = [-(-s // p) for s, p in zip(dims, patch)]
    n_tok = grid[0] * grid[1] * grid[2]
    pos = torch.randn(1, n_tok, hidden, device="cuda", requires_grad=True) if vit else None
    out_dtype = torch.float32 if vit else torch.bfloat16
    y = ops.patch_embed(x, w, b, pos, grid, out_dtype)
    pads = []
    for s_, p_ in zip(reversed(dims), reversed(patch)):
        pads += [0, (p_ - s_ % p_) % p_]
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv3d(F.pad(x, pads), w, b, stride=patch).flatten(2).transpose(1, 2)
        if pos is not None:
            ref = ref + pos
        dout = torch.randn(ref.shape, device="cuda", dtype=out_dtype)
        grads = torch.autograd.grad(ref, (w, b) + ((pos,) if vit else ()), dout.float())
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert y.shape == ref.shape and y.dtype == out_dtype
    if vit:
        assert _mr(y.detach(), ref.detach()) < FP32_TOL
    else:
        assert _mr(y.detach().float(), ref.detach().to(torch.bfloat16).float()) < BF16_TOL
    y.backward(dout)
    assert _mr(w.grad, grads[0]) < 2e-4          # fp32 sums over up to 2.6e5 patches in a different order
    assert _mr(b.grad, grads[1]) < 2e-4
    if vit:
        assert _mr(pos.grad, grads[2]) < FP32_TOL


def test_patch_embed_workspace_contract_through_the_c_abi():
    """lcbi_patch_embed_fwd_ws: a short or misaligned workspace is refused with a message (no silent fallback onto a
    different kernel), NULL selects the workspace-free kernels, and both give the same result as the full workspace."""
    from long_context_biomedical_imaging_b200 import _lib

    lib = _lib.load()
    torch.manual_seed(2)
    B, dims, patch, N = 2, (1, 32, 48), (1, 8, 8), 64
    grid = (1, 4, 6)
    x = torch.randn(B, 1, *dims, device="cuda")
    w = torch.randn(N, 64, device="cuda") * 0.1
    b = torch.randn(N, device="cuda")
    need = lib.lcbi_patch_embed_workspace_bytes(B, 1, _lib.int3(patch), _lib.int3(grid), N)
    assert need > 0
    ws = torch.empty(need + 256, dtype=torch.uint8, device="cuda")
    outs = []
    for ws_ptr, ws_bytes in ((ws.data_ptr(), need), (None, 0)):
        out = torch.empty(B, 24, N, device="cuda")
        rc = lib.lcbi_patch_embed_fwd_ws(x.data_ptr(), 0, w.data_ptr(), b.data_ptr(), None, out.data_ptr(), 0, B, 1,
                                         _lib.int3(dims), _lib.int3(patch), _lib.int3(grid), N, ws_ptr, ws_bytes, None)
        assert rc == 0, lib.lcbi_last_error()
        outs.append(out)
    torch.cuda.synchronize()
    assert max_rel(outs[0].cpu(), outs[1].cpu()) < FP32_TOL
    out = torch.empty(B, 24, N, device="cuda")
    for ws_ptr, ws_bytes in ((ws.data_ptr(), need - 1024), (ws.data_ptr() + 16, need)):
        rc = lib.lcbi_patch_embed_fwd_ws(x.data_ptr(), 0, w.data_ptr(), b.data_ptr(), None, out.data_ptr(), 0, B, 1,
                                         _lib.int3(dims), _lib.int3(patch), _lib.int3(grid), N, ws_ptr, ws_bytes, None)
        assert rc != 0 and b"workspace" in lib.lcbi_last_error()


def test_patch_embed_streaming_kernels_random_geometries():
    """K in {4, 8, 16} streaming forward / narrow-row backward against F.conv over random image sizes (with and without
    trailing pad), patch shapes, channel counts and feature widths: catches index-map slips the named configs miss."""
    import random

    import torch.nn.functional as F
    from long_context_biomedical_imaging_b200 import ops

    rng = random.Random(7)
    patches = [(1, 2, 2), (1, 4, 4), (1, 2, 4), (1, 4, 2), (2, 2, 2), (2, 2, 4), (1, 1, 4), (4, 2, 2), (1, 2, 8), (1, 4, 1)]
    for trial in range(14):
        patch = rng.choice(patches)
        cin = rng.choice([1, 1, 2, 4])
        k = cin * patch[0] * patch[1] * patch[2]
        if k not in (4, 8, 16):
            continue
        dims = tuple(p * rng.randint(2, 40 if i == 2 else 6) + (rng.randint(0, p - 1) if rng.random() < 0.5 else 0)
                     for i, p in enumerate(patch))
        if patch[0] == 1:
            dims = (1,) + dims[1:]
        hidden = rng.choice([8, 24, 48, 96, 132, 256, 320])
        B = rng.randint(1, 3)
        vit = rng.random() < 0.5
        torch.manual_seed(trial)
        x = torch.randn(B, cin, *dims, device="cuda")
        w = (torch.randn(hidden, cin, *patch, device="cuda") * 0.3).requires_grad_(True)
        b = torch.randn(hidden, device="cuda", requires_grad=True)
        grid = [-(-s_ // p_) for s_, p_ in zip(dims, patch)]
        pos = torch.randn(1, grid[0] * grid[1] * grid[2], hidden, device="cuda", requires_grad=True) if vit else None
        y = ops.patch_embed(x, w, b, pos, grid, torch.float32)
        pads = []
        for s_, p_ in zip(reversed(dims), reversed(patch)):
            pads += [0, (p_ - s_ % p_) % p_]
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ref = F.conv3d(F.pad(x, pads).double(), w.double(), b.double(), stride=patch).flatten(2).transpose(1, 2)
            if pos is not None:
                ref = ref + pos.double()
            dout = torch.randn(ref.shape, device="cuda")
            grads = torch.autograd.grad(ref, (w, b) + ((pos,) if vit else ()), dout.double())
        finally:
            torch.backends.cudnn.allow_tf32 = prev
        tag = (trial, patch, cin, dims, hidden, B, vit)
        assert max_rel(y.detach().cpu(), ref.detach().float().cpu()) < FP32_TOL, tag
        y.backward(dout)
        assert max_rel(w.grad.cpu(), grads[0].float().cpu()) < FP32_TOL, tag
        assert max_rel(b.grad.cpu(), grads[1].float().cpu()) < FP32_TOL, tag
        if vit:
            assert max_rel(pos.grad.cpu(), grads[2].float().cpu()) < FP32_TOL, tag


def _vit_cfg(hidden, mlp, layers, heads, patch, t, h, w, task="seg"):
    return types.SimpleNamespace(ViT=types.SimpleNamespace(size="custom", hidden_size=hidden, mlp_dim=mlp, num_layers=layers,
                                                           num_heads=heads, patch_size=list(patch), use_hyena=False,
                                                           use_mamba=False), time=t, height=h, width=w, task_type=task)


VIT_CASES = {
    "vit2d_seg": (dict(hidden=128, mlp=256, layers=2, heads=2, patch=(1, 8, 8), t=1, h=32, w=48), (2, 1, 1, 32, 48)),
    "vit2d_cls": (dict(hidden=64, mlp=128, layers=2, heads=1, patch=(1, 8, 8), t=1, h=32, w=32, task="class"), (2, 3, 1, 32, 32)),
    "vit3d_seg": (dict(hidden=64, mlp=128, layers=2, heads=1, patch=(4, 8, 8), t=8, h=16, w=16), (1, 1, 8, 16, 16)),
}


@pytest.mark.parametrize("name", list(VIT_CASES))
def test_vit_encoder_drop_in_vs_golden(name):
    """Same config + same (seeded) parameters as the reference encoder -> same state_dict layout, same
    hidden-state list, same parameter gradients (every parameter receives one: DDP find_unused=False)."""
    from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT

    g = load_golden("encoders.npz")
    kw, in_shape = VIT_CASES[name]
    model, chans = custom_ViT(_vit_cfg(**kw), in_shape[1])
    assert chans == [kw["hidden"]] * 13
    assert list(model.state_dict().keys()) == [str(k) for k in g[f"{name}/state_keys"]]
    assert [",".join(map(str, v.shape)) for v in model.state_dict().values()] == [str(s) for s in g[f"{name}/state_shapes"]]
    ao.fill_parameters_(model, 31)
    model = model.cuda()
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(*in_shape, generator=gen).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    assert len(outs) == int(g[f"{name}/n_out"])
    for i, o in enumerate(outs):
        want = g[f"{name}/out{i}"]
        assert tuple(o.shape) == want.shape
        assert max_rel(o.detach().float().cpu(), want) < BF16_TOL, (name, i)
    loss = sum((o.float() * torch.linspace(-1, 1, o.numel(), device="cuda").reshape(o.shape)).sum() for o in outs[1:])
    loss.backward()
    for pname, p in model.named_parameters():
        assert p.grad is not None, pname
        want_norm = float(g[f"{name}/gradnorm/{pname}"])
        got_norm = float(p.grad.double().norm())
        assert abs(got_norm - want_norm) <= 3e-2 * max(want_norm, 1e-6), (pname, got_norm, want_norm)
        key = f"{name}/grad/{pname}"
        if key in g.files:
            assert max_rel(p.grad.cpu(), g[key]) < 3e-2, pname


def test_vit_encoder_inference_mode_and_errors():
    from long_context_biomedical_imaging_b200 import ops
    from long_context_biomedical_imaging_b200.backbone_vit import SABlock, custom_ViT

    kw, in_shape = VIT_CASES["vit2d_seg"]
    model, _ = custom_ViT(_vit_cfg(**kw), in_shape[1])
    model = model.cuda().eval()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(torch.randn(*in_shape, device="cuda"))
    assert len(outs) == 4
    with pytest.raises(ValueError):
        SABlock(False, False, 100, 3)
    with pytest.raises(ValueError):
        custom_ViT(types.SimpleNamespace(ViT=types.SimpleNamespace(size="huge"), time=1, height=8, width=8, task_type="seg"), 1)
    with pytest.raises(RuntimeError):   # CPU tensors are rejected: no fallback
        ops.dense_attention_qkv(torch.randn(1, 4, 192), 1)
    with pytest.raises(ValueError):     # head_dim != 64
        ops.dense_attention_qkv(torch.randn(1, 4, 96, device="cuda"), 1)


def test_dense_full_size_properties():
    """cfg3 full size (B=4): rows of P sum to one => with V == 1 the output is exactly 1; and the output is
    invariant to a permutation of the keys/values (size-independent properties, no N^2 oracle needed)."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(3)
    B, N, H, d = 4, 1728, 12, 64
    q, k = [torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16) for _ in range(2)]
    v1 = torch.ones(B, N, H, d, device="cuda", dtype=torch.bfloat16)
    o, _ = ops.dense_attn_fwd(q, k, v1, 0.125)
    assert (o.float() - 1).abs().max().item() < 1e-2
    v = torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16)
    perm = torch.randperm(N, device="cuda")
    o_a, lse_a = ops.dense_attn_fwd(q, k, v, 0.125)
    o_b, lse_b = ops.dense_attn_fwd(q, k[:, perm].contiguous(), v[:, perm].contiguous(), 0.125)
    assert max_rel(o_b.float().cpu(), o_a.float().cpu()) < 1e-2
    assert max_rel(lse_b.cpu(), lse_a.cpu()) < 1e-5


def test_dense_backward_full_size_properties():
    """cfg3 at the bench size (B=16): size-independent identities of the attention gradient.
    (1) with every value row equal, O does not depend on the scores, so dQ = dK = 0;
    (2) rows of P sum to one, so sum_keys dV[key, :] = sum_queries dO[query, :]."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(5)
    B, N, H, d = 16, 1728, 12, 64
    q, k, v = [torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16) for _ in range(3)]
    d_o = torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16)
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    dq, dk, dv = ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
    want = d_o.float().sum(dim=1)                       # (B, H, d)
    got = dv.float().sum(dim=1)
    assert max_rel(got.cpu(), want.cpu()) < 1e-2
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all()

    v_same = torch.randn(B, 1, H, d, device="cuda").to(torch.bfloat16).expand(B, N, H, d).contiguous()
    o2, lse2 = ops.dense_attn_fwd(q, k, v_same, 0.125)
    assert max_rel(o2.float().cpu(), v_same.float().cpu()) < 1e-2
    dq2, dk2, _ = ops.dense_attn_bwd(q, k, v_same, o2, d_o, lse2, 0.125)
    scale_ref = float(dq.float().abs().max())            # gradient magnitude of the generic case
    assert float(dq2.float().abs().max()) < 2e-2 * scale_ref
    assert float(dk2.float().abs().max()) < 2e-2 * float(dk.float().abs().max())


def test_dense_forward_with_carried_state_equals_one_pass():
    """lcbi_dense_attn_fwd_state: folding the K/V sequence in 3 ragged shards through the carried (m, l, O) state gives
    the result of one pass over all keys (same kernel arithmetic; lse to fp32 rounding, o to one bf16 ulp)."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(11)
    B, H, Nq, d = 2, 3, 333, 64
    cuts = [0, 200, 264, 777]
    q = torch.randn(B, Nq, H, d, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, cuts[-1], H, d, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, cuts[-1], H, d, device="cuda").to(torch.bfloat16)
    k[:, 300:] *= 3.0                       # later shards raise the running max: the rescale path runs
    o_ref, lse_ref = ops.dense_attn_fwd(q, k, v, 0.125)
    state = (torch.empty(B, Nq, H, d, device="cuda"), torch.empty(B, H, Nq, device="cuda"),
             torch.empty(B, H, Nq, device="cuda"))
    out = lse = None
    for i in range(3):
        ks, vs = k[:, cuts[i]:cuts[i + 1]], v[:, cuts[i]:cuts[i + 1]]
        out, lse = ops.dense_attn_fwd_state(q, ks, vs, 0.125, state, first=i == 0, last=i == 2)
    assert max_rel(lse.cpu(), lse_ref.cpu()) < 1e-5
    assert max_rel(out.float().cpu(), o_ref.float().cpu()) < 1e-2
    qf, kf, vf = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    ref = ao.dense_attention(qf, kf, vf, 0.125).permute(0, 2, 1, 3)
    assert max_rel(out.float().cpu(), ref.cpu()) < BF16_TOL


def test_dense_rejects_operands_on_mixed_devices_and_bad_dtypes():
    from long_context_biomedical_imaging_b200 import ops

    q, k, v = [torch.randn(1, 64, 1, 64, device="cuda").to(torch.bfloat16) for _ in range(3)]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    with pytest.raises(ValueError):          # fp32 `o` (e.g. a ring accumulator passed by mistake)
        ops.dense_attn_bwd(q, k, v, o.float(), torch.randn_like(o), lse, 0.125)
    with pytest.raises(ValueError):          # bf16 dk with accumulate_dkv
        ops.dense_attn_bwd(q, k, v, o, torch.randn_like(o), lse, 0.125, dk=torch.zeros_like(k), dv=torch.zeros_like(v),
                           accumulate_dkv=True)
    with pytest.raises(ValueError):
        ops.attn_merge(torch.zeros(1, 64, 1, 64, device="cuda"), lse.clone(), o.float(), lse, True)
    if torch.cuda.device_count() > 1:
        with pytest.raises(ValueError):
            ops.dense_attn_fwd(q, k.to("cuda:1"), v, 0.125)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_ops_run_on_the_operands_device_not_the_current_one():
    """a model on cuda:1 while the process default is cuda:0 must launch on cuda:1's stream."""
    from long_context_biomedical_imaging_b200 import ops

    torch.cuda.set_device(0)
    torch.manual_seed(3)
    q, k, v = [torch.randn(1, 200, 2, 64, device="cuda:1").to(torch.bfloat16) for _ in range(3)]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    assert o.device == q.device and torch.cuda.current_device() == 0
    qf, kf, vf = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    ref = ao.dense_attention(qf, kf, vf, 0.125).permute(0, 2, 1, 3)
    assert max_rel(o.float().cpu(), ref.cpu()) < BF16_TOL


def test_sablock_under_fp16_autocast_with_grad_scaler():
    """What the unmodified reference harness selects on a B200 (utils/status.py:50-58 does not list it as bf16-capable:
    fp16 autocast + GradScaler, trainer_base.py:116,166-171). The kernels compute in bf16 (fp16 qkv is cast), gradients
    arrive scaled by the GradScaler's factor and are un-scaled by it: results must match the fp32 reference."""
    from long_context_biomedical_imaging_b200.backbone_vit import SABlock

    g = load_golden("sablock.npz")
    H = int(g["a/heads"])
    C = g["a/x"].shape[-1]
    blk = SABlock(False, False, C, H).cuda()
    with torch.no_grad():
        blk.qkv.weight.copy_(_t(g["a/w_qkv"]))
        blk.out_proj.weight.copy_(_t(g["a/w_out"]))
        blk.out_proj.bias.copy_(_t(g["a/b_out"]))
    x = _t(g["a/x"]).requires_grad_(True)
    opt = torch.optim.SGD(blk.parameters(), lr=0.0)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    with torch.autocast("cuda", dtype=torch.float16):
        y = blk(x)
    loss = (y.float() * _t(g["a/dout"])).sum()
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    assert max_rel(y.detach().float().cpu(), g["a/y"]) < BF16_TOL
    assert max_rel(x.grad.cpu() / 1024.0, g["a/dx"]) < BF16_TOL          # x is not an optimizer parameter: still scaled
    assert max_rel(blk.qkv.weight.grad.cpu(), g["a/dw_qkv"]) < BF16_TOL
    assert max_rel(blk.out_proj.weight.grad.cpu(), g["a/dw_out"]) < BF16_TOL
    assert max_rel(blk.out_proj.bias.grad.cpu(), g["a/db_out"]) < BF16_TOL


def test_vit_encoder_under_ddp_without_unused_parameters():
    """trainer_base.py:94-98 wraps the model in DDP(find_unused_parameters=False): every parameter must receive a
    gradient in every step or the reducer raises on the next forward. One-process NCCL group, two steps."""
    import os
    import socket

    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT

    cfg = types.SimpleNamespace(ViT=types.SimpleNamespace(size="custom", hidden_size=128, mlp_dim=256, num_layers=2,
                                                          num_heads=2, patch_size=[1, 8, 8], use_hyena=False,
                                                          use_mamba=False),
                                time=1, height=32, width=32, task_type="class")
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        model, _ = custom_ViT(cfg, 1)
        ddp = DDP(model.cuda(), find_unused_parameters=False)
        for _ in range(2):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs = ddp(torch.randn(2, 1, 1, 32, 32, device="cuda"))
            sum(o.float().sum() for o in outs[1:]).backward()
        assert all(p.grad is not None for p in model.parameters())
    finally:
        dist.destroy_process_group()


def test_backward_as_first_cuda_call_of_the_autograd_thread():
    """The TMA tensor maps are encoded through the driver API, which needs a context bound to the calling thread. When
    the attention backward is the FIRST node autograd's worker thread executes, that thread has made no runtime call yet
    (found by bench.py's parity_check in round 2: CUDA_ERROR_INVALID_CONTEXT); the library binds the context and retries.
    Run in a fresh interpreter so that the worker thread is new."""
    import subprocess
    import sys

    code = (
        "import torch, sys; sys.path.insert(0, %r)\n"
        "from long_context_biomedical_imaging_b200 import ops\n"
        "x = torch.randn(1, 256, 3 * 128, device='cuda').to(torch.bfloat16).requires_grad_(True)\n"
        "ops.dense_attention_qkv(x, 2).backward(torch.ones(1, 256, 128, device='cuda', dtype=torch.bfloat16))\n"
        "torch.cuda.synchronize(); assert x.grad is not None; print('ok')\n" % __import__("conftest").ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
