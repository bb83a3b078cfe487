"""GPU parity tests (through the C ABI) for the Swin path: index maps (bit-exact), shifted-window attention block
forward/backward against the golden fixtures of the unmodified reference, whole Swin encoders, full-size checks.
Tolerance: bf16 path max-rel <= 2e-2 (BASELINE.json north_star); index maps bit-exact."""
import types

import numpy as np
import pytest
import torch

from conftest import check_stored, load_golden, max_rel
from oracle import attention_oracle as ao
from oracle import window_maps as wm

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _t(a):
    return torch.from_numpy(np.asarray(a)).cuda()


def test_window_maps_bit_exact_vs_reference_golden():
    from long_context_biomedical_imaging_b200 import ops

    g = load_golden("window_maps.npz")
    for case in g["cases"]:
        grid = tuple(int(x) for x in g[f"{case}/grid"])
        window = tuple(int(x) for x in g[f"{case}/window"])
        shift = tuple(int(x) for x in g[f"{case}/shift"])
        gather, region, relidx = ops.window_maps(grid, window, shift)
        check_stored(g, f"{case}/gather_map", gather.cpu().numpy().astype(np.int64))
        reg = region.cpu().numpy()
        if reg.shape[0] * reg.shape[1] ** 2 <= 2e7:
            mask = np.where(reg[:, None, :] != reg[:, :, None], np.float32(-100.0), np.float32(0.0)).astype(np.float32)
            check_stored(g, f"{case}/mask", mask)
        else:   # the big masks are pinned on the CPU (test_oracle_golden); compare region ids with that oracle
            assert np.array_equal(reg, wm.region_ids(grid, window, shift))
        n = gather.shape[1]
        assert np.array_equal(relidx.cpu().numpy().astype(np.int64), wm.rel_pos_index_used(window, n))
    for key in [k for k in g.files if k.startswith("relidx/") and k.endswith("/shape")]:
        window = tuple(int(x) for x in key.split("/")[1].split("x"))
        grid = tuple(2 * w for w in window)
        _, _, relidx = ops.window_maps(grid, window, tuple(0 for _ in window))
        check_stored(g, "relidx/" + key.split("/")[1], relidx.cpu().numpy().astype(np.int64))


def _swin_case(g, name):
    meta = [int(v) for v in g[f"{name}/meta"]]
    B, C, H, K = meta[:4]
    return B, C, H, tuple(meta[4:4 + K]), tuple(meta[4 + K:4 + 2 * K]), tuple(meta[4 + 2 * K:4 + 3 * K])


@pytest.mark.parametrize("name", ["2d_shift", "2d_noshift", "2d_w4", "3d_shift", "3d_clamp", "3d_allclamp"])
@pytest.mark.parametrize("autocast", [True, False])
def test_swin_block_part1_vs_golden(name, autocast):
    from long_context_biomedical_imaging_b200.backbone_swin import SwinTransformerBlock

    g = load_golden("swin_part1.npz")
    B, C, H, grid, window, shift = _swin_case(g, name)
    blk = SwinTransformerBlock(False, False, dim=C, num_heads=H, window_size=window, shift_size=shift).cuda()
    a = blk.attn
    with torch.no_grad():
        for param, key in ((blk.norm1.weight, "norm_w"), (blk.norm1.bias, "norm_b"), (a.qkv.weight, "w_qkv"),
                           (a.qkv.bias, "b_qkv"), (a.relative_position_bias_table, "table"), (a.proj.weight, "w_proj"),
                           (a.proj.bias, "b_proj")):
            param.copy_(_t(g[f"{name}/{key}"]))
    x = _t(g[f"{name}/x"]).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = blk.forward_part1(x, None)
    assert y.shape == x.shape
    assert max_rel(y.detach().float().cpu(), g[f"{name}/y"]) < BF16_TOL
    y.float().backward(_t(g[f"{name}/dout"]))
    got = dict(dx=x.grad, dnorm_w=blk.norm1.weight.grad, dnorm_b=blk.norm1.bias.grad, dw_qkv=a.qkv.weight.grad,
               db_qkv=a.qkv.bias.grad, dtable=a.relative_position_bias_table.grad, dw_proj=a.proj.weight.grad,
               db_proj=a.proj.bias.grad)
    for key, val in got.items():
        assert val is not None, key
        assert max_rel(val.float().cpu(), g[f"{name}/{key}"]) < BF16_TOL, (name, key)


def _swin_cfg(embed, heads, patch, window, t, h, w, size="custom", depths=(2, 2, 2, 2)):
    return types.SimpleNamespace(Swin=types.SimpleNamespace(size=size, embed_dim=embed, depths=list(depths),
                                                            num_heads=list(heads), patch_size=list(patch),
                                                            window_size=list(window), use_hyena=False, use_mamba=False),
                                 time=t, height=h, width=w, task_type="seg")


SWIN_CASES = {
    "swin2d": (dict(embed=16, heads=(1, 2, 4, 8), patch=(1, 2, 2), window=(1, 4, 4), t=1, h=48, w=40), (2, 1, 1, 48, 40)),
    "swin3d": (dict(embed=32, heads=(2, 4, 8, 16), patch=(2, 2, 2), window=(3, 3, 3), t=16, h=24, w=20), (1, 1, 16, 24, 20)),
}


@pytest.mark.parametrize("name", list(SWIN_CASES))
def test_swin_encoder_drop_in_vs_golden(name):
    from long_context_biomedical_imaging_b200.backbone_swin import custom_Swin

    g = load_golden("encoders.npz")
    kw, in_shape = SWIN_CASES[name]
    model, chans = custom_Swin(_swin_cfg(**kw), in_shape[1])
    assert chans == [kw["embed"] * 2 ** i for i in range(5)]
    assert list(model.state_dict().keys()) == [str(k) for k in g[f"{name}/state_keys"]]
    assert [",".join(map(str, v.shape)) for v in model.state_dict().values()] == [str(s) for s in g[f"{name}/state_shapes"]]
    ao.fill_parameters_(model, 31)
    model = model.cuda()
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(*in_shape, generator=gen).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    assert len(outs) == int(g[f"{name}/n_out"]) == 6
    # the reference's `self.patch_embed(x)` seam (backbone_swin.py:885) returns MONAI's channel-first grid
    import torch.nn.functional as F
    xs = x.squeeze(2) if model.spatial_dims == 2 else x
    pe = model.patch_embed(xs)
    conv = F.conv2d if model.spatial_dims == 2 else F.conv3d
    pads = []
    for s_, p_ in zip(reversed(xs.shape[2:]), reversed(model.patch_embed.patch_size)):
        pads += [0, (p_ - s_ % p_) % p_]
    want_pe = conv(F.pad(xs, pads), model.patch_embed.proj.weight, model.patch_embed.proj.bias, stride=model.patch_embed.patch_size)
    assert tuple(pe.shape) == tuple(want_pe.shape) and pe.shape[1] == kw["embed"]
    assert max_rel(pe.detach().float().cpu(), want_pe.detach().float().cpu()) < 1e-3
    for i, o in enumerate(outs):
        want = g[f"{name}/out{i}"]
        assert tuple(o.shape) == want.shape, i
        assert max_rel(o.detach().float().cpu(), want) < 3e-2, (name, i)
    loss = sum((o.float() * torch.linspace(-1, 1, o.numel(), device="cuda").reshape(o.shape)).sum() for o in outs[1:])
    loss.backward()
    for pname, p in model.named_parameters():
        assert p.grad is not None, pname   # DDP(find_unused_parameters=False) needs every parameter to get a grad
        want_norm = float(g[f"{name}/gradnorm/{pname}"])
        got_norm = float(p.grad.double().norm())
        assert abs(got_norm - want_norm) <= 5e-2 * max(want_norm, 1e-3), (pname, got_norm, want_norm)


@pytest.mark.parametrize("grid,window,shift,C,H,B", [
    ((128, 128), (7, 7), (3, 3), 96, 3, 2),      # cfg2 stage 1 geometry (B reduced), d = 32
    ((16, 16), (7, 7), (3, 3), 768, 24, 2),      # cfg2 stage 4
    ((16, 16, 16), (7, 7, 7), (3, 3, 3), 192, 12, 1),   # cfg4 stage 3 ('unetr', d = 16)
    ((8, 8, 8), (7, 7, 7), (3, 3, 3), 384, 24, 1),      # cfg4 stage 4: 8^3 grid padded to 14^3
    ((16, 16, 16), (8, 8, 8), (4, 4, 4), 64, 2, 1),     # window 8^3 = 512 tokens (shipped 3-D scripts), d = 32
    # full BASELINE geometries (VERDICT r1 "untested geometries"): the fp32 oracle runs on the GPU (<= 10 GB)
    ((64, 64, 64), (7, 7, 7), (3, 3, 3), 48, 3, 1),     # cfg4 stage 1 'unetr': 1000 windows x 343, d = 16
    ((64, 64, 64), (7, 7, 7), (3, 3, 3), 96, 3, 1),     # cfg4 stage 1 'tiny', d = 32
    ((32, 32, 32), (7, 7, 7), (3, 3, 3), 96, 6, 1),     # cfg4 stage 2 'unetr': 125 windows
    ((32, 32, 32), (7, 7, 7), (3, 3, 3), 192, 6, 1),    # cfg4 stage 2 'tiny'
    ((64, 64, 64), (7, 7, 7), (0, 0, 0), 48, 3, 1),     # cfg4 stage 1 W-MSA block (no shift, no mask)
    ((64, 64), (7, 7), (3, 3), 192, 6, 16),             # cfg2 stage 2, batch 16
    ((32, 32), (7, 7), (3, 3), 384, 12, 16),            # cfg2 stage 3
    ((128, 128), (7, 7), (3, 3), 96, 3, 16),            # cfg2 stage 1 at the bench batch
])
@pytest.mark.parametrize("mode", ["tcgen05", "generic"])
def test_window_attention_op_vs_oracle_at_config_geometry(grid, window, shift, C, H, B, mode):
    """The raw fused op (no Linear layers) against the oracle's gather/attention/scatter on the GPU in fp32, once with
    the tcgen05 / TMA kernels forced wherever they apply (3-D windows of 128..512 tokens) and once with the generic
    kernels only (the default picks per shape by measured speed)."""
    from long_context_biomedical_imaging_b200 import ops

    if mode == "generic" and (len(grid) == 2 or B > 2):
        pytest.skip("2-D windows have one kernel family; covered by the tcgen05-mode run")
    ops.set_window_kernel_mode(mode)
    try:
        _window_attention_vs_oracle(grid, window, shift, C, H, B)
    finally:
        ops.set_window_kernel_mode("auto")


def _window_attention_vs_oracle(grid, window, shift, C, H, B):
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(0)
    d = C // H
    qkv = (torch.randn(B, *grid, 3 * C, device="cuda") * 0.7).to(torch.bfloat16).requires_grad_(True)
    bias = (torch.randn(3 * C, device="cuda") * 0.5).requires_grad_(True)
    rows = int(np.prod([2 * w - 1 for w in window]))
    table = (torch.randn(rows, H, device="cuda") * 0.5).requires_grad_(True)
    out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
    d_out = torch.randn_like(out)
    out.backward(d_out)
    torch.cuda.synchronize()

    # oracle: identity "Linear"s so that swin_part1 reduces to gather -> attention(+bias rows for pads) -> scatter.
    # The pad-token rows are produced by feeding x = qkv - bias through F.linear(I, bias).
    qkv_ref = qkv.detach().float().requires_grad_(True)
    bias_ref = bias.detach().clone().requires_grad_(True)
    table_ref = table.detach().clone().requires_grad_(True)
    win, sh = wm.resolve_window(grid, window, shift)
    n = int(np.prod(win))
    gmap = torch.from_numpy(wm.gather_map(grid, window, shift)).cuda()
    valid = gmap >= 0
    flat = qkv_ref.reshape(B, -1, 3 * C)
    xw = torch.where(valid[None, :, :, None], flat[:, gmap.clamp(min=0).reshape(-1)].reshape(B, *gmap.shape, 3 * C),
                     bias_ref.to(torch.bfloat16).float())
    mask = (torch.from_numpy(wm.shift_mask(grid, window, shift)).cuda() if any(s > 0 for s in sh) else None)
    index_nn = torch.from_numpy(wm.rel_pos_index_used(window, n)).cuda()
    # direct evaluation (window_attention applies a qkv Linear we do not want here)
    t = xw.reshape(-1, n, 3, H, d).permute(2, 0, 3, 1, 4)
    s = (t[0] * d ** -0.5) @ t[1].transpose(-2, -1)
    s = s + table_ref[index_nn.reshape(-1)].reshape(n, n, H).permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        s = (s.view(B, nw, H, n, n) + mask[None, :, None]).view(-1, H, n, n)
    yw = (s.softmax(-1) @ t[2]).transpose(1, 2).reshape(B, -1, C)
    ref = torch.zeros(B, flat.shape[1], C, device="cuda")
    sel = valid.reshape(-1)
    ref[:, gmap.reshape(-1)[sel]] = yw[:, sel]
    ref = ref.reshape(B, *grid, C)
    ref.backward(d_out.float())
    assert max_rel(out.detach().float().cpu(), ref.detach().cpu()) < BF16_TOL
    assert max_rel(qkv.grad.float().cpu(), qkv_ref.grad.cpu()) < BF16_TOL
    assert max_rel(table.grad.cpu(), table_ref.grad.cpu()) < BF16_TOL
    if bias_ref.grad is not None and float(bias_ref.grad.abs().max()) > 0:
        assert max_rel(bias.grad.cpu(), bias_ref.grad.cpu()) < BF16_TOL
    else:
        assert float(bias.grad.abs().max()) == 0.0


def test_swin_inference_mode_and_rejections():
    from long_context_biomedical_imaging_b200 import ops
    from long_context_biomedical_imaging_b200.backbone_swin import custom_Swin

    kw, in_shape = SWIN_CASES["swin2d"]
    model, _ = custom_Swin(_swin_cfg(**kw), in_shape[1])
    model = model.cuda().eval()
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(torch.randn(*in_shape, device="cuda"))
    assert len(outs) == 6 and outs[1].shape == (2, 16, 1, 24, 20)
    with pytest.raises(ValueError):   # head_dim 8 unsupported
        ops.window_attention(torch.randn(1, 8, 8, 3 * 16, device="cuda"), None, torch.zeros(169, 2, device="cuda"),
                             (8, 8), (7, 7), (0, 0), 2)


@pytest.mark.parametrize("case", [
    dict(B=2, grid=(14, 21), window=(7, 7), shift=(3, 3), heads=3, d=32),          # fused small-window kernels
    dict(B=1, grid=(9, 10, 8), window=(7, 7, 7), shift=(3, 3, 3), heads=2, d=16),  # 343-token windows: dkdv + dq kernels
])
def test_window_ranges_partition_the_block(case):
    """lcbi_win_attn_{fwd,bwd}_range: the results of a partition of the (batch, window) list are disjoint in out / dqkv
    and add up to the full-block result (what window_parallel.py relies on); each range also matches the oracle."""
    from long_context_biomedical_imaging_b200 import ops
    from long_context_biomedical_imaging_b200.window_parallel import count_windows, shard_range
    from oracle import attention_oracle as ao

    torch.manual_seed(7)
    B, grid, heads, d = case["B"], case["grid"], case["heads"], case["d"]
    C = heads * d
    qkv = (torch.randn(B, *grid, 3 * C, device="cuda") * 0.7).to(torch.bfloat16).requires_grad_(True)
    bias = torch.randn(3 * C, device="cuda").requires_grad_(True)
    n_tab = 1
    for w in case["window"]:
        n_tab *= 2 * w - 1
    table = (torch.randn(n_tab, heads, device="cuda") * 0.5).requires_grad_(True)
    d_out = torch.randn(B, *grid, C, device="cuda").to(torch.bfloat16)

    full = ops.window_attention(qkv, bias, table, grid, case["window"], case["shift"], heads)
    g_full = torch.autograd.grad(full, (qkv, bias, table), d_out)
    total = B * count_windows(grid, case["window"])
    out_sum, g_sum = torch.zeros_like(full, dtype=torch.float32), [torch.zeros_like(t, dtype=torch.float32) for t in g_full]
    for r in range(3):
        rng = shard_range(total, 3, r)
        part = ops.window_attention(qkv, bias, table, grid, case["window"], case["shift"], heads, win_range=rng)
        want = ao.window_attention_core(qkv.detach().float().cpu(), bias.detach().cpu(), table.detach().cpu(), grid,
                                        case["window"], case["shift"], heads, win_range=rng)
        assert max_rel(part.detach().float().cpu(), want) < BF16_TOL
        assert int(((part != 0) & (out_sum != 0)).sum()) == 0          # disjoint token rows
        out_sum += part.detach().float()
        for acc, g in zip(g_sum, torch.autograd.grad(part, (qkv, bias, table), d_out)):
            acc += g.float()
    assert torch.equal(out_sum.to(torch.bfloat16), full.detach())       # the same kernel wrote every row: bit-identical
    assert torch.equal(g_sum[0].to(torch.bfloat16), g_full[0])           # dqkv rows are disjoint too
    assert max_rel(g_sum[1].cpu(), g_full[1].float().cpu()) < 1e-3      # fp32 atomics: order differs
    assert max_rel(g_sum[2].cpu(), g_full[2].float().cpu()) < 1e-3


@pytest.mark.parametrize("case", [
    dict(B=16, grid=(128, 128), window=(7, 7), C=96, heads=3),          # cfg2 stage 1, bench size
    dict(B=1, grid=(64, 64, 64), window=(7, 7, 7), C=48, heads=3),      # cfg4 stage 1 ('unetr'): 1000 windows of 343
])
def test_window_attention_full_size_properties(case):
    """BASELINE full sizes, no n^2 oracle: when every value row (including the pad tokens' bias row) is the same
    vector, each softmax row sums to one and the output equals that vector at every token whatever the bias table and
    the shift mask do; consequently d(q), d(k) and d(table) vanish and d(v) sums to sum(dO) per window set."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(9)
    B, grid, heads, C = case["B"], case["grid"], case["heads"], case["C"]
    window = case["window"]
    shift = tuple(w // 2 for w in window)
    qkv = (torch.randn(B, *grid, 3 * C, device="cuda") * 0.7).to(torch.bfloat16)
    vrow = torch.randn(C, device="cuda").to(torch.bfloat16)
    qkv[..., 2 * C:] = vrow
    qkv.requires_grad_(True)
    bias = torch.randn(3 * C, device="cuda")
    bias[2 * C:] = vrow.float()
    bias.requires_grad_(True)
    n_tab = 1
    for w in window:
        n_tab *= 2 * w - 1
    table = (torch.randn(n_tab, heads, device="cuda") * 0.5).requires_grad_(True)
    out = ops.window_attention(qkv, bias, table, grid, window, shift, heads)
    assert max_rel(out.detach().float().cpu(), vrow.float().expand_as(out).cpu()) < 1e-2
    d_out = torch.randn_like(out)
    out.backward(d_out)
    g = qkv.grad.float()
    ref_mag = float(g[..., 2 * C:].abs().max())          # dv carries the whole gradient
    assert float(g[..., :2 * C].abs().max()) < 2e-2 * ref_mag
    assert float(table.grad.abs().max()) < 2e-2 * ref_mag * 50
    # sum over tokens of dv (+ the pad tokens' share in d(bias_v)) = sum over tokens of dO
    total_dv = g[..., 2 * C:].sum(dim=tuple(range(len(grid) + 1))) + bias.grad[2 * C:]
    total_do = d_out.float().sum(dim=tuple(range(len(grid) + 1)))
    assert max_rel(total_dv.cpu(), total_do.cpu()) < 2e-2


def test_row_gather_scatter_match_torch_indexing():
    """csrc/row_copy.cu (the window-sharded exchange) against index_select / index_copy_, incl. 96-byte rows (cfg4 'unetr'
    output rows) and the dump-row convention of window_parallel._exchange_disjoint_rows."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(4)
    for rows, feat, dtype in ((1000, 48, torch.bfloat16), (777, 144, torch.bfloat16), (64, 8, torch.float32)):
        src = torch.randn(rows, feat, device="cuda").to(dtype)
        ids = torch.randint(0, rows, (rows + 13,), device="cuda")
        assert torch.equal(ops.gather_rows(src, ids), src.index_select(0, ids))
        perm = torch.randperm(rows, device="cuda")
        dump = torch.cat([perm, torch.full((5,), rows, device="cuda", dtype=torch.int64)])      # 5 rows into the dump row
        packed = torch.randn(rows + 5, feat, device="cuda").to(dtype)
        got = ops.scatter_rows(packed, dump, rows + 1)[:rows]
        want = torch.empty(rows, feat, device="cuda", dtype=dtype)
        want.index_copy_(0, perm, packed[:rows])
        assert torch.equal(got, want)
    with pytest.raises(ValueError):
        ops.gather_rows(torch.randn(4, 3, device="cuda"), torch.zeros(2, dtype=torch.int64, device="cuda"))   # 12-byte rows
