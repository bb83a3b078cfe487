"""GPU parity of the kernels around the attention core — LayerNorm (lcbi_layer_norm_fwd / _bwd through ops.layer_norm),
the residual add fused into it (lcbi_add_layer_norm_*, ops.add_layer_norm) and the bias gradient of the projections
(lcbi_bias_grad behind ops.linear). LayerNorm is checked against the CPU oracle's
written-out LayerNorm (oracle.attention_oracle.layer_norm_rows, pinned against torch.nn.LayerNorm in
tests/test_oracle_golden.py): the norm1 / norm2 of the encoder blocks (reference backbone_vit.py:260-263,
backbone_swin.py:437,489).

Tolerances (max-rel = max|a-b| / max|b|): fp32 in / fp32 out 1e-5 (budget of the fp32-accumulate path: 1e-4);
bf16 output or bf16 input 1e-2 (budget of the bf16 path: 2e-2; one bf16 rounding is 3.9e-3)."""
import pytest
import torch

from conftest import max_rel
from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu


def _case(rows_shape, C, x_dtype, out_dtype, affine, seed=0):
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(seed)
    x = (torch.randn(*rows_shape, C) * 1.7 + 0.3).to(x_dtype)
    w = (1 + 0.3 * torch.randn(C)) if affine else None
    b = (0.3 * torch.randn(C)) if affine else None
    g = torch.randn(*rows_shape, C).to(out_dtype)

    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True) if affine else None
    br = b.double().requires_grad_(True) if affine else None
    yr = ao.layer_norm_rows(xr, wr, br, 1e-5)
    ref_grads = torch.autograd.grad(yr, (xr, wr, br) if affine else (xr,), g.double())

    xd = x.cuda().requires_grad_(True)
    wd = w.cuda().requires_grad_(True) if affine else None
    bd = b.cuda().requires_grad_(True) if affine else None
    y = ops.layer_norm(xd, wd, bd, 1e-5, out_dtype=out_dtype)
    assert y.dtype == out_dtype and y.shape == x.shape
    grads = torch.autograd.grad(y, (xd, wd, bd) if affine else (xd,), g.cuda())
    tol = 1e-5 if (x_dtype == torch.float32 and out_dtype == torch.float32) else 1e-2
    assert max_rel(y.detach().float().cpu(), yr.detach()) < tol
    assert grads[0].dtype == x_dtype
    for got, ref, name in zip(grads, ref_grads, ("dx", "dgamma", "dbeta")):
        assert max_rel(got.float().cpu(), ref) < tol, (name, max_rel(got.float().cpu(), ref))


@pytest.mark.parametrize("rows_shape,C", [((1,), 4), ((77,), 48), ((2, 197), 192), ((3, 5, 7), 96), ((1, 1728), 768),
                                          ((300,), 1536), ((9,), 3072), ((33,), 32), ((130,), 64), ((1000,), 128), ((5000,), 132)])
@pytest.mark.parametrize("affine", [True, False])
def test_layer_norm_fp32_vs_oracle(rows_shape, C, affine):
    _case(rows_shape, C, torch.float32, torch.float32, affine)


@pytest.mark.parametrize("x_dtype,out_dtype", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                               (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("rows_shape,C", [((2, 197), 192), ((1, 1728), 768), ((64,), 1536)])
def test_layer_norm_mixed_dtypes_vs_oracle(rows_shape, C, x_dtype, out_dtype):
    _case(rows_shape, C, x_dtype, out_dtype, True, seed=1)


def test_layer_norm_rejects_what_the_kernel_does_not_take():
    from long_context_biomedical_imaging_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layer_norm(torch.randn(4, 8), None, None)
    with pytest.raises(ValueError, match="multiple of 4"):
        ops.layer_norm(torch.randn(4, 6, device="cuda"), None, None)


def test_empty_rows_are_not_an_error():
    from long_context_biomedical_imaging_b200 import ops

    x = torch.zeros(0, 96, device="cuda", requires_grad=True)
    w = torch.ones(96, device="cuda", requires_grad=True)
    b = torch.zeros(96, device="cuda", requires_grad=True)
    y = ops.layer_norm(x, w, b)
    assert y.shape == (0, 96)
    z = ops.linear(y, torch.randn(8, 96, device="cuda", requires_grad=True), torch.zeros(8, device="cuda", requires_grad=True))
    z.sum().backward()
    assert x.grad.shape == (0, 96) and float(w.grad.abs().max()) == 0.0
    assert float(ops.bias_grad(torch.zeros(0, 8, device="cuda")).abs().max()) == 0.0


def test_layer_norm_under_autocast_feeds_the_linear_in_bf16_and_matches_torch():
    """Inside bf16 autocast the op emits bf16 — the rounding torch's autocast applies to its fp32 LayerNorm output in
    front of a Linear — so block-level results agree with nn.LayerNorm -> nn.Linear to bf16 rounding."""
    from long_context_biomedical_imaging_b200.blocks import apply_layer_norm

    torch.manual_seed(2)
    ln = torch.nn.LayerNorm(192).cuda()
    lin = torch.nn.Linear(192, 64).cuda()
    x = torch.randn(2, 50, 192, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = apply_layer_norm(ln, x)
        assert y.dtype == torch.bfloat16
        ours = lin(y)
        theirs = lin(ln(x))
        hidden = apply_layer_norm(ln, x, out_dtype=x.dtype)
    assert hidden.dtype == torch.float32
    assert max_rel(ours.detach().float().cpu(), theirs.detach().float().cpu()) < 1e-2
    assert max_rel(hidden.detach().cpu(), ln(x).detach().cpu()) < 1e-5


def test_layer_norm_full_size_properties():
    """cfg3 encoder size (16 volumes x 1728 tokens x 768 channels): rows come out with zero mean / unit variance
    (no affine), the backward is linear in dy, d(beta) is the column sum of dy, and dx is orthogonal to the
    constant vector and to xhat row by row."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(4)
    rows, C = 16 * 1728, 768
    x = (torch.randn(rows, C, device="cuda") * 3 + 1).requires_grad_(True)
    y = ops.layer_norm(x, None, None, 1e-5, out_dtype=torch.float32)
    assert float(y.mean(-1).abs().max()) < 1e-5
    assert float((y.var(-1, unbiased=False) - 1).abs().max()) < 1e-4
    w = (1 + 0.2 * torch.randn(C, device="cuda")).requires_grad_(True)
    b = torch.zeros(C, device="cuda", requires_grad=True)
    g1, g2 = torch.randn(rows, C, device="cuda"), torch.randn(rows, C, device="cuda")
    out = ops.layer_norm(x, w, b, 1e-5, out_dtype=torch.float32)
    d1 = torch.autograd.grad(out, (x, w, b), g1, retain_graph=True)
    d2 = torch.autograd.grad(out, (x, w, b), g2, retain_graph=True)
    d12 = torch.autograd.grad(out, (x, w, b), 2 * g1 - 3 * g2)
    for a, c, both in zip(d1, d2, d12):
        assert max_rel((2 * a - 3 * c).cpu(), both.cpu()) < 1e-5
    assert max_rel(d1[2].cpu(), g1.sum(0).cpu()) < 1e-5
    assert float(d1[0].sum(-1).abs().max()) < 1e-3 * float(d1[0].abs().max()) * C ** 0.5
    assert float((d1[0] * y.detach()).sum(-1).abs().max()) < 1e-3 * float(d1[0].abs().max()) * C ** 0.5


@pytest.mark.parametrize("rows_shape,cin,cout", [((3, 50), 96, 288), ((1, 1728), 768, 768), ((700,), 48, 144),
                                                 ((5, 7, 9), 64, 4), ((2049,), 32, 520)])
@pytest.mark.parametrize("autocast", [True, False])
def test_linear_with_bias_matches_nn_linear(rows_shape, cin, cout, autocast):
    """ops.linear = nn.Linear's GEMMs (same dtype flow) + lcbi_bias_grad for db: outputs and the three gradients
    against torch.nn.functional.linear on the same tensors, same autocast state. bf16: 1e-2 max-rel (the bias
    gradient is summed in fp32 here, in bf16-rounded form by autograd); fp32: 1e-5."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(5)
    x = torch.randn(*rows_shape, cin, device="cuda", requires_grad=True)
    w = (torch.randn(cout, cin, device="cuda") * cin ** -0.5).requires_grad_(True)
    b = torch.randn(cout, device="cuda", requires_grad=True)
    g = torch.randn(*rows_shape, cout, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        y = ops.linear(x, w, b)
        yr = torch.nn.functional.linear(x, w, b)
    assert y.dtype == yr.dtype
    grads = torch.autograd.grad(y, (x, w, b), g.to(y.dtype))
    ref = torch.autograd.grad(yr, (x, w, b), g.to(yr.dtype))
    tol = 1e-2 if autocast else 1e-5
    assert max_rel(y.detach().float().cpu(), yr.detach().float().cpu()) < tol
    for got, want, name in zip(grads, ref, ("dx", "dw", "db")):
        assert got.dtype == want.dtype, name
        assert max_rel(got.float().cpu(), want.float().cpu()) < tol, (name, max_rel(got.float().cpu(), want.float().cpu()))
    # the bias gradient against an fp64 column sum of exactly the dy the kernel saw
    exact = g.to(y.dtype).double().reshape(-1, cout).sum(0)
    assert max_rel(grads[2].double().cpu(), exact.cpu()) < 1e-5


def test_bias_grad_full_size():
    """Swin cfg2 stage-1 qkv gradient size (16 x 128 x 128 tokens x 288) in bf16, and a ViT cfg3 MLP size."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(6)
    for rows, C in [(16 * 128 * 128, 288), (16 * 1728, 3072)]:
        dy = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
        got = ops.bias_grad(dy)
        want = dy.double().sum(0)
        assert max_rel(got.double().cpu(), want.cpu()) < 1e-5
        assert torch.equal(got, ops.bias_grad(dy))   # deterministic


@pytest.mark.parametrize("x_dtype,d_dtype,out_dtype", [(torch.float32, torch.bfloat16, torch.bfloat16),
                                                       (torch.float32, torch.float32, torch.float32),
                                                       (torch.bfloat16, torch.bfloat16, torch.bfloat16),
                                                       (torch.float32, torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("rows_shape,C", [((2, 197), 128), ((3, 50), 192), ((1, 1728), 768), ((70,), 1536), ((40,), 96)])
def test_fused_residual_add_layer_norm_vs_oracle(rows_shape, C, x_dtype, d_dtype, out_dtype):
    """(x, delta) -> (x + delta, LayerNorm(x + delta)) — the residual add of a block fused into the norm that follows
    (reference backbone_vit.py:261-262). Both outputs feed the loss, as in the encoder (hidden state + next branch).
    C = 96 takes the unfused fallback (narrow rows) and must agree all the same."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(7)
    x = (torch.randn(*rows_shape, C) * 1.5 + 0.2).to(x_dtype)
    d = torch.randn(*rows_shape, C).to(d_dtype)
    w, b = 1 + 0.3 * torch.randn(C), 0.3 * torch.randn(C)
    g_sum = torch.randn(*rows_shape, C).to(x_dtype)
    g_y = torch.randn(*rows_shape, C).to(out_dtype)

    xr, dr, wr, br = (t.double().requires_grad_(True) for t in (x, d, w, b))
    # value rounded like the stored sum, gradient of the plain add
    sum_r = (xr + dr).detach().to(x_dtype).double() + ((xr + dr) - (xr + dr).detach())
    y_r = ao.layer_norm_rows(sum_r, wr, br, 1e-5)
    ref = torch.autograd.grad((sum_r, y_r), (xr, dr, wr, br), (g_sum.double(), g_y.double()))

    xd, dd = x.cuda().requires_grad_(True), d.cuda().requires_grad_(True)
    wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    xsum, y = ops.add_layer_norm(xd, dd, wd, bd, 1e-5, out_dtype=out_dtype)
    assert xsum.dtype == x_dtype and y.dtype == out_dtype
    got = torch.autograd.grad((xsum, y), (xd, dd, wd, bd), (g_sum.cuda(), g_y.cuda()))
    exact = x_dtype == torch.float32 and d_dtype == torch.float32 and out_dtype == torch.float32
    tol = 1e-5 if exact else 1e-2
    assert torch.equal(xsum.detach().cpu(), (x.cuda() + d.cuda()).to(x_dtype).cpu())      # the add itself is bit-exact
    assert max_rel(y.detach().float().cpu(), y_r.detach()) < tol
    assert got[0].dtype == x_dtype and got[1].dtype == d_dtype
    for a, r, name in zip(got, ref, ("dx", "ddelta", "dgamma", "dbeta")):
        assert max_rel(a.float().cpu(), r) < tol, (name, max_rel(a.float().cpu(), r))


def test_vit_block_and_encoder_agree_between_fused_and_per_block_paths():
    """ViT_with_alt_ops.forward folds every residual add into the following norm across block boundaries; calling the
    blocks one by one (TransformerBlock.forward, the per-block seam) must give the same hidden states."""
    import types

    from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT

    cfg = types.SimpleNamespace(ViT=types.SimpleNamespace(size="custom", hidden_size=192, mlp_dim=384, num_layers=3,
                                                          num_heads=3, patch_size=[1, 8, 8], use_hyena=False,
                                                          use_mamba=False), time=1, height=32, width=48, task_type="seg")
    model, _ = custom_ViT(cfg, 1)
    ao.fill_parameters_(model, 13)
    model = model.cuda()
    x = torch.randn(2, 1, 1, 32, 48, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
        t = model.patch_embedding(x.squeeze(2))
        per_block = []
        for blk in model.blocks:
            t = blk(t)
            per_block.append(t)
    assert len(outs) == 1 + 3 + 1
    for got, want in zip(outs[1:-1], per_block):
        assert got.dtype == want.dtype
        assert max_rel(got.detach().float().cpu(), want.detach().float().cpu()) < 1e-2
