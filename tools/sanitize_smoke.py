"""Smallest shape of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck) runs:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Each op runs forward and backward once; nothing is timed or checked numerically here (the parity tests do that)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "dense"):
    # dense attention fwd / bwd (+ state-carrying forward, merge): N = 200 (ragged tile), 2 heads
    q, k, v = [torch.randn(1, 200, 2, 64, device=dev).to(torch.bfloat16) for _ in range(3)]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    ops.dense_attn_bwd(q, k, v, o, torch.randn_like(o), lse, 0.125)
    st = (torch.empty(1, 200, 2, 64, device=dev), torch.empty(1, 2, 200, device=dev), torch.empty(1, 2, 200, device=dev))
    ops.dense_attn_fwd_state(q, k[:, :100], v[:, :100], 0.125, st, True, False)
    ops.dense_attn_fwd_state(q, k[:, 100:], v[:, 100:], 0.125, st, False, True)
    acc, la = torch.empty(1, 200, 2, 64, device=dev), torch.empty(1, 2, 200, device=dev)
    ops.attn_merge(acc, la, o, lse, True)
    ops.attn_merge(acc, la, o, lse, False)
    print("dense ok")

if which in ("all", "swin"):
    # window attention: 2-D small windows (fused small kernels), 3-D 343-token windows with padding + shift (tcgen05
    # kernels incl. the gather4 path), and a clamped 3-D window (generic mma.sync kernels)
    for grid, window, C, H in (((9, 10), (7, 7), 64, 2), ((9, 10, 8), (7, 7, 7), 32, 2), ((4, 9, 9), (7, 7, 7), 32, 2)):
        shift = tuple(w // 2 for w in window)
        qkv = torch.randn(1, *grid, 3 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
        bias = torch.randn(3 * C, device=dev, requires_grad=True)
        rows = 1
        for w in window:
            rows *= 2 * w - 1
        table = torch.randn(rows, H, device=dev, requires_grad=True)
        out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
        out.backward(torch.randn_like(out))
    print("swin ok")

if which in ("all", "misc"):
    # patch embedding (fp32 CUDA-core path and the tensor-core path), LayerNorm, bias gradient
    for shape, patch, n in (((1, 1, 16, 16), (2, 2), 32), ((1, 1, 16, 16, 16), (8, 8, 8), 128)):
        img = torch.randn(*shape, device=dev, requires_grad=True)
        w = torch.randn(n, 1, *patch, device=dev, requires_grad=True)
        b = torch.randn(n, device=dev, requires_grad=True)
        grid = tuple(s // p for s, p in zip(shape[2:], patch))
        npos = 1
        for g in grid:
            npos *= g
        pos = torch.randn(1, npos, n, device=dev, requires_grad=True)
        ops.patch_embed(img, w, b, pos, grid).sum().backward()
    x = torch.randn(70, 192, device=dev, requires_grad=True)
    g, be = torch.randn(192, device=dev, requires_grad=True), torch.randn(192, device=dev, requires_grad=True)
    ops.layer_norm(x, g, be).sum().backward()
    xs, y = ops.add_layer_norm(x, torch.randn_like(x), g, be)
    (xs.sum() + y.sum()).backward()
    ops.bias_grad(torch.randn(70, 192, device=dev))
    print("misc ok")
torch.cuda.synchronize()
print("sanitize smoke done")
