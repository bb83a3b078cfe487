"""Patch-embedding throughput at the BASELINE.json config shapes (fwd and fwd+bwd), CUDA events, vs HBM bytes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import ops  # noqa: E402

CASES = [
    ("cfg1 ViT 224^2 p16 h192", 2, 1, (224, 224), (16, 16), 192, True),
    ("cfg2 Swin 512^2 p4 c96", 16, 1, (512, 512), (4, 4), 96, False),
    ("cfg3 ViT 96^3 p8 h768", 16, 1, (96, 96, 96), (8, 8, 8), 768, True),
    ("cfg4 Swin 128^3 p2 c48", 1, 1, (128, 128, 128), (2, 2, 2), 48, False),
    ("cfg5 ViT 1024^2 p2 h768", 1, 1, (1024, 1024), (2, 2), 768, True),
]


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""

for name, B, Cin, img, patch, N, vit in CASES:
    if ONLY and not name.startswith(ONLY):
        continue
    torch.manual_seed(0)
    x = torch.randn(B, Cin, *img, device="cuda")
    w = torch.randn(N, Cin, *patch, device="cuda", requires_grad=True)
    b = torch.randn(N, device="cuda", requires_grad=True)
    grid = [s // p for s, p in zip(img, patch)]
    Np = int(np.prod(grid))
    pos = torch.randn(1, Np, N, device="cuda", requires_grad=True) if vit else None
    out_dtype = torch.float32 if vit else torch.bfloat16
    d_out = torch.randn(B, Np, N, device="cuda", dtype=out_dtype)

    def fwd():
        with torch.no_grad():
            return ops.patch_embed(x, w, b, pos, grid, out_dtype)

    def fwdbwd():
        y = ops.patch_embed(x, w, b, pos, grid, out_dtype)
        y.backward(d_out)
        w.grad = None
        b.grad = None
        if pos is not None:
            pos.grad = None

    def torch_ref():
        conv = torch.nn.functional.conv3d if len(img) == 3 else torch.nn.functional.conv2d
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            y = conv(x, w, b, stride=patch).flatten(2).transpose(-1, -2)
            return y + pos if pos is not None else y.contiguous()

    t_f, t_fb, t_ref = timeit(fwd), timeit(fwdbwd), timeit(torch_ref)
    out_bytes = B * Np * N * (4 if vit else 2)
    in_bytes = x.numel() * 4 + w.numel() * 4 + (Np * N * 4 if vit else 0)
    print(f"{name:26s} fwd {t_f * 1e3:8.1f} us ({(in_bytes + out_bytes) / t_f / 1e6:7.1f} GB/s)  fwd+bwd {t_fb * 1e3:8.1f} us"
          f"  | torch conv(bf16 autocast)+transpose(+pos) fwd {t_ref * 1e3:8.1f} us", flush=True)
