"""Swin window-attention op throughput at the cfg2 / cfg4 stage geometries (fwd and fwd+bwd), CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import _lib
if os.environ.get('LCBI_LIB'):
    _lib.LIB_PATH = os.environ['LCBI_LIB']   # timing experiments with variant builds
from long_context_biomedical_imaging_b200 import ops  # noqa: E402

CASES = [
    ("cfg2 st1", 16, (128, 128), (7, 7), 96, 3),
    ("cfg2 st2", 16, (64, 64), (7, 7), 192, 6),
    ("cfg2 st3", 16, (32, 32), (7, 7), 384, 12),
    ("cfg2 st4", 16, (16, 16), (7, 7), 768, 24),
    ("cfg4 st1 unetr", 1, (64, 64, 64), (7, 7, 7), 48, 3),
    ("cfg4 st2 unetr", 1, (32, 32, 32), (7, 7, 7), 96, 6),
    ("cfg4 st3 unetr", 1, (16, 16, 16), (7, 7, 7), 192, 12),
    ("cfg4 st4 unetr", 1, (8, 8, 8), (7, 7, 7), 384, 24),
    ("cfg4 st1 tiny", 1, (64, 64, 64), (7, 7, 7), 96, 3),
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


ONLY = os.environ.get("SWIN_ONLY")
for name, B, grid, window, C, H in CASES:
    if ONLY and ONLY not in name:
        continue
    for shifted in ((True,) if ONLY else (False, True)):
        shift = tuple(w // 2 if shifted else 0 for w in window)
        torch.manual_seed(0)
        qkv = torch.randn(B, *grid, 3 * C, device="cuda").to(torch.bfloat16).requires_grad_(True)
        bias = torch.randn(3 * C, device="cuda", requires_grad=True)
        rows = int(np.prod([2 * w - 1 for w in window]))
        table = torch.randn(rows, H, device="cuda", requires_grad=True)
        d_out = torch.randn(B, *grid, C, device="cuda").to(torch.bfloat16)

        def fwd():
            with torch.no_grad():
                return ops.window_attention(qkv, bias, table, grid, window, shift, H)

        def fwdbwd():
            out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
            out.backward(d_out)
            qkv.grad = None

        t_f = timeit(fwd)
        t_fb = timeit(fwdbwd)
        T = B * int(np.prod(grid))
        bytes_alg = 24 * C * T          # fwd+bwd bytes per layer-token
        print(f"{name:16s} shift={int(shifted)} fwd {t_f * 1e3:8.1f} us  fwd+bwd {t_fb * 1e3:8.1f} us  "
              f"{T / t_fb / 1e3:8.2f} Mtok/s  {bytes_alg / t_fb / 1e6:7.1f} GB/s(alg)", flush=True)
