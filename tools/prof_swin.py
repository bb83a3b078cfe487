import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import ops
B, grid, window, C, H = 16, (128, 128), (7, 7), 96, 3
if os.environ.get("PROF_CFG4"):
    B, grid, window, C, H = 1, (64, 64, 64), (7, 7, 7), 48, 3
shift = tuple(w // 2 for w in window)
qkv = torch.randn(B, *grid, 3 * C, device="cuda").to(torch.bfloat16).requires_grad_(True)
bias = torch.randn(3 * C, device="cuda", requires_grad=True)
table = torch.randn(int(np.prod([2 * w - 1 for w in window])), H, device="cuda", requires_grad=True)
d_out = torch.randn(B, *grid, C, device="cuda").to(torch.bfloat16)
for _ in range(3):
    out = ops.window_attention(qkv, bias, table, grid, window, shift, H)
    out.backward(d_out)
torch.cuda.synchronize()
print("done")
