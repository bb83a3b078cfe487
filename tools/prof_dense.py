"""Minimal driver for ncu: cfg3 (B=16, N=1728, H=12, d=64) dense attention fwd + bwd, a few iterations."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import ops  # noqa: E402

B, N, H, d = int(os.environ.get("PROF_B", 16)), 1728, 12, 64
torch.manual_seed(0)
qkv = torch.randn(B, N, 3, H, d, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
d_o = torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16)
for _ in range(int(os.environ.get("PROF_ITERS", 3))):
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
torch.cuda.synchronize()
print("done")
