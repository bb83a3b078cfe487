import sys

import torch
sys.path.insert(0, "/root/repo")
from long_context_biomedical_imaging_b200 import ops
def t(B, N, kind="fwd"):
    qkv = torch.randn(B, N, 3, 12, 64, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    d_o = torch.randn_like(o)
    fn = (lambda: ops.dense_attn_fwd(q, k, v, 0.125, out=o)) if kind == "fwd" else (lambda: ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
for kind in sys.argv[1:] or ["fwd"]:
    for B, N in ((16, 1728), (8, 3456), (4, 6912), (32, 864), (64, 432), (16, 1664), (16, 1792), (18, 1728), (21, 1728)):
        print(kind, B, N, f"{t(B, N, kind):.4f} ms", flush=True)
