"""Timing of the per-rank kernel calls of one ring step at cfg5 on 8 GPUs (Nq = Nk = 32768, B = 1, accumulate modes),
on ONE GPU: isolates kernel speed at that shape from communication effects. LCBI_LIB selects a variant build."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import _lib
if os.environ.get("LCBI_LIB"):
    _lib.LIB_PATH = os.environ["LCBI_LIB"]
from long_context_biomedical_imaging_b200 import ops  # noqa: E402

N, H, d = int(os.environ.get("RS_N", 32768)), 12, 64
torch.manual_seed(0)
q, k, v, d_o = [torch.randn(1, N, H, d, device="cuda").to(torch.bfloat16) for _ in range(4)]
o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
dq, dk, dv = [torch.zeros(1, N, H, d, device="cuda") for _ in range(3)]


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(os.path.basename(_lib.LIB_PATH), f"N={N}",
      f"fwd {timeit(lambda: ops.dense_attn_fwd(q, k, v, 0.125)):.3f} ms",
      f"bwd(accumulate) {timeit(lambda: ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125, dq=dq, dk=dk, dv=dv, accumulate_dkv=True, accumulate_dq=True)):.3f} ms",
      f"bwd(plain) {timeit(lambda: ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)):.3f} ms", flush=True)
