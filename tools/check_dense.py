"""GPU smoke/diagnostic for the dense attention kernels: parity vs fp32 torch + CUDA-event timing."""
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import _lib  # noqa: E402
if os.environ.get("LCBI_LIB_PATH"):          # an experimental build of the library (tools/ablate_dense.py build)
    _lib.LIB_PATH = os.environ["LCBI_LIB_PATH"]
from long_context_biomedical_imaging_b200 import ops  # noqa: E402


def ref(q, k, v, scale):
    qf, kf, vf = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    s = torch.einsum("bhxd,bhyd->bhxy", qf, kf) * scale
    p = s.softmax(-1)
    o = torch.einsum("bhxy,bhyd->bhxd", p, vf)
    return o.permute(0, 2, 1, 3), torch.logsumexp(s, -1)


def maxrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def run(B, H, N, do_bwd=True, time_it=False):
    torch.manual_seed(0)
    d = 64
    qkv = torch.randn(B, N, 3, H, d, device="cuda", dtype=torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    scale = d ** -0.5
    o, lse = ops.dense_attn_fwd(q, k, v, scale)
    torch.cuda.synchronize()
    msg = f"B={B} H={H} N={N}: "
    if N <= 4096:
        qr = qkv.float().requires_grad_(True)
        o_ref, lse_ref = ref(qr[:, :, 0], qr[:, :, 1], qr[:, :, 2], scale)
        msg += f"fwd o {maxrel(o, o_ref):.2e} lse {maxrel(lse, lse_ref):.2e} "
        if do_bwd:
            d_o = torch.randn_like(o)
            dq, dk, dv = ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale)
            torch.cuda.synchronize()
            (g,) = torch.autograd.grad(o_ref, qr, d_o.float())
            msg += f"| dq {maxrel(dq, g[:, :, 0]):.2e} dk {maxrel(dk, g[:, :, 1]):.2e} dv {maxrel(dv, g[:, :, 2]):.2e} "
            msg += f"nan={bool(torch.isnan(dq.float()).any() or torch.isnan(dk.float()).any() or torch.isnan(dv.float()).any())}"
    if time_it:
        d_o = torch.randn_like(o)
        for name, fn in (("fwd", lambda: ops.dense_attn_fwd(q, k, v, scale, out=o)),
                         ("bwd", lambda: ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale))):
            if name == "bwd" and not do_bwd:
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 30
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            flops = (4 if name == "fwd" else 8) * B * H * N * N * d
            msg += f"| {name} {ms:.3f} ms {flops / ms / 1e9:.0f} TF/s(alg) "
    print(msg, flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    do_bwd = which != "fwd"
    t0 = time.time()
    run(1, 1, 128, do_bwd)
    run(1, 2, 256, do_bwd)
    run(2, 3, 196, do_bwd)
    run(2, 3, 197, do_bwd)
    run(1, 2, 1000, do_bwd)
    run(1, 12, 1728, do_bwd, time_it=True)
    run(16, 12, 1728, do_bwd=do_bwd, time_it=True)
    run(1, 12, 8192, do_bwd=do_bwd, time_it=True)
    print("elapsed", time.time() - t0, flush=True)
