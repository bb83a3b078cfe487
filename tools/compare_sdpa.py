"""Context only (not a contract number): torch SDPA (library kernels) and the reference's unfused eager attention
at the cfg3 shape, next to the lcbi_b200 kernels. CUDA events, fwd+bwd."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import ops  # noqa: E402

B, N, H, d = 16, 1728, 12, 64
torch.manual_seed(0)
qkv = torch.randn(B, N, 3, H, d, device="cuda").to(torch.bfloat16)
d_o = torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16)


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ours():
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)


def sdpa(backend):
    from torch.nn.attention import SDPBackend, sdpa_kernel

    x = qkv.detach().clone().requires_grad_(True)

    def run():
        q, k, v = [x[:, :, i].transpose(1, 2) for i in range(3)]
        with sdpa_kernel(backend):
            o = F.scaled_dot_product_attention(q, k, v)
        o.backward(d_o.transpose(1, 2))
        x.grad = None
    return run


def eager():
    x = qkv.detach().clone().requires_grad_(True)

    def run():
        q, k, v = [x[:, :, i].transpose(1, 2) for i in range(3)]
        att = (torch.einsum("blxd,blyd->blxy", q, k) * 0.125).softmax(dim=-1)
        o = torch.einsum("bhxy,bhyd->bhxd", att, v)
        o.backward(d_o.transpose(1, 2))
        x.grad = None
    return run


flops = 12.0 * B * N * N * H * d
print(f"lcbi_b200            {timeit(ours):8.3f} ms")
from torch.nn.attention import SDPBackend
for name, be in (("sdpa flash", SDPBackend.FLASH_ATTENTION), ("sdpa cudnn", SDPBackend.CUDNN_ATTENTION),
                 ("sdpa efficient", SDPBackend.EFFICIENT_ATTENTION)):
    try:
        print(f"{name:20s} {timeit(sdpa(be)):8.3f} ms")
    except Exception as e:  # backend unavailable
        print(f"{name:20s} unavailable: {str(e)[:80]}")
print(f"reference eager bf16 {timeit(eager(), 5):8.3f} ms   (materialises the B*H*N*N matrix, backbone_vit.py:193)")
print(f"algorithmic FLOPs per step {flops / 1e9:.1f} GF")
