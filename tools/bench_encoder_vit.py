"""Context only (not a contract number): whole ViT encoder fwd+bwd, drop-in modules vs a plain-PyTorch encoder with the
reference's arithmetic (strided conv patch embedding, unfused einsum -> softmax -> einsum attention as in
backbone_vit.py:191-201, LayerNorm / MLP identical), both under bf16 autocast on the same parameters.

    python tools/bench_encoder_vit.py            # cfg3: ViT-B, 96^3 volume, patch 8, batch 4
"""
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT  # noqa: E402


class EagerBlock(nn.Module):
    def __init__(self, c, mlp, heads):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(c), nn.LayerNorm(c)
        self.qkv, self.out_proj = nn.Linear(c, 3 * c, bias=False), nn.Linear(c, c)
        self.linear1, self.linear2 = nn.Linear(c, mlp), nn.Linear(mlp, c)
        self.h, self.scale = heads, (c // heads) ** -0.5

    def forward(self, x, mode):
        B, N, C = x.shape
        qkv = self.qkv(self.norm1(x)).view(B, N, 3, self.h, C // self.h).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        if mode == "eager":
            att = (torch.einsum("blxd,blyd->blxy", q, k) * self.scale).softmax(dim=-1)
            o = torch.einsum("bhxy,bhyd->bhxd", att.to(v.dtype), v)
        else:
            o = F.scaled_dot_product_attention(q, k, v)
        x = x + self.out_proj(o.transpose(1, 2).reshape(B, N, C))
        return x + self.linear2(F.gelu(self.linear1(self.norm2(x))))


class EagerViT(nn.Module):
    def __init__(self, cin, patch, grid, c, mlp, layers, heads):
        super().__init__()
        self.conv = nn.Conv3d(cin, c, patch, patch)
        self.pos = nn.Parameter(torch.zeros(1, grid, c))
        self.blocks = nn.ModuleList(EagerBlock(c, mlp, heads) for _ in range(layers))
        self.norm = nn.LayerNorm(c)

    def forward(self, x, mode):
        h = self.conv(x).flatten(2).transpose(1, 2) + self.pos
        for b in self.blocks:
            h = b(h, mode)
        return self.norm(h)


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


if __name__ == "__main__":
    B = int(os.environ.get("ENC_B", 4))
    T = Hh = W = 96
    patch = (8, 8, 8)
    cfg = types.SimpleNamespace(ViT=types.SimpleNamespace(size="base", patch_size=list(patch), use_hyena=False, use_mamba=False),
                                time=T, height=Hh, width=W, task_type="seg")
    torch.manual_seed(0)
    ours, _ = custom_ViT(cfg, 1)
    ours = ours.cuda()
    ref = EagerViT(1, patch, (T // 8) * (Hh // 8) * (W // 8), 768, 3072, 12, 12).cuda()
    x = torch.randn(B, 1, T, Hh, W, device="cuda")

    def run_ours():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ours(x)[-1]
        out.float().square().mean().backward()
        ours.zero_grad(set_to_none=True)

    def run_ref(mode):
        def f():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = ref(x, mode)
            out.float().square().mean().backward()
            ref.zero_grad(set_to_none=True)
        return f

    print(f"ViT-B encoder, {T}^3 volume, patch 8 -> 1728 tokens, batch {B}, bf16 autocast, fwd+bwd")
    print(f"  drop-in modules (lcbi_b200 kernels)            {timeit(run_ours):8.2f} ms")
    print(f"  plain PyTorch, SDPA attention (library kernels) {timeit(run_ref('sdpa')):8.2f} ms")
    try:
        print(f"  plain PyTorch, reference's unfused attention   {timeit(run_ref('eager')):8.2f} ms")
    except torch.OutOfMemoryError:
        print("  plain PyTorch, reference's unfused attention   out of memory")
