"""Parity + timing of experimental builds of the dense backward, all in one process on the GPU box.

    python tools/try_variants.py build            # here (no GPU): one .so per entry of VARIANTS
    python tools/try_variants.py [name ...]       # on the GPU box

Every variant is checked against an fp32 torch reference at a few shapes (ragged lengths, cross lengths, the fp32
accumulate modes the ring uses) and timed at cfg3 (16 volumes x 12 heads x 1728 tokens), the backward launch group
alone (prep + main + finish), hot L2, CUDA events.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "long_context_biomedical_imaging_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-lineinfo",
         "-Xcompiler", "-fPIC"]
OBJS = ["capi", "dense_attn_fwd", "dense_attn_bwd", "window_attn", "window_attn_small", "patch_embed",
        "patch_embed_mma", "attn_merge", "layer_norm"]
# name -> (source file, extra defines)
VARIANTS = {
    "base": ("dense_attn_bwd", []),
    "red1": ("dense_attn_bwd", ["-DLCBI_BWD_DQ_RED=1"]),
    "red2": ("dense_attn_bwd", ["-DLCBI_BWD_DQ_RED=2"]),
    "red1s6": ("dense_attn_bwd", ["-DLCBI_BWD_DQ_RED=1", "-DLCBI_BWD_QSTAGES=6"]),
    "red2s6": ("dense_attn_bwd", ["-DLCBI_BWD_DQ_RED=2", "-DLCBI_BWD_QSTAGES=6"]),
    "poly4": ("dense_attn_bwd", ["-DLCBI_BWD_POLY_EXP=4"]),
    "poly8": ("dense_attn_bwd", ["-DLCBI_BWD_POLY_EXP=8"]),
    "poly16": ("dense_attn_bwd", ["-DLCBI_BWD_POLY_EXP=16"]),
    "red2s6poly8": ("dense_attn_bwd", ["-DLCBI_BWD_DQ_RED=2", "-DLCBI_BWD_QSTAGES=6", "-DLCBI_BWD_POLY_EXP=8"]),
    "split": ("dense_attn_bwd", ["-DLCBI_BWD_SPLIT_STEPS=1"]),
    "chunked": ("dense_attn_bwd", ["-DLCBI_BWD_CHUNKED=1"]),
    "q32": ("dense_attn_bwd", ["-DLCBI_BWD_Q32=1"]),
    "il2": ("dense_attn_bwd", ["-DLCBI_BWD_INTERLEAVE=2"]),
    "il3": ("dense_attn_bwd", ["-DLCBI_BWD_INTERLEAVE=3"]),
    "il4": ("dense_attn_bwd", ["-DLCBI_BWD_INTERLEAVE=4"]),
}


def lib_path(name):
    return os.path.join(CSRC, "liblcbi_b200.so" if name == "base" else f"liblcbi_b200_{name}.so")


def build(names):
    from long_context_biomedical_imaging_b200.build import build_library
    build_library()
    for name in names:
        if name == "base":
            continue
        src, defs = VARIANTS[name]
        obj = os.path.join(CSRC, f"_var_{name}.o")
        subprocess.run(["nvcc"] + FLAGS + defs + ["-c", os.path.join(CSRC, src + ".cu"), "-o", obj], check=True)
        objs = [obj if o == src else os.path.join(CSRC, o + ".o") for o in OBJS]
        subprocess.run(["nvcc", "-shared", "-o", lib_path(name)] + objs + ["-lcudart"], check=True,
                       stderr=subprocess.DEVNULL)
        os.remove(obj)
        print("built", lib_path(name))


def maxrel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def reference(qkv_q, qkv_k, qkv_v, d_o, scale):
    import torch
    q, k, v = [t.float().detach().requires_grad_(True) for t in (qkv_q, qkv_k, qkv_v)]
    s = torch.einsum("bxhd,byhd->bhxy", q, k) * scale
    o = torch.einsum("bhxy,byhd->bxhd", s.softmax(-1), v)
    return torch.autograd.grad(o, (q, k, v), d_o.float())


def run(name):
    import torch
    from long_context_biomedical_imaging_b200 import _lib
    _lib._lib = None
    _lib.LIB_PATH = lib_path(name)
    from long_context_biomedical_imaging_b200 import ops
    d, scale = 64, 0.125
    worst = 0.0
    for (B, H, Nq, Nk) in [(2, 3, 197, 197), (1, 2, 64, 300), (1, 2, 333, 130), (1, 12, 1728, 1728), (1, 1, 1, 1)]:
        torch.manual_seed(Nq * 7 + Nk)
        q = torch.randn(B, Nq, H, d, device="cuda").to(torch.bfloat16)
        k = torch.randn(B, Nk, H, d, device="cuda").to(torch.bfloat16)
        v = torch.randn(B, Nk, H, d, device="cuda").to(torch.bfloat16)
        o, lse = ops.dense_attn_fwd(q, k, v, scale)
        d_o = torch.randn_like(o)
        g = reference(q, k, v, d_o, scale)
        dq, dk, dv = ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale)
        errs = [maxrel(a, b) for a, b in zip((dq, dk, dv), g)]
        # fp32 accumulate modes (ring): accumulate twice into zeroed buffers -> 2 x gradient
        aq = torch.zeros(B, Nq, H, d, device="cuda")
        ak, av = torch.zeros(B, Nk, H, d, device="cuda"), torch.zeros(B, Nk, H, d, device="cuda")
        for _ in range(2):
            ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=aq, dk=ak, dv=av, accumulate_dkv=True, accumulate_dq=True)
        errs += [maxrel(a / 2, b) for a, b in zip((aq, ak, av), g)]
        bad = any(not (e < 2e-2) for e in errs)
        worst = max(worst, max(errs))
        if bad:
            print(f"  {name}: PARITY FAIL at B={B} H={H} Nq={Nq} Nk={Nk}: {['%.2e' % e for e in errs]}", flush=True)
    B, H, N = 16, 12, 1728
    torch.manual_seed(0)
    qkv = torch.randn(B, N, 3, H, d, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.dense_attn_fwd(q, k, v, scale)
    d_o = torch.randn_like(o)
    dq, dk, dv = (torch.empty_like(q) for _ in range(3))
    fn = lambda: ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=dq, dk=dk, dv=dv)   # noqa: E731
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print(f"{name:14s} worst max-rel {worst:.2e}   bwd group {best:.4f} ms  "
          f"({8 * B * H * N * N * d / best / 1e9:.0f} TF/s algorithmic)", flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "build":
        build(args[1:] or list(VARIANTS))
    else:
        for n in (args or list(VARIANTS)):
            try:
                run(n)
            except Exception as exc:   # a trapped kernel poisons the context: stop here
                print(f"{n}: FAILED {type(exc).__name__}: {exc}", flush=True)
                break
