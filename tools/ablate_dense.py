"""Timing-only ablations of the dense kernels: where does the step time go?

    python tools/ablate_dense.py build            # here (no GPU): one .so per ablation mask
    python tools/ablate_dense.py                  # on the GPU box: time every variant at cfg3 (B=16)

Each variant recompiles one kernel with -DLCBI_BWD_ABLATE=<mask> / -DLCBI_FWD_ABLATE=<mask> (see the kernels for the bit
meanings). Results of ablated variants are numerically wrong by design; only the timings mean anything.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "long_context_biomedical_imaging_b200", "csrc")
BWD_MASKS = [int(x) for x in os.environ.get("ABL_BWD", "").split(",") if x != ""]   # backward switches existed up to commit a07535e
FWD_MASKS = [int(x) for x in os.environ.get("ABL_FWD", "0,1,2,3,4").split(",") if x != ""]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-lineinfo", "-Xcompiler", "-fPIC"]
OTHERS = ["capi", "dense_attn_fwd", "dense_attn_bwd", "window_attn", "window_attn_small", "patch_embed", "patch_embed_mma", "attn_merge", "layer_norm"]


def lib_path(kind, mask):
    return os.path.join(CSRC, f"liblcbi_b200_abl_{kind}{mask}.so")


def build():
    from long_context_biomedical_imaging_b200.build import build_library
    build_library()
    for kind, masks, src, macro in (("bwd", BWD_MASKS, "dense_attn_bwd", "LCBI_BWD_ABLATE"), ("fwd", FWD_MASKS, "dense_attn_fwd", "LCBI_FWD_ABLATE")):
        for m in masks:
            obj = os.path.join(CSRC, f"_abl_{kind}{m}.o")
            extra = os.environ.get("ABL_DEFS", "").split()
            subprocess.run(["nvcc"] + FLAGS + extra + [f"-D{macro}={m}", "-c", os.path.join(CSRC, src + ".cu"), "-o", obj], check=True)
            objs = [obj if o == src else os.path.join(CSRC, o + ".o") for o in OTHERS]
            subprocess.run(["nvcc", "-shared", "-o", lib_path(kind, m)] + objs + ["-lcudart"], check=True)
            os.remove(obj)
            print("built", lib_path(kind, m))


def run_one(path, kind):
    import torch
    from long_context_biomedical_imaging_b200 import _lib
    _lib.LIB_PATH = path
    from long_context_biomedical_imaging_b200 import ops
    B, H, N, d = int(os.environ.get("ABL_B", 16)), 12, int(os.environ.get("ABL_N", 1728)), 64
    torch.manual_seed(0)
    qkv = torch.randn(B, N, 3, H, d, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
    d_o = torch.randn_like(o)
    fn = (lambda: ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)) if kind == "bwd" else (lambda: ops.dense_attn_fwd(q, k, v, 0.125, out=o))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    msg = ""
    if path.endswith(("fwd0.so", "bwd0.so")):      # un-ablated variant: also check the numbers
        import torch.nn.functional as F
        qf, kf, vf = [t[:1].float().permute(0, 2, 1, 3).requires_grad_(True) for t in (q, k, v)]
        ref = F.scaled_dot_product_attention(qf, kf, vf, scale=0.125)
        o2, lse2 = ops.dense_attn_fwd(q, k, v, 0.125)
        rel = lambda a, b: ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
        msg = f" | o max-rel {rel(o2[:1].permute(0, 2, 1, 3), ref):.2e}"
        if kind == "bwd":
            gq, gk, gv = torch.autograd.grad(ref, (qf, kf, vf), d_o[:1].float().permute(0, 2, 1, 3))
            dq, dk, dv = ops.dense_attn_bwd(q, k, v, o2, d_o, lse2, 0.125)
            msg += f" dq {rel(dq[:1].permute(0, 2, 1, 3), gq):.2e} dk {rel(dk[:1].permute(0, 2, 1, 3), gk):.2e} dv {rel(dv[:1].permute(0, 2, 1, 3), gv):.2e}"
    print(f"{os.path.basename(path):40s} {kind} {e0.elapsed_time(e1) / iters:.4f} ms{msg}", flush=True)


if __name__ == "__main__":
    if sys.argv[1:2] == ["build"]:
        build()
    elif sys.argv[1:2] == ["one"]:
        run_one(sys.argv[2], sys.argv[3])
    else:
        for kind, masks in (("bwd", BWD_MASKS), ("fwd", FWD_MASKS)):
            for m in masks:
                p = lib_path(kind, m)
                if os.path.exists(p):
                    subprocess.run(["timeout", "120", sys.executable, os.path.abspath(__file__), "one", p, kind])
