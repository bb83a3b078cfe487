"""Debug: per-role clock64 timeline of CTA (0,0,0) of the dense forward kernel (needs the -DLCBI_TRACE build)."""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "long_context_biomedical_imaging_b200", "csrc")
TRACE_LIB = os.path.join(CSRC, "liblcbi_b200_trace.so")


def build():
    srcs = [os.path.join(CSRC, f) for f in ("capi.cu", "dense_attn_fwd.cu", "dense_attn_bwd.cu", "window_attn.cu", "window_attn_small.cu",
                                            "patch_embed.cu", "patch_embed_mma.cu", "attn_merge.cu")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--use_fast_math", "-lineinfo",
           "-DLCBI_TRACE", "-Xcompiler", "-fPIC", "-shared", "-o", TRACE_LIB] + srcs + ["-lcudart"]
    subprocess.run(cmd, check=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["build"]:
        build()
        sys.exit(0)
    from long_context_biomedical_imaging_b200 import _lib
    _lib.LIB_PATH = TRACE_LIB
    from long_context_biomedical_imaging_b200 import ops
    lib = _lib.load()
    B, H, N, d = int(os.environ.get("TR_B", 16)), 12, 1728, 64
    qkv = torch.randn(B, N, 3, H, d, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    for _ in range(2):
        ops.dense_attn_fwd(q, k, v, 0.125)
    trace = torch.zeros(3 * 32 * 8, dtype=torch.int64, device="cuda")
    lib.lcbi_debug_set_fwd_trace.argtypes = [ctypes.c_void_p]
    assert lib.lcbi_debug_set_fwd_trace(ctypes.c_void_p(trace.data_ptr())) == 0
    ops.dense_attn_fwd(q, k, v, 0.125)
    torch.cuda.synchronize()
    if os.environ.get("TR_ITEMS"):
        # wall-clock phases (globaltimer, compute thread 0) of every item of every persistent backward CTA
        o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
        d_o = torch.randn_like(o)
        for _ in range(2):
            ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
        n_cta = 148
        ct = torch.zeros(n_cta * 32 * 8, dtype=torch.int64, device="cuda")
        lib.lcbi_debug_set_bwd_item_times.argtypes = [ctypes.c_void_p]
        assert lib.lcbi_debug_set_bwd_item_times(ctypes.c_void_p(ct.data_ptr())) == 0
        ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
        torch.cuda.synchronize()
        t = ct.cpu().view(n_cta, 32, 8).double()
        valid = t[:, :, 0] > 0
        names = ["item start -> step 1", "step 1 -> step 2", "step 2 -> last step done", "copy next K/V to TMEM", "wait all GEMMs retired",
                 "read dV/dK + stage + store"]
        for i, nm in enumerate(names):
            d = (t[:, :, i + 1] - t[:, :, i])[valid]
            print(f"{nm:28s} mean {d.mean():8.0f} ns  median {d.median():8.0f}  p90 {d.quantile(0.9):8.0f}")
        tot = (t[:, :, 6] - t[:, :, 0])[valid]
        print(f"item total                   mean {tot.mean():8.0f} ns  median {tot.median():8.0f}  items {int(valid.sum())}")
        nxt = (t[:, 1:, 0] - t[:, :-1, 6])[valid[:, 1:]]
        print(f"store issue -> next start    mean {nxt.mean():8.0f} ns")
        first = t[:, 0, 0][valid[:, 0]]
        last = t[:, :, 6].max(dim=1).values
        print(f"first item starts spread {first.max() - first.min():.0f} ns; CTA busy span mean {(last - t[:, 0, 0]).mean():.0f} max {(last - t[:, 0, 0]).max():.0f} ns")
        print("CTA 0 items:", [int(x) for x in (t[0, :, 6] - t[0, :, 0])[valid[0]]])
        sys.exit(0)
    if os.environ.get("TR_BWD"):
        o, lse = ops.dense_attn_fwd(q, k, v, 0.125)
        d_o = torch.randn_like(o)
        for _ in range(2):
            ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
        btrace = torch.zeros(4 * 16 * 8, dtype=torch.int64, device="cuda")
        lib.lcbi_debug_set_bwd_trace.argtypes = [ctypes.c_void_p]
        assert lib.lcbi_debug_set_bwd_trace(ctypes.c_void_p(btrace.data_ptr())) == 0
        ops.dense_attn_bwd(q, k, v, o, d_o, lse, 0.125)
        torch.cuda.synchronize()
        bt = btrace.cpu().view(4, 16, 8)
        t0 = int(bt[bt > 0].min())
        for role, name in enumerate(["CMP half0", "CMP half1", "MMA", "DRAIN"]):
            print(name)
            for step in range(16):
                row = [int(x) - t0 if x > 0 else -1 for x in bt[role, step]]
                print(f"  tile {step:2d}: " + " ".join(f"{x:7d}" for x in row))
        sys.exit(0)
    t = trace.cpu().view(3, 32, 8)
    t0 = int(t[t > 0].min())
    names = {0: "SOFTMAX", 1: "MMA", 2: "-"}
    for role in range(2):
        print(names[role])
        for step in range(int(os.environ.get('TR_STEPS', 27))):
            row = [int(x) - t0 if x > 0 else -1 for x in t[role, step]]
            print(f"  step {step:2d}: " + " ".join(f"{x:7d}" for x in row))
