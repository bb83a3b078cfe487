"""LayerNorm kernels at encoder sizes: CUDA-event timings, algorithmic HBM bytes / time against the measured copy
bandwidth (MEASURED_PEAKS.json, else the profiling guide's fallback), next to torch.nn.functional.layer_norm under the
same dtypes (fp32 rows in, bf16 out as under bf16 autocast, i.e. layer_norm + the cast the next Linear applies).

    python tools/bench_layer_norm.py
"""
import json
import os
import re
import sys
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from long_context_biomedical_imaging_b200 import ops  # noqa: E402


def hbm_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        for key in ("hbm_gbs", "hbm_GBps", "hbm_copy_gbs", "hbm_bw_gbs"):
            if key in peaks:
                return float(peaks[key]), "measured"
        for k, v in peaks.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), "measured"
    except (OSError, ValueError):
        pass
    return 6539.9, "SURVEY 8d figure"


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best * 1e-3


def main():
    warnings.filterwarnings("ignore")
    peak, src = hbm_peak_gbs()
    print(f"HBM roofline {peak:.0f} GB/s ({src}); inputs of every case exceed the 126 MB L2 or rotate over 4 copies")
    cases = [("cfg3 ViT-B block norm, 16 x 1728 tokens", 16 * 1728, 768, torch.float32),
             ("cfg5 ViT-B block norm, 262144 tokens", 262144, 768, torch.float32),
             ("cfg2 Swin-T stage 1, 16 x 128 x 128 tokens", 16 * 128 * 128, 96, torch.bfloat16),
             ("cfg4 SwinUNETR stage 1, 64^3 tokens", 64 ** 3, 48, torch.bfloat16)]
    for name, rows, C, x_dtype in cases:
        n_copies = 4
        xs = [(torch.randn(rows, C, device="cuda") * 2 + 0.5).to(x_dtype).requires_grad_(True) for _ in range(n_copies)]
        w = torch.ones(C, device="cuda", requires_grad=True)
        b = torch.zeros(C, device="cuda", requires_grad=True)
        gs = [torch.randn(rows, C, device="cuda").to(torch.bfloat16) for _ in range(n_copies)]
        xb = 4 if x_dtype == torch.float32 else 2
        state = {"i": 0}

        def ours_fwd():
            state["i"] = (state["i"] + 1) % n_copies
            return ops.layer_norm(xs[state["i"]], w, b, 1e-5, out_dtype=torch.bfloat16)

        def torch_fwd():
            state["i"] = (state["i"] + 1) % n_copies
            return F.layer_norm(xs[state["i"]].float(), (C,), w, b, 1e-5).to(torch.bfloat16)

        def fwd_bwd(f, params=True):
            def run():
                y = f()
                torch.autograd.grad(y, (xs[state["i"]], w, b) if params else (xs[state["i"]],), gs[state["i"]])
            return run

        t_f, t_tf = timeit(ours_fwd), timeit(torch_fwd)
        t_fb, t_tfb = timeit(fwd_bwd(ours_fwd)), timeit(fwd_bwd(torch_fwd))
        fwd_bytes = rows * C * (xb + 2)
        bwd_bytes = rows * C * (2 + xb + xb) + rows * C * (2 + xb)   # dx pass + the d(gamma)/d(beta) pass
        print(f"{name}: rows {rows} x C {C}, x {str(x_dtype).split('.')[-1]} -> bf16")
        print(f"   fwd      {t_f * 1e6:8.1f} us  {fwd_bytes / t_f / 1e9:7.0f} GB/s = {fwd_bytes / t_f / 1e9 / peak:5.1%} of roofline"
              f"   | torch layer_norm + cast {t_tf * 1e6:8.1f} us")
        t_b = t_fb - t_f
        # the op-level timings above include the Python / autograd launch path (tens of microseconds per call, more
        # than the narrow-row kernels themselves take): per-kernel device times from the profiler beside them
        from torch.profiler import ProfilerActivity, profile
        run = fwd_bwd(ours_fwd)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(8):
                run()
            torch.cuda.synchronize()
        kern = {re.search(r"(ln_\w+)", e.key).group(1): e.device_time_total / e.count
                for e in prof.key_averages() if re.search(r"(ln_\w+)", e.key)}
        print(f"   fwd+bwd  {t_fb * 1e6:8.1f} us  (bwd {t_b * 1e6:.1f} us, {bwd_bytes / t_b / 1e9:.0f} GB/s = "
              f"{bwd_bytes / t_b / 1e9 / peak:5.1%})   | torch {t_tfb * 1e6:8.1f} us")
        dev = {k: v for k, v in kern.items()}
        dx_b, par_b = rows * C * (2 + xb + xb), rows * C * (2 + xb)
        line = "   device time per kernel: " + ", ".join(f"{k} {v:.1f} us" for k, v in sorted(dev.items()))
        print(line)
        for k, nbytes in (("ln_fwd", fwd_bytes), ("ln_bwd_dx", dx_b), ("ln_bwd_params", par_b)):
            hit = [v for kk, v in dev.items() if kk.startswith(k) and "finish" not in kk]
            if hit:
                print(f"      {k:14s} {nbytes / hit[0] / 1e3:7.0f} GB/s = {nbytes / hit[0] / 1e3 / peak:5.1%} of roofline")


if __name__ == "__main__":
    main()
