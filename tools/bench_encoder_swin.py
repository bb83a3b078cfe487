"""Context only (not a contract number): whole Swin encoder fwd+bwd through the drop-in modules under bf16 autocast, with
the top CUDA kernels of one step.

    python tools/bench_encoder_swin.py           # cfg2: Swin-T 2-D, 512 x 512, patch 4, window 7, batch 16
    SWIN_CFG=cfg4 python tools/bench_encoder_swin.py   # cfg4: 3-D 'unetr' width, 128^3 volume, patch 2, window 7, batch 1
"""
import os
import sys
import types
import warnings

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200.backbone_swin import custom_Swin  # noqa: E402

warnings.filterwarnings("ignore")
which = os.environ.get("SWIN_CFG", "cfg2")
if which == "cfg2":
    cfg = types.SimpleNamespace(Swin=types.SimpleNamespace(size="tiny", patch_size=[1, 4, 4], window_size=[1, 7, 7],
                                                           use_hyena=False, use_mamba=False),
                                time=1, height=512, width=512, task_type="seg")
    x = torch.randn(16, 1, 1, 512, 512, device="cuda")
    label = "Swin-T 2-D encoder, 512 x 512, patch 4, window 7, batch 16"
else:
    cfg = types.SimpleNamespace(Swin=types.SimpleNamespace(size="unetr", patch_size=[2, 2, 2], window_size=[7, 7, 7],
                                                           use_hyena=False, use_mamba=False),
                                time=128, height=128, width=128, task_type="seg")
    x = torch.randn(1, 1, 128, 128, 128, device="cuda")
    label = "Swin 3-D encoder ('unetr' width), 128^3 volume, patch 2, window 7, batch 1"
torch.manual_seed(0)
model, _ = custom_Swin(cfg, 1)
model = model.cuda()


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    sum(o.float().square().mean() for o in outs[1:]).backward()
    model.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print(f"{label}, bf16 autocast, fwd+bwd: {e0.elapsed_time(e1) / 5:.2f} ms per step")

if os.environ.get("SWIN_GRAPH", "1") != "0":
    # The same step captured once into a CUDA graph and replayed: every library entry point only enqueues on the current
    # stream and allocates through PyTorch's (graph-aware) caching allocator, so the whole fwd+bwd is capturable; this
    # removes the host-side launch path, which is what bounds the small late stages.
    model.zero_grad(set_to_none=False)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs = model(x)
            sum(o.float().square().mean() for o in outs[1:]).backward()
    torch.cuda.current_stream().wait_stream(side)
    for prm in model.parameters():
        if prm.grad is not None:
            prm.grad.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x)
        loss = sum(o.float().square().mean() for o in outs[1:])
        loss.backward()
    eager_grads = None
    graph.replay()
    torch.cuda.synchronize()
    g_first = [prm.grad.clone() for prm in model.parameters() if prm.grad is not None]
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"   the same step replayed from a CUDA graph: {e0.elapsed_time(e1) / 10:.2f} ms per step "
          f"(loss {float(loss):.6f}, {len(g_first)} parameter gradients accumulated in place)")
    model.zero_grad(set_to_none=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=64))
