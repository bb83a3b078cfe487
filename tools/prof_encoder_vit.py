"""Top CUDA kernels of one drop-in ViT-B encoder fwd+bwd step (cfg3, bf16 autocast) via torch.profiler."""
import os
import sys
import types

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT  # noqa: E402

B = int(os.environ.get("ENC_B", 16))
cfg = types.SimpleNamespace(ViT=types.SimpleNamespace(size="base", patch_size=[8, 8, 8], use_hyena=False, use_mamba=False),
                            time=96, height=96, width=96, task_type="seg")
model, _ = custom_ViT(cfg, 1)
model = model.cuda()
x = torch.randn(B, 1, 96, 96, 96, device="cuda")


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x)[-1]
    out.float().square().mean().backward()
    model.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
