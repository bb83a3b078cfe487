"""Where the window-sharded Swin block spends its time (cfg4 stage 1), per rank, with torch.profiler.
torchrun --nproc-per-node 2 tools/prof_window_sharded.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from long_context_biomedical_imaging_b200 import ops, window_parallel  # noqa: E402

rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(rank)
dist.init_process_group("nccl")
dev = torch.device("cuda", rank)
B, grid, C, H, window = 1, (64, 64, 64), 48, 3, (7, 7, 7)
shift = (3, 3, 3)
torch.manual_seed(0)
qkv = torch.randn(B, *grid, 3 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
bias = torch.randn(3 * C, device=dev).requires_grad_(True)
table = (torch.randn(13 ** 3, H, device=dev) * 0.5).requires_grad_(True)
d_out = torch.randn(B, *grid, C, device=dev).to(torch.bfloat16)


def step():
    out = window_parallel.window_attention_sharded(qkv, bias, table, grid, window, shift, H)
    out.backward(d_out)
    qkv.grad = bias.grad = table.grad = None


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world {dist.get_world_size()}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per step")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
dist.destroy_process_group()
