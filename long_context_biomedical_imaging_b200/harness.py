"""Trainer glue for the reference harness: the three places outside the encoders that have to know
about a B200 and about sequence / window parallelism.

  reference                                          here
  utils/status.py:50-58    support_bfloat16          support_bfloat16: by compute capability, so a B200 gets bf16 autocast
                                                     instead of fp16 + GradScaler (same DISABLE_FLOAT16_INFERENCE switch)
  setup/setup_utils.py:65-82 setup_ddp               make_parallel_layout: after init_process_group, split the world into
                                                     data-parallel replicas x model-parallel (sequence or window) groups
  trainer/trainer_base.py:94-98  DDP(model, ...)     wrap_ddp: DDP over the DATA group only, plus sync_model_parallel_grads
                                                     for the partial parameter gradients of a sequence-parallel ViT

Layout: rank = data_index * model_parallel + model_index, so the ranks of one model-parallel group are neighbours
(same node, NVLink) and the data-parallel all-reduce runs between equal model_index ranks.
"""
from __future__ import annotations

import dataclasses
import os

import torch
import torch.distributed as dist


def support_bfloat16(device=None) -> bool:
    """Reference semantics (utils/status.py:50-58: env switch, then a device test), with the device test by compute
    capability >= 8.0 instead of the name list {A100, H100} - a B200 reports 10.0."""
    if os.environ.get("DISABLE_FLOAT16_INFERENCE", "False") == "True":
        return False
    if not torch.cuda.is_available():
        return False
    dev = torch.device(device) if device is not None else torch.device("cuda")
    if dev.type != "cuda":
        return False
    return torch.cuda.get_device_capability(dev)[0] >= 8


@dataclasses.dataclass
class ParallelLayout:
    world: int
    rank: int
    model_parallel: int           # ranks that share one sample: the sequence_group (ViT) / window_group (Swin)
    data_parallel: int            # replicas that see different samples
    model_index: int
    data_index: int
    model_group: object           # ProcessGroup or None when model_parallel == 1
    data_group: object            # ProcessGroup or None when data_parallel == 1


def make_parallel_layout(model_parallel: int = 1) -> ParallelLayout:
    """Call after dist.init_process_group (reference setup_ddp). Every rank must call it with the same argument: all
    groups are created on all ranks, as torch.distributed requires."""
    if not (dist.is_available() and dist.is_initialized()):
        return ParallelLayout(1, 0, 1, 1, 0, 0, None, None)
    world, rank = dist.get_world_size(), dist.get_rank()
    if model_parallel < 1 or world % model_parallel != 0:
        raise ValueError(f"world size {world} is not a multiple of model_parallel {model_parallel}")
    data_parallel = world // model_parallel
    model_group = data_group = None
    for d in range(data_parallel):
        ranks = list(range(d * model_parallel, (d + 1) * model_parallel))
        grp = dist.new_group(ranks) if model_parallel > 1 else None
        if rank in ranks:
            model_group = grp
    for m in range(model_parallel):
        ranks = list(range(m, world, model_parallel))
        grp = dist.new_group(ranks) if data_parallel > 1 else None
        if rank in ranks:
            data_group = grp
    return ParallelLayout(world, rank, model_parallel, data_parallel, rank % model_parallel, rank // model_parallel,
                          model_group, data_group)


def apply_layout(encoder, layout: ParallelLayout):
    """Point an encoder of this package at its model-parallel group (set_sequence_group for ViT_with_alt_ops,
    set_window_group for SwinTransformer_with_alt_ops)."""
    if layout.model_parallel == 1:
        return encoder
    if hasattr(encoder, "set_sequence_group"):
        encoder.set_sequence_group(layout.model_group)
    elif hasattr(encoder, "set_window_group"):
        encoder.set_window_group(layout.model_group)
    else:
        raise TypeError(f"{type(encoder).__name__} has neither set_sequence_group nor set_window_group")
    return encoder


def wrap_ddp(model, layout: ParallelLayout, device_ids=None):
    """The reference's `DDP(model, device_ids=[rank], find_unused_parameters=False)` over the data-parallel group only:
    the ranks of a model-parallel group hold the SAME sample, their gradients are not independent replicas."""
    if layout.data_parallel == 1:
        return model
    from torch.nn.parallel import DistributedDataParallel as DDP

    return DDP(model, device_ids=device_ids, process_group=layout.data_group, find_unused_parameters=False)


def sync_model_parallel_grads(model, layout: ParallelLayout, sequence_parallel: bool = True):
    """After backward. A sequence-parallel ViT computes every parameter gradient from the LOCAL token rows, so the
    ranks of a sequence group hold partial sums: add them up (one flat all-reduce). A window-parallel Swin needs
    nothing here - its activations are replicated and window_parallel.py already reduces the two attention-internal
    parameter gradients."""
    if layout.model_parallel == 1 or not sequence_parallel:
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=layout.model_group)
    offset = 0
    for g in grads:
        g.copy_(flat[offset:offset + g.numel()].view_as(g))
        offset += g.numel()
