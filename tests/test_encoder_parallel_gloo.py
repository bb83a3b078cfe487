"""CPU, world_size 2 over gloo: the sequence-parallel ViT encoder (`sequence_group`) and the window-parallel Swin
encoder (`window_group`) against the same encoders run un-sharded in one process. The product operators have no CPU
path, so the test substitutes the oracle's math for them (the host-side sharding, the ring driver, the all-gather of
owned window rows and the gradient plumbing are what is being exercised)."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _install_cpu_ops():
    """Oracle-backed stand-ins for the CUDA operators (test infrastructure only)."""
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    for path in (here, os.path.dirname(here)):
        if path not in sys.path:
            sys.path.insert(0, path)
    import test_ring_gloo as trg
    from long_context_biomedical_imaging_b200 import ops, ring
    from oracle import attention_oracle as ao

    ops._require_cuda = lambda *t: next((x.device for x in t if x is not None), None)
    ops.layer_norm = lambda x, w, b, eps=1e-5, out_dtype=None: F.layer_norm(x, (x.shape[-1],), w, b, eps)

    def add_layer_norm(x, delta, w, b, eps=1e-5, out_dtype=None):
        xs = x + delta
        return xs, F.layer_norm(xs, (xs.shape[-1],), w, b, eps)

    ops.add_layer_norm = add_layer_norm
    ops.linear = F.linear

    def patch_embed(img, weight, bias, pos, grid, out_dtype=torch.float32):
        if pos is not None:
            return ao.patch_embed_vit(img, weight, bias, pos)
        return ao.patch_embed_swin(img, weight, bias).movedim(1, -1).reshape(img.shape[0], -1, weight.shape[0])

    ops.patch_embed = patch_embed

    def dense_attention_qkv(qkv, num_heads, scale=None):
        q, k, v = ao.split_qkv_vit(qkv, num_heads)
        B, N, C3 = qkv.shape
        o = ao.dense_attention(q, k, v, scale if scale is not None else (C3 // 3 // num_heads) ** -0.5)
        return o.permute(0, 2, 1, 3).reshape(B, N, C3 // 3)

    ops.dense_attention_qkv = dense_attention_qkv
    ops.window_attention = ao.window_attention_core
    ring._cuda_fwd_state = trg._fwd_state
    ring._cuda_bwd = trg._bwd


def _vit_cfg(group):
    return types.SimpleNamespace(ViT=types.SimpleNamespace(size="custom", hidden_size=32, mlp_dim=64, num_layers=2,
                                                           num_heads=2, patch_size=[1, 2, 2], use_hyena=False,
                                                           use_mamba=False, sequence_group=group),
                                 time=1, height=16, width=12, task_type="seg")


def _swin_cfg(group):
    return types.SimpleNamespace(Swin=types.SimpleNamespace(size="custom", embed_dim=8, depths=[2, 2, 2, 2],
                                                            num_heads=[1, 2, 2, 2], patch_size=[1, 2, 2],
                                                            window_size=[1, 4, 4], use_hyena=False, use_mamba=False,
                                                            window_group=group),
                                 time=1, height=40, width=36, task_type="seg")


def _worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_cpu_ops()
        from long_context_biomedical_imaging_b200.backbone_swin import custom_Swin
        from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT
        from oracle import attention_oracle as ao

        errs = {}
        # ---------------- ViT: sequence-parallel vs one process
        torch.manual_seed(0)
        x = torch.randn(2, 1, 1, 16, 12)
        sharded, _ = custom_ViT(_vit_cfg("world"), 1)
        plain, _ = custom_ViT(_vit_cfg(None), 1)
        ao.fill_parameters_(sharded, 3)
        plain.load_state_dict(sharded.state_dict())
        outs_s, outs_p = sharded(x), plain(x)
        assert len(outs_s) == len(outs_p) == 4
        w = [torch.linspace(-1, 1, o.numel()).reshape(o.shape) for o in outs_p[1:]]
        sum((o * wi).sum() for o, wi in zip(outs_s[1:], w)).backward()
        sum((o * wi).sum() for o, wi in zip(outs_p[1:], w)).backward()
        errs["vit_out"] = max(float((a - b).abs().max()) for a, b in zip(outs_s[1:], outs_p[1:]))
        g_err = 0.0
        for (name, ps), pp in zip(sharded.named_parameters(), plain.parameters()):
            g = ps.grad.clone() if ps.grad is not None else torch.zeros_like(ps)
            dist.all_reduce(g)          # parameter gradients are partial sums over the local tokens
            g_err = max(g_err, float((g - pp.grad).abs().max() / pp.grad.abs().max().clamp_min(1e-6)))
        errs["vit_grad"] = g_err
        # ---------------- Swin: window-parallel vs one process
        torch.manual_seed(1)
        x = torch.randn(1, 1, 1, 40, 36)
        sharded, _ = custom_Swin(_swin_cfg("world"), 1)
        plain, _ = custom_Swin(_swin_cfg(None), 1)
        ao.fill_parameters_(sharded, 5)
        plain.load_state_dict(sharded.state_dict())
        outs_s, outs_p = sharded(x), plain(x)
        w = [torch.linspace(-1, 1, o.numel()).reshape(o.shape) for o in outs_p[1:]]
        sum((o * wi).sum() for o, wi in zip(outs_s[1:], w)).backward()
        sum((o * wi).sum() for o, wi in zip(outs_p[1:], w)).backward()
        errs["swin_out"] = max(float((a - b).abs().max()) for a, b in zip(outs_s[1:], outs_p[1:]))
        errs["swin_grad"] = max(float((ps.grad - pp.grad).abs().max() / pp.grad.abs().max().clamp_min(1e-6))
                                for ps, pp in zip(sharded.parameters(), plain.parameters()))
        queue.put((rank, errs))
    finally:
        dist.destroy_process_group()


def test_sequence_parallel_vit_and_window_parallel_swin_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    for _ in range(600):                          # up to 5 minutes, but fail at once when a worker dies
        try:
            results.append(queue.get(timeout=0.5))
        except Exception:  # noqa: BLE001 - queue.Empty
            if any(p.exitcode not in (None, 0) for p in procs):
                break
        if len(results) == world:
            break
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(results) == world
    for rank, errs in results:
        assert all(e < 1e-4 for e in errs.values()), (rank, errs)
