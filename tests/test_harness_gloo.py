"""CPU, world_size 4 over gloo: the trainer glue of harness.py - 2 data-parallel replicas x 2
sequence-parallel ranks of the ViT encoder, DDP over the data group, partial parameter gradients summed over the
sequence group - against one process that sees both samples. Oracle-backed operator stand-ins as in
test_encoder_parallel_gloo.py (the product operators have no CPU path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_encoder_parallel_gloo import _free_port, _install_cpu_ops, _vit_cfg  # noqa: E402


def _worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_cpu_ops()
        from long_context_biomedical_imaging_b200 import harness
        from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT
        from oracle import attention_oracle as ao

        layout = harness.make_parallel_layout(model_parallel=2)
        assert (layout.data_parallel, layout.model_index, layout.data_index) == (2, rank % 2, rank // 2)
        assert dist.get_world_size(layout.model_group) == 2 and dist.get_world_size(layout.data_group) == 2

        torch.manual_seed(0)
        x_all = torch.randn(2, 1, 1, 1, 16, 12)              # one sample per data-parallel replica
        model, _ = custom_ViT(_vit_cfg(None), 1)
        plain, _ = custom_ViT(_vit_cfg(None), 1)
        ao.fill_parameters_(model, 3)
        plain.load_state_dict(model.state_dict())
        harness.apply_layout(model, layout)
        ddp = harness.wrap_ddp(model, layout)

        outs = ddp(x_all[layout.data_index])
        w = [torch.linspace(-1, 1, o.numel()).reshape(o.shape) for o in outs[1:]]
        sum((o * wi).sum() for o, wi in zip(outs[1:], w)).backward()
        harness.sync_model_parallel_grads(model, layout)

        want = [torch.zeros_like(p) for p in plain.parameters()]
        for d in range(2):                                   # DDP averages over the replicas
            plain.zero_grad()
            outs_p = plain(x_all[d])
            sum((o * wi).sum() for o, wi in zip(outs_p[1:], w)).backward()
            for acc, p in zip(want, plain.parameters()):
                acc += p.grad / 2
            if d == layout.data_index:
                out_err = max(float((a - b).abs().max()) for a, b in zip(outs[1:], outs_p[1:]))
        g_err = max(float((p.grad - g).abs().max() / g.abs().max().clamp_min(1e-6)) for p, g in zip(model.parameters(), want))
        queue.put((rank, out_err, g_err))
    finally:
        dist.destroy_process_group()


def test_support_bfloat16_follows_the_reference_switch(monkeypatch):
    from long_context_biomedical_imaging_b200 import harness

    monkeypatch.setenv("DISABLE_FLOAT16_INFERENCE", "True")
    assert harness.support_bfloat16("cuda:0") is False
    monkeypatch.delenv("DISABLE_FLOAT16_INFERENCE")
    if not torch.cuda.is_available():
        assert harness.support_bfloat16("cuda:0") is False
    assert harness.support_bfloat16("cpu") is False
    single = harness.make_parallel_layout(1)                 # no process group: a one-rank layout
    assert (single.world, single.model_group, single.data_group) == (1, None, None)


def test_data_parallel_x_sequence_parallel_layout_matches_single_process():
    world = 4
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    for _ in range(600):
        try:
            results.append(queue.get(timeout=0.5))
        except Exception:  # noqa: BLE001 - queue.Empty
            if any(p.exitcode not in (None, 0) for p in procs):
                break
        if len(results) == world:
            break
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    assert len(results) == world, [p.exitcode for p in procs]
    for rank, out_err, g_err in results:
        assert out_err < 1e-4, (rank, out_err)
        assert g_err < 1e-4, (rank, g_err)
