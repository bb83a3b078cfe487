"""The drop-in encoders are CUDA-graph capturable: every C-ABI entry point only enqueues on the caller's stream, allocates
nothing itself (buffers come from PyTorch's graph-aware caching allocator) and never synchronises, so a whole
forward + backward step can be captured once and replayed. That is how the launch-bound configurations (the late
Swin stages: a handful of windows per kernel) get rid of the host-side launch path.

A replayed step must give the hidden states and parameter gradients of the eager step on the same inputs (bf16 path:
the bias-table and dQ gradients are summed with fp32 atomics, so not bit-identical; 1e-2 max-rel)."""
import types

import pytest
import torch

from conftest import max_rel
from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu


def _build(kind):
    if kind == "vit":
        from long_context_biomedical_imaging_b200.backbone_vit import custom_ViT

        cfg = types.SimpleNamespace(ViT=types.SimpleNamespace(size="custom", hidden_size=128, mlp_dim=256, num_layers=2,
                                                              num_heads=2, patch_size=[1, 8, 8], use_hyena=False,
                                                              use_mamba=False), time=1, height=32, width=48, task_type="seg")
        model, _ = custom_ViT(cfg, 1)
        shape = (2, 1, 1, 32, 48)
    else:
        from long_context_biomedical_imaging_b200.backbone_swin import custom_Swin

        cfg = types.SimpleNamespace(Swin=types.SimpleNamespace(size="custom", embed_dim=16, depths=[2, 2, 2, 2],
                                                               num_heads=[1, 2, 4, 8], patch_size=[1, 2, 2],
                                                               window_size=[1, 4, 4], use_hyena=False, use_mamba=False),
                                    time=1, height=48, width=40, task_type="seg")
        model, _ = custom_Swin(cfg, 1)
        shape = (2, 1, 1, 48, 40)
    ao.fill_parameters_(model, 11)
    return model.cuda(), shape


def _step(model, x):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    loss = sum(o.float().square().mean() for o in outs[1:])
    loss.backward()
    return outs, loss


@pytest.mark.parametrize("kind", ["vit", "swin"])
def test_encoder_step_replays_from_a_cuda_graph(kind):
    model, shape = _build(kind)
    torch.manual_seed(3)
    x = torch.randn(*shape, device="cuda")

    outs, loss = _step(model, x)
    eager_out = [o.detach().float().clone() for o in outs]
    eager_loss = float(loss.detach())
    eager_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    assert all(g is not None for g in eager_grads.values())
    del outs, loss      # nothing of the eager autograd graph (its AccumulateGrad nodes live on the default stream) may survive

    # warm-up on a side stream, static gradient buffers, then capture
    model.zero_grad(set_to_none=False)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            _step(model, x)
    torch.cuda.current_stream().wait_stream(side)
    static_x = x.clone()
    for p in model.parameters():
        p.grad.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_outs, g_loss = _step(model, static_x)

    for trial in range(2):      # second replay with a different input written into the static buffer and back
        for p in model.parameters():
            p.grad.zero_()
        static_x.copy_(x if trial == 0 else x * 0.5)
        graph.replay()
        torch.cuda.synchronize()
        if trial == 0:
            assert abs(float(g_loss.detach()) - eager_loss) <= 1e-2 * abs(eager_loss)
            for got, want in zip(g_outs, eager_out):
                assert max_rel(got.detach().float().cpu(), want.cpu()) < 1e-2
            for n, p in model.named_parameters():
                want = eager_grads[n]
                scale = float(want.abs().max())
                if scale == 0.0:
                    assert float(p.grad.abs().max()) == 0.0, n
                else:
                    assert max_rel(p.grad.float().cpu(), want.float().cpu()) < 1e-2, n
        else:
            # the new input was really used (the loss itself is nearly input-independent for layer-normed outputs)
            assert max_rel(g_outs[1].detach().float().cpu(), eager_out[1].cpu()) > 1e-3
