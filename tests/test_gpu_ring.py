"""GPU: the ring-attention building blocks on ONE device — the P ranks are emulated as a loop over shards (a ring step
never waits on another rank's kernel, so no concurrency is needed): fused attention per visiting K/V shard, the CUDA
log-sum-exp merge kernel, fp32 dq / travelling dk,dv accumulation modes of the backward. Must equal global attention.
With >= 2 GPUs the real NCCL ring is also run."""
import os
import socket

import pytest
import torch

from conftest import max_rel
from oracle import attention_oracle as ao

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,N", [(2, 512), (4, 1024), (8, 8 * 200)])
def test_emulated_ring_equals_global_attention(P, N):
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(P)
    B, H, d = 1, 3, 64
    q, k, v, d_o = [torch.randn(B, N, H, d, device="cuda").to(torch.bfloat16) for _ in range(4)]
    n = N // P
    shard = lambda t, r: t[:, r * n:(r + 1) * n].contiguous()
    outs, lses = [], []
    for r in range(P):
        acc = torch.empty(B, n, H, d, device="cuda")
        lse = torch.empty(B, H, n, device="cuda")
        for s in range(P):
            src = (r - s) % P
            o_s, lse_s = ops.dense_attn_fwd(shard(q, r), shard(k, src), shard(v, src), 0.125)
            ops.attn_merge(acc, lse, o_s, lse_s, s == 0)
        outs.append(acc)
        lses.append(lse)
    o_full = torch.cat(outs, 1)
    qf, kf, vf = [t.float().requires_grad_(True) for t in (q, k, v)]
    ref = ao.dense_attention(qf.permute(0, 2, 1, 3), kf.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), 0.125).permute(0, 2, 1, 3)
    gq, gk, gv = torch.autograd.grad(ref, [qf, kf, vf], d_o.float())
    assert max_rel(o_full.cpu(), ref.detach().cpu()) < 2e-2

    dq = [torch.zeros(B, n, H, d, device="cuda") for _ in range(P)]
    dk = [torch.zeros(B, n, H, d, device="cuda") for _ in range(P)]   # accumulator of K/V shard `src`, wherever it is
    dv = [torch.zeros(B, n, H, d, device="cuda") for _ in range(P)]
    for r in range(P):
        o_bf16 = outs[r].to(torch.bfloat16)
        for s in range(P):
            src = (r - s) % P
            ops.dense_attn_bwd(shard(q, r), shard(k, src), shard(v, src), o_bf16, shard(d_o, r), lses[r], 0.125,
                               dq=dq[r], dk=dk[src], dv=dv[src], accumulate_dkv=True, accumulate_dq=True)
    assert max_rel(torch.cat(dq, 1).cpu(), gq.cpu()) < 2e-2
    assert max_rel(torch.cat(dk, 1).cpu(), gk.cpu()) < 2e-2
    assert max_rel(torch.cat(dv, 1).cpu(), gv.cpu()) < 2e-2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, queue):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from long_context_biomedical_imaging_b200 import ring

        torch.manual_seed(0)
        B, N, H, d = 1, 256 * world, 2, 64
        qkv = torch.randn(B, N, 3 * H * d, device="cuda").to(torch.bfloat16)
        d_out = torch.randn(B, N, H * d, device="cuda").to(torch.bfloat16)
        n = N // world
        local = qkv[:, rank * n:(rank + 1) * n].clone().requires_grad_(True)
        out = ring.ring_attention_qkv(local, H, ring.RingComm())
        out.backward(d_out[:, rank * n:(rank + 1) * n])
        full = qkv.float().requires_grad_(True)
        qf, kf, vf = ao.split_qkv_vit(full, H)
        ref = ao.dense_attention(qf, kf, vf, d ** -0.5).permute(0, 2, 1, 3).reshape(B, N, H * d)
        ref.backward(d_out.float())
        e_o = max_rel(out.detach().float().cpu(), ref[:, rank * n:(rank + 1) * n].detach().cpu())
        e_g = max_rel(local.grad.float().cpu(), full.grad[:, rank * n:(rank + 1) * n].cpu())
        queue.put((rank, e_o, e_g))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_ring_two_gpus():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, e_o, e_g in results:
        assert e_o < 2e-2 and e_g < 2e-2, (rank, e_o, e_g)
