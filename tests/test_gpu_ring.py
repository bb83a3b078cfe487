"""GPU: the ring-attention building blocks on ONE device — the P ranks are emulated as a loop over shards (a ring step
never waits on another rank's kernel, so no concurrency is needed): the fused forward carrying the online-softmax
state from shard to shard (lcbi_dense_attn_fwd_state), the log-sum-exp merge kernel, fp32 dq / travelling dk,dv
accumulation modes of the backward. Must equal global attention, up to the full cfg5 size (sampled rows).
With >= 2 GPUs the real NCCL ring is also run (and bench.py runs it as `parity_check` at every N > 1)."""
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
        # (a) separate launches merged with the log-sum-exp kernel, (b) one state carried through the launches
        acc = torch.empty(B, n, H, d, device="cuda")
        lse = torch.empty(B, H, n, device="cuda")
        state = (torch.empty(B, n, H, d, device="cuda"), torch.empty(B, H, n, device="cuda"),
                 torch.empty(B, H, n, device="cuda"))
        for s in range(P):
            src = (r - s) % P
            o_s, lse_s = ops.dense_attn_fwd(shard(q, r), shard(k, src), shard(v, src), 0.125)
            ops.attn_merge(acc, lse, o_s, lse_s, s == 0)
            o_c, lse_c = ops.dense_attn_fwd_state(shard(q, r), shard(k, src), shard(v, src), 0.125, state, s == 0,
                                                  s == P - 1)
        assert max_rel(lse_c.cpu(), lse.cpu()) < 1e-5
        assert max_rel(o_c.float().cpu(), acc.cpu()) < 1e-2
        outs.append(o_c.float())
        lses.append(lse_c)
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


def test_cfg5_size_ring_sampled_rows():
    """BASELINE configs[4] at full size: 262,144 tokens, 12 heads x 64, ring of 8 shards of 32,768 tokens emulated on
    one GPU (64 forward and 64 backward launches of the 32,768 x 32,768 step). The N x N matrix cannot be materialised,
    so sampled rows are compared with the oracle's chunked fp64 evaluation (oracle.attention_oracle.dense_attention_rows):
    o, lse and dq rows fully independently; dk / dv rows with the (row-checked) lse and D = rowsum(dO o O) of the run."""
    from long_context_biomedical_imaging_b200 import ops

    torch.manual_seed(5)
    P, n, H, d = 8, 32768, 12, 64
    N = P * n
    scale = d ** -0.5
    qkv = torch.randn(1, N, 3, H, d, device="cuda").to(torch.bfloat16)
    d_o = torch.randn(1, N, H, d, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    sl = lambda r: slice(r * n, (r + 1) * n)
    o = torch.empty(1, N, H, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(1, H, N, device="cuda")
    state = (torch.empty(1, n, H, d, device="cuda"), torch.empty(1, H, n, device="cuda"), torch.empty(1, H, n, device="cuda"))
    for r in range(P):
        for s in range(P):
            src = (r - s) % P
            res = ops.dense_attn_fwd_state(q[:, sl(r)], k[:, sl(src)], v[:, sl(src)], scale, state, s == 0, s == P - 1)
        o[:, sl(r)] = res[0]
        lse[:, :, sl(r)] = res[1]
    dq = torch.zeros(1, N, H, d, device="cuda")
    dk = torch.zeros(1, N, H, d, device="cuda")
    dv = torch.zeros(1, N, H, d, device="cuda")
    for r in range(P):
        lse_r = lse[:, :, sl(r)].contiguous()
        for s in range(P):
            src = (r - s) % P
            ops.dense_attn_bwd(q[:, sl(r)], k[:, sl(src)], v[:, sl(src)], o[:, sl(r)], d_o[:, sl(r)], lse_r, scale,
                               dq=dq[:, sl(r)], dk=dk[:, sl(src)], dv=dv[:, sl(src)], accumulate_dkv=True, accumulate_dq=True)
    torch.cuda.synchronize()

    gen = torch.Generator().manual_seed(0)
    rows = torch.randint(0, N, (32,), generator=gen)
    for h in (0, 7):
        qh, kh, vh, doh = [t[0, :, h].float().cpu() for t in (q, k, v, d_o)]
        # ---- query rows: o, lse, dq
        o_ref, lse_ref = ao.dense_attention_rows(qh[rows], kh, vh, scale)
        assert max_rel(o[0, rows.cuda(), h].float().cpu(), o_ref) < 2e-2, h
        assert max_rel(lse[0, h, rows.cuda()].cpu(), lse_ref) < 1e-4, h
        s_rows = (qh[rows].double() @ kh.double().T) * scale
        p_rows = torch.exp(s_rows - lse_ref[:, None])
        dp_rows = doh[rows].double() @ vh.double().T
        D_rows = (doh[rows].double() * o_ref).sum(-1, keepdim=True)
        dq_ref = scale * ((p_rows * (dp_rows - D_rows)) @ kh.double())
        assert max_rel(dq[0, rows.cuda(), h].cpu(), dq_ref) < 2e-2, h
        # ---- key rows: dk, dv over ALL queries, with the run's lse and D
        lse_all = lse[0, h].double().cpu()
        D_all = (d_o[0, :, h].float() * o[0, :, h].float()).sum(-1).double().cpu()
        dk_ref = torch.zeros(len(rows), d, dtype=torch.float64)
        dv_ref = torch.zeros(len(rows), d, dtype=torch.float64)
        for c in range(0, N, 32768):
            qc, doc = qh[c:c + 32768].double(), doh[c:c + 32768].double()
            pt = torch.exp((kh[rows].double() @ qc.T) * scale - lse_all[None, c:c + 32768])       # (rows, chunk)
            dpt = vh[rows].double() @ doc.T
            dv_ref += pt @ doc
            dk_ref += scale * ((pt * (dpt - D_all[None, c:c + 32768])) @ qc)
        assert max_rel(dk[0, rows.cuda(), h].cpu(), dk_ref) < 2e-2, h
        assert max_rel(dv[0, rows.cuda(), h].cpu(), dv_ref) < 2e-2, h


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
