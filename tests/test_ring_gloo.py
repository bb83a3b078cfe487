"""CPU, world_size 2 and 3 over gloo: the ring (sequence-parallel) attention driver — K/V rotation, the online-softmax
state carried from step to step, travelling dK/dV sums added on arrival — with the attention math supplied by the
oracle instead of the CUDA kernels.
The result must equal single-process global attention (forward and all gradients)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attention_oracle as ao


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fwd(q, k, v, scale):
    qh, kh, vh = [t.permute(0, 2, 1, 3) for t in (q, k, v)]
    s = torch.einsum("bhxd,bhyd->bhxy", qh, kh) * scale
    lse = torch.logsumexp(s, -1)
    return ao.dense_attention(qh, kh, vh, scale).permute(0, 2, 1, 3).contiguous(), lse


def _fwd_state(q, k, v, scale, state, first, last, out, lse):
    """CPU stand-in for lcbi_dense_attn_fwd_state: same contract (state = un-normalised output, running max in raw score
    units, running sum; `last` normalises into out / lse)."""
    st_o, st_m, st_l = state
    qh, kh, vh = [t.permute(0, 2, 1, 3) for t in (q, k, v)]
    s = torch.einsum("bhxd,bhyd->bhxy", qh, kh)                      # raw scores (B,H,Nq,Nk)
    m_new = s.max(-1).values
    if not first:
        m_new = torch.maximum(m_new, st_m)
    p = torch.exp((s - m_new.unsqueeze(-1)) * scale)
    o_part = torch.einsum("bhxy,bhyd->bhxd", p, vh).permute(0, 2, 1, 3)
    if first:
        st_o.copy_(o_part)
        st_l.copy_(p.sum(-1))
    else:
        alpha = torch.exp((st_m - m_new) * scale)
        st_o.copy_(st_o * alpha.permute(0, 2, 1).unsqueeze(-1) + o_part)
        st_l.copy_(st_l * alpha + p.sum(-1))
    st_m.copy_(m_new)
    if last:
        out.copy_(st_o / st_l.permute(0, 2, 1).unsqueeze(-1))
        lse.copy_(st_m * scale + torch.log(st_l))


def _bwd(q, k, v, o, d_o, lse, scale, dq, dk, dv):
    qh, kh, vh, oh, doh = [t.permute(0, 2, 1, 3).float() for t in (q, k, v, o, d_o)]
    p = torch.exp(torch.einsum("bhxd,bhyd->bhxy", qh, kh) * scale - lse.unsqueeze(-1))
    dsum = (oh * doh).sum(-1, keepdim=True)
    ds = p * (torch.einsum("bhxd,bhyd->bhxy", doh, vh) - dsum)
    dq += (scale * torch.einsum("bhxy,bhyd->bhxd", ds, kh)).permute(0, 2, 1, 3)
    dk += (scale * torch.einsum("bhxy,bhxd->bhyd", ds, qh)).permute(0, 2, 1, 3)
    dv += torch.einsum("bhxy,bhxd->bhyd", p, doh).permute(0, 2, 1, 3)


def _worker(rank, world, port, n_local, result_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from long_context_biomedical_imaging_b200 import ring

        torch.manual_seed(0)
        B, H, d = 2, 2, 8
        N = n_local * world
        q, k, v, d_o = [torch.randn(B, N, H, d) for _ in range(4)]
        sl = slice(rank * n_local, (rank + 1) * n_local)
        comm = ring.RingComm()
        assert comm.world == world
        acc, lse = ring.ring_attention_forward(q[:, sl], k[:, sl], v[:, sl], 0.3, comm, fwd_state_fn=_fwd_state)
        dq, dk, dv = ring.ring_attention_backward(q[:, sl], k[:, sl], v[:, sl], acc, d_o[:, sl], lse, 0.3, comm,
                                                  bwd_fn=_bwd)
        # single-process reference
        qf, kf, vf = [t.clone().requires_grad_(True) for t in (q, k, v)]
        ref = ao.dense_attention(qf.permute(0, 2, 1, 3), kf.permute(0, 2, 1, 3), vf.permute(0, 2, 1, 3), 0.3)
        ref = ref.permute(0, 2, 1, 3)
        gq, gk, gv = torch.autograd.grad(ref, [qf, kf, vf], d_o)
        errs = [float((acc - ref[:, sl].detach()).abs().max()), float((dq - gq[:, sl]).abs().max()),
                float((dk - gk[:, sl]).abs().max()), float((dv - gv[:, sl]).abs().max())]
        result_queue.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_local", [(2, 5), (3, 4)])
def test_ring_attention_matches_global_attention(world, n_local):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in results) == list(range(world))
    for rank, errs in results:
        assert max(errs) < 1e-5, (rank, errs)


def test_ring_single_rank_degenerates_to_one_step():
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from long_context_biomedical_imaging_b200 import ring

        torch.manual_seed(1)
        q, k, v = [torch.randn(1, 7, 2, 8) for _ in range(3)]
        acc, lse = ring.ring_attention_forward(q, k, v, 0.5, ring.RingComm(), fwd_state_fn=_fwd_state)
        ref, ref_lse = _fwd(q, k, v, 0.5)
        assert torch.allclose(acc, ref, atol=1e-6) and torch.allclose(lse, ref_lse, atol=1e-6)
    finally:
        dist.destroy_process_group()
