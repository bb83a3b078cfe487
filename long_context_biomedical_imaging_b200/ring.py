"""Sequence-parallel (ring) global attention for very long ViT sequences on one multi-GPU box.

New functionality relative to the reference, which has data parallelism only and materialises the N x N matrix
(backbone_vit.py:193), so it cannot run configs[4] (262,144 tokens) at all. Rank r owns tokens
[r*N/P, (r+1)*N/P) of q, k, v. The K/V shards rotate around the ring (rank r sends to r+1, receives from r-1) with
`torch.distributed` point-to-point ops over NCCL/NVLink on a side stream, double-buffered, while the fused tcgen05
attention kernel processes the shard that is already resident:

  forward   P launches of the fused kernel on (q_local, k_visiting, v_visiting) that CARRY the online-softmax state
            (running max, running sum, un-normalised fp32 output) from launch to launch (`lcbi_dense_attn_fwd_state`):
            no per-step merge pass and no fp32 round trip of partial outputs; the last launch normalises and writes
            the bf16 output and the log-sum-exp;
  backward  P launches of the flash backward on (q_local, k_visiting, v_visiting) with the FINAL lse: dq accumulates
            locally in fp32; the fp32 (dk, dv) partial sums TRAVEL with their K/V shard, and are added on arrival: a
            step's kernel accumulates this rank's contribution into a zeroed buffer while the running sum of the same
            shard is still in flight from the previous rank, the two are added afterwards and sent on while the NEXT
            step computes. Only the last hop (the sums returning home) is exposed.

The attention callables are injectable so that the rotation and bookkeeping can be exercised on CPU with the gloo
backend (tests/test_ring_gloo.py); the defaults are the CUDA kernels and there is no fallback.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist

from . import ops


def _cuda_fwd_state(q, k, v, scale, state, first, last, out, lse):
    ops.dense_attn_fwd_state(q, k, v, scale, state, first, last, out=out, lse=lse)


def _cuda_bwd(q, k, v, o, d_o, lse, scale, dq_acc, dk_acc, dv_acc):
    ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=dq_acc, dk=dk_acc, dv=dv_acc, accumulate_dkv=True,
                       accumulate_dq=True)


@contextlib.contextmanager
def _leave_sms_for_comm(comm, device):
    """The dense kernels are persistent (they hold every SM for a whole launch), so the NCCL send/recv kernels of the
    side stream may only start once a launch retires and the exchange does not overlap the compute. While a ring
    pass runs, the kernels leave `comm.reserved_sms` SMs of THIS device unused (`lcbi_set_reserved_sms`; env
    LCBI_RING_RESERVED_SMS; the default depends on the ring size, see RingComm); the previous setting is restored."""
    n = comm.reserved_sms if (comm.world > 1 and device.type == "cuda") else 0
    if not n:
        yield
        return
    from . import _lib

    lib = _lib.load()
    with torch.cuda.device(device):
        previous = lib.lcbi_get_reserved_sms()
        _lib.check(lib.lcbi_set_reserved_sms(n), "lcbi_set_reserved_sms")
    try:
        yield
    finally:
        with torch.cuda.device(device):
            lib.lcbi_set_reserved_sms(previous)


class RingComm:
    """Double-buffered neighbour exchange: send tensors to rank+1 and receive from rank-1 on a side stream.

    `exchange` returns a ticket; `wait(ticket)` makes the current stream wait for exactly that exchange, so that two
    exchanges (K/V prefetch and the travelling dK/dV sums) can be in flight at once."""

    def __init__(self, group=None, reserved_sms=None):
        self.group = group
        world = dist.get_world_size(group)
        if reserved_sms is None:
            # The exchange kernels need SMs of their own beside the persistent attention kernels. Measured at cfg5
            # (round 1, default NCCL channel count): 8 ranks 134.0 ms with 0, 122.7 ms with 8, 128.7 ms with 16 reserved
            # SMs; 2 ranks 459.8 / 474.1 / 490.7 ms. A ring hop moves 100-300 MB inside a multi-millisecond step
            # (< 40 GB/s needed), so a few channels suffice: with NCCL_MAX_P2P_NCHANNELS capped (bench.py sets 4)
            # as many SMs are enough.
            cap = os.environ.get("NCCL_MAX_P2P_NCHANNELS")
            default = 8 if world >= 8 else (4 if world >= 4 else 0)
            if cap is not None and world >= 4:
                default = max(2, min(default, int(cap)))
            reserved_sms = int(os.environ.get("LCBI_RING_RESERVED_SMS", default))
        self.reserved_sms = int(reserved_sms)
        self.rank = dist.get_rank(group)
        self.world = world
        nxt, prv = (self.rank + 1) % world, (self.rank - 1) % world
        self.send_to = dist.get_global_rank(group, nxt) if group is not None else nxt
        self.recv_from = dist.get_global_rank(group, prv) if group is not None else prv
        self._stream = None

    def _side_stream(self, device):
        if device.type != "cuda":
            return None
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=device)
        return self._stream

    def exchange(self, send_tensors, recv_tensors):
        """Starts sending `send_tensors` to the next rank and receiving into `recv_tensors` from the previous one.
        Returns a ticket for `wait`."""
        if self.world == 1:
            for s, r in zip(send_tensors, recv_tensors):
                r.copy_(s)
            return None
        p2p = []
        for s, r in zip(send_tensors, recv_tensors):
            p2p.append(dist.P2POp(dist.isend, s, self.send_to, self.group))
            p2p.append(dist.P2POp(dist.irecv, r, self.recv_from, self.group))
        device = send_tensors[0].device
        side = self._side_stream(device)
        if side is None:
            return (dist.batch_isend_irecv(p2p), None)
        side.wait_stream(torch.cuda.current_stream(device))   # the buffers being sent are complete
        with torch.cuda.stream(side):
            reqs = dist.batch_isend_irecv(p2p)
            for r in reqs:
                r.wait()                                      # stream-ordered on the side stream, not a host block
            done = torch.cuda.Event()
            done.record(side)
        return (None, done)

    def wait(self, ticket):
        if ticket is None:
            return
        reqs, done = ticket
        if reqs is not None:
            for r in reqs:
                r.wait()
        if done is not None:
            torch.cuda.current_stream().wait_event(done)


def ring_attention_forward(q, k, v, scale, comm, fwd_state_fn=None):
    """q,k,v: local shards (B, N_local, H, d). Returns (out (B,N_local,H,d) in q.dtype, lse fp32 (B,H,N_local))."""
    fwd_state_fn = fwd_state_fn or _cuda_fwd_state
    P = comm.world
    B, N, H, d = q.shape
    k_cur, v_cur = k.contiguous(), v.contiguous()
    k_nxt, v_nxt = (torch.empty_like(k_cur), torch.empty_like(v_cur)) if P > 1 else (None, None)
    out = torch.empty((B, N, H, d), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=q.device)
    # carried state: un-normalised fp32 output, running max, running sum (untouched when P == 1)
    state = (torch.empty((B, N, H, d), dtype=torch.float32, device=q.device),
             torch.empty((B, H, N), dtype=torch.float32, device=q.device),
             torch.empty((B, H, N), dtype=torch.float32, device=q.device))
    with _leave_sms_for_comm(comm, q.device):
        for step in range(P):
            ticket = comm.exchange([k_cur, v_cur], [k_nxt, v_nxt]) if step + 1 < P else None
            fwd_state_fn(q, k_cur, v_cur, scale, state, step == 0, step == P - 1, out, lse)
            if step + 1 < P:
                comm.wait(ticket)
                k_cur, k_nxt = k_nxt, k_cur
                v_cur, v_nxt = v_nxt, v_cur
    return out, lse


def ring_attention_backward(q, k, v, o, d_o, lse, scale, comm, bwd_fn=None):
    """Gradients for the local shards. o: the forward output, lse: final (B,H,N_local).
    Returns (dq, dk, dv) in fp32, each (B, N_local, H, d)."""
    bwd_fn = bwd_fn or _cuda_bwd
    P = comm.world
    dev = q.device
    k_cur, v_cur = k.contiguous(), v.contiguous()
    dq = torch.zeros(q.shape, dtype=torch.float32, device=dev)
    if P == 1:
        dk, dv = torch.zeros(k.shape, dtype=torch.float32, device=dev), torch.zeros(k.shape, dtype=torch.float32, device=dev)
        bwd_fn(q, k_cur, v_cur, o, d_o, lse, scale, dq, dk, dv)
        return dq, dk, dv
    f32 = dict(dtype=torch.float32, device=dev)
    k_nxt, v_nxt = torch.empty_like(k_cur), torch.empty_like(v_cur)
    # (dk, dv) live stacked in one tensor so that a hop is one send: `local` = this step's contribution, `recv` = the
    # running sum of the visiting shard arriving from the previous rank, `send` = their total on its way to the next rank
    local = torch.empty((2,) + tuple(k.shape), **f32)
    recv = torch.empty_like(local)
    send = torch.empty_like(local)
    t_sum = None
    with _leave_sms_for_comm(comm, dev):
        for step in range(P):
            t_kv = comm.exchange([k_cur, v_cur], [k_nxt, v_nxt]) if step + 1 < P else None
            local.zero_()
            bwd_fn(q, k_cur, v_cur, o, d_o, lse, scale, dq, local[0], local[1])
            if step > 0:
                comm.wait(t_sum)                      # the visiting shard's running sum (sent one step ago) is here
                torch.add(local, recv, out=local)
            send, local = local, send                 # `send` now holds the total; its old storage is free again:
            t_sum = comm.exchange([send], [recv])     # the previous send out of it completed before `recv` arrived
            if step + 1 < P:
                comm.wait(t_kv)
                k_cur, k_nxt = k_nxt, k_cur
                v_cur, v_nxt = v_nxt, v_cur
        comm.wait(t_sum)                              # last hop: the sums of this rank's own shard come home
    return dq, recv[0], recv[1]


class _RingAttentionQKV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, scale, comm):
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        o, lse = ring_attention_forward(q, k, v, scale, comm)
        ctx.save_for_backward(qkv, o, lse)
        ctx.scale, ctx.comm = scale, comm
        B, N, H, d = o.shape
        return o.view(B, N, H * d)

    @staticmethod
    def backward(ctx, d_out):
        qkv, o, lse = ctx.saved_tensors
        B, N, _, H, d = qkv.shape
        d_o = d_out.to(qkv.dtype).contiguous().view(B, N, H, d)
        dq, dk, dv = ring_attention_backward(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, d_o, lse, ctx.scale, ctx.comm)
        dqkv = torch.empty_like(qkv)
        dqkv[:, :, 0].copy_(dq)
        dqkv[:, :, 1].copy_(dk)
        dqkv[:, :, 2].copy_(dv)
        return dqkv, None, None


def ring_attention_qkv(qkv, num_heads, comm, scale=None):
    """Sequence-parallel counterpart of ops.dense_attention_qkv: qkv is the LOCAL (B, N/P, 3*C) slice of the qkv
    Linear output (token-wise layers need no communication); returns the local (B, N/P, C) attention output."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // num_heads
    if scale is None:
        scale = d ** -0.5
    # the CUDA kernels compute in bf16; CPU tensors only occur in the gloo tests, whose injected math keeps their dtype
    x = qkv if (qkv.dtype == torch.bfloat16 or not qkv.is_cuda) else qkv.to(torch.bfloat16)
    o = _RingAttentionQKV.apply(x.view(B, N, 3, num_heads, d).contiguous(), float(scale), comm)
    return o if o.dtype == qkv.dtype else o.to(qkv.dtype)


# --------------------------------------------------------------------------------------------------
# sequence sharding helpers for the encoder modules (backbone_vit.py `sequence_group`)
# --------------------------------------------------------------------------------------------------
def shard_bounds(n_tokens, world, rank):
    """Contiguous token range [begin, end) of `rank`; the ring needs equal shards, so n_tokens % world must be 0."""
    if n_tokens % world != 0:
        raise ValueError(f"sequence-parallel attention needs a token count divisible by the group size "
                         f"({n_tokens} tokens, {world} ranks)")
    per = n_tokens // world
    return rank * per, (rank + 1) * per


class _GatherSequence(torch.autograd.Function):
    """(B, N/P, C) local token slices -> (B, N, C) on every rank. The consumer (a decoder head) is REPLICATED: every
    rank evaluates the same loss on the same gathered tensor, so the gradient of that one logical loss with respect to
    this rank's tokens is the local slice of the incoming gradient (no reduction: summing the P identical copies would
    scale it by P)."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        ctx.group, ctx.rank, ctx.n = group, dist.get_rank(group), x.shape[1]
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous(), group=group)
        return torch.cat(parts, dim=1)

    @staticmethod
    def backward(ctx, g):
        return g[:, ctx.rank * ctx.n:(ctx.rank + 1) * ctx.n].contiguous(), None


def gather_sequence(x, group=None):
    """Differentiable all-gather of the local token slices along dim 1 (hidden states handed to a replicated decoder)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _GatherSequence.apply(x, group)
