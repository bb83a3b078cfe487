"""Sequence-parallel (ring) global attention for very long ViT sequences on one multi-GPU box.

New functionality relative to the reference, which has data parallelism only and materialises the N x N matrix
(backbone_vit.py:193), so it cannot run configs[4] (262,144 tokens) at all. Rank r owns tokens
[r*N/P, (r+1)*N/P) of q, k, v. The K/V shards rotate around the ring (rank r sends to r+1, receives from r-1) with
`torch.distributed` point-to-point ops over NCCL/NVLink on a side stream, double-buffered, while the fused tcgen05
attention kernel processes the shard that is already resident:

  forward   P steps of  (o_s, lse_s) = attn(q_local, k_visiting, v_visiting), merged with a log-sum-exp combine
            kernel into an fp32 running output;
  backward  P steps of  the flash backward on (q_local, k_visiting, v_visiting) with the FINAL lse: dq accumulates
            locally in fp32, the fp32 (dk, dv) accumulators travel with their K/V shard and arrive home after the
            last rotation.

The attention / merge callables are injectable so that the rotation and bookkeeping can be exercised on CPU with the
gloo backend (tests/test_ring_gloo.py); the defaults are the CUDA kernels and there is no fallback.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist

from . import ops


def _cuda_fwd(q, k, v, scale):
    return ops.dense_attn_fwd(q, k, v, scale)


def _cuda_bwd(q, k, v, o, d_o, lse, scale, dq_acc, dk_acc, dv_acc):
    ops.dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=dq_acc, dk=dk_acc, dv=dv_acc, accumulate_dkv=True,
                       accumulate_dq=True)


def _cuda_merge(acc, lse_acc, o_s, lse_s, first):
    ops.attn_merge(acc, lse_acc, o_s, lse_s, first)


@contextlib.contextmanager
def _leave_sms_for_comm(comm, device):
    """The dense kernels are persistent (they hold every SM for a whole launch), so the NCCL send/recv kernels of the
    side stream may only start once a launch retires and the exchange does not overlap the compute. While a ring
    pass runs, the kernels leave `comm.reserved_sms` SMs unused (`lcbi_set_reserved_sms`; env
    LCBI_RING_RESERVED_SMS; the default depends on the ring size, see RingComm)."""
    n = comm.reserved_sms if (comm.world > 1 and device.type == "cuda") else 0
    if n:
        from . import _lib
        _lib.check(_lib.load().lcbi_set_reserved_sms(n), "lcbi_set_reserved_sms")
    try:
        yield
    finally:
        if n:
            _lib.load().lcbi_set_reserved_sms(0)


class RingComm:
    """Double-buffered neighbour exchange: send tensors to rank+1 and receive from rank-1 on a side stream."""

    def __init__(self, group=None, reserved_sms=None):
        self.group = group
        world = dist.get_world_size(group)
        if reserved_sms is None:
            # measured at cfg5: 8 ranks 134.0 ms with 0, 122.7 ms with 8, 128.7 ms with 16 reserved SMs;
            # 2 ranks 459.8 / 474.1 / 490.7 ms (the exchange is small next to a step's compute there)
            default = 8 if world >= 8 else (4 if world >= 4 else 0)
            reserved_sms = int(os.environ.get("LCBI_RING_RESERVED_SMS", default))
        self.reserved_sms = int(reserved_sms)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.send_to = dist.get_global_rank(group, (self.rank + 1) % self.world) if group is not None else (self.rank + 1) % self.world
        self.recv_from = dist.get_global_rank(group, (self.rank - 1) % self.world) if group is not None else (self.rank - 1) % self.world
        self._reqs = []
        self._stream = None

    def _side_stream(self, device):
        if device.type != "cuda":
            return None
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=device)
        return self._stream

    def exchange(self, send_tensors, recv_tensors):
        """Starts sending `send_tensors` to the next rank and receiving into `recv_tensors` from the previous one."""
        if self.world == 1:
            for s, r in zip(send_tensors, recv_tensors):
                r.copy_(s)
            return
        p2p = []
        for s, r in zip(send_tensors, recv_tensors):
            p2p.append(dist.P2POp(dist.isend, s, self.send_to, self.group))
            p2p.append(dist.P2POp(dist.irecv, r, self.recv_from, self.group))
        side = self._side_stream(send_tensors[0].device)
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())   # the buffers being sent are complete
            with torch.cuda.stream(side):
                self._reqs = dist.batch_isend_irecv(p2p)
        else:
            self._reqs = dist.batch_isend_irecv(p2p)

    def wait(self):
        for r in self._reqs:
            r.wait()
        self._reqs = []
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)


def ring_attention_forward(q, k, v, scale, comm, fwd_fn=_cuda_fwd, merge_fn=_cuda_merge):
    """q,k,v: local shards (B, N_local, H, d). Returns (out fp32 (B,N_local,H,d), lse fp32 (B,H,N_local))."""
    P = comm.world
    k_cur, v_cur = k.contiguous(), v.contiguous()
    k_nxt, v_nxt = (torch.empty_like(k_cur), torch.empty_like(v_cur)) if P > 1 else (None, None)
    acc = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    lse = torch.empty((q.shape[0], q.shape[2], q.shape[1]), dtype=torch.float32, device=q.device)
    with _leave_sms_for_comm(comm, q.device):
        for step in range(P):
            if step + 1 < P:
                comm.exchange([k_cur, v_cur], [k_nxt, v_nxt])
            o_s, lse_s = fwd_fn(q, k_cur, v_cur, scale)
            merge_fn(acc, lse, o_s, lse_s, step == 0)
            if step + 1 < P:
                comm.wait()
                k_cur, k_nxt = k_nxt, k_cur
                v_cur, v_nxt = v_nxt, v_cur
    return acc, lse


def ring_attention_backward(q, k, v, o, d_o, lse, scale, comm, bwd_fn=_cuda_bwd):
    """Gradients for the local shards. o: the merged forward output (bf16 for the CUDA path), lse: final (B,H,N_local).
    Returns (dq, dk, dv) in fp32, each (B, N_local, H, d)."""
    P = comm.world
    k_cur, v_cur = k.contiguous(), v.contiguous()
    dq = torch.zeros(q.shape, dtype=torch.float32, device=q.device)
    dk_cur = torch.zeros(k.shape, dtype=torch.float32, device=q.device)
    dv_cur = torch.zeros(k.shape, dtype=torch.float32, device=q.device)
    if P > 1:
        k_nxt, v_nxt = torch.empty_like(k_cur), torch.empty_like(v_cur)
        dk_nxt, dv_nxt = torch.empty_like(dk_cur), torch.empty_like(dv_cur)
    with _leave_sms_for_comm(comm, q.device):
        for step in range(P):
            if step + 1 < P:
                comm.exchange([k_cur, v_cur], [k_nxt, v_nxt])      # K/V prefetch overlaps this step's compute
            bwd_fn(q, k_cur, v_cur, o, d_o, lse, scale, dq, dk_cur, dv_cur)
            if P > 1:
                if step + 1 < P:
                    comm.wait()
                comm.exchange([dk_cur, dv_cur], [dk_nxt, dv_nxt])  # accumulators follow their shard (last hop: home)
                comm.wait()
                dk_cur, dk_nxt = dk_nxt, dk_cur
                dv_cur, dv_nxt = dv_nxt, dv_cur
                if step + 1 < P:
                    k_cur, k_nxt = k_nxt, k_cur
                    v_cur, v_nxt = v_nxt, v_cur
    return dq, dk_cur, dv_cur


class _RingAttentionQKV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, scale, comm):
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        acc, lse = ring_attention_forward(q, k, v, scale, comm)
        o = acc.to(torch.bfloat16)
        ctx.save_for_backward(qkv, o, lse)
        ctx.scale, ctx.comm = scale, comm
        B, N, H, d = o.shape
        return o.view(B, N, H * d)

    @staticmethod
    def backward(ctx, d_out):
        qkv, o, lse = ctx.saved_tensors
        B, N, _, H, d = qkv.shape
        d_o = d_out.to(torch.bfloat16).contiguous().view(B, N, H, d)
        dq, dk, dv = ring_attention_backward(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, d_o, lse, ctx.scale, ctx.comm)
        return torch.stack((dq, dk, dv), dim=2).to(qkv.dtype), None, None


def ring_attention_qkv(qkv, num_heads, comm, scale=None):
    """Sequence-parallel counterpart of ops.dense_attention_qkv: qkv is the LOCAL (B, N/P, 3*C) slice of the qkv
    Linear output (token-wise layers need no communication); returns the local (B, N/P, C) attention output."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // num_heads
    if scale is None:
        scale = d ** -0.5
    x = qkv if qkv.dtype == torch.bfloat16 else qkv.to(torch.bfloat16)
    o = _RingAttentionQKV.apply(x.view(B, N, 3, num_heads, d).contiguous(), float(scale), comm)
    return o if o.dtype == qkv.dtype else o.to(qkv.dtype)
