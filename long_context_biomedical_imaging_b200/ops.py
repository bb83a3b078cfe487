"""Host-side operators over the C ABI: raw launch wrappers and torch.autograd.Functions.

PyTorch is used for device memory, streams and autograd plumbing only; every attention FLOP runs in the
hand-written sm_100a kernels of liblcbi_b200.so. CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("lcbi_b200 operators need CUDA tensors: there is no CPU fallback for the attention hot path")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _bnhd_strides(t):
    """(batch, row, head) element strides of a (B, N, H, d) view; d must be contiguous."""
    if t.dim() != 4 or t.stride(3) != 1:
        raise ValueError("expected a (B, N, H, d) view with contiguous head_dim")
    return _lib.strides3(t.stride(0), t.stride(1), t.stride(2))


# --------------------------------------------------------------------------------------------------
# dense attention (ViT)
# --------------------------------------------------------------------------------------------------
def dense_attn_fwd(q, k, v, scale, out=None):
    """q: (B,Nq,H,64), k/v: (B,Nk,H,64) bf16 views. Returns (o (B,Nq,H,64) bf16, lse (B,H,Nq) fp32)."""
    _require_cuda(q, k, v)
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise ValueError("dense_attn_fwd expects bf16 tensors")
    B, Nq, H, d = q.shape
    Nk = k.shape[1]
    if k.shape != (B, Nk, H, d) or v.shape != (B, Nk, H, d):
        raise ValueError("q/k/v shape mismatch")
    o = out if out is not None else torch.empty((B, Nq, H, d), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    rc = lib.lcbi_dense_attn_fwd(_p(q), _p(k), _p(v), _p(o), _p(lse), B, H, Nq, Nk, d, _bnhd_strides(q),
                                 _bnhd_strides(k), _bnhd_strides(v), _bnhd_strides(o), float(scale), _stream())
    _lib.check(rc, "lcbi_dense_attn_fwd")
    return o, lse


def dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=None, dk=None, dv=None, accumulate_dkv=False):
    """Gradients of dense_attn_fwd. With accumulate_dkv, dk/dv must be fp32 contiguous (B,Nk,H,d) buffers
    that are accumulated into (ring steps)."""
    _require_cuda(q, k, v, o, d_o, lse)
    B, Nq, H, d = q.shape
    Nk = k.shape[1]
    dev = q.device
    if dq is None:
        dq = torch.empty((B, Nq, H, d), dtype=torch.bfloat16, device=dev)
    if accumulate_dkv:
        if dk is None or dv is None or dk.dtype != torch.float32 or not dk.is_contiguous() or not dv.is_contiguous():
            raise ValueError("accumulate_dkv needs contiguous fp32 dk/dv accumulators")
    else:
        if dk is None:
            dk = torch.empty((B, Nk, H, d), dtype=torch.bfloat16, device=dev)
        if dv is None:
            dv = torch.empty((B, Nk, H, d), dtype=torch.bfloat16, device=dev)
    lib = _lib.load()
    ws_bytes = lib.lcbi_dense_attn_bwd_workspace_bytes(B, H, Nq, d)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    rc = lib.lcbi_dense_attn_bwd(_p(q), _p(k), _p(v), _p(o), _p(d_o), _p(lse), _p(dq), _p(dk), _p(dv), B, H, Nq, Nk, d,
                                 _bnhd_strides(q), _bnhd_strides(k), _bnhd_strides(v), _bnhd_strides(o),
                                 _bnhd_strides(d_o), _bnhd_strides(dq), _bnhd_strides(dk), _bnhd_strides(dv),
                                 float(scale), 1 if accumulate_dkv else 0, _p(ws), ws_bytes, _stream())
    _lib.check(rc, "lcbi_dense_attn_bwd")
    return dq, dk, dv


class _DenseAttentionQKV(torch.autograd.Function):
    """qkv: (B, N, 3, H, d) bf16 (the qkv Linear output viewed in place) -> o: (B, N, H*d)."""

    @staticmethod
    def forward(ctx, qkv, scale):
        qkv = qkv.contiguous()
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        o, lse = dense_attn_fwd(q, k, v, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.scale = scale
        B, N, H, d = o.shape
        return o.view(B, N, H * d)

    @staticmethod
    def backward(ctx, d_out):
        qkv, o, lse = ctx.saved_tensors
        B, N, _, H, d = qkv.shape
        d_o = d_out.to(torch.bfloat16).contiguous().view(B, N, H, d)
        dqkv = torch.empty_like(qkv)
        dense_attn_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, d_o, lse, ctx.scale, dq=dqkv[:, :, 0],
                       dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
        return dqkv, None


def dense_attention_qkv(qkv, num_heads, scale=None):
    """Drop-in for the attention core of the reference's SABlock.forward (backbone_vit.py:191-201).
    qkv: (B, N, 3*C) output of the qkv Linear, feature index = s*C + h*d + j. Returns (B, N, C) in qkv.dtype."""
    _require_cuda(qkv)
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // num_heads
    if scale is None:
        scale = d ** -0.5
    x = qkv if qkv.dtype == torch.bfloat16 else qkv.to(torch.bfloat16)
    o = _DenseAttentionQKV.apply(x.view(B, N, 3, num_heads, d), float(scale))
    return o if o.dtype == qkv.dtype else o.to(qkv.dtype)
