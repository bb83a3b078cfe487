"""Host-side operators over the C ABI: raw launch wrappers and torch.autograd.Functions.

PyTorch is used for device memory, streams and autograd plumbing only; every attention FLOP runs in the
hand-written sm_100a kernels of liblcbi_b200.so. CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _require_cuda(*tensors):
    """All operands must live on ONE CUDA device. Returns that device (None when every operand is None)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("lcbi_b200 operators need CUDA tensors: there is no CPU fallback for the attention hot path")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"lcbi_b200 operators need all operands on one device, got {dev} and {t.device}")
    return dev


def _on(dev):
    """Device guard for a launch: the library reads the SM count, its per-device launch caches and the TMA encode from
    the CURRENT device, so the operands' device is made current for the duration of the call."""
    return torch.cuda.device(dev)


def _stream(dev):
    """The caller's current stream ON THE OPERANDS' DEVICE (not the process-wide current device's)."""
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


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
    dev = _require_cuda(q, k, v, out)
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise ValueError("dense_attn_fwd expects bf16 tensors")
    B, Nq, H, d = q.shape
    Nk = k.shape[1]
    if k.shape != (B, Nk, H, d) or v.shape != (B, Nk, H, d):
        raise ValueError("q/k/v shape mismatch")
    if out is not None and (out.dtype != torch.bfloat16 or out.shape != (B, Nq, H, d)):
        raise ValueError("dense_attn_fwd: `out` must be a bf16 (B, Nq, H, d) view")
    o = out if out is not None else torch.empty((B, Nq, H, d), dtype=torch.bfloat16, device=dev)
    lse = torch.empty((B, H, Nq), dtype=torch.float32, device=dev)
    lib = _lib.load()
    with _on(dev):
        rc = lib.lcbi_dense_attn_fwd(_p(q), _p(k), _p(v), _p(o), _p(lse), B, H, Nq, Nk, d, _bnhd_strides(q),
                                     _bnhd_strides(k), _bnhd_strides(v), _bnhd_strides(o), float(scale), _stream(dev))
    _lib.check(rc, "lcbi_dense_attn_fwd")
    return o, lse


def dense_attn_fwd_state(q, k, v, scale, state, first, last, out=None, lse=None):
    """One ring step of the forward (C ABI lcbi_dense_attn_fwd_state): folds the K/V shard (k, v) into the carried
    online-softmax state of the local query rows. state = (o_acc fp32 (B,Nq,H,64), m fp32 (B,H,Nq), l fp32 (B,H,Nq)),
    all contiguous; `first` ignores the incoming state; `last` normalises and writes `out` (bf16) and `lse` (both are
    allocated when not given) and returns them, otherwise only the state is updated and (None, None) is returned."""
    st_o, st_m, st_l = state
    dev = _require_cuda(q, k, v, st_o, st_m, st_l, out, lse)
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise ValueError("dense_attn_fwd_state expects bf16 q/k/v")
    B, Nq, H, d = q.shape
    Nk = k.shape[1]
    if k.shape != (B, Nk, H, d) or v.shape != (B, Nk, H, d):
        raise ValueError("q/k/v shape mismatch")
    if (st_o.dtype != torch.float32 or tuple(st_o.shape) != (B, Nq, H, d) or not st_o.is_contiguous()):
        raise ValueError("dense_attn_fwd_state: state[0] must be a contiguous fp32 (B, Nq, H, d) tensor")
    for t in (st_m, st_l):
        if t.dtype != torch.float32 or tuple(t.shape) != (B, H, Nq) or not t.is_contiguous():
            raise ValueError("dense_attn_fwd_state: state[1], state[2] must be contiguous fp32 (B, H, Nq) tensors")
    if last:
        if out is None:
            out = torch.empty((B, Nq, H, d), dtype=torch.bfloat16, device=dev)
        if lse is None:
            lse = torch.empty((B, H, Nq), dtype=torch.float32, device=dev)
        if out.dtype != torch.bfloat16 or tuple(out.shape) != (B, Nq, H, d):
            raise ValueError("dense_attn_fwd_state: `out` must be a bf16 (B, Nq, H, d) view")
        if lse.dtype != torch.float32 or tuple(lse.shape) != (B, H, Nq) or not lse.is_contiguous():
            raise ValueError("dense_attn_fwd_state: `lse` must be a contiguous fp32 (B, H, Nq) tensor")
    lib = _lib.load()
    with _on(dev):
        rc = lib.lcbi_dense_attn_fwd_state(_p(q), _p(k), _p(v), _p(out) if last else None, _p(lse) if last else None, B, H,
                                           Nq, Nk, d, _bnhd_strides(q), _bnhd_strides(k), _bnhd_strides(v),
                                           _bnhd_strides(out) if last else None, float(scale), _p(st_o), _p(st_m),
                                           _p(st_l), 1 if first else 0, 1 if last else 0, _stream(dev))
    _lib.check(rc, "lcbi_dense_attn_fwd_state")
    return (out, lse) if last else (None, None)


def dense_attn_bwd(q, k, v, o, d_o, lse, scale, dq=None, dk=None, dv=None, accumulate_dkv=False, accumulate_dq=False):
    """Gradients of dense_attn_fwd. With accumulate_dkv (accumulate_dq), dk/dv (dq) must be fp32 contiguous
    (B,N,H,d) buffers that are accumulated into (ring steps)."""
    dev = _require_cuda(q, k, v, o, d_o, lse, dq, dk, dv)
    B, Nq, H, d = q.shape
    Nk = k.shape[1]
    for name, t, shape in (("q", q, (B, Nq, H, d)), ("k", k, (B, Nk, H, d)), ("v", v, (B, Nk, H, d)),
                           ("o", o, (B, Nq, H, d)), ("d_o", d_o, (B, Nq, H, d))):
        if t.dtype != torch.bfloat16 or tuple(t.shape) != shape:
            raise ValueError(f"dense_attn_bwd: {name} must be bf16 {shape}, got {t.dtype} {tuple(t.shape)}")
    if lse.dtype != torch.float32 or tuple(lse.shape) != (B, H, Nq) or not lse.is_contiguous():
        raise ValueError("dense_attn_bwd: lse must be a contiguous fp32 (B, H, Nq) tensor")

    def _check_grad(name, t, rows, accumulate):
        want = torch.float32 if accumulate else torch.bfloat16
        if t.dtype != want or tuple(t.shape) != (B, rows, H, d) or (accumulate and not t.is_contiguous()):
            raise ValueError(f"dense_attn_bwd: {name} must be a {'contiguous fp32 accumulator' if accumulate else 'bf16 view'} "
                             f"of shape {(B, rows, H, d)}, got {t.dtype} {tuple(t.shape)}")

    if accumulate_dq and dq is None:
        raise ValueError("accumulate_dq needs a contiguous fp32 dq accumulator")
    if accumulate_dkv and (dk is None or dv is None):
        raise ValueError("accumulate_dkv needs contiguous fp32 dk/dv accumulators")
    if dq is None:
        dq = torch.empty((B, Nq, H, d), dtype=torch.bfloat16, device=dev)
    if dk is None:
        dk = torch.empty((B, Nk, H, d), dtype=torch.bfloat16, device=dev)
    if dv is None:
        dv = torch.empty((B, Nk, H, d), dtype=torch.bfloat16, device=dev)
    _check_grad("dq", dq, Nq, accumulate_dq)
    _check_grad("dk", dk, Nk, accumulate_dkv)
    _check_grad("dv", dv, Nk, accumulate_dkv)
    lib = _lib.load()
    ws_bytes = lib.lcbi_dense_attn_bwd_workspace_bytes_for(B, H, Nq, d, 1 if accumulate_dq else 0)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with _on(dev):
        rc = lib.lcbi_dense_attn_bwd(_p(q), _p(k), _p(v), _p(o), _p(d_o), _p(lse), _p(dq), _p(dk), _p(dv), B, H, Nq, Nk,
                                     d, _bnhd_strides(q), _bnhd_strides(k), _bnhd_strides(v), _bnhd_strides(o),
                                     _bnhd_strides(d_o), _bnhd_strides(dq), _bnhd_strides(dk), _bnhd_strides(dv),
                                     float(scale), 1 if accumulate_dkv else 0, 1 if accumulate_dq else 0, _p(ws),
                                     ws_bytes, _stream(dev))
    _lib.check(rc, "lcbi_dense_attn_bwd")
    return dq, dk, dv


def attn_merge(acc, lse_acc, o_s, lse_s, first, out_bf16=None):
    """Ring-attention combine: merges the partial (o_s, lse_s) into the running fp32 (acc, lse_acc) in place."""
    dev = _require_cuda(acc, lse_acc, o_s, lse_s, out_bf16)
    B, N, H, d = acc.shape
    if not (acc.is_contiguous() and o_s.is_contiguous() and lse_acc.is_contiguous() and lse_s.is_contiguous()):
        raise ValueError("attn_merge expects contiguous tensors")
    if acc.dtype != torch.float32 or lse_acc.dtype != torch.float32 or lse_s.dtype != torch.float32:
        raise ValueError("attn_merge: acc, lse_acc and lse_s must be fp32")
    if o_s.dtype != torch.bfloat16 or tuple(o_s.shape) != (B, N, H, d):
        raise ValueError("attn_merge: o_s must be bf16 with acc's shape (B, N, H, d)")
    if tuple(lse_acc.shape) != (B, H, N) or tuple(lse_s.shape) != (B, H, N):
        raise ValueError("attn_merge: lse_acc and lse_s must be (B, H, N)")
    if out_bf16 is not None and (out_bf16.dtype != torch.bfloat16 or tuple(out_bf16.shape) != (B, N, H, d) or
                                 not out_bf16.is_contiguous()):
        raise ValueError("attn_merge: out_bf16 must be a contiguous bf16 (B, N, H, d) tensor")
    with _on(dev):
        rc = _lib.load().lcbi_attn_merge(_p(acc), _p(lse_acc), _p(o_s), _p(lse_s),
                                         _p(out_bf16) if out_bf16 is not None else None, B, N, H, d,
                                         1 if first else 0, _stream(dev))
    _lib.check(rc, "lcbi_attn_merge")


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


# --------------------------------------------------------------------------------------------------
# LayerNorm in front of the qkv projection / the MLP
# --------------------------------------------------------------------------------------------------
def _is_bf16(t):
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise ValueError(f"layer_norm handles fp32 and bf16 tensors, got {t.dtype}")


class _LayerNorm(torch.autograd.Function):
    """x: (..., C) fp32 / bf16 -> y in `out_dtype`; statistics in fp32 (lcbi_layer_norm_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
        with _on(x.device):
            rc = _lib.load().lcbi_layer_norm_fwd(_p(x), _is_bf16(x), _p(weight) if weight is not None else None,
                                                 _p(bias) if bias is not None else None, _p(y), _is_bf16(y), _p(mean),
                                                 _p(rstd), rows, C, float(eps), _stream(x.device))
        _lib.check(rc, "lcbi_layer_norm_fwd")
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        C = x.shape[-1]
        rows = x.numel() // C
        dy = dy.contiguous()
        need_dx = ctx.needs_input_grad[0]
        need_w = weight is not None and ctx.needs_input_grad[1]
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_dx else None
        dw = torch.empty((C,), dtype=torch.float32, device=x.device) if need_w else None
        db = torch.empty((C,), dtype=torch.float32, device=x.device) if need_b else None
        lib = _lib.load()
        ws, ws_bytes = None, 0
        if need_w or need_b:
            ws_bytes = lib.lcbi_layer_norm_bwd_workspace_bytes(rows, C)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
        if need_dx or need_w or need_b:
            with _on(x.device):
                rc = lib.lcbi_layer_norm_bwd(_p(dy), _is_bf16(dy), _p(x), _is_bf16(x),
                                             _p(weight) if weight is not None else None, _p(mean), _p(rstd),
                                             _p(dx) if need_dx else None, _p(dw) if need_w else None,
                                             _p(db) if need_b else None, _p(ws) if ws is not None else None, ws_bytes,
                                             rows, C, _stream(x.device))
            _lib.check(rc, "lcbi_layer_norm_bwd")
        return dx, dw, db, None, None


def layer_norm(x, weight, bias, eps=1e-5, out_dtype=None):
    """Drop-in for `nn.LayerNorm(C)(x)` as the encoder blocks call it (reference backbone_vit.py:260-263,
    backbone_swin.py:437,489). `out_dtype` defaults to bf16 under bf16 autocast (what the following Linear would cast
    the fp32 result to anyway) and to x.dtype otherwise."""
    _require_cuda(x, weight, bias)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if out_dtype is None:
        amp_bf16 = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
        out_dtype = torch.bfloat16 if amp_bf16 else x.dtype
    if x.shape[-1] % 4 != 0:
        raise ValueError("layer_norm needs a channel count that is a multiple of 4")
    if x.numel() == 0:      # no rows: nothing to launch (an empty batch shard)
        return torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).to(out_dtype)
    if weight is not None and (weight.dtype != torch.float32 or not weight.is_contiguous()):
        weight = weight.float().contiguous()
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
        bias = bias.float().contiguous()
    return _LayerNorm.apply(x, weight, bias, float(eps), out_dtype)


class _AddLayerNorm(torch.autograd.Function):
    """(x, delta) -> (xsum = x + delta in x's dtype, y = LayerNorm(xsum) in `out_dtype`): the residual add of an encoder
    block fused into the LayerNorm that follows it (lcbi_add_layer_norm_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, delta, weight, bias, eps, out_dtype):
        x, delta = x.contiguous(), delta.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        xsum = torch.empty_like(x)
        y = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
        with _on(x.device):
            rc = _lib.load().lcbi_add_layer_norm_fwd(_p(x), _is_bf16(x), _p(delta), _is_bf16(delta), _p(xsum),
                                                     _p(weight) if weight is not None else None,
                                                     _p(bias) if bias is not None else None, _p(y), _is_bf16(y), _p(mean),
                                                     _p(rstd), rows, C, float(eps), _stream(x.device))
        _lib.check(rc, "lcbi_add_layer_norm_fwd")
        ctx.save_for_backward(xsum, weight, mean, rstd)
        ctx.has_bias = bias is not None
        ctx.delta_dtype = delta.dtype
        ctx.y_dtype = out_dtype
        return xsum, y

    @staticmethod
    def backward(ctx, dxsum, dy):
        xsum, weight, mean, rstd = ctx.saved_tensors
        C = xsum.shape[-1]
        rows = xsum.numel() // C
        if dy is None:
            dy = torch.zeros(xsum.shape, dtype=ctx.y_dtype, device=xsum.device)
        dy = dy.contiguous()
        if dxsum is not None:
            dxsum = dxsum.contiguous()
        need_w = weight is not None and ctx.needs_input_grad[2]
        need_b = ctx.has_bias and ctx.needs_input_grad[3]
        dx = torch.empty_like(xsum)
        dd = torch.empty(xsum.shape, dtype=ctx.delta_dtype, device=xsum.device) if ctx.needs_input_grad[1] else None
        dw = torch.empty((C,), dtype=torch.float32, device=xsum.device) if need_w else None
        db = torch.empty((C,), dtype=torch.float32, device=xsum.device) if need_b else None
        lib = _lib.load()
        ws, ws_bytes = None, 0
        if need_w or need_b:
            ws_bytes = lib.lcbi_layer_norm_bwd_workspace_bytes(rows, C)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=xsum.device)
        with _on(xsum.device):
            rc = lib.lcbi_add_layer_norm_bwd(_p(dy), _is_bf16(dy), _p(dxsum) if dxsum is not None else None, _p(xsum),
                                             _is_bf16(xsum), _p(weight) if weight is not None else None, _p(mean), _p(rstd),
                                             _p(dx), _p(dd) if dd is not None else None,
                                             1 if ctx.delta_dtype == torch.bfloat16 else 0, _p(dw) if need_w else None,
                                             _p(db) if need_b else None, _p(ws) if ws is not None else None, ws_bytes, rows,
                                             C, _stream(xsum.device))
        _lib.check(rc, "lcbi_add_layer_norm_bwd")
        return (dx if ctx.needs_input_grad[0] else None), dd, dw, db, None, None


def add_layer_norm(x, delta, weight, bias, eps=1e-5, out_dtype=None):
    """`x = x + delta; y = LayerNorm(x)` of a pre-norm block (reference backbone_vit.py:261-262 followed by the next
    norm) as one pass. Returns (x + delta, y). Falls back to the two separate steps for rows of < 128 channels or dtype
    combinations where torch's add would promote away from x's dtype."""
    _require_cuda(x, delta, weight, bias)
    fusable = (x.dtype in (torch.float32, torch.bfloat16) and delta.dtype in (torch.float32, torch.bfloat16) and
               torch.promote_types(x.dtype, delta.dtype) == x.dtype and x.shape == delta.shape and
               x.shape[-1] % 4 == 0 and x.shape[-1] >= 128 and x.numel() > 0)
    if not fusable:
        xsum = x + delta
        return xsum, layer_norm(xsum, weight, bias, eps, out_dtype)
    if out_dtype is None:
        amp_bf16 = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
        out_dtype = torch.bfloat16 if amp_bf16 else x.dtype
    if weight is not None and (weight.dtype != torch.float32 or not weight.is_contiguous()):
        weight = weight.float().contiguous()
    if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
        bias = bias.float().contiguous()
    return _AddLayerNorm.apply(x, delta, weight, bias, float(eps), out_dtype)


# --------------------------------------------------------------------------------------------------
# token-wise Linear with a bias: cuBLAS GEMMs, bias gradient through lcbi_bias_grad
# --------------------------------------------------------------------------------------------------
def bias_grad(dy):
    """Column sums of dy viewed as (rows, C): the bias gradient of a token-wise Linear. Returns fp32 (C)."""
    _require_cuda(dy)
    dy = dy.contiguous()
    C = dy.shape[-1]
    rows = dy.numel() // C
    if rows == 0:
        return torch.zeros((C,), dtype=torch.float32, device=dy.device)
    lib = _lib.load()
    out = torch.empty((C,), dtype=torch.float32, device=dy.device)
    ws_bytes = lib.lcbi_layer_norm_bwd_workspace_bytes(rows, C)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dy.device)
    with _on(dy.device):
        rc = lib.lcbi_bias_grad(_p(dy), _is_bf16(dy), _p(out), _p(ws), ws_bytes, rows, C, _stream(dy.device))
    _lib.check(rc, "lcbi_bias_grad")
    return out


class _LinearBias(torch.autograd.Function):
    """y = x W^T + b with the dtype flow of nn.Linear under autocast (operands in the compute dtype, fp32 accumulate in
    cuBLAS, output in the compute dtype). The backward runs the two GEMMs in the same dtype as autograd would and the
    bias gradient through lcbi_bias_grad instead of a generic reduction."""

    @staticmethod
    def forward(ctx, x, weight, bias, compute_dtype):
        xc = x if x.dtype == compute_dtype else x.to(compute_dtype)
        wc = weight if weight.dtype == compute_dtype else weight.to(compute_dtype)
        bc = bias if bias.dtype == compute_dtype else bias.to(compute_dtype)
        ctx.save_for_backward(xc, wc)
        ctx.in_dtypes = (x.dtype, weight.dtype, bias.dtype)
        return torch.nn.functional.linear(xc, wc, bc)

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        x_dtype, w_dtype, b_dtype = ctx.in_dtypes
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (dy2 @ wc).view(xc.shape).to(x_dtype)
        if ctx.needs_input_grad[1]:
            dw = (dy2.t() @ xc.reshape(-1, xc.shape[-1])).to(w_dtype)
        if ctx.needs_input_grad[2]:
            db = bias_grad(dy2).to(b_dtype)
        return dx, dw, db, None


def linear(x, weight, bias):
    """Drop-in for `nn.Linear(...)(x)` with a bias on CUDA tensors (the projections around the attention kernels and
    the MLP: reference backbone_vit.py:167,202,249; backbone_swin.py:309-311,433)."""
    _require_cuda(x)
    if bias is None or bias.shape[0] % 4 != 0 or x.dtype not in (torch.float32, torch.bfloat16):
        return torch.nn.functional.linear(x, weight, bias)      # nothing of ours to run: plain cuBLAS Linear
    if torch.is_autocast_enabled("cuda"):
        compute_dtype = torch.get_autocast_dtype("cuda")
        if compute_dtype != torch.bfloat16:
            return torch.nn.functional.linear(x, weight, bias)
    else:
        compute_dtype = torch.promote_types(x.dtype, weight.dtype)
        if compute_dtype not in (torch.float32, torch.bfloat16):
            return torch.nn.functional.linear(x, weight, bias)
    with torch.autocast("cuda", enabled=False):
        return _LinearBias.apply(x, weight, bias, compute_dtype)


# --------------------------------------------------------------------------------------------------
# patch embedding
# --------------------------------------------------------------------------------------------------
def _triple(vals):
    vals = [int(v) for v in vals]
    return [1] * (3 - len(vals)) + vals


class _PatchEmbed(torch.autograd.Function):
    """img (B,Cin,*sp) -> tokens (B, Np, N). weight: conv weight (N,Cin,*patch); pos: (1,Np,N) or None."""

    @staticmethod
    def forward(ctx, img, weight, bias, pos, grid, out_bf16):
        _require_cuda(img, weight, bias, pos)
        img_c = img.contiguous()
        if img_c.dtype not in (torch.float32, torch.bfloat16):
            img_c = img_c.float()
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous()
        p = pos.detach().float().contiguous() if pos is not None else None
        B, Cin = img_c.shape[:2]
        N = w.shape[0]
        img_dims, patch, grid3 = _triple(img_c.shape[2:]), _triple(w.shape[2:]), _triple(grid)
        Np = grid3[0] * grid3[1] * grid3[2]
        out = torch.empty((B, Np, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=img.device)
        lib = _lib.load()
        ws_bytes = lib.lcbi_patch_embed_workspace_bytes(B, Cin, _lib.int3(patch), _lib.int3(grid3), N)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=img_c.device) if ws_bytes else None   # tcgen05 path
        with _on(img_c.device):
            rc = lib.lcbi_patch_embed_fwd_ws(_p(img_c), int(img_c.dtype == torch.bfloat16), _p(w), _p(b),
                                             _p(p) if p is not None else None, _p(out), int(out_bf16), B, Cin,
                                             _lib.int3(img_dims), _lib.int3(patch), _lib.int3(grid3), N,
                                             _p(ws) if ws is not None else None, ws_bytes, _stream(img_c.device))
        _lib.check(rc, "lcbi_patch_embed_fwd_ws")
        ctx.save_for_backward(img_c, w)
        ctx.geom = (img_dims, patch, grid3, B, Cin, N)
        ctx.meta = (weight.shape, weight.dtype, bias.dtype, None if pos is None else (pos.shape, pos.dtype),
                    img.shape, img.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        img_c, w = ctx.saved_tensors
        img_dims, patch, grid3, B, Cin, N = ctx.geom
        w_shape, w_dtype, b_dtype, pos_meta, img_shape, img_dtype = ctx.meta
        dout = dout.contiguous()
        if dout.dtype not in (torch.float32, torch.bfloat16):
            dout = dout.float()
        dev = dout.device
        Np = grid3[0] * grid3[1] * grid3[2]
        dw = torch.empty((N, w.numel() // N), dtype=torch.float32, device=dev)
        db = torch.empty((N,), dtype=torch.float32, device=dev)
        dpos = torch.empty((Np, N), dtype=torch.float32, device=dev) if (pos_meta and ctx.needs_input_grad[3]) else None
        dimg = torch.empty(img_c.shape, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        lib = _lib.load()
        ws_bytes = lib.lcbi_patch_embed_workspace_bytes(B, Cin, _lib.int3(patch), _lib.int3(grid3), N)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None     # tcgen05 dW path
        with _on(dev):
            rc = lib.lcbi_patch_embed_bwd_ws(_p(img_c), int(img_c.dtype == torch.bfloat16), _p(w), _p(dout),
                                             int(dout.dtype == torch.bfloat16), _p(dw), _p(db),
                                             _p(dpos) if dpos is not None else None, _p(dimg) if dimg is not None else None,
                                             B, Cin, _lib.int3(img_dims), _lib.int3(patch), _lib.int3(grid3), N,
                                             _p(ws) if ws is not None else None, ws_bytes, _stream(dev))
        _lib.check(rc, "lcbi_patch_embed_bwd_ws")
        return (dimg.to(img_dtype).view(img_shape) if dimg is not None else None, dw.view(w_shape).to(w_dtype),
                db.to(b_dtype), dpos.view(pos_meta[0]).to(pos_meta[1]) if dpos is not None else None, None, None)


def patch_embed(img, weight, bias, pos, grid, out_dtype=torch.float32):
    """Fused strided-conv patch projection (+bias, +position embedding): (B,Cin,*sp) -> (B, prod(grid), N)."""
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("patch_embed output dtype must be float32 or bfloat16")
    return _PatchEmbed.apply(img, weight, bias, pos, tuple(int(g) for g in grid), out_dtype == torch.bfloat16)


# --------------------------------------------------------------------------------------------------
# Row gather / scatter (the exchange of the window-sharded Swin block, window_parallel.py)
# --------------------------------------------------------------------------------------------------
def _row_args(x2d, ids):
    dev = _require_cuda(x2d, ids)
    if x2d.dim() != 2 or not x2d.is_contiguous() or ids.dtype != torch.int64 or not ids.is_contiguous():
        raise ValueError("row copy: a contiguous (rows, F) tensor and contiguous int64 ids are required")
    row_bytes = x2d.shape[1] * x2d.element_size()
    if row_bytes % 16 != 0:
        raise ValueError(f"row copy: rows must be a multiple of 16 bytes, got {row_bytes}")
    return dev, row_bytes


def gather_rows(src, ids, out=None):
    """out[i] = src[ids[i]] (no autograd: the callers are autograd Functions). `out`: a contiguous (len(ids), F) tensor
    of src's dtype on the same device to write into."""
    dev, row_bytes = _row_args(src, ids)
    if out is None:
        out = torch.empty((ids.numel(), src.shape[1]), dtype=src.dtype, device=dev)
    elif (out.device != dev or out.dtype != src.dtype or tuple(out.shape) != (ids.numel(), src.shape[1])
          or not out.is_contiguous()):
        raise ValueError("gather_rows: out must be a contiguous (len(ids), F) tensor of src's dtype and device")
    with _on(dev):
        rc = _lib.load().lcbi_gather_rows(_p(src), _p(ids), _p(out), ids.numel(), row_bytes, _stream(dev))
    _lib.check(rc, "lcbi_gather_rows")
    return out


def scatter_rows(src, ids, n_out_rows):
    """out[ids[i]] = src[i] into a fresh (n_out_rows, F) tensor; rows no id points at are left uninitialised."""
    dev, row_bytes = _row_args(src, ids)
    if ids.numel() != src.shape[0]:
        raise ValueError("scatter_rows: one id per source row")
    out = torch.empty((n_out_rows, src.shape[1]), dtype=src.dtype, device=dev)
    with _on(dev):
        rc = _lib.load().lcbi_scatter_rows(_p(src), _p(ids), _p(out), ids.numel(), row_bytes, _stream(dev))
    _lib.check(rc, "lcbi_scatter_rows")
    return out


# --------------------------------------------------------------------------------------------------
# Swin window attention
# --------------------------------------------------------------------------------------------------
class _WindowAttention(torch.autograd.Function):
    """qkv (B, T, 3, H, d) bf16 on the un-padded token grid -> attention output (B, T, H*d) bf16.
    qkv_bias (3*H*d) supplies the q/k/v rows of pad tokens; table is relative_position_bias_table."""

    @staticmethod
    def forward(ctx, qkv, qkv_bias, table, grid, window, shift, scale, win_range=None):
        _require_cuda(qkv, qkv_bias, table)
        B, T, _, H, d = qkv.shape
        qkv = qkv.contiguous()
        bias_f = qkv_bias.detach().float().contiguous() if qkv_bias is not None else None
        table_f = table.detach().float().contiguous()
        # a window range leaves the rows of all other windows' tokens untouched: they must read as zero
        alloc = torch.empty if win_range is None else torch.zeros
        out = alloc((B, T, H * d), dtype=torch.bfloat16, device=qkv.device)
        lse2 = alloc((B, T, H), dtype=torch.float32, device=qkv.device)
        begin, count = (0, -1) if win_range is None else (int(win_range[0]), int(win_range[1]))
        lib = _lib.load()
        g, w, s = _lib.int_array(grid), _lib.int_array(window), _lib.int_array(shift)
        with _on(qkv.device):
            rc = lib.lcbi_win_attn_fwd_range(len(grid), g, w, s, B, H, d, float(scale), _p(qkv),
                                             _p(bias_f) if bias_f is not None else None, _p(table_f), _p(out), _p(lse2),
                                             begin, count, _stream(qkv.device))
        _lib.check(rc, "lcbi_win_attn_fwd_range")
        ctx.save_for_backward(qkv, bias_f, table_f, out, lse2)
        ctx.win_range = (begin, count)
        ctx.geom = (tuple(grid), tuple(window), tuple(shift), float(scale))
        ctx.meta = (None if qkv_bias is None else qkv_bias.dtype, table.dtype)
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, bias_f, table_f, out, lse2 = ctx.saved_tensors
        grid, window, shift, scale = ctx.geom
        B, T, _, H, d = qkv.shape
        d_out = d_out.to(torch.bfloat16).contiguous()
        dev = qkv.device
        begin, count = ctx.win_range
        dqkv = torch.empty_like(qkv) if count < 0 else torch.zeros_like(qkv)
        dsum = torch.empty((B, T, H), dtype=torch.float32, device=dev)
        dbias = torch.zeros((3 * H * d,), dtype=torch.float32, device=dev) if bias_f is not None else None
        dtable = torch.zeros_like(table_f)
        lib = _lib.load()
        g, w, s = _lib.int_array(grid), _lib.int_array(window), _lib.int_array(shift)
        with _on(dev):
            rc = lib.lcbi_win_attn_bwd_range(len(grid), g, w, s, B, H, d, scale, _p(qkv),
                                             _p(bias_f) if bias_f is not None else None, _p(table_f), _p(out), _p(lse2),
                                             _p(d_out), _p(dsum), _p(dqkv), _p(dbias) if dbias is not None else None,
                                             _p(dtable), begin, count, _stream(dev))
        _lib.check(rc, "lcbi_win_attn_bwd_range")
        bias_dtype, table_dtype = ctx.meta
        return (dqkv, dbias.to(bias_dtype) if dbias is not None else None, dtable.to(table_dtype), None, None, None, None,
                None)


def window_attention(qkv, qkv_bias, table, grid, window, shift, num_heads, scale=None, win_range=None):
    """Fused Swin (shifted-)window attention on the token grid.

    win_range = (begin, count) restricts the call to a range of the flattened (batch, window) list (window-sharded
    execution, `window_parallel.py`): rows of tokens outside those windows come back as zeros, and the parameter
    gradients are this range's contribution only.

    qkv: (B, *grid, 3*C) output of the qkv Linear applied to the un-padded, un-shifted tokens (feature index
    s*C + h*d + j, reference backbone_swin.py:339); qkv_bias: the Linear's bias (q/k/v of pad tokens) or None;
    table: relative_position_bias_table; window / shift: CONSTRUCTOR sizes (clamping happens inside).
    Returns (B, *grid, C) in qkv.dtype — the input of the proj Linear, already back at the source positions."""
    _require_cuda(qkv)
    grid = tuple(int(x) for x in grid)
    B, C3 = qkv.shape[0], qkv.shape[-1]
    C = C3 // 3
    d = C // num_heads
    T = 1
    for x in grid:
        T *= x
    if tuple(qkv.shape[1:-1]) != grid:
        raise ValueError("qkv must be (B, *grid, 3*C)")
    if scale is None:
        scale = d ** -0.5
    x = qkv if qkv.dtype == torch.bfloat16 else qkv.to(torch.bfloat16)
    out = _WindowAttention.apply(x.reshape(B, T, 3, num_heads, d), qkv_bias, table, grid,
                                 tuple(int(v) for v in window), tuple(int(v) for v in shift), float(scale), win_range)
    out = out.view(B, *grid, C)
    return out if out.dtype == qkv.dtype else out.to(qkv.dtype)


def set_window_kernel_mode(mode):
    """'auto' (default: tcgen05 kernels where measured at least as fast), 'tcgen05' or 'generic' for 3-D windows of
    128..512 tokens (C ABI lcbi_set_window_kernel_mode); used by the parity tests to cover both kernel families."""
    code = {"auto": 0, "tcgen05": 1, "generic": 2}[mode]
    _lib.check(_lib.load().lcbi_set_window_kernel_mode(code), "lcbi_set_window_kernel_mode")


def window_maps(grid, window, shift, device="cuda"):
    """(gather_map (nW,n) int32, region_ids (nW,n) int32, rel_pos_index (n,n) int32) computed by the CUDA code
    path's own closed forms — for bit-exactness tests against the reference's pad/roll/partition/compute_mask."""
    lib = _lib.load()
    g, w, s = _lib.int_array(grid), _lib.int_array(window), _lib.int_array(shift)
    n, nw = ctypes.c_int32(0), ctypes.c_int32(0)
    rc = lib.lcbi_window_maps(len(grid), g, w, s, None, None, None, ctypes.byref(n), ctypes.byref(nw), None)
    _lib.check(rc, "lcbi_window_maps")
    gather = torch.empty((nw.value, n.value), dtype=torch.int32, device=device)
    region = torch.empty((nw.value, n.value), dtype=torch.int32, device=device)
    relidx = torch.empty((n.value, n.value), dtype=torch.int32, device=device)
    with _on(gather.device):
        rc = lib.lcbi_window_maps(len(grid), g, w, s, _p(gather), _p(region), _p(relidx), ctypes.byref(n), ctypes.byref(nw),
                                  _stream(gather.device))
    _lib.check(rc, "lcbi_window_maps")
    return gather, region, relidx
