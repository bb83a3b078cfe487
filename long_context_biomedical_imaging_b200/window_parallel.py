"""Window-sharded Swin attention over torch.distributed (SURVEY.md 8e, row "Swin windows", option (a)).

The windows of one W-MSA / SW-MSA block are independent, so the flattened (batch, window) list is split evenly over
the ranks of a process group: every rank holds the full (replicated) qkv of the block and runs the fused kernel on its
window range only (`ops.window_attention(..., win_range=...)`, C ABI `lcbi_win_attn_{fwd,bwd}_range`). Every token
belongs to exactly one window, so the ranks' results are DISJOINT token rows: they are exchanged with ONE all-gather of
the owned rows (packed window-major by `lcbi_gather_rows`, 1/P of the tensor per rank) followed by a scatter to token
order (`lcbi_scatter_rows`) — not an all-reduce of the zero-padded full tensor, which moves twice the bytes. Backward
mirrors it inside the same autograd node: the incoming gradient is replicated, every rank back-propagates through its
windows, and one all-gather carries the dqkv rows (disjoint again) together with the rank's partial gradients of the
two small parameters (qkv.bias through the pad tokens, the relative-position table), which are then summed locally.

Consecutive blocks use different (shifted) partitions, so tokens must be visible to every rank between blocks: this is
the simple formulation of 8e (a). Per-GPU compute at cfg4 stage 1 is of the order of 200 microseconds at 8 GPUs, so
collective latency and not the kernels bounds this configuration - bench.py reports it as measured.
"""
from __future__ import annotations

import functools

import torch
import torch.distributed as dist
import torch.nn.functional as F


def _used_window(grid, window, shift):
    """Reference get_window_size clamp (backbone_swin.py:200-224)."""
    win = [int(g) if g <= w else int(w) for g, w in zip(grid, window)]
    sh = [0 if g <= w else int(s) for g, w, s in zip(grid, window, shift)]
    return win, sh


def count_windows(grid, window):
    """Windows per image for a token grid and a CONSTRUCTOR window (clamp, then ceil division of the padded grid)."""
    total = 1
    for g, w in zip(grid, window):
        w_used = g if g <= w else w
        total *= -(-int(g) // int(w_used))
    return total


def shard_range(total, world, rank):
    """Contiguous, near-even split of `total` units; trailing ranks may get an empty range."""
    per = -(-total // world)
    begin = min(rank * per, total)
    return begin, min(per, total - begin)


@functools.lru_cache(maxsize=64)
def _window_token_ids(grid, window, shift):
    """(nW, n) int64 on the CPU: token index (raster order of the un-padded grid) feeding each window slot, -1 for the
    zero-pad tokens — the reference's pad -> roll(-shift) -> window_partition chain (backbone_swin.py:441-468) applied
    to an index tensor."""
    win, sh = _used_window(grid, window, shift)
    t = 1
    for g in grid:
        t *= g
    ids = torch.arange(t, dtype=torch.int64).reshape(grid)
    pads = []
    for g, w in zip(reversed(grid), reversed(win)):
        pads += [0, (w - g % w) % w]
    ids = F.pad(ids, pads, value=-1)
    if any(s > 0 for s in sh):
        ids = torch.roll(ids, shifts=[-s for s in sh], dims=list(range(len(grid))))
    k = len(grid)
    nwin = [p // w for p, w in zip(ids.shape, win)]
    shape = []
    for a, b in zip(nwin, win):
        shape += [a, b]
    ids = ids.reshape(shape).permute(*range(0, 2 * k, 2), *range(1, 2 * k, 2))
    n = 1
    for w in win:
        n *= w
    return ids.reshape(-1, n).contiguous()


_ROW_ID_CACHE = {}


def _row_ids(batch, grid, window, shift, world, device):
    """Cached on the device per geometry (built once: the map is static for a block):
      gather_ids  (world, per*n) int64  flattened (batch*T) token row of every slot of every rank's window range, padded
                  to `per` windows per rank; pad tokens and the dummy windows that even out the last ranks read row 0
      scatter_ids (world*per*n) int64   the same, with those slots pointing at a dump row (index batch*T)."""
    key = (batch, tuple(grid), tuple(window), tuple(shift), world, str(device))
    hit = _ROW_ID_CACHE.get(key)
    if hit is None:
        ids, n_rows = _build_row_ids(batch, grid, window, shift, world)
        flat = ids.reshape(-1)
        hit = (ids.clamp(min=0).to(device), torch.where(flat >= 0, flat, torch.full_like(flat, n_rows)).to(device), n_rows)
        if len(_ROW_ID_CACHE) > 64:
            _ROW_ID_CACHE.clear()
        _ROW_ID_CACHE[key] = hit
    return hit


def _build_row_ids(batch, grid, window, shift, world):
    ids = _window_token_ids(tuple(grid), tuple(window), tuple(shift))            # (nW, n)
    nW, n = ids.shape
    t = 1
    for g in grid:
        t *= g
    full = torch.cat([torch.where(ids >= 0, ids + b * t, ids) for b in range(batch)], 0)      # (B*nW, n)
    per = -(-(batch * nW) // world)
    padded = torch.full((world * per, n), -1, dtype=torch.int64)
    padded[:batch * nW] = full
    return padded.reshape(world, per * n), batch * t


def _all_gather_rows(rows, group, world):
    out = rows.new_empty((world * rows.shape[0],) + tuple(rows.shape[1:]))
    try:
        dist.all_gather_into_tensor(out, rows, group=group)
    except (RuntimeError, NotImplementedError):          # a backend without the flat variant
        dist.all_gather(list(out.chunk(world, 0)), rows, group=group)
    return out


_EXT_IDS_CACHE = {}


def _scatter_ids_with_tail(scatter_ids, world, tail_rows, n_rows):
    """scatter ids of the gathered buffer when every rank's chunk carries `tail_rows` extra rows behind its window rows:
    those rows are sent to the dump row (index n_rows) like the pad / dummy slots."""
    if tail_rows == 0:
        return scatter_ids
    key = (scatter_ids.data_ptr(), world, tail_rows)
    hit = _EXT_IDS_CACHE.get(key)
    if hit is None:
        per = scatter_ids.numel() // world
        ext = torch.full((world, per + tail_rows), n_rows, dtype=scatter_ids.dtype, device=scatter_ids.device)
        ext[:, :per] = scatter_ids.view(world, per)
        hit = (ext.reshape(-1), scatter_ids)          # keeps scatter_ids alive so that its data_ptr stays unique
        if len(_EXT_IDS_CACHE) > 64:
            _EXT_IDS_CACHE.clear()
        _EXT_IDS_CACHE[key] = hit
    return hit[0]


def _exchange_disjoint_rows(x2d, ids, n_rows, rank, group, world, extra=None):
    """x2d: (n_rows, F) whose rows owned by this rank are valid. Returns the (n_rows, F) tensor in which every rank's
    owned rows are filled in: pack own rows -> ONE all-gather -> scatter by token id. ids = (gather_ids, scatter_ids).
    `extra` (a flat float tensor, e.g. this rank's partial parameter gradients) travels in the same all-gather as whole
    extra rows behind the rank's window rows; the sum over the ranks is returned as the second value."""
    gather_ids, scatter_ids = ids
    feat, row_bytes = x2d.shape[1], x2d.shape[1] * x2d.element_size()
    my_ids = gather_ids[rank]
    n = my_ids.numel()
    extra_bytes = 0 if extra is None else extra.numel() * extra.element_size()
    # the tail starts at a 16-byte aligned row boundary: round the rows up when a row is not a multiple of 16 bytes
    head_rows = n
    while (head_rows * row_bytes) % 16:
        head_rows += 1
    tail_rows = -(-extra_bytes // row_bytes)
    send = torch.empty((head_rows + tail_rows, feat), dtype=x2d.dtype, device=x2d.device)
    fast = x2d.is_cuda and row_bytes % 16 == 0
    if fast:
        # streaming row copies (csrc/row_copy.cu): torch's index_select / index_copy_ ran at ~0.2 TB/s on these rows and
        # were a quarter of the sharded step at cfg4 stage 1
        from . import ops
        ops.gather_rows(x2d.contiguous(), my_ids, out=send[:n])
    else:                                     # CPU tensors (the gloo tests of the host-side plumbing)
        torch.index_select(x2d, 0, my_ids, out=send[:n])
    if extra is not None:
        send.view(-1).view(torch.uint8)[head_rows * row_bytes:head_rows * row_bytes + extra_bytes].view(extra.dtype).copy_(extra)
    everyone = _all_gather_rows(send, group, world)                      # (world * (head_rows + tail_rows), F)
    ext_ids = scatter_ids
    if head_rows + tail_rows != n:
        if head_rows != n:                    # (odd row sizes: CPU tests only) drop the alignment rows first
            everyone = everyone.view(world, head_rows + tail_rows, feat)
            everyone = torch.cat([everyone[:, :n], everyone[:, head_rows:]], 1).reshape(-1, feat)
        ext_ids = _scatter_ids_with_tail(scatter_ids, world, tail_rows, n_rows)
    if fast:
        out = ops.scatter_rows(everyone, ext_ids, n_rows + 1)[:n_rows]   # last row swallows pad / dummy / tail rows
    else:
        out = x2d.new_empty((n_rows + 1, feat))
        out.index_copy_(0, ext_ids, everyone)
        out = out[:n_rows]
    if extra is None:
        return out, None
    tails = everyone.view(world, n + tail_rows, feat)[:, n:].reshape(world, -1).view(torch.uint8)[:, :extra_bytes]
    return out, tails.contiguous().view(extra.dtype).view(world, -1).sum(0)


class _ShardedWindowAttention(torch.autograd.Function):
    """The whole sharded block as ONE autograd node: forward = range attention on this rank's windows + one all-gather of
    the owned output rows; backward = range backward + one all-gather that carries the owned dqkv rows AND the rank's
    partial gradients of the two small parameters (qkv.bias through the pad tokens, the relative-position table), which
    every rank then sums locally - no separate all-reduce (265 us of a 1.17 ms step at 8 GPUs)."""

    @staticmethod
    def forward(ctx, qkv, qkv_bias, table, attn_fn, call, ids, n_rows, rank, group, world):
        inputs = [qkv, qkv_bias, table]
        local = [None if t is None else t.detach().requires_grad_(t.requires_grad) for t in inputs]
        with torch.enable_grad():
            part = attn_fn(local[0], local[1], local[2], *call[0], **call[1])
        full, _ = _exchange_disjoint_rows(part.detach().reshape(n_rows, part.shape[-1]), ids, n_rows, rank, group, world)
        ctx.local, ctx.part = local, part
        ctx.meta = (ids, n_rows, rank, group, world)
        return full.reshape(part.shape)

    @staticmethod
    def backward(ctx, g):
        ids, n_rows, rank, group, world = ctx.meta
        local, part = ctx.local, ctx.part
        ctx.local = ctx.part = None
        wanted = [t for t in local if t is not None and t.requires_grad]
        grads = iter(torch.autograd.grad(part, wanted, g.reshape(part.shape), allow_unused=True)) if wanted else iter(())
        got = [next(grads) if (t is not None and t.requires_grad) else None for t in local]
        dqkv, dbias, dtable = got
        small = [x for x in (dbias, dtable) if x is not None]
        wide = torch.float64 if any(x.dtype == torch.float64 for x in small) else torch.float32
        extra = torch.cat([x.reshape(-1).to(wide) for x in small]) if small else None
        out = [None, None, None]
        if dqkv is not None:
            full, total = _exchange_disjoint_rows(dqkv.reshape(n_rows, dqkv.shape[-1]), ids, n_rows, rank, group, world, extra)
            out[0] = full.reshape(dqkv.shape)
        elif extra is not None:                               # parameters only (qkv detached): plain all-reduce
            total = extra.clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        else:
            total = None
        offset = 0
        for i, x in ((1, dbias), (2, dtable)):
            if x is not None:
                out[i] = total[offset:offset + x.numel()].view(x.shape).to(x.dtype)
                offset += x.numel()
        return out[0], out[1], out[2], None, None, None, None, None, None, None


def window_attention_sharded(qkv, qkv_bias, table, grid, window, shift, num_heads, scale=None, group=None, attn_fn=None):
    """Drop-in for `ops.window_attention` when the block's windows are split over the ranks of `group`.
    Every rank must pass the same (replicated) qkv / parameters and receives the same full output.
    `attn_fn` (default: the CUDA op) must accept `win_range=(begin, count)`; tests inject the CPU oracle."""
    if attn_fn is None:
        from . import ops
        attn_fn = ops.window_attention
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return attn_fn(qkv, qkv_bias, table, grid, window, shift, num_heads, scale)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    grid = tuple(int(g) for g in grid)
    batch = qkv.shape[0]
    total = batch * count_windows(grid, window)
    begin, count = shard_range(total, world, rank)
    g_ids, s_ids, n_rows = _row_ids(batch, grid, tuple(int(w) for w in window), tuple(int(s) for s in shift), world,
                                    qkv.device)
    call = ((grid, window, shift, num_heads, scale), {"win_range": (begin, count)})
    return _ShardedWindowAttention.apply(qkv, qkv_bias, table, attn_fn, call, (g_ids, s_ids), n_rows, rank, group, world)
