"""Window-sharded Swin attention over torch.distributed (SURVEY.md 8e, row "Swin windows", option (a)).

The windows of one W-MSA / SW-MSA block are independent, so the flattened (batch, window) list is split evenly over
the ranks of a process group: every rank holds the full (replicated) qkv of the block and runs the fused kernel on its
window range only (`ops.window_attention(..., win_range=...)`, C ABI `lcbi_win_attn_{fwd,bwd}_range`). Every token
belongs to exactly one window, so the ranks' results are DISJOINT token rows: they are exchanged with ONE all-gather of
the owned rows (window-major, 1/P of the tensor per rank) followed by a scatter to token order — not an all-reduce of
the zero-padded full tensor, which moves twice the bytes. Backward mirrors it: the incoming gradient is replicated,
every rank back-propagates through its windows, the dqkv rows (disjoint again) are all-gathered, and the two small
parameter gradients (qkv.bias through the pad tokens, the relative-position table) share one all-reduce.

Consecutive blocks use different (shifted) partitions, so tokens must be visible to every rank between blocks: this is
the simple formulation of 8e (a). Per-GPU compute at cfg4 stage 1 is of the order of 100 microseconds, so collective
latency and not the kernels bounds this configuration - bench.py reports it as measured.
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


def _exchange_disjoint_rows(x2d, ids, n_rows, rank, group, world):
    """x2d: (n_rows, F) whose rows owned by this rank are valid. Returns the (n_rows, F) tensor in which every rank's
    owned rows are filled in: gather own rows -> all-gather -> scatter by token id. ids = (gather_ids, scatter_ids)."""
    gather_ids, scatter_ids = ids
    if x2d.is_cuda and (x2d.shape[1] * x2d.element_size()) % 16 == 0:
        # streaming row copies (csrc/row_copy.cu): torch's index_select / index_copy_ ran at ~0.2 TB/s on these rows and
        # were a quarter of the sharded step at cfg4 stage 1
        from . import ops
        mine = ops.gather_rows(x2d.contiguous(), gather_ids[rank])
        everyone = _all_gather_rows(mine, group, world)
        return ops.scatter_rows(everyone, scatter_ids, n_rows + 1)[:n_rows]   # last row swallows pad / dummy slots
    mine = x2d.index_select(0, gather_ids[rank])              # CPU tensors (the gloo tests of the host-side plumbing)
    everyone = _all_gather_rows(mine, group, world)
    out = x2d.new_empty((n_rows + 1, x2d.shape[1]))
    out.index_copy_(0, scatter_ids, everyone)
    return out[:n_rows]


class _ReplicatedRows(torch.autograd.Function):
    """Identity on the replicated qkv; the gradient rows each rank produced (those of its windows) are all-gathered."""

    @staticmethod
    def forward(ctx, x, ids, n_rows, rank, group, world):
        ctx.meta = (ids, n_rows, rank, group, world)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ids, n_rows, rank, group, world = ctx.meta
        full = _exchange_disjoint_rows(g.reshape(n_rows, g.shape[-1]), ids, n_rows, rank, group, world)
        return full.reshape(g.shape), None, None, None, None, None


class _ReplicatedInput(torch.autograd.Function):
    """Identity on a small tensor every rank holds a copy of; its gradient is the sum of the ranks' partial gradients."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


class _GatherOutputs(torch.autograd.Function):
    """Partial output (own windows' rows valid) -> full output on every rank. The incoming gradient is replicated and
    the range backward reads only its own windows' rows of it, so it is passed through unchanged."""

    @staticmethod
    def forward(ctx, part, ids, n_rows, rank, group, world):
        full = _exchange_disjoint_rows(part.reshape(n_rows, part.shape[-1]), ids, n_rows, rank, group, world)
        return full.reshape(part.shape)

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None, None, None


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
    ids = (g_ids, s_ids)
    qkv_r = _ReplicatedRows.apply(qkv, ids, n_rows, rank, group, world)
    if qkv_bias is not None:                 # one all-reduce for both small parameter gradients
        packed = _ReplicatedInput.apply(torch.cat([qkv_bias.reshape(-1), table.reshape(-1).to(qkv_bias.dtype)]), group)
        bias_r = packed[:qkv_bias.numel()].reshape(qkv_bias.shape)
        table_r = packed[qkv_bias.numel():].reshape(table.shape).to(table.dtype)
    else:
        bias_r, table_r = None, _ReplicatedInput.apply(table, group)
    part = attn_fn(qkv_r, bias_r, table_r, grid, window, shift, num_heads, scale, win_range=(begin, count))
    return _GatherOutputs.apply(part, ids, n_rows, rank, group, world)
