"""CPU, world_size 2 and 3 over gloo: the window-sharded Swin attention driver (range split, all-gather of the owned
output / dqkv rows, one all-reduce for the small parameter gradients) with the attention math supplied by the oracle instead of the CUDA
kernels. Every rank must end up with the single-process result, forward and all gradients."""
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


CASE = dict(B=1, grid=(6, 9, 5), window=(4, 4, 4), shift=(2, 2, 2), heads=2, d=4)   # padded, shifted, 3-D: 2*3*2 windows


def _inputs():
    torch.manual_seed(11)
    c = CASE
    C = c["heads"] * c["d"]
    qkv = torch.randn(c["B"], *c["grid"], 3 * C, dtype=torch.float64)
    bias = torch.randn(3 * C, dtype=torch.float64)
    n_tab = 1
    for w in c["window"]:
        n_tab *= 2 * w - 1
    table = torch.randn(n_tab, c["heads"], dtype=torch.float64) * 0.5
    d_out = torch.randn(c["B"], *c["grid"], C, dtype=torch.float64)
    return qkv, bias, table, d_out


def _worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from long_context_biomedical_imaging_b200 import window_parallel as wp

        c = CASE
        qkv, bias, table, d_out = _inputs()
        for t in (qkv, bias, table):
            t.requires_grad_(True)
        out = wp.window_attention_sharded(qkv, bias, table, c["grid"], c["window"], c["shift"], c["heads"],
                                          attn_fn=ao.window_attention_core)
        out.backward(d_out)
        # numpy arrays are pickled by value: a tensor would travel as a shared-memory handle that dies with this process
        queue.put((rank, out.detach().numpy(), qkv.grad.numpy(), bias.grad.numpy(), table.grad.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_window_sharded_attention_equals_single_process(world):
    c = CASE
    qkv, bias, table, d_out = _inputs()
    qkv.requires_grad_(True); bias.requires_grad_(True); table.requires_grad_(True)
    want = ao.window_attention_core(qkv, bias, table, c["grid"], c["window"], c["shift"], c["heads"])
    want.backward(d_out)

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, *arrays in results:
        out, dqkv, dbias, dtable = [torch.from_numpy(a) for a in arrays]
        torch.testing.assert_close(out, want.detach(), rtol=1e-10, atol=1e-10)
        torch.testing.assert_close(dqkv, qkv.grad, rtol=1e-9, atol=1e-10)
        torch.testing.assert_close(dbias, bias.grad, rtol=1e-9, atol=1e-10)
        torch.testing.assert_close(dtable, table.grad, rtol=1e-9, atol=1e-10)


def test_oracle_core_matches_the_golden_pinned_block_oracle():
    """window_attention_core (qkv in, attention out) is the middle of swin_part1, which is pinned against the
    reference by tests/golden/swin_part1.npz: with an identity projection the two must agree."""
    import torch.nn.functional as F

    torch.manual_seed(3)
    grid, window, shift, heads, d = (5, 9), (4, 4), (2, 2), 2, 4
    C = heads * d
    x = torch.randn(2, *grid, C, dtype=torch.float64)
    w_qkv, b_qkv = torch.randn(3 * C, C, dtype=torch.float64) * 0.3, torch.randn(3 * C, dtype=torch.float64)
    table = torch.randn((2 * 4 - 1) ** 2, heads, dtype=torch.float64)
    want = ao.swin_part1(x, window, shift, w_qkv, b_qkv, table, torch.eye(C, dtype=torch.float64),
                         torch.zeros(C, dtype=torch.float64), heads)
    got = ao.window_attention_core(F.linear(x, w_qkv, b_qkv), b_qkv, table, grid, window, shift, heads)
    torch.testing.assert_close(got, want, rtol=1e-10, atol=1e-10)
    # a partition of the window list sums to the whole
    n_units = 2 * 2 * 3
    parts = sum(ao.window_attention_core(F.linear(x, w_qkv, b_qkv), b_qkv, table, grid, window, shift, heads,
                                         win_range=(b, 5)) for b in (0, 5, 10) if b < n_units)
    torch.testing.assert_close(parts, want, rtol=1e-10, atol=1e-10)


def test_shard_range_and_window_count():
    from long_context_biomedical_imaging_b200 import window_parallel as wp

    assert wp.count_windows((64, 64, 64), (7, 7, 7)) == 1000          # cfg4 stage 1
    assert wp.count_windows((4, 4, 4), (7, 7, 7)) == 1                # clamped window
    assert wp.count_windows((128, 128), (7, 7)) == 361                # cfg2 stage 1
    for total, world in ((1000, 8), (27, 8), (8, 8), (5, 8)):
        covered = []
        for r in range(world):
            b, n = wp.shard_range(total, world, r)
            covered += list(range(b, b + n))
        assert covered == list(range(total))
