"""CPU: the oracle restatement (oracle/) against the golden fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py). Bit-exact for the integer index maps; fp32 tolerance 1e-5 max-rel for floating point
(same fp32 arithmetic, different summation order)."""
import numpy as np
import pytest
import torch

from conftest import check_stored, load_golden, max_rel
from oracle import attention_oracle as ao
from oracle import window_maps as wm

FP32_TOL = 2e-5


def test_window_maps_bit_exact():
    g = load_golden("window_maps.npz")
    for case in g["cases"]:
        grid = tuple(int(x) for x in g[f"{case}/grid"])
        window = tuple(int(x) for x in g[f"{case}/window"])
        shift = tuple(int(x) for x in g[f"{case}/shift"])
        win, sh = wm.resolve_window(grid, window, shift)
        assert win == tuple(int(x) for x in g[f"{case}/win_used"])
        assert sh == tuple(int(x) for x in g[f"{case}/shift_used"])
        check_stored(g, f"{case}/gather_map", wm.gather_map(grid, window, shift))
        check_stored(g, f"{case}/mask", wm.shift_mask(grid, window, shift))


def test_rel_pos_index_bit_exact():
    g = load_golden("window_maps.npz")
    for key in [k for k in g.files if k.startswith("relidx/") and k.endswith("/shape")]:
        window = tuple(int(x) for x in key.split("/")[1].split("x"))
        check_stored(g, "relidx/" + key.split("/")[1], wm.rel_pos_index(window))


def test_region_ids_match_mask_semantics():
    # mask == 0 exactly where region ids agree, for a shifted + padded case
    grid, window, shift = (10, 9), (7, 7), (3, 3)
    ids = wm.region_ids(grid, window, shift)
    mask = wm.shift_mask(grid, window, shift)
    assert np.array_equal(mask == 0, ids[:, None, :] == ids[:, :, None])
    assert set(np.unique(mask)) <= {0.0, -100.0}


@pytest.mark.parametrize("case", ["a", "b"])
def test_sablock(case):
    g = load_golden("sablock.npz")
    t = {k: torch.from_numpy(g[f"{case}/{k}"]) for k in ["x", "w_qkv", "w_out", "b_out", "y", "dout", "dx", "dw_qkv",
                                                        "dw_out", "db_out"]}
    x, w_qkv, w_out, b_out = [t[k].clone().requires_grad_(True) for k in ["x", "w_qkv", "w_out", "b_out"]]
    y = ao.sablock(x, w_qkv, w_out, b_out, int(g[f"{case}/heads"]))
    assert max_rel(y.detach(), t["y"]) < FP32_TOL
    grads = torch.autograd.grad(y, [x, w_qkv, w_out, b_out], t["dout"])
    for got, key in zip(grads, ["dx", "dw_qkv", "dw_out", "db_out"]):
        assert max_rel(got, t[key]) < FP32_TOL, key


def _swin_case(g, name):
    meta = [int(v) for v in g[f"{name}/meta"]]
    B, C, H, K = meta[:4]
    grid = tuple(meta[4:4 + K])
    window = tuple(meta[4 + K:4 + 2 * K])
    shift = tuple(meta[4 + 2 * K:4 + 3 * K])
    return B, C, H, grid, window, shift


def test_swin_part1_all_cases():
    g = load_golden("swin_part1.npz")
    for name in g["cases"]:
        B, C, H, grid, window, shift = _swin_case(g, name)
        keys = ["x", "norm_w", "norm_b", "w_qkv", "b_qkv", "table", "w_proj", "b_proj"]
        leaves = [torch.from_numpy(g[f"{name}/{k}"]).clone().requires_grad_(True) for k in keys]
        x, nw, nb, wq, bq, tb, wp, bp = leaves
        y = ao.swin_part1(x, window, shift, wq, bq, tb, wp, bp, H, norm_w=nw, norm_b=nb)
        assert max_rel(y.detach(), g[f"{name}/y"]) < FP32_TOL, name
        grads = torch.autograd.grad(y, leaves, torch.from_numpy(g[f"{name}/dout"]))
        for got, key in zip(grads, ["dx", "dnorm_w", "dnorm_b", "dw_qkv", "db_qkv", "dtable", "dw_proj", "db_proj"]):
            assert max_rel(got, g[f"{name}/{key}"]) < 5e-5, (name, key)


def test_patch_embed():
    g = load_golden("patch_embed.npz")
    for name in ["vit2d", "vit3d", "vit2d_p2"]:
        x, w, b, pos = [torch.from_numpy(g[f"{name}/{k}"]).clone().requires_grad_(True) for k in ["x", "w", "b", "pos"]]
        y = ao.patch_embed_vit(x, w, b, pos)
        assert max_rel(y.detach(), g[f"{name}/y"]) < FP32_TOL
        grads = torch.autograd.grad(y, [x, w, b, pos], torch.from_numpy(g[f"{name}/dout"]))
        for got, key in zip(grads, ["dx", "dw", "db", "dpos"]):
            assert max_rel(got, g[f"{name}/{key}"]) < FP32_TOL, (name, key)
    for name in ["swin2d", "swin3d"]:
        x, w, b = [torch.from_numpy(g[f"{name}/{k}"]).clone().requires_grad_(True) for k in ["x", "w", "b"]]
        y = ao.patch_embed_swin(x, w, b)
        assert max_rel(y.detach(), g[f"{name}/y"]) < FP32_TOL
        grads = torch.autograd.grad(y, [x, w, b], torch.from_numpy(g[f"{name}/dout"]))
        for got, key in zip(grads, ["dx", "dw", "db"]):
            assert max_rel(got, g[f"{name}/{key}"]) < FP32_TOL, (name, key)


def test_dense_attention_rows_matches_full():
    torch.manual_seed(0)
    q, k, v = torch.randn(3, 300, 64).unbind(0)
    full = ao.dense_attention(q[None, None], k[None, None], v[None, None], 0.125)[0, 0]
    rows, lse = ao.dense_attention_rows(q[[0, 17, 299]], k, v, 0.125, chunk=64)
    assert max_rel(rows.float(), full[[0, 17, 299]]) < 1e-5
    ref_lse = torch.logsumexp((q[[0, 17, 299]] @ k.T) * 0.125, -1)
    assert max_rel(lse.float(), ref_lse) < 1e-5


def test_layer_norm_restatement_matches_the_reference_module_class():
    """The blocks' norm1 / norm2 are torch.nn.LayerNorm (reference backbone_vit.py:249-258): the oracle's written-out
    LayerNorm must agree with that class, outputs and all three gradients, in fp64."""
    import torch

    from oracle import attention_oracle as ao

    torch.manual_seed(3)
    for rows, C in [(5, 48), (7, 192), (3, 768), (2, 1536)]:
        ln = torch.nn.LayerNorm(C).double()
        with torch.no_grad():
            ln.weight.normal_(1.0, 0.3)
            ln.bias.normal_(0.0, 0.3)
        x = (torch.randn(rows, C, dtype=torch.float64) * 2 + 0.5).requires_grad_(True)
        g = torch.randn(rows, C, dtype=torch.float64)
        ref = ln(x)
        ref_grads = torch.autograd.grad(ref, (x, ln.weight, ln.bias), g)
        w, b = ln.weight.detach().clone().requires_grad_(True), ln.bias.detach().clone().requires_grad_(True)
        x2 = x.detach().clone().requires_grad_(True)
        out = ao.layer_norm_rows(x2, w, b, ln.eps)
        grads = torch.autograd.grad(out, (x2, w, b), g)
        assert max_rel(out.detach().numpy(), ref.detach().numpy()) < 1e-12
        for a, r in zip(grads, ref_grads):
            assert max_rel(a.numpy(), r.numpy()) < 1e-11


def _partition_by_copies(grid, window, shift):
    """The reference's data movement spelled out with array copies (backbone_swin.py:441-468, :135-165): clamp the
    window, zero-pad the far ends, roll by -shift, cut raster-ordered windows. Applied to an arange token grid it gives
    the gather map; applied to the region-label image of compute_mask (:604-624) it gives the region ids."""
    win, sh = wm.resolve_window(grid, window, shift)
    K = len(grid)
    pads = [(0, (-g) % w) for g, w in zip(grid, win)]

    def cut(img):
        shape = []
        for n_k, w in zip(img.shape, win):
            shape += [n_k // w, w]
        t = img.reshape(shape)
        t = t.transpose([2 * k for k in range(K)] + [2 * k + 1 for k in range(K)])
        return t.reshape(-1, int(np.prod(win)))

    tokens = np.arange(int(np.prod(grid)), dtype=np.int64).reshape(grid)
    padded = np.pad(tokens, pads, constant_values=-1)
    rolled = np.roll(padded, [-s for s in sh], axis=tuple(range(K))) if any(sh) else padded
    gather = cut(rolled)

    labels = np.zeros(padded.shape, dtype=np.int64)
    cnt = 0
    import itertools
    slabs = [(slice(-w), slice(-w, -s), slice(-s, None)) for w, s in zip(win, sh)]
    for combo in itertools.product(*slabs):
        labels[combo] = cnt
        cnt += 1
    return gather, cut(labels)


def test_closed_form_maps_equal_pad_roll_partition_on_random_geometries():
    """Beyond the 22 golden geometries: the closed forms agree with the copy-based restatement for random 2-D / 3-D
    grids, windows and shifts (clamped axes, zero shifts on some axes, non-divisible grids)."""
    from hypothesis import given, settings, strategies as st

    @st.composite
    def geometry(draw):
        k = draw(st.sampled_from([2, 3]))
        grid = tuple(draw(st.integers(1, 11 if k == 3 else 20)) for _ in range(k))
        window = tuple(draw(st.integers(1, 7)) for _ in range(k))
        shift = tuple(draw(st.integers(0, w - 1)) if draw(st.booleans()) else 0 for w in window)
        return grid, window, shift

    @settings(max_examples=150, deadline=None)
    @given(geometry())
    def check(geo):
        grid, window, shift = geo
        gather, labels = _partition_by_copies(grid, window, shift)
        assert np.array_equal(wm.gather_map(grid, window, shift), gather)
        ids = wm.region_ids(grid, window, shift)
        same_ref = labels[:, :, None] == labels[:, None, :]
        same_ours = ids[:, :, None] == ids[:, None, :]
        assert np.array_equal(same_ref, same_ours)       # the mask only depends on which slots share a region

    check()
