"""CPU model of the 3xFP16 operand format: what the staging kernels write and what the epilogues
undo, restated in numpy, so that the design's error bounds and the scale algebra are checked without a GPU.

This is a model of OUR numerics (row scaling, fp16 hi/lo split, the batch-wide scale of the backward operand), not of
the reference; the GPU parity tests compare the kernels themselves with the reference oracle under both precisions."""
import re

import numpy as np
import pytest

from conftest import ROOT

TOP = 14          # staged rows have their largest magnitude in [2^14, 2^15)  (F16_TOP_EXP in som_b200.cu)


def row_scale(v):
    """2^e per row: the power of two that lifts the row maximum into [2^TOP, 2^(TOP+1)); rows of zeros keep 1."""
    amax = np.abs(v).max(axis=1)
    e = np.where(amax > 0, TOP - np.floor(np.log2(np.where(amax > 0, amax, 1.0))), 0.0)
    return np.exp2(e).astype(np.float32)


def stage(v):
    """(hi, lo, scale): fp16 split of the row-scaled matrix, as prep_rows_kernel<GROUP, true> writes it."""
    s = row_scale(v)
    vs = (v * s[:, None]).astype(np.float32)            # exact: power of two
    hi = vs.astype(np.float16)
    lo = (vs - hi.astype(np.float32)).astype(np.float16)
    return hi, lo, s


def dot3(a_hi, a_lo, b_hi, b_lo):
    """hi.hi + hi.lo + lo.hi with exact products (fp16 x fp16 fits fp32) and a wide accumulator."""
    ah, al, bh, bl = (t.astype(np.float64) for t in (a_hi, a_lo, b_hi, b_lo))
    return ah @ bh.T + ah @ bl.T + al @ bh.T


def test_python_and_header_agree_on_the_precision_flag():
    from vit_som_b200 import ops
    hdr = open(f"{ROOT}/include/som_b200.h").read()
    assert int(re.search(r"#define\s+SOM_PREC_FP16X3\s+(\d+)", hdr).group(1)) == ops.PREC_FP16X3 == ops.PREC["fp16x3"]
    assert ops.PREC["tf32x3"] == 0 and ops.DEFAULT_PRECISION in ops.PREC
    assert ops.is_f16(ops.MODE["cosine"] | ops.PREC_FP16X3) and not ops.is_f16(ops.MODE["cosine"])
    # the mode word keeps the distance in bit 0
    assert (ops.MODE["cosine"] | ops.PREC_FP16X3) & 1 == 1 and (ops.MODE["euclidean"] | ops.PREC_FP16X3) & 1 == 0


def test_staging_layout_of_both_precisions():
    """Offsets of hi | lo | aux inside one allocation (ops.Staging): fp16 matrices are half the bytes, aux holds
    3 * rows + 4 floats, everything stays 16-byte aligned."""
    import torch
    from vit_som_b200 import ops
    for rows, dim in [(7, 50), (16, 3136), (5, 12)]:
        t = ops.Staging(rows, dim, ops.MODE["euclidean"], torch.device("cpu"))
        f = ops.Staging(rows, dim, ops.MODE["euclidean"] | ops.PREC_FP16X3, torch.device("cpu"))
        assert t.ld == (dim + 3) // 4 * 4 and f.ld == (dim + 7) // 8 * 8
        assert t.lo - t.hi == 4 * rows * t.ld and t.aux - t.hi == 8 * rows * t.ld and t.buf.numel() == 2 * rows * t.ld + rows
        assert f.lo - f.hi == 2 * rows * f.ld and f.aux - f.hi == 4 * rows * f.ld
        assert f.buf.numel() == rows * f.ld + 3 * rows + 4
        assert all(p % 16 == 0 for p in (f.hi, f.lo, f.aux, t.hi, t.lo, t.aux))
        hi, lo = f.halves()
        assert hi.shape == lo.shape == (rows, f.ld) and hi.dtype == torch.float16
        assert f.aux_tensor().numel() == rows and f.scale_tensor().numel() == rows
        assert f.aux_tensor().data_ptr() == f.aux and f.scale_tensor().data_ptr() == f.aux + 8 * rows


def test_split_carries_22_bits_where_lo_is_normal_and_an_absolute_floor_elsewhere():
    rng = np.random.default_rng(0)
    v = (rng.standard_normal((64, 3136)) * np.exp2(rng.integers(-20, 20, size=(64, 1)))).astype(np.float32)
    v[:, ::7] *= 1e-7                                   # elements far below the row maximum
    hi, lo, s = stage(v)
    assert np.all(np.abs(hi.astype(np.float32)).max(axis=1) >= 2.0 ** TOP) and np.all(np.isfinite(hi.astype(np.float32)))
    rec = (hi.astype(np.float64) + lo.astype(np.float64)) / s[:, None]
    err = np.abs(rec - v)
    big = np.abs(v) * s[:, None] >= 2.0 ** -3            # lo normal: element within 2^-17 of the (scaled) row maximum
    assert np.all(err[big] <= np.abs(v[big]) * 2.0 ** -22)
    floor = (2.0 ** -24) / s[:, None]                    # subnormal spacing of lo, un-scaled: 2^-39 of the row maximum
    assert np.all(err <= np.maximum(np.abs(v) * 2.0 ** -22, floor * np.ones_like(v)))
    assert np.all(floor[:, 0] <= np.abs(v).max(axis=1) * 2.0 ** -38)


@pytest.mark.parametrize("positive", [False, True])
def test_three_product_split_is_fp32_accurate(positive):
    """x.W^T through the staged operands, un-scaled as the distance epilogue does it, against fp64."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((96, 3136)).astype(np.float32)
    W = rng.random((160, 3136)).astype(np.float32)
    if positive:
        x = np.abs(x) + 0.5
    xh, xl, xs = stage(x)
    wh, wl, ws = stage(W)
    acc = dot3(xh, xl, wh, wl)
    got = acc / xs[:, None].astype(np.float64) / ws[None, :].astype(np.float64)
    ref = x.astype(np.float64) @ W.astype(np.float64).T
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6
    # rows of wildly different magnitude do not disturb each other (the scales are per row)
    x2 = x * np.exp2(rng.integers(-30, 30, size=(96, 1))).astype(np.float32)
    xh, xl, xs = stage(x2)
    got2 = dot3(xh, xl, wh, wl) / xs[:, None].astype(np.float64) / ws[None, :].astype(np.float64)
    ref2 = x2.astype(np.float64) @ W.astype(np.float64).T
    assert np.all(np.abs(got2 - ref2).max(axis=1) <= np.abs(ref2).max(axis=1) * 2e-6)


def r_scale(stat, inv_count):
    """f16_r_scale of som_b200.cu: the power of two that puts inv_count * stat into [2^(TOP-1), 2^TOP)."""
    t = np.float32(stat) * np.float32(inv_count)
    if not (t > 0) or not np.isfinite(t):
        return np.float32(1.0)
    _, p = np.frexp(t)
    return np.float32(np.exp2(np.clip(TOP - p, -120, 120)))


def test_backward_operand_one_scale_serves_both_gradient_gemms():
    """R^[b,k] = R[b,k] * 2^-e_b * 2^-g_k * S: stays inside the fp16 range, and both gradient contractions come back
    exactly scaled by what their epilogue folds into beta (2^e_b / S for dx rows, 2^g_k / S for dW rows)."""
    rng = np.random.default_rng(2)
    B, K, D = 128, 96, 320
    x = (rng.standard_normal((B, D)) * np.exp2(rng.integers(-3, 4, size=(B, 1)))).astype(np.float32)
    W = rng.random((K, D)).astype(np.float32)
    d = np.sqrt(np.maximum((x.astype(np.float64) ** 2).sum(1)[:, None] + (W.astype(np.float64) ** 2).sum(1)[None, :]
                           - 2 * x.astype(np.float64) @ W.astype(np.float64).T, 0)).astype(np.float32)
    bmu = d.argmin(1)
    w = np.exp(-((bmu[:, None] - np.arange(K)[None, :]) ** 2) / 50.0).astype(np.float32)    # any weights in (0, 1]
    inv_count = np.float32(1.0 / (B * K))
    R = (inv_count * w / d).astype(np.float32)
    xh, xl, xs = stage(x)
    wh, wl, ws = stage(W)
    stat = np.max((1.0 / xs) / d[np.arange(B), bmu]) * np.max(1.0 / ws)       # bmu_decode_stat_kernel
    S = r_scale(stat, inv_count)
    Rs = (R * (1.0 / xs)[:, None]) * ((1.0 / ws) * S)[None, :]
    assert Rs.max() < 2.0 ** TOP * 1.01 and Rs.max() >= 2.0 ** (TOP - 12)    # the bound is tight to the spread of 1/d
    rh = Rs.astype(np.float16)
    rl = (Rs - rh.astype(np.float32)).astype(np.float16)
    assert np.all(np.isfinite(rh.astype(np.float32)))
    # dx = R . W  (reduce over k): A = R^ rows b, B = W^ read MN-major;  un-scale rows by 2^e_b / S
    acc_dx = dot3(rh, rl, wh.T, wl.T)
    dx = acc_dx * (xs.astype(np.float64) / S)[:, None]
    ref_dx = R.astype(np.float64) @ W.astype(np.float64)
    assert np.abs(dx - ref_dx).max() / np.abs(ref_dx).max() < 2e-6
    # dW = R^T . x (reduce over b): un-scale rows k by 2^g_k / S
    acc_dw = dot3(rh.T, rl.T, xh.T, xl.T)
    dw = acc_dw * (ws.astype(np.float64) / S)[:, None]
    ref_dw = R.astype(np.float64).T @ x.astype(np.float64)
    assert np.abs(dw - ref_dw).max() / np.abs(ref_dw).max() < 2e-6
