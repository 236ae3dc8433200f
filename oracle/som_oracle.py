"""CPU oracle of the ViT-SOM SOM-layer hot path  —  TEST INFRASTRUCTURE ONLY.

This module is the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it; the product package ``vit_som_b200`` never does (there is no CPU fallback).

It restates, in numpy, what ``/root/reference/models/som_layer.py`` computes through PyTorch ATen
(pinned by the reference to torch 2.2, Dockerfile:1; the arithmetic lives in ATen, which is not under
/root/reference).  Every function cites the reference lines it follows.  Parity pinning: the reference
ships no golden vectors or runnable tests for this path, so the oracle is pinned
against outputs of the *unmodified reference module* run in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, checked by ``tests/test_oracle_golden.py``).

Two precisions are offered: ``np.float32`` mirrors the reference's arithmetic formula by formula
(including ATen's augmented-GEMM evaluation of cdist), ``np.float64`` is the exact-ish restatement
used to classify best-matching-unit disagreements as fp32 near-ties.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

EUCLIDEAN, COSINE = "euclidean", "cosine"
NORMALIZE_EPS = 1e-12          # torch.nn.functional.normalize default eps


# ------------------------------------------------------------------------------------------------
# grid, temperature
# ------------------------------------------------------------------------------------------------
def grid_positions(map_size, topology="square") -> np.ndarray:
    """[K,2] fp32 grid coordinates, k = row*cols + col  (models/som_layer.py:60-81)."""
    rows, cols = int(map_size[0]), int(map_size[1])
    if topology == "square":
        gy, gx = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")        # :61-66
        return np.stack([gy, gx], axis=-1).reshape(-1, 2).astype(np.float32)         # :67
    if topology == "hexa":
        pos = np.zeros((rows * cols, 2), dtype=np.float32)                           # :70
        for i in range(rows * cols):
            r, c = divmod(i, cols)                                                   # :72-73
            pos[i, 0] = c                                                            # :74
            pos[i, 1] = np.float32(r * np.sqrt(3) / 2)                               # :75
            if r % 2 == 1:
                pos[i, 0] += 0.5                                                     # :76-77
        return pos
    raise ValueError(f"Unsupported topology: {topology}")                            # :79


def total_iterations(n_samples: int, batch_size: int, total_epochs: int) -> float:
    """(len(dataset) / batch_size) * total_epochs — a float, world size ignored (models/som_layer.py:131)."""
    return (n_samples / batch_size) * total_epochs


def temperature(iteration, Tmax: float, Tmin: float, total_iters: float):
    """T = Tmax * (Tmin/Tmax) ** (it / (total - 1))  (models/som_layer.py:132).

    With a python/np integer `iteration` the result is a python float (double arithmetic); with an
    fp32 array it follows torch's tensor semantics (fp32 pow, fp32 multiply)."""
    if isinstance(iteration, np.ndarray):
        e = (iteration.astype(np.float32) / np.float32(total_iters - 1)).astype(np.float32)
        return (np.float32(Tmax) * np.power(np.float32(Tmin / Tmax), e, dtype=np.float32)).astype(np.float32)
    return Tmax * (Tmin / Tmax) ** (iteration / (total_iters - 1))


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def l2_normalize(a: np.ndarray, dtype=np.float32):
    """F.normalize(a, p=2, dim=1): a / max(||a||, eps)  (models/som_layer.py:120-121)."""
    a = a.astype(dtype, copy=False)
    nrm = np.sqrt((a * a).sum(axis=1, dtype=dtype)).astype(dtype)
    denom = np.maximum(nrm, dtype(NORMALIZE_EPS))
    return (a / denom[:, None]).astype(dtype), denom


def distances(x: np.ndarray, W: np.ndarray, fcn: str = EUCLIDEAN, dtype=np.float32) -> np.ndarray:
    """compute_distances (models/som_layer.py:111-125).

    euclidean: torch.cdist(x, W, p=2) (:118).  For B > 25 or K > 25 ATen's `_euclidean_dist` evaluates
      sqrt(clamp_min([-2x, |x|^2, 1] . [W, 1, |W|^2]^T, 0)); for smaller problems a direct per-pair kernel
      sqrt(sum (x-w)^2) is used.  Both are restated here.
    cosine: 1 - normalize(x) . normalize(W)^T (:120-122)."""
    x = np.ascontiguousarray(x, dtype=dtype)
    W = np.ascontiguousarray(W, dtype=dtype)
    if fcn == EUCLIDEAN:
        B, K = x.shape[0], W.shape[0]
        if dtype == np.float64:
            d2 = (x * x).sum(1)[:, None] - 2.0 * (x @ W.T) + (W * W).sum(1)[None, :]
            # the exact difference form is better conditioned when it is affordable
            if B * K * x.shape[1] <= 2 ** 28:
                d2 = ((x[:, None, :] - W[None, :, :]) ** 2).sum(-1)
            return np.sqrt(np.maximum(d2, 0.0))
        if B <= 25 and K <= 25:
            diff = x[:, None, :] - W[None, :, :]
            return np.sqrt((diff * diff).sum(-1, dtype=dtype)).astype(dtype)
        xn = (x * x).sum(axis=1, keepdims=True, dtype=dtype)
        wn = (W * W).sum(axis=1, keepdims=True, dtype=dtype)
        x_aug = np.concatenate([dtype(-2) * x, xn, np.ones_like(xn)], axis=1)
        w_aug = np.concatenate([W, np.ones_like(wn), wn], axis=1)
        return np.sqrt(np.maximum(x_aug @ w_aug.T, dtype(0))).astype(dtype)
    if fcn == COSINE:
        xh, _ = l2_normalize(x, dtype)
        wh, _ = l2_normalize(W, dtype)
        return (dtype(1) - xh @ wh.T).astype(dtype)
    raise ValueError(f"Unsupported distance function: {fcn}")                        # :124


def bmu(d: np.ndarray) -> np.ndarray:
    """torch.argmin(distances, dim=1): first minimal index per row, int64 (models/som_layer.py:88)."""
    return np.argmin(d, axis=1).astype(np.int64)


def index_to_position(indices: np.ndarray, map_size) -> np.ndarray:
    """(idx // cols, idx % cols) as float (models/som_layer.py:134-135)."""
    cols = int(map_size[1])
    return np.stack([indices // cols, indices % cols], axis=1).astype(np.float32)


def weights(bmu_idx: np.ndarray, pos: np.ndarray, T, dtype=np.float32) -> np.ndarray:
    """compute_weights: exp(-||p_k - p_bmu||^2 / (2 T^2)), norm first then squared (models/som_layer.py:148-150)."""
    pos = pos.astype(dtype)
    pb = pos[bmu_idx]                                                                 # :148
    diff = pos[None, :, :] - pb[:, None, :]
    g = np.sqrt((diff * diff).sum(-1, dtype=dtype)).astype(dtype)                     # :149 torch.norm(dim=2)
    if isinstance(T, (np.floating, np.ndarray)) and np.asarray(T).dtype == np.float32:
        T32 = np.float32(T)                       # T is a 0-dim fp32 tensor in ViTSOM (vit_som.py:65,84): fp32 pow/mul
        two_t2 = dtype(np.float32(2) * (T32 * T32))
    else:
        two_t2 = dtype(2 * float(T) ** 2)         # python-float T: `2 * T ** 2` in double, cast once by the division
    return np.exp(-(g * g) / two_t2).astype(dtype)                                    # :150


def som_loss(w: np.ndarray, d: np.ndarray, dtype=np.float32):
    """mean(w * d) over all B*K entries (models/som_layer.py:141-142)."""
    return dtype((w.astype(np.float64) * d.astype(np.float64)).mean()) if dtype == np.float64 \
        else np.float32((w * d).mean(dtype=np.float64))


# ------------------------------------------------------------------------------------------------
# backward (closed forms of the autograd graph MeanBackward0 -> MulBackward0 -> distance backward)
# ------------------------------------------------------------------------------------------------
def backward(x, W, d, w, fcn=EUCLIDEAN, g_out=1.0, dtype=np.float32):
    """Gradients of  g_out * mean(w * d)  w.r.t. x and W; w is a constant (no grad flows through
    compute_weights: integer indices and a buffer).

    euclidean (ATen _euclidean_dist_backward):  G = g_out*w/(B K); R = G/d, R[d==0] = 0
        dx = x * rowsum(R) - R W        dW = W * colsum(R) - R^T x
    cosine (MmBackward0 + normalize backward):  Gs = -G; dxh = Gs Wh; dWh = Gs^T xh
        dx = (dxh - (dxh . xh) xh) / max(|x|, eps)   (same for W)."""
    x = x.astype(dtype); W = W.astype(dtype); d = d.astype(dtype); w = w.astype(dtype)
    B, K = d.shape
    G = (w * dtype(g_out / (B * K))).astype(dtype)
    if fcn == EUCLIDEAN:
        with np.errstate(divide="ignore", invalid="ignore"):
            R = np.where(d == 0, dtype(0), G / d).astype(dtype)
        dx = x * R.sum(1, keepdims=True) - R @ W
        dW = W * R.sum(0)[:, None] - R.T @ x
        return dx.astype(dtype), dW.astype(dtype)
    if fcn == COSINE:
        xh, xden = l2_normalize(x, dtype)
        wh, wden = l2_normalize(W, dtype)
        dxh = -(G @ wh)
        dwh = -(G.T @ xh)
        dx = (dxh - (dxh * xh).sum(1, keepdims=True) * xh) / xden[:, None]
        dW = (dwh - (dwh * wh).sum(1, keepdims=True) * wh) / wden[:, None]
        return dx.astype(dtype), dW.astype(dtype)
    raise ValueError(f"Unsupported distance function: {fcn}")


# ------------------------------------------------------------------------------------------------
# one full step, and the BMU near-tie classifier
# ------------------------------------------------------------------------------------------------
@dataclass
class StepResult:
    distances: np.ndarray
    bmu: np.ndarray
    weights: np.ndarray
    loss: np.floating
    grad_x: np.ndarray
    grad_w: np.ndarray


def step(x, W, pos, T, fcn=EUCLIDEAN, g_out=1.0, dtype=np.float32, bmu_override=None) -> StepResult:
    """forward -> compute_weights -> som_loss -> backward, the sequence of models/vit_som.py:82-86."""
    if x.ndim > 2:
        x = x.reshape(x.shape[0], -1)                                                 # models/som_layer.py:84-85
    d = distances(x, W, fcn, dtype)
    b = bmu(d) if bmu_override is None else np.asarray(bmu_override, dtype=np.int64)
    w = weights(b, pos, T, dtype)
    loss = som_loss(w, d, dtype)
    gx, gw = backward(x, W, d, w, fcn, g_out, dtype)
    return StepResult(d, b, w, loss, gx, gw)


def classify_bmu_mismatches(x, W, got, fcn=EUCLIDEAN, rel_tol=4e-6):
    """Compare BMU indices `got` with the fp64 oracle.  A mismatch is an *fp32 near-tie* when the fp64
    distances of the two candidates differ by less than rel_tol relative (fp32 cannot order them
    reliably: the squared distance is a difference of O(|x|^2+|w|^2) terms rounded at 6e-8 relative).
    Returns (n_mismatch, n_not_near_tie, worst_rel_gap)."""
    d64 = distances(x, W, fcn, np.float64)
    want = bmu(d64)
    got = np.asarray(got, dtype=np.int64)
    bad = np.nonzero(want != got)[0]
    worst, hard = 0.0, 0
    for r in bad:
        a, b_ = d64[r, want[r]], d64[r, got[r]]
        scale = max(abs(a), abs(b_), 1e-30) if fcn == EUCLIDEAN else 1.0
        gap = abs(b_ - a) / scale
        worst = max(worst, gap)
        if gap > rel_tol:
            hard += 1
    return len(bad), hard, worst


def rel_err(a, b) -> float:
    """Norm-wise relative error ||a-b|| / ||b|| (the 1e-5 gate of BASELINE.json's north_star)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
