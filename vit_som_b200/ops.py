"""Tensor-level wrappers over the C-ABI (include/som_b200.h) and the two autograd functions.

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and the
autograd graph.  All arithmetic of the hot path happens in ``libsom_b200.so``; nothing in this file
computes distances, weights, losses or gradients with torch ops, and CPU tensors are rejected.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import SomError, check, ptr, stream_ptr

MODE = {"euclidean": 0, "cosine": 1}

# bench.py sets this to a list to get per-launch CUDA-event timings of the three tensor-core GEMMs
# (entries: (name, start_event, end_event) recorded on the launching stream).
GEMM_TIMERS = None


def _gemm(name, call):
    if GEMM_TIMERS is None:
        return call()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = call()
    e.record()
    GEMM_TIMERS.append((name, s, e))
    return rc


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise SomError(f"{name} is on {t.device}: the SOM hot path runs on a B200 only (no CPU path, no fallback)")
    if t.dtype != torch.float32:
        t = t.float()            # bf16/fp16 latents (autocast ViT) are widened exactly before the fp32-accurate math
    return t


class StagedOperand:
    """tf32 hi/lo split of a row-major matrix plus its per-row aux vector (|row|^2 or 1/max(|row|,eps))."""
    __slots__ = ("hi", "lo", "aux", "rows", "dim", "ld", "mode")

    def __init__(self, src: torch.Tensor, mode: int):
        src = _require_cuda_f32(src, "operand")
        if src.dim() != 2:
            raise ValueError("operand must be 2-D")
        if src.stride(1) != 1:
            src = src.contiguous()
        rows, dim = src.shape
        if rows == 0 or dim == 0:
            raise ValueError("empty operand")
        ld = _pad4(dim)
        self.hi = torch.empty((rows, ld), device=src.device, dtype=torch.float32)
        self.lo = torch.empty((rows, ld), device=src.device, dtype=torch.float32)
        self.aux = torch.empty((rows,), device=src.device, dtype=torch.float32)
        self.rows, self.dim, self.ld, self.mode = rows, dim, ld, mode
        L = _lib.lib()
        check(L.som_prep_rows(ptr(src), rows, dim, src.stride(0), mode, ptr(self.hi), ptr(self.lo), ld,
                              ptr(self.aux), stream_ptr()), "som_prep_rows")


def fwd_distances(xs: StagedOperand, ws: StagedOperand, want_dist: bool = True, idx_offset: int = 0,
                  packed: torch.Tensor | None = None):
    """Distances [B, K] (view of a [B, pad4(K)] buffer) and the packed (key, index) minima [B]."""
    if xs.dim != ws.dim or xs.mode != ws.mode:
        raise ValueError("latent and prototype staging do not match")
    B, K, D = xs.rows, ws.rows, xs.dim
    dev = xs.hi.device
    L = _lib.lib()
    if packed is None:
        packed = torch.empty((B,), device=dev, dtype=torch.int64)
        check(L.som_bmu_init(ptr(packed), B, stream_ptr()), "som_bmu_init")
    ldd = _pad4(K)
    dist_buf = torch.empty((B, ldd), device=dev, dtype=torch.float32) if want_dist else None
    check(_gemm("fwd", lambda: L.som_fwd_distances(
        ptr(xs.hi), ptr(xs.lo), xs.ld, ptr(xs.aux), ptr(ws.hi), ptr(ws.lo), ws.ld, ptr(ws.aux),
        B, K, D, xs.mode, idx_offset, ptr(dist_buf), ldd, ptr(packed), stream_ptr())), "som_fwd_distances")
    dist = dist_buf[:, :K] if want_dist else None
    return dist, packed


def bmu_decode(packed: torch.Tensor, k_total: int, want_min: bool = False):
    B = packed.shape[0]
    bmu = torch.empty((B,), device=packed.device, dtype=torch.int64)
    mn = torch.empty((B,), device=packed.device, dtype=torch.float32) if want_min else None
    check(_lib.lib().som_bmu_decode(ptr(packed), B, k_total, ptr(bmu), ptr(mn), stream_ptr()), "som_bmu_decode")
    return (bmu, mn) if want_min else bmu


def neighbourhood(bmu: torch.Tensor, grid_pos: torch.Tensor, T_dev: torch.Tensor, K: int, k_offset: int = 0):
    B = bmu.shape[0]
    ldw = _pad4(K)
    w = torch.empty((B, ldw), device=bmu.device, dtype=torch.float32)
    check(_lib.lib().som_neighbourhood(ptr(bmu), ptr(grid_pos), B, K, k_offset, ptr(T_dev), ptr(w), ldw, stream_ptr()),
          "som_neighbourhood")
    return w[:, :K]


_scratch = {}


def _loss_scratch(device, B: int, K: int) -> torch.Tensor:
    n = int(_lib.lib().som_loss_scratch_floats(B, K))
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.zeros((max(n, 4096),), device=device, dtype=torch.float32)   # zero state is restored by the kernel
        _scratch[key] = buf
    return buf


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """A 2-D view with unit inner stride and non-overlapping rows (what the kernels index with an ld)."""
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


class WeightedLossFn(torch.autograd.Function):
    """loss = inv_count * sum(w(bmu, T) * dist)  — models/som_layer.py:137-152 fused; w is never stored."""

    @staticmethod
    def forward(ctx, dist, bmu, grid_pos, T_dev, inv_count, k_offset):
        dist_c = _rowmajor(_require_cuda_f32(dist, "distances"))
        B, K = dist_c.shape
        loss = torch.empty((), device=dist_c.device, dtype=torch.float32)
        scratch = _loss_scratch(dist_c.device, B, K)
        check(_lib.lib().som_weighted_loss(ptr(dist_c), dist_c.stride(0), ptr(bmu), ptr(grid_pos), B, K, k_offset,
                                           ptr(T_dev), inv_count, ptr(scratch), ptr(loss), stream_ptr()),
              "som_weighted_loss")
        ctx.save_for_backward(bmu, grid_pos, T_dev)
        ctx.shape = (B, K)
        ctx.inv_count = inv_count
        ctx.k_offset = k_offset
        return loss

    @staticmethod
    def backward(ctx, g_out):
        bmu, grid_pos, T_dev = ctx.saved_tensors
        B, K = ctx.shape
        g = g_out.reshape(1).to(torch.float32).contiguous()
        ldg = _pad4(K)
        G = torch.empty((B, ldg), device=g.device, dtype=torch.float32)
        check(_lib.lib().som_weighted_loss_grad(ptr(bmu), ptr(grid_pos), B, K, ctx.k_offset, ptr(T_dev), ptr(g),
                                                ctx.inv_count, ptr(G), ldg, stream_ptr()), "som_weighted_loss_grad")
        return G[:, :K], None, None, None, None, None


class DistanceFn(torch.autograd.Function):
    """(distances, packed minima) = f(x, W) with the closed-form backward of ATen's cdist / normalize+mm
    (models/som_layer.py:111-125 and their autograd), both directions on the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, x, W, mode, w_staged, idx_offset):
        xs = StagedOperand(x, mode)
        ws = w_staged if w_staged is not None else StagedOperand(W, mode)
        dist, packed = fwd_distances(xs, ws, True, idx_offset)
        ctx.mode = mode
        ctx.xs, ctx.ws = xs, ws
        ctx.save_for_backward(x, W, dist)
        ctx.mark_non_differentiable(packed)
        return dist, packed

    @staticmethod
    def backward(ctx, g_dist, _g_packed):
        x, W, dist = ctx.saved_tensors
        xs, ws, mode = ctx.xs, ctx.ws, ctx.mode
        B, K, D = xs.rows, ws.rows, xs.dim
        dev = dist.device
        L = _lib.lib()
        G = _rowmajor(_require_cuda_f32(g_dist, "grad_distances"))
        ldr = _pad4(K)
        r_hi = torch.empty((B, ldr), device=dev, dtype=torch.float32)
        r_lo = torch.empty((B, ldr), device=dev, dtype=torch.float32)
        coef = torch.zeros((2 * B + 2 * K,), device=dev, dtype=torch.float32)
        ax, bx, aw, bw = coef[:B], coef[B:2 * B], coef[2 * B:2 * B + K], coef[2 * B + K:]
        check(L.som_bwd_coeffs(ptr(G), G.stride(0), ptr(dist), dist.stride(0), B, K, mode, ptr(xs.aux), ptr(ws.aux),
                               ptr(r_hi), ptr(r_lo), ldr, ptr(ax), ptr(bx), ptr(aw), ptr(bw), stream_ptr()),
              "som_bwd_coeffs")
        dx = dw = None
        if ctx.needs_input_grad[0]:
            xf = _rowmajor(_require_cuda_f32(x.reshape(B, D), "x"))
            dx = torch.empty((B, D), device=dev, dtype=torch.float32)
            check(_gemm("dx", lambda: L.som_bwd_dx(ptr(r_hi), ptr(r_lo), ldr, ptr(ws.hi), ptr(ws.lo), ws.ld, ptr(xf),
                                                   xf.stride(0), ptr(ax), ptr(bx), B, K, D, ptr(dx), D,
                                                   stream_ptr())), "som_bwd_dx")
            if dx.dtype != x.dtype:
                dx = dx.to(x.dtype)
        if ctx.needs_input_grad[1]:
            Wf = _rowmajor(_require_cuda_f32(W, "prototypes"))
            dw = torch.empty((K, D), device=dev, dtype=torch.float32)
            check(_gemm("dw", lambda: L.som_bwd_dw(ptr(r_hi), ptr(r_lo), ldr, ptr(xs.hi), ptr(xs.lo), xs.ld, ptr(Wf),
                                                   Wf.stride(0), ptr(aw), ptr(bw), B, K, D, ptr(dw), D,
                                                   stream_ptr())), "som_bwd_dw")
        return dx, dw, None, None, None
