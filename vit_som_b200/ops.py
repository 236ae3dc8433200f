"""Tensor-level wrappers over the C-ABI (include/som_b200.h) and the autograd functions of the SOM layer.

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and the
autograd graph.  All arithmetic of the hot path happens in ``libsom_b200.so``; nothing in this file
computes distances, weights, losses or gradients with torch ops, and CPU tensors are rejected.

Launches per training step (reference call sequence models/vit_som.py:82-86 + backward):

    forward   som_forward      staging kernel (x and, when stale, W) -> tcgen05 GEMM + distance/argmin epilogue
                               -> BMU decode                                                       3 launches
    loss      som_loss_fused     loss + backward staging (R hi/lo, partial row/column sums), one pass  1 launch
    backward  som_backward_fused both gradient GEMMs (R^T x~ and R W~, gradient epilogue)              1 launch

Every wrapper runs on the device of its tensors (device guard + that device's current stream), whatever the
process-wide current device is.
"""
from __future__ import annotations

import contextlib
import ctypes
import os

import torch

from . import _lib
from ._lib import SomError, check, ptr, stream_ptr

MODE = {"euclidean": 0, "cosine": 1}
# operand precision, OR-ed into the mode word of every C-ABI call (include/som_b200.h: SOM_PREC_FP16X3)
PREC = {"tf32x3": 0, "fp16x3": 16}
PREC_FP16X3 = 16
# what a layer without an explicit ``precision`` attribute computes in (environment override for A/B runs)
DEFAULT_PRECISION = os.environ.get("SOM_B200_PRECISION", "fp16x3")


def is_f16(mode: int) -> bool:
    return bool(mode & PREC_FP16X3)

# bench.py sets this to a list to get per-launch CUDA-event timings of the three tensor-core GEMMs
# (entries: (name, start_event, end_event) recorded on the launching stream).
GEMM_TIMERS = None
# one persistent launch for both gradient GEMMs (som_backward_fused); False issues them separately
FUSE_BACKWARD = True


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


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _guard(device):
    """Make ``device`` current for the enclosed C-ABI calls (the library launches on the current device)."""
    if device.index is None or torch.cuda.current_device() == device.index:
        return contextlib.nullcontext()
    return torch.cuda.device(device)


def _grid_f32(grid_pos: torch.Tensor, device) -> torch.Tensor:
    """grid_positions as the contiguous fp32 tensor on ``device`` the kernels read through a raw pointer
    (module.half() / .double() casts the registered buffer)."""
    if grid_pos.dtype != torch.float32 or grid_pos.device != device or not grid_pos.is_contiguous():
        grid_pos = grid_pos.to(device=device, dtype=torch.float32).contiguous()
    return grid_pos


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise SomError(f"{name} is on {t.device}: the SOM hot path runs on a B200 only (no CPU path, no fallback)")
    if t.dtype != torch.float32:
        t = t.float()            # bf16/fp16 latents (autocast ViT) are widened exactly before the fp32-accurate math
    return t


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """A 2-D view with unit inner stride and non-overlapping rows (what the kernels index with an ld)."""
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


class Staging:
    """Staged form of a row-major [rows, dim] matrix, carved out of one flat allocation: hi | lo | aux.

    * 3xTF32 (default): hi / lo are the exact tf32 split in fp32 containers, row pitch ``ld`` = pad4(dim) floats; aux holds
      one float per row (|row|^2 for euclidean, 1/max(|row|, eps) for cosine).
    * 3xFP16 (``mode & PREC_FP16X3``): hi / lo are fp16 matrices of the row-scaled operand, row pitch ``ld`` = pad8(dim)
      halves; aux holds 3 * rows + 4 floats: aux | 2^-e | 2^e | 4 statistics words (include/som_b200.h)."""
    __slots__ = ("buf", "rows", "dim", "ld", "mode", "hi", "lo", "aux", "key", "from_optimizer", "aux_off")

    def __init__(self, rows: int, dim: int, mode: int, device):
        if rows <= 0 or dim <= 0:
            raise ValueError("empty operand")
        self.rows, self.dim, self.mode = rows, dim, mode
        if is_f16(mode):
            self.ld = _pad8(dim)
            n = rows * self.ld                # halves per matrix = n / 2 floats; hi + lo together n floats
            self.buf = torch.empty((n + 3 * rows + 4,), device=device, dtype=torch.float32)
            base = self.buf.data_ptr()
            self.hi, self.lo, self.aux = base, base + 2 * n, base + 4 * n
            self.aux_off = n
        else:
            self.ld = _pad4(dim)
            n = rows * self.ld
            self.buf = torch.empty((2 * n + rows,), device=device, dtype=torch.float32)
            base = self.buf.data_ptr()
            self.hi, self.lo, self.aux = base, base + 4 * n, base + 8 * n
            self.aux_off = 2 * n
        self.key = None                       # identity of the tensor version staged here (prototype cache)
        self.from_optimizer = False           # filled by the fused optimizer kernel for the parameter version in `key`

    def aux_tensor(self) -> torch.Tensor:
        """The per-row aux vector (norms / reciprocal norms)."""
        return self.buf[self.aux_off:self.aux_off + self.rows]

    def scale_tensor(self) -> torch.Tensor:
        """3xFP16 only: the 2^e row scales."""
        return self.buf[self.aux_off + 2 * self.rows:self.aux_off + 3 * self.rows]

    def dense(self) -> torch.Tensor:
        """The staged matrix put together again, [rows, ld] fp32 (hi + lo, row scales taken out): what the GEMMs see."""
        n = self.rows * self.ld
        if is_f16(self.mode):
            hi, lo = self.halves()
            return (hi.float() + lo.float()) / self.scale_tensor()[:, None]
        return (self.buf[:n] + self.buf[n:2 * n]).view(self.rows, self.ld)

    def halves(self):
        """3xFP16 only: (hi, lo) as [rows, ld] fp16 views."""
        n = self.rows * self.ld
        h = self.buf[:n].view(torch.float16)
        return h[:n].view(self.rows, self.ld), h[n:2 * n].view(self.rows, self.ld)


def stage_rows(src: torch.Tensor, mode: int) -> Staging:
    """Stand-alone staging of one matrix (som_prep_rows)."""
    src = _rowmajor(_require_cuda_f32(src, "operand"))
    if src.dim() != 2:
        raise ValueError("operand must be 2-D")
    st = Staging(src.shape[0], src.shape[1], mode, src.device)
    with _guard(src.device):
        check(_lib.lib().som_prep_rows(ptr(src), st.rows, st.dim, src.stride(0), mode, st.hi, st.lo, st.ld, st.aux,
                                       stream_ptr(src.device)), "som_prep_rows")
    return st


class ForwardState:
    """What one forward leaves behind for the loss and the backward: raw and staged operands and the distances."""
    __slots__ = ("x", "W", "xs", "ws", "mode", "B", "K", "D", "dist_buf", "ldd", "packed", "idx_offset",
                 "x_in", "W_in", "grad_accum", "dw_out", "dx_out", "nvls_dx", "layer", "dx_hook", "dx_exchanged")


def forward(x: torch.Tensor, W: torch.Tensor, mode: int, ws: Staging | None = None, stage_w: bool = True,
            want_dist: bool = True, want_bmu: bool = True, idx_offset: int = 0, k_total: int | None = None):
    """SOMLayer.forward through ``som_forward``.  Returns (state, bmu or None).

    ``ws`` is a prototype staging owned by the caller (the layer caches it across steps); it is (re)filled when
    ``stage_w`` is true."""
    xf = _rowmajor(_require_cuda_f32(x, "x"))
    Wf = _rowmajor(_require_cuda_f32(W, "prototypes"))
    if xf.dim() != 2 or Wf.dim() != 2 or xf.shape[1] != Wf.shape[1]:
        raise ValueError("latents [B, D] and prototypes [K, D] do not match")
    if xf.device != Wf.device:
        raise SomError(f"latents on {xf.device} but prototypes on {Wf.device}")
    B, D = xf.shape
    K = Wf.shape[0]
    if B == 0:
        raise ValueError("empty batch")
    dev = xf.device
    with _guard(dev):
        sp = stream_ptr(dev)
        xs = Staging(B, D, mode, dev)
        if ws is None:
            ws, stage_w = Staging(K, D, mode, dev), True
        elif ws.rows != K or ws.dim != D or ws.mode != mode:
            raise ValueError("prototype staging does not match the prototypes")
        ldd = _pad4(K)
        dist_buf = torch.empty((B, ldd), device=dev, dtype=torch.float32) if want_dist else None
        packed = torch.empty((B,), device=dev, dtype=torch.int64)
        bmu = torch.empty((B,), device=dev, dtype=torch.int64) if want_bmu else None
        L = _lib.lib()
        gws, gws_n = gemm_workspace(dev)
        if GEMM_TIMERS is None:
            check(L.som_forward(
                ptr(xf), xf.stride(0), ptr(Wf), Wf.stride(0), B, K, D, mode, 1 if stage_w else 0, idx_offset,
                xs.hi, xs.lo, xs.aux, ws.hi, ws.lo, ws.aux, xs.ld, ptr(dist_buf), ldd, ptr(packed), ptr(bmu),
                k_total if k_total is not None else K, gws, gws_n, sp), "som_forward")
        else:
            # instrumented run (bench.py roofline leg): the same launches issued one by one so that the CUDA events
            # bracket the tensor-core kernel alone
            check(L.som_prep_rows(ptr(xf), B, D, xf.stride(0), mode, xs.hi, xs.lo, xs.ld, xs.aux, sp), "som_prep_rows")
            if stage_w:
                check(L.som_prep_rows(ptr(Wf), K, D, Wf.stride(0), mode, ws.hi, ws.lo, ws.ld, ws.aux, sp),
                      "som_prep_rows")
            check(L.som_bmu_init(ptr(packed), B, sp), "som_bmu_init")
            check(_gemm("fwd", lambda: L.som_fwd_distances(xs.hi, xs.lo, xs.ld, xs.aux, ws.hi, ws.lo, ws.ld, ws.aux,
                                                           B, K, D, mode, idx_offset, ptr(dist_buf), ldd, ptr(packed),
                                                           gws, gws_n, sp)), "som_fwd_distances")
            if bmu is not None and is_f16(mode):
                check(L.som_bmu_decode_scaled(ptr(packed), B, k_total if k_total is not None else K, ptr(bmu), None,
                                              xs.aux, ws.aux, K, mode, sp), "som_bmu_decode_scaled")
            elif bmu is not None:
                check(L.som_bmu_decode(ptr(packed), B, k_total if k_total is not None else K, ptr(bmu), None, sp),
                      "som_bmu_decode")
    st = ForwardState()
    st.x, st.W, st.xs, st.ws, st.mode = xf, Wf, xs, ws, mode
    st.B, st.K, st.D, st.dist_buf, st.ldd, st.packed, st.idx_offset = B, K, D, dist_buf, ldd, packed, idx_offset
    st.grad_accum = None
    st.dw_out = None             # optional caller-owned [K, D] destination of dW (data parallel: a symmetric buffer)
    st.dx_out = None             # optional caller-owned [B, D] destination of dx (prototype shards: a symmetric buffer)
    st.nvls_dx = None
    st.layer = None
    st.dx_hook = None            # prototype shards: exchange of dx started by the dx-complete counter of the fused backward
    st.dx_exchanged = False
    return st, bmu


def bmu_decode(packed: torch.Tensor, k_total: int, want_min: bool = False, state=None):
    """packed (key, index) minima -> int64 BMUs.  ``state``: the forward these minima belong to; with 3xFP16 stagings the
    decode also leaves the scaling statistic of the backward staging in the latent staging (som_bmu_decode_scaled)."""
    B = packed.shape[0]
    dev = packed.device
    bmu = torch.empty((B,), device=dev, dtype=torch.int64)
    mn = torch.empty((B,), device=dev, dtype=torch.float32) if want_min else None
    with _guard(dev):
        if state is not None and is_f16(state.mode):
            check(_lib.lib().som_bmu_decode_scaled(ptr(packed), B, k_total, ptr(bmu), ptr(mn), state.xs.aux, state.ws.aux,
                                                   state.K, state.mode, stream_ptr(dev)), "som_bmu_decode_scaled")
        else:
            check(_lib.lib().som_bmu_decode(ptr(packed), B, k_total, ptr(bmu), ptr(mn), stream_ptr(dev)), "som_bmu_decode")
    return (bmu, mn) if want_min else bmu


def neighbourhood(bmu: torch.Tensor, grid_pos: torch.Tensor, T_dev: torch.Tensor, K: int, k_offset: int = 0):
    B = bmu.shape[0]
    dev = bmu.device
    ldw = _pad4(K)
    grid_pos = _grid_f32(grid_pos, dev)
    w = torch.empty((B, ldw), device=dev, dtype=torch.float32)
    with _guard(dev):
        check(_lib.lib().som_neighbourhood(ptr(bmu), ptr(grid_pos), B, K, k_offset, ptr(T_dev), ptr(w), ldw,
                                           stream_ptr(dev)), "som_neighbourhood")
    return w[:, :K]


_scratch = {}
_gemm_ws = {}


def gemm_workspace(device):
    """(pointer, floats) of the per-(device, stream) GEMM workspace that lets the CTA-pair kernel schedule stream-K."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _gemm_ws.get(key)
    if buf is None:
        # zero-filled once: the head of the workspace holds the stream-K hand-over flags (the kernels restore zeros)
        with _guard(device):
            n = int(_lib.lib().som_gemm_workspace_floats())
        buf = torch.zeros((n,), device=device, dtype=torch.float32)
        _gemm_ws[key] = buf
    return buf.data_ptr(), buf.numel()


def _loss_scratch(device, n: int) -> torch.Tensor:
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.zeros((max(n, 4096),), device=device, dtype=torch.float32)   # zero state is restored by the kernel
        _scratch[key] = buf
    return buf


def loss_parts(B: int, K: int, device) -> tuple[int, int]:
    """(n_row_parts, n_col_parts): sizes of the partial-sum tables ``som_loss_fused`` writes for this shape."""
    nr, nc = ctypes.c_int64(0), ctypes.c_int64(0)
    with _guard(device):
        check(_lib.lib().som_loss_fused_parts(B, K, ctypes.addressof(nr), ctypes.addressof(nc)), "som_loss_fused_parts")
    return int(nr.value), int(nc.value)


def _grad_scalar(g_out: torch.Tensor) -> torch.Tensor:
    g = g_out.reshape(1)
    if g.dtype != torch.float32:
        g = g.float()
    return g.contiguous()


class FusedLossFn(torch.autograd.Function):
    """loss = inv_count * sum_k w(bmu, T)[b,k] * dist(x, W)[b,k] as ONE autograd node over (x, W).

    Forward is one kernel over the distances the layer's forward already produced (``state``): it emits the loss
    and, when a gradient will be needed, the staged backward operand R and its partial row/column sums.  Backward is
    the two tcgen05 gradient GEMMs in one launch; the upstream gradient enters through their epilogue.  The B x K
    weight matrix and the B x K upstream-gradient matrix of the reference's autograd graph are never formed.
    (models/som_layer.py:137-152 + MeanBackward0 -> MulBackward0 -> EuclideanDistBackward0 | MmBackward0)

    ``dw_hook`` (data parallel, see distributed.DataParallelSOM): an object with ``gemm_sm_limit``, ``counter_ptr()``,
    ``exchange_counted(dw, expected)`` and ``exchange_after(dw)``; the last two return a join callable."""

    @staticmethod
    def forward(ctx, x, W, state, bmu, grid_pos, T_dev, inv_count, k_offset, want_grad, dw_hook, grid_dims=(0, 0)):
        B, K = state.B, state.K
        dev = state.dist_buf.device
        L = _lib.lib()
        grid_pos = _grid_f32(grid_pos, dev)
        if T_dev.device != dev or T_dev.dtype != torch.float32:
            T_dev = T_dev.to(device=dev, dtype=torch.float32)
        if dw_hook is not None and state.grad_accum is not None:
            raise SomError("grad_accumulator and a data-parallel wrapper are both set: the in-place accumulated "
                           "prototype gradient would never be exchanged (use DataParallelSOM.reduce_accumulator() "
                           "after the last chunk and detach the hook, or drop the accumulator)")
        loss = torch.empty((), device=dev, dtype=torch.float32)
        f16 = is_f16(state.mode)
        ldr = _pad8(K) if f16 else _pad4(K)
        with _guard(dev):
            scratch = _loss_scratch(dev, int(L.som_loss_fused_scratch_floats(B, K)))
            if want_grad:
                nrp, ncp = loss_parts(B, K, dev)
                # R hi | R lo | row parts (+ 4: 1 / S of the fp16 staging) | column parts (+ 4)
                r_floats = B * ldr // 2 if f16 else B * ldr          # floats per R matrix (fp16: two halves per float)
                row_floats = _pad4(B * nrp + 4)
                rbuf = torch.empty((2 * r_floats + row_floats + ncp * K + 4,), device=dev, dtype=torch.float32)
                base = rbuf.data_ptr()
                r_hi, r_lo = base, base + 4 * r_floats
                row_part = base + 8 * r_floats
                col_part = row_part + 4 * row_floats
            else:
                rbuf, r_hi, r_lo, row_part, col_part, nrp, ncp = None, None, None, None, None, 0, 0
            check(L.som_loss_fused(ptr(state.dist_buf), state.ldd, ptr(bmu), ptr(grid_pos), int(grid_dims[0]),
                                   int(grid_dims[1]), B, K, k_offset, ptr(T_dev),
                                   inv_count, state.mode, r_hi, r_lo, ldr, row_part, col_part, ptr(scratch), ptr(loss),
                                   state.xs.aux if f16 else None, state.ws.aux if f16 else None,
                                   stream_ptr(dev)), "som_loss_fused")
        ctx.state, ctx.rbuf, ctx.ptrs, ctx.ldr = state, rbuf, (r_hi, r_lo, row_part, nrp, col_part, ncp), ldr
        ctx.x_dtype, ctx.x_shape, ctx.dw_hook = x.dtype, x.shape, dw_hook
        ctx.save_for_backward(bmu, grid_pos, T_dev)     # keeps them alive; the kernels no longer need them
        return loss

    @staticmethod
    def backward(ctx, g_out):
        st = ctx.state
        if ctx.rbuf is None:
            raise SomError("som_loss was evaluated without gradient staging (no_grad) but backward was requested")
        r_hi, r_lo, row_part, nrp, col_part, ncp = ctx.ptrs
        B, K, D, mode = st.B, st.K, st.D, st.mode
        dev = st.dist_buf.device
        L = _lib.lib()
        g = _grad_scalar(g_out)
        if g.device != dev:
            g = g.to(dev)
        hook = ctx.dw_hook
        dxh = st.dx_hook                        # at most one of the two results is counted (dW: data parallel, dx: shards)
        if hook is not None and dxh is not None:
            raise SomError("a data-parallel dW hook and a prototype-shard dx hook cannot be combined on one layer")
        counted = hook if hook is not None else dxh
        sm_limit = int(getattr(counted, "gemm_sm_limit", 0) or 0) if counted is not None else 0
        dx = dw = join = None
        acc_buf = st.grad_accum                 # row-chunked batches: the GEMM epilogue adds into this [K, D] buffer
        nothing = (None,) * 9
        with _guard(dev):
            sp = stream_ptr(dev)
            gws, gws_n = gemm_workspace(dev)
            split = hook is not None and getattr(hook, "overlap", "counter") == "split"
            if ctx.needs_input_grad[0] and ctx.needs_input_grad[1] and FUSE_BACKWARD and not split:
                # both gradients: ONE persistent launch over the tiles of both GEMMs (dW tiles first)
                dw = acc_buf if acc_buf is not None else (
                    st.dw_out if st.dw_out is not None else torch.empty((K, D), device=dev, dtype=torch.float32))
                dx = st.dx_out if st.dx_out is not None else torch.empty((B, D), device=dev, dtype=torch.float32)
                counter = counted.counter_ptr() if counted is not None else None
                expected = ctypes.c_int64(-1)
                check(_gemm("dw+dx", lambda: L.som_backward_fused(
                    r_hi, r_lo, ctx.ldr, st.xs.hi, st.xs.lo, st.ws.hi, st.ws.lo, st.xs.ld, ptr(st.x), st.x.stride(0),
                    ptr(st.W), st.W.stride(0), row_part, nrp, col_part, ncp, st.xs.aux, st.ws.aux, ptr(g), B, K, D, mode,
                    ptr(dw), dw.stride(0), 1 if acc_buf is not None else 0, ptr(dx), dx.stride(0), sm_limit,
                    1 if dxh is not None else 0, counter, ctypes.addressof(expected), gws, gws_n, sp)),
                    "som_backward_fused")
                if hook is not None:
                    # data parallel: the exchange of dW starts as soon as its last tile is written (a stream-ordered
                    # wait on the counter the dW epilogues raise), under the dx tiles of the same launch
                    join = (hook.exchange_counted(dw, int(expected.value)) if expected.value >= 0
                            else hook.exchange_after(dw))
                elif dxh is not None:
                    # prototype shards: the same with the roles swapped - dx tiles first, their exchange under the dW tiles
                    dx = (dxh.exchange_counted(dx, int(expected.value)) if expected.value >= 0
                          else dxh.exchange_after(dx))
                    st.dx_exchanged = True
                if acc_buf is not None:
                    dw = None                   # already accumulated in place: nothing for autograd to add
            else:
                if ctx.needs_input_grad[1]:
                    dw = acc_buf if acc_buf is not None else (
                        st.dw_out if st.dw_out is not None else torch.empty((K, D), device=dev, dtype=torch.float32))
                    check(_gemm("dw", lambda: L.som_backward_dw(
                        r_hi, r_lo, ctx.ldr, st.xs.hi, st.xs.lo, st.xs.ld, ptr(st.W), st.W.stride(0), col_part, ncp,
                        st.ws.aux, ptr(g), B, K, D, mode, ptr(dw), dw.stride(0), 1 if acc_buf is not None else 0, 0,
                        gws, gws_n, sp)), "som_backward_dw")
                    if acc_buf is not None:
                        dw = None
                    elif hook is not None:
                        join = hook.exchange_after(dw)      # runs beside the dx GEMM below
                if ctx.needs_input_grad[0]:
                    dx = st.dx_out if st.dx_out is not None else torch.empty((B, D), device=dev, dtype=torch.float32)
                    check(_gemm("dx", lambda: L.som_backward_dx(
                        r_hi, r_lo, ctx.ldr, st.ws.hi, st.ws.lo, st.ws.ld, ptr(st.x), st.x.stride(0), row_part, nrp,
                        st.xs.aux, ptr(g), B, K, D, mode, ptr(dx), dx.stride(0), 0, sm_limit if join is not None else 0,
                        gws, gws_n, sp)), "som_backward_dx")
                    if dxh is not None:
                        dx = dxh.exchange_after(dx)
                        st.dx_exchanged = True
            if dx is not None:
                if dx.dtype != ctx.x_dtype:
                    dx = dx.to(ctx.x_dtype)
                dx = dx.view(ctx.x_shape)
            if join is not None:
                join()                          # compute stream waits for the communication stream
        return (dx, dw) + nothing


class WeightedLossFn(torch.autograd.Function):
    """loss = inv_count * sum(w(bmu, T) * dist) for a caller-supplied dense ``dist`` (general path: the distances
    did not come from this layer's forward, or were modified).  models/som_layer.py:137-152; w is never stored."""

    @staticmethod
    def forward(ctx, dist, bmu, grid_pos, T_dev, inv_count, k_offset):
        dist_c = _rowmajor(_require_cuda_f32(dist, "distances"))
        B, K = dist_c.shape
        dev = dist_c.device
        grid_pos = _grid_f32(grid_pos, dev)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        L = _lib.lib()
        with _guard(dev):
            scratch = _loss_scratch(dev, int(L.som_loss_scratch_floats(B, K)))
            check(L.som_weighted_loss(ptr(dist_c), dist_c.stride(0), ptr(bmu), ptr(grid_pos), B, K, k_offset,
                                      ptr(T_dev), inv_count, ptr(scratch), ptr(loss), stream_ptr(dev)),
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
        dev = bmu.device
        g = _grad_scalar(g_out)
        ldg = _pad4(K)
        G = torch.empty((B, ldg), device=dev, dtype=torch.float32)
        with _guard(dev):
            check(_lib.lib().som_weighted_loss_grad(ptr(bmu), ptr(grid_pos), B, K, ctx.k_offset, ptr(T_dev), ptr(g),
                                                    ctx.inv_count, ptr(G), ldg, stream_ptr(dev)),
                  "som_weighted_loss_grad")
        return G[:, :K], None, None, None, None, None


class DistanceFn(torch.autograd.Function):
    """distances = f(x, W) with the closed-form backward of ATen's cdist / normalize+mm for an arbitrary upstream
    gradient (models/som_layer.py:111-125 and their autograd).  In the training step of ViT-SOM no gradient reaches
    this node (``FusedLossFn`` differentiates the loss directly); it serves callers that use ``distances`` in
    their own expressions."""

    @staticmethod
    def forward(ctx, x, W, state):
        ctx.state = state
        ctx.x_dtype, ctx.x_shape = x.dtype, x.shape
        return state.dist_buf[:, :state.K]

    @staticmethod
    def backward(ctx, g_dist):
        st = ctx.state
        B, K, D, mode = st.B, st.K, st.D, st.mode
        dev = st.dist_buf.device
        L = _lib.lib()
        G = _rowmajor(_require_cuda_f32(g_dist, "grad_distances"))
        xs, ws = st.xs, st.ws
        if is_f16(mode):
            # general path (an arbitrary upstream gradient of the distances): its kernels take tf32 stagings
            mode &= 1
            xs, ws = stage_rows(st.x, mode), stage_rows(st.W, mode)
        ldr = _pad4(K)
        r_hi = torch.empty((B, ldr), device=dev, dtype=torch.float32)
        r_lo = torch.empty((B, ldr), device=dev, dtype=torch.float32)
        coef = torch.zeros((2 * B + 2 * K,), device=dev, dtype=torch.float32)
        ax, bx, aw, bw = coef[:B], coef[B:2 * B], coef[2 * B:2 * B + K], coef[2 * B + K:]
        dx = dw = None
        with _guard(dev):
            sp = stream_ptr(dev)
            check(L.som_bwd_coeffs(ptr(G), G.stride(0), ptr(st.dist_buf), st.ldd, B, K, mode, xs.aux, ws.aux,
                                   ptr(r_hi), ptr(r_lo), ldr, ptr(ax), ptr(bx), ptr(aw), ptr(bw), sp), "som_bwd_coeffs")
            gws, gws_n = gemm_workspace(dev)
            if ctx.needs_input_grad[0]:
                dx = torch.empty((B, D), device=dev, dtype=torch.float32)
                check(_gemm("dx", lambda: L.som_bwd_dx(ptr(r_hi), ptr(r_lo), ldr, ws.hi, ws.lo, ws.ld, ptr(st.x),
                                                       st.x.stride(0), ptr(ax), ptr(bx), B, K, D, ptr(dx), D,
                                                       gws, gws_n, sp)), "som_bwd_dx")
                if dx.dtype != ctx.x_dtype:
                    dx = dx.to(ctx.x_dtype)
                dx = dx.view(ctx.x_shape)
            if ctx.needs_input_grad[1]:
                dw = torch.empty((K, D), device=dev, dtype=torch.float32)
                check(_gemm("dw", lambda: L.som_bwd_dw(ptr(r_hi), ptr(r_lo), ldr, xs.hi, xs.lo, xs.ld, ptr(st.W),
                                                       st.W.stride(0), ptr(aw), ptr(bw), B, K, D, ptr(dw), D,
                                                       gws, gws_n, sp)), "som_bwd_dw")
        return dx, dw, None
