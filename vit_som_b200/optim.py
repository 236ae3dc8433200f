"""Fused prototype optimizer.

The reference optimises ``som_layer.parameters()`` with ``torch.optim.AdamW`` in a parameter group without explicit
weight decay, i.e. the default 0.01 (``/root/reference/models/vit_som.py:140-151``).  On the prototypes that step is a
pure HBM pass over four ``[K, D]`` arrays - and the next forward starts with another pass over ``W`` (the tf32
staging).  ``FusedPrototypeAdamW`` does both in ONE kernel (``som_adamw_step``): it reads ``W, dW, m, v`` and writes
``W, m, v`` and the staged operands ``W_hi, W_lo, |w|^2`` (or ``1/max(|w|, eps)``) of the NEW prototypes, so that the
next ``SOMLayer.forward`` stages the latents only.  Same update rule, defaults and state-dict layout as
``torch.optim.AdamW`` (decoupled weight decay, bias-corrected moments, no amsgrad); learning-rate schedulers work
(they edit ``param_groups[0]['lr']``).  Step count and learning rate live in device memory, so the step can be
captured in a CUDA graph together with the forward / backward.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import SomError, check, ptr


class FusedPrototypeAdamW(torch.optim.Optimizer):
    """``opt = FusedPrototypeAdamW(layer, lr=...)``; ``loss.backward(); opt.step(); opt.zero_grad()``.

    ``layer`` is a :class:`vit_som_b200.SOMLayer` (or a prototype shard); only ``layer.prototypes`` is optimised -
    hand the other parameters of the model to their own optimizer.  The gradient is taken from
    ``layer.prototypes.grad`` or, for row-chunked batches, from ``layer.grad_accumulator``.  ``grad_scale`` multiplies
    the gradient inside the kernel (e.g. ``1 / world`` when the exchange left a sum instead of a mean)."""

    def __init__(self, layer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 grad_scale: float = 1.0, stage: bool = True):
        W = layer.prototypes
        if not W.is_cuda:
            raise SomError("FusedPrototypeAdamW runs on a B200 only: move the layer to cuda first (no CPU path)")
        if W.dtype != torch.float32 or not W.is_contiguous():
            raise SomError("prototypes must be a contiguous fp32 parameter")
        if lr < 0 or eps < 0 or weight_decay < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__([W], dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.layer, self.stage = layer, stage
        self._hp = torch.tensor([lr, 0.0, grad_scale], device=W.device, dtype=torch.float32)   # lr, step count, grad scale
        self._lr_on_device = float(lr)
        self._staging = None
        self.state[W] = {"step": self._hp[1], "exp_avg": torch.zeros_like(W), "exp_avg_sq": torch.zeros_like(W)}

    def _moments(self):
        st = self.state[self.layer.prototypes]
        return st["exp_avg"], st["exp_avg_sq"]

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        W = self.layer.prototypes
        st = self.state[W]
        step = st.get("step", 0.0)
        self._hp[1] = float(step.item() if torch.is_tensor(step) else step)
        st["step"] = self._hp[1]                          # the kernel reads the count from the hyper-parameter block
        for k in ("exp_avg", "exp_avg_sq"):
            # private copies: torch's load_state_dict keeps tensors that already have the right dtype and device, so the
            # moments would alias the buffers of the optimizer the state came from
            st[k] = st[k].detach().to(device=W.device, dtype=torch.float32).clone(memory_format=torch.contiguous_format)
        self._lr_on_device = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        layer = self.layer
        W = layer.prototypes
        grad = W.grad if W.grad is not None else layer.grad_accumulator
        if grad is None:
            return loss
        if grad.dtype != torch.float32 or grad.device != W.device or grad.stride(1) != 1:
            raise SomError("prototype gradient must be fp32, row-major, on the prototypes' device")
        group = self.param_groups[0]
        lr = float(group["lr"])
        if lr != self._lr_on_device:                      # a scheduler moved it: refresh the device copy (no sync)
            self._hp[0] = lr
            self._lr_on_device = lr
        self._hp[1] += 1.0
        m, v = self._moments()
        K, D = W.shape
        dev = W.device
        mode = layer._mode()
        st = None
        if self.stage:
            # one persistent staging buffer, rewritten in place by every step: a captured graph (forward reading it,
            # optimizer writing it) stays consistent across replays.  A backward that still needs the staging of the
            # previous parameter version must run before the step - which is the order of any training loop.
            st = self._staging
            if st is None or st.rows != K or st.dim != D or st.mode != mode or st.buf.device != dev:
                st = self._staging = ops.Staging(K, D, mode, dev)
        with ops._guard(dev):
            check(_lib.lib().som_adamw_step(
                ptr(W), W.stride(0), ptr(grad), grad.stride(0), ptr(m), ptr(v), m.stride(0), K, D, ptr(self._hp),
                float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                mode, st.hi if st else None, st.lo if st else None, st.ld if st else 0, st.aux if st else None,
                _lib.stream_ptr(dev)), "som_adamw_step")
        torch.autograd.graph.increment_version(W)         # the kernel wrote the parameter through a raw pointer
        if st is not None:
            st.key = layer._staging_key(mode)
            st.from_optimizer = True
            layer._w_cache = st
        else:
            layer.invalidate_staging()
        return loss
