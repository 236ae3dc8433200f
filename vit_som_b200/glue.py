"""Sync-free glue for the caller of the SOM layer.

The reference's ``ViTSOM.training_step`` (``/root/reference/models/vit_som.py:80-105``) forces one device -> host
synchronisation per step: the gamma ramp reads ``self.iteration.item()`` (``:90``).  With the SOM step down to
~0.2 ms that stall is of the order of the step itself.  The two helpers below keep the same arithmetic on the device;
the SOM loss then enters the total loss as ``gamma_t * som_loss`` with ``gamma_t`` a 0-dim tensor, and the backward
of the SOM layer reads that factor from device memory in the epilogue of its gradient GEMMs (no host round trip
anywhere on the path).  Plain torch tensor arithmetic on scalars - no kernels of the hot path live here.
"""
from __future__ import annotations

import torch


def gamma_ramp(iteration, ramp_up_end_step: int, gamma: float):
    """``gamma * min(1.0, iteration / ramp_up_end_step)`` (models/vit_som.py:89-90).

    ``iteration`` may be the module's 0-dim integer buffer (any device): the result is then a 0-dim float32 tensor on
    that device and nothing synchronises.  A Python number gives a Python float, exactly as in the reference."""
    if torch.is_tensor(iteration):
        frac = iteration.to(torch.float32) / float(ramp_up_end_step)
        return float(gamma) * torch.clamp(frac, max=1.0)
    return gamma * min(1.0, iteration / ramp_up_end_step)


def som_input(cls_token: torch.Tensor, patches: torch.Tensor, use_reduced: bool) -> torch.Tensor:
    """The latent the SOM layer scores (models/vit_som.py:69-73): the CLS token, or all patch tokens flattened.

    ``patches`` usually is the view ``x[:, 1:]`` of the encoder output (models/vit.py:222); flattening its last two
    dimensions is again a view, with row stride (N + 1) * E.  ``SOMLayer.forward`` takes that stride as the leading
    dimension of the staging kernel and of the gradient epilogue - no contiguous() copy of the [B, N * E] latent is
    made anywhere (a caller that adds ``.contiguous()`` pays 2 * B * N * E * 4 bytes of HBM traffic per step)."""
    if use_reduced:
        return cls_token
    return patches.flatten(start_dim=1)          # a view (torch merges the two trailing dimensions), not a copy
