"""Torch-CPU restatement of the reference SOM step  —  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's arithmetic is a sequence of ATen calls (``/root/reference/models/som_layer.py``):
``torch.cdist`` / ``F.normalize``+``torch.mm`` (:111-125), ``torch.argmin`` (:88), ``torch.norm``+``torch.exp``
(:148-150), ``torch.mean`` (:141-142) and autograd's backward through them.  ATen ships with the torch wheel
on the GPU box, /root/reference does not, so the CPU baseline that ``bench.py`` times there is this functional
restatement issuing the *same* ATen calls in the same order (kind = "port").  It is pinned against the
reference's outputs by ``tests/test_oracle_golden.py::test_torch_port_matches_reference``.

Only ``tests/`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def distances(x: torch.Tensor, W: torch.Tensor, fcn: str) -> torch.Tensor:
    """compute_distances (models/som_layer.py:111-125)."""
    if fcn == "euclidean":
        return torch.cdist(x, W, p=2)
    if fcn == "cosine":
        return 1 - torch.mm(F.normalize(x, p=2, dim=1), F.normalize(W, p=2, dim=1).T)
    raise ValueError(f"Unsupported distance function: {fcn}")


def neighbourhood(bmu: torch.Tensor, grid_positions: torch.Tensor, T) -> torch.Tensor:
    """compute_weights (models/som_layer.py:144-152)."""
    pb = grid_positions[bmu]
    g = torch.norm(grid_positions.unsqueeze(0) - pb.unsqueeze(1), dim=2)
    return torch.exp(-g ** 2 / (2 * T ** 2))


def step(x: torch.Tensor, W: torch.Tensor, grid_positions: torch.Tensor, T, fcn: str, g_out: float = 1.0):
    """forward -> compute_weights -> som_loss -> backward (models/vit_som.py:82-86 + loss.backward()).
    Returns (distances, bmu, loss, grad_x, grad_w)."""
    x = x.detach().requires_grad_(True)
    W = W.detach().requires_grad_(True)
    xf = x.flatten(start_dim=1) if x.dim() > 2 else x
    d = distances(xf, W, fcn)
    bmu = torch.argmin(d, dim=1)
    w = neighbourhood(bmu, grid_positions, T)
    loss = torch.mean(w * d)
    (loss * g_out).backward()
    return d.detach(), bmu, loss.detach(), x.grad, W.grad
