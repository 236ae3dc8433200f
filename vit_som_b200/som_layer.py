"""Drop-in replacement for the reference ``SOMLayer`` (``/root/reference/models/som_layer.py:8-152``).

Same constructor dict, same attributes (``prototypes``, ``grid_positions``, ``map_size``, ``n_prototypes``,
``latent_dim``, ``Tmax``, ``Tmin``, ``topology``, ``distance_fcn``, ``current_temperature``, ``use_reduced``),
same state-dict keys and the same four calls ``ViTSOM`` makes (``models/vit_som.py:75,84-86``):

    distances, bmu = layer(x)                 # tcgen05 3xTF32 GEMM + fused distance/argmin epilogue
    layer.update_temperature(iteration)       # same schedule, stays a device tensor (no host sync)
    weights = layer.compute_weights(bmu)      # lazy handle: materialised only if somebody reads it
    loss = layer.som_loss(weights, distances) # fused neighbourhood-weighted reduction, fused backward

Everything numeric runs in ``libsom_b200.so``; CPU tensors and a missing library raise.
``manhattan`` (DESOM only, not a contraction) is out of scope and raises ``NotImplementedError``.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import SomError

try:                                                    # the reference derives from pl.LightningModule so that
    import pytorch_lightning as _pl                     # `.trainer` propagates to the child (som_layer.py:131)
    _Base = _pl.LightningModule
    if getattr(_pl, "__som_stub__", False):             # the oracle's import stub is not Lightning
        raise ImportError
except Exception:                                       # noqa: BLE001  (Lightning is not installed in this image)
    _Base = nn.Module


class NeighbourhoodWeights:
    """Lazy result of ``compute_weights``: (bmu, temperature snapshot).  ``som_loss`` consumes it without ever
    writing the B x K weight matrix; any other use (``w * d``, ``w.sum()``, ``w[:3]``, torch.* functions)
    materialises it once through the CUDA kernel ``som_neighbourhood``."""

    def __init__(self, layer: "SOMLayer", bmu: torch.Tensor, T_dev: torch.Tensor):
        self._layer, self.bmu, self.T_dev = layer, bmu, T_dev
        self._dense = None

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            lay = self._layer
            self._dense = ops.neighbourhood(self.bmu, lay.grid_positions, self.T_dev, lay.n_prototypes)
        return self._dense

    @property
    def shape(self):
        return torch.Size((self.bmu.shape[0], self._layer.n_prototypes))

    requires_grad = False                                # no grad flows through compute_weights

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.materialize() if isinstance(a, cls) else a   # noqa: E731
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def __mul__(self, other):
        return self.materialize() * other

    __rmul__ = __mul__

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __getattr__(self, name):                         # .sum(), .cpu(), .dtype, ... of the dense tensor
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class SOMLayer(_Base):
    """Self-organising-map layer, B200-native (see module docstring)."""

    def __init__(self, config):
        super().__init__()
        hp = config["hyperparameters"]
        self.model_arch = hp["model_arch"]
        som_hp = hp["som"]
        vit_hp = hp["vit"] if self.model_arch == "vit_som" else None
        data_hp = config["data"]

        self.total_epochs = hp["total_epochs"]
        self.batch_size = hp["batch_size"]
        self.map_size = som_hp["map_size"]
        self.Tmax = som_hp["Tmax"]
        self.Tmin = som_hp["Tmin"]
        self.topology = som_hp["topology"]
        self.distance_fcn = som_hp["distance_fcn"]
        self.n_prototypes = int(np.prod(self.map_size))

        self.use_reduced = som_hp["use_reduced"] if self.model_arch == "vit_som" else False
        latent_dim = vit_hp["emb_dim"] if self.model_arch == "vit_som" else hp["ae"]["encoder_dims"][-1]
        if not self.use_reduced and self.model_arch == "vit_som":
            latent_dim *= (data_hp["input_size"] // vit_hp["patch_size"]) ** 2
        self.latent_dim = latent_dim
        self.current_temperature = self.Tmax

        if self.distance_fcn not in ("euclidean", "cosine", "manhattan"):
            raise ValueError(f"Unsupported distance function: {self.distance_fcn}")
        init = torch.rand(self.n_prototypes, self.latent_dim)            # same RNG draw as the reference (:46-56)
        if self.distance_fcn == "cosine":
            init = torch.nn.functional.normalize(init, p=2, dim=1)
        self.prototypes = nn.Parameter(init)
        self.create_grid_positions()

        # optional, non-reference knobs
        self.total_iterations = som_hp.get("total_iterations")            # overrides the trainer-derived count
        if _Base is nn.Module:
            self.trainer = None
        self._w_cache = None
        # optional [K, D] fp32 buffer: when set, backward ADDS the prototype gradient into it from the GEMM epilogue
        # and returns no gradient for `prototypes` to autograd (row-chunked batches accumulate without extra passes)
        self.grad_accumulator = None
        # optional: rows of the WHOLE batch when it is fed in row chunks - som_loss then divides by batch_rows * K, so
        # the chunk losses (and their gradients) simply add up to the loss of the whole batch (None: rows of the call)
        self.batch_rows = None
        # operand precision of the tensor-core contractions: "tf32x3" (3xTF32) or "fp16x3" (3xFP16 on row-scaled operands:
        # the same 22-bit split at twice the tensor-core rate); None = ops.DEFAULT_PRECISION.  Both are fp32-accurate.
        self.precision = som_hp.get("precision")
        self._dw_hook = None                                              # data-parallel wrapper: called with dW as soon as it is enqueued
        self._dw_out = None                                               # data-parallel wrapper: [K, D] buffer the dW GEMM writes (NVLS symmetric memory)

    # ---- construction helpers -----------------------------------------------------------------
    def create_grid_positions(self):
        rows, cols = int(self.map_size[0]), int(self.map_size[1])
        if self.topology == "square":
            r = torch.arange(rows).repeat_interleave(cols)
            c = torch.arange(cols).repeat(rows)
            positions = torch.stack([r, c], dim=1).float()                 # (row, col), k = row * cols + col
        elif self.topology == "hexa":
            k = torch.arange(rows * cols)
            r, c = k // cols, k % cols
            positions = torch.stack([c.float() + 0.5 * (r % 2).float(),
                                     (r.double() * np.sqrt(3) / 2).float()], dim=1)
        else:
            raise ValueError(f"Unsupported topology: {self.topology}")
        self.register_buffer("grid_positions", positions)

    # ---- hot path -----------------------------------------------------------------------------
    def _mode(self) -> int:
        if self.distance_fcn == "manhattan":
            raise NotImplementedError(
                "distance_fcn='manhattan' (DESOM, cdist p=1) is not a contraction and is out of scope of the "
                "B200 hot path; use 'euclidean' or 'cosine'")
        prec = getattr(self, "precision", None) or ops.DEFAULT_PRECISION
        if prec not in ops.PREC:
            raise ValueError(f"precision must be one of {sorted(ops.PREC)}, got {prec!r}")
        return ops.MODE[self.distance_fcn] | ops.PREC[prec]

    def _staging_key(self, mode: int):
        W = self.prototypes
        return (W.data_ptr(), W._version, mode, tuple(W.shape), W.device)

    def invalidate_staging(self):
        """Forget the cached tf32 staging of the prototypes.  The cache follows the parameter's version counter, which
        every in-place torch op on the parameter advances (optimizer steps, ``copy_``, ``load_state_dict``) - but writes
        through ``prototypes.data`` (or through a raw pointer) do not: call this after such a write."""
        self._w_cache = None

    def _staged_prototypes(self, mode: int):
        """(staging, needs_refill): tf32 hi/lo split (+ norms) of the prototypes.

        * training mode with gradients enabled: the prototypes change every step, so the staging is part of the step
          (always refilled) - unless the fused optimizer (``vit_som_b200.FusedPrototypeAdamW``) produced the staging of
          the current parameter version in its own pass;
        * under CUDA-graph capture: always refilled (a replay cannot re-run this Python check, and a captured
          ``stage_w = 0`` would freeze the prototypes of capture time) - again unless the optimizer owns the staging;
        * eval / no_grad: refilled only when the parameter changed (version counter, ``invalidate_staging``).
        A fresh buffer is used on every refill so that a backward still holding the previous staging is not overwritten."""
        key = self._staging_key(mode)
        ws = self._w_cache
        if ws is not None and ws.key == key:
            if ws.from_optimizer:
                return ws, False
            volatile = (self.training and torch.is_grad_enabled()) or (
                self.prototypes.is_cuda and torch.cuda.is_current_stream_capturing())
            if not volatile:
                return ws, False
        W = self.prototypes
        ws = ops.Staging(W.shape[0], W.shape[1], mode, W.device)
        ws.key = key
        self._w_cache = ws
        return ws, True

    def _forward_impl(self, x, want_dist=True):
        if x.dim() > 2:
            x = x.flatten(start_dim=1)
        if not x.is_cuda or not self.prototypes.is_cuda:
            raise SomError("SOMLayer runs on a B200 only: move the module and its input to cuda (no CPU path)")
        if x.dim() != 2 or x.shape[1] != self.latent_dim:
            raise ValueError(f"latent shape {tuple(x.shape)} does not end in latent_dim {self.latent_dim}")
        mode = self._mode()
        ws, refill = self._staged_prototypes(mode)
        try:
            state, bmu = ops.forward(x, self.prototypes, mode, ws, refill, want_dist=want_dist)
        except Exception:
            self.invalidate_staging()                    # never keep a staging that may not have been filled
            raise
        if not want_dist:
            return None, bmu
        state.x_in, state.W_in = x, self.prototypes
        state.grad_accum = self.grad_accumulator
        state.dw_out = self._dw_out
        dist = ops.DistanceFn.apply(x, self.prototypes, state)
        dist._som_state = state                          # lets som_loss take the fused path (no B x K autograd edge)
        return dist, bmu

    def forward(self, x):
        return self._forward_impl(x)

    def compute_distances(self, x):
        return self._forward_impl(x)[0]

    def best_matching_units(self, x):
        """argmin-only inference path (tools/evaluation.py:29-42 keeps only the BMUs): no B x K store."""
        with torch.no_grad():
            return self._forward_impl(x, want_dist=False)[1]

    def _temperature_tensor(self) -> torch.Tensor:
        T = self.current_temperature
        dev = self.prototypes.device
        if torch.is_tensor(T):
            return T.detach().to(device=dev, dtype=torch.float32).reshape(1)
        cached = getattr(self, "_t_cache", None)
        if cached is None or cached[0] != (float(T), dev):
            self._t_cache = ((float(T), dev), torch.full((1,), float(T), device=dev, dtype=torch.float32))
        return self._t_cache[1]

    def update_temperature(self, iteration):
        """T = Tmax (Tmin/Tmax)^(it / (total - 1)), total = len(dataset)/batch_size*epochs (som_layer.py:127-132)."""
        if self.total_iterations is not None:
            total_iterations = self.total_iterations
        else:
            total_iterations = (len(self.trainer.train_dataloader.dataset) / self.batch_size) * self.total_epochs
        self.current_temperature = self.Tmax * (self.Tmin / self.Tmax) ** (iteration / (total_iterations - 1))

    def index_to_position(self, indices):
        return torch.stack((indices // self.map_size[1], indices % self.map_size[1]), dim=1).float()

    def compute_weights(self, bmu_indices):
        if not bmu_indices.is_cuda:
            raise SomError("bmu_indices must live on the GPU (no CPU path)")
        return NeighbourhoodWeights(self, bmu_indices.to(torch.int64).contiguous(), self._temperature_tensor())

    def _square_grid_dims(self):
        """(rows, cols) when grid_positions is the canonical integer grid of the square topology (then the fused loss
        kernel evaluates the neighbourhood weight in its factorised form), else (0, 0)."""
        if self.topology == "square":
            return int(self.map_size[0]), int(self.map_size[1])
        return 0, 0

    def som_loss(self, weights, distances):
        B, K = distances.shape
        rows = B if self.batch_rows is None else int(self.batch_rows)
        if isinstance(weights, NeighbourhoodWeights) and weights._dense is None:
            state = getattr(distances, "_som_state", None)
            if state is not None and state.B == B and state.K == K:
                # distances are this layer's own forward output: loss and its backward as one node over (x, W)
                want_grad = torch.is_grad_enabled() and (state.x_in.requires_grad or state.W_in.requires_grad)
                return ops.FusedLossFn.apply(state.x_in, state.W_in, state, weights.bmu, self.grid_positions,
                                             weights.T_dev, 1.0 / (rows * K), 0, want_grad, self._dw_hook,
                                             self._square_grid_dims())
            return ops.WeightedLossFn.apply(distances, weights.bmu, self.grid_positions, weights.T_dev,
                                            1.0 / (rows * K), 0)
        dense = weights.materialize() if isinstance(weights, NeighbourhoodWeights) else weights
        loss = torch.mean(dense * distances)             # caller supplied its own weights: plain composition
        return loss if rows == B else loss * (B / rows)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate_staging()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)        # .to() / .cuda() / .float(): new storage, new staging
        self.invalidate_staging()
        self._t_cache = None
        return out

    # ---- the reference's own (dead) Lightning hooks, kept for API completeness -------------------
    def training_step(self, batch, batch_idx):
        x, _ = batch
        distances, bmu_indices = self.forward(x.view(x.size(0), -1))
        weights = self.compute_weights(bmu_indices)
        loss = self.som_loss(weights, distances)
        if hasattr(self, "log"):
            self.log("train_loss", loss)
        return loss
