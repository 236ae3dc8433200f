"""Inference-side helpers around the argmin-only kernel path.

The reference's clustering evaluation (``/root/reference/tools/evaluation.py:18-52``) runs the full model per batch,
keeps only the BMU indices, moves them to the host one batch at a time and scores purity / NMI with Python loops and
scikit-learn; its prototype visualisation decodes the K prototypes one by one in a Python loop (``:181-183``).  Here:

* ``assign_bmus``       latents -> BMU indices through ``SOMLayer.best_matching_units`` (tcgen05 GEMM + argmin epilogue
                        only: no B x K distance store, no loss), in row chunks, results stay on the device;
* ``purity_nmi``        purity (majority vote per map cell, ``calculate_purity``) and NMI (arithmetic normalisation,
                        scikit-learn's default) from ONE contingency table built with ``bincount`` on the device;
* ``decode_prototypes`` all prototypes through the ViT decoder in batches instead of K single-row passes.
The metrics are plain tensor arithmetic (they also run on CPU tensors, which is how the unit tests pin them against
scikit-learn); the BMU search has no CPU path.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def assign_bmus(layer, latents: torch.Tensor, row_chunk: int = 16384) -> torch.Tensor:
    """BMU index (int64, on the latents' device) of every row of ``latents`` [N, D] (or [N, ...] flattened)."""
    out = []
    for r0 in range(0, latents.shape[0], row_chunk):
        out.append(layer.best_matching_units(latents[r0:r0 + row_chunk]))
    return torch.cat(out) if len(out) != 1 else out[0]


def contingency(cells: torch.Tensor, labels: torch.Tensor, n_cells: int, n_classes: int) -> torch.Tensor:
    """[n_cells, n_classes] int64 counts of (BMU cell, true label) pairs."""
    cells, labels = cells.reshape(-1).long(), labels.reshape(-1).long()
    if cells.numel() != labels.numel():
        raise ValueError(f"{cells.numel()} assignments but {labels.numel()} labels")
    flat = torch.bincount(cells * n_classes + labels, minlength=n_cells * n_classes)
    return flat.view(n_cells, n_classes)


def purity_nmi(cells: torch.Tensor, labels: torch.Tensor, n_cells: int | None = None, n_classes: int | None = None):
    """(purity, nmi) as 0-dim float64 tensors on the inputs' device.

    purity = (1/N) sum_cells max_class count   (tools/evaluation.py:132-152: accuracy after majority voting)
    nmi    = I(cell; label) / ((H(cell) + H(label)) / 2)   (sklearn.metrics.normalized_mutual_info_score default)"""
    n_cells = int(cells.max().item()) + 1 if n_cells is None else n_cells
    n_classes = int(labels.max().item()) + 1 if n_classes is None else n_classes
    table = contingency(cells, labels, n_cells, n_classes).double()
    n = table.sum()
    purity = table.max(dim=1).values.sum() / n
    pc, py = table.sum(1) / n, table.sum(0) / n

    def entropy(p):
        p = p[p > 0]
        return -(p * p.log()).sum()
    joint = table / n
    nz = joint > 0
    outer = pc[:, None] * py[None, :]
    mi = (joint[nz] * (joint[nz] / outer[nz]).log()).sum()
    hc, hy = entropy(pc), entropy(py)
    denom = (hc + hy) / 2
    # sklearn: a single cluster on both sides is a perfect match; otherwise MI / mean entropy, clipped at 0
    nmi = torch.where(denom > 0, mi.clamp_min(0) / denom.clamp_min(1e-300), torch.ones_like(mi))
    return purity, nmi


@torch.no_grad()
def decode_prototypes(vit, prototypes: torch.Tensor, num_patches: int, embed_dim: int, batch: int = 256):
    """[K, C, H, W] images of all prototypes (tools/evaluation.py:153-222 decodes them one at a time): each prototype
    [num_patches * embed_dim] is reshaped to patch tokens, a zero class-token slot is prepended and the decoder of
    ``vit`` (a module with ``decode(tokens)``, e.g. ``vit_som_b200.vit_som.ViTAutoencoder``) reconstructs the image."""
    K = prototypes.shape[0]
    if prototypes.shape[1] != num_patches * embed_dim:
        raise ValueError("Prototype dimensions mismatch for decoding.")
    out = []
    for k0 in range(0, K, batch):
        p = prototypes[k0:k0 + batch].reshape(-1, num_patches, embed_dim)
        tokens = torch.cat([p.new_zeros(p.shape[0], 1, embed_dim), p], dim=1)
        out.append(vit.decode(tokens).float())
    return torch.cat(out)
