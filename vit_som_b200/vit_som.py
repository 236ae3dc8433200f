"""ViT-SOM training step around the B200 SOM layer: the harness behind the
"ViT-SOM train img/s" half of BASELINE.json's metric.

Only the SOM layer of this model is a kernel target of the repository.  The ViT autoencoder below is plain PyTorch
with the reference's architecture and hyper-parameters (``/root/reference/models/vit.py:65-240``: conv patch embedding,
fixed 2-D sin-cos position table, class token, pre-norm blocks with qkv bias, 4x GELU MLP, LayerNorm eps 1e-6, a light
decoder that predicts pixels per patch) run under bf16 autocast with ``scaled_dot_product_attention``; it is written
from that description, not translated.  What the module adds on top of the reference's ``ViTSOM``
(``/root/reference/models/vit_som.py:67-105``):

* the SOM input is the strided view ``x[:, 1:].flatten(1)`` of the encoder output (``som_input``: no copy),
* the gamma ramp stays on the device (``gamma_ramp``: no ``.item()`` sync per step),
* the prototypes are stepped by ``FusedPrototypeAdamW`` (update + operand staging for the next forward in one pass),
* under data parallelism the ViT parameters go through torch DDP and the prototype gradient through
  ``DataParallelSOM`` (NVLS exchange started from inside the SOM backward, hidden under the ViT backward).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .glue import gamma_ramp, som_input
from .optim import FusedPrototypeAdamW
from .som_layer import SOMLayer


def sincos_table_2d(dim: int, side: int) -> torch.Tensor:
    """[1 + side*side, dim] fixed position table (row 0, the class token, is zero): half of the channels encode the
    x coordinate, half the y coordinate, each as sin | cos over geometrically spaced frequencies (MAE convention)."""
    if dim % 4:
        raise ValueError("embedding dim must be a multiple of 4 for the 2-D sin-cos table")
    freqs = 1.0 / (10000.0 ** (torch.arange(dim // 4, dtype=torch.float64) / (dim / 4.0)))
    coords = torch.arange(side, dtype=torch.float64)
    gy, gx = torch.meshgrid(coords, coords, indexing="ij")

    def axis(c):
        ang = c.reshape(-1, 1) * freqs.reshape(1, -1)
        return torch.cat([ang.sin(), ang.cos()], dim=1)
    table = torch.cat([axis(gx), axis(gy)], dim=1)
    return torch.cat([torch.zeros(1, dim, dtype=torch.float64), table], dim=0).float()


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.heads = heads
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        hidden = int(dim * mlp_ratio)
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v)
        x = x + self.proj(a.transpose(1, 2).reshape(B, N, C))
        return x + self.fc2(F.gelu(self.fc1(self.norm2(x))))


class ViTAutoencoder(nn.Module):
    """Encoder -> (class token, patch tokens); decoder -> reconstructed image."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, decoder_embed_dim, decoder_depth,
                 decoder_num_heads, mlp_ratio=4.0):
        super().__init__()
        if img_size % patch_size:
            raise ValueError("image size must be a multiple of the patch size")
        self.patch, self.side, self.chans = patch_size, img_size // patch_size, in_chans
        self.embed = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.register_buffer("pos", sincos_table_2d(embed_dim, self.side).unsqueeze(0), persistent=False)
        self.blocks = nn.ModuleList(_Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth))
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.dec_embed = nn.Linear(embed_dim, decoder_embed_dim)
        self.register_buffer("dec_pos", sincos_table_2d(decoder_embed_dim, self.side).unsqueeze(0), persistent=False)
        self.dec_blocks = nn.ModuleList(_Block(decoder_embed_dim, decoder_num_heads, mlp_ratio)
                                        for _ in range(decoder_depth))
        self.dec_norm = nn.LayerNorm(decoder_embed_dim, eps=1e-6)
        self.dec_pred = nn.Linear(decoder_embed_dim, patch_size * patch_size * in_chans)
        nn.init.normal_(self.cls_token, std=0.02)
        nn.init.xavier_uniform_(self.embed.weight.view(embed_dim, -1))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def encode(self, img):
        t = self.embed(img).flatten(2).transpose(1, 2) + self.pos[:, 1:]
        cls = (self.cls_token + self.pos[:, :1]).expand(t.shape[0], -1, -1)
        t = torch.cat([cls.to(t.dtype), t], dim=1)
        for blk in self.blocks:
            t = blk(t)
        return self.norm(t)

    def decode(self, tokens):
        t = self.dec_embed(tokens) + self.dec_pos
        for blk in self.dec_blocks:
            t = blk(t)
        pix = self.dec_pred(self.dec_norm(t))[:, 1:]                       # drop the class token's prediction
        B, p, s, c = pix.shape[0], self.patch, self.side, self.chans
        return pix.view(B, s, s, p, p, c).permute(0, 5, 1, 3, 2, 4).reshape(B, c, s * p, s * p)

    def forward(self, img, reconstruct: bool = True):
        tokens = self.encode(img)
        recon = self.decode(tokens) if reconstruct else None
        return tokens[:, 0], tokens[:, 1:], recon


class ViTSOM(nn.Module):
    """ViT autoencoder + SOM layer (+ classification head when ``data.num_classes > 0``), config dict in the
    reference's YAML schema (``/root/reference/configs/vit_som/*.yaml``)."""

    def __init__(self, config, som_layer: SOMLayer | None = None):
        super().__init__()
        hp, data = config["hyperparameters"], config["data"]
        vit = hp["vit"]
        self.config = config
        self.gamma = hp["gamma"]
        self.use_reduced = hp["som"]["use_reduced"]
        self.classification = data["num_classes"] > 0
        self.vit = ViTAutoencoder(data["input_size"], vit["patch_size"], data["num_channels"], vit["emb_dim"], vit["depth"],
                                  vit["heads"], vit["dec_emb_dim"], vit["dec_depth"], vit["heads"], vit["mlp_ratio"])
        self.som_layer = som_layer if som_layer is not None else SOMLayer(config)
        if self.classification:
            self.cls_head = nn.Linear(vit["emb_dim"], data["num_classes"])
            nn.init.normal_(self.cls_head.weight, std=0.02)
        self.smoothing = hp.get("optimizer", {}).get("smoothing", 0.0)
        self.register_buffer("iteration", torch.tensor(0))
        self.ramp_up_end_step = 1                                      # trainer.estimated_stepping_batches // 2 upstream

    def forward(self, img):
        """(class token, reconstruction or None, logits or None, distances, bmu) - the reference's return tuple."""
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=img.is_cuda):
            cls, patches, recon = self.vit(img, reconstruct=not self.classification or not self.training)
            logits = self.cls_head(cls) if self.classification else None
        # the SOM layer computes in fp32-accurate arithmetic: bf16 latents are widened exactly by its staging kernel
        distances, bmu = self.som_layer(som_input(cls, patches, self.use_reduced))
        return cls, recon, logits, distances, bmu

    def training_loss(self, img, labels=None):
        """Total loss of one training step (models/vit_som.py:80-105), sync-free; advances ``iteration``."""
        _, recon, logits, distances, bmu = self.forward(img)
        som = self.som_layer
        som.update_temperature(self.iteration)
        som_loss = som.som_loss(som.compute_weights(bmu), distances)
        g = gamma_ramp(self.iteration, self.ramp_up_end_step, self.gamma)
        if self.classification:
            task = F.cross_entropy(logits.float(), labels.view(-1), label_smoothing=self.smoothing)
        else:
            task = F.l1_loss(recon.float(), img)
        self.iteration += 1
        return task + g * som_loss, som_loss.detach()


def build_optimizers(model: ViTSOM, fused_prototypes: bool = True, capturable: bool = False):
    """AdamW as the reference configures it (models/vit_som.py:126-151): lr = yaml lr * batch / 256, betas from the
    YAML, weight decay on the >= 2-D ViT parameters, the default 0.01 on the SOM prototypes / classification head.
    Returns (optimizer of the ViT and head, optimizer of the prototypes).  ``capturable``: keep the step counts of the
    ViT optimizer on the device so that the whole training step can be captured in a CUDA graph (the prototype
    optimizer always is)."""
    hp = model.config["hyperparameters"]
    opt = hp["optimizer"]
    lr = opt["lr"] * hp["batch_size"] / 256
    betas = (opt["beta_1"], opt["beta_2"])
    decay = [p for p in model.vit.parameters() if p.requires_grad and p.ndim > 1]
    plain = [p for p in model.vit.parameters() if p.requires_grad and p.ndim <= 1]
    groups = [{"params": decay, "weight_decay": opt["weight_decay"]}, {"params": plain, "weight_decay": 0.0}]
    if model.classification:
        groups.append({"params": list(model.cls_head.parameters())})
    if not fused_prototypes:
        groups.append({"params": list(model.som_layer.parameters())})
        return torch.optim.AdamW(groups, lr=lr, betas=betas, fused=True, capturable=capturable), None
    return (torch.optim.AdamW(groups, lr=lr, betas=betas, fused=True, capturable=capturable),
            FusedPrototypeAdamW(model.som_layer, lr=lr, betas=betas))


class FlatGradBucket:
    """Gradients of a parameter list as views of ONE flat buffer, averaged over the ranks with one all-reduce.

    What DDP's bucketing does for the (small) ViT of ViT-SOM, reduced to a form that can be captured in a CUDA graph
    together with the rest of the training step: ``zero()`` before backward, ``all_reduce_mean()`` after it."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(n, device=first.device, dtype=first.dtype)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)


def reference_yaml_config(dataset: str, map_size, batch_size: int):
    """The shipped ViT-SOM YAML of ``dataset`` as a dict (only the keys the model reads), with the map size and batch
    size BASELINE.json names.  cifar-10: configs/vit_som/vit_som_cifar-10.yaml; tiny-imagenet: ..._tiny-imagenet.yaml."""
    shapes = {"cifar-10": (32, 10, 4.0, 0.01, 0.0005), "tiny-imagenet": (64, 200, 14.0, 0.5, 0.001)}
    if dataset not in shapes:
        raise ValueError(f"no YAML transcription for {dataset}")
    size, classes, tmax, gamma, lr = shapes[dataset]
    return {
        "hyperparameters": {
            "model_arch": "vit_som", "total_epochs": 500, "batch_size": batch_size, "gamma": gamma,
            "som": {"map_size": list(map_size), "Tmax": float(max(map_size)) if dataset != "cifar-10" else tmax,
                    "Tmin": 0.1, "distance_fcn": "cosine", "topology": "square", "use_reduced": False},
            "vit": {"patch_size": 4, "emb_dim": 192, "depth": 12, "dec_emb_dim": 96, "dec_depth": 2, "heads": 3,
                    "mlp_ratio": 4},
            "optimizer": {"type": "adamw", "lr": lr, "beta_1": 0.9, "beta_2": 0.999, "weight_decay": 0.05,
                          "smoothing": 0.1},
        },
        "data": {"dataset": dataset, "num_classes": classes, "num_channels": 3, "input_size": size},
    }


def steps_per_epoch_estimate(n_samples: int, batch_size: int) -> int:
    return max(1, math.ceil(n_samples / batch_size))
