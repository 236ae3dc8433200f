"""GPU parity at the large BASELINE.json shapes (``python -m pytest tests -m gpu``):

* config 4 at its FULL shape - 512 rows x 40x40 map x 49 152 latent dims (Tiny-ImageNet-shaped ViT latents,
  ``/root/reference/configs/vit_som/vit_som_tiny-imagenet.yaml:4,16-17,46`` with BASELINE's 40x40 map), euclidean and
  cosine, against the fp64 oracle;
* a slice of config 5 - 128x128 map (K = 16 384), D = 256, 8 192 rows processed as two 4 096-row module calls that
  accumulate the prototype gradient in place through ``layer.grad_accumulator`` (the path ``bench.py`` runs config 5
  on), against the fp64 oracle evaluated in row chunks.

Tolerances (north_star): BMU exact except fp32 near-ties (classified with the fp64 oracle), loss / gradients 1e-5.
"""
import numpy as np
import pytest
import torch

from oracle import som_oracle as O
from oracle.ref_import import make_config

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("som_precision")]

LOSS_TOL = 1e-5
GRAD_TOL = 1e-5


def oracle_in_chunks(x, W, pos, T, fcn, bmu, rows=2048):
    """fp64 oracle of the GLOBAL-batch step evaluated `rows` rows at a time (bounded memory): loss = mean over all
    B x K, gradients of that mean; BMUs given.  Returns (loss, grad_x, grad_w, hard mismatches, worst gap)."""
    B = x.shape[0]
    loss, gw, gx = 0.0, np.zeros(W.shape, np.float64), np.empty(x.shape, np.float64)
    hard, worst = 0, 0.0
    for r0 in range(0, B, rows):
        xs, bs = x[r0:r0 + rows], bmu[r0:r0 + rows]
        frac = xs.shape[0] / B
        r = O.step(xs, W, pos, T, fcn, frac, np.float64, bmu_override=bs)      # g_out = share of the global mean
        loss += float(r.loss) * frac
        gx[r0:r0 + rows] = r.grad_x
        gw += r.grad_w
        want = O.bmu(r.distances)
        for i in np.nonzero(want != bs)[0]:
            a, b_ = r.distances[i, want[i]], r.distances[i, bs[i]]
            gap = abs(b_ - a) / (max(abs(a), abs(b_), 1e-30) if fcn == "euclidean" else 1.0)
            worst = max(worst, gap)
            hard += gap > 4e-6
    return loss, gx, gw, hard, worst


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_config4_full_shape(fcn, cuda_device):
    from vit_som_b200 import SOMLayer
    B, ms, D, T = 512, (40, 40), 49152, 20.0
    torch.manual_seed(0)
    layer = SOMLayer(make_config(list(ms), D, fcn, Tmax=T)).cuda()
    layer.current_temperature = T
    x = torch.randn(B, D)
    xg = x.cuda().requires_grad_(True)
    d, bmu = layer(xg)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    loss.backward()
    torch.cuda.synchronize()
    W = layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions(ms)
    bmu_np = bmu.cpu().numpy()
    ref_loss, ref_gx, ref_gw, hard, worst = oracle_in_chunks(x.numpy(), W, pos, T, fcn, bmu_np, rows=256)
    assert hard == 0, f"BMU mismatch that is not an fp32 near-tie (relative gap {worst:.3e})"
    assert abs(loss.item() - ref_loss) <= LOSS_TOL * abs(ref_loss)
    assert O.rel_err(xg.grad.cpu().numpy(), ref_gx) < GRAD_TOL
    assert O.rel_err(layer.prototypes.grad.cpu().numpy(), ref_gw) < GRAD_TOL
    # distances of a row band against the fp64 oracle
    d64 = O.distances(x.numpy()[:64], W, fcn, np.float64)
    assert O.rel_err(d.detach().cpu().numpy()[:64], d64) < 3e-6


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_config5_slice_through_grad_accumulator(fcn, cuda_device):
    """Two 4096-row chunks of the 128x128 map: in-place accumulation of dW in the GEMM epilogue (accumulate = 1)."""
    from vit_som_b200 import SOMLayer
    B, chunk, ms, D, T = 8192, 4096, (128, 128), 256, 64.0
    torch.manual_seed(1)
    layer = SOMLayer(make_config(list(ms), D, fcn, Tmax=T)).cuda()
    layer.current_temperature = T
    layer.grad_accumulator = torch.zeros(ms[0] * ms[1], D, device="cuda")
    x = torch.randn(B, D)
    chunks = [x[r0:r0 + chunk].cuda().requires_grad_(True) for r0 in range(0, B, chunk)]
    total, bmus = 0.0, []
    if fcn == "cosine":
        layer.batch_rows = B                                  # the layer divides by the rows of the whole batch itself
    for xc in chunks:
        d, bmu = layer(xc)
        loss = layer.som_loss(layer.compute_weights(bmu), d) * (1.0 if fcn == "cosine" else chunk / B)
        loss.backward()
        total += loss.item()
        bmus.append(bmu.cpu().numpy())
    torch.cuda.synchronize()
    assert layer.prototypes.grad is None                      # the gradient went into the accumulator, not to autograd
    W = layer.prototypes.detach().cpu().numpy()
    bmu_np = np.concatenate(bmus)
    ref_loss, ref_gx, ref_gw, hard, worst = oracle_in_chunks(x.numpy(), W, O.grid_positions(ms), T, fcn, bmu_np, rows=1024)
    assert hard == 0, f"BMU mismatch that is not an fp32 near-tie (relative gap {worst:.3e})"
    assert abs(total - ref_loss) <= LOSS_TOL * abs(ref_loss)
    gx = np.concatenate([xc.grad.cpu().numpy() for xc in chunks])
    assert O.rel_err(gx, ref_gx) < GRAD_TOL
    assert O.rel_err(layer.grad_accumulator.cpu().numpy(), ref_gw) < GRAD_TOL
    # accumulating the same two chunks again doubles the buffer exactly (deterministic kernels, exact fp32 doubling)
    first = layer.grad_accumulator.clone()
    for xc in chunks:
        xc.grad = None
        d, bmu = layer(xc)
        (layer.som_loss(layer.compute_weights(bmu), d) * (1.0 if fcn == "cosine" else chunk / B)).backward()
    torch.cuda.synchronize()
    assert O.rel_err(layer.grad_accumulator.cpu().numpy(), 2.0 * first.cpu().numpy()) < 2e-7
