"""GPU tests of the paths around the main training step (``python -m pytest tests -m gpu``):

* the copy-free strided SOM input of ``ViTSOM`` (``x[:, 1:].flatten(1)``, models/vit_som.py:69-73) through the staging
  kernel and the gradient epilogue,
* the general autograd path (``DistanceFn.backward`` with an arbitrary upstream gradient of ``distances``,
  ``WeightedLossFn`` on caller-supplied distances),
* CUDA-graph replay with prototype updates between replays (``StepGraph``),
* run-to-run bit-identical gradients (no float atomics anywhere on the path),
* programmatic dependent launch on / off,
* the fused prototype AdamW against ``torch.optim.AdamW``,
* the argmin-only evaluation helpers and the ViT-SOM training step harness.
"""
import numpy as np
import pytest
import torch

from oracle import som_oracle as O
from oracle.ref_import import make_config

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("som_precision")]

LOSS_TOL = 1e-5
GRAD_TOL = 1e-5


def make_layer(ms, D, fcn, T=None, topology="square"):
    from vit_som_b200 import SOMLayer
    layer = SOMLayer(make_config(list(ms), D, fcn, topology=topology)).cuda()
    if T is not None:
        layer.current_temperature = T
    return layer


def step(layer, x, g_out=1.0):
    layer.prototypes.grad = None
    if x.grad is not None:
        x.grad = None
    d, bmu = layer(x)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    (loss * g_out).backward()
    return d, bmu, loss


# ------------------------------------------------------------------------------------------------------------------
# strided (copy-free) SOM input
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
@pytest.mark.parametrize("tokens", [(49, 16), (16, 24), (5, 7)])
def test_strided_patch_token_view(fcn, tokens, cuda_device):
    """The SOM input is the view x[:, 1:].flatten(1) of the encoder output [B, N + 1, E]: row stride (N + 1) * E, base
    offset E floats.  (5, 7): a row pitch that is not a multiple of 4 floats - the staging kernel's scalar path.)"""
    from vit_som_b200 import som_input
    N, E = tokens
    B, ms, T = 96, (9, 8), 2.5
    torch.manual_seed(3)
    layer = make_layer(ms, N * E, fcn, T)
    enc = torch.randn(B, N + 1, E, device="cuda", requires_grad=True)        # encoder output incl. the class token
    view = som_input(enc[:, 0], enc[:, 1:], use_reduced=False)
    assert view.data_ptr() == enc.data_ptr() + 4 * E and view.stride(0) == (N + 1) * E and not view.is_contiguous()
    d, bmu = layer(view)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    (0.3 * loss).backward()
    torch.cuda.synchronize()
    x_np = enc.detach()[:, 1:].reshape(B, -1).cpu().numpy()
    W = layer.prototypes.detach().cpu().numpy()
    ref = O.step(x_np, W, O.grid_positions(ms), T, fcn, 0.3, np.float64, bmu_override=bmu.cpu().numpy())
    _, hard, worst = O.classify_bmu_mismatches(x_np, W, bmu.cpu().numpy(), fcn)
    assert hard == 0, worst
    assert O.rel_err(d.detach().cpu().numpy(), ref.distances) < 3e-6
    assert abs(loss.item() - float(ref.loss)) <= LOSS_TOL * abs(float(ref.loss))
    g = enc.grad.cpu().numpy()
    assert np.all(g[:, 0] == 0)                                               # the class token gets no SOM gradient
    assert O.rel_err(g[:, 1:].reshape(B, -1), ref.grad_x) < GRAD_TOL
    assert O.rel_err(layer.prototypes.grad.cpu().numpy(), ref.grad_w) < GRAD_TOL
    # same numbers as the contiguous copy the reference would have made
    xc = enc.detach()[:, 1:].reshape(B, -1).contiguous().requires_grad_(True)
    d2, bmu2, _ = step(layer, xc, 0.3)
    assert torch.equal(bmu, bmu2) and torch.equal(d, d2)


# ------------------------------------------------------------------------------------------------------------------
# general autograd path
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_distance_backward_with_arbitrary_upstream_gradient(fcn, cuda_device):
    """A caller that uses `distances` in its own expression: autograd reaches DistanceFn.backward with a dense
    upstream gradient (kernels bwd_coeffs + the two gradient GEMMs with explicit coefficients)."""
    B, ms, D = 150, (7, 9), 200
    K = ms[0] * ms[1]
    torch.manual_seed(5)
    layer = make_layer(ms, D, fcn, 2.0)
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    M = torch.randn(B, K, device="cuda")
    d, _ = layer(x)
    (d * M).sum().backward()
    torch.cuda.synchronize()
    # fp64 torch autograd of the reference's own formula (models/som_layer.py:117-122)
    x64 = x.detach().double().cpu().requires_grad_(True)
    W64 = layer.prototypes.detach().double().cpu().requires_grad_(True)
    if fcn == "euclidean":
        d64 = torch.cdist(x64, W64, p=2)
    else:
        d64 = 1 - torch.nn.functional.normalize(x64, p=2, dim=1) @ torch.nn.functional.normalize(W64, p=2, dim=1).t()
    (d64 * M.double().cpu()).sum().backward()
    assert O.rel_err(d.detach().cpu().numpy(), d64.detach().numpy()) < 3e-6
    assert O.rel_err(x.grad.cpu().numpy(), x64.grad.numpy()) < GRAD_TOL
    assert O.rel_err(layer.prototypes.grad.cpu().numpy(), W64.grad.numpy()) < GRAD_TOL


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_weighted_loss_on_caller_supplied_distances(fcn, cuda_device):
    """som_loss(lazy weights, a distance tensor that is NOT this layer's forward output): WeightedLossFn (loss kernel +
    dense gradient kernel), composed by autograd with the caller's own ops and with DistanceFn.backward."""
    B, ms, D, T = 130, (6, 11), 96, 1.7
    torch.manual_seed(6)
    layer = make_layer(ms, D, fcn, T)
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    d, bmu = layer(x)
    d_mod = d * 1.5 + 0.25                                         # a new tensor: the fused path does not apply
    loss = layer.som_loss(layer.compute_weights(bmu), d_mod)
    loss.backward()
    torch.cuda.synchronize()
    x_np, W = x.detach().cpu().numpy(), layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions(ms)
    ref = O.step(x_np, W, pos, T, fcn, 1.5, np.float64, bmu_override=bmu.cpu().numpy())
    w = O.weights(bmu.cpu().numpy(), pos, T, np.float64)
    ref_loss = float((w * (1.5 * ref.distances + 0.25)).mean())
    assert abs(loss.item() - ref_loss) <= LOSS_TOL * abs(ref_loss)
    assert O.rel_err(x.grad.cpu().numpy(), ref.grad_x) < GRAD_TOL                     # d(loss)/dd = 1.5 * w / (B K)
    assert O.rel_err(layer.prototypes.grad.cpu().numpy(), ref.grad_w) < GRAD_TOL


def test_use_reduced_false_construction(cuda_device):
    """latent_dim = emb_dim * (input_size // patch_size)^2 when use_reduced is False (models/som_layer.py:35-40)."""
    from vit_som_b200 import SOMLayer
    cfg = make_config([4, 4], 16, "cosine")
    cfg["hyperparameters"]["som"]["use_reduced"] = False
    cfg["hyperparameters"]["vit"]["patch_size"] = 4
    cfg["data"]["input_size"] = 28
    layer = SOMLayer(cfg).cuda()
    assert layer.latent_dim == 16 * 49 and tuple(layer.prototypes.shape) == (16, 784) and layer.use_reduced is False
    d, bmu = layer(torch.randn(8, 49, 16, device="cuda"))         # [B, N, E] patch tokens: flattened by forward
    assert d.shape == (8, 16) and bmu.shape == (8,)


# ------------------------------------------------------------------------------------------------------------------
# CUDA graph replay, determinism, PDL
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_step_graph_sees_prototype_updates(fcn, cuda_device):
    """Capture the module-API step, update the prototypes eagerly (SGD), replay: the replay must compute with the NEW
    prototypes (round-1 defect: the captured step kept the staging of capture time)."""
    from vit_som_b200 import StepGraph
    B, ms, D, T = 256, (12, 10), 320, 3.0
    torch.manual_seed(8)
    layer = make_layer(ms, D, fcn, T)
    layer.train()
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    side = torch.cuda.Stream()

    def fn():
        layer.prototypes.grad = None
        x.grad = None
        d, bmu = layer(x)
        loss = layer.som_loss(layer.compute_weights(bmu), d)
        loss.backward()
        return loss.detach(), bmu
    with torch.cuda.stream(side):
        g = StepGraph(fn, warmup=2, stream=side)
    side.synchronize()
    gw_static, gx_static = layer.prototypes.grad, x.grad          # graph-owned: every replay rewrites them
    for it in range(3):
        with torch.no_grad():                                      # an optimizer step outside the captured function
            layer.prototypes.add_(torch.randn_like(layer.prototypes) * 0.05)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            loss_g, bmu_g = g.replay()
        side.synchronize()
        gw_graph, gx_graph = gw_static.clone(), gx_static.clone()
        loss_graph, bmu_graph = loss_g.clone(), bmu_g.clone()
        xe = x.detach().clone().requires_grad_(True)
        d_e, bmu_e, loss_e = step(layer, xe)
        torch.cuda.synchronize()
        assert torch.equal(bmu_graph, bmu_e), f"replay {it} used stale prototypes"
        assert loss_graph.item() == loss_e.item()
        assert torch.equal(gw_graph, layer.prototypes.grad) and torch.equal(gx_graph, xe.grad)


@pytest.mark.parametrize("shape", [(1024, (40, 40), 3136), (700, (24, 24), 520), (4096, (64, 64), 256)])
def test_gradients_are_bit_identical_run_to_run(shape, cuda_device):
    """No float atomics on the path: loss, BMUs, distances AND both gradients repeat bit for bit."""
    B, ms, D = shape
    torch.manual_seed(9)
    layer = make_layer(ms, D, "euclidean", 4.0)
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    outs = []
    for _ in range(3):
        d, bmu, loss = step(layer, x, 0.7)
        torch.cuda.synchronize()
        outs.append((d.detach().clone(), bmu.clone(), loss.item(), x.grad.clone(), layer.prototypes.grad.clone()))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and o[2] == outs[0][2]
        assert torch.equal(o[3], outs[0][3]), "dx differs between runs"
        assert torch.equal(o[4], outs[0][4]), "dW differs between runs"


def test_programmatic_dependent_launch_does_not_change_results(cuda_device):
    from vit_som_b200 import _lib
    B, ms, D = 512, (20, 20), 768
    torch.manual_seed(10)
    layer = make_layer(ms, D, "cosine", 3.0)
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    L = _lib.lib()
    res = {}
    for on in (1, 0, 1):
        L.som_set_pdl(on)
        try:
            d, bmu, loss = step(layer, x)
            torch.cuda.synchronize()
        finally:
            L.som_set_pdl(1)
        cur = (d.detach().clone(), bmu.clone(), loss.item(), x.grad.clone(), layer.prototypes.grad.clone())
        if on in res:
            continue
        res[on] = cur
    for a, b in zip(res[1], res[0]):
        assert (a == b) if isinstance(a, float) else torch.equal(a, b)


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
@pytest.mark.parametrize("shape", [(1024, (40, 40), 512), (300, (24, 20), 96), (4096, (64, 64), 64), (37, (8, 12), 40)])
def test_fast_loss_kernel_matches_generic(shape, fcn, cuda_device):
    """The straight-line loss kernel for aligned square maps against the generic one (som_set_debug bit 5 forces it):
    same BMUs and distances, loss and gradients equal to rounding; both deterministic."""
    from vit_som_b200 import _lib
    B, ms, D = shape
    torch.manual_seed(15)
    layer = make_layer(ms, D, fcn, 2.5)
    x = torch.randn(B, D, device="cuda", requires_grad=True)
    L = _lib.lib()
    out = {}
    for bits in (0, 32, 0):
        L.som_set_debug(bits)
        try:
            d, bmu, loss = step(layer, x, 0.9)
            torch.cuda.synchronize()
        finally:
            L.som_set_debug(0)
        cur = (d.detach().clone(), bmu.clone(), loss.item(), x.grad.clone(), layer.prototypes.grad.clone())
        if bits in out:
            for a, b in zip(cur, out[bits]):                       # the fast path repeats bit for bit
                assert (a == b) if isinstance(a, float) else torch.equal(a, b)
        out[bits] = cur
    fast, gen = out[0], out[32]
    assert torch.equal(fast[0], gen[0]) and torch.equal(fast[1], gen[1])
    assert abs(fast[2] - gen[2]) <= 2e-6 * abs(gen[2])
    assert O.rel_err(fast[3].cpu().numpy(), gen[3].cpu().numpy()) < 2e-6
    assert O.rel_err(fast[4].cpu().numpy(), gen[4].cpu().numpy()) < 2e-6
    r64 = O.step(x.detach().cpu().numpy(), layer.prototypes.detach().cpu().numpy(), O.grid_positions(ms), 2.5, fcn, 0.9,
                 np.float64, bmu_override=fast[1].cpu().numpy())
    assert abs(fast[2] - float(r64.loss)) <= LOSS_TOL * abs(float(r64.loss))
    assert O.rel_err(fast[3].cpu().numpy(), r64.grad_x) < GRAD_TOL
    assert O.rel_err(fast[4].cpu().numpy(), r64.grad_w) < GRAD_TOL


# ------------------------------------------------------------------------------------------------------------------
# fused prototype AdamW
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
@pytest.mark.parametrize("shape", [((10, 10), 512), ((5, 7), 67), ((4, 4), 12288)])
def test_fused_adamw_matches_torch(fcn, shape, cuda_device):
    """Same parameters, moments and next-forward results as torch.optim.AdamW (default weight decay 0.01, the group the
    reference puts som_layer.parameters() in, models/vit_som.py:140-151) over several steps with a changing lr."""
    from vit_som_b200 import FusedPrototypeAdamW, SOMLayer
    ms, D = shape
    B, T = 96, 2.0
    torch.manual_seed(12)
    ours = make_layer(ms, D, fcn, T)
    ref = SOMLayer(make_config(list(ms), D, fcn)).cuda()
    ref.current_temperature = T
    with torch.no_grad():
        ref.prototypes.copy_(ours.prototypes)
    W_start = ours.prototypes.detach().clone()
    opt_o = FusedPrototypeAdamW(ours, lr=2e-3, betas=(0.9, 0.999))
    opt_r = torch.optim.AdamW(ref.parameters(), lr=2e-3, betas=(0.9, 0.999))
    for it in range(5):
        x = torch.randn(B, D, device="cuda", requires_grad=True)
        step(ours, x)
        # the same gradient for both optimizers (two layers that differ by an ulp may pick different BMUs on near-ties)
        ref.prototypes.grad = ours.prototypes.grad.clone()
        for opt in (opt_o, opt_r):
            for g in opt.param_groups:
                g["lr"] = 2e-3 * (1.0 - 0.1 * it)                  # what a scheduler does between steps
            opt.step()
        torch.cuda.synchronize()
        assert O.rel_err(ours.prototypes.detach().cpu().numpy(), ref.prototypes.detach().cpu().numpy()) < 1e-6
        so, sr = opt_o.state[ours.prototypes], opt_r.state[ref.prototypes]
        assert O.rel_err(so["exp_avg"].cpu().numpy(), sr["exp_avg"].cpu().numpy()) < 2e-6
        assert O.rel_err(so["exp_avg_sq"].cpu().numpy(), sr["exp_avg_sq"].cpu().numpy()) < 2e-6
        assert float(so["step"].item()) == it + 1
        # an update really happened, and it is AdamW-sized (about lr per element and step)
        assert 1e-4 < (ours.prototypes.detach() - W_start).abs().max().item() < 0.1
    # the staging written by the optimizer kernel is what the staging kernel produces from the new prototypes (row norms
    # are reduced in a different order by the two kernels: equal to rounding)
    ws = ours._w_cache
    assert ws is not None and ws.from_optimizer and ws.key == ours._staging_key(ours._mode())
    from vit_som_b200 import ops
    fresh = ops.stage_rows(ours.prototypes.detach(), ours._mode())
    torch.cuda.synchronize()
    n = ws.rows * ws.ld
    assert O.rel_err(ws.dense().cpu().numpy(), fresh.dense().cpu().numpy()) < 1e-6
    assert O.rel_err(ws.aux_tensor().cpu().numpy(), fresh.aux_tensor().cpu().numpy()) < 1e-6
    if ops.is_f16(ws.mode):
        # row scales are powers of two that lift the row maximum into [2^14, 2^15)
        sc = ws.scale_tensor()
        assert torch.equal(sc, torch.exp2(torch.log2(sc).round()))
        top = ws.halves()[0].float().abs().amax(dim=1)
        assert float(top.min()) >= 2.0 ** 14 * 0.999 and float(top.max()) <= 2.0 ** 15
    else:
        hi = ws.buf[:n].view(torch.int32)
        assert int((hi & 0x1FFF).abs().max().item()) == 0        # hi parts are exact tf32 values (13 low mantissa bits clear)
    # and the next forward uses it (no W staging in the step) with the same results as a freshly staged layer
    x = torch.randn(B, D, device="cuda")
    ours.eval()
    d_o, b_o = ours(x)
    assert ours._w_cache is ws                                    # not restaged
    ours.invalidate_staging()
    d_f, b_f = ours(x)
    assert O.rel_err(d_o.detach().cpu().numpy(), d_f.detach().cpu().numpy()) < 1e-6
    assert (b_o != b_f).float().mean().item() < 0.05


def test_fused_adamw_state_dict_round_trip(cuda_device):
    from vit_som_b200 import FusedPrototypeAdamW
    layer = make_layer((6, 6), 64, "euclidean", 2.0)
    opt = FusedPrototypeAdamW(layer, lr=1e-3)
    for _ in range(3):
        step(layer, torch.randn(32, 64, device="cuda", requires_grad=True))
        opt.step()
    sd = opt.state_dict()
    layer2 = make_layer((6, 6), 64, "euclidean", 2.0)
    layer2.load_state_dict(layer.state_dict())
    opt2 = FusedPrototypeAdamW(layer2, lr=1e-3)
    opt2.load_state_dict(sd)
    x = torch.randn(32, 64, device="cuda")
    for lay, o in ((layer, opt), (layer2, opt2)):
        step(lay, x.clone().requires_grad_(True))
        o.step()
    torch.cuda.synchronize()
    assert O.rel_err(layer2.prototypes.detach().cpu().numpy(), layer.prototypes.detach().cpu().numpy()) < 1e-6
    assert float(opt2.state[layer2.prototypes]["step"].item()) == 4.0


# ------------------------------------------------------------------------------------------------------------------
# evaluation helpers, ViT-SOM harness
# ------------------------------------------------------------------------------------------------------------------
def test_assign_bmus_and_cluster_metrics(cuda_device):
    from sklearn.metrics import normalized_mutual_info_score
    from vit_som_b200.evaluation import assign_bmus, purity_nmi
    ms, D, N = (8, 8), 128, 5000
    torch.manual_seed(13)
    layer = make_layer(ms, D, "euclidean", 1.0).eval()
    lat = torch.randn(N, D, device="cuda")
    cells = assign_bmus(layer, lat, row_chunk=1536)                 # ragged last chunk
    d, bmu = layer(lat)
    assert torch.equal(cells, bmu)
    labels = (cells % 5 + (torch.rand(N, device="cuda") < 0.2).long()) % 5
    purity, nmi = purity_nmi(cells, labels, 64, 5)
    c, y = cells.cpu().numpy(), labels.cpu().numpy()
    table = np.zeros((64, 5), np.int64)
    np.add.at(table, (c, y), 1)
    assert abs(purity.item() - table.max(1).sum() / N) < 1e-12
    assert abs(nmi.item() - normalized_mutual_info_score(y, c)) < 1e-9


@pytest.mark.parametrize("classes", [10, 0])
def test_vit_som_training_step(classes, cuda_device):
    """The harness behind the img/s records: ViT autoencoder (bf16 autocast) -> strided SOM input -> SOM loss with the
    device-side gamma ramp -> backward -> AdamW (ViT) + fused AdamW (prototypes).  The SOM part is checked against the
    oracle on the latents the ViT produced; the loss must go down over a few steps of the same batch."""
    from vit_som_b200.vit_som import ViTSOM, build_optimizers, reference_yaml_config
    cfg = reference_yaml_config("cifar-10", (4, 4), 32)
    cfg["hyperparameters"]["vit"].update(depth=2, dec_depth=1)
    cfg["data"]["num_classes"] = classes
    torch.manual_seed(14)
    model = ViTSOM(cfg).cuda().train()
    model.som_layer.total_iterations = 1000
    model.ramp_up_end_step = 4
    opt_vit, opt_som = build_optimizers(model)
    img = torch.randn(32, 3, 32, 32, device="cuda")
    labels = torch.randint(0, max(classes, 1), (32,), device="cuda")
    # SOM parity on the real latents
    with torch.no_grad():
        _, _, _, d, bmu = model(img)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cls, patches, _ = model.vit(img)
    lat = patches.flatten(1).float().cpu().numpy()
    W = model.som_layer.prototypes.detach().cpu().numpy()
    _, hard, worst = O.classify_bmu_mismatches(lat, W, bmu.cpu().numpy(), "cosine")
    assert hard == 0, worst
    losses = []
    for it in range(6):
        total, som = model.training_loss(img, labels)
        opt_vit.zero_grad(set_to_none=True)
        opt_som.zero_grad(set_to_none=True)
        total.backward()
        assert model.som_layer.prototypes.grad is not None
        if it >= 1:                                                # gamma(0) = 0: no SOM gradient reaches the ViT at step 0
            assert model.vit.blocks[0].qkv.weight.grad is not None
        opt_vit.step()
        opt_som.step()
        losses.append(total.item())
        assert np.isfinite(losses[-1])
    assert int(model.iteration.item()) == 6
    assert torch.is_tensor(model.som_layer.current_temperature) and model.som_layer.current_temperature.is_cuda
    assert losses[-1] < losses[0] + 0.5 * abs(losses[0])           # training does not diverge on a fixed batch
