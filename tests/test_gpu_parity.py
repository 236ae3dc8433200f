"""GPU parity tests (run on a B200: ``python -m pytest tests -m gpu``).

The CUDA path (``vit_som_b200.SOMLayer`` -> ctypes -> libsom_b200.so) is compared with
  * the committed golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  * the CPU oracle (oracle/som_oracle.py) on seeded inputs at BASELINE.json's shapes.
Tolerances (BASELINE.json north_star): BMU indices bit-exact except documented fp32 near-ties (classified
with the fp64 oracle), loss and gradients within 1e-5 relative.
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import som_oracle as O
from oracle.ref_import import make_config

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("som_precision")]

LOSS_TOL = 1e-5     # relative, north_star
GRAD_TOL = 1e-5     # norm-wise relative, north_star
DIST_TOL = 3e-6     # norm-wise relative on the distance matrix (fp32 rounding of O(|x|^2) terms)

FULL = [n for n in golden_names() if not n.startswith("cfg1_")]


def make_layer(map_size, D, fcn, topology="square", W=None, T=None):
    from vit_som_b200 import SOMLayer
    layer = SOMLayer(make_config(list(map_size), D, fcn, topology=topology)).cuda()
    if W is not None:
        with torch.no_grad():
            layer.prototypes.copy_(torch.as_tensor(W))
    if T is not None:
        layer.current_temperature = T
    return layer


def run_layer(layer, x_np, g_out=1.0):
    x = torch.as_tensor(x_np).cuda().requires_grad_(True)
    layer.prototypes.grad = None
    d, bmu = layer(x)
    w = layer.compute_weights(bmu)
    loss = layer.som_loss(w, d)
    (loss * g_out).backward()
    torch.cuda.synchronize()
    return dict(distances=d.detach().cpu().numpy(), bmu=bmu.cpu().numpy(), loss=loss.item(),
                grad_x=x.grad.cpu().numpy(), grad_w=layer.prototypes.grad.cpu().numpy(), weights_handle=w)


def check_against(out, ref_d, ref_bmu, ref_loss, ref_gx, ref_gw, x2d, W, fcn, pos, T, g_out, exact=False):
    """Compare a GPU result with a reference result.  When BMUs differ on fp32 near-ties the loss/grad
    reference is recomputed by the fp64 oracle with the GPU's BMUs (the tie changes the neighbourhood)."""
    n_bad = int((out["bmu"] != ref_bmu).sum())
    if n_bad:
        _, hard, worst = O.classify_bmu_mismatches(x2d, W, out["bmu"], fcn)
        assert hard == 0, f"{n_bad} BMU mismatches, worst relative gap {worst:.3e} is not an fp32 near-tie"
        r = O.step(x2d, W, pos, T, fcn, g_out, np.float64, bmu_override=out["bmu"])
        ref_d, ref_loss, ref_gx, ref_gw = r.distances, float(r.loss), r.grad_x, r.grad_w
    if exact:
        np.testing.assert_array_equal(out["distances"], ref_d)
    assert O.rel_err(out["distances"], ref_d) < DIST_TOL
    assert abs(out["loss"] - ref_loss) <= LOSS_TOL * abs(ref_loss) + 1e-30
    assert O.rel_err(out["grad_x"].reshape(ref_gx.shape), ref_gx) < GRAD_TOL
    assert O.rel_err(out["grad_w"], ref_gw) < GRAD_TOL
    return n_bad


@pytest.mark.parametrize("name", FULL)
def test_golden_vectors(name, cuda_device):
    fx = load_golden(name)
    fcn, topo = str(fx["distance_fcn"]), str(fx["topology"])
    T = float(fx["T"])
    x2d = fx["x"].reshape(fx["x"].shape[0], -1)
    layer = make_layer(fx["map_size"], x2d.shape[1], fcn, topo, W=fx["W"], T=T)
    out = run_layer(layer, fx["x"], float(fx["g_out"]))
    assert out["distances"].shape == fx["distances"].shape and out["bmu"].dtype == np.int64
    if name.startswith("edge"):
        # exact-arithmetic / degenerate inputs: duplicates -> lowest index, x == W_k -> d == 0 -> masked gradient
        np.testing.assert_array_equal(out["bmu"], fx["bmu"])
    if name == "edge_int_euclidean":
        np.testing.assert_array_equal(out["distances"], fx["distances"])
        assert out["distances"][0, 3] == 0.0 and out["bmu"][0] == 3
    if name == "tiny_euclidean":
        # ATen's direct per-pair kernel (B,K <= 25) rounds differently from the expansion: compare in fp64 terms
        ref = O.step(x2d, fx["W"], fx["grid_positions"], T, fcn, float(fx["g_out"]), np.float64, bmu_override=fx["bmu"])
        check_against(out, ref.distances, fx["bmu"], float(ref.loss), ref.grad_x, ref.grad_w, x2d, fx["W"], fcn,
                      fx["grid_positions"], T, float(fx["g_out"]))
        return
    if name == "edge_cosine":
        # row 5 is the zero vector: gradient magnitudes ~1e12 * tiny; compare the well-conditioned rows
        keep = np.ones(x2d.shape[0], bool); keep[5] = False
        assert np.all(np.isfinite(out["grad_x"]))
        assert O.rel_err(out["grad_x"][keep], fx["grad_x"][keep]) < GRAD_TOL
        assert abs(out["loss"] - float(fx["loss"])) <= LOSS_TOL * abs(float(fx["loss"]))
        assert O.rel_err(out["distances"], fx["distances"]) < DIST_TOL
        return
    check_against(out, fx["distances"], fx["bmu"], float(fx["loss"]), fx["grad_x"].reshape(x2d.shape), fx["grad_w"],
                  x2d, fx["W"], fcn, fx["grid_positions"], T, float(fx["g_out"]))
    # the lazily materialised weights equal the reference's compute_weights
    if (out["bmu"] == fx["bmu"]).all():
        w = out["weights_handle"].materialize().cpu().numpy()
        np.testing.assert_allclose(w, fx["weights"], rtol=5e-6, atol=1e-37)


SHAPES = {
    # BASELINE.json configs (B, map, D) — cfg4/cfg5 are exercised at reduced B in test_gpu_large.py
    "cfg1": (256, (24, 24), 3136, 12.0),
    "cfg2": (1024, (40, 40), 3136, 20.0),
    "cfg3": (128, (4, 4), 12288, 4.0),
}


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
@pytest.mark.parametrize("cfg", list(SHAPES))
def test_baseline_shapes_vs_oracle(cfg, fcn, cuda_device):
    B, ms, D, T = SHAPES[cfg]
    torch.manual_seed(0)
    layer = make_layer(ms, D, fcn, T=T)
    x = torch.randn(B, D)
    W = layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions(ms)
    out = run_layer(layer, x.numpy())
    ref = O.step(x.numpy(), W, pos, T, fcn, 1.0, np.float32)
    n_bad = check_against(out, ref.distances, ref.bmu, float(ref.loss), ref.grad_x, ref.grad_w, x.numpy(), W, fcn,
                          pos, T, 1.0)
    # fp64 cross-check of the gradients (the fp32 oracle itself carries ~1e-6 error)
    r64 = O.step(x.numpy(), W, pos, T, fcn, 1.0, np.float64, bmu_override=out["bmu"])
    assert O.rel_err(out["grad_x"], r64.grad_x) < GRAD_TOL
    assert O.rel_err(out["grad_w"], r64.grad_w) < GRAD_TOL
    assert abs(out["loss"] - float(r64.loss)) <= LOSS_TOL * abs(float(r64.loss))
    print(f"{cfg}/{fcn}: BMU near-tie mismatches {n_bad}/{B}")


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
@pytest.mark.parametrize("T", [20.0, 0.1414, 1e-3])
def test_temperature_extremes(fcn, T, cuda_device):
    """T = Tmax / geometric mid / Tmin of the 40x40 YAML: weights underflow to exactly 0 away from the BMU."""
    torch.manual_seed(1)
    layer = make_layer((16, 16), 512, fcn, T=T)
    x = torch.randn(300, 512).numpy()
    W = layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions((16, 16))
    out = run_layer(layer, x, g_out=0.5)
    ref = O.step(x, W, pos, T, fcn, 0.5, np.float64, bmu_override=out["bmu"])
    _, hard, _ = O.classify_bmu_mismatches(x, W, out["bmu"], fcn)
    assert hard == 0
    assert abs(out["loss"] - float(ref.loss)) <= LOSS_TOL * abs(float(ref.loss))
    assert O.rel_err(out["grad_x"], ref.grad_x) < GRAD_TOL
    assert O.rel_err(out["grad_w"], ref.grad_w) < GRAD_TOL
    if T == 1e-3:   # only the BMU cell survives: loss = mean of the row minima / K
        dmin = out["distances"].min(1).astype(np.float64)
        assert abs(out["loss"] - dmin.sum() / out["distances"].size) <= 1e-6 * abs(out["loss"])


def test_device_temperature_schedule_no_sync(cuda_device):
    """update_temperature with a 0-dim int64 device tensor keeps T on the device (vit_som.py:65,84)."""
    layer = make_layer((8, 8), 96, "euclidean")
    layer.total_iterations = (160 / 16) * 3
    layer.Tmax, layer.Tmin = 20.0, 1e-3
    it = torch.tensor(17, device="cuda")
    layer.update_temperature(it)
    assert torch.is_tensor(layer.current_temperature) and layer.current_temperature.is_cuda
    fx = load_golden("sched_euclidean")
    assert abs(layer.current_temperature.item() - float(fx["T"])) <= 2e-6 * float(fx["T"])
    with torch.no_grad():
        layer.prototypes.copy_(torch.as_tensor(fx["W"]))
    out = run_layer(layer, fx["x"])
    assert abs(out["loss"] - float(fx["loss"])) <= LOSS_TOL * abs(float(fx["loss"]))
    assert O.rel_err(out["grad_w"], fx["grad_w"]) < GRAD_TOL


@pytest.mark.parametrize("shape", [(1, (3, 3), 5), (7, (5, 7), 67), (130, (9, 11), 33), (129, (1, 17), 129)])
@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_ragged_shapes(shape, fcn, cuda_device):
    """Sizes that are not multiples of any tile dimension (rows, prototypes, latent dim, odd leading dims)."""
    B, ms, D = shape
    torch.manual_seed(2)
    layer = make_layer(ms, D, fcn, T=1.3)
    x = torch.randn(B, D).numpy()
    W = layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions(ms)
    out = run_layer(layer, x)
    ref = O.step(x, W, pos, 1.3, fcn, 1.0, np.float64, bmu_override=out["bmu"])
    _, hard, _ = O.classify_bmu_mismatches(x, W, out["bmu"], fcn)
    assert hard == 0
    assert O.rel_err(out["distances"], ref.distances) < DIST_TOL
    assert abs(out["loss"] - float(ref.loss)) <= LOSS_TOL * abs(float(ref.loss))
    assert O.rel_err(out["grad_x"], ref.grad_x) < GRAD_TOL
    assert O.rel_err(out["grad_w"], ref.grad_w) < GRAD_TOL


def test_module_interface(cuda_device):
    """Attributes, state-dict keys, eval path, flatten, bf16 latents, plain-tensor weights, error behaviour."""
    from vit_som_b200 import SOMLayer, SomError
    layer = make_layer((6, 6), 6 * 20, "euclidean", T=2.0)
    assert list(layer.state_dict().keys()) == ["prototypes", "grid_positions"]
    assert layer.n_prototypes == 36 and layer.latent_dim == 120 and layer.map_size == [6, 6]
    x3 = torch.randn(10, 6, 20, device="cuda")
    d, b = layer(x3)                                             # flatten (som_layer.py:84-85)
    assert d.shape == (10, 36) and b.shape == (10,) and b.dtype == torch.int64
    assert torch.equal(b, layer.best_matching_units(x3))         # argmin-only inference path
    assert torch.equal(b, torch.argmin(d, dim=1))                # first-index semantics on our own distances
    w = layer.compute_weights(b)
    dense = w.materialize()
    l_fused = layer.som_loss(layer.compute_weights(b), d)
    l_plain = layer.som_loss(dense, d)                           # caller-provided dense weights still work
    assert abs(l_fused.item() - l_plain.item()) <= 2e-6 * abs(l_plain.item())
    assert torch.allclose((w * d).mean(), l_plain)               # lazy handle composes with torch ops
    assert torch.equal(layer.index_to_position(torch.tensor([10]))[0], torch.tensor([1.0, 4.0]))
    # bf16 latents are widened exactly
    xb = torch.randn(10, 120, device="cuda").bfloat16()
    d_b, _ = layer(xb)
    d_f, _ = layer(xb.float())
    assert torch.equal(d_b, d_f)
    with pytest.raises(SomError):
        layer(torch.randn(4, 120))                               # CPU tensor: loud failure, no fallback
    with pytest.raises(ValueError):
        SOMLayer(make_config([4, 4], 8, "euclidean", topology="triangle"))
    with pytest.raises(ValueError):
        SOMLayer(make_config([4, 4], 8, "chebyshev"))
    with pytest.raises(NotImplementedError):
        make_layer((4, 4), 8, "manhattan")(torch.randn(2, 8, device="cuda"))


def test_prototype_cache_tracks_updates(cuda_device):
    """The staged (tf32-split) prototypes are refreshed after an optimizer step, not before."""
    layer = make_layer((5, 5), 64, "euclidean", T=1.0)
    opt = torch.optim.SGD(layer.parameters(), lr=0.5)
    x = torch.randn(40, 64, device="cuda")
    d0, b0 = layer(x)
    layer.som_loss(layer.compute_weights(b0), d0).backward()
    opt.step()
    d1, _ = layer(x)
    ref = O.distances(x.cpu().numpy(), layer.prototypes.detach().cpu().numpy(), "euclidean", np.float64)
    assert O.rel_err(d1.detach().cpu().numpy(), ref) < DIST_TOL
    assert not torch.equal(d0, d1)


def test_gradient_linearity_and_determinism(cuda_device):
    """Size-independent properties: gradients scale linearly with the upstream gradient; loss is run-to-run
    bit-identical (fixed-order reduction); BMUs are invariant under a permutation of the prototypes."""
    torch.manual_seed(4)
    layer = make_layer((20, 20), 768, "euclidean", T=5.0)
    x = torch.randn(512, 768).numpy()
    a = run_layer(layer, x, 1.0)
    b = run_layer(layer, x, 1.0)
    c = run_layer(layer, x, 3.0)
    assert a["loss"] == b["loss"] and np.array_equal(a["bmu"], b["bmu"])
    assert O.rel_err(c["grad_x"], 3.0 * a["grad_x"]) < 2e-6
    assert O.rel_err(c["grad_w"], 3.0 * a["grad_w"]) < 2e-6
    perm = np.random.RandomState(0).permutation(400)
    layer2 = make_layer((20, 20), 768, "euclidean", W=layer.prototypes.detach().cpu().numpy()[perm], T=5.0)
    b2 = layer2.best_matching_units(torch.as_tensor(x).cuda()).cpu().numpy()
    assert np.array_equal(perm[b2], a["bmu"])


@pytest.mark.parametrize("shape", [(1024, (40, 40), 3136), (300, (17, 19), 520), (2048, (32, 32), 256), (129, (20, 13), 67)])
@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_fused_backward_matches_separate_launches(shape, fcn, cuda_device):
    """som_backward_fused (both gradient GEMMs in one stream-K launch) against som_backward_dw + som_backward_dx and
    the fp64 oracle."""
    from vit_som_b200 import ops
    B, ms, D = shape
    torch.manual_seed(11)
    layer = make_layer(ms, D, fcn, T=3.0)
    x = torch.randn(B, D).numpy()
    assert ops.FUSE_BACKWARD
    a = run_layer(layer, x, 0.7)
    b = run_layer(layer, x, 0.7)
    ops.FUSE_BACKWARD = False
    try:
        c = run_layer(layer, x, 0.7)
    finally:
        ops.FUSE_BACKWARD = True
    # (row / column sums of the loss kernel are float atomics: run-to-run equal to rounding, not bitwise)
    assert O.rel_err(a["grad_x"], b["grad_x"]) < 5e-7 and O.rel_err(a["grad_w"], b["grad_w"]) < 5e-7
    assert O.rel_err(a["grad_x"], c["grad_x"]) < 2e-6
    assert O.rel_err(a["grad_w"], c["grad_w"]) < 2e-6
    W = layer.prototypes.detach().cpu().numpy()
    r64 = O.step(x, W, O.grid_positions(ms), 3.0, fcn, 0.7, np.float64, bmu_override=a["bmu"])
    assert O.rel_err(a["grad_x"], r64.grad_x) < GRAD_TOL
    assert O.rel_err(a["grad_w"], r64.grad_w) < GRAD_TOL


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_streamk_forced_matches_oracle(fcn, cuda_device):
    """Every CTA-pair GEMM forced onto the stream-K schedule (tiles cut across pairs, partial tiles handed over
    through the workspace): same BMUs / loss / gradients."""
    from vit_som_b200 import _lib
    B, ms, D, T = 1024, (24, 24), 1000, 6.0
    torch.manual_seed(5)
    layer = make_layer(ms, D, fcn, T=T)
    x = torch.randn(B, D).numpy()
    W = layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions(ms)
    L = _lib.lib()
    L.som_set_streamk(1)
    L.som_set_cta_group(2)
    try:
        out = run_layer(layer, x)
        again = run_layer(layer, x)
    finally:
        L.som_set_streamk(0)
        L.som_set_cta_group(0)
    ref = O.step(x, W, pos, T, fcn, 1.0, np.float32)
    check_against(out, ref.distances, ref.bmu, float(ref.loss), ref.grad_x, ref.grad_w, x, W, fcn, pos, T, 1.0)
    assert out["loss"] == again["loss"] and np.array_equal(out["distances"], again["distances"])
    assert O.rel_err(out["grad_w"], again["grad_w"]) < 5e-7


@pytest.mark.parametrize("shape", [(20, (10, 16), 64), (40, (8, 20), 96), (160, (12, 16), 32), (33, (16, 16), 160)])
def test_pair_kernel_with_short_reductions(shape, cuda_device):
    """CTA-pair kernel forced on shapes whose reduction dimension is shorter than one k-block or one TMA box
    (MN-major operands through 3-D tensor maps with k < 32, zero-filled out-of-range panels and k rows)."""
    from vit_som_b200 import _lib
    B, ms, D = shape
    torch.manual_seed(3)
    layer = make_layer(ms, D, "euclidean", T=2.0)
    x = torch.randn(B, D).numpy()
    W = layer.prototypes.detach().cpu().numpy()
    pos = O.grid_positions(ms)
    L = _lib.lib()
    L.som_set_cta_group(2)
    try:
        out = run_layer(layer, x)
    finally:
        L.som_set_cta_group(0)
    ref = O.step(x, W, pos, 2.0, "euclidean", 1.0, np.float32)
    check_against(out, ref.distances, ref.bmu, float(ref.loss), ref.grad_x, ref.grad_w, x, W, "euclidean", pos, 2.0, 1.0)
