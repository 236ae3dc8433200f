"""Pin the CPU oracle (oracle/som_oracle.py) against the golden vectors produced by the unmodified
reference SOMLayer (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import som_oracle as O

FULL = [n for n in golden_names() if not n.startswith("cfg1_")]


def _T(fx):
    # the reference's temperature was either a python float or a 0-dim fp32 tensor (sched_* fixtures)
    return np.float32(fx["T"]) if bool(fx["T_is_tensor"]) else float(fx["T"])


@pytest.mark.parametrize("name", FULL)
def test_oracle_fp32_matches_reference(name):
    fx = load_golden(name)
    fcn, topo = str(fx["distance_fcn"]), str(fx["topology"])
    pos = O.grid_positions(fx["map_size"], topo)
    np.testing.assert_array_equal(pos, fx["grid_positions"])
    x = fx["x"].reshape(fx["x"].shape[0], -1)
    res = O.step(fx["x"], fx["W"], pos, _T(fx), fcn, float(fx["g_out"]), np.float32)
    # distances: same formula, BLAS accumulation order may differ from ATen's by a few ulp
    assert O.rel_err(res.distances, fx["distances"]) < 2e-6
    if fcn == "euclidean" and name.startswith("edge_int"):
        np.testing.assert_array_equal(res.distances, fx["distances"])     # exact arithmetic case
    np.testing.assert_array_equal(res.bmu, fx["bmu"])
    # weights from the reference's own BMUs: exp() implementations may differ in the last ulp
    np.testing.assert_allclose(res.weights, fx["weights"], rtol=2e-6, atol=1e-37)
    assert abs(float(res.loss) - float(fx["loss"])) <= 2e-6 * abs(float(fx["loss"])) + 1e-12
    assert O.rel_err(res.grad_x, fx["grad_x"].reshape(res.grad_x.shape)) < 5e-6
    assert O.rel_err(res.grad_w, fx["grad_w"]) < 5e-6
    assert x.shape[1] == fx["W"].shape[1]


@pytest.mark.parametrize("name", FULL)
def test_oracle_fp64_agrees_with_reference(name):
    """The fp64 restatement is the arbiter for near-ties: it must agree with the fp32 reference to fp32 accuracy."""
    fx = load_golden(name)
    fcn = str(fx["distance_fcn"])
    pos = O.grid_positions(fx["map_size"], str(fx["topology"]))
    res = O.step(fx["x"], fx["W"], pos, float(fx["T"]), fcn, float(fx["g_out"]), np.float64, bmu_override=fx["bmu"])
    # near-zero distances (x == W_k) amplify rounding through the sqrt: compare away from them
    far = fx["distances"] > 1e-2
    np.testing.assert_allclose(res.distances[far], fx["distances"][far], rtol=3e-5, atol=3e-6)
    assert abs(float(res.loss) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"])) + 1e-12
    n_bad, hard, worst = O.classify_bmu_mismatches(fx["x"].reshape(len(fx["bmu"]), -1), fx["W"], fx["bmu"], fcn)
    assert hard == 0, (n_bad, worst)
    if not name.startswith("edge"):
        assert O.rel_err(res.grad_x, fx["grad_x"].reshape(res.grad_x.shape)) < 1e-5
        assert O.rel_err(res.grad_w, fx["grad_w"]) < 1e-5


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_oracle_cfg1_shape(fcn):
    """BASELINE config 1 shape (24x24, D=3136, B=256): inputs regenerated from seeds, outputs from the reference."""
    torch = pytest.importorskip("torch")
    fx = load_golden(f"cfg1_{fcn}")
    torch.manual_seed(int(fx["x_seed"]))
    x = torch.randn(256, 3136).numpy()
    torch.manual_seed(int(fx["ctor_seed"]))
    W = torch.rand(576, 3136)
    if fcn == "cosine":
        W = torch.nn.functional.normalize(W, p=2, dim=1)
    W = W.numpy()
    if abs(np.abs(x).astype(np.float64).sum() - float(fx["x_checksum"])) > 1e-6 * float(fx["x_checksum"]):
        pytest.skip("torch RNG stream differs from the one that generated the fixture")
    assert abs(np.abs(W).astype(np.float64).sum() - float(fx["W_checksum"])) < 1e-6 * float(fx["W_checksum"])
    pos = O.grid_positions(fx["map_size"], "square")
    res = O.step(x, W, pos, float(fx["T"]), fcn, 1.0, np.float32)
    n_bad = int((res.bmu != fx["bmu"]).sum())
    if n_bad:
        _, hard, worst = O.classify_bmu_mismatches(x, W, fx["bmu"], fcn)
        assert hard == 0, (n_bad, worst)
    np.testing.assert_allclose(res.distances.min(1), fx["dist_row_min"], rtol=2e-6)
    np.testing.assert_allclose(res.distances.astype(np.float64).sum(1), fx["dist_row_sum"], rtol=1e-6)
    assert abs(float(res.loss) - float(fx["loss"])) <= 2e-6 * abs(float(fx["loss"]))
    assert abs(np.linalg.norm(res.grad_x.astype(np.float64)) - float(fx["grad_x_norm"])) < 5e-6 * float(fx["grad_x_norm"])
    assert abs(np.linalg.norm(res.grad_w.astype(np.float64)) - float(fx["grad_w_norm"])) < 5e-6 * float(fx["grad_w_norm"])
    np.testing.assert_allclose(res.grad_x[::37, ::211], fx["grad_x_sample"], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(res.grad_w[::53, ::197], fx["grad_w_sample"], rtol=2e-4, atol=1e-9)


def test_temperature_schedule_matches_reference():
    for fcn in ("euclidean", "cosine"):
        fx = load_golden(f"sched_{fcn}")
        tmax, tmin = fx["Tmax_Tmin"]
        # reference: python floats ** (int64 0-dim tensor / python float) -> fp32 tensor
        T = O.temperature(np.asarray(int(fx["iteration"])), float(tmax), float(tmin), float(fx["total_iterations"]))
        assert bool(fx["T_is_tensor"])
        assert abs(float(T) - float(fx["T"])) <= 2e-6 * float(fx["T"])
        T_py = O.temperature(int(fx["iteration"]), float(tmax), float(tmin), float(fx["total_iterations"]))
        assert abs(T_py - float(fx["T"])) <= 2e-6 * float(fx["T"])


def test_index_to_position_known_answer():
    """The only known-answer assertion in the reference's tests (experiments/tests/unit_test.py:11-16)."""
    np.testing.assert_array_equal(O.index_to_position(np.asarray([10]), (8, 8)), [[1.0, 2.0]])


def test_micro_vector_matches_survey_appendix_b():
    fx = load_golden("micro_euclidean")
    assert fx["bmu"].tolist() == [9, 32, 1, 3, 13, 34, 35, 3, 22, 34, 3, 34, 17, 7, 35, 34, 24, 7, 3, 7, 16, 34, 23,
                                  16, 16, 34, 35, 1, 16, 34, 0, 35]
    assert abs(float(fx["loss"]) - 1.519182324) < 1e-6
    fx = load_golden("micro_cosine")
    assert abs(float(fx["loss"]) - 0.243644103) < 1e-6


@pytest.mark.parametrize("name", FULL)
def test_torch_port_matches_reference(name):
    """oracle/som_torch_ref.py (the timed CPU baseline) issues the same ATen calls as the reference: bit-level match."""
    torch = pytest.importorskip("torch")
    from oracle import som_torch_ref as R
    fx = load_golden(name)
    T = torch.tensor(float(fx["T"]), dtype=torch.float32) if bool(fx["T_is_tensor"]) else float(fx["T"])
    d, b, loss, gx, gw = R.step(torch.as_tensor(fx["x"]), torch.as_tensor(fx["W"]), torch.as_tensor(fx["grid_positions"]),
                                T, str(fx["distance_fcn"]), float(fx["g_out"]))
    np.testing.assert_array_equal(b.numpy(), fx["bmu"])
    np.testing.assert_allclose(d.numpy(), fx["distances"], rtol=1e-6, atol=1e-7)
    assert abs(loss.item() - float(fx["loss"])) <= 1e-6 * abs(float(fx["loss"]))
    assert O.rel_err(gx.numpy(), fx["grad_x"]) < 1e-6
    assert O.rel_err(gw.numpy(), fx["grad_w"]) < 1e-6
