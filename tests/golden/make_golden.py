"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference SOMLayer.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Every fixture is produced by importing /root/reference/models/som_layer.py (behind the
pytorch_lightning stub of oracle/ref_import.py), CPU fp32, torch as installed in this image, by the
call sequence of models/vit_som.py:82-86:  forward -> [temperature] -> compute_weights -> som_loss ->
backward.  Inputs are stored next to the outputs so the fixtures do not depend on torch's RNG streams.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference_som_layer, make_config  # noqa: E402


def run_reference(SOMLayer, cfg, x, W=None, T=None, g_out=1.0, seed=0, iteration=None, n_samples=None):
    torch.manual_seed(seed)
    layer = SOMLayer(cfg)
    if W is not None:
        with torch.no_grad():
            layer.prototypes.copy_(torch.as_tensor(W))
    x = torch.as_tensor(x).clone().requires_grad_(True)
    d, b = layer(x)
    if iteration is not None:
        class _DL:                       # what update_temperature dereferences (models/som_layer.py:131)
            dataset = range(n_samples)
        class _Tr:
            train_dataloader = _DL()
        layer.trainer = _Tr()
        layer.update_temperature(iteration)
    elif T is not None:
        layer.current_temperature = T
    w = layer.compute_weights(b)
    loss = layer.som_loss(w, d)
    (loss * g_out).backward()
    Tcur = layer.current_temperature
    out = {
        "x": x.detach().numpy(), "W": layer.prototypes.detach().numpy().copy(),
        "grid_positions": layer.grid_positions.numpy().copy(),
        "distances": d.detach().numpy(), "bmu": b.numpy(), "weights": w.detach().numpy(),
        "loss": np.float32(loss.item()), "grad_x": x.grad.numpy(), "grad_w": layer.prototypes.grad.numpy(),
        "T": np.asarray(Tcur.item() if torch.is_tensor(Tcur) else Tcur, dtype=np.float64),
        "T_is_tensor": np.asarray(bool(torch.is_tensor(Tcur))),
        "g_out": np.float64(g_out),
        "map_size": np.asarray(cfg["hyperparameters"]["som"]["map_size"]),
        "distance_fcn": np.asarray(cfg["hyperparameters"]["som"]["distance_fcn"]),
        "topology": np.asarray(cfg["hyperparameters"]["som"]["topology"]),
    }
    return out


def main():
    SOMLayer = load_reference_som_layer()
    g = torch.Generator().manual_seed(1234)
    fixtures = {}

    for fcn in ("euclidean", "cosine"):
        # 1. the micro known-answer vector of SURVEY.md Appendix B (RNG order matters: ctor then randn)
        torch.manual_seed(0)
        cfg = make_config([6, 6], 32, fcn, Tmax=1.5, Tmin=0.1)
        layer = SOMLayer(cfg)
        x = torch.randn(32, 32)
        fixtures[f"micro_{fcn}"] = run_reference(SOMLayer, cfg, x.numpy(), W=layer.prototypes.detach().numpy())

        # 2. 10x10 square map, D=200, B=40 at three temperatures and a non-unit upstream gradient
        cfg = make_config([10, 10], 200, fcn, Tmax=10.0, Tmin=0.1)
        x = torch.randn(40, 200, generator=g).numpy()
        for tag, T, go in (("Tmax", 10.0, 1.0), ("Tmid", 1.0, 0.37), ("Tmin", 0.1, 1.0), ("Ttiny", 1e-3, 2.5)):
            fixtures[f"sq10_{tag}_{fcn}"] = run_reference(SOMLayer, cfg, x, T=T, g_out=go, seed=5)

        # 3. hexagonal topology, odd sizes
        cfg = make_config([5, 7], 67, fcn, topology="hexa", Tmax=3.0, Tmin=0.1)
        x = torch.randn(33, 67, generator=g).numpy()
        fixtures[f"hexa_{fcn}"] = run_reference(SOMLayer, cfg, x, T=1.7, seed=6)

        # 4. B <= 25 and K <= 25: ATen's cdist switches to the direct per-pair kernel
        cfg = make_config([4, 4], 48, fcn, Tmax=4.0, Tmin=0.1)
        x = torch.randn(16, 48, generator=g).numpy()
        fixtures[f"tiny_{fcn}"] = run_reference(SOMLayer, cfg, x, T=4.0, seed=7)

        # 5. temperature schedule through update_temperature with a 0-dim int64 tensor iteration (vit_som.py:65,84)
        cfg = make_config([8, 8], 96, fcn, Tmax=20.0, Tmin=1e-3, total_epochs=3, batch_size=16)
        x = torch.randn(24, 96, generator=g).numpy()
        fixtures[f"sched_{fcn}"] = run_reference(SOMLayer, cfg, x, seed=8, iteration=torch.tensor(17), n_samples=160)
        fixtures[f"sched_{fcn}"]["iteration"] = np.asarray(17)
        fixtures[f"sched_{fcn}"]["total_iterations"] = np.float64((160 / 16) * 3)
        fixtures[f"sched_{fcn}"]["Tmax_Tmin"] = np.asarray([20.0, 1e-3])

        # 6. 3-D input is flattened (models/som_layer.py:84-85)
        cfg = make_config([7, 9], 6 * 20, fcn, Tmax=5.0, Tmin=0.1)
        x = torch.randn(30, 6, 20, generator=g).numpy()
        fixtures[f"flatten3d_{fcn}"] = run_reference(SOMLayer, cfg, x, T=2.0, seed=9)

    # 7. exact-arithmetic edge cases (small integers: every product and sum is exact in fp32 and tf32)
    cfg = make_config([6, 6], 40, "euclidean", Tmax=2.0, Tmin=0.1)
    W = torch.randint(0, 3, (36, 40), generator=g).float()
    W[7] = W[3]; W[20] = W[3]                                  # duplicate prototypes -> tie -> lowest index
    x = torch.randint(0, 3, (30, 40), generator=g).float()
    x[0] = W[3]; x[1] = W[11]; x[2] = W[35]                    # x == W_k -> d == 0 -> masked gradient
    fixtures["edge_int_euclidean"] = run_reference(SOMLayer, cfg, x.numpy(), W=W.numpy(), T=1.2, seed=10)
    cfg = make_config([6, 6], 40, "cosine", Tmax=2.0, Tmin=0.1)
    Wc = torch.rand(36, 40, generator=g)
    Wc[9] = Wc[2]
    xc = torch.randn(28, 40, generator=g)
    xc[5] = 0.0                                                # zero vector under cosine: eps clamp
    xc[6] = 3.0 * Wc[2]                                        # parallel to a duplicated prototype
    fixtures["edge_cosine"] = run_reference(SOMLayer, cfg, xc.numpy(), W=Wc.numpy(), T=1.2, seed=11)

    # 8. BASELINE config 1 shape (24x24 map, D=3136, B=256): outputs only, inputs from seeds
    for fcn in ("euclidean", "cosine"):
        cfg = make_config([24, 24], 3136, fcn, Tmax=12.0, Tmin=0.1)
        torch.manual_seed(21)
        x = torch.randn(256, 3136)
        r = run_reference(SOMLayer, cfg, x.numpy(), T=12.0, seed=22)
        d = r["distances"]
        fixtures[f"cfg1_{fcn}"] = {
            "x_seed": np.asarray(21), "ctor_seed": np.asarray(22),
            "x_checksum": np.float64(np.abs(r["x"]).astype(np.float64).sum()),
            "W_checksum": np.float64(np.abs(r["W"]).astype(np.float64).sum()),
            "bmu": r["bmu"], "loss": r["loss"], "T": r["T"], "g_out": r["g_out"], "T_is_tensor": r["T_is_tensor"],
            "map_size": r["map_size"], "distance_fcn": r["distance_fcn"], "topology": r["topology"],
            "dist_row_min": d.min(1), "dist_row_sum": d.astype(np.float64).sum(1),
            "grad_x_norm": np.float64(np.linalg.norm(r["grad_x"].astype(np.float64))),
            "grad_w_norm": np.float64(np.linalg.norm(r["grad_w"].astype(np.float64))),
            "grad_x_sample": r["grad_x"][::37, ::211].copy(), "grad_w_sample": r["grad_w"][::53, ::197].copy(),
        }

    for name, fx in fixtures.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **fx)
    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"wrote {len(fixtures)} fixtures, {total / 1e6:.2f} MB; torch {torch.__version__}")


if __name__ == "__main__":
    main()
