"""Import the UNMODIFIED reference SOMLayer from /root/reference  —  TEST INFRASTRUCTURE ONLY.

Used in the build container to pin the oracle (``tests/golden/make_golden.py``) and by
``bench.py --impl reference`` when the reference tree happens to be present.  /root/reference does
not exist on the GPU box; callers must handle ``reference_available() == False``.

``models/som_layer.py:5`` imports ``pytorch_lightning`` only for its base class, and Lightning is not
installed in this image, so a stub module (LightningModule = nn.Module + a no-op ``log``) is placed
in ``sys.modules`` before importing the file.  No reference source is copied into this repository.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SOM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "som_layer.py"))


def _install_lightning_stub() -> None:
    if "pytorch_lightning" in sys.modules:
        return
    import torch.nn as nn

    class LightningModule(nn.Module):
        trainer = None

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

        def save_hyperparameters(self, *a, **k):
            pass

    stub = types.ModuleType("pytorch_lightning")
    stub.LightningModule = LightningModule
    stub.__som_stub__ = True
    sys.modules["pytorch_lightning"] = stub


def load_reference_som_layer():
    """Return the reference's SOMLayer class, loaded from its own file, unchanged."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT}/models/som_layer.py not found")
    _install_lightning_stub()
    path = os.path.join(REFERENCE_ROOT, "models", "som_layer.py")
    spec = importlib.util.spec_from_file_location("_reference_som_layer", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SOMLayer


def make_config(map_size, latent_dim, distance_fcn="euclidean", topology="square", Tmax=10.0, Tmin=0.1,
                total_epochs=1, batch_size=1):
    """A config dict in the reference's schema (configs/vit_som/*.yaml) that yields the requested latent_dim
    through the use_reduced branch (models/som_layer.py:35-40)."""
    return {
        "hyperparameters": {
            "model_arch": "vit_som", "total_epochs": total_epochs, "batch_size": batch_size,
            "som": {"map_size": list(map_size), "Tmax": Tmax, "Tmin": Tmin, "topology": topology,
                    "distance_fcn": distance_fcn, "use_reduced": True},
            "vit": {"emb_dim": int(latent_dim), "patch_size": 1},
        },
        "data": {"input_size": 1},
    }
