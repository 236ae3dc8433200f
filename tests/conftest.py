import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu through gpurun)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    import numpy as np
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


# Every GPU test module runs under both operand precisions of the tensor-core contractions (3xTF32 and 3xFP16): the
# modules ask for the fixture with `pytest.mark.usefixtures("som_precision")`, layers without an explicit `precision`
# attribute follow ops.DEFAULT_PRECISION.
PRECISIONS = ["tf32x3", "fp16x3"]


def pytest_generate_tests(metafunc):
    if "som_precision" in metafunc.fixturenames:
        metafunc.parametrize("som_precision", PRECISIONS, indirect=True)


@pytest.fixture
def som_precision(request):
    from vit_som_b200 import ops
    old = ops.DEFAULT_PRECISION
    ops.DEFAULT_PRECISION = request.param
    yield request.param
    ops.DEFAULT_PRECISION = old
