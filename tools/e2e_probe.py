"""Where does the end-to-end leg of bench.py spend its time?  (run on a B200 through gpurun)
Prints the pinned-host -> device copy time of one cfg2 batch, the host time of one eager step through the module
API (launch-only, no synchronisation) and the GPU time of the same step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.ref_import import make_config  # noqa: E402
from vit_som_b200 import SOMLayer  # noqa: E402

B, D, ms = 1024, 3136, (40, 40)
dev = torch.device("cuda", 0)
layer = SOMLayer(make_config(list(ms), D, "euclidean", Tmax=20.0, Tmin=1e-3)).to(dev)
xh = torch.randn(B, D).pin_memory()
xd = torch.empty(B, D, device=dev, requires_grad=True)


def ev():
    return torch.cuda.Event(enable_timing=True)


# 1. H2D
for _ in range(3):
    with torch.no_grad():
        xd.copy_(xh, non_blocking=True)
torch.cuda.synchronize()
a, b = ev(), ev()
a.record()
for _ in range(20):
    with torch.no_grad():
        xd.copy_(xh, non_blocking=True)
b.record()
torch.cuda.synchronize()
ms_copy = a.elapsed_time(b) / 20
print(f"H2D {xh.numel() * 4 / 1e6:.1f} MB pinned: {ms_copy * 1e3:.1f} us  ({xh.numel() * 4 / ms_copy / 1e6:.1f} GB/s)")


def step():
    layer.invalidate_staging()
    layer.prototypes.grad = None
    xd.grad = None
    d, bmu = layer(xd)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
# 2. host time per eager step (GPU idle-free: queue stays ahead because we sleep the GPU first)
torch.cuda._sleep(200_000_000)
t0 = time.perf_counter()
a.record()
for _ in range(50):
    step()
b.record()
t_host = (time.perf_counter() - t0) / 50
torch.cuda.synchronize()
print(f"eager step: host {t_host * 1e6:.1f} us per step (launch only)")
a.record()
for _ in range(50):
    step()
b.record()
torch.cuda.synchronize()
print(f"eager step: GPU-side {a.elapsed_time(b) / 50 * 1e3:.1f} us per step (back to back, L2 warm)")
