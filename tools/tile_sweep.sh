for bn in 128 192 256; do echo "== SOM_BN=$bn"; SOM_BN=$bn python tools/step_time.py 2>&1 | grep "GEMM" | cut -c1-120; done
echo "== no PDL check"; python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from bench import make_cfg
from vit_som_b200 import SOMLayer, _lib
L = _lib.lib()
layer = SOMLayer(make_cfg((40, 40), 3136, "euclidean", 20.0)).cuda().train(); layer.current_temperature = 20.0
x = torch.randn(1024, 3136, device="cuda", requires_grad=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step():
    layer.prototypes.grad = None; x.grad = None
    d, b = layer(x); layer.som_loss(layer.compute_weights(b), d).backward()
for pdl in (1, 0, 1, 0):
    L.som_set_pdl(pdl)
    for _ in range(5): step()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        step(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s): step()
        ts = []
        for _ in range(50):
            flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); print("pdl", pdl, "median step us", round(ts[25] * 1e3, 1))
PY
