"""Timeline of the data-parallel backward (run under torchrun on >= 2 GPUs through gpurun):

    python -m torch.distributed.run --nproc-per-node 2 tools/dp_probe.py [overlap] [gemm_sms]

Prints, per rank, when the exchange kernel starts / ends relative to the start of the fused gradient-GEMM launch, the
duration of that launch, and the stand-alone durations of the NVLS kernel and of NCCL's all-reduce on the same buffer.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import WORKLOADS, make_cfg  # noqa: E402
from vit_som_b200 import SOMLayer, ops  # noqa: E402
from vit_som_b200.distributed import DataParallelSOM, all_reduce_mean  # noqa: E402


def main():
    overlap = sys.argv[1] if len(sys.argv) > 1 else "counter"
    sms = int(sys.argv[2]) if len(sys.argv) > 2 else 136
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    desc, B, ms, D, T, fcn = WORKLOADS[os.environ.get("SOM_WORKLOAD", "cfg2")]
    torch.manual_seed(1 + rank)
    layer = SOMLayer(make_cfg(ms, D, fcn, T)).to(dev).train()
    layer.current_temperature = T
    dp = DataParallelSOM(layer, gemm_sm_limit=sms if sms > 0 else None, overlap=overlap)
    x = torch.randn(B, D, device=dev, requires_grad=True)
    stream = torch.cuda.Stream(dev, priority=-1)
    torch.cuda.set_stream(stream)
    marks = []
    orig = dp._reduce

    def timed_reduce(dw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(dw)
        b.record()
        marks.append((a, b))
    dp._reduce = timed_reduce

    def step():
        layer.prototypes.grad = None
        x.grad = None
        d, bmu = layer(x)
        layer.som_loss(layer.compute_weights(bmu), d).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    ops.GEMM_TIMERS = []
    marks.clear()
    n = 20
    ends = []
    for _ in range(n):
        torch.cuda._sleep(4_000_000)
        step()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ends.append(e)
    torch.cuda.synchronize()
    gem = [(s, e) for nm, s, e in ops.GEMM_TIMERS if nm in ("dw+dx", "dw")]
    dxs = [(s, e) for nm, s, e in ops.GEMM_TIMERS if nm == "dx"]
    ops.GEMM_TIMERS = None
    rows = []
    for i in range(n):
        g0, g1 = gem[i]
        a, b = marks[i]
        last = dxs[i][1] if dxs else g1
        rows.append((g0.elapsed_time(g1), g0.elapsed_time(a), g0.elapsed_time(b), g0.elapsed_time(last), g0.elapsed_time(ends[i])))
    med = [sorted(r[j] for r in rows)[n // 2] * 1e3 for j in range(5)]
    # stand-alone exchange timings on the same buffer
    dw = dp.nvls["dw"] if dp.nvls is not None else torch.zeros_like(layer.prototypes)
    def bench(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 20 * 1e3
    t_nvls = bench(lambda: orig(dw)) if dp.nvls is not None else float("nan")
    plain = torch.zeros_like(layer.prototypes)
    t_nccl = bench(lambda: all_reduce_mean(plain))
    print(f"rank {rank} overlap={overlap} gemm_sms={sms} world={dist.get_world_size()} {'NVLS' if dp.nvls is not None else 'NCCL'}: "
          f"first gradient GEMM launch {med[0]:.1f} us | exchange enqueued-start {med[1]:.1f} us, end {med[2]:.1f} us | last GEMM end "
          f"{med[3]:.1f} us | backward joined {med[4]:.1f} us  (all from the GEMM launch's start; medians of {n}) || stand-alone: "
          f"NVLS kernel {t_nvls:.1f} us, NCCL all-reduce {t_nccl:.1f} us for {dw.numel() * 4 / 1e6:.1f} MB", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
