"""bench.py — SOM-layer hot path throughput on B200 (BASELINE.json metric: SOM fwd+bwd samples/sec, % roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A *step* is one pass of the hot path over one synthetic batch: operand staging (latents AND prototypes, which
change every training step) -> tcgen05 distance GEMM + argmin -> neighbourhood-weighted loss -> backward
(R staging + the two gradient GEMMs) -> for N > 1 the prototype-gradient all-reduce (batch-sharded data
parallel, weak scaling: every rank owns a full batch).  Inputs are resident in HBM for `value`; `e2e` repeats
the measurement through the public module API with the batch coming from pinned host memory each step (copied on a
side stream, double buffered) and the loss read back.  L2 is flushed between the timed steps of `value`.  One JSON line is printed by rank 0.

`--impl reference` times the reference's CPU implementation of the same step on the host cores (the unmodified
reference module when /root/reference is present, else oracle/som_torch_ref.py which issues the same ATen calls).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, batch, map, latent dim, Tmax, distance)
    "cfg1": ("ViT-SOM 24x24 clustering, MNIST-shaped latent (196 patches x 16)", 256, (24, 24), 3136, 12.0, "euclidean"),
    "cfg2": ("ViT-SOM 40x40 clustering, Fashion-MNIST-shaped latent (196 patches x 16), batch 1024", 1024, (40, 40), 3136,
             20.0, "euclidean"),
    "cfg3": ("ViT-SOM-cls 4x4 map, CIFAR-10-shaped latent (64 patches x 192), batch 128 per GPU", 128, (4, 4), 12288, 4.0,
             "euclidean"),
    "cfg4": ("ViT-SOM-cls 40x40 map, Tiny-ImageNet-shaped latent (256 patches x 192), batch 512", 512, (40, 40), 49152,
             20.0, "euclidean"),
    "cfg5": ("SOM microbench 128x128 map, latent 256, batch 65536 (processed in row chunks)", 65536, (128, 128), 256,
             64.0, "euclidean"),
}
METRIC = "som_fwd_bwd_samples_per_sec"
UNIT = "samples/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def workload_config(name, wl, world, chunk, sharded=False):
    """The `config` object both arms print (same workload description, shapes and settings)."""
    desc, B, ms, D, T, fcn = wl
    if world == 1:
        par = "single GPU"
    elif sharded:
        par = f"prototype-sharded x{world}: (min,index) all-reduce, loss and dx all-reduce (one global batch)"
    else:
        par = f"dp{world}: batch-sharded, prototype-gradient all-reduce"
    return {"workload": f"{name}: {desc}", "B_per_gpu": B, "K": ms[0] * ms[1], "D": D, "distance": fcn, "T": T,
            "row_chunk": chunk, "l2_flush_between_steps": True,
            "parallelism": par,
            "prototype_staging_in_step": True}



def make_cfg(map_size, D, fcn, Tmax):
    """Config dict in the reference's schema (configs/vit_som/*.yaml; parsed by SOMLayer.__init__,
    models/som_layer.py:18-40) giving latent_dim = D through the use_reduced branch."""
    return {
        "hyperparameters": {
            "model_arch": "vit_som", "total_epochs": 1, "batch_size": 1,
            "som": {"map_size": list(map_size), "Tmax": Tmax, "Tmin": 1e-3, "topology": "square",
                    "distance_fcn": fcn, "use_reduced": True},
            "vit": {"emb_dim": int(D), "patch_size": 1},
        },
        "data": {"input_size": 1},
    }


# ------------------------------------------------------------------------------------------------
# clocks (NVML sampler thread, runs during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        busy = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's implementation of the same step
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(B, map_size, D, T, fcn):
    """Returns (callable running one full CPU step, kind, description)."""
    import torch
    from oracle.ref_import import reference_available
    torch.manual_seed(0)
    x = torch.randn(B, D)
    if reference_available():
        from oracle.ref_import import load_reference_som_layer
        layer = load_reference_som_layer()(make_cfg(map_size, D, fcn, T))
        layer.current_temperature = T

        def run():
            xx = x.clone().requires_grad_(True)
            layer.prototypes.grad = None
            d, b = layer(xx)
            loss = layer.som_loss(layer.compute_weights(b), d)
            loss.backward()
            return loss
        return run, "reference", "unmodified /root/reference/models/som_layer.py SOMLayer (torch CPU)"
    from oracle import som_oracle as O
    from oracle import som_torch_ref as R
    W = torch.rand(map_size[0] * map_size[1], D)
    pos = torch.as_tensor(O.grid_positions(map_size))

    def run():
        return R.step(x, W, pos, T, fcn)[2]
    return run, "port", "oracle/som_torch_ref.py (same ATen calls as the reference, torch CPU)"


def time_cpu(run, max_seconds, min_steps, max_steps, warmup=1):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(warmup):
        run()
    times = []
    t_all = time.perf_counter()
    while len(times) < max_steps and (len(times) < min_steps or time.perf_counter() - t_all < max_seconds):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
    return times


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    desc, B, ms, D, T, fcn = wl
    chunk = min(B, 4096)           # cfg5 materialises [B,K,2] on the CPU: timed in row chunks
    run, kind, what = cpu_step_fn(chunk, ms, D, T, fcn)
    for _ in range(max(args.warmup, 1)):
        run()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = chunk * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.gpus, min(B, 4096)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{args.steps} full steps of {chunk} rows; {what}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--distance", default=None, choices=["euclidean", "cosine"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager module calls instead of a CUDA-graph replay")
    ap.add_argument("--shard", default="auto", choices=["auto", "batch", "prototypes"],
                    help="multi-GPU partitioning: batch-sharded DP, or prototype-sharded (auto: prototypes for cfg5)")
    ap.add_argument("--row-chunk", type=int, default=0,
                    help="rows per module call (0 = 4096, times the world size when prototypes are sharded)")
    ap.add_argument("--gemm-sms", type=int, default=-1,
                    help="data parallel: SMs the dx GEMM may occupy while the dW exchange runs (0 = all, -1 = 136)")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.distance:
        wl[5] = args.distance
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from vit_som_b200 import SOMLayer, _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_, K_ = max(args.warmup, 3), args.steps
    desc, B, ms, D, T, fcn = wl
    L = _lib.lib()

    # Partitioning over ranks: batch-sharded data parallel (weak scaling, every rank a full batch) for the ViT-SOM
    # shapes; the 128 x 128 map of cfg5 is prototype-sharded (strong scaling: one global batch, replicated latents).
    sharded = world > 1 and (args.shard == "prototypes" or (args.shard == "auto" and args.workload == "cfg5"))
    # cfg5: rows processed in chunks so the B x K_local scratch stays bounded (same bytes per chunk at any world size)
    chunk = min(B, args.row_chunk if args.row_chunk > 0 else 4096 * (world if sharded else 1))
    if sharded:
        from vit_som_b200.distributed import PrototypeShardedSOM
        torch.manual_seed(1234)                # identical full-map draw on every rank, each keeps its block
        layer = PrototypeShardedSOM(make_cfg(ms, D, fcn, T)).to(dev)
        torch.manual_seed(4321)                # replicated latents
    else:
        torch.manual_seed(1234 + rank)
        layer = SOMLayer(make_cfg(ms, D, fcn, T)).to(dev)
    layer.current_temperature = T
    K_local = layer.prototypes.shape[0]
    x_host = torch.randn(B, D).pin_memory()
    x_full = torch.randn(B, D, device=dev)
    # one leaf tensor per row chunk (a chunk is what one call of the module sees)
    x_dev = [x_full[r0:r0 + chunk].clone().requires_grad_(True) for r0 in range(0, B, chunk)]
    x_stage = [[torch.empty_like(c).requires_grad_(True) for c in x_dev] for _ in range(2)]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    copy_stream = torch.cuda.Stream(dev)
    if len(x_dev) > 1:                         # chunked: the dW GEMM epilogue accumulates in place across chunks
        layer.grad_accumulator = torch.zeros(K_local, D, device=dev)

    dp = None
    if world > 1 and not sharded:              # batch-sharded DP: prototype-gradient all-reduce over NVLink,
        from vit_som_b200.distributed import DataParallelSOM
        # issued on a side stream from inside backward (runs under the dx GEMM, which leaves 20 SMs to NCCL)
        gemm_sms = 136 if args.gemm_sms < 0 else args.gemm_sms
        dp = DataParallelSOM(layer, gemm_sm_limit=gemm_sms if gemm_sms > 0 else None)

    def hot_path(xs):
        """One step through the public module API (vit_som.py:82-86 call sequence + backward) over all row chunks."""
        layer._w_cache = None                  # prototypes change every training step: their staging is in the step
        layer.prototypes.grad = None
        if len(xs) == 1:
            x = xs[0]
            x.grad = None
            d, bmu = layer(x)
            loss = layer.som_loss(layer.compute_weights(bmu), d)
            loss.backward()
            return loss.detach()
        layer.grad_accumulator.zero_()
        total = None
        for x in xs:
            x.grad = None
            d, bmu = layer(x)
            loss = layer.som_loss(layer.compute_weights(bmu), d) * (x.shape[0] / B)
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if dp is not None:                     # chunked + data parallel: one all-reduce of the accumulated gradient
            from vit_som_b200.distributed import all_reduce_mean
            all_reduce_mean(layer.grad_accumulator)
        return total

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # Everything below runs on a high-priority compute stream, so that under data parallelism the (lowest-priority)
    # communication stream of DataParallelSOM only takes SMs the GEMMs leave free.
    compute_stream = torch.cuda.Stream(dev, priority=-1)
    compute_stream.wait_stream(torch.cuda.current_stream(dev))
    torch.cuda.set_stream(compute_stream)

    # ---- warm-up ----
    for _ in range(W_):
        hot_path(x_dev)
    barrier()

    # ---- timed region: K steps, HBM-resident inputs, L2 flushed (untimed) between steps ----
    # The step (6 kernels + 1 memset, no host dependence) is captured once in a CUDA graph and replayed: the eager
    # Python / autograd path costs 150-250 us of host time per step, which is of the order of the GPU time and would
    # make the number depend on the host.  --no-graph times the eager module calls instead.
    L.som_launch_count_reset()
    hot_path(x_dev)
    torch.cuda.synchronize(dev)
    launches_per_step = int(L.som_launch_count())
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=compute_stream):
                graph_loss = hot_path(x_dev)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize(dev)
        except Exception as exc:  # noqa: BLE001
            print(f"bench: CUDA graph capture failed ({exc!r}); timing the eager path", file=sys.stderr)
            graph = None
            torch.cuda.synchronize(dev)

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            hot_path(x_dev)

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K_)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        for i in range(K_):
            flush_buf.zero_()
            evs[i][0].record()
            run_step()
            evs[i][1].record()
        barrier()
    launches = launches_per_step * K_
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)

    # ---- roofline leg: the same steps with CUDA events around each tensor-core GEMM launch ----
    ops.GEMM_TIMERS = []
    n_inst = min(K_, 30)
    inst_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_inst)]
    barrier()
    for i in range(n_inst):
        flush_buf.zero_()
        # a GPU-side delay lets the host enqueue the whole eager step before the first kernel starts, so the events
        # around each GEMM launch see GPU time only (not the ~50 us the host needs between two eager launches)
        torch.cuda._sleep(3_000_000)
        inst_evs[i][0].record()
        hot_path(x_dev)
        inst_evs[i][1].record()
    barrier()
    inst_ms = sum(a.elapsed_time(b) for a, b in inst_evs)
    gemm_ms = {}
    for name, s, e in ops.GEMM_TIMERS:
        gemm_ms.setdefault(name, []).append(s.elapsed_time(e))
    ops.GEMM_TIMERS = None

    # ---- end to end: every step's batch comes from pinned host memory, the loss is read back ----
    # The copy of batch i+1 is issued on a copy stream while step i computes (double-buffered staging, what a
    # data loader with a prefetch queue does); all K copies and all K read-backs lie inside the timed window.
    main_stream = torch.cuda.current_stream(dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(j):
        copy_stream.wait_event(consumed[j])            # the step that last read this buffer has finished
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for ci, c in enumerate(x_stage[j]):
                c.copy_(x_host[ci * chunk:ci * chunk + c.shape[0]], non_blocking=True)
        copied[j].record(copy_stream)

    # The step over each of the two staging buffers is captured once (vit_som_b200.StepGraph: the public helper a
    # user of the layer calls) and replayed, so the host costs one graph launch per step; --no-graph issues the
    # eager module calls.  Copies and read-backs stay outside the graphs, on the copy / compute streams.
    e2e_graphs = None
    if graph is not None:
        try:
            from vit_som_b200 import StepGraph
            e2e_graphs = [StepGraph(lambda j=j: hot_path(x_stage[j]), warmup=1, stream=compute_stream) for j in range(2)]
            torch.cuda.synchronize(dev)
        except Exception as exc:  # noqa: BLE001
            print(f"bench: e2e graph capture failed ({exc!r}); eager module calls", file=sys.stderr)
            e2e_graphs = None
            torch.cuda.synchronize(dev)
    for j in range(2):
        consumed[j].record(main_stream)
    e2e_start, e2e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    flush_buf.zero_()
    e2e_start.record()
    prefetch(0)
    for i in range(K_):
        j = i & 1
        if i + 1 < K_:
            prefetch(j ^ 1)
        main_stream.wait_event(copied[j])
        loss = e2e_graphs[j].replay() if e2e_graphs is not None else hot_path(x_stage[j])
        consumed[j].record(main_stream)
        loss_host.copy_(loss, non_blocking=True)
    e2e_end.record()
    barrier()
    e2e_ms = e2e_start.elapsed_time(e2e_end)

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = t.tolist()
    def finish():
        """Leave without waiting on NCCL teardown: with collectives captured in a CUDA graph destroy_process_group has
        been seen to block at exit; every rank synchronises, rank 0 has printed, then the process exits hard."""
        if world > 1:
            sys.stdout.flush()
            sys.stderr.flush()
            torch.cuda.synchronize(dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = load_peaks()
    Kp = ms[0] * ms[1]
    # algorithmic work of the timed tensor-core launches: 2 * rows * K_local * D flop per GEMM; the fused backward
    # launch ("dw+dx") carries two GEMMs
    n_launch = sum(len(v) for v in gemm_ms.values())
    gemm_total_ms = sum(sum(v) for v in gemm_ms.values())
    avg_gemm_ms = gemm_total_ms / max(n_launch, 1)
    per_gemm_flops = 2.0 * chunk * K_local * D
    total_flops = sum(len(v) * per_gemm_flops * (2 if "+" in k else 1) for k, v in gemm_ms.items())
    per_launch_flops = total_flops / max(n_launch, 1)
    achieved_tf = total_flops / (gemm_total_ms * 1e-3) / 1e12
    tf32_peak = peaks["bf16_tflops"] / 2.0                # tf32 dense = half the bf16 rate on the same tensor pipe
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
    roofline = {
        "bound": "tensor", "kernel": "som_gemm3x_pair_kernel (fwd launch + fused dw/dx launch)", "achieved": achieved_tf,
        "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved_tf / tf32_peak,
        "peak_basis": f"{peaks['source']} bf16 burst {peaks['bf16_tflops']} TFLOP/s / 2 (tf32 dense rate)",
        "frac_of_3xtf32_bound": achieved_tf / (tf32_peak / 3.0),
        "algorithmic_flops_per_launch": per_launch_flops,
        "avg_launch_ms": avg_gemm_ms,
        "per_gemm_ms": {k: sum(v) / len(v) for k, v in gemm_ms.items()},
        "gemm_share_of_step": sum(sum(v) for v in gemm_ms.values()) / inst_ms,
        "traffic": traffic,
    }
    samples_per_step = B if sharded else B * world       # prototype sharding: one global batch (strong scaling)
    value = samples_per_step * K_ / (total_ms * 1e-3)
    e2e_value = samples_per_step * K_ / (e2e_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": total_ms / K_, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None,
        "dtype": "fp32 (3xTF32 tensor-core products, fp32 accumulate)", "data": "synthetic",
        "config": dict(workload_config(args.workload, wl, world, chunk, sharded), cuda_graph=graph is not None),
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": 4,
                "path": "vit_som_b200.StepGraph replay of the module-API step" if e2e_graphs is not None
                else "eager module-API calls"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    if not args.no_cpu_baseline and world == 1:
        cchunk = min(B, 4096)
        run, kind, what = cpu_step_fn(cchunk, ms, D, T, fcn)
        times = time_cpu(run, max_seconds=12.0, min_steps=3, max_steps=200)
        line["cpu_baseline"] = {"value": cchunk / statistics.median(times), "unit": UNIT,
                                "cores": torch.get_num_threads(), "kind": kind,
                                "sample": f"{len(times)} full steps of {cchunk} rows (median); {what}"}
    print(json.dumps(line), flush=True)
    finish()


if __name__ == "__main__":
    main()
