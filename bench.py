"""bench.py — SOM-layer hot path throughput on B200 (BASELINE.json metric: SOM fwd+bwd samples/sec and ViT-SOM train
img/s at 1/2/4/8 B200, % roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A *step* is one pass of the hot path over one synthetic batch: operand staging (latents AND prototypes, which
change every training step) -> tcgen05 distance GEMM + argmin -> neighbourhood-weighted loss -> backward
(R staging + both gradient GEMMs in one launch) -> for N > 1 the prototype-gradient exchange (batch-sharded data
parallel, weak scaling: every rank owns a full batch).  Inputs are resident in HBM for `value`; `e2e` repeats
the measurement through the public module API with the batch coming from pinned host memory each step (copied on a
side stream, double buffered) and the loss read back.  L2 is flushed between the timed steps of `value`.  One JSON
line is printed by rank 0.  Besides the headline (config 2) the same line carries, measured in the same run:

  "cfg5"          config 5 (128x128 map, D=256, batch 65536): one GPU at N=1, prototypes sharded over the N GPUs
                  (cross-GPU (min, index) reduction, loss and dx exchange) at N>1 - strong scaling
  "vit_som"       ViT-SOM training img/s for the two data-parallel configs (3: CIFAR-10-shaped 4x4 map, 4: Tiny-
                  ImageNet-shaped 40x40 map): bf16 ViT autoencoder + fused SOM layer + AdamW, N-way data parallel
  "parity_check"  an untimed check of the single-GPU, data-parallel and prototype-sharded steps of THIS process group
                  against the CPU oracle on a small shape (the oracle is the checker here, never the thing measured)
  "adamw"         the fused prototype optimizer step (update + staging of the next forward) timed next to the step

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
# clocks (NVML sampler thread: started before the warm-up, stopped after the timed region)
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

    def pin_to_gpu_numa_node(self):
        """Bind this process to the CPUs next to the GPU (the pinned staging buffers are then allocated on that NUMA
        node: the end-to-end leg is bound by the host -> device copy).  Returns a short description."""
        if not self.nv:
            return "nvml unavailable"
        try:
            self.nv.nvmlDeviceSetCpuAffinity(self.h)
            return f"{len(os.sched_getaffinity(0))} cpus (nvmlDeviceSetCpuAffinity)"
        except Exception as exc:  # noqa: BLE001
            return f"not pinned ({type(exc).__name__})"

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
            time.sleep(0.002)

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
                "samples": len(self.samples), "window": "warm-up + timed steps"}


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
class Env:
    """What every measurement of one bench process shares."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.clocks = ClockSampler(self.local_rank)
        self.numa = self.clocks.pin_to_gpu_numa_node()          # before any pinned allocation
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        # Everything runs on a high-priority compute stream, so that under data parallelism the (lowest-priority)
        # communication stream of DataParallelSOM only takes SMs the GEMMs leave free.
        self.compute_stream = torch.cuda.Stream(self.dev, priority=-1)
        self.compute_stream.wait_stream(torch.cuda.current_stream(self.dev))
        torch.cuda.set_stream(self.compute_stream)
        self.peaks = load_peaks()
        self.tf32_peak = None

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def _matmul_rate(self, dtype):
        torch = self.torch
        n = 8192
        a = torch.randn(n, n, device=self.dev).to(dtype)
        b = torch.randn(n, n, device=self.dev).to(dtype)
        for _ in range(3):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.matmul(a, b)
            e.record()
            e.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12

    def measure_tf32_peak(self):
        """Dense TF32 tensor-core rate of THIS GPU: torch.matmul (cuBLAS, allow_tf32) on 8192^3, best of 10 - the
        tensor roofline of a kind::tf32 kernel (a 3xTF32 product issues three of these per algorithmic MMA)."""
        torch = self.torch
        if self.tf32_peak is not None:
            return self.tf32_peak
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            self.tf32_peak = self._matmul_rate(torch.float32)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return self.tf32_peak

    def measure_fp16_peak(self):
        """Dense fp16 tensor-core rate of THIS GPU (torch.matmul, fp16 in / fp32 accumulate, 8192^3, best of 10): the
        tensor roofline of a kind::f16 kernel (a 3xFP16 product issues three of these per algorithmic MMA)."""
        if getattr(self, "fp16_peak", None) is None:
            self.fp16_peak = self._matmul_rate(self.torch.float16)
        return self.fp16_peak


def measure_som(env, name, wl, steps, warmup, sharded=False, with_e2e=False, with_roofline=True, clocks=None):
    """Times `steps` steps of workload `wl` on this process group.  Returns a dict of raw results (rank-local times
    already reduced with MAX over ranks)."""
    torch, dist, args = env.torch, env.dist, env.args
    from vit_som_b200 import SOMLayer, _lib, ops
    world, rank, dev = env.world, env.rank, env.dev
    desc, B, ms, D, T, fcn = wl
    L = _lib.lib()
    K_, W_ = steps, max(warmup, 3)
    sharded = sharded and world > 1
    # cfg5: rows processed in chunks so the B x K_local scratch stays bounded (same bytes per chunk at any world size;
    # measured on one GPU: 4096-row chunks 10.38 ms/step, 8192 9.91, 16384 and 32768 10.0)
    chunk = min(B, args.row_chunk if args.row_chunk > 0 else 8192 * (world if sharded else 1))
    if sharded:
        from vit_som_b200.distributed import PrototypeShardedSOM
        torch.manual_seed(1234)                # identical full-map draw on every rank, each keeps its block
        layer = PrototypeShardedSOM(make_cfg(ms, D, fcn, T)).to(dev)
        layer.async_dx = not args.sync_dx      # the dx exchange of a row chunk runs under the next chunk's kernels
        layer.async_loss = not args.sync_dx    # and the scalar loss sum leaves the critical path (joined by wait_dx)
        torch.manual_seed(4321)                # replicated latents
    else:
        torch.manual_seed(1234 + rank)
        layer = SOMLayer(make_cfg(ms, D, fcn, T)).to(dev)
    layer.train()
    layer.current_temperature = T
    K_local = layer.prototypes.shape[0]
    x_full = torch.randn(B, D, device=dev)
    # one leaf tensor per row chunk (a chunk is what one call of the module sees)
    x_dev = [x_full[r0:r0 + chunk].clone().requires_grad_(True) for r0 in range(0, B, chunk)]
    del x_full
    chunked = len(x_dev) > 1
    if chunked:                                # chunked: the dW GEMM epilogue accumulates in place across chunks,
        layer.grad_accumulator = torch.zeros(K_local, D, device=dev)
        layer.batch_rows = B                   # and every chunk's loss is its share of the mean over the whole batch

    dp = None
    dp_mode = None
    if world > 1 and not sharded:              # batch-sharded DP: prototype-gradient exchange over NVLink
        from vit_som_b200.distributed import DataParallelSOM
        gemm_sms = 136 if args.gemm_sms < 0 else args.gemm_sms
        dp = DataParallelSOM(layer, gemm_sm_limit=gemm_sms if gemm_sms > 0 else None, overlap=args.dp_overlap)
        dp_mode = dp.overlap
        if chunked:
            dp.detach()                        # chunked + data parallel: one exchange of the accumulated gradient

    # (worth its price - a second phase on 136 SMs - when the exchange of one chunk takes longer than ~60 us at the
    # NVSwitch's all-reduce rate of ~274 GB/s, i.e. from 4 GPUs up at config 5)
    kernel_last = sharded and not args.no_kernel_overlap and chunk * D * 4 / 274e3 > 60.0

    def hot_path(xs):
        """One step through the public module API (vit_som.py:82-86 call sequence + backward) over all row chunks."""
        layer.invalidate_staging()             # prototypes change every training step: their staging is in the step
        layer.prototypes.grad = None
        if not chunked:
            x = xs[0]
            x.grad = None
            d, bmu = layer(x)
            if sharded and layer.async_dx:
                layer.dx_overlap = "kernel" if kernel_last else "stream"
            loss = layer.som_loss(layer.compute_weights(bmu), d)
            loss.backward()
            if sharded:
                layer.wait_dx()
            return loss.detach()
        layer.grad_accumulator.zero_()
        losses = []
        for i, x in enumerate(xs):
            x.grad = None
            if sharded and layer.async_dx:
                # the last row chunk has no successor to hide its dx exchange under: its backward computes the dx tiles
                # first and the exchange runs beside the dW tiles of the same launch
                layer.dx_overlap = "kernel" if i == len(xs) - 1 and kernel_last else "stream"
            d, bmu = layer(x)
            loss = layer.som_loss(layer.compute_weights(bmu), d)
            loss.backward()
            losses.append(loss.detach())
        if dp is not None:
            dp.reduce_accumulator()
        if sharded:
            layer.wait_dx()                    # join the asynchronous dx exchanges: the step ends with complete gradients
        return torch.stack(losses).sum()

    for _ in range(W_):
        hot_path(x_dev)
    env.barrier()

    # ---- timed region: K steps, HBM-resident inputs, L2 flushed (untimed) between steps ----
    # The step (5 kernels, no host dependence) is captured once in a CUDA graph and replayed: the eager Python /
    # autograd path costs 150-250 us of host time per step, which is of the order of the GPU time and would make the
    # number depend on the host.  --no-graph times the eager module calls instead.
    L.som_launch_count_reset()
    hot_path(x_dev)
    torch.cuda.synchronize(dev)
    launches_per_step = int(L.som_launch_count())
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=env.compute_stream):
                hot_path(x_dev)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize(dev)
        except Exception as exc:  # noqa: BLE001
            print(f"bench[{name}]: CUDA graph capture failed ({exc!r}); timing the eager path", file=sys.stderr)
            graph = None
            torch.cuda.synchronize(dev)
    graph_ok = env.max_over_ranks(0.0 if graph is not None else 1.0)[0] == 0.0
    if not graph_ok:
        graph = None                           # every rank takes the same path (the step contains collectives)

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            hot_path(x_dev)

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K_)]
    env.barrier()
    for i in range(K_):
        env.flush_buf.zero_()
        evs[i][0].record()
        run_step()
        evs[i][1].record()
    env.barrier()
    if clocks is not None:
        clocks.__exit__()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    out = {"name": name, "chunk": chunk, "sharded": sharded, "K_local": K_local, "launches_per_step": launches_per_step,
           "graph": graph is not None, "steps": K_, "warmup": W_, "dp_mode": dp_mode,
           "nvls": bool(dp is not None and dp.nvls is not None) or bool(sharded and getattr(layer, "use_nvls", False) and
                                                                        any(v is not None for v in layer._nvls_dx.values()))}

    # ---- roofline leg: the same steps with CUDA events around each tensor-core GEMM launch ----
    gemm_ms = {}
    if with_roofline:
        ops.GEMM_TIMERS = []
        n_inst = min(K_, 30)
        env.barrier()
        for i in range(n_inst):
            env.flush_buf.zero_()
            # a GPU-side delay lets the host enqueue the whole eager step before the first kernel starts, so the events
            # around each GEMM launch see GPU time only (not the ~50 us the host needs between two eager launches)
            torch.cuda._sleep(3_000_000)
            hot_path(x_dev)
        env.barrier()
        for nm, s, e in ops.GEMM_TIMERS:
            gemm_ms.setdefault(nm, []).append(s.elapsed_time(e))
        ops.GEMM_TIMERS = None
    out["gemm_ms"] = gemm_ms

    # ---- fused prototype AdamW (update + staging of the next forward), timed on its own after one step ----
    if not sharded and not chunked and rank == 0 and world == 1:
        from vit_som_b200 import FusedPrototypeAdamW
        hot_path(x_dev)
        opt = FusedPrototypeAdamW(layer, lr=1e-3)
        for _ in range(3):
            opt.step()
        ts = []
        for _ in range(10):
            env.flush_buf.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            opt.step()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        # the step with the prototype staging taken over by the optimizer kernel (staging kernel handles x only)
        def train_step():
            layer.prototypes.grad = None
            x = x_dev[0]
            x.grad = None
            d, bmu = layer(x)
            loss = layer.som_loss(layer.compute_weights(bmu), d)
            loss.backward()
            opt.step()
            return loss.detach()
        for _ in range(3):
            train_step()
        torch.cuda.synchronize(dev)
        tg = None
        try:
            tg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tg, stream=env.compute_stream):
                train_step()
            tg.replay()
            torch.cuda.synchronize(dev)
        except Exception as exc:  # noqa: BLE001
            print(f"bench[{name}]: train-step graph capture failed ({exc!r})", file=sys.stderr)
            tg = None
            torch.cuda.synchronize(dev)
        tt = []
        for _ in range(min(K_, 30)):
            env.flush_buf.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            tg.replay() if tg is not None else train_step()
            e.record()
            e.synchronize()
            tt.append(s.elapsed_time(e))
        # reads W, dW, m, v; writes W, m, v and the staging W_hi, W_lo (fp32 containers, or fp16 under fp16x3)
        nbytes = (8 if getattr(env, "precision", "tf32x3") == "fp16x3" else 9) * K_local * D * 4
        out["adamw"] = {"kernel": "adamw_stage_kernel (AdamW update + operand staging of the new prototypes)",
                        "ms": statistics.median(ts), "algorithmic_bytes": nbytes,
                        "achieved_gbs": nbytes / (statistics.median(ts) * 1e-3) / 1e9,
                        "frac_of_hbm_peak": nbytes / (statistics.median(ts) * 1e-3) / 1e9 / env.peaks["hbm_gbs"],
                        "train_step_ms": statistics.median(tt),
                        "train_step": "forward (stages x only) + loss + backward + fused AdamW, CUDA graph" if tg is not None
                        else "forward + loss + backward + fused AdamW, eager"}
        layer.invalidate_staging()
        del opt

    # ---- end to end: every step's batch comes from pinned host memory, the loss is read back ----
    e2e_ms = None
    if with_e2e:
        x_host = torch.randn(B, D).pin_memory()
        x_stage = [[torch.empty_like(c).requires_grad_(True) for c in x_dev] for _ in range(2)]
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(dev)
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
                e2e_graphs = [StepGraph(lambda j=j: hot_path(x_stage[j]), warmup=1, stream=env.compute_stream)
                              for j in range(2)]
                torch.cuda.synchronize(dev)
            except Exception as exc:  # noqa: BLE001
                print(f"bench[{name}]: e2e graph capture failed ({exc!r}); eager module calls", file=sys.stderr)
                e2e_graphs = None
                torch.cuda.synchronize(dev)
        if env.max_over_ranks(0.0 if e2e_graphs is not None else 1.0)[0] != 0.0:
            e2e_graphs = None
        for j in range(2):
            consumed[j].record(main_stream)
        e2e_start, e2e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        env.flush_buf.zero_()
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
        env.barrier()
        e2e_ms = e2e_start.elapsed_time(e2e_end)
        out["e2e_path"] = ("vit_som_b200.StepGraph replay of the module-API step" if e2e_graphs is not None
                           else "eager module-API calls")
        del e2e_graphs

    total_ms, e2e_max = env.max_over_ranks(total_ms, e2e_ms if e2e_ms is not None else 0.0)
    out["total_ms"], out["e2e_ms"] = total_ms, (e2e_max if e2e_ms is not None else None)
    if dp is not None:
        dp.detach()
    del graph
    return out


def roofline_of(env, wl, res, traffic=None):
    """The roofline object of one measurement: tensor-bound (TF32 tensor peak measured on this GPU; a 3xTF32 product
    issues three tensor-core MMAs per algorithmic one) or HBM-bound (arithmetic intensity
    0.75 B K / (B + K) flop per byte against the 3xTF32-adjusted ridge)."""
    desc, B, ms, D, T, fcn = wl
    peaks = env.peaks
    chunk, K_local = res["chunk"], res["K_local"]
    tf32_peak = env.measure_tf32_peak()
    f16 = getattr(env, "precision", "tf32x3") == "fp16x3"
    kind_peak = env.measure_fp16_peak() if f16 else tf32_peak      # dense rate of the instruction kind the kernels issue
    ms_per_step = res["total_ms"] / res["steps"]
    intensity = 0.75 * chunk * K_local / (chunk + K_local)
    ridge = (kind_peak / 3.0) * 1e12 / (peaks["hbm_gbs"] * 1e9)
    rows = B                                                  # per GPU: a shard scores ALL rows against its K_local prototypes
    step_flops = 6.0 * rows * K_local * D
    step_bytes = 4.0 * (2 * rows * D + 2 * K_local * D) + 8 * rows + 4
    gemm_ms = res["gemm_ms"]
    n_launch = sum(len(v) for v in gemm_ms.values())
    gemm_total_ms = sum(sum(v) for v in gemm_ms.values())
    per_step_gemm_ms = sum(sum(v) / len(v) for v in gemm_ms.values()) * (B // chunk if B > chunk else 1) if gemm_ms else None
    if intensity < ridge:
        achieved = step_bytes / (ms_per_step * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": f"whole step ({res['launches_per_step']} launches; arithmetic intensity "
                f"{intensity:.1f} flop/B is below the ridge {ridge:.0f}; at {step_bytes / 1e6:.1f} MB per step the step "
                "is launch-latency bound rather than bandwidth bound)",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "peak_basis": f"{peaks['source']} HBM copy bandwidth (MEASURED_PEAKS.json)",
                "algorithmic_bytes_per_step": step_bytes, "traffic": traffic}
    per_gemm_flops = 2.0 * chunk * K_local * D
    total_flops = sum(len(v) * per_gemm_flops * (2 if "+" in k else 1) for k, v in gemm_ms.items())
    achieved_tf = total_flops / (gemm_total_ms * 1e-3) / 1e12 if gemm_total_ms else 0.0
    step_tf = step_flops / (ms_per_step * 1e-3) / 1e12
    # The bound of a 3-product split is a third of the dense rate of the instruction kind it issues.  Both bounds are
    # reported: against the kind the kernels use (fp16x3: the measured fp16 rate; the driver's bf16 burst figure is the
    # same datapath) and against the TF32 rate, which is what rounds 1-2 quoted (a 3xFP16 kernel may exceed THAT bound:
    # it is the ceiling of the 3xTF32 formulation, not of the hardware).
    basis = ("fp16 dense rate measured on this GPU in this run (torch.matmul fp16, 8192^3, best of 10); the kernels issue "
             "tcgen05.mma.kind::f16, three per algorithmic MMA (3xFP16 on row-scaled operands)" if f16 else
             "TF32 dense rate measured on this GPU in this run (torch.matmul allow_tf32, 8192^3, best of 10); the kernels "
             "issue tcgen05.mma.kind::tf32, three per algorithmic MMA (3xTF32)")
    basis += (f"; driver-measured bf16 burst {peaks['bf16_tflops']} TFLOP/s ({peaks['source']}), TF32 dense measured here "
              f"{tf32_peak:.1f} TFLOP/s")
    return {
        "bound": "tensor", "kernel": "som_gemm3x_pair_kernel (forward launch + fused dW/dx launch)",
        "precision": "fp16x3" if f16 else "tf32x3",
        "achieved": achieved_tf, "peak": kind_peak, "unit": "TFLOP/s", "frac": achieved_tf / kind_peak,
        "peak_basis": basis,
        "frac_of_3x_bound": achieved_tf / (kind_peak / 3.0),
        "frac_of_3xtf32_bound": achieved_tf / (tf32_peak / 3.0),
        # the same against half of the driver-measured bf16 burst rate (the basis of round 1's fractions; cuBLAS' TF32
        # GEMM itself stays ~10 % below it)
        "frac_of_3xtf32_bound_vs_bf16_half": achieved_tf / (peaks["bf16_tflops"] / 2.0 / 3.0),
        "algorithmic_flops_per_launch": total_flops / max(n_launch, 1),
        "avg_launch_ms": gemm_total_ms / max(n_launch, 1),
        "per_gemm_ms": {k: sum(v) / len(v) for k, v in gemm_ms.items()},
        "gemm_share_of_step": (per_step_gemm_ms / ms_per_step) if per_step_gemm_ms else None,
        "whole_step": {"achieved": step_tf, "frac": step_tf / kind_peak, "frac_of_3x_bound": step_tf / (kind_peak / 3.0),
                       "frac_of_3xtf32_bound": step_tf / (tf32_peak / 3.0),
                       "frac_of_3xtf32_bound_vs_bf16_half": step_tf / (peaks["bf16_tflops"] / 2.0 / 3.0),
                       "algorithmic_flops_per_step": step_flops},
        "traffic": traffic,
    }


def load_traffic(name, world, precision="tf32x3"):
    """DRAM bytes per launch of the dominant kernel from an `ncu --set full` capture of THIS workload at THIS GPU count
    and operand precision (profiles/traffic_r02.json, key "<workload>@n<N>" for tf32x3, "<workload>@n<N>@fp16x3"
    otherwise); null when no such capture exists."""
    path = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if not os.path.exists(path):
        return None
    key = f"{name}@n{world}" + ("" if precision == "tf32x3" else f"@{precision}")
    with open(path) as f:
        return json.load(f).get(key, {}).get("dram_bytes_per_launch")


# ------------------------------------------------------------------------------------------------
# parity check of this process group against the CPU oracle (untimed)
# ------------------------------------------------------------------------------------------------
def parity_check(env):
    """Single-GPU, data-parallel and prototype-sharded step on a small shape, compared with the fp64 oracle's
    single-process answer on the full problem.  Every rank computes its own verdict; rank 0's is printed, and `ok`
    is the AND over ranks."""
    import numpy as np
    torch, dist = env.torch, env.dist
    from oracle import som_oracle as O
    from vit_som_b200 import SOMLayer
    world, rank, dev = env.world, env.rank, env.dev
    # > 128 local rows and > 128 local prototypes at up to 8 GPUs: the fused CTA-pair backward (and its counted,
    # two-phase variant) is the path that is checked
    ms, D, B, T = (40, 32), 192, 256 * world, 3.0
    K = ms[0] * ms[1]
    pos = O.grid_positions(ms)
    res, ok = {}, True
    x_np = np.random.RandomState(1).randn(B, D).astype(np.float32)

    def rel(a, b):
        return float(O.rel_err(a, b))

    # (1) single GPU: this rank alone on the full batch
    torch.manual_seed(7)
    layer = SOMLayer(make_cfg(ms, D, "euclidean", T)).to(dev)
    layer.current_temperature = T
    W = layer.prototypes.detach().cpu().numpy()
    x = torch.as_tensor(x_np).to(dev).requires_grad_(True)
    d, bmu = layer(x)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    loss.backward()
    torch.cuda.synchronize(dev)
    n_bad, hard, _ = O.classify_bmu_mismatches(x_np, W, bmu.cpu().numpy(), "euclidean")
    ref = O.step(x_np, W, pos, T, "euclidean", 1.0, np.float64, bmu_override=bmu.cpu().numpy())
    res["single"] = {"bmu_mismatch": int(n_bad), "bmu_not_near_tie": int(hard),
                     "loss_rel": abs(loss.item() - float(ref.loss)) / abs(float(ref.loss)),
                     "dx_rel": rel(x.grad.cpu().numpy(), ref.grad_x),
                     "dw_rel": rel(layer.prototypes.grad.cpu().numpy(), ref.grad_w)}
    ok &= hard == 0 and max(res["single"]["loss_rel"], res["single"]["dx_rel"], res["single"]["dw_rel"]) < 1e-5
    if world > 1:
        # (2) batch-sharded data parallel: local rows, averaged dW == global-batch dW
        from vit_som_b200.distributed import DataParallelSOM, PrototypeShardedSOM, shard_range
        dp = DataParallelSOM(layer, gemm_sm_limit=136, overlap=env.args.dp_overlap)
        r0, r1 = rank * B // world, (rank + 1) * B // world
        worst = 0.0
        for rep in range(2):                                   # twice: the exchange state resets itself
            xl = torch.as_tensor(x_np[r0:r1]).to(dev).requires_grad_(True)
            layer.prototypes.grad = None
            d, bmu_l = layer(xl)
            layer.som_loss(layer.compute_weights(bmu_l), d).backward()
            torch.cuda.synchronize(dev)
            all_bmu = [torch.empty_like(bmu_l) for _ in range(world)]
            dist.all_gather(all_bmu, bmu_l)
            full_bmu = torch.cat(all_bmu).cpu().numpy()
            refg = O.step(x_np, W, pos, T, "euclidean", 1.0, np.float64, bmu_override=full_bmu)
            worst = max(worst, rel(layer.prototypes.grad.cpu().numpy(), refg.grad_w))
        loc = O.step(x_np[r0:r1], W, pos, T, "euclidean", 1.0, np.float64, bmu_override=bmu_l.cpu().numpy())
        res["dp"] = {"exchange": ("NVLS multimem kernel" if dp.nvls is not None else "NCCL all-reduce") + f", {dp.overlap}",
                     "dw_rel": worst, "dx_rel": rel(xl.grad.cpu().numpy(), loc.grad_x)}
        ok &= worst < 1e-5 and res["dp"]["dx_rel"] < 1e-5
        dp.detach()
        # (3) prototype-sharded: global BMU by the packed (min, index) reduction, global loss, summed dx, local dW shard
        torch.manual_seed(7)
        sh = PrototypeShardedSOM(make_cfg(ms, D, "euclidean", T)).to(dev)
        sh.current_temperature = T
        k0, k1 = shard_range(K, world, rank)
        xs = torch.as_tensor(x_np).to(dev).requires_grad_(True)
        d_loc, bmu_s = sh(xs)
        loss_s = sh.som_loss(sh.compute_weights(bmu_s), d_loc)
        loss_s.backward()
        torch.cuda.synchronize(dev)
        n_bad, hard, _ = O.classify_bmu_mismatches(x_np, W, bmu_s.cpu().numpy(), "euclidean")
        refs = O.step(x_np, W, pos, T, "euclidean", 1.0, np.float64, bmu_override=bmu_s.cpu().numpy())
        res["sharded"] = {"bmu_mismatch": int(n_bad), "bmu_not_near_tie": int(hard),
                          "loss_rel": abs(loss_s.item() - float(refs.loss)) / abs(float(refs.loss)),
                          "dx_rel": rel(xs.grad.cpu().numpy(), refs.grad_x),
                          "dw_shard_rel": rel(sh.prototypes.grad.cpu().numpy(), refs.grad_w[k0:k1])}
        ok &= hard == 0 and max(res["sharded"]["loss_rel"], res["sharded"]["dx_rel"], res["sharded"]["dw_shard_rel"]) < 1e-5
        # (4) the same layer the way the cfg5 record runs it: two row chunks, in-place dW accumulation, asynchronous loss
        # sum and dx exchange - the first chunk's behind its backward launch, the last chunk's from inside it
        sh.async_dx = sh.async_loss = True
        sh.batch_rows = B
        sh.grad_accumulator = torch.zeros_like(sh.prototypes)
        sh.prototypes.grad = None
        half = B // 2
        xa = [torch.as_tensor(x_np[r0:r0 + half]).to(dev).requires_grad_(True) for r0 in (0, half)]
        losses, bmus = [], []
        for i, xc in enumerate(xa):
            sh.dx_overlap = "kernel" if i == 1 else "stream"
            d_c, bmu_c = sh(xc)
            loss_c = sh.som_loss(sh.compute_weights(bmu_c), d_c)
            loss_c.backward()
            losses.append(loss_c)
            bmus.append(bmu_c)
        sh.wait_dx()
        torch.cuda.synchronize(dev)
        bmu_a = torch.cat(bmus).cpu().numpy()
        refa = O.step(x_np, W, pos, T, "euclidean", 1.0, np.float64, bmu_override=bmu_a)
        loss_a = float(sum(l.item() for l in losses))
        res["sharded_async"] = {"loss_rel": abs(loss_a - float(refa.loss)) / abs(float(refa.loss)),
                                "dx_rel": rel(torch.cat([xc.grad for xc in xa]).cpu().numpy(), refa.grad_x),
                                "dw_shard_rel": rel(sh.grad_accumulator.cpu().numpy(), refa.grad_w[k0:k1])}
        ok &= max(res["sharded_async"].values()) < 1e-5
        del sh
    res["ok"] = env.max_over_ranks(0.0 if ok else 1.0)[0] == 0.0
    res["shape"] = {"B": B, "K": K, "D": D, "T": T, "distance": "euclidean"}
    res["tolerance"] = "BMU exact except fp32 near-ties (fp64 oracle); loss / gradients 1e-5 relative"
    return res


# ------------------------------------------------------------------------------------------------
# ViT-SOM training img/s (configs 3 and 4)
# ------------------------------------------------------------------------------------------------
def measure_vit_som(env, tag, dataset, map_size, batch, steps, warmup):
    torch, dist = env.torch, env.dist
    from vit_som_b200.vit_som import ViTSOM, build_optimizers, reference_yaml_config
    world, dev = env.world, env.dev
    cfg = reference_yaml_config(dataset, map_size, batch)
    torch.manual_seed(99)
    model = ViTSOM(cfg).to(dev).train()
    som = model.som_layer
    som.total_iterations = 1000.0 * 500          # (len(dataset) / batch) * epochs of the YAML, order of magnitude
    model.ramp_up_end_step = 10
    if model.classification:                     # the decoder is not part of the classification loss: no gradients
        for n, p in model.vit.named_parameters():
            if n.startswith("dec_"):
                p.requires_grad_(False)
    dp = bucket = None
    if world > 1:
        # data parallel: the prototype gradient is exchanged from inside the SOM backward (DataParallelSOM); the ViT
        # and head gradients live in one flat bucket that is averaged with ONE all-reduce after backward (what DDP's
        # bucketing amounts to for a 5 M parameter model, in a form a CUDA graph can capture)
        from vit_som_b200.distributed import DataParallelSOM
        from vit_som_b200.vit_som import FlatGradBucket
        with torch.no_grad():
            for p in list(model.vit.parameters()) + (list(model.cls_head.parameters()) if model.classification else []):
                dist.broadcast(p, src=0)
        dp = DataParallelSOM(som, gemm_sm_limit=136, overlap=env.args.dp_overlap)
        bucket = FlatGradBucket(list(model.vit.parameters()) +
                                (list(model.cls_head.parameters()) if model.classification else []))
    opt_vit, opt_som = build_optimizers(model, capturable=True)
    size, chans = cfg["data"]["input_size"], cfg["data"]["num_channels"]
    img = torch.randn(batch, chans, size, size, device=dev)
    labels = torch.randint(0, max(cfg["data"]["num_classes"], 1), (batch,), device=dev)

    def train_step():
        total, _ = model.training_loss(img, labels)
        if bucket is not None:
            bucket.zero()
        else:
            opt_vit.zero_grad(set_to_none=True)
        opt_som.zero_grad(set_to_none=True)
        total.backward()
        if bucket is not None:
            bucket.all_reduce_mean()
        opt_vit.step()
        opt_som.step()
        return total

    for _ in range(max(warmup, 3)):
        train_step()
    env.barrier()
    # The whole training step (ViT forward / backward, SOM kernels, gradient exchanges, both optimizers; ~700 launches)
    # is captured in a CUDA graph - at these model sizes the eager step is bound by the host.
    graph = None
    if not env.args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=env.compute_stream):
                graph_loss = train_step()
            graph.replay()
            torch.cuda.synchronize(dev)
        except Exception as exc:  # noqa: BLE001
            print(f"bench[vit {tag}]: CUDA graph capture failed ({exc!r}); eager step", file=sys.stderr)
            graph = None
            torch.cuda.synchronize(dev)
    if env.max_over_ranks(0.0 if graph is not None else 1.0)[0] != 0.0:
        graph = None                           # every rank takes the same path (the step contains collectives)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        if graph is not None:
            graph.replay()
            loss = graph_loss
        else:
            loss = train_step()
    e.record()
    env.barrier()
    ms = env.max_over_ranks(s.elapsed_time(e))[0] / steps
    finite = bool(torch.isfinite(loss).item())
    K, D = som.prototypes.shape
    out = {"workload": f"{tag}: ViT-SOM-cls {dataset}-shaped {size}x{size}x{chans}, {map_size[0]}x{map_size[1]} map, "
                       f"batch {batch} per GPU (emb 192, depth 12, heads 3, patch 4; cosine SOM on {D}-dim patch latents)",
           "img_per_s": batch * world / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "n_gpus": world,
           "precision": f"bf16 autocast ViT (SDPA), fp32-accurate ({getattr(env, 'precision', 'tf32x3')}) SOM layer, fp32 AdamW",
           "step": "forward + CE/SOM loss (device-side gamma ramp, strided SOM input) + backward + AdamW (ViT) + fused "
                   "AdamW (prototypes); " + ("one CUDA graph per step" if graph is not None else "eager launches"),
           "parallelism": "single GPU" if world == 1 else
                          f"dp{world}: one flat-bucket NCCL all-reduce (ViT + head gradients) + DataParallelSOM "
                          f"({'NVLS' if dp.nvls is not None else 'NCCL'} prototype-gradient exchange inside the SOM backward)",
           "loss_finite": finite}
    if dp is not None:
        dp.detach()
    return out


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
                    help="rows per module call (0 = 8192, times the world size when prototypes are sharded)")
    ap.add_argument("--gemm-sms", type=int, default=-1,
                    help="data parallel: SMs the gradient GEMMs may occupy while the dW exchange runs (0 = all, -1 = 136)")
    ap.add_argument("--dp-overlap", default="split", choices=["counter", "split", "after"],
                    help="data parallel: how the dW exchange overlaps the backward (see DataParallelSOM)")
    ap.add_argument("--sync-dx", action="store_true",
                    help="prototype-sharded: exchange the latent gradients inside backward (default: asynchronous, joined "
                         "at the end of the step)")
    ap.add_argument("--no-kernel-overlap", action="store_true",
                    help="prototype-sharded: do not overlap the last chunk's dx exchange inside its backward launch")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip cfg5, ViT-SOM img/s and the parity check")
    ap.add_argument("--no-vit", action="store_true", help="skip the ViT-SOM img/s records")
    ap.add_argument("--precision", default=None, choices=["fp16x3", "tf32x3"],
                    help="operand precision of the SOM layer's tensor-core contractions (default: the library default)")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.distance:
        wl[5] = args.distance
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    env = Env(args)
    torch, dist = env.torch, env.dist
    world, rank, dev = env.world, env.rank, env.dev
    desc, B, ms, D, T, fcn = wl
    from vit_som_b200 import ops as _ops
    if args.precision:
        _ops.DEFAULT_PRECISION = args.precision           # every layer built below follows it
    env.precision = _ops.DEFAULT_PRECISION
    sharded = world > 1 and (args.shard == "prototypes" or (args.shard == "auto" and args.workload == "cfg5"))

    parity = None
    if not args.no_extras:
        try:
            parity = parity_check(env)
        except Exception as exc:  # noqa: BLE001  (recorded in the line; the timed measurements still run)
            parity = {"ok": False, "error": repr(exc)}
            print(f"bench: parity check failed ({exc!r})", file=sys.stderr)

    env.clocks.__enter__()                      # sampling from the warm-up on; stopped at the end of the timed region
    head = measure_som(env, args.workload, wl, args.steps, args.warmup, sharded=sharded, with_e2e=True,
                       clocks=env.clocks)
    extras = {}
    if not args.no_extras:
        if args.workload != "cfg5":
            c5 = list(WORKLOADS["cfg5"])
            r5 = measure_som(env, "cfg5", c5, max(5, min(args.steps, 10)), 3, sharded=world > 1, with_e2e=False)
            if rank == 0:
                ms5 = r5["total_ms"] / r5["steps"]
                extras["cfg5"] = {
                    "workload": f"cfg5: {c5[0]}", "ms_per_step": ms5, "samples_per_s": c5[1] / (ms5 * 1e-3),
                    "steps": r5["steps"], "scaling": "strong (one global batch of 65536 rows at every N)",
                    "parallelism": workload_config("cfg5", c5, world, r5["chunk"], world > 1)["parallelism"],
                    "row_chunk": r5["chunk"], "cuda_graph": r5["graph"], "launches_per_step": r5["launches_per_step"],
                    "dx_exchange": ("NVLS multimem kernel" if r5["nvls"] else "NCCL all-reduce") if world > 1 else None,
                    "roofline": roofline_of(env, c5, r5, load_traffic("cfg5", world, env.precision))}
        if not args.no_vit:
            vit = {}
            for tag, dataset, msz, bsz in (("cfg3", "cifar-10", (4, 4), 128), ("cfg4", "tiny-imagenet", (40, 40), 512)):
                try:
                    vit[tag] = measure_vit_som(env, tag, dataset, msz, bsz, 20, 5)
                except Exception as exc:  # noqa: BLE001
                    print(f"bench: ViT-SOM {tag} failed ({exc!r})", file=sys.stderr)
                    vit[tag] = {"error": repr(exc)}
            # share of the SOM layer: the SOM step of the same shape (cosine, as the YAMLs select), timed standalone
            if world == 1:
                for tag, key in (("cfg3", "cfg3"), ("cfg4", "cfg4")):
                    if "ms_per_step" not in vit[tag]:
                        continue
                    w = list(WORKLOADS[key])
                    w[5] = "cosine"
                    r = measure_som(env, key, w, 20, 3, with_e2e=False, with_roofline=False)
                    som_ms = r["total_ms"] / r["steps"]
                    vit[tag]["som_step_ms_standalone"] = som_ms
                    vit[tag]["som_share_of_step"] = som_ms / vit[tag]["ms_per_step"]
            extras["vit_som"] = vit

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

    K_ = head["steps"]
    total_ms, e2e_ms = head["total_ms"], head["e2e_ms"]
    samples_per_step = B if head["sharded"] else B * world       # prototype sharding: one global batch (strong scaling)
    value = samples_per_step * K_ / (total_ms * 1e-3)
    e2e_value = samples_per_step * K_ / (e2e_ms * 1e-3)
    h2d = B * D * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": head["warmup"],
        "ms_per_step": total_ms / K_, "higher_is_better": True, "scaling": "strong" if head["sharded"] else "weak",
        "vs_baseline": None,
        "dtype": ("fp32 (3xFP16 tensor-core products on row-scaled operands, fp32 accumulate)" if env.precision == "fp16x3"
                  else "fp32 (3xTF32 tensor-core products, fp32 accumulate)"),
        "precision": env.precision, "data": "synthetic",
        "config": workload_config(args.workload, wl, world, head["chunk"], head["sharded"]),
        "cuda_graph": head["graph"],
        "roofline": roofline_of(env, wl, head, load_traffic(args.workload, world, env.precision)),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "h2d_gbs_per_gpu": h2d * K_ / (e2e_ms * 1e-3) / 1e9, "host_pinning": env.numa,
                "path": head["e2e_path"]},
        "gpu_launches": head["launches_per_step"] * K_,
        "launches_per_step": head["launches_per_step"],
        "clocks": env.clocks.summary(),
    }
    if head["dp_mode"]:
        line["dp_exchange"] = ("NVLS multimem kernel" if head["nvls"] else "NCCL all-reduce") + f", overlap={head['dp_mode']}"
    if "adamw" in head:
        line["adamw"] = head["adamw"]
    line.update(extras)
    if parity is not None:
        line["parity_check"] = parity
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
