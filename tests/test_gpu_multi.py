"""Multi-GPU parity (needs >= 2 B200s: ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``).

One process per GPU over NCCL.  The prototype-sharded layer and the batch-sharded data-parallel wrapper are compared
with the CPU oracle's single-process answer on the full problem (tolerances of BASELINE.json: BMU exact up to fp32
near-ties, loss / gradients 1e-5 relative).
"""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import som_oracle as O
from oracle.ref_import import make_config

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, args):
    import datetime
    import traceback
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=90))
    try:
        fn(rank, world, *args)
        torch.cuda.synchronize()
        dist.destroy_process_group()
    except BaseException:  # noqa: BLE001
        # a failed rank must not wait for its peer inside a collective or in destroy_process_group: leave at once
        traceback.print_exc()
        os._exit(1)


def spawn(fn, *args, world=2):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.spawn(_run, args=(world, _free_port(), fn, args), nprocs=world, join=False)
    import time
    deadline = time.time() + 240
    while not ctx.join(timeout=5):                       # raises if a rank failed (and terminates the others)
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail("multi-GPU worker timed out")


def _sharded_worker(rank, world, fcn, dx_mode="sync"):
    from vit_som_b200.distributed import PrototypeShardedSOM, shard_range
    ms, D, B, T = (24, 20), 160, 333, 3.0
    K = ms[0] * ms[1]
    torch.manual_seed(7)                                  # same seed on every rank: identical full-map draw
    layer = PrototypeShardedSOM(make_config(list(ms), D, fcn, Tmax=T)).cuda()
    if dx_mode != "sync":                                 # asynchronous exchange of the latent gradients
        layer.async_dx = True
        layer.dx_overlap = dx_mode                        # "stream": behind the backward launch; "kernel": inside it
    torch.manual_seed(7)
    W_full = torch.rand(K, D)
    if fcn == "cosine":
        W_full = torch.nn.functional.normalize(W_full, p=2, dim=1)
    k0, k1 = shard_range(K, world, rank)
    pos = O.grid_positions(ms)
    for it in range(3 if dx_mode != "sync" else 1):       # repeated calls: the two symmetric buffers alternate
        x_np = np.random.RandomState(1 + it).randn(B, D).astype(np.float32)
        x = torch.as_tensor(x_np).cuda().requires_grad_(True)
        layer.prototypes.grad = None
        d_loc, bmu = layer(x)
        loss = layer.som_loss(layer.compute_weights(bmu), d_loc)
        (loss * 0.5).backward()
        layer.wait_dx()
        bmu_only = layer.best_matching_units(x)            # last collective: every check below is local
        torch.cuda.synchronize()
        assert torch.equal(layer.prototypes.detach().cpu(), W_full[k0:k1])
        assert d_loc.shape == (B, k1 - k0)
        ref = O.step(x_np, W_full.numpy(), pos, T, fcn, 0.5, np.float64, bmu_override=bmu.cpu().numpy())
        _, hard, worst = O.classify_bmu_mismatches(x_np, W_full.numpy(), bmu.cpu().numpy(), fcn)
        assert hard == 0, worst
        assert O.rel_err(d_loc.detach().cpu().numpy(), ref.distances[:, k0:k1]) < 3e-6
        assert abs(loss.item() - float(ref.loss)) <= 1e-5 * abs(float(ref.loss))
        assert O.rel_err(x.grad.cpu().numpy(), ref.grad_x) < 1e-5, f"dx, call {it}, mode {dx_mode}"
        assert O.rel_err(layer.prototypes.grad.cpu().numpy(), ref.grad_w[k0:k1]) < 1e-5
        assert torch.equal(bmu, bmu_only)


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_prototype_sharded_matches_oracle(fcn):
    spawn(_sharded_worker, fcn)


@pytest.mark.parametrize("dx_mode", ["stream", "kernel"])
def test_prototype_sharded_async_dx_exchange(dx_mode):
    """The asynchronous exchange of the latent gradients (joined by wait_dx): enqueued behind the backward launch, or
    started from inside it by the dx-complete counter of the dx-first two-phase schedule."""
    spawn(_sharded_worker, "euclidean", dx_mode)


def _dp_worker(rank, world, fcn):
    from vit_som_b200 import SOMLayer
    from vit_som_b200.distributed import DataParallelSOM
    ms, D, B, T = (12, 12), 200, 192, 2.5
    torch.manual_seed(50 + rank)                          # different initial prototypes: the wrapper broadcasts rank 0's
    layer = SOMLayer(make_config(list(ms), D, fcn, Tmax=T)).cuda()
    DataParallelSOM(layer)
    W = layer.prototypes.detach().cpu().numpy()
    x_np = np.random.RandomState(2).randn(B, D).astype(np.float32)
    r0, r1 = rank * B // world, (rank + 1) * B // world
    x = torch.as_tensor(x_np[r0:r1]).cuda().requires_grad_(True)
    d, bmu = layer(x)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    loss.backward()
    import torch.distributed as dist
    gathered = [torch.empty_like(layer.prototypes.grad) for _ in range(world)]
    dist.all_gather(gathered, layer.prototypes.grad)       # last collective: every check below is local
    torch.cuda.synchronize()
    pos = O.grid_positions(ms)
    full_bmu = O.bmu(O.distances(x_np, W, fcn, np.float64))
    loc = O.step(x_np[r0:r1], W, pos, T, fcn, 1.0, np.float64, bmu_override=bmu.cpu().numpy())
    ref = O.step(x_np, W, pos, T, fcn, 1.0, np.float64)
    assert O.rel_err(x.grad.cpu().numpy(), loc.grad_x) < 1e-5              # dx stays local (scaled by the local mean)
    if (bmu.cpu().numpy() == full_bmu[r0:r1]).all():
        assert O.rel_err(layer.prototypes.grad.cpu().numpy(), ref.grad_w) < 1e-5   # averaged dW = global-batch dW
    assert all(torch.equal(g, gathered[0]) for g in gathered)              # every rank holds the same averaged gradient


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_data_parallel_matches_oracle(fcn):
    spawn(_dp_worker, fcn)


def _dp_repeat_worker(rank, world, mode):
    """Several steps through the same DataParallelSOM (the NVLS barrier flags reset themselves; the symmetric dW buffer
    is reused), NVLS against the NCCL all-reduce on the same inputs."""
    import torch.distributed as dist
    from vit_som_b200 import SOMLayer
    from vit_som_b200.distributed import DataParallelSOM
    ms, D, B, T = (16, 20), 264, 512, 3.0                 # K * D = 84480 floats: several blocks per rank slice
    torch.manual_seed(7)
    layer = SOMLayer(make_config(list(ms), D, "euclidean", Tmax=T)).cuda()
    dp = DataParallelSOM(layer, nvls=None if mode == "auto" else False)
    used_nvls = dp.nvls is not None
    grads = []
    for step in range(4):
        x_np = np.random.RandomState(100 + step).randn(B, D).astype(np.float32)
        r0, r1 = rank * B // world, (rank + 1) * B // world
        x = torch.as_tensor(x_np[r0:r1]).cuda().requires_grad_(True)
        layer.prototypes.grad = None
        d, bmu = layer(x)
        layer.som_loss(layer.compute_weights(bmu), d).backward()
        torch.cuda.synchronize()
        g = layer.prototypes.grad.clone()
        W = layer.prototypes.detach().cpu().numpy()
        ref = O.step(x_np, W, O.grid_positions(ms), T, "euclidean", 1.0, np.float64)
        full_bmu = O.bmu(O.distances(x_np, W, "euclidean", np.float64))
        same = torch.tensor([int((bmu.cpu().numpy() == full_bmu[r0:r1]).all())], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if int(same.item()):
            assert O.rel_err(g.cpu().numpy(), ref.grad_w) < 1e-5, f"step {step} nvls={used_nvls}"
        gathered = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gathered, g)
        assert all(torch.equal(t, gathered[0]) for t in gathered)
        grads.append(g)
    torch.cuda.synchronize()
    print(f"rank {rank}: data-parallel exchange path = {'NVLS multimem kernel' if used_nvls else 'NCCL all-reduce'}"
          f"{'' if used_nvls or mode != 'auto' else ' (' + str(getattr(dp, 'nvls_error', 'no multicast')) + ')'}", flush=True)


@pytest.mark.parametrize("mode", ["auto", "nccl"])
def test_data_parallel_repeated_steps(mode):
    spawn(_dp_repeat_worker, mode)


def test_layer_on_a_device_that_is_not_current():
    """Single process, two GPUs: the layer lives on cuda:1 while cuda:0 is the current device - every wrapper must run
    on the tensors' device and that device's current stream (round-1 finding: launches went to the current device)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from vit_som_b200 import FusedPrototypeAdamW, SOMLayer
    torch.cuda.set_device(0)
    ms, D, B, T = (12, 16), 200, 300, 2.0
    torch.manual_seed(3)
    layer = SOMLayer(make_config(list(ms), D, "euclidean", Tmax=T)).to("cuda:1")
    opt = FusedPrototypeAdamW(layer, lr=1e-3)
    x_np = np.random.RandomState(4).randn(B, D).astype(np.float32)
    x = torch.as_tensor(x_np).to("cuda:1").requires_grad_(True)
    assert torch.cuda.current_device() == 0
    d, bmu = layer(x)
    loss = layer.som_loss(layer.compute_weights(bmu), d)
    loss.backward()
    W = layer.prototypes.detach().cpu().numpy()
    torch.cuda.synchronize("cuda:1")
    assert d.device.index == 1 and x.grad.device.index == 1 and torch.cuda.current_device() == 0
    ref = O.step(x_np, W, O.grid_positions(ms), T, "euclidean", 1.0, np.float64, bmu_override=bmu.cpu().numpy())
    assert abs(loss.item() - float(ref.loss)) <= 1e-5 * abs(float(ref.loss))
    assert O.rel_err(x.grad.cpu().numpy(), ref.grad_x) < 1e-5
    assert O.rel_err(layer.prototypes.grad.cpu().numpy(), ref.grad_w) < 1e-5
    opt.step()                                               # the optimizer kernel on cuda:1 as well
    torch.cuda.synchronize("cuda:1")
    assert not np.array_equal(layer.prototypes.detach().cpu().numpy(), W)
    w = layer.compute_weights(bmu).materialize()
    assert w.device.index == 1
