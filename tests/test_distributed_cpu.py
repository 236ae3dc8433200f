"""World-size-2 tests of the multi-GPU exchange logic (vit_som_b200/distributed.py) on CPU over gloo.

The CUDA kernels cannot run here, so each rank's *local* numbers come from the CPU oracle (test infrastructure)
and the test drives the same exchange helpers the product classes call, in the same order:

  prototype-sharded : local distances of the shard -> packed (key, global index) minima -> reduce_packed_min ->
                      unpack_bmu -> local partial loss -> all_reduce_sum -> partial dx -> all_reduce_sum
  batch-sharded DP  : local step on half the batch -> DataParallelSOM's dW hook (all_reduce_mean)

and checks the combined result against the oracle's single-process answer on the full problem.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import som_oracle as O
from oracle.ref_import import make_config


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def pack_keys_np(key: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """numpy restatement of the device-side pack_key (csrc/som_gemm.cuh): order-preserving float32 -> int32 map in
    the high word, global index in the low word."""
    i = key.astype(np.float32).view(np.int32).astype(np.int64)
    i = np.where(i < 0, i ^ 0x7FFFFFFF, i)
    return (i << 32) | idx.astype(np.int64)


def _run(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def spawn(fn, *args, world=2):
    mp.spawn(_run, args=(world, _free_port(), fn, args), nprocs=world, join=True)


def test_shard_range_partitions_any_map():
    from vit_som_b200.distributed import shard_range
    for n in (1, 2, 7, 16, 1600, 16384):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_pack_keys_order():
    """The packed ordering is the (key, index) lexicographic order, including negative keys (cosine can be -1e-7)."""
    key = np.array([0.5, -1e-7, 0.0, 0.5, 3.0, -2.0], np.float32)
    idx = np.array([3, 1, 2, 0, 4, 5])
    p = pack_keys_np(key, idx)
    order = np.argsort(p, kind="stable")
    assert list(order) == [5, 1, 2, 3, 0, 4]


def _sharded_worker(rank, world, fcn, seed):
    from vit_som_b200 import distributed as D
    rng = np.random.RandomState(seed)
    ms, Dlat, B, T = (6, 7), 24, 37, 1.7
    K = ms[0] * ms[1]
    x = rng.randn(B, Dlat).astype(np.float32)
    W = rng.rand(K, Dlat).astype(np.float32)
    W[5] = W[29]                                        # duplicate prototypes in different shards: lowest index must win
    x[3] = W[29]
    pos = O.grid_positions(ms)
    ref = O.step(x, W, pos, T, fcn, 1.0, np.float64)

    k0, k1 = D.shard_range(K, world, rank)
    d_loc = O.distances(x, W[k0:k1], fcn, np.float64)
    key = d_loc * d_loc if fcn == "euclidean" else d_loc
    j = np.argmin(key, axis=1)
    packed = torch.from_numpy(pack_keys_np(key[np.arange(B), j], j + k0))
    D.reduce_packed_min(packed)
    bmu = D.unpack_bmu(packed, K).numpy()
    np.testing.assert_array_equal(bmu, ref.bmu)
    assert bmu[3] == 5

    w_loc = O.weights(bmu, pos, T, np.float64)[:, k0:k1]
    loss = torch.tensor((w_loc * d_loc).sum() / (B * K))
    D.all_reduce_sum(loss)
    assert abs(loss.item() - float(ref.loss)) < 1e-12 * abs(float(ref.loss)) + 1e-15

    # partial gradients of the local shard (same closed forms as the kernels, oracle arithmetic)
    if fcn == "euclidean":
        with np.errstate(divide="ignore", invalid="ignore"):
            R = np.where(d_loc == 0, 0.0, (w_loc / (B * K)) / d_loc)
        dx_part = x * R.sum(1, keepdims=True) - R @ W[k0:k1]
        dw_loc = W[k0:k1] * R.sum(0)[:, None] - R.T @ x
    else:
        xh, xden = O.l2_normalize(x, np.float64)
        wh, wden = O.l2_normalize(W[k0:k1], np.float64)
        G = w_loc / (B * K)
        c = G * (1.0 - d_loc)
        dx_part = (c.sum(1, keepdims=True) * xh - G @ wh) / xden[:, None]
        dw_loc = (c.sum(0)[:, None] * wh - G.T @ xh) / wden[:, None]
    dx = torch.from_numpy(dx_part.copy())
    D.all_reduce_sum(dx)
    assert O.rel_err(dx.numpy(), ref.grad_x) < 1e-10
    assert O.rel_err(dw_loc, ref.grad_w[k0:k1]) < 1e-10


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_prototype_sharded_exchange_world2(fcn):
    spawn(_sharded_worker, fcn, 11)


def test_prototype_sharded_exchange_world3_uneven():
    spawn(_sharded_worker, "euclidean", 5, world=3)     # 42 cells over 3 ranks, and 37 rows: nothing divides evenly


def _dp_worker(rank, world, fcn):
    from vit_som_b200 import SOMLayer
    from vit_som_b200.distributed import DataParallelSOM
    torch.manual_seed(100 + rank)                       # ranks start with DIFFERENT prototypes: broadcast must fix it
    ms, Dlat, B, T = (5, 4), 16, 24, 2.0
    layer = SOMLayer(make_config(list(ms), Dlat, fcn))
    dp = DataParallelSOM(layer)
    W = layer.prototypes.detach().numpy().copy()
    gathered = [torch.empty_like(layer.prototypes.data) for _ in range(world)]
    dist.all_gather(gathered, layer.prototypes.data)
    assert all(torch.equal(g, gathered[0]) for g in gathered)

    rng = np.random.RandomState(3)
    x = rng.randn(B, Dlat).astype(np.float32)
    pos = O.grid_positions(ms)
    ref = O.step(x, W, pos, T, fcn, 1.0, np.float64)
    r0, r1 = rank * B // world, (rank + 1) * B // world
    loc = O.step(x[r0:r1], W, pos, T, fcn, 1.0, np.float64)         # local loss = mean over the local rows (DDP)
    dw = torch.from_numpy(loc.grad_w.copy())
    join = dp._on_dw(dw)                                            # the hook FusedLossFn.backward calls
    assert join is None
    assert O.rel_err(dw.numpy(), ref.grad_w) < 1e-10                # mean of local-mean gradients = global-mean gradient
    np.testing.assert_array_equal(loc.bmu, ref.bmu[r0:r1])
    dp.detach()
    assert layer._dw_hook is None


@pytest.mark.parametrize("fcn", ["euclidean", "cosine"])
def test_data_parallel_gradient_average_world2(fcn):
    spawn(_dp_worker, fcn)


def _bucket_worker(rank, world):
    """FlatGradBucket (vit_som.py): gradients are views of one flat buffer, one all-reduce averages all of them."""
    from vit_som_b200.vit_som import FlatGradBucket
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    bucket = FlatGradBucket(list(lin.parameters()) + [frozen])
    assert bucket.flat.numel() == 5 * 3 + 3 and frozen.grad is None
    assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    for step in range(2):
        bucket.zero()
        x = torch.full((4, 5), float(rank + 1 + step))
        lin(x).sum().backward()                             # autograd accumulates INTO the views
        assert lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
        local = bucket.flat.clone()
        bucket.all_reduce_mean()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat, torch.stack(gathered).mean(0))
        assert torch.allclose(lin.bias.grad, torch.full((3,), 4.0))


def test_flat_grad_bucket_world2():
    spawn(_bucket_worker)
