"""CPU checks of the host-side pieces around the hot path that need no GPU: clustering metrics of
vit_som_b200/evaluation.py against scikit-learn / the reference's purity, the ViT position table of the harness
against the reference's generator (when /root/reference is mounted), batched prototype decoding against the one-by-one
loop of the reference, and the sharded layer's full-map state dict over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch


def test_purity_and_nmi_match_sklearn():
    from sklearn.metrics import normalized_mutual_info_score
    from vit_som_b200.evaluation import purity_nmi
    rng = np.random.RandomState(0)
    for n_cells, n_classes, n in [(16, 10, 2000), (576, 10, 5000), (4, 2, 50), (1, 1, 10)]:
        cells = rng.randint(0, n_cells, n)
        labels = (cells + rng.randint(0, 3, n)) % n_classes
        purity, nmi = purity_nmi(torch.as_tensor(cells), torch.as_tensor(labels), n_cells, n_classes)
        table = np.zeros((n_cells, n_classes), np.int64)
        np.add.at(table, (cells, labels), 1)
        assert abs(purity.item() - table.max(1).sum() / n) < 1e-12      # accuracy after majority voting (evaluation.py:132-152)
        assert abs(nmi.item() - normalized_mutual_info_score(labels, cells)) < 1e-9


def test_sincos_table_matches_reference_generator():
    ref_root = "/root/reference"
    if not os.path.exists(os.path.join(ref_root, "tools", "utils.py")):
        pytest.skip("reference not mounted")
    import importlib.util
    import types
    # tools/utils.py imports torchvision transforms at module level: available here
    spec = importlib.util.spec_from_file_location("_ref_utils", os.path.join(ref_root, "tools", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception as exc:  # noqa: BLE001
        pytest.skip(f"reference utils not importable here: {exc!r}")
    from vit_som_b200.vit_som import sincos_table_2d
    for dim, side in [(16, 14), (192, 8), (96, 16)]:
        ref = mod.get_2d_sincos_pos_embed(dim, side, cls_token=True)
        ours = sincos_table_2d(dim, side).numpy()
        np.testing.assert_allclose(ours, ref, rtol=0, atol=2e-6)


def test_vit_autoencoder_shapes_and_batched_prototype_decoding():
    from vit_som_b200.evaluation import decode_prototypes
    from vit_som_b200.vit_som import ViTAutoencoder
    torch.manual_seed(0)
    vit = ViTAutoencoder(16, 4, 3, 32, 2, 2, 16, 1, 2).eval()
    img = torch.randn(5, 3, 16, 16)
    cls, patches, recon = vit(img)
    assert cls.shape == (5, 32) and patches.shape == (5, 16, 32) and recon.shape == img.shape
    assert patches.flatten(1).stride(0) == 17 * 32                   # the copy-free strided SOM input
    # unpatchify is the inverse of the reference's patchify (vit.py:123-135): check on a known pattern
    p, s = 4, 4
    x = img.reshape(5, 3, s, p, s, p).permute(0, 2, 4, 3, 5, 1).reshape(5, s * s, p * p * 3)      # nchpwq -> nhwpqc
    back = x.view(5, s, s, p, p, 3).permute(0, 5, 1, 3, 2, 4).reshape(5, 3, s * p, s * p)
    assert torch.equal(back, img)
    protos = torch.randn(7, 16 * 32)
    batched = decode_prototypes(vit, protos, 16, 32, batch=3)
    one_by_one = torch.cat([vit.decode(torch.cat([torch.zeros(1, 1, 32), q.reshape(1, 16, 32)], dim=1)) for q in protos])
    assert batched.shape == (7, 3, 16, 16)
    assert torch.allclose(batched, one_by_one, atol=1e-5)
    with pytest.raises(ValueError):
        decode_prototypes(vit, torch.randn(2, 100), 16, 32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sharded_state_worker(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.ref_import import make_config
        from vit_som_b200 import SOMLayer
        from vit_som_b200.distributed import PrototypeShardedSOM, shard_range
        cfg = make_config([5, 7], 12, "euclidean")
        torch.manual_seed(3)
        full = SOMLayer(cfg)
        torch.manual_seed(3)
        shard = PrototypeShardedSOM(cfg)
        k0, k1 = shard_range(35, world, rank)
        assert torch.equal(shard.prototypes.detach(), full.prototypes.detach()[k0:k1])
        sd = shard.state_dict()                                       # collective: the FULL map under the reference's key
        assert list(sd.keys()) == ["prototypes", "grid_positions"] and tuple(sd["prototypes"].shape) == (35, 12)
        assert torch.equal(sd["prototypes"], full.prototypes.detach())
        full2 = SOMLayer(cfg)
        full2.load_state_dict(sd)                                     # a sharded checkpoint loads into the plain layer
        assert torch.equal(full2.prototypes, full.prototypes)
        with torch.no_grad():
            full.prototypes.mul_(2.0)
        shard.load_state_dict(full.state_dict())                      # and an unsharded checkpoint loads into the shard
        assert torch.equal(shard.prototypes.detach(), full.prototypes.detach()[k0:k1])
        assert shard._w_cache is None
        # nested under a parent module (prefix handling)
        parent = torch.nn.Module()
        parent.som_layer = shard
        psd = parent.state_dict()
        assert tuple(psd["som_layer.prototypes"].shape) == (35, 12)
        parent.load_state_dict(psd)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_layer_state_dict_holds_the_full_map(world):
    import torch.multiprocessing as mp
    mp.spawn(_sharded_state_worker, args=(world, _free_port()), nprocs=world, join=True)


def test_training_mode_always_restages_and_eval_caches():
    """Host logic of the staging cache (no kernels run: the decision is taken before the launch)."""
    from oracle.ref_import import make_config
    from vit_som_b200 import SOMLayer, ops
    layer = SOMLayer(make_config([3, 3], 8))
    made = []
    real = ops.Staging

    class FakeStaging:
        def __init__(self, rows, dim, mode, device):
            self.rows, self.dim, self.mode, self.key, self.from_optimizer = rows, dim, mode, None, False
            made.append(self)
    ops.Staging = FakeStaging
    try:
        layer.train()
        a, fill_a = layer._staged_prototypes(0)
        b, fill_b = layer._staged_prototypes(0)
        assert fill_a and fill_b and a is not b                       # training: staging is part of every step
        layer.eval()
        c, fill_c = layer._staged_prototypes(0)
        d, fill_d = layer._staged_prototypes(0)
        assert not fill_c and c is b and d is b and not fill_d        # eval: cached while the parameter is unchanged
        with torch.no_grad():
            layer.prototypes.add_(1.0)                                # in-place op on the parameter: version moves
        e, fill_e = layer._staged_prototypes(0)
        assert fill_e and e is not b
        layer.prototypes.data.mul_(0.5)                               # .data writes do not move the version ...
        f, fill_f = layer._staged_prototypes(0)
        assert not fill_f and f is e
        layer.invalidate_staging()                                    # ... which is what invalidate_staging() is for
        g, fill_g = layer._staged_prototypes(0)
        assert fill_g and g is not e
        g.from_optimizer = True                                       # staging produced by the fused optimizer: trusted
        layer.train()
        h, fill_h = layer._staged_prototypes(0)
        assert h is g and not fill_h
    finally:
        ops.Staging = real
