"""Host-side glue of SURVEY section 8f rank 1 (device-side gamma ramp, copy-free SOM input) against the reference's
expressions (models/vit_som.py:69-73,88-90).  Pure tensor arithmetic: runs on the CPU."""
import torch

from vit_som_b200 import gamma_ramp, som_input


def test_gamma_ramp_matches_the_reference_expression():
    gamma, end = 0.37, 1250
    for it in [0, 1, 17, 624, 625, 1249, 1250, 1251, 10 ** 6]:
        ref = gamma * min(1.0, it / end)                               # vit_som.py:90 with iteration.item()
        assert gamma_ramp(it, end, gamma) == ref                       # python number in, python float out
        t = gamma_ramp(torch.tensor(it), end, gamma)                   # the module's int64 buffer in, tensor out
        assert torch.is_tensor(t) and t.dim() == 0 and t.dtype == torch.float32
        assert abs(t.item() - ref) <= 1e-6 * max(ref, 1e-30) + 1e-12


def test_gamma_ramp_keeps_the_graph_differentiable():
    som_loss = torch.tensor(2.5, requires_grad=True)
    total = gamma_ramp(torch.tensor(300), 1000, 0.5) * som_loss
    total.backward()
    assert abs(som_loss.grad.item() - 0.15) < 1e-7


def test_som_input_is_a_view_of_the_encoder_output():
    B, N, E = 4, 9, 16
    enc = torch.randn(B, N + 1, E)
    cls, patches = enc[:, 0], enc[:, 1:]                               # models/vit.py:221-222
    z = som_input(cls, patches, use_reduced=False)
    assert z.shape == (B, N * E) and z.stride() == ((N + 1) * E, 1)
    assert z.data_ptr() == patches.data_ptr()                          # no copy
    assert torch.equal(z, patches.flatten(start_dim=1))                # same values as vit_som.py:73
    assert som_input(cls, patches, use_reduced=True) is cls
    weird = patches.transpose(1, 2)                                    # not mergeable: falls back to flatten
    assert torch.equal(som_input(cls, weird, False), weird.flatten(start_dim=1))
