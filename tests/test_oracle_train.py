"""CPU: the oracle's training step (oracle.train_step_grads / adam_step) against recorded steps of the unmodified
reference (tests/golden/train_gf2.npz: 4 bands, train_wv3.npz: 8 bands, made by tests/golden/make_golden_train.py from
/root/reference): train() mode with the recorded nn.Dropout masks, L1 loss, loss.backward(), Adam(lr=1.5e-3).  Pins the checker
the GPU training tests rely on."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_weights
from oracle import lgteun_oracle as O


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def golden():
    return _load("train_gf2.npz")


def _masks(g):
    return [torch.from_numpy(g[f"mask{i}"].astype(np.float32)) / 0.9 for i in range(5)]


def test_train_step_matches_reference(golden):
    torch.set_num_threads(1)
    sd = load_weights(4)
    ms, pan, gt = (torch.from_numpy(golden[k]) for k in ("ms", "pan", "gt"))
    out, loss, grads = O.train_step_grads(sd, ms, pan, gt, _masks(golden))
    assert (out - torch.from_numpy(golden["out"])).abs().max().item() <= 1e-5
    assert abs(loss.item() - float(golden["loss"])) <= 1e-6
    live = {k[5:] for k in golden if k.startswith("grad/")}
    assert live == {k for k, g in grads.items() if g is not None}
    assert len(live) == 133                      # 14 shared + 119 of the last prior
    assert all(k.startswith("prior_module.0.") for k, g in grads.items() if g is None)
    for k in sorted(live):
        ref = torch.from_numpy(golden["grad/" + k])
        tol = 1e-5 * max(1.0, ref.abs().max().item())
        assert (grads[k] - ref).abs().max().item() <= tol, k


def test_train_step_matches_reference_8_bands():
    """The same pin at 8 bands (BASELINE configs[4]: WV-2 / WV-3 band count), batch 1."""
    g8 = _load("train_wv3.npz")
    torch.set_num_threads(1)
    sd = load_weights(8)
    ms, pan, gt = (torch.from_numpy(g8[k]) for k in ("ms", "pan", "gt"))
    assert ms.shape == (1, 8, 8, 8)
    out, loss, grads = O.train_step_grads(sd, ms, pan, gt, _masks(g8))
    assert (out - torch.from_numpy(g8["out"])).abs().max().item() <= 1e-5
    assert abs(loss.item() - float(g8["loss"])) <= 1e-6
    live = {k[5:] for k in g8 if k.startswith("grad/")}
    assert live == {k for k, g in grads.items() if g is not None} and len(live) == 133
    for k in sorted(live):
        ref = torch.from_numpy(g8["grad/" + k])
        assert (grads[k] - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), k
    for k in (kk[6:] for kk in g8 if kk.startswith("after/")):
        g = torch.from_numpy(g8["grad/" + k])
        p, _, _ = O.adam_step(sd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, 1.5e-3)
        assert (p - torch.from_numpy(g8["after/" + k])).abs().max().item() <= 1e-7, k


def test_adam_step_matches_reference(golden):
    sd = load_weights(4)
    for k in ("eta.1", "R.weight", "prior_module.1.tail.1.weight", "prior_module.1.bottleneck.blocks.0.0.fn.fn.local_mixer.pos_emb"):
        g = torch.from_numpy(golden["grad/" + k])
        p, _, _ = O.adam_step(sd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, 1.5e-3)
        assert (p - torch.from_numpy(golden["after/" + k])).abs().max().item() <= 1e-7, k


def test_dropout_masks_are_the_live_prior_only(golden):
    shapes = [golden[f"mask{i}"].shape for i in range(5)]
    assert shapes == [(2, 32, 32, 16)] * 2 + [(2, 16, 16, 32)] + [(2, 32, 32, 16)] * 2
    keep = np.mean([golden[f"mask{i}"].mean() for i in range(5)])
    assert 0.88 < keep < 0.92                    # nn.Dropout(0.1)
