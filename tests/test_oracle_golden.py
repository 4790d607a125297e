"""Pin the oracle restatement against the committed outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from conftest import load_case, load_weights
from oracle import lgteun_oracle as O
from oracle import metrics_oracle as M

# fp32 summation order differs between 1 and 4 CPU threads (up to 2.7e-4 end to end);
# the goldens were generated single-threaded.
TOL_FWD = 5e-4
PRIOR = "prior_module.1"
BLK0 = PRIOR + ".encoder_layers.0.0.blocks.0"


@pytest.mark.parametrize("name,bands", [("gf2_small", 4), ("wv3_small", 8), ("gf2_batch", 4), ("gf2_rect", 4),
                                        ("gf2_metric", 4)])
def test_forward_matches_reference_output(name, bands):
    sd, case = load_weights(bands), load_case(name)
    out = O.forward(sd, case["ms"], case["pan"])
    assert out.shape == case["out"].shape
    assert (out - case["out"]).abs().max().item() <= TOL_FWD


def test_forward_single_thread_is_bit_exact(weights4):
    case = load_case("gf2_small")
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        out = O.forward(weights4, case["ms"], case["pan"])
        both = O.forward(weights4, case["ms"], case["pan"], skip_dead_priors=False)
    finally:
        torch.set_num_threads(n)
    assert torch.equal(out, case["out"])
    assert torch.equal(both, case["out"])          # dead prior does not change the result


def test_full_size_config0(weights4):
    """BASELINE.json configs[0]: GF-2 shape PAN 256x256 + LrMS 64x64x4, batch 1."""
    case = load_case("gf2_full")
    out = O.forward(weights4, case["ms"], case["pan"])
    assert (out - case["out"]).abs().max().item() <= TOL_FWD


def test_per_operator_goldens(weights4):
    sd, g = weights4, load_case("gf2_small")
    tol = 2e-5
    trace = {}
    O.forward(sd, g["ms"], g["pan"], trace=trace)
    assert (trace["z1"] - g["prior_in"]).abs().max() <= tol                       # two data steps
    pe = O.patch_embed(sd, PRIOR + ".patch_embed", g["prior_in"])
    assert (pe - g["pe"]).abs().max() <= tol
    loc = O.local_mixer(sd, BLK0 + ".0.fn.fn.local_mixer", g["enc0_local_in"])
    b, h, w, c2 = g["enc0_local_in"].shape
    ref_loc = g["enc0_local"].reshape(b, h // 8, w // 8, 8, 8, c2).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c2)
    assert (loc - ref_loc).abs().max() <= tol
    glo = O.global_mixer(sd, BLK0 + ".0.fn.fn.global_mixer", g["enc0_global_in"])
    assert (glo - g["enc0_global"]).abs().max() <= tol
    mix = O.lg_mixer(sd, BLK0 + ".0.fn.fn", O.layer_norm(sd, BLK0 + ".0.fn.norm", g["pe"])) + g["pe"]
    assert (mix - g["enc0_mixer"]).abs().max() <= tol
    ffn = O.feed_forward(sd, BLK0 + ".1.fn.fn", O.layer_norm(sd, BLK0 + ".1.fn.norm", g["enc0_mixer"])) + g["enc0_mixer"]
    assert (ffn - g["enc0_ffn"]).abs().max() <= tol
    enc = O.lgb(sd, PRIOR + ".encoder_layers.0.0", g["pe"], 2)
    assert (enc - g["enc"]).abs().max() <= 1e-4
    down = O.pconv(sd, PRIOR + ".encoder_layers.0.1.1", O.bicubic(g["enc"], 0.5))
    assert (down - g["down"]).abs().max() <= tol
    bott = O.lgb(sd, PRIOR + ".bottleneck", g["down"].permute(0, 2, 3, 1), 1)
    assert (bott - g["bott"]).abs().max() <= 1e-4
    up = O.pconv(sd, PRIOR + ".decoder_layers.0.0.1", O.bicubic(g["bott"], 2))
    assert (up - g["up"]).abs().max() <= tol
    fuse = O.pconv(sd, PRIOR + ".decoder_layers.0.1", torch.cat([g["up"], g["enc"]], 1))
    assert (fuse - g["fuse"]).abs().max() <= tol
    dec = O.lgb(sd, PRIOR + ".decoder_layers.0.2", g["fuse"].permute(0, 2, 3, 1), 2)
    assert (dec - g["dec"]).abs().max() <= 1e-4
    out = O.pconv(sd, PRIOR + ".tail.1", O.bicubic(g["dec"], 1)) + g["prior_in"]
    assert (out - g["out"]).abs().max() <= tol


def test_metrics_match_reference_functions(weights4):
    g = load_case("gf2_metric")
    ours = M.evaluate(g["out"].numpy(), g["gt"].numpy())
    np.testing.assert_allclose(ours, g["ref_metrics_psnr_sam_ergas"].numpy(), rtol=0, atol=1e-9)


def test_param_counts_match_paper(weights4, weights8):
    """paper Table 4: 202.2 K (4-band) / 540.0 K (8-band) parameters at K=2."""
    assert sum(v.numel() for v in weights4.values()) == 202183
    assert sum(v.numel() for v in weights8.values()) == 540043
    assert len(weights4) == 252 and O.num_stages(weights4) == 2
