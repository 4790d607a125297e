"""The explicit restatements (what the CUDA kernels implement) agree with the torch-op oracle."""
import torch

from conftest import load_case
from oracle import lgteun_oracle as O

BLK0 = "prior_module.1.encoder_layers.0.0.blocks.0"


def test_cubic_taps_are_dyadic():
    assert O.cubic_taps(0.5) == [-3 / 32, 19 / 32, 19 / 32, -3 / 32]
    assert O.cubic_taps(0.75) == [-0.03515625, 0.26171875, 0.87890625, -0.10546875]
    assert O.cubic_taps(0.25) == [-0.10546875, 0.87890625, 0.26171875, -0.03515625]
    assert O.cubic_taps(0.0) == [0.0, 1.0, 0.0, 0.0]       # scale 1 is the identity (LGT.py:302 tail)
    for t in (0.125, 0.375, 0.625, 0.875):
        assert abs(sum(O.cubic_taps(t)) - 1.0) < 1e-15


def test_bicubic_explicit_all_scales():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 16, 32, generator=g, dtype=torch.float64)
    for s in (4, 2, 0.5, 1):
        a, b = O.bicubic(x, s), O.bicubic_explicit(x, s)
        assert a.shape == b.shape
        assert (a - b).abs().max() < 1e-14


def test_local_mixer_explicit(weights4):
    g = load_case("gf2_small")
    x = g["enc0_local_in"][:, :16, :24].contiguous()
    p = BLK0 + ".0.fn.fn.local_mixer"
    assert (O.local_mixer(weights4, p, x) - O.local_mixer_explicit(weights4, p, x)).abs().max() < 2e-6


def test_global_mixer_three_pass(weights4):
    g = load_case("gf2_small")
    x = g["enc0_global_in"]
    p = BLK0 + ".0.fn.fn.global_mixer"
    assert (O.global_mixer(weights4, p, x) - O.global_mixer_3pass(weights4, p, x)).abs().max() < 2e-5


def test_real_bins_have_positive_zero_imag():
    """the oracle's rfft2 yields exactly +0.0 imaginary parts at the four real bins."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 8, 64, 64, generator=g)
    f = torch.fft.rfft2(x)
    for ky in (0, 32):
        for kx in (0, 32):
            im = f.imag[..., ky, kx]
            assert torch.all(im == 0) and not torch.any(torch.signbit(im))


def test_gelu_is_erf_form():
    x = torch.linspace(-6, 6, 1001)
    assert torch.equal(O.gelu_erf(x).float(), O.gelu_erf(x)) and (O.gelu_erf(x) - torch.nn.functional.gelu(x)).abs().max() < 1e-6


def test_exact_real_bins_switch_is_a_no_op_at_powers_of_two_and_stable_elsewhere():
    """oracle.EXACT_REAL_BINS (imaginary part of the four purely real rfft2 bins := +0.0): identical to the literal
    restatement at power-of-two sizes, where torch already returns +0.0; at column lengths such as 72 or 24 the literal
    fp32 restatement is off by 1e-3 .. 3e-2 from its own fp64 evaluation (arbitrary sign of a rounding residue under
    angle()) while the exact-bin form agrees between fp32 and fp64 to fp32 rounding."""
    import torch
    from conftest import load_weights
    from oracle import lgteun_oracle as O
    sd = load_weights(4)
    sd64 = {k: v.double() for k, v in sd.items()}
    pre = "prior_module.1.encoder_layers.0.0.blocks.0.0.fn.fn.global_mixer"
    x = torch.randn(1, 64, 32, 8, generator=torch.Generator().manual_seed(3))
    lit = O.global_mixer(sd, pre, x)
    O.EXACT_REAL_BINS = True
    try:
        assert (O.global_mixer(sd, pre, x) - lit).abs().max().item() <= 2e-5
        for H in (72, 24):
            xs = torch.randn(1, H, 64, 8, generator=torch.Generator().manual_seed(3))
            a = O.global_mixer(sd, pre, xs).double()
            b = O.global_mixer(sd64, pre, xs.double())
            assert (a - b).abs().max().item() <= 1e-4
    finally:
        O.EXACT_REAL_BINS = False
