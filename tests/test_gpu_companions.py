"""GPU parity of the companion operators (run with `-m gpu` on a B200) through the C ABI
(lgteun_op_freprocess) behind lgteun_b200.Freprocess: against the recorded outputs of the unmodified reference class
(tests/golden/freprocess_c8.npz) and against the CPU oracle on other seeded shapes and channel counts.
Tolerance: max |delta| <= 1e-3 x max(1, max|ref|) — the bound of the hot path (BASELINE north_star), scaled because the
un-normalised random-init output of this operator is O(10)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test collected without a CUDA device")
    import lgteun_b200
    lgteun_b200._abi.lib()
    return lgteun_b200


def _tol(ref):
    return 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", ["sq", "rect"])
def test_freprocess_matches_recorded_reference(lib, case):
    z = np.load(os.path.join(GOLDEN, "freprocess_c8.npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    net = lib.Freprocess(8)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(torch.from_numpy(z[f"{case}/msf"]).cuda(), torch.from_numpy(z[f"{case}/panf"]).cuda()).cpu()
    ref = torch.from_numpy(z[f"{case}/out"])
    err = (out - ref).abs().max().item()
    print(f"freprocess {case}: max|delta| {err:.3e} (max|ref| {ref.abs().max().item():.3f})")
    assert err <= _tol(ref)


@pytest.mark.parametrize("channels,n,h,w", [(8, 3, 128, 128), (4, 2, 8, 16), (16, 1, 64, 32), (8, 1, 256, 256), (8, 2, 256, 64),
                                               (4, 1, 32, 128), (16, 1, 128, 256), (4, 2, 256, 128), (16, 1, 256, 128), (8, 2, 128, 256)])
def test_freprocess_matches_oracle(lib, channels, n, h, w):
    from oracle import companions_oracle as CO
    torch.manual_seed(100 + channels + h)
    net = lib.Freprocess(channels)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(h * w + channels)
    msf, panf = torch.rand(n, channels, h, w, generator=g), torch.rand(n, channels, h, w, generator=g)
    ref = CO.freprocess_forward(sd, msf, panf)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(msf.cuda(), panf.cuda())
        out2 = net(msf.cuda(), panf.cuda())
    assert torch.equal(out, out2)                              # deterministic, inputs untouched
    err = (out.cpu() - ref).abs().max().item()
    print(f"freprocess C={channels} {n}x{h}x{w}: max|delta| {err:.3e} (max|ref| {ref.abs().max().item():.3f})")
    assert err <= _tol(ref)


def test_freprocess_rejects_what_it_does_not_support(lib):
    net = lib.Freprocess(8).cuda().eval()
    with torch.no_grad():
        with pytest.raises(ValueError):
            net(torch.zeros(1, 8, 24, 24, device="cuda"), torch.zeros(1, 8, 24, 24, device="cuda"))   # not a power of two
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 8, 16, 16, device="cuda"), torch.zeros(1, 8, 16, 16, device="cuda"))       # autograd on


# ---- PanFormer WindowAttention (models/common/modules.py:341-422) ------------------------------------------------------------
WINATT_CASES = {   # mirrors tests/golden/make_golden_companions.py
    "regular": (False, False, True, 4, 16, 64),
    "shifted": (True, False, True, 4, 16, 64),
    "cross": (False, True, True, 4, 16, 64),
    "cross_shifted": (True, True, True, 4, 16, 64),
    "dense_pos": (True, False, False, 2, 8, 32),
}


@pytest.mark.parametrize("case", sorted(WINATT_CASES))
def test_window_attention_matches_recorded_reference(lib, case):
    shifted, cross, rel, heads, hd, dim = WINATT_CASES[case]
    z = np.load(os.path.join(GOLDEN, "window_attention.npz"))
    pre = f"{case}/w/"
    sd = {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
    net = lib.WindowAttention(dim=dim, heads=heads, head_dim=hd, shifted=shifted, window_size=4, relative_pos_embedding=rel,
                              cross_attn=cross)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    x = torch.from_numpy(z[f"{case}/x"]).cuda()
    with torch.no_grad():
        out = net(x, torch.from_numpy(z[f"{case}/y"]).cuda()) if cross else net(x)
    ref = torch.from_numpy(z[f"{case}/out"])
    err = (out.cpu() - ref).abs().max().item()
    print(f"window attention {case}: max|delta| {err:.3e} (max|ref| {ref.abs().max().item():.3f})")
    assert err <= 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("shifted,cross,heads,hd,dim,b,nh,nw", [(True, False, 4, 16, 64, 3, 64, 64), (False, True, 2, 32, 96, 1, 32, 16),
                                                                (True, True, 3, 8, 20, 2, 4, 4), (True, False, 4, 16, 64, 1, 128, 128)])
def test_window_attention_matches_oracle(lib, shifted, cross, heads, hd, dim, b, nh, nw):
    from oracle import companions_oracle as CO
    torch.manual_seed(7 + dim + nh)
    net = lib.WindowAttention(dim=dim, heads=heads, head_dim=hd, shifted=shifted, window_size=4, relative_pos_embedding=True,
                              cross_attn=cross)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(nh * nw)
    x = torch.randn(b, nh, nw, dim, generator=g)
    y = torch.randn(b, nh, nw, dim, generator=g) if cross else None
    ref = CO.window_attention_forward(sd, x, y, heads, hd, 4, shifted, True)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(x.cuda(), y.cuda()) if cross else net(x.cuda())
    err = (out.cpu() - ref).abs().max().item()
    print(f"window attention dim={dim} heads={heads}x{hd} {b}x{nh}x{nw}: max|delta| {err:.3e}")
    assert err <= 1e-4 * max(1.0, ref.abs().max().item())


def test_window_attention_rejects_what_it_does_not_support(lib):
    net = lib.WindowAttention(dim=64, heads=4, head_dim=16, shifted=False, window_size=8, relative_pos_embedding=True,
                              cross_attn=False).cuda().eval()
    with torch.no_grad(), pytest.raises(ValueError):
        net(torch.zeros(1, 16, 16, 64, device="cuda"))                   # window 8 is not built
    net = lib.WindowAttention(dim=64, heads=4, head_dim=16, shifted=False, window_size=4, relative_pos_embedding=True,
                              cross_attn=False).cuda().eval()
    with torch.no_grad(), pytest.raises(ValueError):
        net(torch.zeros(1, 10, 16, 64, device="cuda"))                   # 10 is not a multiple of the window
