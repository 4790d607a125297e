"""CPU: the companion-operator oracle (oracle/companions_oracle.py) against the golden vectors recorded from the unmodified
reference class (tests/golden/make_golden_companions.py), and the host-side mirror's interface."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _golden():
    z = np.load(os.path.join(GOLDEN, "freprocess_c8.npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    return z, sd


@pytest.mark.parametrize("case", ["sq", "rect"])
def test_freprocess_oracle_matches_recorded_reference(case):
    from oracle import companions_oracle as CO
    z, sd = _golden()
    torch.set_num_threads(1)
    out = CO.freprocess_forward(sd, torch.from_numpy(z[f"{case}/msf"]), torch.from_numpy(z[f"{case}/panf"]))
    ref = torch.from_numpy(z[f"{case}/out"])
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())


def test_freprocess_module_mirrors_the_reference_interface():
    import lgteun_b200
    from lgteun_b200.companions import FREPROCESS_KEYS
    _, sd = _golden()
    torch.manual_seed(19971118)
    net = lgteun_b200.Freprocess(8)
    own = net.state_dict()
    assert list(own) == list(sd) == list(FREPROCESS_KEYS)          # same keys, same order as the reference's state_dict
    for k in sd:
        assert own[k].shape == sd[k].shape, k
        assert torch.equal(own[k], sd[k]), k                       # same construction order => same default init under the seed
    net.load_state_dict(sd)
    x = torch.zeros(1, 8, 16, 16)
    with pytest.raises(RuntimeError):                              # no CPU fallback
        net(x, x)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 4, 16, 16), torch.zeros(1, 4, 16, 16))


def test_freprocess_golden_matches_the_reference_when_present():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference checkout not present (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_companions", os.path.join(GOLDEN, "make_golden_companions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sf = mod.load_sfiin()
    z, sd = _golden()
    net = sf.Freprocess(8).eval()
    net.load_state_dict(sd)
    torch.set_num_threads(1)
    with torch.no_grad():
        y = net(torch.from_numpy(z["sq/msf"]), torch.from_numpy(z["sq/panf"]))
    assert torch.equal(y, torch.from_numpy(z["sq/out"]))


# ---- PanFormer WindowAttention (models/common/modules.py:341-422) ------------------------------------------------------------
WINATT_CASES = {   # mirrors tests/golden/make_golden_companions.py
    "regular": (False, False, True, 4, 16, 64),
    "shifted": (True, False, True, 4, 16, 64),
    "cross": (False, True, True, 4, 16, 64),
    "cross_shifted": (True, True, True, 4, 16, 64),
    "dense_pos": (True, False, False, 2, 8, 32),
}


def _winatt(case):
    z = np.load(os.path.join(GOLDEN, "window_attention.npz"))
    pre = f"{case}/w/"
    sd = {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
    x = torch.from_numpy(z[f"{case}/x"])
    y = torch.from_numpy(z[f"{case}/y"]) if f"{case}/y" in z.files else None
    return sd, x, y, torch.from_numpy(z[f"{case}/out"])


@pytest.mark.parametrize("case", sorted(WINATT_CASES))
def test_window_attention_oracle_matches_recorded_reference(case):
    from oracle import companions_oracle as CO
    shifted, cross, rel, heads, hd, dim = WINATT_CASES[case]
    sd, x, y, ref = _winatt(case)
    torch.set_num_threads(1)
    out = CO.window_attention_forward(sd, x, y, heads, hd, 4, shifted, rel)
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", sorted(WINATT_CASES))
def test_window_attention_module_mirrors_the_reference_interface(case):
    import lgteun_b200
    shifted, cross, rel, heads, hd, dim = WINATT_CASES[case]
    sd, x, y, _ = _winatt(case)
    torch.manual_seed(19971118)
    net = lgteun_b200.WindowAttention(dim=dim, heads=heads, head_dim=hd, shifted=shifted, window_size=4,
                                      relative_pos_embedding=rel, cross_attn=cross)
    own = net.state_dict()
    assert list(own) == list(sd)                                   # same keys in the same order
    for k in sd:
        assert own[k].shape == sd[k].shape, k
        assert torch.equal(own[k], sd[k]), k                       # same default init under the seed, same -inf masks
    net.load_state_dict(sd)
    with pytest.raises(RuntimeError):                              # no CPU fallback
        net(x, y) if cross else net(x)
    with pytest.raises(ValueError):
        net(x, None if cross else x)


def test_swap_companions_inside_unmodified_reference_networks():
    """lgteun_b200.swap_companions on the reference's own SpaFre (models/SFIIN.py:240-273) and SwinModule
    (models/common/modules.py:458-505): the operators are replaced, the networks' state_dicts do not change, and the
    reference forward now reaches the lgteun_b200 operators (which refuse CPU tensors: there is no fallback)."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference checkout not present (GPU box)")
    import importlib.util
    import lgteun_b200
    spec = importlib.util.spec_from_file_location("make_golden_companions", os.path.join(GOLDEN, "make_golden_companions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sf, md = mod.load_sfiin(), mod.load_modules()

    torch.manual_seed(3)
    spafre = sf.SpaFre(8).eval()
    before = {k: v.clone() for k, v in spafre.state_dict().items()}
    assert lgteun_b200.swap_companions(spafre) == {"Freprocess": 1, "WindowAttention": 0}
    assert isinstance(spafre.fre_process, lgteun_b200.Freprocess) and not spafre.fre_process.training
    after = spafre.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        spafre(torch.zeros(1, 8, 16, 16), torch.zeros(1, 8, 16, 16))

    swin = md.SwinModule(in_channels=64, hidden_dimension=64, layers=2, downscaling_factor=1, num_heads=4, head_dim=16,
                         window_size=4, relative_pos_embedding=True, cross_attn=True).eval()     # models/panformer.py:51-53
    before = {k: v.clone() for k, v in swin.state_dict().items()}
    assert lgteun_b200.swap_companions(swin) == {"Freprocess": 0, "WindowAttention": 2}          # regular + shifted block
    kinds = [type(m).__module__ for m in swin.modules() if type(m).__name__ == "WindowAttention"]
    assert kinds == ["lgteun_b200.companions"] * 2
    after = swin.state_dict()
    assert list(after) == list(before) and all(torch.equal(after[k], before[k]) for k in before)
    assert lgteun_b200.swap_companions(swin) == {"Freprocess": 0, "WindowAttention": 0}          # idempotent
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        swin(torch.zeros(1, 64, 8, 8), torch.zeros(1, 64, 8, 8))

    big = md.WindowAttention(dim=64, heads=4, head_dim=16, shifted=False, window_size=8, relative_pos_embedding=True,
                             cross_attn=False)
    holder = torch.nn.Sequential(big)
    assert lgteun_b200.swap_companions(holder) == {"Freprocess": 0, "WindowAttention": 0}        # window 8 is left alone
    assert holder[0] is big
