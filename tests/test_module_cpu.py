"""Host-side mirror of the reference interface (no GPU needed): constructor, state_dict ABI, init parity,
pickling, registry hook, loud failure without CUDA."""
import io
from types import SimpleNamespace

import pytest
import torch

import lgteun_b200
from conftest import load_weights
from oracle import ref_import


@pytest.mark.parametrize("bands,count", [(4, 202183), (8, 540043)])
def test_state_dict_abi(bands, count):
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=bands), None, stage=2)
    sd = net.state_dict()
    golden = load_weights(bands)
    assert list(sd.keys()) == list(golden.keys())                       # same keys, same order as the reference
    assert all(sd[k].shape == golden[k].shape for k in golden)
    assert sorted(sd.keys()) == sorted(lgteun_b200.expected_state_dict_keys(bands, 2))
    assert sum(p.numel() for p in net.parameters()) == count == lgteun_b200.param_count(bands, 2)
    net.load_state_dict(golden)                                         # strict


def test_default_stage_and_ctor_signature():
    net = lgteun_b200.Pansharpening(cfg=SimpleNamespace(ms_chans=4), logger=None)      # stage defaults to 5
    assert net.stage == 5 and len(net.prior_module) == 5 and len(net.eta) == 5
    with pytest.raises(ValueError):
        lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=3), None, stage=2)


def test_same_seed_same_init_as_golden_weights():
    """Golden weights were produced by the reference under seed 19971118; the mirror reproduces them exactly."""
    torch.manual_seed(19971118)
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=4), None, stage=2)
    golden = load_weights(4)
    assert all(torch.equal(v, golden[k]) for k, v in net.state_dict().items())


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_init_matches_reference_class_and_registry_hook():
    ul, _, _ = ref_import.load()
    ref = ref_import.build(8, stages=2, seed=7)
    torch.manual_seed(7)
    ours = lgteun_b200.Pansharpening(ref_import.Config(ms_chans=8), None, stage=2)
    a, b = ours.state_dict(), ref.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in b)
    # registry/builder path: UnlgFormer.__init__ looks Pansharpening up in models.unlg_former (unlg_former.py:70-78)
    assert "UnlgFormer" in ul.MODELS
    mod = lgteun_b200.install()
    try:
        assert mod is ul and ul.Pansharpening is lgteun_b200.Pansharpening
        core = ul.Pansharpening(cfg=ref_import.Config(ms_chans=4), logger=None, **dict(stage=2))
        assert isinstance(core, torch.nn.Module) and core.stage == 2
        core.load_state_dict(ref_import.build(4).state_dict())          # checkpoint path, base_model.py:102-114
    finally:
        from lgteun_b200.register import uninstall
        uninstall()
    assert ul.Pansharpening is not lgteun_b200.Pansharpening


def test_pickle_roundtrip_drops_runtime():
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=4), None, stage=2)
    net._rt[0] = {"handle": object(), "sig": None}
    buf = io.BytesIO()
    torch.save(net, buf)                                                # base_model.py:362-368 pickles whole modules
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert back._rt == {} and list(back.state_dict()) == list(net.state_dict())


def test_no_cpu_fallback():
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=4), None, stage=2).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        net(torch.rand(1, 4, 16, 16), torch.rand(1, 1, 64, 64))
    with pytest.raises(ValueError):
        net(torch.rand(1, 4, 16, 16), torch.rand(1, 1, 32, 32))


def test_product_never_imports_the_oracle():
    import os
    import re
    from conftest import ROOT
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lgteun_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_weight_collection_inside_a_dataparallel_replica():
    """The reference wraps the module in nn.DataParallel when several GPUs are visible (models/base/base_model.py:90-96).
    A replica has empty `_parameters` and carries the broadcast copies as plain attributes; the weight table handed to the
    C ABI must still be complete there.  The replica is built here the way torch/nn/parallel/replicate.py builds it,
    without the CUDA broadcast."""
    from collections import OrderedDict
    from types import SimpleNamespace
    import torch
    import lgteun_b200

    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=4), None, stage=2)
    modules = list(net.modules())
    index = {m: i for i, m in enumerate(modules)}
    copies = []
    for m in modules:
        rep = m._replicate_for_data_parallel()
        rep._former_parameters = OrderedDict()
        copies.append(rep)
    for i, m in enumerate(modules):
        rep = copies[i]
        for key, child in m._modules.items():
            rep._modules[key] = None if child is None else copies[index[child]]
        for key, p in m._parameters.items():
            if p is None:
                rep._parameters[key] = None
                continue
            t = p.detach().clone()
            setattr(rep, key, t)
            rep._former_parameters[key] = t
    replica = copies[0]
    assert list(replica.parameters()) == [] and len(replica.state_dict()) == 0       # why parameters() cannot be used
    got = dict(replica._weight_items())
    want = net.state_dict()
    assert list(got) == list(want) and sorted(got) == sorted(lgteun_b200.expected_state_dict_keys(4, 2))
    assert all(torch.equal(got[k], want[k]) for k in want)
    assert [k for k, _ in net._weight_items()] == list(want)
