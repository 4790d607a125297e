"""Import the UNMODIFIED reference (read-only at /root/reference) inside the build container.

Only used by tests/golden/make_golden.py and by container-only tests that are skipped when
/root/reference is absent (it does not exist on the GPU box).  Nothing is copied: the reference
files are imported from where they lie, behind stubs for the modules this image lacks
(mmcv, gdal, osr, tifffile) and a package shim that skips models/__init__.py (which would import
every other network and pywt).  TEST INFRASTRUCTURE ONLY."""
import importlib
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("LGTEUN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "unlg_former.py"))


class _Registry:
    def __init__(self, name):
        self.name, self._d = name, {}

    def register_module(self, name=None):
        def deco(cls):
            self._d[name or cls.__name__] = cls
            return cls
        return deco

    def __contains__(self, k):
        return k in self._d

    def get(self, k):
        return self._d.get(k)


class Config(dict):
    __getattr__ = dict.get


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns (unlg_former module, LGT module, metrics module)."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if "mmcv" not in sys.modules:
        mm = _stub("mmcv", Config=Config, mkdir_or_exist=lambda p: None, Timer=object)
        mm.utils = _stub("mmcv.utils", Registry=_Registry,
                         get_logger=lambda *a, **k: logging.getLogger("ref"))
    for n in ("gdal", "osr", "tifffile"):
        sys.modules.setdefault(n, types.ModuleType(n))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "models" not in sys.modules or not hasattr(sys.modules["models"], "__path__"):
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REF_ROOT, "models")]
        sys.modules["models"] = pkg
    ul = importlib.import_module("models.unlg_former")
    lgt = importlib.import_module("models.common.LGT")
    mtc = importlib.import_module("models.base.metrics")
    return ul, lgt, mtc


def build(bands: int, stages: int = 2, seed: int = 19971118):
    """Reference Pansharpening with its default init under the config's seed (configs/unlg_former.py:66)."""
    import torch
    ul, _, _ = load()
    torch.manual_seed(seed)
    return ul.Pansharpening(Config(ms_chans=bands), None, stage=stages).eval()
