"""Plug the B200 forward behind the reference's mmcv-style registry/config.

The reference builds its runner with ``build_model(cfg.model_type, cfg, logger, loaders...)``
(models/base/builder.py:17-24, main.py:90) -> ``MODELS.get('UnlgFormer')`` ->
``UnlgFormer.__init__`` which instantiates ``Pansharpening(cfg=cfg, logger=logger, **G_cfg)`` looked up in
the namespace of ``models.unlg_former`` (models/unlg_former.py:70-78).  ``install()`` rebinds that one
name, so ``main.py`` and ``configs/unlg_former.py`` run unchanged and every other code path of the runner
(set_cuda / load_checkpoint / test, models/base/base_model.py) keeps working on an ordinary nn.Module with
the same state_dict."""
import importlib
import sys

from .module import Pansharpening


def install(reference_module: str = "models.unlg_former"):
    """Rebind ``Pansharpening`` inside the (already importable) reference module. Returns that module."""
    mod = sys.modules.get(reference_module) or importlib.import_module(reference_module)
    if getattr(mod, "Pansharpening", None) is not Pansharpening:
        mod._reference_Pansharpening = getattr(mod, "Pansharpening", None)
        mod.Pansharpening = Pansharpening
    return mod


def uninstall(reference_module: str = "models.unlg_former"):
    mod = sys.modules.get(reference_module)
    if mod is not None and getattr(mod, "_reference_Pansharpening", None) is not None:
        mod.Pansharpening = mod._reference_Pansharpening
        del mod._reference_Pansharpening
