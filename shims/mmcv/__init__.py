"""Minimal stand-in for the parts of mmcv 1.x that the reference's entry point touches (see shims/README.md)."""
import os
import os.path as osp
import pprint
import time

from . import utils  # noqa: F401


class ConfigDict(dict):
    """dict with attribute access (mmcv.utils.config.ConfigDict): cfg.optim_cfg['core_module'].type etc."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def copy(self):
        return ConfigDict(dict.copy(self))


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


def _merge(base, over):
    """Recursive dict merge, the override wins (mmcv Config._merge_a_into_b without the _delete_ key)."""
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load_py(path):
    path = osp.abspath(osp.expanduser(path))
    with open(path) as f:
        src = f.read()
    scope = {"__file__": path}
    exec(compile(src, path, "exec"), scope)
    cfg = {k: v for k, v in scope.items()
           if not k.startswith("__") and not isinstance(v, type(os)) and not callable(v)}
    base = cfg.pop("_base_", None)
    if base is not None:
        merged = {}
        for b in ([base] if isinstance(base, str) else list(base)):
            b = b if osp.isabs(b) else osp.join(osp.dirname(path), b)
            merged = _merge(merged, _load_py(b))
        cfg = _merge(merged, cfg)
    return cfg


class Config:
    """cfg = Config.fromfile('configs/unlg_former.py'): keys as attributes, `in`, get / setdefault / copy of sub-dicts."""

    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict or {}))
        object.__setattr__(self, "_filename", filename)

    @staticmethod
    def fromfile(filename):
        return Config(_load_py(filename), filename=filename)

    @property
    def filename(self):
        return self._filename

    @property
    def pretty_text(self):
        return pprint.pformat(dict(self._cfg_dict), width=120)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setitem__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def setdefault(self, key, default=None):
        return self._cfg_dict.setdefault(key, _wrap(default))

    def __repr__(self):
        return f"Config (path: {self._filename}): {dict(self._cfg_dict)!r}"


def mkdir_or_exist(dir_name, mode=0o777):
    if dir_name == "":
        return
    os.makedirs(osp.expanduser(dir_name), mode=mode, exist_ok=True)


class Timer:
    """mmcv.Timer: since_start() / since_last_check()."""

    def __init__(self, start=True):
        self._is_running = False
        if start:
            self.start()

    def start(self):
        if not self._is_running:
            self._t_start = time.time()
            self._is_running = True
        self._t_last = time.time()

    def since_start(self):
        if not self._is_running:
            raise RuntimeError("timer is not running")
        self._t_last = time.time()
        return self._t_last - self._t_start

    def since_last_check(self):
        if not self._is_running:
            raise RuntimeError("timer is not running")
        dur = time.time() - self._t_last
        self._t_last = time.time()
        return dur
