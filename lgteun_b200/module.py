"""Host-side mirror of the reference's hot-path module: a drop-in ``Pansharpening`` nn.Module.

Reference interface mirrored (paths relative to /root/reference):
  * ``Pansharpening(cfg, logger, stage=5)``           models/unlg_former.py:21-48   (reads only cfg.ms_chans)
  * ``forward(ms, pan) -> HrMS``                      models/unlg_former.py:50-67
  * ``state_dict()`` key grammar / shapes             SURVEY.md Appendix B (the weight ABI; checkpoints are
    loaded with ``module.load_state_dict(ckpt[name].state_dict())``, models/base/base_model.py:102-114)

The module owns ordinary ``nn.Parameter``s with the reference's names, shapes and default
initialisation (same construction order, hence the same values under the same seed), but it has no
PyTorch compute path: ``forward`` hands raw device pointers to the C ABI of ``include/lgteun.h``
(hand-written sm_100a kernels chained in one CUDA graph).  If the CUDA library is missing or the
inputs are not on a CUDA device it raises — there is no CPU / eager fallback."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import _abi

WINDOW = 8      # unlg_former.py:47
HEADS = 2       # unlg_former.py:48
UP_FACTOR = 4   # unlg_former.py:26


class _Slot(nn.Module):
    """Parameter-free placeholder that keeps the reference's child indices (e.g. the resize units at
    positions 0 and 2 of ``D``/``DT`` and position 0 of the down/up/tail Sequentials)."""

    def __init__(self, note: str = ""):
        super().__init__()
        self.note = note

    def extra_repr(self):
        return self.note


class _Node(nn.Module):
    """Plain named container: children are attached by key so state_dict paths match the reference."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.attach(k, v)

    def attach(self, name: str, child):
        if isinstance(child, nn.Parameter):
            self.register_parameter(name, child)
        else:
            self.add_module(name, child)
        return child

    def __getitem__(self, idx):
        return self._modules[str(idx)]

    def __len__(self):
        return len(self._modules)


def _seq(*mods):
    node = _Node()
    for i, m in enumerate(mods):
        node.attach(str(i), m)
    return node


def _pw(cin, cout):          # bmu.point_conv (basic_module_unformer_v2.py:13-14) — parameter holder
    return nn.Conv2d(cin, cout, 1, 1, 0)


def _dw(ch, k):              # bmu.dep_conv (basic_module_unformer_v2.py:17-18) — parameter holder
    return nn.Conv2d(ch, ch, k, 1, k // 2, groups=ch)


def _mixer(ch):
    """LGMixer parameters in the reference's creation order (LGT.py:183-198, 112-128, 149-160)."""
    half = ch // 2
    local = _Node()
    local.attach("to_qkv", _pw(half, 3 * half))
    pos = torch.empty(1, HEADS, WINDOW * WINDOW, WINDOW * WINDOW)
    nn.init.trunc_normal_(pos, mean=0.0, std=1.0, a=-2.0, b=2.0)      # LGT.py:127-128 (same algorithm as :21-42)
    local.attach("pos_emb", nn.Parameter(pos))
    glob = _Node(conv_amp=_seq(_dw(half, 1)), conv_pha=_seq(_dw(half, 1)))
    node = _Node(local_mixer=local, global_mixer=glob)
    node.attach("proj", _pw(ch, ch))
    return node


def _ffn(ch):
    """feed_forward parameters (LGT.py:91-101; depthwise_conv bmu:37-48)."""
    net = _Node()
    net.attach("0", _pw(ch, 4 * ch))
    net.attach("1", _Slot("GELU"))
    dwc = _Node()
    dwc.attach("point_conv", _pw(4 * ch, 4 * ch))
    dwc.attach("depth_conv", _dw(4 * ch, 3))
    net.attach("2", dwc)
    net.attach("3", _Slot("GELU"))
    net.attach("4", _pw(4 * ch, ch))
    return _Node(net=net)


def _prenorm_residual(ch, fn):
    """residual(pre_norm(channels, fn)) -> keys '<idx>.fn.norm.*' and '<idx>.fn.fn.*' (LGT.py:45-61)."""
    inner = _Node()
    inner.attach("fn", fn)                       # fn is built before the LayerNorm in the reference
    inner.attach("norm", nn.LayerNorm(ch))
    return _Node(fn=inner)


def _lgb(ch, blocks):
    """LGB (LGT.py:222-239)."""
    bl = _Node()
    for j in range(blocks):
        mixer = _prenorm_residual(ch, _mixer(ch))
        ffn = _prenorm_residual(ch, _ffn(ch))
        bl.attach(str(j), _seq(mixer, ffn))
    return _Node(blocks=bl)


def _lgt(bands):
    """LGT with num_block=[2,1], embed = 4*bands (unlg_former.py:44-48; LGT.py:251-303)."""
    c = 4 * bands
    node = _Node()
    pe = _Node()
    pe.attach("proj", _seq(_dw(bands, 1), _pw(bands, c)))
    pe.attach("norm", nn.LayerNorm(c))
    node.attach("patch_embed", pe)
    enc = _seq(_seq(_lgb(c, 2), _seq(_Slot("bicubic 1/2"), _pw(c, 2 * c))))
    node.attach("encoder_layers", enc)
    node.attach("bottleneck", _lgb(2 * c, 1))
    up = _seq(_Slot("bicubic x2"), _pw(2 * c, c))
    fuse = _pw(2 * c, c)
    dec = _seq(_seq(up, fuse, _lgb(c, 2)))
    node.attach("decoder_layers", dec)
    node.attach("tail", _seq(_Slot("bicubic x1"), _pw(c, bands)))
    return node


class Pansharpening(nn.Module):
    """Drop-in for ``models.unlg_former.Pansharpening`` backed by the sm_100a kernels.

    ``skip_dead_priors`` (default True): the reference never feeds a prior's output back into Z, only
    the last prior reaches the returned tensor (unlg_former.py:63-67).  Skipping priors 0..K-2 returns
    the identical tensor; set it to False to execute them anyway (for like-for-like timing)."""

    def __init__(self, cfg, logger=None, stage: int = 5, skip_dead_priors: bool = True):
        super().__init__()
        self.in_channels = int(cfg.ms_chans if hasattr(cfg, "ms_chans") else cfg["ms_chans"])
        self.stage = int(stage)
        self.up_factor = UP_FACTOR
        self.skip_dead_priors = bool(skip_dead_priors)
        if self.in_channels not in (4, 8):
            raise ValueError("lgteun_b200 supports ms_chans in {4, 8} (GF-2/WV-2 and WV-3, configs/unlg_former.py:12-19)")
        b = self.in_channels
        self.D = _seq(_Slot("bicubic 1/2"), _dw(b, 3), _Slot("bicubic 1/2"), _dw(b, 3))        # unlg_former.py:29-30
        self.DT = _seq(_Slot("bicubic x2"), _dw(b, 3), _Slot("bicubic x2"), _dw(b, 3))         # unlg_former.py:32-33
        self.R = _pw(b, 1)                                                                   # unlg_former.py:36
        self.RT = _pw(1, b)                                                                  # unlg_former.py:37
        self.eta = nn.ParameterList([nn.Parameter(torch.tensor(0.1)) for _ in range(self.stage)])   # :40
        self.prior_module = nn.ModuleList([_lgt(b) for _ in range(self.stage)])              # :42-48
        self._rt: Dict[int, dict] = {}      # device index -> {"handle", "sig"}; runtime only, never pickled
        self._flat = None                   # train.FlatParameters once a training-mode forward has run
        self.dropout_p = 0.1                # nn.Dropout(0.1) after the mixer projection (LGT.py:198)
        self._train_calls = 0

    # -- pickling / replication: runtime handles are per process ------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_rt"] = {}
        state["_flat"] = None
        return state

    def _invalidate_runtime(self):
        """Parameters were changed behind autograd's back (fused optimizer step): re-snapshot on the next eval forward."""
        for rt in self._rt.values():
            rt["sig"] = None

    def _flat_parameters(self):
        """The flat parameter / gradient buffers of the training step (created on first use; rebuilt if the parameters
        were moved, e.g. by .to() or a fresh load of the module)."""
        from .train import FlatParameters
        flat = self._flat
        if flat is not None:
            base = flat.param.data_ptr()
            params = dict(self.named_parameters())
            if not all(params[k].data_ptr() == base + 4 * off for k, off, _ in flat.layout):
                flat = None
        if flat is None:
            flat = self._flat = FlatParameters(self)
        return flat

    # -- runtime ---------------------------------------------------------------------------------------------
    def _weight_items(self):
        """(state_dict key, tensor) of every parameter.  Also valid inside an ``nn.DataParallel`` replica (the reference
        wraps the module when more than one GPU is visible, models/base/base_model.py:90-96): a replica's ``_parameters``
        are empty and its submodules hold the broadcast copies as plain attributes (``_former_parameters``,
        torch/nn/parallel/replicate.py), so ``parameters()`` / ``state_dict()`` cannot be used there."""
        for prefix, m in self.named_modules():
            former = getattr(m, "_former_parameters", None) or {}
            for k in list(m._parameters.keys()) + [k for k in former if k not in m._parameters]:
                t = m._parameters.get(k)
                if t is None:
                    t = former.get(k)
                if t is not None:
                    yield (prefix + "." if prefix else "") + k, t

    def _runtime(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        rt = self._rt.get(idx)
        if rt is None:
            rt = {"handle": _abi.Handle(idx, self.in_channels, self.stage), "sig": None}
            self._rt[idx] = rt
        items = list(self._weight_items())
        sig = tuple((t.data_ptr(), t._version) for _, t in items)
        if rt["sig"] != sig:                 # load_state_dict / optimizer step / .to() / a new replica: refresh the packed copy
            tensors = {}
            for name, p in items:
                t = p.detach()
                if t.device.type != "cuda" or t.device.index != idx:
                    raise RuntimeError(f"parameter {name} lives on {t.device}, inputs on cuda:{idx}; call .cuda() first")
                if t.dtype != torch.float32:
                    raise TypeError(f"parameter {name} must be float32, got {t.dtype}")
                tensors[name] = t.contiguous()
            rt["handle"].load_weights(tensors, torch.cuda.current_stream(device).cuda_stream)
            rt["sig"] = sig
        return rt["handle"]

    def forward(self, ms: torch.Tensor, pan: torch.Tensor) -> torch.Tensor:
        if ms.dim() != 4 or pan.dim() != 4:
            raise ValueError("expected ms [N,B,h,w] and pan [N,1,4h,4w]")
        n, b, h, w = ms.shape
        if b != self.in_channels or tuple(pan.shape) != (n, 1, UP_FACTOR * h, UP_FACTOR * w):
            raise ValueError(f"shape mismatch: ms {tuple(ms.shape)}, pan {tuple(pan.shape)}, bands {self.in_channels}")
        if ms.device.type != "cuda" or pan.device != ms.device:
            raise RuntimeError("lgteun_b200.Pansharpening runs on CUDA (sm_100a) tensors only; there is no CPU path")
        if ms.dtype != torch.float32 or pan.dtype != torch.float32:
            raise TypeError("ms and pan must be float32")
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for _, t in self._weight_items())
        if self.training or needs_grad:
            # train_iter path (models/unlg_former.py:94): Dropout(0.1) is active in train() mode (LGT.py:198,216) and the
            # output carries the autograd graph of the hand-written backward (train.cu)
            if ms.requires_grad or pan.requires_grad:
                raise NotImplementedError("gradients with respect to ms / pan are not computed (the reference never needs them)")
            if not self._parameters and not any(True for _ in self.parameters()):
                raise NotImplementedError("training inside nn.DataParallel replicas is not supported: use lgteun_b200.Trainer "
                                          "with one process per GPU (torch.distributed)")
            from .train import autograd_forward
            self._train_calls += 1
            p = self.dropout_p if self.training else 0.0
            rank = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()      # data-parallel ranks must not drop the same positions
            seed = (torch.initial_seed() + self._train_calls + 0x9E3779B1 * rank) & 0x7FFFFFFFFFFFFFFF
            if not needs_grad:
                with torch.no_grad():
                    return autograd_forward(self, ms, pan, p, seed)
            return autograd_forward(self, ms, pan, p, seed)
        with torch.cuda.device(ms.device):
            handle = self._runtime(ms.device)
            ms_c, pan_c = ms.contiguous(), pan.contiguous()
            out = torch.empty((n, b, UP_FACTOR * h, UP_FACTOR * w), dtype=torch.float32, device=ms.device)
            flags = 0 if self.skip_dead_priors else _abi.RUN_DEAD_PRIORS
            handle.forward(ms_c.data_ptr(), pan_c.data_ptr(), out.data_ptr(), n, h, w, flags,
                           torch.cuda.current_stream(ms.device).cuda_stream)
        return out

    def evaluate(self, pred: torch.Tensor, gt: torch.Tensor, bit_depth: int = 11, dynamic_range: float = 2047.5) -> torch.Tensor:
        """PSNR / SAM / ERGAS per image on the device, float64 — the reduced-resolution metrics the reference's test
        loop computes per image with numpy (models/base/base_model.py:304-327, models/base/metrics.py).  pred / gt are
        the normalised [N,B,H,W] CUDA tensors of the eval loop; returns a float64 CUDA tensor [N, 3].
        ``bit_depth`` is the de-normalisation scale 2**bit_depth - .5 (dataset/utils.py:252-263); ``dynamic_range`` is the
        PSNR peak, which the reference keeps at the module constant 2047.5 whatever cfg.bit_depth says
        (models/base/metrics.py:19,39)."""
        if pred.shape != gt.shape or pred.dim() != 4 or pred.shape[1] != self.in_channels:
            raise ValueError("pred and gt must both be [N,B,H,W]")
        if pred.device.type != "cuda" or gt.device != pred.device or pred.dtype != torch.float32 or gt.dtype != torch.float32:
            raise RuntimeError("evaluate() needs float32 CUDA tensors on one device")
        with torch.cuda.device(pred.device):
            handle = self._runtime(pred.device)
            n, _, h, w = pred.shape
            out = torch.empty((n, 3), dtype=torch.float64, device=pred.device)
            pc, gc = pred.contiguous(), gt.contiguous()
            max_value = float(2 ** bit_depth - 0.5)
            handle.op("metrics", pc.data_ptr(), gc.data_ptr(), out.data_ptr(), n, h, w, max_value,
                      stream=torch.cuda.current_stream(pred.device).cuda_stream)
            if float(dynamic_range) != max_value:        # the kernel's PSNR peak is max_value: 20 log10(R'/R) moves it
                import math
                out[:, 0] += 20.0 * math.log10(float(dynamic_range) / max_value)
        return out

    def normalize(self, raw: torch.Tensor, bit_depth: int = 11) -> torch.Tensor:
        """``data_normalize`` of the reference's loops (dataset/utils.py:232-248) on the device: raw / (2**bit_depth - .5)."""
        if raw.device.type != "cuda" or raw.dtype != torch.float32:
            raise RuntimeError("normalize() needs a float32 CUDA tensor")
        with torch.cuda.device(raw.device):
            handle = self._runtime(raw.device)
            src = raw.contiguous()
            out = torch.empty_like(src)
            handle.op("normalize", src.data_ptr(), out.data_ptr(), src.numel(), float(2 ** bit_depth - 0.5),
                      stream=torch.cuda.current_stream(raw.device).cuda_stream)
        return out

    def to_numpy_layout(self, x: torch.Tensor, bit_depth=None) -> torch.Tensor:
        """``torch2np`` (+ ``data_denormalize`` when bit_depth is given) of the reference's test loop on the device
        (models/base/utils.py:28-39, dataset/utils.py:252-263): [N,C,H,W] -> [N,H,W,C] (C == 1: [N,H,W]), still a
        CUDA tensor, so that one contiguous D2H copy replaces the strided host transpose."""
        if x.dim() != 4 or x.device.type != "cuda" or x.dtype != torch.float32:
            raise RuntimeError("to_numpy_layout() needs a float32 CUDA tensor [N,C,H,W]")
        n, c, h, w = x.shape
        with torch.cuda.device(x.device):
            handle = self._runtime(x.device)
            src = x.contiguous()
            out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
            handle.op("to_nhwc", src.data_ptr(), out.data_ptr(), n, c, h, w,
                      1.0 if bit_depth is None else float(2 ** bit_depth - 0.5),
                      stream=torch.cuda.current_stream(x.device).cuda_stream)
        return out.squeeze(-1) if c == 1 else out

    def extra_repr(self):
        return f"bands={self.in_channels}, stage={self.stage}, backend=sm_100a C-ABI ({_abi.LIB_PATH})"


def expected_state_dict_keys(bands: int, stages: int):
    """The key grammar of SURVEY.md Appendix B, generated independently of the module tree (used by tests)."""
    c = 4 * bands
    keys = []
    for m in ("D.1", "D.3", "DT.1", "DT.3", "R", "RT"):
        keys += [f"{m}.weight", f"{m}.bias"]
    keys += [f"eta.{i}" for i in range(stages)]

    def block(p):
        out = [f"{p}.0.fn.norm.weight", f"{p}.0.fn.norm.bias", f"{p}.0.fn.fn.local_mixer.pos_emb"]
        for m in ("local_mixer.to_qkv", "global_mixer.conv_amp.0", "global_mixer.conv_pha.0", "proj"):
            out += [f"{p}.0.fn.fn.{m}.weight", f"{p}.0.fn.fn.{m}.bias"]
        out += [f"{p}.1.fn.norm.weight", f"{p}.1.fn.norm.bias"]
        for m in ("net.0", "net.2.point_conv", "net.2.depth_conv", "net.4"):
            out += [f"{p}.1.fn.fn.{m}.weight", f"{p}.1.fn.fn.{m}.bias"]
        return out

    for i in range(stages):
        p = f"prior_module.{i}"
        for m in ("patch_embed.proj.0", "patch_embed.proj.1", "patch_embed.norm"):
            keys += [f"{p}.{m}.weight", f"{p}.{m}.bias"]
        for j in range(2):
            keys += block(f"{p}.encoder_layers.0.0.blocks.{j}")
        keys += [f"{p}.encoder_layers.0.1.1.weight", f"{p}.encoder_layers.0.1.1.bias"]
        keys += block(f"{p}.bottleneck.blocks.0")
        keys += [f"{p}.decoder_layers.0.0.1.weight", f"{p}.decoder_layers.0.0.1.bias",
                 f"{p}.decoder_layers.0.1.weight", f"{p}.decoder_layers.0.1.bias"]
        for j in range(2):
            keys += block(f"{p}.decoder_layers.0.2.blocks.{j}")
        keys += [f"{p}.tail.1.weight", f"{p}.tail.1.bias"]
    return keys


def param_count(bands: int, stages: int = 2) -> int:
    """Closed-form parameter count (paper Table 4: 202,183 for 4 bands, 540,043 for 8 bands at K=2)."""
    c = 4 * bands

    def block(ch):
        h = ch // 2
        return (2 * ch + 2 * 64 * 64 + 3 * h * h + 3 * h + 4 * h + ch * ch + ch + 2 * ch
                + 4 * ch * ch + 4 * ch + 16 * ch * ch + 4 * ch + 36 * ch + 4 * ch + 4 * ch * ch + ch)

    prior = (2 * bands + c * bands + c + 2 * c + 4 * block(c) + block(2 * c) + 2 * c * c + 2 * c
             + 2 * (2 * c * c + c) + bands * c + bands)
    shared = 4 * (9 * bands + bands) + bands + 1 + bands + bands
    return shared + stages + stages * prior
