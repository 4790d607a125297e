"""Host-side mirror of the companion operators next to the hot path.

  * ``Freprocess(channels)`` / ``forward(msf, panf)``       models/SFIIN.py:210-236

Same constructor, the same ``state_dict`` keys, shapes and default initialisation (the reference's construction order:
pre1, pre2, amp_fuse, pha_fuse, post), so a SFIIN checkpoint's ``fre_process.*`` tensors load unchanged.  The module has
no PyTorch compute path: ``forward`` hands device pointers to ``lgteun_op_freprocess`` (include/lgteun.h) and raises on
CPU tensors, unsupported shapes, under autograd, or if the CUDA library is missing."""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _abi

FREPROCESS_KEYS = ("pre1.weight", "pre1.bias", "pre2.weight", "pre2.bias", "amp_fuse.0.weight", "amp_fuse.0.bias",
                   "amp_fuse.2.weight", "amp_fuse.2.bias", "pha_fuse.0.weight", "pha_fuse.0.bias", "pha_fuse.2.weight",
                   "pha_fuse.2.bias", "post.weight", "post.bias")


class _Slot(nn.Module):
    """Parameter-free placeholder at index 1 of amp_fuse / pha_fuse (the reference's LeakyReLU(0.1))."""


def _conv1x1(cin, cout):
    return nn.Conv2d(cin, cout, 1, 1, 0)     # parameter container with torch's default init; never called


class Freprocess(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = int(channels)
        self.pre1 = _conv1x1(channels, channels)
        self.pre2 = _conv1x1(channels, channels)
        self.amp_fuse = nn.Sequential(_conv1x1(2 * channels, channels), _Slot(), _conv1x1(channels, channels))
        self.pha_fuse = nn.Sequential(_conv1x1(2 * channels, channels), _Slot(), _conv1x1(channels, channels))
        self.post = _conv1x1(channels, channels)
        self._ws = None

    def __getstate__(self):                  # the runner pickles module objects (models/base/base_model.py:362-368)
        state = self.__dict__.copy()
        state["_ws"] = None                  # scratch memory is not part of the model
        return state

    def forward(self, msf, panf):
        if msf.shape != panf.shape or msf.dim() != 4 or msf.shape[1] != self.channels:
            raise ValueError(f"Freprocess: expected two [N,{self.channels},H,W] tensors, got {tuple(msf.shape)} and "
                             f"{tuple(panf.shape)}")
        if not (msf.is_cuda and panf.is_cuda) or msf.dtype != torch.float32 or panf.dtype != torch.float32:
            raise RuntimeError("Freprocess: inputs must be float32 CUDA tensors (there is no CPU / eager fallback)")
        if torch.is_grad_enabled() and (msf.requires_grad or panf.requires_grad
                                        or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("Freprocess: forward only — call it under torch.no_grad()")
        sd = self.state_dict()
        dev = msf.device
        ws_t = [sd[k].detach().to(dev, torch.float32).contiguous() for k in FREPROCESS_KEYS]
        n, c, h, w = msf.shape
        lib = _abi.lib()
        need = lib.lgteun_op_freprocess_workspace_bytes(n, c, h, w)
        if self._ws is None or self._ws.device != dev or self._ws.numel() * 4 < need:
            self._ws = torch.empty(max(1, (need + 3) // 4), dtype=torch.float32, device=dev)
        msf, panf = msf.contiguous(), panf.contiguous()
        out = torch.empty_like(msf)
        ptrs = (ctypes.c_void_p * 14)(*[t.data_ptr() for t in ws_t])
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _abi.check(lib.lgteun_op_freprocess(dev.index or 0, msf.data_ptr(), panf.data_ptr(), out.data_ptr(), n, c, h, w,
                                                ptrs, self._ws.data_ptr(), self._ws.numel() * 4, ctypes.c_void_p(stream)))
        for t in ws_t:                       # the kernels run asynchronously on `stream`
            t.record_stream(torch.cuda.current_stream(dev))
        return out


def _shift_mask(window_size, displacement, upper_lower):
    """The -inf masks of a shifted block (models/common/modules.py:318-333): a token of the last `displacement` rows
    (upper_lower) or columns (left_right) of a window only attends to tokens of the same side."""
    t = torch.arange(window_size * window_size)
    side = ((t // window_size) if upper_lower else (t % window_size)) >= window_size - displacement
    mask = torch.zeros(window_size ** 2, window_size ** 2)
    mask[side[:, None] != side[None, :]] = float("-inf")
    return mask


class WindowAttention(nn.Module):
    """PanFormer's window attention, models/common/modules.py:341-422 (same constructor, state_dict keys and default
    initialisation order); ``forward(x, y=None)`` on channels-last [b, n_h, n_w, dim] CUDA tensors through
    ``lgteun_op_window_attention``.  No PyTorch compute path."""

    def __init__(self, dim, heads, head_dim, shifted, window_size, relative_pos_embedding, cross_attn):
        super().__init__()
        inner = head_dim * heads
        self.dim, self.heads, self.head_dim = int(dim), int(heads), int(head_dim)
        self.scale = head_dim ** -0.5
        self.window_size, self.shifted = int(window_size), bool(shifted)
        self.relative_pos_embedding, self.cross_attn = bool(relative_pos_embedding), bool(cross_attn)
        if self.shifted:
            d = window_size // 2
            self.upper_lower_mask = nn.Parameter(_shift_mask(window_size, d, True), requires_grad=False)
            self.left_right_mask = nn.Parameter(_shift_mask(window_size, d, False), requires_grad=False)
        if not self.cross_attn:
            self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        else:
            self.to_kv = nn.Linear(dim, inner * 2, bias=False)
            self.to_q = nn.Linear(dim, inner, bias=False)
        n = 2 * window_size - 1 if self.relative_pos_embedding else window_size ** 2
        self.pos_embedding = nn.Parameter(torch.randn(n, n))
        self.to_out = nn.Linear(inner, dim)

    def forward(self, x, y=None):
        if self.cross_attn != (y is not None):
            raise ValueError("WindowAttention: y is required exactly when cross_attn=True")
        if x.dim() != 4 or x.shape[-1] != self.dim or (y is not None and y.shape != x.shape):
            raise ValueError(f"WindowAttention: expected [b, n_h, n_w, {self.dim}] inputs, got {tuple(x.shape)}")
        if not x.is_cuda or x.dtype != torch.float32 or (y is not None and (not y.is_cuda or y.dtype != torch.float32)):
            raise RuntimeError("WindowAttention: inputs must be float32 CUDA tensors (there is no CPU / eager fallback)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("WindowAttention: forward only — call it under torch.no_grad()")
        dev = x.device
        inner = self.heads * self.head_dim

        def dv(t):
            return t.detach().to(dev, torch.float32).contiguous()

        if self.cross_attn:
            wq, wkv = dv(self.to_q.weight), dv(self.to_kv.weight)
            wq_ptr, wkv_ptr = wq.data_ptr(), wkv.data_ptr()
        else:
            wq = wkv = dv(self.to_qkv.weight)
            wq_ptr, wkv_ptr = wq.data_ptr(), wq.data_ptr() + inner * self.dim * 4
        wo, bo, pos = dv(self.to_out.weight), dv(self.to_out.bias), dv(self.pos_embedding)
        ul = dv(self.upper_lower_mask) if self.shifted else None
        lr = dv(self.left_right_mask) if self.shifted else None
        x = x.contiguous()
        y = y.contiguous() if y is not None else None
        out = torch.empty_like(x)
        b, n_h, n_w, _ = x.shape
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            _abi.check(_abi.lib().lgteun_op_window_attention(
                dev.index or 0, x.data_ptr(), y.data_ptr() if y is not None else None, out.data_ptr(), b, n_h, n_w, self.dim,
                self.heads, self.head_dim, self.window_size, int(self.shifted), int(self.relative_pos_embedding),
                float(self.scale), wq_ptr, wkv_ptr, wo.data_ptr(), bo.data_ptr(), pos.data_ptr(),
                ul.data_ptr() if ul is not None else None, lr.data_ptr() if lr is not None else None,
                ctypes.c_void_p(cur.cuda_stream)))
        for t in (wq, wkv, wo, bo, pos, ul, lr):
            if t is not None:
                t.record_stream(cur)
        return out


def swap_companions(root: nn.Module):
    """Drop-in inside the reference's own networks: replace every ``Freprocess`` (models/SFIIN.py:210, a child of ``SpaFre``
    :247) and every ``WindowAttention`` (models/common/modules.py:341, inside ``SwinBlock.attention_block`` :428) of ``root``
    by the lgteun_b200 module with the same constructor arguments, weights, device and train / eval flag.  The rest of the
    network keeps running in PyTorch; the swapped operators run under ``torch.no_grad()`` (inference).  Instances whose
    configuration the kernels do not cover (window size other than 4, unsupported channel counts) are left alone.
    Returns ``{"Freprocess": n, "WindowAttention": m}``."""
    counts = {"Freprocess": 0, "WindowAttention": 0}
    for parent in list(root.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, (Freprocess, WindowAttention)):
                continue
            kind = type(child).__name__
            if kind == "Freprocess" and isinstance(getattr(child, "pre1", None), nn.Conv2d):
                channels = child.pre1.in_channels
                if channels not in (4, 8, 16):
                    continue
                new = Freprocess(channels)
            elif kind == "WindowAttention" and isinstance(getattr(child, "to_out", None), nn.Linear):
                heads, inner, dim = int(child.heads), child.to_out.in_features, child.to_out.out_features
                head_dim = inner // heads
                if (int(child.window_size) != 4 or head_dim not in (8, 16, 32) or dim % 4 or dim > 128 or inner > 128
                        or 4 * dim * inner * 4 + 32 * 1024 > 227 * 1024):
                    continue
                new = WindowAttention(dim=dim, heads=heads, head_dim=head_dim, shifted=bool(child.shifted), window_size=4,
                                      relative_pos_embedding=bool(child.relative_pos_embedding), cross_attn=bool(child.cross_attn))
            else:
                continue
            new.load_state_dict(child.state_dict())
            first = next(child.parameters(), None)
            if first is not None:
                new.to(first.device)
            new.train(child.training)
            parent._modules[name] = new
            counts[kind] += 1
    return counts
