"""CPU oracle for the LGTEUN forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch CPU ops, fp32 by default, fp64 on request) of the
reference's stage-wise unfolding forward.  It is NOT part of the product path: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it.  The product (``lgteun_b200``) never imports anything under ``oracle/`` and fails loudly
when its CUDA extension is missing.

Parity pinning: the reference ships no tests and no golden vectors.  The oracle
is therefore pinned against outputs of the *reference itself*, executed in the build container from
``/root/reference`` by ``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` replays them).  All arithmetic that the reference delegates to
PyTorch (conv, layer-norm, bicubic interpolate, softmax, FFT; torch 2.11 in this image, the
reference pins 1.9.1) is delegated to the same PyTorch calls here, so the restatement agrees with
the reference bit-for-bit on the goldens.

Weights are taken as a flat ``state_dict`` with the reference's key grammar.
Every function cites the reference lines it follows (paths relative to /root/reference).

A second group of functions (``*_explicit``) restates the same maths with explicit indexing/taps
(no F.interpolate / einops / rfft2) — they document exactly what the CUDA kernels implement and are
checked against the torch forms in ``tests/test_oracle_explicit.py``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

WINDOW = 8          # models/unlg_former.py:47  window_size=8
HEADS = 2           # models/unlg_former.py:48  num_heads=2
LN_EPS = 1e-5       # nn.LayerNorm default, models/common/LGT.py:58
CUBIC_A = -0.75     # ATen upsample_bicubic2d coefficient


# --------------------------------------------------------------------------------------------
# small building blocks
# --------------------------------------------------------------------------------------------
def _w(sd: SD, key: str) -> Tensor:
    return sd[key]


def bicubic(x: Tensor, scale: float) -> Tensor:
    """models/common/basic_module_unformer_v2.py:21-23 and :32-34 (sampling_ / sampling_unit_)."""
    return F.interpolate(x, scale_factor=scale, mode="bicubic", align_corners=False,
                         recompute_scale_factor=False)


def pconv(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """1x1 convolution, basic_module_unformer_v2.py:13-14 (point_conv)."""
    return F.conv2d(x, _w(sd, prefix + ".weight"), _w(sd, prefix + ".bias"))


def dconv(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """depthwise k x k convolution with k//2 zero padding, basic_module_unformer_v2.py:17-18."""
    w = _w(sd, prefix + ".weight")
    return F.conv2d(x, w, _w(sd, prefix + ".bias"), padding=w.shape[-1] // 2, groups=w.shape[0])


def layer_norm(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """nn.LayerNorm over the last (channel) axis, LGT.py:58-61."""
    w = _w(sd, prefix + ".weight")
    return F.layer_norm(x, (w.shape[0],), w, _w(sd, prefix + ".bias"), LN_EPS)


# --------------------------------------------------------------------------------------------
# data module   (models/unlg_former.py:29-40, 58-61)
# --------------------------------------------------------------------------------------------
def degrade(sd: SD, z: Tensor) -> Tensor:
    """D: [bicubic 1/2 -> depthwise 3x3] twice, unlg_former.py:29-30."""
    z = dconv(sd, "D.1", bicubic(z, 0.5))
    return dconv(sd, "D.3", bicubic(z, 0.5))


def degrade_adjoint(sd: SD, r: Tensor) -> Tensor:
    """DT: [bicubic x2 -> depthwise 3x3] twice, unlg_former.py:32-33."""
    r = dconv(sd, "DT.1", bicubic(r, 2))
    return dconv(sd, "DT.3", bicubic(r, 2))


def data_step(sd: SD, z: Tensor, ms: Tensor, pan: Tensor, stage: int) -> Tensor:
    """One proximal-gradient step, unlg_former.py:58-61."""
    ms_term = degrade_adjoint(sd, degrade(sd, z) - ms)
    pan_term = pconv(sd, "RT", pconv(sd, "R", z) - pan)
    return z - _w(sd, f"eta.{stage}") * (ms_term + pan_term)


# --------------------------------------------------------------------------------------------
# Local-Global Transformer prior   (models/common/LGT.py)
# --------------------------------------------------------------------------------------------
def patch_embed(sd: SD, p: str, x: Tensor) -> Tensor:
    """LGT.py:72-88: depthwise 1x1 -> 1x1 B->C -> NHWC -> LayerNorm(C)."""
    x = dconv(sd, p + ".proj.0", x)
    x = pconv(sd, p + ".proj.1", x).permute(0, 2, 3, 1)
    return layer_norm(sd, p + ".norm", x)


def local_mixer(sd: SD, p: str, x: Tensor) -> Tensor:
    """Window multi-head self-attention on the local channel half, LGT.py:130-146 and the window
    merge of LGT.py:207-208.  x: [b,h,w,c2] -> [b,h,w,c2]."""
    b, h, w, c2 = x.shape
    nh, nw = h // WINDOW, w // WINDOW
    # 'b (h i) (w j) c -> b c (h w) (i j)'   (LGT.py:135)
    xw = x.reshape(b, nh, WINDOW, nw, WINDOW, c2).permute(0, 5, 1, 3, 2, 4).reshape(b, c2, nh * nw, 64)
    qkv = pconv(sd, p + ".to_qkv", xw)
    q, k, v = qkv.chunk(3, dim=1)                                       # LGT.py:136
    d = c2 // HEADS

    def heads(t: Tensor) -> Tensor:                                     # 'b (h c) m n -> (b m) h n c'  (LGT.py:138)
        return t.reshape(b, HEADS, d, nh * nw, 64).permute(0, 3, 1, 4, 2).reshape(b * nh * nw, HEADS, 64, d)

    q, k, v = heads(q), heads(k), heads(v)
    q = q * (d ** -0.5)                                                 # LGT.py:119,139
    sim = torch.einsum("bhic,bhjc->bhij", q, k) + _w(sd, p + ".pos_emb")  # LGT.py:140-141
    att = torch.softmax(sim, dim=-1)                                    # LGT.py:142
    out = torch.einsum("bhij,bhjc->bhic", att, v)                       # LGT.py:143
    out = out.permute(0, 2, 1, 3).reshape(b * nh * nw, 64, c2)          # 'b h m c -> b m (h c)'  (LGT.py:144)
    # '(b h w) (i j) c -> b (h i) (w j) c'  (LGT.py:207-208)
    return out.reshape(b, nh, nw, WINDOW, WINDOW, c2).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c2)


# torch.fft.rfft2 returns an exact +0.0 imaginary part in the four purely real bins for power-of-two sizes, but a rounding
# residue of arbitrary sign at many other lengths (8 x odd, 16 x 5, 16 x 7 ..., in fp32 AND in fp64): where the real part is
# negative, angle() then flips between +pi and -pi and conv_pha spreads the difference over the whole map.  With
# this switch on, global_mixer evaluates the same formula with the imaginary part of those bins set to +0.0 (what exact
# arithmetic gives): the well-defined reference for sizes that are not powers of two (tests only; default = literal restatement).
EXACT_REAL_BINS = False


def global_mixer(sd: SD, p: str, x: Tensor) -> Tensor:
    """FFT amplitude/phase mixer on the global channel half, LGT.py:162-180."""
    if EXACT_REAL_BINS:
        return global_mixer_3pass(sd, p, x)
    b, h, w, c2 = x.shape
    x = x.permute(0, 3, 1, 2)
    fre = torch.fft.rfft2(x, norm="backward")                           # LGT.py:166
    amp = dconv(sd, p + ".conv_amp.0", torch.abs(fre))                  # LGT.py:168,171
    pha = dconv(sd, p + ".conv_pha.0", torch.angle(fre))                # LGT.py:169,172
    real = amp * torch.cos(pha) + 1e-8                                  # LGT.py:174
    imag = amp * torch.sin(pha) + 1e-8                                  # LGT.py:175
    out = torch.complex(real, imag) + 1e-8                              # LGT.py:177
    out = torch.abs(torch.fft.irfft2(out, s=(h, w), norm="backward"))   # LGT.py:178
    return out.permute(0, 2, 3, 1)


def lg_mixer(sd: SD, p: str, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """LGMixer, LGT.py:200-219.  ``mask`` (NHWC, values 0 or 1/(1-p)) is the train-mode nn.Dropout(0.1) of LGT.py:198,216
    applied to the projection output; None = eval mode (identity)."""
    c = x.shape[-1]
    x1 = local_mixer(sd, p + ".local_mixer", x[..., : c // 2].contiguous())
    x2 = global_mixer(sd, p + ".global_mixer", x[..., c // 2:].contiguous())
    out = torch.cat((x1, x2), dim=-1).permute(0, 3, 1, 2)
    out = pconv(sd, p + ".proj", out).permute(0, 2, 3, 1)
    return out if mask is None else out * mask


def feed_forward(sd: SD, p: str, x: Tensor) -> Tensor:
    """conv-FFN, LGT.py:95-109 (+ depthwise_conv bmu:37-53): 1x1 -> GELU -> 1x1 -> dw3x3 -> GELU -> 1x1."""
    t = pconv(sd, p + ".net.0", x.permute(0, 3, 1, 2))
    t = F.gelu(t)
    t = pconv(sd, p + ".net.2.point_conv", t)
    t = dconv(sd, p + ".net.2.depth_conv", t)
    t = F.gelu(t)
    t = pconv(sd, p + ".net.4", t)
    return t.permute(0, 2, 3, 1)


def lgb_block(sd: SD, p: str, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """One (mixer, ffn) pair with pre-norm + residual, LGT.py:231-247, :45-61."""
    x = lg_mixer(sd, p + ".0.fn.fn", layer_norm(sd, p + ".0.fn.norm", x), mask) + x
    x = feed_forward(sd, p + ".1.fn.fn", layer_norm(sd, p + ".1.fn.norm", x)) + x
    return x


def lgb(sd: SD, p: str, x: Tensor, num_blocks: int, masks: Optional[List[Tensor]] = None) -> Tensor:
    """LGB: num_blocks blocks, NHWC in, NCHW out (LGT.py:240-248)."""
    for j in range(num_blocks):
        x = lgb_block(sd, f"{p}.blocks.{j}", x, None if masks is None else masks[j])
    return x.permute(0, 3, 1, 2)


def lgt(sd: SD, p: str, x: Tensor, masks: Optional[List[Tensor]] = None) -> Tensor:
    """LGT U-Net with num_block=[2,1] (unlg_former.py:47-48), LGT.py:314-344.  ``masks``: the five dropout masks of the
    blocks in execution order (encoder 0, 1; bottleneck; decoder 0, 1), None = eval mode."""
    m = (lambda a, b: None) if masks is None else (lambda a, b: masks[a:b])
    fea = patch_embed(sd, p + ".patch_embed", x)
    skip = lgb(sd, p + ".encoder_layers.0.0", fea, 2, m(0, 2))                           # LGT.py:326-327
    fea = pconv(sd, p + ".encoder_layers.0.1.1", bicubic(skip, 0.5)).permute(0, 2, 3, 1)  # LGT.py:328-329
    fea = lgb(sd, p + ".bottleneck", fea, 1, m(2, 3))                                    # LGT.py:332
    fea = pconv(sd, p + ".decoder_layers.0.0.1", bicubic(fea, 2))                        # LGT.py:336
    fea = pconv(sd, p + ".decoder_layers.0.1", torch.cat([fea, skip], dim=1))            # LGT.py:337-338
    fea = lgb(sd, p + ".decoder_layers.0.2", fea.permute(0, 2, 3, 1), 2, m(3, 5))        # LGT.py:339
    return pconv(sd, p + ".tail.1", bicubic(fea, 1)) + x                                 # LGT.py:342


# --------------------------------------------------------------------------------------------
# the unfolding network   (models/unlg_former.py:50-67)
# --------------------------------------------------------------------------------------------
def num_stages(sd: SD) -> int:
    return sum(1 for k in sd if k.startswith("eta."))


def forward(sd: SD, ms: Tensor, pan: Tensor, stages: Optional[int] = None,
            skip_dead_priors: bool = True, trace: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """Pansharpening.forward.  The reference never feeds a prior's output back into Z
    (unlg_former.py:63-67: ``Z_`` is appended to a list, only the last entry is returned), so the
    priors of stages 0..K-2 do not influence the result; ``skip_dead_priors`` omits them (bitwise
    the same returned tensor).  ``trace`` collects named intermediates for per-op tests."""
    K = num_stages(sd) if stages is None else stages
    with torch.no_grad():
        z = bicubic(ms, 4)                                              # unlg_former.py:53
        out = z
        for i in range(K):
            z = data_step(sd, z, ms, pan, i)                            # unlg_former.py:58-61
            if trace is not None:
                trace[f"z{i}"] = z
            if i == K - 1 or not skip_dead_priors:
                out = lgt(sd, f"prior_module.{i}", z)                   # unlg_former.py:63
        return out


def forward_train(sd: SD, ms: Tensor, pan: Tensor, masks: Optional[List[Tensor]] = None) -> Tensor:
    """Pansharpening.forward with autograd enabled and the live prior's dropout masks given explicitly (train() mode,
    unlg_former.py:50-67 + LGT.py:198,216).  Dead priors are skipped: their outputs never reach the loss, so torch leaves
    their parameters' .grad at None (unlg_former.py:63-67)."""
    K = num_stages(sd)
    z = bicubic(ms, 4)
    for i in range(K):
        z = data_step(sd, z, ms, pan, i)
    return lgt(sd, f"prior_module.{K - 1}", z, masks)


def train_step_grads(sd: SD, ms: Tensor, pan: Tensor, gt: Tensor, masks: Optional[List[Tensor]] = None,
                     loss_weight: float = 1.0):
    """One UnlgFormer.train_iter up to loss.backward() (models/unlg_former.py:87-110): nn.L1Loss(out, gt) * w
    (models/base/losses.py:29,39; configs/unlg_former.py:88-90).  Returns (out, loss, {key: grad or None})."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = forward_train(leaf, ms, pan, masks)
    loss = F.l1_loss(out, gt) * loss_weight
    loss.backward()
    return out.detach(), loss.detach(), {k: v.grad for k, v in leaf.items()}


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1: float = 0.9, b2: float = 0.999,
              eps: float = 1e-8):
    """torch.optim.Adam single-tensor update (base_model.py:121-122: Adam(betas=(0.9, 0.999), lr)), no weight decay."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    denom = v.sqrt() / math.sqrt(1 - b2 ** step) + eps
    return p - (lr / (1 - b1 ** step)) * m / denom, m, v


# --------------------------------------------------------------------------------------------
# explicit restatements (what the CUDA kernels implement)
# --------------------------------------------------------------------------------------------
def cubic_taps(t: float) -> List[float]:
    """Keys cubic convolution weights (A=-0.75) for fractional offset t in [0,1): taps at
    floor-1, floor, floor+1, floor+2 (ATen UpSample.h get_cubic_upsample_coefficients)."""
    A = CUBIC_A

    def c1(x):  # |x| <= 1
        return ((A + 2) * x - (A + 3)) * x * x + 1

    def c2(x):  # 1 < |x| < 2
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A

    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def resize_matrix(n_in: int, scale: float, dtype=np.float64) -> np.ndarray:
    """Dense [n_out, n_in] matrix of the 1-D bicubic resize with align_corners=False:
    src = (dst + 0.5) / scale - 0.5, 4 taps, tap indices clamped to [0, n_in-1]."""
    n_out = int(math.floor(n_in * scale))
    m = np.zeros((n_out, n_in), dtype=dtype)
    for o in range(n_out):
        src = (o + 0.5) / scale - 0.5
        f = math.floor(src)
        taps = cubic_taps(src - f)
        for k in range(4):
            idx = min(max(f - 1 + k, 0), n_in - 1)
            m[o, idx] += taps[k]
    return m


def bicubic_explicit(x: Tensor, scale: float) -> Tensor:
    """Separable bicubic resize as two dense matrix products (fp64 internally)."""
    h, w = x.shape[-2:]
    my = torch.from_numpy(resize_matrix(h, scale))
    mx = torch.from_numpy(resize_matrix(w, scale))
    y = torch.einsum("oh,...hw->...ow", my, x.double())
    y = torch.einsum("pw,...ow->...op", mx, y)
    return y.to(x.dtype)


def local_mixer_explicit(sd: SD, p: str, x: Tensor) -> Tensor:
    """Window attention with explicit loops over windows/heads (semantics of SURVEY.md §8a-notes)."""
    b, h, w, c2 = x.shape
    d = c2 // HEADS
    wq = _w(sd, p + ".to_qkv.weight").reshape(3 * c2, c2)
    bq = _w(sd, p + ".to_qkv.bias")
    pos = _w(sd, p + ".pos_emb")[0]
    out = torch.empty_like(x)
    for n in range(b):
        for wy in range(h // WINDOW):
            for wx in range(w // WINDOW):
                tok = x[n, wy * 8:wy * 8 + 8, wx * 8:wx * 8 + 8, :].reshape(64, c2)   # token = i*8+j
                qkv = tok @ wq.t() + bq
                for hd in range(HEADS):
                    q = qkv[:, hd * d:(hd + 1) * d] * (d ** -0.5)
                    k = qkv[:, c2 + hd * d:c2 + (hd + 1) * d]
                    v = qkv[:, 2 * c2 + hd * d:2 * c2 + (hd + 1) * d]
                    att = torch.softmax(q @ k.t() + pos[hd], dim=-1)
                    out[n, wy * 8:wy * 8 + 8, wx * 8:wx * 8 + 8, hd * d:(hd + 1) * d] = (att @ v).reshape(8, 8, d)
    return out


def global_mixer_3pass(sd: SD, p: str, x: Tensor) -> Tensor:
    """global_mixer as the three passes the CUDA path runs: row rFFT along W; column FFT along H +
    pointwise amp/phase + inverse column FFT; row C2R along W that ignores Im of bins 0 and W/2;
    1/(H*W) scaling; abs.  The four purely-real bins get an exact +0.0 imaginary part."""
    b, h, w, c2 = x.shape
    xr = x.permute(0, 3, 1, 2)
    rows = torch.fft.rfft(xr, dim=-1)                       # [b,c2,h,w/2+1]
    spec = torch.fft.fft(rows, dim=-2)
    re, im = spec.real.clone(), spec.imag.clone()
    for ky in (0, h // 2):
        for kx in (0, w // 2):
            im[..., ky, kx] = 0.0
    wa = _w(sd, p + ".conv_amp.0.weight").reshape(1, c2, 1, 1)
    ba = _w(sd, p + ".conv_amp.0.bias").reshape(1, c2, 1, 1)
    wp = _w(sd, p + ".conv_pha.0.weight").reshape(1, c2, 1, 1)
    bp = _w(sd, p + ".conv_pha.0.bias").reshape(1, c2, 1, 1)
    amp = torch.sqrt(re * re + im * im) * wa + ba
    pha = torch.atan2(im, re) * wp + bp
    o_re = (amp * torch.cos(pha) + 1e-8) + 1e-8
    o_im = amp * torch.sin(pha) + 1e-8
    cols = torch.fft.ifft(torch.complex(o_re, o_im), dim=-2, norm="forward")    # unnormalised inverse
    cre, cim = cols.real, cols.imag.clone()
    cim[..., 0] = 0.0
    cim[..., w // 2] = 0.0
    full = torch.fft.irfft(torch.complex(cre, cim), n=w, dim=-1, norm="forward")
    return (full / (h * w)).abs().permute(0, 2, 3, 1)


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default (exact erf form), LGT.py:97,99."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))
