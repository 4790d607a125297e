"""Generate the committed golden fixtures by executing the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference is not on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/weights_b{4,8}.npz and tests/golden/case_*.npz.  Each case holds the inputs,
the reference output of Pansharpening.forward (models/unlg_former.py:50-67) and named intermediates
captured with forward hooks on the reference's own sub-modules (LGT.py classes), so that both the
oracle restatement and the CUDA kernels can be pinned per operator."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402

SEED = 19971118   # configs/unlg_former.py:66


def smooth_scene(n, bands, size, gen):
    """Metric-set inputs: low-passed noise in [0.05,0.95], Wald protocol LrMS/PAN."""
    gt = torch.rand(n, bands, size, size, generator=gen)
    k = torch.arange(-4, 5, dtype=torch.float32)
    g = torch.exp(-k * k / 8.0)
    g = (g / g.sum())
    for _ in range(2):
        gt = F.conv2d(F.pad(gt, (4, 4, 0, 0), mode="reflect"), g.view(1, 1, 1, 9).repeat(bands, 1, 1, 1), groups=bands)
        gt = F.conv2d(F.pad(gt, (0, 0, 4, 4), mode="reflect"), g.view(1, 1, 9, 1).repeat(bands, 1, 1, 1), groups=bands)
    lo, hi = gt.amin(dim=(2, 3), keepdim=True), gt.amax(dim=(2, 3), keepdim=True)
    gt = 0.05 + 0.9 * (gt - lo) / (hi - lo)
    pan = gt.mean(dim=1, keepdim=True)
    ms = F.interpolate(gt, scale_factor=0.25, mode="bicubic", align_corners=False, recompute_scale_factor=False)
    return ms.contiguous(), pan.contiguous(), gt.contiguous()


def capture(net, ms, pan):
    """Run the reference and collect intermediates of the LAST prior (the live one)."""
    got = {}
    prior = net.prior_module[-1]
    blk0 = prior.encoder_layers[0][0].blocks[0]
    bott = prior.bottleneck.blocks[0]
    taps = {
        "pe": prior.patch_embed,
        "enc0_mixer_in": None,
        "enc0_local": blk0[0].fn.fn.local_mixer,
        "enc0_global": blk0[0].fn.fn.global_mixer,
        "enc0_mixer": blk0[0],
        "enc0_ffn": blk0[1],
        "enc": prior.encoder_layers[0][0],
        "down": prior.encoder_layers[0][1],
        "bott_local": bott[0].fn.fn.local_mixer,
        "bott_global": bott[0].fn.fn.global_mixer,
        "bott": prior.bottleneck,
        "up": prior.decoder_layers[0][0],
        "fuse": prior.decoder_layers[0][1],
        "dec": prior.decoder_layers[0][2],
        "prior_in": None,
    }
    hooks = []
    for name, mod in taps.items():
        if mod is None:
            continue
        hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: got.__setitem__(name, o.detach().clone())))
    hooks.append(prior.register_forward_pre_hook(lambda m, i: got.__setitem__("prior_in", i[0].detach().clone())))
    hooks.append(blk0[0].fn.fn.local_mixer.register_forward_pre_hook(
        lambda m, i: got.__setitem__("enc0_local_in", i[0].detach().clone())))
    hooks.append(blk0[0].fn.fn.global_mixer.register_forward_pre_hook(
        lambda m, i: got.__setitem__("enc0_global_in", i[0].detach().clone())))
    with torch.no_grad():
        out = net(ms, pan)
    for h in hooks:
        h.remove()
    got["out"] = out
    return got


def main():
    torch.set_num_threads(1)           # one thread: the deterministic fp32 summation order
    _, _, mtc = ref_import.load()
    nets = {}
    for bands in (4, 8):
        net = ref_import.build(bands, stages=2, seed=SEED)
        nets[bands] = net
        sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
        np.savez(os.path.join(HERE, f"weights_b{bands}.npz"), **sd)
        print("weights", bands, len(sd), sum(v.size for v in sd.values()))

    ALL = None
    cases = [  # name, bands, N, h, w, kind, intermediates kept (None = all)
        ("gf2_small", 4, 1, 16, 16, "rand", ALL),
        ("wv3_small", 8, 1, 16, 16, "rand", ("prior_in", "enc0_global_in", "enc0_global", "bott_local", "bott", "out")),
        ("gf2_batch", 4, 3, 16, 16, "rand", ("out",)),
        ("gf2_rect", 4, 1, 16, 32, "rand", ("prior_in", "out")),
        ("gf2_metric", 4, 1, 32, 32, "scene", ("out",)),
        ("gf2_full", 4, 1, 64, 64, "rand", ("out",)),       # BASELINE.json configs[0]
    ]
    for name, bands, n, h, w, kind, keep in cases:
        gen = torch.Generator().manual_seed(0)
        extra = {}
        if kind == "rand":
            ms = torch.rand(n, bands, h, w, generator=gen)
            pan = torch.rand(n, 1, 4 * h, 4 * w, generator=gen)
        else:
            ms, pan, gt = smooth_scene(n, bands, 4 * h, gen)
            extra["gt"] = gt.numpy()
        got = capture(nets[bands], ms, pan)
        if keep is not None:                 # keep the fixtures small on disk
            got = {k: got[k] for k in keep}
        if kind == "scene":
            p = got["out"][0].permute(1, 2, 0).numpy() * 2047.5
            g = gt[0].permute(1, 2, 0).numpy() * 2047.5
            extra["ref_metrics_psnr_sam_ergas"] = np.array([mtc.psnr(p, g), mtc.sam(p, g), mtc.ergas(p, g)])
        arrays = {k: v.numpy() for k, v in got.items()}
        arrays.update(extra)
        arrays["ms"], arrays["pan"] = ms.numpy(), pan.numpy()
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
