"""Golden fixture of ONE training step of the UNMODIFIED reference (container only; /root/reference is not on the GPU box):
    python tests/golden/make_golden_train.py
Runs models.unlg_former.Pansharpening in train() mode exactly as UnlgFormer.train_iter does (models/unlg_former.py:87-110):
out = G(lr, pan); loss = nn.L1Loss()(out, gt) * 1.0; loss.backward(); Adam(lr=1.5e-3).step() (configs/unlg_former.py:82-90).
The five nn.Dropout(0.1) masks of the live prior (LGT.py:198,216) are captured with forward hooks so that the oracle and
the CUDA path can replay the same step.  Writes tests/golden/train_gf2.npz (4 bands) and train_wv3.npz (8 bands): inputs, masks
(NHWC), output, loss, every gradient (absent key = .grad is None) and the parameters after the Adam step."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_import  # noqa: E402


def record(bands, n, h, fname, after_keys=None):
    """after_keys: None = store every updated parameter, else only these (keeps the 8-band fixture small)."""
    torch.set_num_threads(1)
    net = ref_import.build(bands, stages=2, seed=19971118)
    net.train()
    gen = torch.Generator().manual_seed(1)
    ms = torch.rand(n, bands, h, h, generator=gen)
    pan = torch.rand(n, 1, 4 * h, 4 * h, generator=gen)
    gt = torch.rand(n, bands, 4 * h, 4 * h, generator=gen)
    masks = []
    hooks = []
    prior = net.prior_module[-1]
    for m in prior.modules():
        if isinstance(m, torch.nn.Dropout):
            # proj_drop sees NCHW [N,c,H,W]; mask = out / in where in != 0 (values 0 or 1/0.9)
            def hook(mod, inp, out):
                x = inp[0]
                mk = torch.where(x != 0, out / x, torch.full_like(x, float("nan")))
                masks.append(mk.detach().permute(0, 2, 3, 1).contiguous())
            hooks.append(m.register_forward_hook(hook))
    torch.manual_seed(7)
    opt = torch.optim.Adam(net.parameters(), betas=(0.9, 0.999), lr=1.5e-3)
    out = net(ms, pan)
    loss = torch.nn.L1Loss()(out, gt) * 1.0
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
    opt.step()
    for hk in hooks:
        hk.remove()
    assert len(masks) == 5, len(masks)
    arrays = {"ms": ms.numpy(), "pan": pan.numpy(), "gt": gt.numpy(), "out": out.detach().numpy(),
              "loss": np.array(loss.item(), dtype=np.float64)}
    for i, mk in enumerate(masks):
        assert not torch.isnan(mk).any()
        vals = torch.unique(mk)
        assert all(abs(v.item()) < 1e-6 or abs(v.item() - 1 / 0.9) < 1e-4 for v in vals), vals
        arrays[f"mask{i}"] = (mk > 0.5).numpy()          # keep flags; the scale is 1/(1-p) with p = 0.1
    for k, g in grads.items():
        arrays["grad/" + k] = g.numpy()
    for k, p in net.named_parameters():
        if k in grads and (after_keys is None or k in after_keys):
            arrays["after/" + k] = p.detach().numpy()
    np.savez_compressed(os.path.join(HERE, fname), **arrays)
    print(fname, "loss", loss.item(), "live grads", len(grads), "of", len(list(net.parameters())))


AFTER_KEYS_WV = ("eta.1", "R.weight", "prior_module.1.tail.1.weight",
                 "prior_module.1.bottleneck.blocks.0.0.fn.fn.local_mixer.pos_emb")


def main():
    record(4, 2, 8, "train_gf2.npz")                      # GF-2 / WV-2 band count
    record(8, 1, 8, "train_wv3.npz", AFTER_KEYS_WV)       # 8 bands (BASELINE configs[4]); batch 1 keeps the fixture small


if __name__ == "__main__":
    main()
