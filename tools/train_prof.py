"""Per-kernel device time of the training step (CUPTI through torch.profiler; launch-by-launch path, warm caches).
Usage: python tools/train_prof.py [--pan 256] [--bands 8] [--batch 4] [--steps 5] [--top 40]"""
import argparse
import collections
import os
import sys
from types import SimpleNamespace

os.environ["LGTEUN_TRAIN_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import lgteun_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pan", type=int, default=256)
    ap.add_argument("--bands", type=int, default=8)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    h = a.pan // 4
    torch.manual_seed(0)
    net = lgteun_b200.Pansharpening(SimpleNamespace(ms_chans=a.bands), None, stage=2).cuda().train()
    tr = lgteun_b200.Trainer(net, lr=1.5e-3, dropout_p=0.1)
    ms, pan, gt = torch.rand(a.batch, a.bands, h, h).cuda(), torch.rand(a.batch, 1, a.pan, a.pan).cuda(), torch.rand(a.batch, a.bands, a.pan, a.pan).cuda()
    for _ in range(3):
        tr.step(ms, pan, gt)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            tr.step(ms, pan, gt)
        torch.cuda.synchronize()
    tot = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            t = tot[e.name[:70]]
            t[0] += 1
            t[1] += e.device_time
    total = sum(v[1] for v in tot.values())
    print(f"# {a.steps} steps, PAN {a.pan}, {a.bands} bands, batch {a.batch}: kernel time {total / a.steps / 1000:.3f} ms per step")
    print(f"{'kernel':72s} {'n/step':>7s} {'us/step':>9s} {'share':>6s} {'avg_us':>8s}")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"{k:72s} {n / a.steps:7.1f} {us / a.steps:9.1f} {100 * us / total:5.1f}% {us / n:8.1f}")


if __name__ == "__main__":
    main()
