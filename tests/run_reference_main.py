"""Run the reference's UNMODIFIED entry point (`main.py -c <config>`, main.py:136-146) in a fresh interpreter.

    python tests/run_reference_main.py --ref-root baseline/_ref --workdir DIR [--install] [--cpu]

Builds a tiny synthetic WV-3-shaped dataset ({id}_lr/_pan/_mul.tif, 11-bit), a checkpoint that pickles the reference's own
`Pansharpening` module with the golden weights (the format `Base_model.load_checkpoint` reads, base_model.py:99-106) and a
config that inherits `configs/unlg_former.py` unchanged and overrides PATHS only (plus `cuda = False` with --cpu), puts
shims/ (mmcv, gdal, osr, tifffile, pywt stand-ins) behind the reference root on sys.path, optionally calls
`lgteun_b200.install()` (rebinding `models.unlg_former.Pansharpening` to the CUDA module), and then executes main.py as
`__main__`.  Prints one JSON line with the metric values main.py logged.  TEST INFRASTRUCTURE."""
import argparse
import json
import os
import re
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_dataset(d, bands=8, n=4, h=64, seed=2023):
    import numpy as np
    import tifffile
    os.makedirs(d, exist_ok=True)
    rng = np.random.RandomState(seed)
    H = 4 * h
    yy, xx = np.mgrid[0:H, 0:H].astype(np.float64) / H
    for i in range(n):
        chans = []
        for b in range(bands):
            f = rng.uniform(1.0, 6.0, size=4)
            p = rng.uniform(0, 6.28, size=4)
            img = 0.5 + 0.2 * np.sin(6.28 * f[0] * xx + p[0]) * np.cos(6.28 * f[1] * yy + p[1]) \
                + 0.15 * np.sin(6.28 * (f[2] * xx + f[3] * yy) + p[2]) + 0.05 * rng.rand(H, H)
            chans.append(np.clip(img, 0.02, 0.98))
        mul = np.stack(chans, axis=-1)                                       # [H,W,B] in (0,1)
        pan = mul.mean(axis=-1)
        lr = mul.reshape(h, 4, h, 4, bands).mean(axis=(1, 3))               # box-filtered LrMS
        to16 = lambda a: np.round(a * 2047.0).astype(np.uint16)
        tifffile.imwrite(os.path.join(d, f"{i}_mul.tif"), to16(mul))
        tifffile.imwrite(os.path.join(d, f"{i}_pan.tif"), to16(pan))
        tifffile.imwrite(os.path.join(d, f"{i}_lr.tif"), to16(lr))


def make_checkpoint(path, bands=8):
    """{'core_module': <reference Pansharpening with the golden weights>, 'iter_num': 30000}, pickled whole like
    Base_model.save does (base_model.py:355-368)."""
    import numpy as np
    import torch
    import models.unlg_former as ul
    ref_cls = getattr(ul, "_reference_Pansharpening", None) or ul.Pansharpening
    from mmcv import Config
    net = ref_cls(cfg=Config({"ms_chans": bands}), logger=None, stage=2)
    z = np.load(os.path.join(ROOT, "tests", "golden", f"weights_b{bands}.npz"))
    net.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
    torch.save({"core_module": net, "iter_num": 30000}, path)


def write_config(path, ref_root, work, cpu):
    data = os.path.join(work, "data")
    with open(path, "w") as f:
        f.write(f"""# paths-only override of the reference's shipped config (inherited unchanged)
_base_ = {os.path.join(ref_root, 'configs', 'unlg_former.py')!r}
currentPath = {work!r}
work_dir = {os.path.join(work, 'model_out', 'LGTEUN')!r}
log_dir = {os.path.join(work, 'logs')!r}
log_file = {os.path.join(work, 'logs', 'LGTEUN.log')!r}
checkpoint = {os.path.join(work, 'model_iter_30000.pth')!r}
train_set_cfg = dict(dataset=dict(image_dirs=[{data!r}]), num_workers=0)
test_set0_cfg = dict(dataset=dict(image_dirs=[{data!r}]))
test_set1_cfg = dict(dataset=dict(image_dirs=[{data!r}]))
""")
        if cpu:
            f.write("cuda = False\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-root", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--install", action="store_true", help="lgteun_b200.install(): the CUDA module behind the registry")
    ap.add_argument("--cpu", action="store_true", help="cuda = False (the reference class on the host)")
    args = ap.parse_args()
    ref_root = os.path.abspath(args.ref_root)
    work = os.path.abspath(args.workdir)
    os.makedirs(work, exist_ok=True)
    # torch >= 2.6 defaults torch.load to weights_only=True; the reference's checkpoints are pickled modules
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
    for p in (os.path.join(ROOT, "shims"), ROOT, ref_root):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)                     # reference root first, then the repo, then the shims
    os.chdir(ref_root)
    import warnings
    warnings.filterwarnings("ignore")
    make_dataset(os.path.join(work, "data"))
    make_checkpoint(os.path.join(work, "model_iter_30000.pth"))
    cfg_path = os.path.join(work, "override_config.py")
    write_config(cfg_path, ref_root, work, args.cpu)
    core = None
    if args.install:
        import lgteun_b200
        mod = lgteun_b200.install()
        core = mod.Pansharpening.__module__ + "." + mod.Pansharpening.__name__
    sys.argv = ["main.py", "-c", cfg_path]
    runpy.run_path(os.path.join(ref_root, "main.py"), run_name="__main__")
    log = open(os.path.join(work, "logs", "LGTEUN.log")).read()
    vals = {m: (float(a), float(b)) for m, a, b in re.findall(r"(\w+) metric value: ([-\d.einf]+) \+- ([-\d.einfa]+)", log)}
    err = re.findall(r"ERROR - (.*)", log)
    import models.unlg_former as ul
    loaded = sorted(m for m in sys.modules if m.startswith("lgteun_b200"))
    print("MAIN_RESULT " + json.dumps({"metrics": vals, "finished": "Finish !!!" in log, "errors": err[:3],
                                       "core_class": ul.Pansharpening.__module__ + "." + ul.Pansharpening.__name__,
                                       "installed": core, "lgteun_modules": len(loaded),
                                       "outputs": sorted(os.listdir(os.path.join(work, "model_out", "LGTEUN", "WV-3", "test_out1",
                                                                                 "iter_35000")))
                                       if os.path.isdir(os.path.join(work, "model_out", "LGTEUN", "WV-3", "test_out1", "iter_35000")) else []}),
          flush=True)


if __name__ == "__main__":
    main()
