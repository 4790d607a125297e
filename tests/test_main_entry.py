"""The reference's UNMODIFIED entry point (`main.py -c <config>`, main.py:61-146) end to end: config -> datasets ->
`build_model('UnlgFormer')` -> `load_checkpoint` -> `runner.test(ref=True)` (models/base/base_model.py:267-352), with only
the five missing third-party modules shimmed (shims/) and a config that overrides paths.

  * CPU (here): the reference's own `Pansharpening` class on the host reproduces tests/golden/main_entry_metrics.json
    (recorded by tests/golden/make_golden_main.py from the same run) — the shims, fixture dataset and checkpoint are sound.
  * GPU (`-m gpu`): the same command after `lgteun_b200.install()` (the CUDA module behind the registry, cfg.cuda = True as
    shipped) logs PSNR / SAM / ERGAS within 0.01 of the reference-on-CPU run (north_star's bar), and SSIM / Q within 0.01.

The unmodified reference lives under git-ignored baseline/_ref/ (tools/vendor_reference.py, run by build()); the tests are
skipped where it is absent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden", "main_entry_metrics.json")
have_ref = os.path.isfile(os.path.join(REF, "main.py"))


def run_main(tmp, *flags):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_main.py"), "--ref-root", REF, "--workdir",
                        str(tmp), *flags], capture_output=True, text=True, timeout=1500)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("MAIN_RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(lines[-1][len("MAIN_RESULT "):])


def check(res, gold, tol):
    assert res["finished"] and not res["errors"], res
    assert res["outputs"] == [f"{i}_mul_hat.tif" for i in range(4)]          # save_image through the gdal shim
    for m in ("PSNR", "SSIM", "Q", "SAM", "ERGAS"):
        assert abs(res["metrics"][m][0] - gold["metrics"][m][0]) <= tol, (m, res["metrics"][m], gold["metrics"][m])


@pytest.mark.skipif(not have_ref, reason="baseline/_ref not vendored (run tools/vendor_reference.py where /root/reference exists)")
def test_unmodified_main_runs_on_cpu_with_the_reference_class(tmp_path):
    res = run_main(tmp_path, "--cpu")
    assert res["core_class"] == "models.unlg_former.Pansharpening" and res["lgteun_modules"] == 0
    check(res, json.load(open(GOLD)), 2e-4)


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref, reason="baseline/_ref not vendored")
def test_unmodified_main_runs_on_the_cuda_module(tmp_path):
    res = run_main(tmp_path, "--install")
    assert res["core_class"] == "lgteun_b200.module.Pansharpening" and res["installed"] == res["core_class"]
    check(res, json.load(open(GOLD)), 0.01)
