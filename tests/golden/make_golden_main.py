"""Records what the reference's unmodified main.py logs on the synthetic fixture with its OWN Pansharpening class on the
CPU (tests/run_reference_main.py --cpu): tests/golden/main_entry_metrics.json.  Needs /root/reference (build container)."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.vendor_reference import vendor  # noqa: E402

if __name__ == "__main__":
    assert vendor(), "reference tree not present"
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_main.py"), "--workdir", tmp, "--cpu"],
                           capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("MAIN_RESULT ")][-1]
    res = json.loads(line[len("MAIN_RESULT "):])
    assert res["finished"] and not res["errors"], res
    res["how"] = "python tests/run_reference_main.py --cpu: unmodified /root/reference main.py + configs/unlg_former.py (index 2: WV-3), " \
                 "reference Pansharpening on the CPU, 4 synthetic 8-band pairs, golden weights_b8"
    with open(os.path.join(ROOT, "tests", "golden", "main_entry_metrics.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(res)
