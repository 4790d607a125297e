"""Place an UNMODIFIED copy of the reference's python entry point under baseline/_ref/ (git-ignored, travels to the GPU
box with the snapshot) so that `tests/test_main_entry.py` can run the reference's own `main.py` there.

The reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` (the base contract's
install) has nothing to build; this script is that install: main.py, configs/, models/, dataset/ byte for byte, nothing
else, never into tracked paths.  Run by `__graft_entry__.build()` whenever /root/reference is present."""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LGTEUN_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ["main.py", "configs", "models", "dataset"]


def vendor(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "main.py")):
        if verbose:
            print(f"[vendor_reference] {SRC} not present: nothing to do")
        return False
    os.makedirs(DST, exist_ok=True)
    n = 0
    for part in PARTS:
        s = os.path.join(SRC, part)
        if os.path.isfile(s):
            shutil.copy2(s, os.path.join(DST, part))
            n += 1
            continue
        for base, dirs, files in os.walk(s):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            rel = os.path.relpath(base, SRC)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    src, dst = os.path.join(base, f), os.path.join(DST, rel, f)
                    if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                        shutil.copy2(src, dst)
                    n += 1
    if verbose:
        print(f"[vendor_reference] {n} files of the unmodified reference under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
