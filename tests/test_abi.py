"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/lgteun.h
declares (no compute calls without a GPU), and the ctypes signatures cover every declaration."""
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "lgteun.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgteun_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from lgteun_b200 import _abi, build
    build.build()                      # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    return _abi.lib()


def test_header_declares_the_forward_boundary():
    names = declared_functions()
    for must in ("lgteun_create", "lgteun_destroy", "lgteun_load_weights", "lgteun_forward", "lgteun_forward_host",
                 "lgteun_last_error", "lgteun_op_data_step", "lgteun_op_ffn", "lgteun_op_local_mixer",
                 "lgteun_op_global_mixer"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    for name in declared_functions():
        assert hasattr(built_lib, name), f"{name} declared in include/lgteun.h but not exported"


def test_ctypes_signatures_cover_the_header(built_lib):
    from lgteun_b200 import _abi
    assert sorted(_abi.SIGNATURES) == declared_functions()
    assert built_lib.lgteun_abi_version() == 1


def test_library_is_sm100a_native():
    """The shipped cubin is sm_100a (no PTX-JIT of some other arch)."""
    import shutil
    import subprocess
    from lgteun_b200 import _abi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out
