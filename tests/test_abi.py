"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what include/muopdb_gpu.h
declares, fails loudly without a device, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "muopdb_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from muopdb_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/muopdb_gpu.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "python binding table and header disagree"


def test_version_and_no_device_behaviour():
    import torch
    from muopdb_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.mgpu_version()
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.mgpu_init(0, ctypes.byref(h)) == _lib.ERR_NO_DEVICE  # no CPU fallback
        import muopdb_b200 as M
        with pytest.raises(M.NoDevice):
            M.Context(0)


def test_product_package_does_not_use_the_oracle():
    pkg = os.path.join(ROOT, "muopdb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


def test_sass_is_sm100a():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "muopdb_b200", "libmuopdb_gpu.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out
