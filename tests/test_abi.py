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


def _build_harness(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "harness", "abi_smoke.cpp"),
                           "-ldl", "-o", exe])
    return exe


def test_cpp_harness_loads_the_library_without_python(tmp_path):
    """A plain C++ program (tests/harness/abi_smoke.cpp: dlopen + include/muopdb_gpu.h, no torch, no CUDA headers) resolves the
    entry points, runs the host-only ones, and -- without a GPU -- sees MGPU_ERR_NO_DEVICE from mgpu_init."""
    import subprocess
    import torch
    exe = _build_harness(tmp_path)
    p = subprocess.run([exe, os.path.join(ROOT, "muopdb_b200", "libmuopdb_gpu.so")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    if not torch.cuda.is_available():
        assert p.stdout.strip() == "NO_DEVICE"


@pytest.mark.gpu
def test_cpp_harness_reference_golden_on_gpu(tmp_path):
    """The same program on a GPU box: the reference's golden case (spann/index.rs:335-366, 412-444) through the C ABI from C++,
    flat scores bit-identical to the inline restatement of l2.rs."""
    import subprocess
    exe = _build_harness(tmp_path)
    p = subprocess.run([exe, os.path.join(ROOT, "muopdb_b200", "libmuopdb_gpu.so")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.startswith("OK"), p.stdout + p.stderr


def test_rust_shim_binds_every_header_entry_point():
    """ffi/muopdb_gpu.rs: the extern block is generated from the header (tools/gen_rust_ffi.py) and must be in sync; the shim
    implements the reference traits the boundary keeps (Quantizer, DistanceCalculator, CalculateSquared)."""
    import subprocess
    import sys
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]).returncode == 0
    src = open(os.path.join(ROOT, "ffi", "muopdb_gpu.rs")).read()
    for s in _header_symbols():
        assert f"pub fn {s}(" in src, s
    for needle in ("impl quantization::quantization::Quantizer for GpuProductQuantizer", "impl utils::DistanceCalculator for $t",
                   "impl utils::CalculateSquared for $t", "pub fn get_doc_ids", "pub fn get_vector", "pub fn get_point_id"):
        assert needle in src, needle
