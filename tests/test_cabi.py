"""The product library: builds for sm_100a, loads, and exports every symbol include/lbc_b200.h declares.
No compute is called here (no GPU in the CPU container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "lbc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lbc_[a-z0-9_]+)\s*\(", txt)))


def test_cuda_library_exports_header():
    from learningbycheating_b200 import build
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "liblbc_b200.so does not export %s" % s
    lib.lbc_device_kind.restype = ctypes.c_int
    assert lib.lbc_device_kind() == 1
    lib.lbc_build_info.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.lbc_build_info()


def test_binding_covers_header():
    from learningbycheating_b200 import _lib, build
    import subprocess, sys
    # run in a subprocess: the session may be bound to the host-emulation build
    code = ("import sys; sys.path.insert(0, %r); from learningbycheating_b200 import _lib; _lib.lib(); "
            "print(' '.join(_lib.EXPORTED_SYMBOLS))" % ROOT)
    build.build_cuda()
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert sorted(out.stdout.split()) == _header_symbols()


def test_no_silent_cpu_path_without_gpu():
    """On a machine without a GPU the product library must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import subprocess, sys
    code = ("import sys, ctypes; sys.path.insert(0, %r); from learningbycheating_b200 import _lib; L=_lib.lib(); "
            "h=ctypes.c_void_p(); rc=L.lbc_net_create(0,0,2,ctypes.byref(h)); print(rc, L.lbc_last_error().decode())" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("1 ") and "CUDA device" in out.stdout


def test_hot_path_kernels_are_in_the_binary():
    """cuobjdump lists the sm_100a kernels the engine launches (no PTX-only / JIT dependence)."""
    import shutil, subprocess
    from learningbycheating_b200 import build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    path = build.build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
