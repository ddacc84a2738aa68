"""Host-emulation unit checks of data-movement kernels (bit-exact on the CPU)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pair_walking_weight_pack_equals_element_walk(tmp_path):
    exe = str(tmp_path / "pack_check")
    cmd = ["g++", "-O1", "-std=c++17", "-fopenmp", "-DLBC_HOST_EMU", "-I", os.path.join(ROOT, "learningbycheating_b200", "csrc"),
           os.path.join(ROOT, "tests", "hostemu", "pack_check.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pack_all_pairs == pack_all" in r.stdout


def test_pair_walking_pack_in_the_engine_cpu(backend):
    """with the fast kernels on, the per-forward weight pack runs pack_all_pairs (off: the element walk): same predictions."""
    if backend != "cpu":
        pytest.skip("host-emulation check")
    import learningbycheating_b200 as lbc
    from learningbycheating_b200 import _lib
    from lbc_testing import build_models, batch_on
    L = _lib.lib()
    outs = []
    try:
        for bits in (0, 1):
            _lib.check(L.lbc_set_fast_kernels(bits))
            s, _ = build_models(backend, "fp32")
            s.eval()
            b = batch_on(backend, 2)
            with torch.no_grad():
                outs.append(s(b["rgb"], b["speed"], lbc.one_hot(b["command"].cpu()).to(backend))[1].clone())
    finally:
        _lib.check(L.lbc_set_fast_kernels(1))
    assert torch.equal(outs[0], outs[1])
