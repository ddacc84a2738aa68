"""Build recipes for the native library (explicit nvcc, in-tree outputs).

  build_cuda()    -> learningbycheating_b200/liblbc_b200.so   (sm_100a; THE product library)
  build_hostemu() -> tests/hostemu/liblbc_hostemu.so          (g++ host emulation of the
                     correctness-first kernels; only `pytest -m "not gpu"` loads it)
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SOURCES = ["lbc_capi.cu", "lbc_net.cu", "lbc_fast.cu", "lbc_fast_conv.cu", "lbc_fast_elem.cu", "lbc_fast_head.cu"]
CUDA_LIB = os.path.join(PKG, "liblbc_b200.so")
EMU_LIB = os.path.join(ROOT, "tests", "hostemu", "liblbc_hostemu.so")


def _digest(deps):
    """Content hash of the sources (mtimes do not survive the snapshot that carries the tree to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(deps):
        if d.endswith((".o", ".so")):
            continue
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _newer(target, deps):
    """True when `target` is missing or was not built from the current contents of `deps`."""
    stamp = target + ".srchash"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != _digest(deps)


def _stamp(target, deps):
    with open(target + ".srchash", "w") as f:
        f.write(_digest(deps))


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(ROOT, "include", "lbc_b200.h"))
    return d


def _run(cmd):
    print("+ " + " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_cuda(force=False, verbose=False):
    if not force and not _newer(CUDA_LIB, _deps()):
        return CUDA_LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
               "--extended-lambda", "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        out = _run(cmd)
        if verbose:
            print(out)
        objs.append(obj)
    _run([nvcc, "-shared", "-o", CUDA_LIB] + objs)   # static cudart (nvcc default)
    _stamp(CUDA_LIB, _deps())
    return CUDA_LIB


def build_hostemu(force=False):
    if not force and not _newer(EMU_LIB, _deps()):
        return EMU_LIB
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-DLBC_HOST_EMU", "-x", "c++"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", EMU_LIB]
    _run(cmd)
    _stamp(EMU_LIB, _deps())
    return EMU_LIB


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("cuda", "all"):
        print(build_cuda(force=True, verbose="-v" in sys.argv))
    if what in ("hostemu", "all"):
        print(build_hostemu(force=True))
