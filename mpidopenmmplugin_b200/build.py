"""Builds the sm_100a shared library (C ABI of include/mpidb200.h) in-tree with nvcc.

`python -m mpidopenmmplugin_b200.build` or `build()`; nvcc cross-compiles without a GPU.  The library is
git-ignored but travels to the GPU box with the repo snapshot."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmpidb200.so")
SOURCES = ["mpid_engine.cu"]
HEADERS = ["mpid_math.h", "mpid_kernels.cuh", "mpid_fft.cuh", os.path.join("..", "..", "include", "mpidb200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


RECORD = os.path.join(HERE, "build_record.json")     # what the last build() call did (git-ignored)


def source_hash():
    """SHA-256 over the CUDA sources and headers the library is compiled from."""
    import hashlib
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    """True unless the library exists AND was compiled from exactly the sources that are in the tree now (the hash of
    the sources is stored beside the library; file times do not survive a snapshot copy)."""
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".srchash"):
        return True
    with open(LIB + ".srchash") as fh:
        return fh.read().strip() != source_hash()


def _record(mode, seconds=0.0):
    import json
    import time
    with open(RECORD, "w") as fh:
        json.dump(dict(build_mode=mode, seconds=seconds, source_hash=source_hash(), when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
                       library=os.path.basename(LIB)), fh)


def build(force=False, verbose=False):
    force = force or os.environ.get("MPIDB200_FORCE_BUILD") == "1"
    if not force and not needs_build():
        _record("up-to-date (library matches the source hash)")
        return LIB
    import time
    t0 = time.time()
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    cmd += ["-o", LIB, "-lcufft", "-ldl", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    with open(LIB + ".srchash", "w") as fh:
        fh.write(source_hash())
    _record("compiled with nvcc -gencode arch=compute_100a,code=sm_100a", time.time() - t0)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
