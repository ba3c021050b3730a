"""Shared test plumbing: fixtures, the oracle (reference stack compiled unmodified into
oracle/_ref/libmpidref.so), and the host-side driver of mpid_math.h.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "_build")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


import sys
sys.path.insert(0, ROOT)
from mpidopenmmplugin_b200.workloads import FlatSystem as System, water_box, make_kernel, subset_waters  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def load_fixture(name):
    with open(os.path.join(GOLDEN, "fixtures.json")) as f:
        d = json.load(f)[name]
    n = d["n"]
    s = System(n)
    s.pos = np.array(d["positions"], dtype=np.float64)
    s.covalent = [[[] for _ in range(8)] for _ in range(n)]
    for i, m in enumerate(d["multipoles"]):
        s.charges[i] = m["charge"]
        s.dipoles[i] = m["dipole"]
        s.quadrupoles[i] = m["quadrupole"]
        s.octopoles[i] = m["octopole"]
        s.axis[i] = m["axisType"]
        s.atomZ[i] = m["atomZ"]
        s.atomX[i] = m["atomX"]
        s.atomY[i] = m["atomY"]
        s.tholes[i] = m["thole"]
        s.alphas[i] = m["alpha"]
        s.covalent[i] = [list(c) for c in m["covalent"]]
    s.box = np.diag([d["box"]]*3)
    return s


# ---------------------------------------------------------------------------------------------------
# host driver of mpid_math.h
# ---------------------------------------------------------------------------------------------------
_emul = None


def emul_lib():
    global _emul
    if _emul is None:
        os.makedirs(BUILD, exist_ok=True)
        so = os.path.join(BUILD, "libemul.so")
        src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
        hdr = os.path.join(ROOT, "mpidopenmmplugin_b200", "csrc", "mpid_math.h")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", so])
        _emul = ctypes.CDLL(so)
    return _emul


def emul_evaluate(s, use_float=False):
    lib = emul_lib()
    off, idx = s.cov_csr()
    off = np.ascontiguousarray(off, dtype=np.int32); idx = np.ascontiguousarray(idx, dtype=np.int32)
    box = np.ascontiguousarray(s.box, dtype=np.float64).reshape(-1)
    coefs = np.ascontiguousarray(s.coefs, dtype=np.float64)
    e = ctypes.c_double()
    it = ctypes.c_int()
    f = np.zeros((s.n, 3))
    mu = np.zeros((s.n, 3))
    pos = np.ascontiguousarray(s.pos, dtype=np.float64)
    lib.emul_evaluate(ctypes.c_int(s.n), dp(pos), dp(s.charges), dp(np.ascontiguousarray(s.dipoles)),
                      dp(np.ascontiguousarray(s.quadrupoles)), dp(np.ascontiguousarray(s.octopoles)),
                      ip(s.axis), ip(s.atomZ), ip(s.atomX), ip(s.atomY), dp(s.tholes),
                      dp(np.ascontiguousarray(s.alphas)), ip(off), ip(idx),
                      ctypes.c_int(s.method), ctypes.c_int(s.polarization), ctypes.c_double(s.cutoff),
                      ctypes.c_double(s.alpha), ctypes.c_int(int(s.grid[0])), ctypes.c_int(int(s.grid[1])), ctypes.c_int(int(s.grid[2])),
                      ctypes.c_double(s.default_thole), ctypes.c_double(s.scale14), ctypes.c_int(s.max_iter),
                      ctypes.c_double(s.epsilon), ctypes.c_int(len(coefs)), dp(coefs), dp(box),
                      ctypes.c_int(1 if use_float else 0), ctypes.byref(e), dp(f), dp(mu), ctypes.byref(it))
    return e.value, f, mu, it.value


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b))/max(np.linalg.norm(np.asarray(b)), 1e-300))


# ---------------------------------------------------------------------------------------------------
# the reference's own test configurations (platforms/reference/tests/TestReferenceMPIDForce.cpp)
# ---------------------------------------------------------------------------------------------------
def water_dimer(method, polarization):
    """make_waterbox(6, 2.0): :109-660; PME settings :1456-1467."""
    s = load_fixture("water_dimer")
    s.method = method
    s.polarization = polarization
    s.default_thole = 3.0
    s.epsilon = 1e-8
    s.max_iter = 500
    s.cutoff = 0.6
    s.alpha = 3.0
    s.grid = (64, 64, 64)
    return s


def methanol_dimer(method, polarization):
    """make_methanolbox(12, 24.61817 A): :662-880; PME settings :1080-1095."""
    s = load_fixture("methanol_dimer")
    s.method = method
    s.polarization = polarization
    s.default_thole = 3.0
    s.epsilon = 1e-9
    s.max_iter = 500
    s.cutoff = 1.2
    s.alpha = 4.5
    s.grid = (64, 64, 64)
    return s
