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


# ---------------------------------------------------------------------------------------------------
# the oracle's pair set, restated with numpy for the bit-exactness tests
# ---------------------------------------------------------------------------------------------------
def pair_classes(s):
    """{(i,j): class} for i<j from the covalent maps, exactly as setupScaleMaps
    (MPIDReferenceForce.cpp:190-225): 1 = scale 0 (1-2, 1-3), 2 = 1-4, later lists override earlier ones."""
    off, idx = s.cov_csr()
    n = s.n
    out = {}
    for t in range(4):
        for i in range(n):
            for j in idx[off[t*(n+1)+i]:off[t*(n+1)+i+1]]:
                j = int(j)
                if j <= i:
                    continue
                out[(i, j)] = 1 if t < 2 else (2 if t == 2 else 0)
    return {k: v for k, v in out.items() if v != 0}


def min_image(s, d):
    """getPeriodicDelta (MPIDReferenceForce.cpp:2671-2676) with separate multiply and add roundings."""
    d = d.copy()
    a, b, c = s.box[0], s.box[1], s.box[2]
    det = a[0]*b[1]*c[2]
    rc2 = (a[0]*b[1])*(1.0/det)
    rb1 = (a[0]*c[2])*(1.0/det)
    ra0 = (b[1]*c[2])*(1.0/det)
    k = np.floor(d[:, 2]*rc2 + 0.5)
    d -= k[:, None]*c[None, :]
    k = np.floor(d[:, 1]*rb1 + 0.5)
    d -= k[:, None]*b[None, :]
    k = np.floor(d[:, 0]*ra0 + 0.5)
    d -= k[:, None]*a[None, :]
    return d


def pair_set_reference(s, chunk=512):
    """[(i, j, class)] with i<j and |minimg(r_j - r_i)|^2 <= rc^2 in FP64 (all pairs when no cutoff)."""
    cls = pair_classes(s)
    n = s.n
    out = []
    rc2 = s.cutoff*s.cutoff
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        d = s.pos[None, :, :] - s.pos[i0:i1, None, :]
        if s.method == 1:
            d = min_image(s, d.reshape(-1, 3)).reshape(i1 - i0, n, 3)
        r2 = d[:, :, 0]*d[:, :, 0] + d[:, :, 1]*d[:, :, 1] + d[:, :, 2]*d[:, :, 2]
        ok = (r2 <= rc2) if s.method == 1 else np.ones_like(r2, dtype=bool)
        ii, jj = np.nonzero(ok)
        for a, b in zip(ii + i0, jj):
            if a < b:
                out.append((int(a), int(b), cls.get((int(a), int(b)), 0)))
    return out
