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


class System:
    """Flat-array description of one MPIDForce system, in the orderings of
    MPIDForce::getMultipoleParameters (reference openmmapi/include/openmm/MPIDForce.h:262-290)."""

    def __init__(self, n):
        self.n = n
        self.pos = np.zeros((n, 3))
        self.charges = np.zeros(n)
        self.dipoles = np.zeros((n, 3))
        self.quadrupoles = np.zeros((n, 6))
        self.octopoles = np.zeros((n, 10))
        self.axis = np.full(n, 5, dtype=np.int32)
        self.atomZ = np.full(n, -1, dtype=np.int32)
        self.atomX = np.full(n, -1, dtype=np.int32)
        self.atomY = np.full(n, -1, dtype=np.int32)
        self.tholes = np.zeros(n)
        self.alphas = np.zeros((n, 3))
        self.covalent = [[[] for _ in range(8)] for _ in range(n)]
        self.box = np.diag([2.0, 2.0, 2.0])
        # method / parameters
        self.method = 0           # 0 NoCutoff, 1 PME
        self.polarization = 0     # 0 Mutual, 1 Direct, 2 Extrapolated
        self.cutoff = 1.0
        self.alpha = 0.0
        self.grid = (0, 0, 0)
        self.ewald_tol = 5e-4
        self.default_thole = 5.0
        self.scale14 = 1.0
        self.max_iter = 60
        self.epsilon = 1e-5
        self.coefs = np.array([-0.154, 0.017, 0.658, 0.474])

    def cov_csr(self):
        n = self.n
        offsets = np.zeros(8*(n+1), dtype=np.int32)
        idx = []
        for t in range(8):
            for i in range(n):
                offsets[t*(n+1)+i] = len(idx)
                idx.extend(self.covalent[i][t])
            offsets[t*(n+1)+n] = len(idx)
        return offsets, np.array(idx if idx else [0], dtype=np.int32)

    def copy(self):
        import copy
        return copy.deepcopy(self)


def load_fixture(name):
    with open(os.path.join(GOLDEN, "fixtures.json")) as f:
        d = json.load(f)[name]
    n = d["n"]
    s = System(n)
    s.pos = np.array(d["positions"], dtype=np.float64)
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
# oracle
# ---------------------------------------------------------------------------------------------------
_ref = None


def oracle_lib():
    global _ref
    if _ref is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libmpidref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference"):
                subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j4"], stdout=subprocess.DEVNULL)
            else:
                raise RuntimeError("oracle/_ref/libmpidref.so missing and /root/reference not present")
        _ref = ctypes.CDLL(path)
        _ref.mpidref_last_error.restype = ctypes.c_char_p
    return _ref


class Oracle:
    """The reference's own MPIDForce -> Reference-platform stack behind flat C calls (oracle/ref_driver.cpp)."""

    def __init__(self, s):
        lib = oracle_lib()
        self.lib = lib
        self.s = s
        off, idx = s.cov_csr()
        h = ctypes.c_void_p()
        box = np.ascontiguousarray(s.box, dtype=np.float64).reshape(-1)
        coefs = np.ascontiguousarray(s.coefs, dtype=np.float64)
        rc = lib.mpidref_create(ctypes.c_int(s.n), dp(s.charges), dp(np.ascontiguousarray(s.dipoles)),
                                dp(np.ascontiguousarray(s.quadrupoles)), dp(np.ascontiguousarray(s.octopoles)),
                                ip(s.axis), ip(s.atomZ), ip(s.atomX), ip(s.atomY), dp(s.tholes),
                                dp(np.ascontiguousarray(s.alphas)), ip(off), ip(idx),
                                ctypes.c_int(s.method), ctypes.c_int(s.polarization), ctypes.c_double(s.cutoff),
                                ctypes.c_double(s.alpha), ctypes.c_int(s.grid[0]), ctypes.c_int(s.grid[1]),
                                ctypes.c_int(s.grid[2]), ctypes.c_double(s.ewald_tol), ctypes.c_double(s.default_thole),
                                ctypes.c_double(s.scale14), ctypes.c_int(s.max_iter), ctypes.c_double(s.epsilon),
                                ctypes.c_int(len(coefs)), dp(coefs), dp(box), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(lib.mpidref_last_error().decode())
        self.h = h

    def execute(self, pos=None):
        pos = np.ascontiguousarray(self.s.pos if pos is None else pos, dtype=np.float64)
        e = ctypes.c_double()
        f = np.zeros((self.s.n, 3))
        rc = self.lib.mpidref_execute(self.h, dp(pos), ctypes.byref(e), dp(f))
        if rc != 0:
            raise RuntimeError(self.lib.mpidref_last_error().decode())
        return e.value, f

    def dipoles(self, which=0, pos=None):
        pos = np.ascontiguousarray(self.s.pos if pos is None else pos, dtype=np.float64)
        out = np.zeros((self.s.n, 3))
        rc = self.lib.mpidref_get_dipoles(self.h, dp(pos), ctypes.c_int(which), dp(out))
        if rc != 0:
            raise RuntimeError(self.lib.mpidref_last_error().decode())
        return out

    def close(self):
        if self.h:
            self.lib.mpidref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
                      ctypes.c_double(s.alpha), ctypes.c_int(s.grid[0]), ctypes.c_int(s.grid[1]), ctypes.c_int(s.grid[2]),
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
