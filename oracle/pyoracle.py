"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the parity oracle.

oracle/_ref/libmpidref.so is the reference's own MPIDForce -> MPIDForceImpl -> Reference-platform stack,
compiled unmodified by oracle/Makefile from the sources under /root/reference (see oracle/ref_driver.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def build():
    """Compile the oracle when the reference tree is present (build container only)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libmpidref.so")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j4"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libmpidref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libmpidref.so missing and /root/reference not present")
        _lib = ctypes.CDLL(path)
        _lib.mpidref_last_error.restype = ctypes.c_char_p
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class Oracle:
    def __init__(self, s):
        L = lib()
        self.lib = L
        self.s = s
        off, idx = s.cov_csr()
        off = np.ascontiguousarray(off, dtype=np.int32); idx = np.ascontiguousarray(idx, dtype=np.int32)
        h = ctypes.c_void_p()
        box = np.ascontiguousarray(s.box, dtype=np.float64).reshape(-1)
        coefs = np.ascontiguousarray(s.coefs, dtype=np.float64)
        c = np.ascontiguousarray
        # The reference maps the torque of every framed site through particleData[atomX]
        # (MPIDReferenceForce.cpp:2124-2127).  For a ZOnly site without an x anchor that is particleData[-1]: an
        # out-of-range read whose value depends on the heap (we have seen it produce NaN forces).  ZOnly frames do
        # not use the x anchor, and for the axially symmetric sites the reference's fixtures use the mapping does not
        # depend on the direction, so hand the oracle a valid placeholder to make it deterministic.
        atomX = np.array(s.atomX, dtype=np.int32, copy=True)
        for i in np.nonzero((np.asarray(s.axis) == 4) & (atomX < 0))[0]:
            for j in range(s.n):
                if j != i and j != s.atomZ[i]:
                    atomX[i] = j
                    break
        rc = L.mpidref_create(ctypes.c_int(s.n), _dp(c(s.charges)), _dp(c(s.dipoles)), _dp(c(s.quadrupoles)), _dp(c(s.octopoles)),
                              _ip(c(s.axis)), _ip(c(s.atomZ)), _ip(atomX), _ip(c(s.atomY)), _dp(c(s.tholes)), _dp(c(s.alphas)),
                              _ip(off), _ip(idx), ctypes.c_int(s.method), ctypes.c_int(s.polarization), ctypes.c_double(s.cutoff),
                              ctypes.c_double(s.alpha), ctypes.c_int(int(s.grid[0])), ctypes.c_int(int(s.grid[1])), ctypes.c_int(int(s.grid[2])),
                              ctypes.c_double(s.ewald_tol), ctypes.c_double(s.default_thole), ctypes.c_double(s.scale14),
                              ctypes.c_int(s.max_iter), ctypes.c_double(s.epsilon), ctypes.c_int(len(coefs)), _dp(coefs), _dp(box),
                              ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(L.mpidref_last_error().decode())
        self.h = h

    def execute(self, pos=None):
        pos = np.ascontiguousarray(self.s.pos if pos is None else pos, dtype=np.float64)
        e = ctypes.c_double()
        f = np.zeros((self.s.n, 3))
        if self.lib.mpidref_execute(self.h, _dp(pos), ctypes.byref(e), _dp(f)) != 0:
            raise RuntimeError(self.lib.mpidref_last_error().decode())
        return e.value, f

    def dipoles(self, which=0, pos=None):
        pos = np.ascontiguousarray(self.s.pos if pos is None else pos, dtype=np.float64)
        out = np.zeros((self.s.n, 3))
        if self.lib.mpidref_get_dipoles(self.h, _dp(pos), ctypes.c_int(which), _dp(out)) != 0:
            raise RuntimeError(self.lib.mpidref_last_error().decode())
        return out

    def pme_parameters(self):
        a = ctypes.c_double(); nx = ctypes.c_int(); ny = ctypes.c_int(); nz = ctypes.c_int()
        if self.lib.mpidref_get_pme_parameters(self.h, ctypes.byref(a), ctypes.byref(nx), ctypes.byref(ny), ctypes.byref(nz)) != 0:
            raise RuntimeError(self.lib.mpidref_last_error().decode())
        return a.value, nx.value, ny.value, nz.value

    def system_moments(self, pos=None):
        pos = np.ascontiguousarray(self.s.pos if pos is None else pos, dtype=np.float64)
        out = np.zeros(13)
        if self.lib.mpidref_get_system_multipole_moments(self.h, _dp(pos), _dp(out)) != 0:
            raise RuntimeError(self.lib.mpidref_last_error().decode())
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpidref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CellOracle:
    """The reference's PME pair functions driven from a cell list (oracle/cell_driver.cpp): same arithmetic as `Oracle`,
    O(N) pair enumeration, optionally several threads.  PME with explicit parameters only."""

    def __init__(self, s, threads=1):
        L = lib()
        L.mpidcell_last_error.restype = ctypes.c_char_p
        self.lib = L
        self.s = s
        self.threads = int(threads)
        if s.method != 1 or s.alpha == 0.0 or int(s.grid[0]) == 0:
            raise ValueError("CellOracle: PME with explicit alpha/grid only")
        off, idx = s.cov_csr()
        off = np.ascontiguousarray(off, dtype=np.int32); idx = np.ascontiguousarray(idx, dtype=np.int32)
        box = np.ascontiguousarray(s.box, dtype=np.float64).reshape(-1)
        coefs = np.ascontiguousarray(s.coefs, dtype=np.float64)
        c = np.ascontiguousarray
        h = ctypes.c_void_p()
        rc = L.mpidcell_create(ctypes.c_int(s.n), _dp(c(s.charges)), _dp(c(s.dipoles)), _dp(c(s.quadrupoles)), _dp(c(s.octopoles)),
                               _ip(c(s.axis)), _ip(c(s.atomZ)), _ip(c(s.atomX)), _ip(c(s.atomY)), _dp(c(s.tholes)), _dp(c(s.alphas)),
                               _ip(off), _ip(idx), ctypes.c_int(s.polarization), ctypes.c_double(s.cutoff),
                               ctypes.c_double(s.alpha), ctypes.c_int(int(s.grid[0])), ctypes.c_int(int(s.grid[1])), ctypes.c_int(int(s.grid[2])),
                               ctypes.c_double(s.default_thole), ctypes.c_double(s.scale14),
                               ctypes.c_int(s.max_iter), ctypes.c_double(s.epsilon), ctypes.c_int(len(coefs)), _dp(coefs), _dp(box),
                               ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(L.mpidcell_last_error().decode())
        self.h = h

    def execute(self, pos=None, threads=None):
        pos = np.ascontiguousarray(self.s.pos if pos is None else pos, dtype=np.float64)
        e = ctypes.c_double()
        f = np.zeros((self.s.n, 3))
        t = self.threads if threads is None else int(threads)
        if self.lib.mpidcell_execute(self.h, _dp(pos), ctypes.c_int(t), ctypes.byref(e), _dp(f)) != 0:
            raise RuntimeError(self.lib.mpidcell_last_error().decode())
        return e.value, f

    def induced(self):
        out = np.zeros((self.s.n, 3))
        if self.lib.mpidcell_get_induced(self.h, _dp(out)) != 0:
            raise RuntimeError(self.lib.mpidcell_last_error().decode())
        return out

    def profile(self):
        sec = np.zeros(5)
        st = np.zeros(4, dtype=np.int64)
        self.lib.mpidcell_get_profile(self.h, _dp(sec), st.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
        return dict(candidates_s=sec[0], fixed_field_s=sec[1], induced_fields_s=sec[2], electrostatics_s=sec[3], total_s=sec[4],
                    candidate_pairs=int(st[0]), induced_field_evaluations=int(st[1]), iterations=int(st[2]), threads=int(st[3]))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpidcell_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
