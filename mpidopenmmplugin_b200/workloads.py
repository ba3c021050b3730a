"""Synthetic MPID workloads named by BASELINE.json: SWM6-MPID water boxes built from the reference's
996-water example box (coordinates fixture: tests/golden/waterbox_31ang.npz, parameters:
reference examples/parameters/swm6.xml:25-39) and tiled to the 96k / 1M atom sizes of BASELINE.md."""
import copy
import os

import numpy as np

from .api import MPIDForce

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FlatSystem:
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
        self.covalent = None          # list[n][8] of lists, or None with cov_offsets/cov_indices set
        self.cov_offsets = None
        self.cov_indices = None
        self.box = np.diag([2.0, 2.0, 2.0])
        self.method = 0               # 0 NoCutoff, 1 PME
        self.polarization = 0         # 0 Mutual, 1 Direct, 2 Extrapolated
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
        if self.cov_offsets is not None:
            return self.cov_offsets, self.cov_indices
        offsets = np.zeros(8*(n+1), dtype=np.int32)
        idx = []
        cov = self.covalent if self.covalent is not None else [[[] for _ in range(8)] for _ in range(n)]
        for t in range(8):
            for i in range(n):
                offsets[t*(n+1)+i] = len(idx)
                idx.extend(cov[i][t])
            offsets[t*(n+1)+n] = len(idx)
        return offsets, np.array(idx if idx else [0], dtype=np.int32)

    def copy(self):
        return copy.deepcopy(self)

    def to_force(self):
        """The same system as an MPIDForce object (small systems; O(N) python calls)."""
        f = MPIDForce()
        f.setNonbondedMethod(self.method)
        f.setPolarizationType(self.polarization)
        f.setCutoffDistance(self.cutoff)
        f.setPMEParameters(self.alpha, *self.grid)
        f.setEwaldErrorTolerance(self.ewald_tol)
        f.setDefaultTholeWidth(self.default_thole)
        f.set14ScaleFactor(self.scale14)
        f.setMutualInducedMaxIterations(self.max_iter)
        f.setMutualInducedTargetEpsilon(self.epsilon)
        f.setExtrapolationCoefficients(list(self.coefs))
        off, idx = self.cov_csr()
        n = self.n
        for i in range(n):
            f.addMultipole(self.charges[i], self.dipoles[i], self.quadrupoles[i], self.octopoles[i], int(self.axis[i]),
                           int(self.atomZ[i]), int(self.atomX[i]), int(self.atomY[i]), self.tholes[i], self.alphas[i])
            for t in range(8):
                lst = idx[off[t*(n+1)+i]:off[t*(n+1)+i+1]]
                if len(lst):
                    f.setCovalentMap(i, t, [int(v) for v in lst])
        return f


def make_kernel(system, precision="mixed", device=0, profiling=False, frameless_alpha_fix=False, solver="diis"):
    """Create an engine for a FlatSystem through the raw C ABI (no per-atom python loops)."""
    import ctypes
    from .api import MPIDB200Kernel, _Config, _dp, _ip
    k = MPIDB200Kernel(precision=precision, device=device, frameless_alpha_fix=frameless_alpha_fix, solver=solver)
    lib = k._lib
    cfg = _Config()
    lib.mpidb200_default_config(ctypes.byref(cfg))
    cfg.num_particles = system.n
    cfg.nonbonded_method = system.method
    cfg.polarization_type = system.polarization
    cfg.cutoff = system.cutoff
    cfg.ewald_alpha = system.alpha
    cfg.grid[0], cfg.grid[1], cfg.grid[2] = [int(g) for g in system.grid]
    cfg.ewald_tolerance = system.ewald_tol
    cfg.default_thole_width = system.default_thole
    cfg.scale14 = system.scale14
    cfg.max_iterations = system.max_iter
    cfg.target_epsilon = system.epsilon
    cfg.num_extrapolation_coefficients = len(system.coefs)
    for i, c in enumerate(system.coefs):
        cfg.extrapolation_coefficients[i] = float(c)
    cfg.precision = k._precision
    cfg.solver = k._solver
    cfg.device = device
    cfg.frameless_alpha_fix = k._fix
    h = ctypes.c_void_p()
    k._check(lib.mpidb200_create(ctypes.byref(cfg), ctypes.byref(h)))
    k._h = h
    k._n = system.n
    k._uses_pme = system.method == 1
    off, idx = system.cov_csr()
    c = np.ascontiguousarray
    k._check(lib.mpidb200_set_particles(h, _dp(c(system.charges)), _dp(c(system.dipoles)), _dp(c(system.quadrupoles)),
                                        _dp(c(system.octopoles)), _ip(c(system.axis)), _ip(c(system.atomZ)), _ip(c(system.atomX)),
                                        _ip(c(system.atomY)), _dp(c(system.tholes)), _dp(c(system.alphas))))
    k._check(lib.mpidb200_set_covalent_maps(h, _ip(c(off)), _ip(c(idx))))
    if k._uses_pme:
        k.setPeriodicBoxVectors(system.box)
    if profiling:
        k.setProfiling(True)
    return k


# SWM6-MPID water (reference examples/parameters/swm6.xml:25-39); API component orders.
SWM6_O = dict(charge=-1.0614, dipole=[0.0, 0.0, -0.023671684],
              quadrupole=[0.000150963, 0.0, 0.00008707, 0.0, 0.0, -0.000238034],
              octopole=[0.0, 0.0, 0.0, 0.0, 0.000000426, 0.0, 0.000000853, 0.0, 0.0, -0.000001279],
              alpha=[0.00088, 0.00088, 0.00088], thole=8.0)
SWM6_H = dict(charge=0.5307, dipole=[0.0]*3, quadrupole=[0.0]*6, octopole=[0.0]*10, alpha=[0.0]*3, thole=0.0)
# anisotropic variant of the O site used to exercise the induced-dipole torque branches
ANISO_ALPHA_O = [0.00100024*0.88, 0.00125025*0.88, 0.00083350*0.88]


def _water_topology(s, anisotropic=False):
    """O H1 H2 per molecule; frames and covalent maps as the reference's generator assigns them
    (python/mpidplugin.i:845-1044; test fixture TestReferenceMPIDForce.cpp:125-146, 626-632)."""
    n = s.n
    nw = n//3
    o = np.arange(nw)*3
    for k, v in (("charges", "charge"), ("tholes", "thole")):
        arr = getattr(s, k)
        arr[o] = SWM6_O[v]; arr[o+1] = SWM6_H[v]; arr[o+2] = SWM6_H[v]
    s.dipoles[o] = SWM6_O["dipole"]
    s.quadrupoles[o] = SWM6_O["quadrupole"]
    s.octopoles[o] = SWM6_O["octopole"]
    s.alphas[o] = ANISO_ALPHA_O if anisotropic else SWM6_O["alpha"]
    s.axis[o] = MPIDForce.Bisector; s.atomZ[o] = o+1; s.atomX[o] = o+2
    s.axis[o+1] = MPIDForce.ZThenX; s.atomZ[o+1] = o; s.atomX[o+1] = o+2
    s.axis[o+2] = MPIDForce.ZThenX; s.atomZ[o+2] = o; s.atomX[o+2] = o+1
    # covalent CSR: type 0 (1-2): O:[H1,H2], H:[O];  type 1 (1-3): H1:[H2], H2:[H1]
    offsets = np.zeros(8*(n+1), dtype=np.int64)
    c12 = np.zeros(n+1, dtype=np.int64)
    cnt12 = np.ones(n, dtype=np.int64); cnt12[o] = 2
    c12[1:] = np.cumsum(cnt12)
    idx12 = np.zeros(c12[-1], dtype=np.int32)
    idx12[c12[o]] = o+1; idx12[c12[o]+1] = o+2; idx12[c12[o+1]] = o; idx12[c12[o+2]] = o
    cnt13 = np.ones(n, dtype=np.int64); cnt13[o] = 0
    c13 = np.zeros(n+1, dtype=np.int64); c13[1:] = np.cumsum(cnt13)
    idx13 = np.zeros(c13[-1], dtype=np.int32)
    idx13[c13[o+1]] = o+2; idx13[c13[o+2]] = o+1
    offsets[0:n+1] = c12
    offsets[n+1:2*(n+1)] = len(idx12) + c13
    offsets[2*(n+1):] = len(idx12) + len(idx13)
    s.cov_offsets = offsets.astype(np.int32)
    s.cov_indices = np.concatenate([idx12, idx13]).astype(np.int32)
    s.covalent = None


def water_box(tiles=(1, 1, 1), jitter=0.005, seed=20261017, polarization=0, epsilon=1e-5, anisotropic=False,
              grid=None, cutoff=0.8, alpha=3.2853, default_thole=8.0):
    """SWM6-MPID water: the 996-water box of examples/waterbox tiled `tiles` times.

    (1,1,1) -> N = 2,988, L = 3.1289 nm, grid 32^3        (config 3)
    (4,4,2) -> N = 95,616, 12.5156 x 12.5156 x 6.2578 nm, grid 128 x 128 x 64   (config 4)
    (7,7,7) -> N = 1,024,884, L = 21.9023 nm, grid 224^3  (config 5)
    Each water of the tiled boxes is rigidly translated by N(0, jitter nm) (seeded) to break the exact
    replication; the (1,1,1) box is left as in the PDB."""
    d = np.load(os.path.join(_ROOT, "tests", "golden", "waterbox_31ang.npz"))
    base = d["milli_angstrom"].astype(np.float64)*1e-4        # nm
    L = d["box_angstrom"].astype(np.float64)*0.1
    tx, ty, tz = tiles
    reps = []
    for ix in range(tx):
        for iy in range(ty):
            for iz in range(tz):
                reps.append(base + np.array([ix*L[0], iy*L[1], iz*L[2]]))
    pos = np.concatenate(reps, axis=0)
    n = len(pos)
    if (tx, ty, tz) != (1, 1, 1) and jitter > 0:
        rng = np.random.default_rng(seed)
        shift = rng.normal(0.0, jitter, size=(n//3, 3))
        pos += np.repeat(shift, 3, axis=0)
    s = FlatSystem(n)
    s.pos = pos
    _water_topology(s, anisotropic)
    s.box = np.diag([tx*L[0], ty*L[1], tz*L[2]])
    s.method = 1
    s.polarization = polarization
    s.cutoff = cutoff
    s.alpha = alpha
    s.grid = tuple(grid) if grid is not None else (32*tx, 32*ty, 32*tz)
    s.default_thole = default_thole
    s.epsilon = epsilon
    s.max_iter = 100
    return s


def subset_waters(s, nwaters):
    """First `nwaters` molecules of a water FlatSystem in the same box (used for bounded CPU samples)."""
    n = 3*nwaters
    t = FlatSystem(n)
    for k in ("pos", "charges", "dipoles", "quadrupoles", "octopoles", "axis", "atomZ", "atomX", "atomY", "tholes", "alphas"):
        setattr(t, k, getattr(s, k)[:n].copy())
    _water_topology(t, anisotropic=bool(np.any(s.alphas[0] != s.alphas[0][0])))
    for k in ("box", "method", "polarization", "cutoff", "alpha", "grid", "ewald_tol", "default_thole", "scale14", "max_iter", "epsilon", "coefs"):
        setattr(t, k, copy.deepcopy(getattr(s, k)))
    return t


def ethane_box(polarization=1, cutoff=0.8, default_thole=8.0):
    """examples/ethane_water_charge_only (BASELINE.json config 2): ethane in 1383 waters, N = 4,157, L = 3.5 nm,
    charges only, isotropic polarizabilities on OW and CT3, Direct polarization (run_ethane.py:12-13,
    ethane_water.xml:51-59).  Every <Multipole> entry of that force field lacks kz/kx, so the reference's generator
    gives every atom NoAxisType with anchors -1 (python/mpidplugin.i:1009-1023) -- and on the Reference platform
    such atoms end up with a zero lab-frame polarizability.  Covalent12/13/14 maps from the bonds
    (mpidplugin.i:1042-1044).  Coordinates: tests/golden/ethane_water.npz (made by tests/golden/make_ethane_fixture.py)."""
    d = np.load(os.path.join(_ROOT, "tests", "golden", "ethane_water.npz"))
    pos = d["milli_angstrom"].astype(np.float64)*1e-4
    z = d["atomic_number"]
    n = len(pos)
    s = FlatSystem(n)
    s.pos = pos
    bonds = [tuple(b) for b in d["ethane_bonds"]]
    for w in range(8, n, 3):
        bonds += [(w, w+1), (w, w+2)]
    nb = [set() for _ in range(n)]
    for a, b in bonds:
        nb[a].add(int(b)); nb[b].add(int(a))
    carbon = {0, 1}
    for i in range(n):
        if i < 8:
            s.charges[i] = -0.27 if i in carbon else 0.09
            if i in carbon:
                s.alphas[i] = 0.00068; s.tholes[i] = 8.0
        elif z[i] == 8:
            s.charges[i] = -0.834; s.alphas[i] = 0.00088; s.tholes[i] = 8.0
        else:
            s.charges[i] = 0.417
    cov = [[[] for _ in range(8)] for _ in range(n)]
    for i in range(n):
        c12 = nb[i]
        c13 = set()
        for j in c12:
            c13 |= nb[j]
        c13 -= c12 | {i}
        c14 = set()
        for j in c13:
            c14 |= nb[j]
        c14 -= c12 | c13 | {i}
        cov[i][0] = sorted(c12); cov[i][1] = sorted(c13); cov[i][2] = sorted(c14)
    s.covalent = cov
    L = d["box_angstrom"].astype(np.float64)*0.1
    s.box = np.diag(L)
    s.method = 1
    s.polarization = polarization
    s.cutoff = cutoff
    s.alpha = 0.0          # automatic PME parameters, as the example script leaves them (ewaldErrorTolerance 5e-4)
    s.grid = (0, 0, 0)
    s.default_thole = default_thole
    s.scale14 = 1.0
    return s
