"""Python host of the MPIDB200 engine (ctypes over the C ABI in include/mpidb200.h).

`MPIDForce` restates the plugin's parameter object (reference: openmmapi/include/openmm/MPIDForce.h:54-489,
openmmapi/src/MPIDForce.cpp) with the same method names, argument meaning, defaults and error behaviour;
`MPIDB200Kernel` restates the kernel contract `CalcMPIDForceKernel`
(reference: openmmapi/include/openmm/mpidKernels.h:50-104) on top of the engine.  The same mapping in C++
(for a real OpenMM build) lives in mpidopenmmplugin_b200/plugin/."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MPIDB200Error(RuntimeError):
    """Counterpart of OpenMM::OpenMMException for this host layer."""


def library_path():
    return os.path.join(_HERE, "libmpidb200.so")


class _Config(ctypes.Structure):
    _fields_ = [("num_particles", ctypes.c_int), ("nonbonded_method", ctypes.c_int), ("polarization_type", ctypes.c_int),
                ("cutoff", ctypes.c_double), ("ewald_alpha", ctypes.c_double), ("grid", ctypes.c_int*3),
                ("ewald_tolerance", ctypes.c_double), ("default_thole_width", ctypes.c_double), ("scale14", ctypes.c_double),
                ("max_iterations", ctypes.c_int), ("target_epsilon", ctypes.c_double),
                ("num_extrapolation_coefficients", ctypes.c_int), ("extrapolation_coefficients", ctypes.c_double*8),
                ("precision", ctypes.c_int), ("solver", ctypes.c_int), ("device", ctypes.c_int),
                ("frameless_alpha_fix", ctypes.c_int)]


STAGE_NAMES = ["sort", "nlist", "fixed_spread", "fft", "fixed_gather", "fixed_real", "ind_spread", "ind_gather", "ind_real",
               "solver", "electrostatics", "finish"]

EXPORTS = ["mpidb200_last_error", "mpidb200_default_config", "mpidb200_create", "mpidb200_destroy", "mpidb200_set_particles",
           "mpidb200_set_covalent_maps", "mpidb200_set_box", "mpidb200_execute", "mpidb200_execute_device",
           "mpidb200_get_dipoles", "mpidb200_get_system_multipole_moments", "mpidb200_get_electrostatic_potential",
           "mpidb200_get_pme_parameters", "mpidb200_get_stats", "mpidb200_set_profiling", "mpidb200_last_launch_count",
           "mpidb200_get_pair_list", "mpidb200_get_pair_class_counts", "mpidb200_nccl_unique_id", "mpidb200_comm_init", "mpidb200_set_stream"]


def load_library():
    """dlopen the engine.  Raises if it has not been built -- there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise MPIDB200Error("%s not found: build it with `python -m mpidopenmmplugin_b200.build` "
                                "(there is no CPU fallback)" % path)
        lib = ctypes.CDLL(path)
        lib.mpidb200_last_error.restype = ctypes.c_char_p
        lib.mpidb200_last_launch_count.restype = ctypes.c_longlong
        lib.mpidb200_last_launch_count.argtypes = [ctypes.c_void_p]
        lib.mpidb200_destroy.argtypes = [ctypes.c_void_p]
        lib.mpidb200_destroy.restype = None
        _LIB = lib
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class MPIDForce:
    """Parameter container with the reference's API (MPIDForce.h:54-489)."""
    # NonbondedMethod (MPIDForce.h:58-71)
    NoCutoff, PME = 0, 1
    # PolarizationType (MPIDForce.h:73-95)
    Mutual, Direct, Extrapolated = 0, 1, 2
    # MultipoleAxisTypes (MPIDForce.h:97)
    ZThenX, Bisector, ZBisect, ThreeFold, ZOnly, NoAxisType, LastAxisTypeIndex = 0, 1, 2, 3, 4, 5, 6
    # CovalentType (MPIDForce.h:99-101)
    (Covalent12, Covalent13, Covalent14, Covalent15, PolarizationCovalent11, PolarizationCovalent12,
     PolarizationCovalent13, PolarizationCovalent14, CovalentEnd) = range(9)

    def __init__(self):
        # defaults: openmmapi/src/MPIDForce.cpp:43-50
        self._method = MPIDForce.NoCutoff
        self._polarization = MPIDForce.Extrapolated
        self._cutoff = 1.0
        self._alpha = 0.0
        self._grid = [0, 0, 0]
        self._ewald_tol = 5e-4
        self._max_iter = 60
        self._epsilon = 1e-5
        self._default_thole = 5.0
        self._scale14 = 1.0
        self._coefs = [-0.154, 0.017, 0.658, 0.474]
        self._multipoles = []
        self._force_group = 0          # OpenMM::Force::getForceGroup / setForceGroup

    # ---- global settings ----------------------------------------------------------------------------
    def getForceGroup(self):
        return self._force_group

    def setForceGroup(self, group):
        if group < 0 or group > 31:
            raise MPIDB200Error("Force group must be between 0 and 31")
        self._force_group = int(group)

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        if method not in (MPIDForce.NoCutoff, MPIDForce.PME):
            raise MPIDB200Error("MPIDForce: Illegal value for nonbonded method")
        self._method = method

    def getPolarizationType(self):
        return self._polarization

    def setPolarizationType(self, t):
        self._polarization = t

    def getCutoffDistance(self):
        return self._cutoff

    def setCutoffDistance(self, d):
        self._cutoff = float(d)

    def getPMEParameters(self):
        return self._alpha, self._grid[0], self._grid[1], self._grid[2]

    def setPMEParameters(self, alpha, nx, ny, nz):
        self._alpha = float(alpha)
        self._grid = [int(nx), int(ny), int(nz)]

    def getAEwald(self):
        return self._alpha

    def setAEwald(self, a):
        self._alpha = float(a)

    def getPmeBSplineOrder(self):
        return 6

    def getPmeGridDimensions(self):
        return list(self._grid)

    def setPmeGridDimensions(self, g):
        self._grid = [int(g[0]), int(g[1]), int(g[2])]

    def getEwaldErrorTolerance(self):
        return self._ewald_tol

    def setEwaldErrorTolerance(self, t):
        self._ewald_tol = float(t)

    def getMutualInducedMaxIterations(self):
        return self._max_iter

    def setMutualInducedMaxIterations(self, n):
        self._max_iter = int(n)

    def getMutualInducedTargetEpsilon(self):
        return self._epsilon

    def setMutualInducedTargetEpsilon(self, e):
        self._epsilon = float(e)

    def getExtrapolationCoefficients(self):
        return list(self._coefs)

    def setExtrapolationCoefficients(self, c):
        self._coefs = [float(x) for x in c]

    def getDefaultTholeWidth(self):
        return self._default_thole

    def setDefaultTholeWidth(self, w):
        self._default_thole = float(w)

    def get14ScaleFactor(self):
        return self._scale14

    def set14ScaleFactor(self, s):
        self._scale14 = float(s)

    def usesPeriodicBoundaryConditions(self):
        return self._method == MPIDForce.PME

    # ---- particles ----------------------------------------------------------------------------------
    def getNumMultipoles(self):
        return len(self._multipoles)

    def addMultipole(self, charge, molecularDipole, molecularQuadrupole, molecularOctopole, axisType,
                     multipoleAtomZ, multipoleAtomX, multipoleAtomY, thole, alphas):
        self._multipoles.append(self._pack(charge, molecularDipole, molecularQuadrupole, molecularOctopole, axisType,
                                           multipoleAtomZ, multipoleAtomX, multipoleAtomY, thole, alphas))
        return len(self._multipoles) - 1

    @staticmethod
    def _pack(charge, d, q, o, axisType, z, x, y, thole, alphas):
        if len(d) != 3 or len(q) != 6 or len(o) != 10 or len(alphas) != 3:
            raise MPIDB200Error("MPIDForce: dipole/quadrupole/octopole/alphas must have 3/6/10/3 entries")
        return dict(charge=float(charge), dipole=[float(v) for v in d], quadrupole=[float(v) for v in q],
                    octopole=[float(v) for v in o], axisType=int(axisType), z=int(z), x=int(x), y=int(y),
                    thole=float(thole), alphas=[float(v) for v in alphas], covalent=[[] for _ in range(8)])

    def getMultipoleParameters(self, index):
        m = self._multipoles[index]
        return (m["charge"], list(m["dipole"]), list(m["quadrupole"]), list(m["octopole"]), m["axisType"],
                m["z"], m["x"], m["y"], m["thole"], list(m["alphas"]))

    def setMultipoleParameters(self, index, charge, molecularDipole, molecularQuadrupole, molecularOctopole, axisType,
                               multipoleAtomZ, multipoleAtomX, multipoleAtomY, thole, alphas):
        cov = self._multipoles[index]["covalent"]
        self._multipoles[index] = self._pack(charge, molecularDipole, molecularQuadrupole, molecularOctopole, axisType,
                                             multipoleAtomZ, multipoleAtomX, multipoleAtomY, thole, alphas)
        self._multipoles[index]["covalent"] = cov

    def setCovalentMap(self, index, typeId, covalentAtoms):
        self._multipoles[index]["covalent"][typeId] = [int(a) for a in covalentAtoms]

    def getCovalentMap(self, index, typeId):
        return list(self._multipoles[index]["covalent"][typeId])

    def getCovalentMaps(self, index):
        return [list(c) for c in self._multipoles[index]["covalent"]]

    # ---- context-style queries (forwarded to the kernel, MPIDForceImpl.cpp:212-240) -------------------
    def getInducedDipoles(self, kernel, positions):
        return kernel.getInducedDipoles(positions)

    def getLabFramePermanentDipoles(self, kernel, positions):
        return kernel.getLabFramePermanentDipoles(positions)

    def getTotalDipoles(self, kernel, positions):
        return kernel.getTotalDipoles(positions)

    def getPMEParametersInContext(self, kernel):
        return kernel.getPMEParameters()

    def getElectrostaticPotential(self, inputGrid, kernel, positions):
        """MPIDForce::getElectrostaticPotential(inputGrid, context, out) (MPIDForce.h): potential at the grid points, kJ/mol/e."""
        return kernel.getElectrostaticPotential(positions, inputGrid)

    def getSystemMultipoleMoments(self, kernel, positions, masses):
        """MPIDForce::getSystemMultipoleMoments(context, out): charge, dipole (Debye), quadrupole (Debye.Angstrom), 13 values."""
        return kernel.getSystemMultipoleMoments(positions, masses)

    def updateParametersInContext(self, kernel):
        kernel.copyParametersToContext(self)

    # ---- validation done by MPIDForceImpl::initialize (openmmapi/src/MPIDForceImpl.cpp:51-149) ---------
    def validate(self, numParticles, boxVectors):
        if numParticles != self.getNumMultipoles():
            raise MPIDB200Error("MPIDForce must have exactly as many particles as the System it belongs to.")
        if self._method == MPIDForce.PME:
            cutoff = self._cutoff
            if cutoff > 0.5*boxVectors[0][0] or cutoff > 0.5*boxVectors[1][1] or cutoff > 0.5*boxVectors[2][2]:
                raise MPIDB200Error("MPIDForce: The cutoff distance cannot be greater than half the periodic box size.")
        for i, m in enumerate(self._multipoles):
            q, o = m["quadrupole"], m["octopole"]
            if abs(q[0] + q[2] + q[5]) > 1e-5:
                raise MPIDB200Error("MPIDForce: The multipole quadrupole trace for particle %d is not zero" % i)
            # octopole API order: XXX XXY XYY YYY XXZ XYZ YYZ XZZ YZZ ZZZ
            if abs(o[0] + o[2] + o[7]) > 1e-5 or abs(o[1] + o[3] + o[8]) > 1e-5 or abs(o[4] + o[6] + o[9]) > 1e-5:
                raise MPIDB200Error("MPIDForce: The multipole octopole trace for particle %d is not zero" % i)
            axis = m["axisType"]
            if axis < 0 or axis >= MPIDForce.LastAxisTypeIndex:
                raise MPIDB200Error("MPIDForce: axis type=%d not currently handled" % axis)
            if axis != MPIDForce.NoAxisType and (m["z"] < 0 or m["z"] >= numParticles):
                raise MPIDB200Error("MPIDForce: invalid z axis particle: %d for particle %d" % (m["z"], i))
            if axis not in (MPIDForce.NoAxisType, MPIDForce.ZOnly) and (m["x"] < 0 or m["x"] >= numParticles):
                raise MPIDB200Error("MPIDForce: invalid x axis particle: %d for particle %d" % (m["x"], i))
            if axis in (MPIDForce.ZBisect, MPIDForce.ThreeFold) and (m["y"] < 0 or m["y"] >= numParticles):
                raise MPIDB200Error("MPIDForce: invalid y axis particle: %d for particle %d" % (m["y"], i))


class MPIDB200Kernel:
    """`CalcMPIDForceKernel` on the B200 engine (mpidKernels.h:50-104).

    initialize(numParticles, force, boxVectors)  <-> initialize(const System&, const MPIDForce&)
    execute(positions, includeForces, includeEnergy, forces=None) -> energy; forces are ACCUMULATED
    """

    @staticmethod
    def Name():
        return "CalcMPIDForce"

    def __init__(self, precision="mixed", device=0, solver="diis", frameless_alpha_fix=False):
        self._lib = load_library()
        self._h = None
        self._precision = {"mixed": 0, "single": 0, "double": 1}[precision]
        self._solver = {"diis": 0, "cg": 1}[solver]
        self._device = device
        self._fix = 1 if frameless_alpha_fix else 0
        self._n = 0

    def _check(self, rc):
        if rc != 0:
            raise MPIDB200Error(self._lib.mpidb200_last_error().decode())

    def initialize(self, numParticles, force, boxVectors=None):
        if boxVectors is None:
            boxVectors = np.diag([2.0, 2.0, 2.0])
        boxVectors = np.asarray(boxVectors, dtype=np.float64).reshape(3, 3)
        force.validate(numParticles, boxVectors)
        if self._h is not None:
            self.close()
        cfg = _Config()
        self._lib.mpidb200_default_config(ctypes.byref(cfg))
        cfg.num_particles = numParticles
        cfg.nonbonded_method = force.getNonbondedMethod()
        cfg.polarization_type = force.getPolarizationType()
        cfg.cutoff = force.getCutoffDistance()
        a, nx, ny, nz = force.getPMEParameters()
        cfg.ewald_alpha = a
        cfg.grid[0], cfg.grid[1], cfg.grid[2] = nx, ny, nz
        cfg.ewald_tolerance = force.getEwaldErrorTolerance()
        cfg.default_thole_width = force.getDefaultTholeWidth()
        cfg.scale14 = force.get14ScaleFactor()
        cfg.max_iterations = force.getMutualInducedMaxIterations()
        cfg.target_epsilon = force.getMutualInducedTargetEpsilon()
        coefs = force.getExtrapolationCoefficients()
        if len(coefs) > 8:
            raise MPIDB200Error("MPIDForce: at most 8 extrapolation coefficients are supported")
        cfg.num_extrapolation_coefficients = len(coefs)
        for i, c in enumerate(coefs):
            cfg.extrapolation_coefficients[i] = c
        cfg.precision = self._precision
        cfg.solver = self._solver
        cfg.device = self._device
        cfg.frameless_alpha_fix = self._fix
        h = ctypes.c_void_p()
        self._check(self._lib.mpidb200_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self._n = numParticles
        self._uses_pme = force.getNonbondedMethod() == MPIDForce.PME
        self.copyParametersToContext(force)
        if self._uses_pme:
            self.setPeriodicBoxVectors(boxVectors)

    def copyParametersToContext(self, force):
        n = self._n
        if force.getNumMultipoles() != n:
            raise MPIDB200Error("updateParametersInContext: The number of multipoles has changed")
        charges = np.zeros(n); dip = np.zeros((n, 3)); quad = np.zeros((n, 6)); octo = np.zeros((n, 10))
        axis = np.zeros(n, dtype=np.int32); z = np.zeros(n, dtype=np.int32); x = np.zeros(n, dtype=np.int32)
        y = np.zeros(n, dtype=np.int32); thole = np.zeros(n); alphas = np.zeros((n, 3))
        offsets = np.zeros(8*(n+1), dtype=np.int32)
        for i in range(n):
            c, d, q, o, at, az, ax, ay, th, al = force.getMultipoleParameters(i)
            charges[i] = c; dip[i] = d; quad[i] = q; octo[i] = o
            axis[i] = at; z[i] = az; x[i] = ax; y[i] = ay; thole[i] = th; alphas[i] = al
        idx = []
        for t in range(8):
            for i in range(n):
                offsets[t*(n+1)+i] = len(idx)
                idx.extend(force.getCovalentMap(i, t))
            offsets[t*(n+1)+n] = len(idx)
        idx = np.array(idx if idx else [0], dtype=np.int32)
        self._check(self._lib.mpidb200_set_particles(self._h, _dp(charges), _dp(dip), _dp(quad), _dp(octo), _ip(axis),
                                                     _ip(z), _ip(x), _ip(y), _dp(thole), _dp(alphas)))
        self._check(self._lib.mpidb200_set_covalent_maps(self._h, _ip(offsets), _ip(idx)))

    def setPeriodicBoxVectors(self, boxVectors):
        b = np.ascontiguousarray(np.asarray(boxVectors, dtype=np.float64).reshape(3, 3))
        self._check(self._lib.mpidb200_set_box(self._h, _dp(b[0].copy()), _dp(b[1].copy()), _dp(b[2].copy())))

    def execute(self, positions, includeForces=True, includeEnergy=True, forces=None):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        if pos.size != 3*self._n:
            raise MPIDB200Error("execute: positions must hold 3*N values")
        e = ctypes.c_double(0.0)
        fptr = None
        if includeForces:
            if forces is None:
                raise MPIDB200Error("execute: a force array is required when includeForces is set")
            if forces.dtype != np.float64 or not forces.flags["C_CONTIGUOUS"] or forces.size != 3*self._n:
                raise MPIDB200Error("execute: forces must be a C-contiguous float64 array of 3*N values")
            fptr = _dp(forces)
        self._check(self._lib.mpidb200_execute(self._h, _dp(pos), ctypes.c_int(1 if includeForces else 0),
                                               ctypes.c_int(1 if includeEnergy else 0), ctypes.byref(e), fptr))
        return e.value

    def execute_device(self, d_positions, includeForces, includeEnergy, d_forces):
        """positions / forces are raw device pointers (ints) to double[3N] on the engine's device."""
        e = ctypes.c_double(0.0)
        self._check(self._lib.mpidb200_execute_device(self._h, ctypes.c_void_p(d_positions), ctypes.c_int(1 if includeForces else 0),
                                                      ctypes.c_int(1 if includeEnergy else 0), ctypes.byref(e),
                                                      ctypes.c_void_p(d_forces)))
        return e.value

    def execute_cuda_context(self, d_posq, posq_is_double, d_posq_correction, d_atom_index, padded_num_atoms, includeForces, includeEnergy, d_force_buffer):
        """mpidb200_execute_cuda_context: raw device pointers (ints) in the layouts of an OpenMM CudaContext."""
        e = ctypes.c_double(0.0)
        self._check(self._lib.mpidb200_execute_cuda_context(self._h, ctypes.c_void_p(d_posq), ctypes.c_int(1 if posq_is_double else 0),
                                                            ctypes.c_void_p(d_posq_correction if d_posq_correction else None), ctypes.c_void_p(d_atom_index),
                                                            ctypes.c_int(padded_num_atoms), ctypes.c_int(1 if includeForces else 0),
                                                            ctypes.c_int(1 if includeEnergy else 0), ctypes.byref(e), ctypes.c_void_p(d_force_buffer if d_force_buffer else None)))
        return e.value

    def _dipoles(self, positions, which):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        out = np.zeros((self._n, 3))
        self._check(self._lib.mpidb200_get_dipoles(self._h, _dp(pos), ctypes.c_int(which), _dp(out)))
        return out

    def getInducedDipoles(self, positions):
        return self._dipoles(positions, 0)

    def getLabFramePermanentDipoles(self, positions):
        return self._dipoles(positions, 1)

    def getTotalDipoles(self, positions):
        return self._dipoles(positions, 2)

    def getSystemMultipoleMoments(self, positions, masses):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        m = np.ascontiguousarray(masses, dtype=np.float64)
        out = np.zeros(13)
        self._check(self._lib.mpidb200_get_system_multipole_moments(self._h, _dp(pos), _dp(m), _dp(out)))
        return out

    def getElectrostaticPotential(self, positions, points):
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1)
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(pts))
        self._check(self._lib.mpidb200_get_electrostatic_potential(self._h, _dp(pos), ctypes.c_int(len(pts)), _dp(pts), _dp(out)))
        return out

    def getPMEParameters(self):
        a = ctypes.c_double(); nx = ctypes.c_int(); ny = ctypes.c_int(); nz = ctypes.c_int()
        self._check(self._lib.mpidb200_get_pme_parameters(self._h, ctypes.byref(a), ctypes.byref(nx), ctypes.byref(ny), ctypes.byref(nz)))
        return a.value, nx.value, ny.value, nz.value

    # ---- diagnostics -------------------------------------------------------------------------------
    def setProfiling(self, enabled):
        self._check(self._lib.mpidb200_set_profiling(self._h, ctypes.c_int(1 if enabled else 0)))

    def getStats(self):
        it = ctypes.c_int(); eps = ctypes.c_double(); ms = (ctypes.c_double*16)(); pairs = ctypes.c_longlong()
        self._check(self._lib.mpidb200_get_stats(self._h, ctypes.byref(it), ctypes.byref(eps), ms, ctypes.byref(pairs)))
        names = STAGE_NAMES
        cls = (ctypes.c_longlong*3)()
        self._check(self._lib.mpidb200_get_pair_class_counts(self._h, cls))
        return dict(iterations=it.value, epsilon=eps.value, pairs=pairs.value,
                    pair_classes=dict(full_full=cls[0], full_charge=cls[1], charge_charge=cls[2]),
                    stage_ms={k: ms[i] for i, k in enumerate(names)},
                    launches=int(self._lib.mpidb200_last_launch_count(self._h)))

    def pinHostBuffer(self, array):
        """Page-lock a numpy array the caller keeps alive (positions / forces passed to execute): mpidb200_pin_host_buffer."""
        self._check(self._lib.mpidb200_pin_host_buffer(self._h, ctypes.c_void_p(array.ctypes.data), ctypes.c_ulonglong(array.nbytes)))

    def setHostIoPartition(self, enable=True):
        """Several ranks: rank r moves only its block of atoms between host and device (mpidb200_set_host_io_partition);
        positions of the other blocks arrive from the other ranks over NVLink, forces of the other blocks stay there."""
        self._check(self._lib.mpidb200_set_host_io_partition(self._h, ctypes.c_int(1 if enable else 0)))

    def getHostIoBlock(self):
        """(first_atom, num_atoms) this rank's execute() reads positions of / accumulates forces into."""
        first, count = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self._lib.mpidb200_get_host_io_block(self._h, ctypes.byref(first), ctypes.byref(count)))
        return first.value, count.value

    def unpinHostBuffer(self, array):
        self._check(self._lib.mpidb200_unpin_host_buffer(self._h, ctypes.c_void_p(array.ctypes.data)))

    def getWorkCounts(self):
        out = (ctypes.c_longlong*8)()
        self._check(self._lib.mpidb200_get_work_counts(self._h, out))
        names = ("pairs", "full_full", "full_charge", "charge_charge", "pol_pol", "fixed_field_directed", "covalent_pairs", "polarizable_sites")
        return {k: int(out[i]) for i, k in enumerate(names)}

    def debugReciprocalPass(self, grid, use_library=False):
        """One reciprocal pass of a float32 [nx][ny][nz] array (copy returned): hand-written kernels or cuFFT."""
        g = np.ascontiguousarray(grid, dtype=np.float32).copy()
        self._check(self._lib.mpidb200_debug_reciprocal_pass(self._h, g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_int(1 if use_library else 0)))
        return g

    def getListStats(self):
        out = (ctypes.c_longlong*2)()
        self._check(self._lib.mpidb200_get_list_stats(self._h, out))
        return dict(builds=int(out[0]), reuses=int(out[1]))

    def setKernelProfiling(self, enabled):
        self._check(self._lib.mpidb200_set_kernel_profiling(self._h, ctypes.c_int(1 if enabled else 0)))

    def getKernelProfile(self):
        """{kernel name: (launches per evaluation, microseconds per evaluation)} accumulated since setKernelProfiling(True)."""
        need = ctypes.c_longlong()
        self._check(self._lib.mpidb200_get_kernel_profile(self._h, None, ctypes.c_longlong(0), ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value + 1)
        self._check(self._lib.mpidb200_get_kernel_profile(self._h, buf, ctypes.c_longlong(need.value + 1), ctypes.byref(need)))
        out = {}
        for line in buf.value.decode().splitlines()[1:]:
            name, rest = line.rsplit('",', 1)
            launches, us, evals = rest.split(",")
            ev = max(int(evals), 1)
            out[name.strip('"')] = (int(launches)/ev, float(us)/ev)
        return out

    @staticmethod
    def measureFp32Peak(device=0):
        lib = load_library()
        t = ctypes.c_double(); sec = ctypes.c_double()
        if lib.mpidb200_measure_fp32_peak(ctypes.c_int(device), ctypes.byref(t), ctypes.byref(sec)) != 0:
            raise MPIDB200Error(lib.mpidb200_last_error().decode())
        return t.value

    def getPairList(self):
        cnt = ctypes.c_longlong()
        self._check(self._lib.mpidb200_get_pair_list(self._h, ctypes.c_longlong(0), None, None, None, ctypes.byref(cnt)))
        m = max(cnt.value, 1)
        pi = np.zeros(m, dtype=np.int32); pj = np.zeros(m, dtype=np.int32); pc = np.zeros(m, dtype=np.int32)
        self._check(self._lib.mpidb200_get_pair_list(self._h, ctypes.c_longlong(m), _ip(pi), _ip(pj), _ip(pc), ctypes.byref(cnt)))
        k = cnt.value
        return pi[:k], pj[:k], pc[:k]

    def setStream(self, cuda_stream):
        """cuda_stream: raw cudaStream_t as int (e.g. torch.cuda.current_stream().cuda_stream), or None."""
        self._check(self._lib.mpidb200_set_stream(self._h, ctypes.c_void_p(cuda_stream if cuda_stream else None)))

    def commInit(self, rank, numRanks, uniqueId):
        buf = (ctypes.c_ubyte*128).from_buffer_copy(bytes(uniqueId))
        self._check(self._lib.mpidb200_comm_init(self._h, ctypes.c_int(rank), ctypes.c_int(numRanks), buf))

    @staticmethod
    def ncclUniqueId():
        lib = load_library()
        buf = (ctypes.c_ubyte*128)()
        if lib.mpidb200_nccl_unique_id(buf) != 0:
            raise MPIDB200Error(lib.mpidb200_last_error().decode())
        return bytes(buf)

    def close(self):
        if self._h is not None:
            self._lib.mpidb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
