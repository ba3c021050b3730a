"""Host-side logic that needs no GPU: C-ABI exports, the MPIDForce mirror, workloads, pair-set helper."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from _common import ROOT, load_fixture, water_box, pair_set_reference
from mpidopenmmplugin_b200 import MPIDForce, MPIDB200Error, MPIDB200Kernel


def test_library_exports_every_declared_symbol(engine_lib):
    hdr = open(os.path.join(ROOT, "include", "mpidb200.h")).read()
    names = set(re.findall(r"\b(mpidb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    for nm in sorted(names):
        assert hasattr(engine_lib, nm), nm
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "mpidopenmmplugin_b200", "libmpidb200.so")],
                         capture_output=True, text=True).stdout
    for nm in names:
        assert re.search(r"\bT %s\b" % nm, out), nm


def test_library_is_sm100a_only(engine_lib):
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "mpidopenmmplugin_b200", "libmpidb200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_device(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mpidopenmmplugin_b200.workloads import make_kernel
    with pytest.raises(MPIDB200Error, match="no CUDA device"):
        make_kernel(water_box((1, 1, 1)))


def test_default_config_matches_mpidforce_defaults(engine_lib):
    """openmmapi/src/MPIDForce.cpp:43-50"""
    from mpidopenmmplugin_b200.api import _Config
    cfg = _Config()
    engine_lib.mpidb200_default_config(ctypes.byref(cfg))
    f = MPIDForce()
    assert cfg.polarization_type == f.getPolarizationType() == MPIDForce.Extrapolated
    assert cfg.nonbonded_method == f.getNonbondedMethod() == MPIDForce.NoCutoff
    assert cfg.cutoff == f.getCutoffDistance() == 1.0
    assert cfg.ewald_tolerance == f.getEwaldErrorTolerance() == 5e-4
    assert cfg.max_iterations == f.getMutualInducedMaxIterations() == 60
    assert cfg.target_epsilon == f.getMutualInducedTargetEpsilon() == 1e-5
    assert cfg.default_thole_width == f.getDefaultTholeWidth() == 5.0
    assert cfg.scale14 == f.get14ScaleFactor() == 1.0
    assert list(cfg.extrapolation_coefficients)[:4] == f.getExtrapolationCoefficients() == [-0.154, 0.017, 0.658, 0.474]
    assert f.getPmeBSplineOrder() == 6


def _force_from(s):
    return s.to_force()


def test_mpidforce_round_trip_and_validation():
    s = load_fixture("water_dimer")
    f = _force_from(s)
    assert f.getNumMultipoles() == 6
    c, d, q, o, ax, z, x, y, th, al = f.getMultipoleParameters(0)
    assert c == s.charges[0] and q == list(s.quadrupoles[0]) and ax == MPIDForce.Bisector and (z, x, y) == (1, 2, -1)
    assert f.getCovalentMap(0, MPIDForce.Covalent12) == [1, 2]
    box = np.diag([2.0]*3)
    f.validate(6, box)
    with pytest.raises(MPIDB200Error, match="exactly as many particles"):
        f.validate(7, box)
    # traceless checks (MPIDForceImpl.cpp:86-116)
    g = _force_from(s)
    c, d, q, o, ax, z, x, y, th, al = g.getMultipoleParameters(0)
    q[0] += 1e-3
    g.setMultipoleParameters(0, c, d, q, o, ax, z, x, y, th, al)
    with pytest.raises(MPIDB200Error, match="quadrupole"):
        g.validate(6, box)
    g = _force_from(s)
    c, d, q, o, ax, z, x, y, th, al = g.getMultipoleParameters(0)
    o[9] += 1e-3
    g.setMultipoleParameters(0, c, d, q, o, ax, z, x, y, th, al)
    with pytest.raises(MPIDB200Error, match="octopole"):
        g.validate(6, box)
    # cutoff vs box (MPIDForceImpl.cpp:61-67)
    g = _force_from(s)
    g.setNonbondedMethod(MPIDForce.PME)
    g.setCutoffDistance(1.2)
    with pytest.raises(MPIDB200Error, match="half the periodic box"):
        g.validate(6, box)
    # axis particle ranges (MPIDForceImpl.cpp:131-149)
    g = _force_from(s)
    c, d, q, o, ax, z, x, y, th, al = g.getMultipoleParameters(1)
    g.setMultipoleParameters(1, c, d, q, o, MPIDForce.ZThenX, 17, x, y, th, al)
    with pytest.raises(MPIDB200Error, match="z axis"):
        g.validate(6, box)


def test_water_box_workloads():
    s = water_box((1, 1, 1))
    assert s.n == 2988 and abs(s.box[0, 0] - 3.1289) < 1e-12 and s.grid == (32, 32, 32)
    off, idx = s.cov_csr()
    n = s.n
    # O:[H1,H2], H1:[O], H2:[O] as 1-2; H1:[H2], H2:[H1] as 1-3
    assert list(idx[off[0]:off[1]]) == [1, 2] and list(idx[off[1]:off[2]]) == [0] and list(idx[off[2]:off[3]]) == [0]
    assert list(idx[off[(n+1)+1]:off[(n+1)+2]]) == [2] and list(idx[off[(n+1)+2]:off[(n+1)+3]]) == [1]
    assert off[(n+1)] == off[(n+1)+1]        # O has no 1-3 partner
    # intramolecular geometry is water-like and molecules are whole
    d = np.linalg.norm(s.pos[1::3] - s.pos[0::3], axis=1)
    assert d.max() < 0.12 and d.min() > 0.08
    big = water_box((2, 1, 1))
    assert big.n == 2*2988 and abs(big.box[0, 0] - 2*3.1289) < 1e-12 and big.grid == (64, 32, 32)
    again = water_box((2, 1, 1))
    assert np.array_equal(big.pos, again.pos)       # seeded jitter is reproducible


def test_pair_set_helper_small_case():
    s = load_fixture("water_dimer")
    s.method = 1
    s.cutoff = 0.6
    pairs = pair_set_reference(s)
    assert len(pairs) == 15 and all(i < j for i, j, c in pairs)
    cls = {(i, j): c for i, j, c in pairs}
    assert cls[(0, 1)] == 1 and cls[(1, 2)] == 1 and cls[(0, 3)] == 0
