"""The oracle (reference stack compiled unmodified) against the reference's own golden vectors
(platforms/reference/tests/TestReferenceMPIDForce.cpp), and the host build of mpid_math.h against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from _common import (ROOT, Oracle, emul_evaluate, load_fixture, methanol_dimer, rel_err, water_dimer)

# energies (kJ/mol) hard-coded in the reference test, file:line = TestReferenceMPIDForce.cpp
GOLDEN = {
    ("water", 0, 1): -1.949902453,     # :1329  NoCutoff Direct
    ("water", 0, 0): -1.952917117,     # :1430  NoCutoff Mutual
    ("water", 0, 2): -1.94668563,      # :1583  NoCutoff Extrapolated
    ("water", 1, 1): -2.523318862,     # :1381  PME Direct
    ("water", 1, 0): -2.533082539,     # :1482  PME Mutual
    ("water", 1, 2): -2.527846018,     # :1535  PME Extrapolated
    ("methanol", 0, 1): 100.1426571,   # :1055
    ("methanol", 0, 0): 100.1424251,   # :1166
    ("methanol", 0, 2): 100.1424271,   # :1275
    ("methanol", 1, 1): 100.048119,    # :1003
    ("methanol", 1, 0): 100.0480699,   # :1113
    ("methanol", 1, 2): 100.0480906,   # :1223
}
# first force vectors printed in the reference test
GOLDEN_F0 = {
    ("water", 1, 0): (-140.0801113, -184.8502938, 30.90206227),     # :1484
    ("water", 0, 0): (-139.7835608, -184.4337529, 35.62953533),     # :1432
    ("methanol", 1, 1): (0.4407512632, 0.9533272891, 0.2662227116),  # :1005
}
MAKERS = {"water": water_dimer, "methanol": methanol_dimer}


def assert_equal_tol(expected, found, tol):
    """ASSERT_EQUAL_TOL of OpenMM's AssertionUtilities: relative with a floor of 1."""
    scale = max(1.0, abs(expected))
    assert abs(expected - found)/scale <= tol, (expected, found)


@pytest.mark.parametrize("key", sorted(GOLDEN.keys()))
def test_oracle_reproduces_reference_goldens(key):
    name, method, pol = key
    s = MAKERS[name](method, pol)
    e, f = Oracle(s).execute()
    assert_equal_tol(GOLDEN[key], e, 1e-4)
    assert abs(e - GOLDEN[key]) < 5e-7*max(1.0, abs(e))     # the printed digits
    if key in GOLDEN_F0:
        for a, b in zip(GOLDEN_F0[key], f[0]):
            assert abs(a - b) < 2e-6*max(1.0, abs(a))


def test_reference_own_test_binary_passes():
    exe = os.path.join(ROOT, "oracle", "_ref", "TestReferenceMPIDForce")
    if not os.path.exists(exe):
        pytest.skip("reference test binary not built (no /root/reference here)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("scale,expected", [(1.0, -1389.35), (0.5, -694.675), (0.0, 0.0)])
def test_oracle_14_scaling(scale, expected):
    """test14ScalingNoCutoff (:1603-1650): four unit charges on a square, only 1-4 pairs interact."""
    s = load_fixture("charge_square")
    s.method = 0
    s.polarization = 1
    s.scale14 = scale
    e, _ = Oracle(s).execute()
    assert abs(e - expected) < 1e-2


@pytest.mark.parametrize("key", sorted(GOLDEN.keys()))
def test_math_header_matches_oracle(key):
    """mpid_math.h (the arithmetic the CUDA kernels inline), compiled for the host, in FP64 and FP32."""
    name, method, pol = key
    s = MAKERS[name](method, pol)
    o = Oracle(s)
    e0, f0 = o.execute()
    mu0 = o.dipoles(0)
    e, f, mu, it = emul_evaluate(s, use_float=False)
    tol = 1e-6 if pol == 0 else 1e-12        # mutual: both sides stop at eps, with different linear solvers
    assert rel_err(f, f0) < tol and rel_err(mu, mu0) < max(tol, 1e-6 if pol == 0 else 1e-12)
    assert abs(e - e0) < 1e-7*max(1.0, abs(e0))
    e, f, mu, it = emul_evaluate(s, use_float=True)
    assert rel_err(f, f0) < 1e-5
    assert rel_err(mu, mu0) < 5e-5


def test_math_header_anisotropic_mutual_water():
    """Anisotropic polarizability on a Bisector site with mutual polarization: exercises the frame-dependent
    induced-induced torque term of the reference (MPIDReferenceForce.cpp:4877-4878)."""
    from mpidopenmmplugin_b200.workloads import ANISO_ALPHA_O
    s = load_fixture("water_375")
    s.method = 1; s.cutoff = 0.9; s.alpha = 3.0; s.grid = (20, 20, 20); s.default_thole = 8.0
    s.epsilon = 1e-9; s.max_iter = 200; s.polarization = 0
    for i in range(0, 375, 3):
        s.alphas[i] = ANISO_ALPHA_O
    o = Oracle(s)
    e0, f0 = o.execute()
    e, f, mu, it = emul_evaluate(s, use_float=False)
    assert rel_err(f, f0) < 1e-6
    assert abs(e - e0) < 1e-6*abs(e0)


def test_math_header_all_axis_types_triclinic():
    """Every axis type, chirality flips, 1-4/1-5 classes and a reduced triclinic box against the oracle."""
    from _common import random_molecule_box
    s = random_molecule_box()
    s.polarization = 1
    o = Oracle(s)
    e0, f0 = o.execute()
    e, f, mu, it = emul_evaluate(s, use_float=False)
    assert rel_err(f, f0) < 1e-10 and abs(e - e0) < 1e-10*abs(e0)
    assert rel_err(mu, o.dipoles(0)) < 1e-10
