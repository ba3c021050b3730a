"""The OpenMM-facing plugin layer (mpidopenmmplugin_b200/plugin): exports, loud failure without a GPU, and -- on the
GPU -- the reference's own CUDA test file (compiled unmodified) and a Reference-platform comparison through
System / MPIDForce / Context / State, both running on the MPIDB200 kernel.

The binaries are prebuilt by __graft_entry__.build() in the build container (they need the MPID API headers of the
reference checkout) and travel to the GPU box in oracle/_ref/."""
import os
import re
import subprocess

import numpy as np
import pytest

from _common import ROOT, water_box, subset_waters

PLUGIN = os.path.join(ROOT, "mpidopenmmplugin_b200", "plugin", "libMPIDPluginB200.so")
REF_CUDA_TEST = os.path.join(ROOT, "oracle", "_ref", "TestCudaMPIDForce_on_MPIDB200")
B200_TEST = os.path.join(ROOT, "oracle", "_ref", "TestB200MPIDForce")


def _need(path):
    if not os.path.exists(path):
        pytest.skip("%s has not been built (needs /root/reference; run __graft_entry__.build())" % os.path.relpath(path, ROOT))


def test_plugin_exports_the_registration_entry_points():
    """registerPlatforms / registerKernelFactories / registerMPID<Platform>KernelFactories: the three-function pattern of
    the reference's platform plugins (platforms/cuda/src/MPIDCudaKernelFactory.cpp:36-66)."""
    _need(PLUGIN)
    out = subprocess.run(["nm", "-D", "--defined-only", PLUGIN], capture_output=True, text=True).stdout
    for nm in ("registerPlatforms", "registerKernelFactories", "registerMPIDB200KernelFactories"):
        assert re.search(r"\bT %s\b" % nm, out), nm
    assert "B200CalcMPIDForceKernel" in subprocess.run(["nm", "-DC", "--defined-only", PLUGIN], capture_output=True, text=True).stdout
    # the plugin carries no arithmetic of its own: it needs the engine's C ABI
    und = subprocess.run(["nm", "-D", "--undefined-only", PLUGIN], capture_output=True, text=True).stdout
    for nm in ("mpidb200_create", "mpidb200_execute", "mpidb200_set_particles", "mpidb200_get_dipoles"):
        assert nm in und, nm


def test_plugin_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _need(REF_CUDA_TEST)
    r = subprocess.run([REF_CUDA_TEST], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device available" in r.stdout and "there is no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_reference_cuda_test_file_passes_on_mpidb200(precision):
    """All 14 cases of the reference's platforms/cuda/tests/TestCudaMPIDForce.cpp (golden energies and forces of the water
    and methanol dimers, NoCutoff and PME, Direct / Mutual / Extrapolated, and the 1-4 scaling energies)."""
    _need(REF_CUDA_TEST)
    r = subprocess.run([REF_CUDA_TEST, precision], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("Done")


def _write_waters(path, s):
    with open(path, "w") as f:
        f.write("%d %.10f %.10f %.10f\n" % (s.n//3, s.box[0][0], s.box[1][1], s.box[2][2]))
        for p in s.pos:
            f.write("%.17g %.17g %.17g\n" % tuple(p))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_plugin_matches_reference_platform_through_context(tmp_path, precision):
    """MPIDB200 vs. the reference's Reference platform in one process: forces, energy, the three dipole queries, PME
    parameters, system multipole moments, electrostatic potential, updateParametersInContext and error behaviour."""
    _need(B200_TEST)
    s = subset_waters(water_box((1, 1, 1)), 332)
    path = os.path.join(str(tmp_path), "waters.txt")
    _write_waters(path, s)
    r = subprocess.run([B200_TEST, path, precision], capture_output=True, text=True, timeout=900)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout and r.stdout.strip().endswith("Done")
