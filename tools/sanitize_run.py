"""Small evaluations for compute-sanitizer (memcheck / racecheck): water dimer (no-cutoff and PME), the 996-water box in
every polarization mode with a second (list-reuse) evaluation, and a 224-point grid edge for the single-buffer transform
kernels.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import water_dimer  # noqa: E402
from mpidopenmmplugin_b200.workloads import make_kernel, water_box  # noqa: E402


def run(s, prec="mixed", steps=2):
    k = make_kernel(s, precision=prec)
    f = np.zeros((s.n, 3))
    for i in range(steps):
        e = k.execute(s.pos + 0.001*i, True, True, f)
    mu = k.getInducedDipoles(s.pos)
    k.close()
    return e, float(np.abs(mu).max())


print("dimer nocutoff", run(water_dimer(0, 0)))
print("dimer pme", run(water_dimer(1, 0), "double"))
for pol in (0, 1, 2):
    print("996 box pol", pol, run(water_box((1, 1, 1), polarization=pol)))
print("996 box double", run(water_box((1, 1, 1), polarization=0), "double", 1))
print("224 grid edge", run(water_box((1, 1, 1), polarization=1, grid=(224, 32, 32)), "mixed", 1))
