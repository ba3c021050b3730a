"""Developer check (gpurun): run-to-run reproducibility of the mutual solver on the 996-water box."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import make_kernel, water_box, rel_err  # noqa: E402

for prec in ("double", "mixed"):
    for eps in (1e-5, 1e-7, 1e-9):
        s = water_box((1, 1, 1), polarization=0, epsilon=eps)
        k = make_kernel(s, precision=prec)
        fs, mus, its = [], [], []
        for r in range(4):
            f = np.zeros((s.n, 3))
            e = k.execute(s.pos, True, True, f)
            st = k.getStats()
            fs.append(f); its.append((st["iterations"], st["epsilon"], e))
            mus.append(k.getInducedDipoles(s.pos))
        print(prec, eps, "iters/eps/E:", its)
        print("   max|dF| vs run0:", [float(np.abs(f - fs[0]).max()) for f in fs[1:]], " rel:", [rel_err(f, fs[0]) for f in fs[1:]])
        print("   rel dmu vs run0:", [rel_err(m, mus[0]) for m in mus[1:]])
        k.close()
