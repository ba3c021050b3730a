"""Developer check (run under gpurun): CUDA engine vs the oracle on the golden configurations and the
996-water box, both precisions; per-stage timings of the large boxes.  Writes gpurun_out/gpu_check.json."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _common import water_dimer, methanol_dimer, rel_err, Oracle, make_kernel, water_box, load_fixture  # noqa: E402

out = []
names = {0: "Mutual", 1: "Direct", 2: "Extrap"}


def run(tag, s, precisions=("double", "mixed"), oracle=True):
    e0 = f0 = mu0 = None
    if oracle:
        t = time.time()
        o = Oracle(s)
        e0, f0 = o.execute()
        mu0 = o.dipoles(0)
        t_or = time.time() - t
    for prec in precisions:
        try:
            k = make_kernel(s, precision=prec)
            f = np.zeros((s.n, 3))
            t = time.time()
            e = k.execute(s.pos, True, True, f)
            dt = time.time() - t
            mu = k.getInducedDipoles(s.pos)
            st = k.getStats()
            rec = dict(tag=tag, prec=prec, n=s.n, E=e, it=st["iterations"], eps=st["epsilon"], pairs=st["pairs"], ms_first=dt*1e3)
            if oracle:
                rec.update(E_ref=e0, dE=abs(e-e0)/max(abs(e0), 1e-30), dF=rel_err(f, f0), dmu=rel_err(mu, mu0) if np.linalg.norm(mu0) > 0 else float(np.linalg.norm(mu)),
                           t_oracle=t_or)
            print(json.dumps(rec), flush=True)
            out.append(rec)
            k.close()
        except Exception as ex:
            traceback.print_exc()
            out.append(dict(tag=tag, prec=prec, error=str(ex)))
            print("ERROR", tag, prec, ex, flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "golden"):
    for mk, nm in ((water_dimer, "water"), (methanol_dimer, "methanol")):
        for method in (0, 1):
            for pol in (1, 0, 2):
                s = mk(method, pol)
                run("%s-%s-%s" % (nm, "PME" if method else "NoCut", names[pol]), s)
    # charge square, 1-4 scaling (TestReferenceMPIDForce.cpp:1603-1699)
    for sc in (1.0, 0.5, 0.0):
        s = load_fixture("charge_square")
        s.method = 0; s.polarization = 1; s.scale14 = sc
        run("square-NoCut-s14=%g" % sc, s)
if which in ("all", "box"):
    for pol, eps in ((1, 1e-5), (2, 1e-5), (0, 1e-6)):
        s = water_box((1, 1, 1), polarization=pol, epsilon=eps)
        run("waterbox996-%s" % names[pol], s)
    s = water_box((1, 1, 1), polarization=0, epsilon=1e-6, anisotropic=True)
    run("waterbox996-aniso-Mutual", s)
if which in ("all", "big"):
    for tiles in ((4, 4, 2),):
        s = water_box(tiles, polarization=0, epsilon=1e-5)
        for prec in ("mixed", "double"):
            k = make_kernel(s, precision=prec, profiling=True)
            f = np.zeros((s.n, 3))
            for rep in range(3):
                f[:] = 0
                t = time.time()
                e = k.execute(s.pos, True, True, f)
                dt = time.time() - t
                st = k.getStats()
                rec = dict(tag="water%dx%dx%d" % tiles, prec=prec, n=s.n, rep=rep, E=e, ms=dt*1e3, **st)
                print(json.dumps(rec), flush=True)
                out.append(rec)
            k.setProfiling(False)
            ts = []
            for rep in range(5):
                f[:] = 0
                t = time.time()
                e = k.execute(s.pos, True, True, f)
                ts.append((time.time() - t)*1e3)
            rec = dict(tag="water%dx%dx%d-noprof" % tiles, prec=prec, ms=ts, fnorm=float(np.linalg.norm(f)))
            print(json.dumps(rec), flush=True)
            out.append(rec)
            k.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check_%s.json" % which), "w") as fh:
    json.dump(out, fh, indent=1)
