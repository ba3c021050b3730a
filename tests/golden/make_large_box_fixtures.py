"""Generates tests/golden/large_box_<name>.npz: energy, a 4,096-atom subset of the forces and induced dipoles, and whole-system
norms of the jittered 95,616-atom and 1,024,884-atom water boxes bench.py times, computed by the reference's own pair
functions driven from a cell list (oracle/cell_driver.cpp; bit-identical to the Reference platform's O(N^2) loops with
one thread, round-off different with several).  Run in the build container (needs /root/reference for oracle/_ref):

    python tests/golden/make_large_box_fixtures.py [name ...]

The GPU parity tests (tests/test_gpu_scale.py) rebuild the same coordinates from the same seed and check their hash."""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpidopenmmplugin_b200.workloads import water_box  # noqa: E402
from oracle.pyoracle import CellOracle  # noqa: E402

CASES = {
    # name: (tiles, polarization, anisotropic, epsilon)
    "96k_mutual": ((4, 4, 2), 0, False, 1e-12),
    "96k_mutual_aniso": ((4, 4, 2), 0, True, 1e-12),
    "96k_direct": ((4, 4, 2), 1, False, 1e-5),
    "96k_extrapolated": ((4, 4, 2), 2, False, 1e-5),
    "96k_extrapolated_aniso": ((4, 4, 2), 2, True, 1e-5),
    "1m_mutual": ((7, 7, 7), 0, False, 1e-12),
}
SUBSET = 4096


def position_hash(pos):
    return hashlib.sha256(np.ascontiguousarray(pos, dtype=np.float64).tobytes()).hexdigest()


def main():
    names = sys.argv[1:] or list(CASES)
    threads = len(os.sched_getaffinity(0))
    for name in names:
        tiles, pol, aniso, eps = CASES[name]
        s = water_box(tiles, polarization=pol, epsilon=eps, anisotropic=aniso)
        o = CellOracle(s, threads=threads)
        t0 = time.time()
        e, f = o.execute()
        mu = o.induced()
        prof = o.profile()
        idx = np.sort(np.random.default_rng(4096).choice(s.n, SUBSET, replace=False)).astype(np.int64)
        out = os.path.join(ROOT, "tests", "golden", "large_box_%s.npz" % name)
        np.savez_compressed(out, energy=e, subset=idx, forces=f[idx], induced=mu[idx],
                            force_norm2=float(np.sum(f*f)), induced_norm2=float(np.sum(mu*mu)), force_sum=f.sum(axis=0),
                            iterations=prof["iterations"], n=s.n, epsilon=eps, pos_sha256=position_hash(s.pos),
                            threads=threads, seconds=time.time() - t0, candidate_pairs=prof["candidate_pairs"])
        print("%s: N=%d E=%.6f iterations=%d %.1f s (%d threads) -> %s" % (name, s.n, e, prof["iterations"], time.time() - t0, threads, out), flush=True)
        o.close()


if __name__ == "__main__":
    main()
