"""Size-independent properties at the BASELINE sizes (the O(N^2) oracle cannot run there)."""
import numpy as np
import pytest

from _common import Oracle, make_kernel, rel_err, water_box

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tiles,pol", [((4, 4, 2), 0), ((4, 4, 2), 1), ((4, 4, 2), 2), ((7, 7, 7), 0)])
def test_tiled_boxes_equal_copies_of_the_oracle_box(tiles, pol):
    """Without jitter a tiling is an exact periodic replication of the 996-water box on a grid that is also
    replicated (128x128x64 = 4x4x2 x 32^3, 224^3 = 7^3 x 32^3), so E = copies x E_996 and every image atom feels the
    same force.  The oracle value for the 996-water box therefore pins the full-size runs of BASELINE.json: config 4
    (95,616 atoms; Mutual, Direct and Extrapolated) and config 5 (1,024,884 atoms)."""
    copies = tiles[0]*tiles[1]*tiles[2]
    base = water_box((1, 1, 1), polarization=pol, epsilon=1e-7)
    e0, f0 = Oracle(base).execute()
    s = water_box(tiles, jitter=0.0, polarization=pol, epsilon=1e-7)
    assert s.n == copies*2988
    k = make_kernel(s, precision="mixed")
    f = np.zeros((s.n, 3))
    e = k.execute(s.pos, True, True, f)
    assert abs(e - copies*e0) < 1e-5*abs(copies*e0)
    fr = f.reshape(copies, 2988, 3)
    assert rel_err(fr.mean(axis=0), f0) < 1e-5
    assert np.abs(fr - fr[0]).max() < 2e-3*np.sqrt(np.mean(f0*f0))*10
    st = k.getStats()
    assert st["pairs"] + copies*2988 == copies*312265           # ordinary pairs + covalently scaled pairs
    k.close()


@pytest.mark.parametrize("prec", ["mixed"])
def test_96k_box_invariances_and_determinism(prec):
    s = water_box((4, 4, 2), polarization=0, epsilon=1e-6)
    k = make_kernel(s, precision=prec)
    f1 = np.zeros((s.n, 3)); f2 = np.zeros((s.n, 3)); f3 = np.zeros((s.n, 3))
    e1 = k.execute(s.pos, True, True, f1)
    e2 = k.execute(s.pos, True, True, f2)
    # fixed-point accumulation of forces, torques and energy in the pair stage; the only float atomics are the
    # grid spreads, so run-to-run differences stay at round-off of the grid
    assert abs(e1 - e2) < 1e-8*abs(e1)
    assert rel_err(f2, f1) < 1e-6
    # rigid translation by a lattice-incommensurate vector + whole-molecule wrapping into another image
    shift = np.array([0.3711, -1.2345, 7.7777])
    pos = s.pos + shift
    mol = np.floor(pos[0::3] @ np.linalg.inv(s.box) + np.array([0.3, 0.0, -0.4]))
    pos -= np.repeat(mol @ s.box, 3, axis=0)
    e3 = k.execute(pos, True, True, f3)
    assert abs(e3 - e1) < 2e-6*abs(e1)
    assert rel_err(f3, f1) < 2e-4           # PME discretisation error is not translation invariant
    # Newton's third law holds for the real-space part exactly and for PME to discretisation accuracy
    assert np.abs(f1.sum(axis=0)).max() < 1e-3*np.abs(f1).sum(axis=0).max()
    k.close()
