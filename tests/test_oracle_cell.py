"""The cell-list driver around the reference's own PME pair functions (oracle/cell_driver.cpp) against the stock
Reference-platform loops (oracle/ref_driver.cpp): with one thread it must be BIT-identical -- same pair functions, same
pairs, same order -- which is what lets it stand in for the O(N^2) platform on the 96k / 1M boxes."""
import numpy as np
import pytest

from _common import Oracle, water_box, load_fixture
from oracle.pyoracle import CellOracle


@pytest.mark.parametrize("pol", [0, 1, 2])
def test_cell_driver_is_bit_identical_to_the_stock_loops(pol):
    s = water_box((1, 1, 1), polarization=pol, epsilon=1e-6)
    e0, f0 = Oracle(s).execute()
    mu0 = Oracle(s).dipoles(0)
    c = CellOracle(s, threads=1)
    e1, f1 = c.execute()
    assert e1 == e0
    assert np.array_equal(f1, f0)
    assert np.array_equal(c.induced(), mu0)
    prof = c.profile()
    assert prof["candidate_pairs"] >= 312265           # superset of the in-cutoff pairs (312,265 for this box)
    assert prof["candidate_pairs"] < 312265 + 64       # ... by the 1e-6 shell only
    # several threads: private accumulators summed in range order -> round-off differences only
    e2, f2 = c.execute(threads=3)
    assert abs(e2 - e0) < 1e-11*abs(e0)
    assert np.abs(f2 - f0).max() < 1e-11*np.abs(f0).max()


def test_cell_driver_anisotropic_and_triclinic():
    s = water_box((1, 1, 1), polarization=0, epsilon=1e-6, anisotropic=True)
    L = s.box[0, 0]
    s.box = np.array([[L, 0, 0], [0.11*L, L, 0], [-0.07*L, 0.05*L, L]])
    e0, f0 = Oracle(s).execute()
    c = CellOracle(s, threads=1)
    e1, f1 = c.execute()
    assert e1 == e0 and np.array_equal(f1, f0)


def test_cell_driver_small_box_where_the_cell_grid_degenerates():
    # 375-atom box of the reference's own test (L = 1.8643 nm): fewer than three cells per axis
    s = load_fixture("water_375")
    s.method = 1; s.polarization = 0; s.cutoff = 0.7; s.alpha = 3.3; s.grid = (24, 24, 24); s.epsilon = 1e-6; s.default_thole = 8.0
    e0, f0 = Oracle(s).execute()
    e1, f1 = CellOracle(s, threads=2).execute()
    assert abs(e1 - e0) < 1e-11*abs(e0)
    assert np.abs(f1 - f0).max() < 1e-11*np.abs(f0).max()
