"""CPU model of redLine6 (mpid_kernels.cuh): a 6-point z line of the spread leaves as reductions on the aligned 4-float
quads that cover it, zeros in the unused lanes; a quad that would pass the end of the row is replaced by scalar adds.
Checks the invariants the kernel relies on for every row length and alignment: the padded writes never pass the row end
(so never the end of the grid), they never start before the grid, and the grid ends up as if six scalars had been added."""
import numpy as np
import pytest


def red_line6(grid, p, v, room):
    """grid: flat float array whose element 0 is 16-byte aligned; p: index of the first point; room: floats to the row end."""
    mis = p & 3
    q = p - mis
    quads = 3 if mis == 3 else 2
    if 4*quads - mis > room:
        for k in range(6):
            grid[p + k] += v[k]
        return [(p + k, 1) for k in range(6)]
    w = np.zeros(12, dtype=grid.dtype)
    w[mis:mis + 6] = v
    for b in range(quads):
        grid[q + 4*b:q + 4*b + 4] += w[4*b:4*b + 4]
    return [(q + 4*b, 4) for b in range(quads)]


@pytest.mark.parametrize("nz", [6, 7, 8, 9, 10, 12, 15, 18, 30, 32, 33, 64])
def test_padded_reductions_stay_inside_the_row(nz):
    rows = 5
    rng = np.random.default_rng(nz)
    for r in range(rows):
        for z0 in range(0, nz - 5):                      # lines that do not wrap in z (the kernel's condition ig.z + 5 < nz)
            grid = np.zeros(rows*nz, dtype=np.float32)
            ref = grid.copy()
            v = rng.normal(size=6).astype(np.float32)
            p = r*nz + z0
            ops = red_line6(grid, p, v, nz - z0)
            ref[p:p + 6] += v
            assert np.array_equal(grid, ref)
            for start, width in ops:
                assert start >= 0 and start + width <= (r + 1)*nz          # never before the grid, never past this row's end
                assert width == 1 or start % 4 == 0                         # vector requests are 16-byte aligned
            if nz % 4 == 0 and z0 + 9 <= nz:
                assert len(ops) <= 3 and all(w == 4 for _, w in ops)        # aligned rows away from the end: 2-3 vector requests
