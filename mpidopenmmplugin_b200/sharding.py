"""Host-side plumbing of the multi-GPU path (one process per GPU): how the evaluation is partitioned and how the
ranks are wired together.  The arithmetic mirrors what the engine does on the device (mpid_engine.cu: buildNeighbors
row partition, k_special_electrostatics pair ownership) so the partition can be tested on CPU with gloo; the data
path itself (NCCL all-reduce of the partial fields, grid and forces) lives in the engine.

The reference has no multi-GPU support (platforms/cuda/src/MPIDCudaKernelFactory.cpp:69-70 uses contexts[0] only)."""
import numpy as np


def row_range(n, rank, world):
    """Contiguous range [begin, end) of SORTED atoms whose neighbour rows / PME atoms rank `rank` owns
    (mpid_engine.cu: P.rowBegin = n*rank/numRanks, P.rowEnd = n*(rank+1)/numRanks)."""
    return (n*rank)//world, (n*(rank + 1))//world


def special_pair_owner(sorted_index_of_lo, row_begins):
    """Owner rank of every covalently scaled (1-2/1-3/1-4) pair: the rank whose rows hold the pair's lower atom
    (mpid_kernels.cuh: k_own_special compacts those pairs per rank; the rank then builds the moments of both partners
    itself).  sorted_index_of_lo[k] = position of pair k's lower atom in the sorted order, row_begins = [b_0, ..., b_R]."""
    return np.searchsorted(np.asarray(row_begins)[1:], np.asarray(sorted_index_of_lo), side="right")


def cell_column_partition(ncx, world):
    """mpid_engine.cu: planHalo() -- first x cell column of every rank (and ncx at the end): rank r owns the atoms of the
    columns [lo[r], lo[r+1]), which are contiguous in the x-major sorted order."""
    return [(r*ncx)//world for r in range(world + 1)]


def owner_computes_exchange(dist, own_begin, own_values, counts):
    """The per-iteration exchange of the partitioned solver (mpid_engine.cu: gatherDipoles): every rank has new values for
    the polarizable sites among its rows (a contiguous piece of the compact site list, own_begin .. own_begin +
    counts[rank]) and needs everybody's.  Pieces differ in size, so it is point-to-point sends of exactly those pieces,
    not a padded all-gather.  Returns the assembled compact array."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    begins = np.concatenate([[0], np.cumsum(counts)])
    assert begins[rank] == own_begin and len(own_values) == counts[rank]
    out = torch.zeros((int(begins[-1]),) + tuple(own_values.shape[1:]), dtype=own_values.dtype)
    out[begins[rank]:begins[rank+1]] = own_values
    reqs = []
    for r in range(world):
        if r == rank:
            continue
        reqs.append(dist.isend(own_values.contiguous(), r))
        reqs.append(dist.irecv(out[begins[r]:begins[r+1]], r))
    for q in reqs:
        q.wait()
    return out


SLAB_FFT_MIN_RANKS = 4      # mpid_engine.cu: slabFftMinRanks (MPIDB200_SLAB_FFT)


def uses_slab_fft(world, grid):
    """mpid_engine.cu: useSlabFft() -- from 4 ranks on, when the x and y grid sizes divide by the rank count."""
    return world >= SLAB_FFT_MIN_RANKS and grid[0] % world == 0 and grid[1] % world == 0


def uses_halo_exchange(world, grid, ncell_x=None):
    """mpid_engine.cu: planHalo() -- an even rank count that divides nx and ny, at least one x cell column per rank and
    halos that fit into a block (checked by the engine; MPIDB200_HALO=0 switches it off)."""
    import os
    if os.environ.get("MPIDB200_HALO", "1") == "0":
        return False
    return world >= 2 and world % 2 == 0 and grid[0] % world == 0 and grid[1] % world == 0 and (ncell_x is None or ncell_x >= world)


def host_io_block(n, world, rank):
    """(first_atom, num_atoms) rank `rank` moves between host and device under partitioned host I/O
    (mpidb200_set_host_io_partition): equal blocks of ceil(n/world) atoms in input order, the last ones clipped to n --
    the layout of one in-place all-gather with equal counts.  The engine's rule (mpid_engine.cu, Engine::ioFirst / ioAtoms)."""
    if world <= 1:
        return 0, n
    blk = (n + world - 1)//world
    first = min(n, rank*blk)
    return first, min(n, (rank + 1)*blk) - first


def reciprocal_mode(world, grid):
    """One-line description of how the reciprocal pass is partitioned at this rank count (bench.py's config line)."""
    if world <= 1:
        return "single GPU"
    if uses_halo_exchange(world, grid):
        return ("slab decomposition with halo exchange: rows = whole x cell columns, every rank spreads into its block of nx/R planes + halo; "
                "halo reduce with the two neighbours, 2-D FFT on own planes, all-to-all, x FFT + influence function on own ky rows, all-to-all back, halo gather")
    if uses_slab_fft(world, grid):
        return ("slab decomposition (reduce-scatter, 2-D FFT on own x planes, all-to-all, x FFT + influence function on own ky rows, "
                "all-to-all back, all-gather)")
    return "all-reduce of the charge grid, FFT replicated"


def collectives_per_evaluation(polarization, field_evaluations, pme=True, world=2, grid=(224, 224, 224), halo=False):
    """Collectives one evaluation issues per rank, by payload (used for the scaling model in DESIGN.md 5):
    list of (what, element count per atom or 'grid' / 'slab' / 'halo' / 'pol' / 'scalars', dtype bytes).

    halo=False (MPIDB200_HALO=0): a reciprocal pass is one all-reduce of the charge grid (every rank then transforms the
    whole grid) or, with the slab decomposition, a reduce-scatter, two all-to-all transposes of this rank's slab and an
    all-gather; partial fields are all-reduced.
    halo=True (the default when the grid divides): halo reduce with the two neighbours, the two all-to-alls (done by
    remote stores of the transform kernels + a flag barrier when the ranks can map each other's memory), halo gather; the
    mutual solver is owner-computes: per iteration a scalar all-reduce of the <= 21 error overlaps and the exchange of
    the new dipoles of the polarizable sites; energy + forces are one all-reduce of 64-bit integers."""
    if halo:
        def hpass(what):
            return [(what + ": halo reduce", "halo", 4), (what + ": all-to-all", "slab", 8), (what + ": all-to-all back", "slab", 8),
                    (what + ": halo gather", "halo", 4)]
        out = []
        if pme:
            out += hpass("fixed charge grid")
        out.append(("dipoles of the polarizable sites (mu0)", "pol", 24))
        for _ in range(field_evaluations):
            if pme:
                out += hpass("induced-dipole grid")
            if polarization == 0:
                out.append(("error overlaps", "scalars", 8))
                out.append(("dipoles of the polarizable sites", "pol", 24))
            else:
                out.append(("partial induced field", 3, 8))
                if polarization == 2:
                    out.append(("partial induced field gradient", 6, 8))
        out.append(("energy + forces (one buffer of 64-bit integers)", 3, 8))
        return out

    def grid_pass(what):
        if uses_slab_fft(world, grid):
            return [(what + ": reduce-scatter", "grid", 4), (what + ": all-to-all", "slab", 8), (what + ": all-to-all back", "slab", 8),
                    (what + ": all-gather", "grid", 4)]
        return [(what, "grid", 4)]
    out = [("fixed field", 3, 8)]
    if pme:
        out += grid_pass("fixed charge grid")
    for _ in range(field_evaluations):
        if pme:
            out += grid_pass("induced-dipole grid")
        out.append(("partial induced field", 3, 8))
        if polarization == 2:
            out.append(("partial induced field gradient", 6, 8))
    out.append(("forces + torques + energy (one buffer of 64-bit integers)", 6, 8))
    return out


def slab_reciprocal_pass(partial_grids, eterm):
    """numpy restatement of Engine::slabReciprocalPass (mpid_engine.cu) with the ranks as list entries, data movement
    and index arithmetic as in k_slab_transpose / k_slab_convolution: partial_grids[r] is rank r's full-size real grid
    (its own atoms spread), eterm the [nx][ny][nz/2+1] influence function.  Returns the full real grid every rank holds
    after the pass; must equal irfftn(eterm * rfftn(sum of the partial grids)) * nx*ny*nz (unnormalised transforms)."""
    R = len(partial_grids)
    nx, ny, nz = partial_grids[0].shape
    nzc = nz//2 + 1
    assert nx % R == 0 and ny % R == 0
    nxl, nyl = nx//R, ny//R
    row = nyl*nzc
    total = np.sum(partial_grids, axis=0)
    # reduce-scatter: rank r receives the summed planes x in [r nxl, (r+1) nxl)
    slab_c = [np.fft.rfft2(total[r*nxl:(r + 1)*nxl], axes=(1, 2)) for r in range(R)]           # 2-D R2C on own planes
    packed = []
    for r in range(R):                          # k_slab_transpose<PACK>: [xl][q][e] -> [q][xl][e]
        src = slab_c[r].reshape(nxl*R*row)
        dst = np.empty_like(src)
        idx = np.arange(src.size)
        e = idx % row
        q = (idx // row) % R
        xl = idx // row // R
        dst[(q*nxl + xl)*row + e] = src[idx]
        packed.append(dst.reshape(R, nxl*row))
    # all-to-all: block q of rank r goes to rank q, lands in slot r
    t = [np.concatenate([packed[src_rank][r] for src_rank in range(R)]).reshape(nx, nyl, nzc) for r in range(R)]
    back = []
    for r in range(R):
        f = np.fft.fft(t[r], axis=0)
        f = f*eterm[:, r*nyl:(r + 1)*nyl, :]                       # k_slab_convolution: ky = ky0 + kyl
        back.append((np.fft.ifft(f, axis=0)*nx).reshape(R, nxl*row))
    out = np.empty((nx, ny, nz))
    for r in range(R):                          # all-to-all back, k_slab_transpose<!PACK>, 2-D C2R, all-gather
        recv = np.concatenate([back[src_rank][r] for src_rank in range(R)])
        dst = np.empty(nxl*R*row, dtype=complex)
        idx = np.arange(dst.size)
        e = idx % row
        q = (idx // row) % R
        xl = idx // row // R
        dst[idx] = recv[(q*nxl + xl)*row + e]
        out[r*nxl:(r + 1)*nxl] = np.fft.irfft2(dst.reshape(nxl, ny, nzc), s=(ny, nz), axes=(1, 2))*(ny*nz)
    return out


def broadcast_unique_id(dist, make_id, device=None):
    """Rank 0 creates the 128-byte ncclUniqueId (mpidb200_nccl_unique_id) and broadcasts it with torch.distributed
    (any backend); every rank returns the same bytes for mpidb200_comm_init."""
    import torch
    rank = dist.get_rank()
    if rank == 0:
        raw = bytes(make_id())
        assert len(raw) == 128
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def max_over_ranks(dist, values, device=None):
    """Timing reduction of bench.py: element-wise maximum of per-rank milliseconds."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


# ---------------------------------------------------------------------------------------------------------------------
# Next step for the reciprocal pass: halo exchange instead of reduce-scatter + all-gather of the full grid.
# Host-side model of the partition and the data movement; the device path is not built yet.
# ---------------------------------------------------------------------------------------------------------------------
PME_ORDER = 6


def grid_plane_of(frac, nx):
    """Last of the six x planes an atom with wrapped fractional coordinate `frac` in [0,1) spreads to.  The reference shifts
    the grid origin by half a box: fr = nx (f - round(f) + 0.5) (computeMPIDBsplines, MPIDReferenceForce.cpp:3049-3075;
    mpid_math.h: pmeAtomCell), so plane = floor(nx ((frac + 0.5) mod 1)); the support is plane-5 .. plane (mod nx)."""
    u = (np.asarray(frac, dtype=float) + 0.5) % 1.0
    return np.minimum((u*nx).astype(int), nx - 1)


def halo_plan(nx, ncx, world):
    """Plane-aligned partition for `world` ranks: rank r takes the cell columns [cell_lo[r], cell_hi[r]) of the x-major
    sort (so its atoms are one contiguous range of sorted rows, as now) and owns the nx/world planes starting at
    block_start[r] -- the uniform blocks of the slab transform, rotated by half a box like the grid origin.  halo_lo /
    halo_hi = how many planes below / above its block the atoms of a rank can reach (spline support + one cell column of
    slack from rounding the cell boundaries), which is what has to be exchanged with the neighbouring ranks."""
    if nx % world or world % 2:
        raise ValueError("halo plan needs an even rank count that divides nx")
    nxl = nx//world
    rot = nx//2                                            # = (world/2) blocks: block r of the transform is rank r's
    cell_lo = [(r*ncx)//world for r in range(world)]
    cell_hi = cell_lo[1:] + [ncx]
    block_start = [(r*nxl + rot) % nx for r in range(world)]
    lo = hi = 0
    for r in range(world):
        # extreme fractional coordinates of the rank's cells (upper end exclusive)
        first = grid_plane_of(cell_lo[r]/ncx, nx) - (PME_ORDER - 1)
        last = grid_plane_of(np.nextafter(cell_hi[r]/ncx, 0.0), nx)
        lo = max(lo, (block_start[r] - first) % nx if first != block_start[r] else 0)
        hi = max(hi, (last - (block_start[r] + nxl - 1)) % nx if (last - block_start[r]) % nx >= nxl else 0)
    if max(lo, hi) > nxl:
        raise ValueError("halo wider than a slab: too many ranks for this grid")
    return dict(world=world, nx=nx, nxl=nxl, rot=rot, cell_lo=cell_lo, cell_hi=cell_hi, block_start=block_start, halo_lo=int(lo), halo_hi=int(hi))


def halo_reciprocal_pass(partial_grids, eterm, plan):
    """numpy model of the reciprocal pass with halo exchange.  partial_grids[r]: rank r's full-size array holding its own
    atoms' spread (non-zero only on its block and halo).  Steps: (1) every rank adds the halo planes its two neighbours
    spread into its block; (2) slab transform on the owned blocks (2-D transforms, all-to-all, x transforms + influence
    function, all-to-all back); (3) every rank fetches the halo planes of the result from its neighbours.  Returns, per
    rank, a full-size array that is valid on that rank's block and halo (NaN elsewhere)."""
    R, nx, nxl = plan["world"], plan["nx"], plan["nxl"]
    lo, hi, start = plan["halo_lo"], plan["halo_hi"], plan["block_start"]
    ny, nz = partial_grids[0].shape[1:]
    planes = lambda first, count: [(first + k) % nx for k in range(count)]
    # (1) halo reduce: rank r's block receives what rank r+1 spread below its own block and rank r-1 above its own
    owned = []
    for r in range(R):
        block = planes(start[r], nxl)
        acc = partial_grids[r][block].copy()
        up, down = (r + 1) % R, (r - 1) % R
        send_from_up = planes(start[up] - lo, lo)                       # the low halo of rank r+1 = top planes of block r
        acc[nxl - lo:] += partial_grids[up][send_from_up]
        send_from_down = planes(start[down] + nxl, hi)                  # the high halo of rank r-1 = first planes of block r
        acc[:hi] += partial_grids[down][send_from_down]
        owned.append(acc)
    # (2) slab transform; block r holds planes start[r].. so the natural x order is a rotation of the rank order
    spec = [np.fft.rfft2(b, axes=(1, 2)) for b in owned]
    nzc = nz//2 + 1
    nyl = ny//R
    full = np.empty((nx, ny, nzc), dtype=complex)
    for r in range(R):
        full[planes(start[r], nxl)] = spec[r]                            # what the all-to-all assembles, ky rows split below
    out_blocks = [np.empty((nxl, ny, nzc), dtype=complex) for _ in range(R)]
    for q in range(R):                                                   # rank q transforms its ky rows along x
        rows = slice(q*nyl, (q + 1)*nyl)
        f = np.fft.ifft(np.fft.fft(full[:, rows, :], axis=0)*eterm[:, rows, :], axis=0)*nx
        for r in range(R):
            out_blocks[r][:, rows, :] = f[planes(start[r], nxl)]         # all-to-all back
    real_blocks = [np.fft.irfft2(b, s=(ny, nz), axes=(1, 2))*(ny*nz) for b in out_blocks]
    # (3) halo gather
    result = []
    for r in range(R):
        g = np.full((nx, ny, nz), np.nan)
        g[planes(start[r], nxl)] = real_blocks[r]
        down, up = (r - 1) % R, (r + 1) % R
        g[planes(start[r] - lo, lo)] = real_blocks[down][nxl - lo:]
        g[planes(start[r] + nxl, hi)] = real_blocks[up][:hi]
        result.append(g)
    return result
