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


def special_pair_owner(num_special, world):
    """Owner rank of every covalently scaled (1-2/1-3/1-4) pair: round robin over the static pair list
    (mpid_kernels.cuh: k_special_electrostatics, `k % numRanks == rank`)."""
    return np.arange(num_special) % world


SLAB_FFT_MIN_RANKS = 4      # mpid_engine.cu: slabFftMinRanks (MPIDB200_SLAB_FFT)


def uses_slab_fft(world, grid):
    """mpid_engine.cu: useSlabFft() -- from 4 ranks on, when the x and y grid sizes divide by the rank count."""
    return world >= SLAB_FFT_MIN_RANKS and grid[0] % world == 0 and grid[1] % world == 0


def collectives_per_evaluation(polarization, field_evaluations, pme=True, world=2, grid=(224, 224, 224)):
    """Collectives one evaluation issues per rank, by payload (used for the scaling model in DESIGN.md 5):
    list of (what, element count per atom or 'grid' / 'slab', dtype bytes).  A reciprocal pass is one all-reduce of
    the charge grid (every rank then transforms the whole grid) or, with the slab decomposition, a reduce-scatter,
    two all-to-all transposes of this rank's slab and an all-gather."""
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
