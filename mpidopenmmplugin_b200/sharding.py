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


def collectives_per_evaluation(polarization, field_evaluations, pme=True):
    """All-reduces one evaluation issues per rank, by payload (used for the scaling model in DESIGN.md 5):
    list of (what, element count per atom or 'grid', dtype bytes)."""
    out = [("fixed field", 3, 8)]
    if pme:
        out.append(("fixed charge grid", "grid", 4))
    for _ in range(field_evaluations):
        if pme:
            out.append(("induced-dipole grid", "grid", 4))
        out.append(("partial induced field", 3, 8))
        if polarization == 2:
            out.append(("partial induced field gradient", 6, 8))
    out += [("forces", 3, 8), ("torques", 3, 8), ("energy", 0, 8)]
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
