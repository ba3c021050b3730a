"""Partitioned host I/O check (run under torch.distributed.run, one rank per GPU): with
mpidb200_set_host_io_partition every rank reads only its block of the positions from its host array (the other blocks
arrive from the other ranks over NVLink) and accumulates only its block of the forces.  Compared here with replicated
I/O of the same engine: the block has to agree, everything outside the block has to stay untouched, the energy and the
induced dipoles have to agree on every rank -- for page-locked and for pageable caller arrays, and with a DIFFERENT,
deliberately wrong position array outside the block (it must not be read).  Every rank prints one JSON line; exit
code non-zero on a mismatch.  Usage: python -m torch.distributed.run --nproc-per-node N tools/io_partition_check.py [tiles] [steps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mpidopenmmplugin_b200 import MPIDB200Kernel, sharding
from mpidopenmmplugin_b200.workloads import water_box, make_kernel


def rel(a, b):
    return float(np.linalg.norm(a - b)/max(np.linalg.norm(b), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tiles = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (2, 2, 2)
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    s = water_box(tiles, polarization=0, epsilon=1e-6)
    n = s.n
    k = make_kernel(s, precision="mixed", device=local)
    uid = torch.tensor(list(MPIDB200Kernel.ncclUniqueId()), dtype=torch.uint8, device="cuda") if rank == 0 else torch.zeros(128, dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    k.commInit(rank, world, bytes(uid.cpu().tolist()))
    pos = np.ascontiguousarray(s.pos, dtype=np.float64)
    f_rep = np.zeros((n, 3))
    e_rep = k.execute(pos, True, True, f_rep)
    mu_rep = k.getInducedDipoles(pos)

    def timed(p, f):
        for _ in range(2):
            k.execute(p, True, True, f)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            k.execute(p, True, True, f)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0)*1e3/steps
        dist.barrier()
        return ms

    pos_pin, f_pin = pos.copy(), np.zeros((n, 3))
    k.pinHostBuffer(pos_pin)
    k.pinHostBuffer(f_pin)
    ms_rep = timed(pos_pin, f_pin)

    k.setHostIoPartition(True)
    first, cnt = k.getHostIoBlock()
    assert (first, cnt) == sharding.host_io_block(n, world, rank), (first, cnt)
    outside = np.ones(n, dtype=bool)
    outside[first:first + cnt] = False
    # positions outside the block are poisoned: they must come from the owners, not from this array
    poisoned = pos.copy()
    poisoned[outside] = 1.0e3
    rec = dict(rank=rank, world=world, n=n, first_atom=first, num_atoms=cnt)
    good = True
    for label, pin in (("pageable", False), ("pinned", True)):
        p = poisoned.copy()
        f = np.full((n, 3), 7.0)                     # the engine ACCUMULATES: 7 + force inside the block, 7 outside
        if pin:
            k.pinHostBuffer(p)
            k.pinHostBuffer(f)
        e = k.execute(p, True, True, f)
        mu = k.getInducedDipoles(p)
        if pin:
            k.unpinHostBuffer(p)
            k.unpinHostBuffer(f)
        rec[label] = dict(dF_block=rel(f[first:first + cnt] - 7.0, f_rep[first:first + cnt]), dE=abs(e - e_rep)/abs(e_rep), dmu=rel(mu, mu_rep),
                          outside_untouched=bool((f[outside] == 7.0).all()))
        r = rec[label]
        good = good and r["dF_block"] < 1e-6 and r["dE"] < 1e-8 and r["dmu"] < 1e-6 and r["outside_untouched"]
    pos_pin[outside] = 1.0e3
    ms_part = timed(pos_pin, f_pin)
    k.setHostIoPartition(False)
    # back to replicated I/O: the full result again
    f_back = np.zeros((n, 3))
    e_back = k.execute(pos, True, True, f_back)
    rec["replicated_again_dF"] = rel(f_back, f_rep)
    good = good and rec["replicated_again_dF"] < 1e-6 and abs(e_back - e_rep)/abs(e_rep) < 1e-8
    rec.update(ms_per_step_replicated_io=ms_rep, ms_per_step_partitioned_io=ms_part, ok=bool(good))
    k.close()
    for r in range(world):
        if r == rank:
            print(json.dumps(rec), flush=True)
        dist.barrier()
    flag = torch.tensor([0 if good else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
