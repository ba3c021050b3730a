"""Multi-rank parity check (run under torch.distributed.run, one rank per GPU): the row-partitioned engine with its
NCCL collectives against the same engine on one GPU, same box, for every polarization type and both reciprocal-pass
strategies (all-reduce + replicated FFT, slab decomposition).  Rank 0 prints one JSON line per case and exits
non-zero on a mismatch.  Usage: python -m torch.distributed.run --nproc-per-node N tools/multirank_check.py [tiles]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mpidopenmmplugin_b200 import MPIDB200Kernel
from mpidopenmmplugin_b200.workloads import water_box, make_kernel


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tiles = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (2, 2, 2)
    ok = True
    for pol, eps in ((0, 1e-6), (1, 1e-6), (2, 1e-6)):          # Mutual, Direct, Extrapolated
        s = water_box(tiles, polarization=pol, epsilon=eps)
        ref_f = np.zeros((s.n, 3))
        k1 = make_kernel(s, precision="mixed", device=local)       # every rank computes the single-GPU answer itself
        ref_e = k1.execute(s.pos, True, True, ref_f)
        ref_mu = k1.getInducedDipoles(s.pos)
        k1.close()
        for slab in (0, 2):                                       # MPIDB200_SLAB_FFT: never / from 2 ranks on
            os.environ["MPIDB200_SLAB_FFT"] = str(slab)
            k = make_kernel(s, precision="mixed", device=local)
            if rank == 0:
                uid = torch.tensor(list(MPIDB200Kernel.ncclUniqueId()), dtype=torch.uint8, device="cuda")
            else:
                uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            dist.broadcast(uid, 0)
            k.commInit(rank, world, bytes(uid.cpu().tolist()))
            f = np.zeros((s.n, 3))
            e = k.execute(s.pos, True, True, f)
            e2 = k.execute(s.pos, True, True, np.zeros((s.n, 3)))          # second call: speculative capacities, same answer
            mu = k.getInducedDipoles(s.pos)
            k.close()
            df = float(np.linalg.norm(f - ref_f)/np.linalg.norm(ref_f))
            dmu = float(np.linalg.norm(mu - ref_mu)/max(np.linalg.norm(ref_mu), 1e-300))
            de = abs(e - ref_e)/abs(ref_e)
            rec = dict(world=world, n=s.n, grid=list(s.grid), polarization=pol, slab_fft=slab, dF=df, dmu=dmu, dE=de, repeat_dE=abs(e2 - e)/abs(e))
            # the grid is summed with single-precision atomics, so a repeated evaluation agrees to ~1e-9, not bitwise
            good = df < 2e-6 and dmu < 2e-6 and de < 1e-8 and rec["repeat_dE"] < 1e-8
            ok = ok and good
            if rank == 0:
                print(json.dumps(dict(rec, ok=good)), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
