"""Several ranks against one (needs >= 2 visible GPUs; skipped on a single-GPU box): tools/multirank_check.py under
torch.distributed.run compares forces, induced dipoles and energy of the row-partitioned engine -- both reciprocal-pass
strategies, all three polarization types -- with the single-GPU engine on the same box."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,tiles", [(2, "2x2x2"), (4, "2x2x2"), (8, "4x4x2")])
def test_ranks_reproduce_the_single_gpu_evaluation(world, tiles):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    # 64^3 grid at 2x2x2 tiles (divisible by 2 and 4), 128x128x64 at 4x4x2 (divisible by 8; compared with the oracle fixtures
    # too): every reciprocal-pass strategy -- all-reduce, slab, slab + halo exchange -- is exercised at every rank count
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "tools", "multirank_check.py"), tiles]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert len(lines) == 9 and all(l["ok"] for l in lines), lines


@pytest.mark.parametrize("world,tiles", [(2, "2x2x2"), (4, "2x2x2"), (8, "4x4x2")])
def test_partitioned_host_io_returns_each_ranks_block(world, tiles):
    """mpidb200_set_host_io_partition: rank r reads only its block of the positions (the rest of its array is poisoned),
    accumulates only its block of the forces; block, energy and dipoles equal the replicated-I/O result."""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world), os.path.join(ROOT, "tools", "io_partition_check.py"), tiles, "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert len(lines) == world and all(l["ok"] for l in lines), lines
