"""Multi-rank host logic on CPU: two processes, gloo backend (the GPU data path uses NCCL inside the engine)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpidopenmmplugin_b200 import sharding  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the unique id made on rank 0 reaches every rank unchanged
        uid = sharding.broadcast_unique_id(dist, lambda: bytes((7*i + 3) % 256 for i in range(128)))
        assert uid == bytes((7*i + 3) % 256 for i in range(128))
        # 2. the row ranges tile [0, n) without gaps or overlap
        b, e = sharding.row_range(n, rank, world)
        owned = torch.zeros(n, dtype=torch.int32)
        owned[b:e] = 1
        dist.all_reduce(owned)
        assert int(owned.min()) == 1 and int(owned.max()) == 1
        # 3. partial sums over the owned rows all-reduce to the full sum (what the engine does with the partial
        #    induced field on every solver iteration); integers, like the fixed-point force buffers -> exact
        rng = np.random.default_rng(1234)
        contrib = rng.integers(-2**40, 2**40, size=(n, 3), dtype=np.int64)       # contribution of row i to atom i
        partial = torch.zeros((n, 3), dtype=torch.int64)
        partial[b:e] = torch.from_numpy(contrib[b:e])
        dist.all_reduce(partial)
        assert torch.equal(partial, torch.from_numpy(contrib))
        # 4. every special pair has exactly one owner: the rank whose rows hold its lower atom
        bounds = [sharding.row_range(n, r, world)[0] for r in range(world)] + [n]
        lo_sorted = rng.integers(0, n, size=1001)
        own = sharding.special_pair_owner(lo_sorted, bounds)
        assert np.all((lo_sorted >= np.asarray(bounds)[own]) & (lo_sorted < np.asarray(bounds)[own + 1]))
        mine = torch.from_numpy((own == rank).astype(np.int32))
        dist.all_reduce(mine)
        assert int(mine.min()) == 1 and int(mine.max()) == 1
        # 4b. owner-computes solver step: each rank owns an uneven piece of the polarizable-site list; the error overlaps
        #     are partial sums that all-reduce to the global ones, and the new dipoles travel as exactly-sized pieces
        npol = n//3
        counts = [npol//world + (3 if r == 0 else 0) for r in range(world)]
        counts[-1] = npol - sum(counts[:-1])
        begins = np.concatenate([[0], np.cumsum(counts)])
        err = np.random.default_rng(99).normal(size=(4, npol, 3))            # 4 history vectors, same on every rank
        dots = torch.tensor([np.sum(err[3, begins[rank]:begins[rank+1]]*err[k, begins[rank]:begins[rank+1]]) for k in range(4)])
        dist.all_reduce(dots)
        assert np.allclose(dots.numpy(), [np.sum(err[3]*err[k]) for k in range(4)], rtol=1e-12)
        new_mu = torch.from_numpy(err[0, begins[rank]:begins[rank+1]] + rank)
        full = sharding.owner_computes_exchange(dist, int(begins[rank]), new_mu, counts)
        expect = np.concatenate([err[0, begins[r]:begins[r+1]] + r for r in range(world)])
        assert np.array_equal(full.numpy(), expect)
        # 4c. partitioned host I/O: every rank uploads only its block of the positions (the rest of its array is poison),
        #     one in-place all-gather of equal padded blocks rebuilds the array on every rank; each rank keeps only its
        #     block of the summed forces, and the blocks tile [0, n)
        first, cnt = sharding.host_io_block(n, world, rank)
        blk = (n + world - 1)//world
        truth = rng.normal(size=(n, 3))
        staged = torch.full((world*blk, 3), float("nan"), dtype=torch.float64)
        staged[first:first + cnt] = torch.from_numpy(truth[first:first + cnt])
        pieces = [torch.empty((blk, 3), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(pieces, staged[rank*blk:(rank + 1)*blk].clone())
        gathered = torch.cat(pieces)[:n]
        assert np.array_equal(gathered.numpy(), truth)
        covered = torch.zeros(n, dtype=torch.int32)
        covered[first:first + cnt] = 1
        dist.all_reduce(covered)
        assert int(covered.min()) == 1 and int(covered.max()) == 1
        # 5. timing reduction = max over ranks
        mx = sharding.max_over_ranks(dist, [1.0 + rank, 5.0 - rank])
        assert mx == [float(world), 5.0]
        out.put((rank, "ok"))
    except Exception as ex:      # pragma: no cover
        out.put((rank, "FAIL: %r" % (ex,)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [2988, 95617])
def test_two_rank_partition_and_plumbing(n):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_collective_inventory_matches_the_engine():
    """Mutual, PME, 6 field evaluations: fixed field + fixed grid + 6 x (grid + field) + one force/torque/energy buffer; from
    4 ranks on every grid all-reduce becomes reduce-scatter + two all-to-all transposes + all-gather."""
    c = sharding.collectives_per_evaluation(0, 6, pme=True, world=2)
    assert len(c) == 2 + 12 + 1
    assert sum(1 for w in c if w[0] == "partial induced field") == 6
    c = sharding.collectives_per_evaluation(2, 3, pme=True, world=2)
    assert sum(1 for w in c if w[0] == "partial induced field gradient") == 3
    c = sharding.collectives_per_evaluation(0, 6, pme=True, world=8, grid=(224, 224, 224))
    assert len(c) == 1 + 4 + 6*(4 + 1) + 1
    assert sum(1 for w in c if w[0].endswith("all-to-all")) == 7
    assert not sharding.uses_slab_fft(8, (225, 224, 224)) and not sharding.uses_slab_fft(2, (224, 224, 224))
    # halo exchange + owner-computes solver (the default at 2/4/8 ranks on the 224^3 grid)
    assert sharding.uses_halo_exchange(8, (224, 224, 224), ncell_x=48) and not sharding.uses_halo_exchange(3, (224, 224, 224))
    c = sharding.collectives_per_evaluation(0, 6, pme=True, world=8, halo=True)
    assert len(c) == 4 + 1 + 6*(4 + 2) + 1
    assert sum(1 for w in c if w[1] == "grid") == 0                  # no full-grid collective is left
    assert sum(1 for w in c if w[0] == "error overlaps") == 6
    assert sharding.cell_column_partition(48, 8) == [0, 6, 12, 18, 24, 30, 36, 42, 48]


@pytest.mark.parametrize("world,shape", [(2, (8, 6, 10)), (4, (8, 12, 6)), (8, (16, 8, 9))])
def test_slab_reciprocal_pass_equals_the_full_transform(world, shape):
    """The slab-decomposed reciprocal pass (reduce-scatter, 2-D transforms on own planes, all-to-all, x transforms and
    influence function on own ky rows, all-to-all back, all-gather) restated in numpy with the engine's index
    arithmetic reproduces forward FFT -> influence function -> backward FFT of the summed grid."""
    rng = np.random.default_rng(7)
    nx, ny, nz = shape
    parts = [rng.normal(size=shape) for _ in range(world)]
    eterm = rng.uniform(0.1, 1.0, size=(nx, ny, nz//2 + 1))
    ref = np.fft.irfftn(eterm*np.fft.rfftn(np.sum(parts, axis=0)), s=shape, axes=(0, 1, 2))*(nx*ny*nz)
    # irfftn drops the imaginary parts a C2R transform drops, so both sides treat a non-Hermitian product alike
    got = sharding.slab_reciprocal_pass(parts, eterm)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("world,nx,ncx", [(2, 32, 7), (4, 32, 9), (8, 224, 54), (4, 64, 15)])
def test_halo_plan_covers_every_atom_of_a_rank(world, nx, ncx):
    """The plane-aligned partition planned for the halo-exchange reciprocal pass: every plane an atom of rank r spreads to
    lies in r's block or its halo, the halo is the spline support plus at most one cell column, and it fits in a slab."""
    plan = sharding.halo_plan(nx, ncx, world)
    assert plan["halo_lo"] <= sharding.PME_ORDER - 1 + -(-nx//ncx) + 1 and plan["halo_hi"] <= 1
    rng = np.random.default_rng(3)
    f = rng.random(20000)
    cell = np.minimum((f*ncx).astype(int), ncx - 1)
    last = sharding.grid_plane_of(f, nx)
    for r in range(world):
        mine = (cell >= plan["cell_lo"][r]) & (cell < plan["cell_hi"][r])
        allowed = {(plan["block_start"][r] - plan["halo_lo"] + k) % nx for k in range(plan["halo_lo"] + plan["nxl"] + plan["halo_hi"])}
        touched = {int((p - k) % nx) for p in last[mine] for k in range(sharding.PME_ORDER)}
        assert touched <= allowed, (r, sorted(touched - allowed))
    assert sorted(plan["cell_lo"]) == plan["cell_lo"] and plan["cell_hi"][-1] == ncx
    if world == 8 and nx == 224:
        assert plan["halo_lo"] <= 11 and plan["nxl"] == 28          # 1,024,884-atom box: 11 + 28 + 1 planes instead of 224


@pytest.mark.parametrize("world,nx,ncx", [(2, 32, 7), (4, 32, 9), (4, 64, 15)])
def test_halo_reciprocal_pass_equals_the_full_transform(world, nx, ncx):
    """Halo reduce -> slab transform on rotated blocks -> halo gather reproduces the full reciprocal pass on every plane a
    rank needs (its block and halo), for atoms spread with a 6-point support under the planned partition."""
    plan = sharding.halo_plan(nx, ncx, world)
    ny, nz = 4*world, 6
    rng = np.random.default_rng(11)
    f = rng.random(300)
    cell = np.minimum((f*ncx).astype(int), ncx - 1)
    last = sharding.grid_plane_of(f, nx)
    parts = [np.zeros((nx, ny, nz)) for _ in range(world)]
    for a in range(len(f)):
        r = next(q for q in range(world) if plan["cell_lo"][q] <= cell[a] < plan["cell_hi"][q])
        for k in range(sharding.PME_ORDER):
            parts[r][(last[a] - k) % nx] += rng.normal(size=(ny, nz))
    eterm = rng.uniform(0.1, 1.0, size=(nx, ny, nz//2 + 1))
    ref = np.fft.irfftn(eterm*np.fft.rfftn(np.sum(parts, axis=0)), s=(nx, ny, nz), axes=(0, 1, 2))*(nx*ny*nz)
    got = sharding.halo_reciprocal_pass(parts, eterm, plan)
    for r in range(world):
        need = [(plan["block_start"][r] - plan["halo_lo"] + k) % nx for k in range(plan["halo_lo"] + plan["nxl"] + plan["halo_hi"])]
        assert np.allclose(got[r][need], ref[need], rtol=1e-9, atol=1e-9), r
        assert np.isnan(got[r]).sum() == (nx - len(set(need)))*ny*nz


@pytest.mark.parametrize("n,world", [(2988, 2), (95616, 8), (1024884, 8), (5, 8), (7, 3), (1, 4)])
def test_host_io_blocks_tile_the_atoms(n, world):
    blocks = [sharding.host_io_block(n, world, r) for r in range(world)]
    blk = (n + world - 1)//world
    pos = 0
    for r, (first, cnt) in enumerate(blocks):
        assert cnt >= 0 and first == min(n, r*blk) and cnt <= blk
        if cnt:
            assert first == pos
            pos += cnt
    assert pos == n
    assert sharding.host_io_block(n, 1, 0) == (0, n)
