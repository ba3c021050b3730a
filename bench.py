#!/usr/bin/env python
"""MPIDB200 benchmark: ms per MPIDForce evaluation and force-evaluation-limited ns/day (2 fs step) on the
synthetic SWM6-MPID water boxes named by BASELINE.json.

  python bench.py --gpus N --steps K --warmup W            # our engine (N=1: 95,616 atoms; N>1: 1,024,884 atoms, sharded)
  python bench.py --impl reference ...                     # the reference's own CPU pair functions on the SAME box

One "step" = one MPIDForce energy+force evaluation (PME, mutual polarization to eps=1e-5, octopoles, anisotropic
polarizability on O) on one coordinate set of a short synthetic ballistic trajectory (every water moves rigidly with
its own seeded thermal velocity, so no step sees the coordinates of the one before).
`value`: positions and forces resident in HBM (mpidb200_execute_device).  `e2e`: the same evaluation through
mpidb200_execute with HOST buffers (H2D of positions, H2D + D2H of the accumulated forces inside the timed call).
`roofline` / `roofline_kernels`: per KERNEL, from a third pass in which every launch runs alone between two CUDA
events (mpidb200_set_kernel_profiling); work = the pairs / bytes that kernel actually processes.
`cpu_baseline` / `--impl reference`: the reference's own pair functions driven from a cell list (oracle/cell_driver.cpp,
bit-identical to the Reference platform's O(N^2) loops) on the same coordinates, all host threads, measured -- not
extrapolated; `parity` compares the GPU result with it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NS_PER_DAY_PER_MS = 172.8          # 2 fs step: ns/day = 172.8 / (ms per evaluation)   (examples/waterbox/run.py:17)
WORKLOADS = {
    "996":  dict(tiles=(1, 1, 1), name="SWM6-MPID water box, 996 waters, N=2988, L=3.1289 nm, grid 32^3"),
    "96k":  dict(tiles=(4, 4, 2), name="synthetic SWM6-MPID water box, N=95616, 12.5156x12.5156x6.2578 nm, grid 128x128x64"),
    "1m":   dict(tiles=(7, 7, 7), name="synthetic SWM6-MPID water box, N=1024884, L=21.9023 nm, grid 224^3"),
}
METRIC = "ns/day (force-evaluation limited, 2 fs) of one MPIDForce eval, waterbox PME + mutual induced"
STEP_SIGMA_NM = 0.00074            # per-step, per-component displacement of every water in the timed trajectory: thermal
                                   # velocity of a water molecule at 300 K, sqrt(kT/m) = 0.37 nm/ps, times the 2 fs step
# Algorithmic flops per unit.  2240 / 430 / 150: SURVEY.md 8(d), counted from the oracle's generic
# routines (430 and 150 cover both directions of a pair; the gather kernels evaluate directions: 215 / 75 each).
# 325 / 48: the pair classes the oracle has no routine for, counted the same way (tools/count_flops.cpp).
FLOP_FULL_FULL, FLOP_FIXED_DIRECTED, FLOP_INDUCED_PAIR, FLOP_FULL_CHARGE, FLOP_CHARGE_CHARGE = 2240.0, 215.0, 150.0, 325.0, 48.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_system(args, world):
    from mpidopenmmplugin_b200.workloads import water_box
    wl = args.workload or ("96k" if world == 1 else "1m")
    s = water_box(WORKLOADS[wl]["tiles"], polarization=0, epsilon=1e-5, anisotropic=(args.variant == "aniso"))
    name = WORKLOADS[wl]["name"] + (", anisotropic polarizability on O (north-star target variant)" if args.variant == "aniso" else ", isotropic polarizability")
    return wl, s, name


def trajectory_shifts(s, count, seed=777):
    """Ballistic synthetic trajectory: every water moves rigidly with a constant seeded thermal velocity; step k uses
    s.pos + k*v.  (Displacements grow linearly, as between two neighbour-list rebuilds of a real MD run.)"""
    rng = np.random.default_rng(seed)
    v = np.repeat(rng.normal(0.0, STEP_SIGMA_NM, size=(s.n//3, 3)), 3, axis=0)
    return [k*v for k in range(count)]


def kernel_rooflines(kprof, work, n, n_pol, rows, G, n_f, hbm_peak, fp32_peak):
    """Per-kernel rooflines.  kprof: {name: (launches per evaluation, us per evaluation)} from the kernel-profile pass
    (each launch alone on the device).  Work per evaluation = what the kernel's launches process (this rank's share)."""
    def find(sub, excl=()):
        hits = [(k, v) for k, v in kprof.items() if sub in k and not any(x in k for x in excl)]
        if not hits:
            return None
        return (sum(v[0] for _, v in hits), sum(v[1] for _, v in hits))

    out = {}

    def add(label, key, bound, work_amount, note=None, excl=()):
        r = find(key, excl)
        if r is None or r[1] <= 0:
            return
        launches, us = r
        if bound == "fp32":
            achieved, peak, unit = work_amount/(us*1e-6)/1e12, fp32_peak, "TFLOP/s"
        else:
            achieved, peak, unit = work_amount/(us*1e-6)/1e9, hbm_peak, "GB/s"
        d = dict(bound=bound, achieved=achieved, peak=peak, unit=unit, frac=achieved/peak, us_per_evaluation=us, launches_per_evaluation=launches,
                 us_per_launch=us/max(launches, 1), work_per_evaluation=work_amount, work_per_launch=work_amount/max(launches, 1))
        if note:
            d["note"] = note
        out[label] = d

    P = float(work["pairs"])
    # pair kernels: FP32 FMA bound
    add("k_electrostatics", "k_electrostatics<", "fp32", work["full_full"]*FLOP_FULL_FULL, "full x full pairs x 2240 flop (quasi-internal frame)", excl=("special",))
    add("k_charge_site_pairs", "k_charge_site_pairs", "fp32", work["full_charge"]*FLOP_FULL_CHARGE, "full x bare-charge pairs x 325 flop (tools/count_flops.cpp)")
    add("k_simple_pairs", "k_simple_pairs", "fp32", work["charge_charge"]*FLOP_CHARGE_CHARGE, "charge x charge pairs x 48 flop (tools/count_flops.cpp)")
    add("k_fixed_field", "k_fixed_field", "fp32", work["fixed_field_directed"]*FLOP_FIXED_DIRECTED, "polarizable sites x their neighbours x 215 flop (one direction of SURVEY's 430)")
    add("k_induced_field", "k_induced_field", "fp32", n_f*work["pol_pol"]*FLOP_INDUCED_PAIR, "%d field evaluations x polarizable x polarizable pairs x 150 flop" % n_f)
    # neighbour search: integer / FP32 issue bound; the bytes figure is the floor (positions in, two list entries per pair + the
    # polarizable list out)
    add("k_neighbor_list_cell", "k_neighbor_list", "hbm", 16.0*rows + 8.0*P + 8.0*work["pol_pol"] + 16.0*rows,
        "issue-slot bound integer/FP32 search (ncu: 66 % issue active); bytes = float4 positions in + 2 list entries per pair + polarizable list out")
    add("k_filter_list", "k_filter_list", "hbm", 4.0*work.get("candidate_entries", 0) + 16.0*rows + 8.0*P + 8.0*work["pol_pol"] + 16.0*rows,
        "candidate entries in + float4 positions + 2 list entries per pair + polarizable list out")
    add("cub_radix_sort", "cub_radix_sort", "hbm", 4*16.0*n, "4 onesweep passes over (key, index) pairs")
    # reciprocal space: HBM class (the grid lives in L2 at these sizes)
    add("k_spread_fixed", "k_spread<real, true>", "hbm", rows*(16 + 19*4) + 4.0*G, "N x (position + 19 fractional moments) + grid")
    add("k_spread_induced", "k_spread<real, false>", "hbm", n_f*(n_pol*(16 + 12) + 4.0*G), "per pass: polarizable sites x (position + dipole) + grid")
    add("k_fft2_planes_forward", "k_fft2_planes_forward", "hbm", (n_f + 1)*8.0*G, "per pass: real grid in, half-complex grid out")
    add("k_fft2_x_convolve", "k_fft2_x_convolve", "hbm", (n_f + 1)*10.0*G, "per pass: half-complex grid in and out + influence function")
    add("k_fft2_planes_backward", "k_fft2_planes_backward", "hbm", (n_f + 1)*8.0*G, "per pass: half-complex grid in, real grid out")
    add("cufft_forward", "cufft_forward", "hbm", (n_f + 1)*8.0*G)
    add("cufft_backward", "cufft_backward", "hbm", (n_f + 1)*8.0*G)
    add("k_convolution", "k_convolution", "hbm", (n_f + 1)*10.0*G)
    add("k_gather_35", "k_gather<real, 4, false>", "hbm", 2*(4.0*G + rows*(16 + 35*4)), "2 launches: grid + 35 derivatives per site")
    add("k_gather_field", "k_gather<real, 1, true>", "hbm", n_f*(4.0*G + n_pol*(16 + 12)), "per pass: grid + field at polarizable sites")
    add("k_diis_step", "k_diis_step", "hbm", n_f*n_pol*(12 + 48 + 24 + 24 + 48 + 48 + 24.0*(n_f + 1)/2),
        "per iteration per polarizable site: reciprocal field, alpha, E_fixed, E_induced, mu in/out, history in/out, error overlaps")
    add("k_lab_frame", "k_lab_frame", "hbm", n*(24 + 20*8 + 16 + 20*8 + 16*8 + 20*4 + 16*4 + 16*8 + 48 + 4.0), "per atom: parameters in, both moment packings (FP64 + FP32), alpha tensor out")
    add("k_reciprocal_terms", "k_reciprocal_terms", "hbm", rows*(2*35*4 + 20*8 + 16*8 + 24 + 48.0), "per atom: phi, phi_dp, moments in; force/torque atomics out")
    return out


def run_reference(args, rank, real_stdout):
    """The reference's own CPU implementation of the path on the SAME box and coordinates as the GPU arm: its PME pair
    functions, reciprocal-space and solver code (oracle/_ref, compiled unmodified) with the three O(N^2) pair loops replaced
    by a cell list that visits the same pairs in the same order (oracle/cell_driver.cpp), on every host thread."""
    if rank != 0:
        return
    from oracle.pyoracle import CellOracle, Oracle
    from mpidopenmmplugin_b200.workloads import water_box
    world = args.gpus
    wl, s, name = build_system(args, world)
    threads = host_threads() if args.ref_threads <= 0 else args.ref_threads
    o = CellOracle(s, threads=threads)
    shifts = trajectory_shifts(s, 4)
    budget = float(os.environ.get("MPIDB200_REF_BUDGET_S", "240"))
    t_begin = time.perf_counter()
    warm = min(args.warmup, 1)
    for _ in range(warm):
        o.execute(s.pos)
    t_warm = time.perf_counter() - t_begin
    per = t_warm if warm else None
    steps_wanted = max(1, args.steps)
    times = []
    for k in range(steps_wanted):
        if times and (time.perf_counter() - t_begin) + np.mean(times) > budget:
            break                                  # bounded: the whole run has to end within minutes (1M box: ~2.5 min per step)
        t0 = time.perf_counter()
        e, f = o.execute(s.pos + shifts[k % len(shifts)])
        times.append(time.perf_counter() - t0)
    steps = len(times)
    ms = float(np.mean(times))*1e3
    prof = o.profile()
    value = NS_PER_DAY_PER_MS/ms
    # the stock O(N^2) loops, for the record: measured on the 996-water box, law checked at N=11,952 (profiles/r02_reference_arm.md)
    stock = None
    if not args.no_stock_sample:
        sb = water_box((1, 1, 1), polarization=0, epsilon=1e-5, anisotropic=(args.variant == "aniso"))
        ob = Oracle(sb)
        ob.execute()
        t0 = time.perf_counter()
        ob.execute()
        t_stock = (time.perf_counter() - t0)*1e3
        stock = dict(measured_ms_at_N2988=t_stock, extrapolated=True, same_config=False,
                     extrapolated_ms_at_this_N=t_stock*(s.n/2988.0)**2,
                     note="stock Reference-platform loops visit all N^2/2 pairs (MPIDReferenceForce.cpp:919-933, 4084-4088, 4932-4946); "
                          "(N/2988)^2 extrapolation, NOT used for value")
    sample = ("full %d-atom box, %d evaluation(s) of %d requested (budget %.0f s), %d host threads, cell-list driver around the reference's own pair functions "
              "(bit-identical to the stock loops with 1 thread: tests/test_oracle_cell.py)" % (s.n, steps, steps_wanted, budget, threads))
    line = dict(metric=METRIC, value=value, unit="ns/day",
                impl="reference", n_gpus=args.gpus, steps=steps, warmup=warm, ms_per_step=ms, higher_is_better=True,
                scaling="strong" if args.gpus > 1 else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=name, polarization="Mutual eps=1e-5 (DIIS, %d field evaluations)" % prof["induced_field_evaluations"],
                            cutoff_nm=s.cutoff, ewald_alpha=s.alpha, trajectory="ballistic: every water moves rigidly with its own N(0, %.5f nm/step) thermal velocity (seed 777)" % STEP_SIGMA_NM),
                cpu_baseline=dict(value=value, unit="ns/day", cores=threads, kind="reference", sample=sample, ms_per_eval=ms,
                                  seconds_by_part={k: float(prof[k]) for k in ("candidates_s", "fixed_field_s", "induced_fields_s", "electrostatics_s", "total_s")}),
                e2e=dict(value=value, unit="ns/day", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                energy_kj_mol=e, stock_loops=stock, wall_s=time.perf_counter() - t_begin)
    emit(line, real_stdout)


def emit(line, real_stdout):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def rel_err(a, b):
    return float(np.linalg.norm(a - b)/max(np.linalg.norm(b), 1e-300))


def main():
    # Libraries loaded below (NCCL's version banner, torch warnings) write to fd 1; the contract is ONE JSON line on
    # stdout, so fd 1 is pointed at stderr for the whole run and the line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS.keys()))
    ap.add_argument("--variant", default="aniso", choices=["aniso", "iso"], help="polarizability of the O site (north star: anisotropic)")
    ap.add_argument("--precision", default="mixed", choices=["mixed", "double"])
    ap.add_argument("--solver", default="diis", choices=["diis", "cg"],
                    help="mutual induced-dipole solver: diis = the reference's (MPIDReferenceForce.cpp:1182-1252, what parity is against); "
                         "cg = the preconditioned conjugate-gradient alternative BASELINE.json config 4 names (same fixed point, same epsilon test)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock-sample", action="store_true")
    ap.add_argument("--no-kernel-profile", action="store_true")
    ap.add_argument("--ref-threads", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from mpidopenmmplugin_b200 import MPIDB200Kernel
    from mpidopenmmplugin_b200.workloads import make_kernel
    from mpidopenmmplugin_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MPIDB200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl, s, wl_name = build_system(args, world)
    n = s.n
    G = float(np.prod(s.grid))
    k = make_kernel(s, precision=args.precision, device=local_rank, solver=args.solver)
    if world > 1:
        if rank == 0:
            uid = torch.tensor(list(MPIDB200Kernel.ncclUniqueId()), dtype=torch.uint8, device="cuda")
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        k.commInit(rank, world, bytes(uid.cpu().tolist()))
    # One explicit (non-default) stream for everything: the engine runs its main work on it (mpidb200_set_stream), so
    # the L2 flush, the zeroing of the force buffer, the timing events and the evaluation are ordered on the device.
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    k.setStream(stream.cuda_stream)
    # the timed trajectory: warm-up and timed steps all see different coordinates
    total = args.warmup + args.steps
    shifts = trajectory_shifts(s, total)
    pos_h = [np.ascontiguousarray(s.pos + sh) for sh in shifts]
    pos_d = [torch.tensor(p, dtype=torch.float64, device="cuda").contiguous() for p in pos_h]
    f_d = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        f_d.zero_()
        k.execute_device(pos_d[i].data_ptr(), True, True, f_d.data_ptr())
    # ---- timed region: device-resident ----------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    energy = 0.0
    iters = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                    # L2 flush between timed iterations (not timed)
        f_d.zero_()
        ev[i][0].record(stream)
        energy = k.execute_device(pos_d[args.warmup + i].data_ptr(), True, True, f_d.data_ptr())
        ev[i][1].record(stream)
        st = k.getStats()
        launches += st["launches"]
        iters.append(st["iterations"])
    barrier()
    t_wall = (time.perf_counter() - t_wall0)*1e3
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)/args.steps
    stats = k.getStats()
    work = k.getWorkCounts()
    ls = k.getListStats()
    list_stats = dict(ls, note="evaluations of warm-up + timed loop that sorted and searched (builds) / reused the order and the skin-padded candidate list "
                               "(reuses); skin 0.1 nm, rebuild when an atom has moved skin/2; the exact in-cutoff list is re-derived every evaluation")
    # ---- same steps with the stage timers on (intervals on three co-resident streams: informational) ----
    k.setProfiling(True)
    stage_sum = {}
    for i in range(args.steps):
        flush.zero_()
        f_d.zero_()
        k.execute_device(pos_d[args.warmup + i].data_ptr(), True, True, f_d.data_ptr())
        for kk, v in k.getStats()["stage_ms"].items():
            stage_sum[kk] = stage_sum.get(kk, 0.0) + v
    barrier()
    k.setProfiling(False)
    # ---- per-kernel times: every launch alone on the device between two events ---------------------------
    kprof = {}
    if not args.no_kernel_profile:
        k.setKernelProfiling(True)
        for i in range(min(args.steps, 5)):
            f_d.zero_()
            k.execute_device(pos_d[args.warmup + i].data_ptr(), True, True, f_d.data_ptr())
        kprof = k.getKernelProfile()
        k.setKernelProfiling(False)
        barrier()
    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------
    # The caller's arrays are page-locked once (mpidb200_pin_host_buffer), as the platform kernel does for the Context's
    # position / force vectors; every step: H2D of that step's coordinates and of the caller's (zeroed) forces, the
    # evaluation, D2H of the accumulated forces and of the energy.
    f_h = np.zeros((n, 3))
    k.pinHostBuffer(f_h)
    for p in pos_h:
        k.pinHostBuffer(p)

    def e2e_loop(steps):
        for i in range(2):
            f_h.fill(0.0)
            k.execute(pos_h[i], True, True, f_h)
        barrier()
        f_h.fill(0.0)
        t0 = time.perf_counter()
        for i in range(steps):
            # (the engine ACCUMULATES into the caller's forces, so they are uploaded, added to on the device and read back
            # every step; clearing them between steps is the host framework's business and is not part of this call)
            k.execute(pos_h[args.warmup + i], True, True, f_h)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0)*1e3/steps
        barrier()
        return ms

    io = None
    if world > 1:
        # every rank moves ALL positions / forces over its own PCIe link (replicated I/O) ...
        e2e_replicated_ms = e2e_loop(min(args.steps, 5))
        f_rep = np.zeros((n, 3))
        e_rep = k.execute(s.pos, True, True, f_rep)
        # ... against partitioned I/O: rank r moves its block of atoms only, position blocks all-gathered over NVLink
        k.setHostIoPartition(True)
        first, cnt = k.getHostIoBlock()
        e2e_ms = e2e_loop(args.steps)
        f_blk = np.zeros((n, 3))
        e_blk = k.execute(s.pos, True, True, f_blk)
        k.setHostIoPartition(False)
        outside = np.ones(n, dtype=bool)
        outside[first:first + cnt] = False
        io = dict(first_atom=first, num_atoms=cnt, dF_block=rel_err(f_blk[first:first + cnt], f_rep[first:first + cnt]),
                  dE=abs(e_blk - e_rep)/abs(e_rep), outside_block_untouched=bool(not f_blk[outside].any()))
        io["ok"] = bool(io["dF_block"] < 1e-6 and io["dE"] < 1e-7 and io["outside_block_untouched"])
    else:
        e2e_ms = e2e_loop(args.steps)
        e2e_replicated_ms = e2e_ms
    for p in pos_h:
        k.unpinHostBuffer(p)
    # pageable (not pinned) caller arrays, for comparison
    f_p = np.zeros((n, 3))
    t0 = time.perf_counter()
    for i in range(min(args.steps, 5)):
        k.execute(pos_h[args.warmup + i], True, True, f_p)
    torch.cuda.synchronize()
    e2e_pageable_ms = (time.perf_counter() - t0)*1e3/min(args.steps, 5)
    barrier()
    sampler.stop_flag = True
    # ---- several ranks: the same workload on ONE GPU in the same run, and parity of the sharded result against it ----
    single = None
    if world > 1:
        f_ref = np.zeros((n, 3))
        f_sh = np.zeros((n, 3))
        e_sh = k.execute(s.pos, True, True, f_sh)
        mu_sh = k.getInducedDipoles(s.pos)
        k1 = make_kernel(s, precision=args.precision, device=local_rank, solver=args.solver)
        k1.setStream(stream.cuda_stream)
        for i in range(3):
            f_d.zero_()
            k1.execute_device(pos_d[i].data_ptr(), True, True, f_d.data_ptr())
        ev1 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for i in range(5):
            flush.zero_()
            f_d.zero_()
            ev1[i][0].record(stream)
            k1.execute_device(pos_d[args.warmup + i].data_ptr(), True, True, f_d.data_ptr())
            ev1[i][1].record(stream)
        torch.cuda.synchronize()
        one_ms = sum(a.elapsed_time(b) for a, b in ev1)/5
        e_one = k1.execute(s.pos, True, True, f_ref)
        mu_one = k1.getInducedDipoles(s.pos)
        k1.close()
        single = dict(ms_per_step=one_ms, value=NS_PER_DAY_PER_MS/one_ms, unit="ns/day", steps=5,
                      note="same box, same coordinates, one engine without a communicator on this rank's GPU, measured in this run",
                      parity_of_sharded_result=dict(dF=rel_err(f_sh, f_ref), dmu=rel_err(mu_sh, mu_one), dE=abs(e_sh - e_one)/abs(e_one)))
        barrier()
    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms, e2e_pageable_ms, e2e_replicated_ms, 0.0 if (io is None or io["ok"]) else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_pageable_ms, e2e_replicated_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    io_partitioned = world > 1 and float(t[4]) == 0.0
    if world > 1 and not io_partitioned:
        e2e_ms = e2e_replicated_ms          # a block that differs from the replicated result is not a result: report replicated I/O
    if rank == 0:
        pk = peaks()
        fp32_peak = MPIDB200Kernel.measureFp32Peak(local_rank)
        stage_avg = {kk: v/args.steps for kk, v in stage_sum.items()}
        n_f = stats["iterations"] + 1
        rows = n//world if world > 1 else n
        roofs = kernel_rooflines(kprof, work, n, work["polarizable_sites"], rows, G, n_f, pk["hbm_gbs"], fp32_peak)
        # dominant kernel = the one with the most device time per evaluation, whatever its bound
        roof = None
        if roofs:
            dom = max(roofs, key=lambda kk: roofs[kk]["us_per_evaluation"])
            roof = dict(roofs[dom], kernel=dom)
            roof["achieved_definition"] = "algorithmic work of the kernel's launches in one evaluation / their summed duration, each launch alone on the device (CUDA events, warm L2)"
            roof["peak_source"] = (pk["source"] if roof["bound"] == "hbm" else
                                   "measured in this run: FP32 FMA-chain micro-benchmark (mpidb200_measure_fp32_peak), nominal 148 SM x 128 lanes x 2 x %.0f MHz = %.1f" % (pk["sm_max_mhz"], 148*128*2*pk["sm_max_mhz"]*1e-6))
            roof["traffic"] = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath) and world == 1:
                tj = json.load(open(tpath)).get(wl, {}).get(dom)
                if tj:
                    roof["traffic"] = tj.get("bytes_per_launch", tj["bytes"])          # DRAM bytes per launch (ncu --set full, cold L2)
                    roof["traffic_source"] = tj["source"]
                    roof["traffic_over_algorithmic"] = roof["traffic"]/max(roof["work_per_launch"], 1.0) if roof["bound"] == "hbm" else None
            fp = {kk: v for kk, v in roofs.items() if v["bound"] == "fp32"}
            hb = {kk: v for kk, v in roofs.items() if v["bound"] == "hbm" and kk != dom}
            if fp:
                kk = max(fp, key=lambda q: fp[q]["us_per_evaluation"])
                roof["fp32_dominant"] = dict(fp[kk], kernel=kk)
            if hb:
                kk = max(hb, key=lambda q: hb[q]["us_per_evaluation"])
                roof["hbm_next"] = dict(hb[kk], kernel=kk)
        kernel_us = {kk: dict(launches=v[0], us=v[1]) for kk, v in sorted(kprof.items(), key=lambda kv: -kv[1][1])}
        line = dict(metric=METRIC,
                    value=NS_PER_DAY_PER_MS/dev_ms, unit="ns/day", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=dev_ms, higher_is_better=True, scaling="strong" if world > 1 else "weak", vs_baseline=None,
                    dtype="f32 pair/grid math, f64 accumulation" if args.precision == "mixed" else "f64", data="synthetic",
                    config=dict(workload=wl_name, polarization="Mutual eps=1e-5 (%s, %d field evaluations)" % ("DIIS" if args.solver == "diis" else "conjugate gradient", n_f),
                                cutoff_nm=s.cutoff, ewald_alpha=s.alpha, l2="256 MB buffer written between timed iterations",
                                trajectory="ballistic: every water moves rigidly with its own N(0, %.5f nm/step) thermal velocity (seed 777); every warm-up and timed step has its own coordinates" % STEP_SIGMA_NM,
                                parallelism=("owner-computes rows x%d (cell columns along x): halo planes of the grid, the solver's overlaps and own dipoles exchanged per iteration "
                                             "(peer-to-peer stores + flag barriers when the ranks can map each other's memory), one NCCL all-reduce of energy + forces per evaluation; "
                                             "reciprocal pass: %s" % (world, sharding.reciprocal_mode(world, s.grid)))
                                if world > 1 else "1 GPU",
                                note=("N>1 runs the 1,024,884-atom box of BASELINE.json config 5 (strong scaling of ONE system); the N=1 default runs the "
                                      "95,616-atom box of config 4, so the same-workload single-GPU time is measured in THIS run: single_gpu_same_workload") if world > 1
                                else "N=1 default = BASELINE.json config 4 (96k atoms, 1 B200)"),
                    e2e=dict(value=NS_PER_DAY_PER_MS/e2e_ms, unit="ns/day", ms_per_step=e2e_ms,
                             h2d_bytes_per_step=48*n if (world == 1 or io_partitioned) else 48*n*world,
                             d2h_bytes_per_step=(24*n + 8*world) if (world == 1 or io_partitioned) else (24*n + 8)*world,
                             ms_per_step_pageable_arrays=e2e_pageable_ms,
                             note="host arrays page-locked once with mpidb200_pin_host_buffer; per step: positions H2D, caller's forces H2D (accumulated on the device), forces D2H, energy D2H; "
                                  "bytes are summed over all ranks; "
                                  "wall-clocked back to back WITHOUT the L2 flush the device-resident loop runs between its steps, which is why it can come out below `value`"),
                    gpu_launches=int(launches), energy_kj_mol=energy, wall_ms_per_step=t_wall/args.steps,
                    solver_field_evaluations=[int(i) + 1 for i in iters],
                    clocks=sampler.summary(), roofline=roof, roofline_kernels=roofs, kernel_us_per_evaluation=kernel_us,
                    work_counts=work, fp32_peak_measured_tflops=fp32_peak, neighbour_list=list_stats,
                    stage_ms_coresident_intervals=stage_avg)
        if single:
            line["single_gpu_same_workload"] = single
        if world > 1:
            line["e2e"]["host_io"] = dict(mode="partitioned" if io_partitioned else "replicated", ms_per_step_replicated_io=e2e_replicated_ms,
                                          rank0_block_check=io,
                                          note="partitioned (mpidb200_set_host_io_partition): every rank passes full-length arrays, moves only its block of atoms over its PCIe link, "
                                               "position blocks are all-gathered over NVLink; checked in this run against the replicated-I/O result of the same coordinates "
                                               "(all ranks must agree, else the replicated time is reported)")
        if world == 1 and not args.no_cpu_baseline:
            # the reference's own pair functions (cell-list driven) on the SAME coordinates, all host threads: measured baseline
            # and parity of the GPU result in one go
            from oracle.pyoracle import CellOracle
            threads = host_threads()
            o = CellOracle(s, threads=threads)
            t0 = time.perf_counter()
            e_ref, f_ref = o.execute(s.pos)
            cpu_ms = (time.perf_counter() - t0)*1e3
            mu_ref = o.induced()
            prof = o.profile()
            f_g = np.zeros((n, 3))
            e_g = k.execute(s.pos, True, True, f_g)
            mu_g = k.getInducedDipoles(s.pos)
            line["cpu_baseline"] = dict(value=NS_PER_DAY_PER_MS/cpu_ms, unit="ns/day", cores=threads, kind="reference", ms_per_eval=cpu_ms,
                                        sample="one evaluation of the full %d-atom box (same coordinates as step 0), %d host threads: the reference's own PME pair functions, "
                                               "reciprocal-space and DIIS code (oracle/_ref, unmodified) driven from a cell list (oracle/cell_driver.cpp)" % (n, threads),
                                        seconds_by_part={kk: float(prof[kk]) for kk in ("candidates_s", "fixed_field_s", "induced_fields_s", "electrostatics_s", "total_s")})
            line["parity"] = dict(against="cpu_baseline evaluation (reference arithmetic, FP64), both at eps=1e-5", dF=rel_err(f_g, f_ref), dmu=rel_err(mu_g, mu_ref),
                                  dE=abs(e_g - e_ref)/abs(e_ref), energy_reference=e_ref, energy_gpu=e_g, tolerance_mixed=1e-5)
        emit(line, real_stdout)
    k.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
