#!/usr/bin/env python
"""MPIDB200 benchmark: ms per MPIDForce evaluation and force-evaluation-limited ns/day (2 fs step) on the
synthetic SWM6-MPID water boxes named by BASELINE.json.

  python bench.py --gpus N --steps K --warmup W            # our engine (N=1: 95,616 atoms; N>1: 1,024,884 atoms, sharded)
  python bench.py --impl reference ...                     # the reference's own CPU path (oracle/_ref) on a bounded sample

One "step" = one MPIDForce energy+force evaluation (PME, mutual polarization to eps=1e-5, octopoles) on one
coordinate set.  `value`: positions and forces resident in HBM (mpidb200_execute_device).  `e2e`: the same
evaluation through mpidb200_execute with host buffers (H2D of positions and D2H of forces inside the call).
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


def stage_rooflines(stats, n, G, pairs, n_f, pk, es_pairs=None):
    """Algorithmic flops / bytes per stage over the measured CUDA-event time of that stage.
    es_pairs: pairs evaluated by the kernels the electrostatics stage timer brackets (full-full + full-charge); the
    charge-charge pairs run beside the solver on another stream and are not credited to it."""
    ms = stats["stage_ms"]
    fp32_peak = 148*128*2*pk["sm_max_mhz"]*1e6/1e12     # TFLOP/s, nominal FP32 FMA peak at max SM clock
    out = {}

    def add(name, bound, work, unit_scale, peak, unit):
        t = ms.get(name, 0.0)
        if t <= 0:
            return
        achieved = work/(t*1e-3)/unit_scale
        out[name] = dict(bound=bound, achieved=achieved, peak=peak, unit=unit, frac=achieved/peak, ms=t)

    add("electrostatics", "fp32", (pairs if es_pairs is None else es_pairs)*2240.0, 1e12, fp32_peak, "TFLOP/s")
    add("fixed_real", "fp32", pairs*430.0, 1e12, fp32_peak, "TFLOP/s")
    add("ind_real", "fp32", n_f*pairs*150.0, 1e12, fp32_peak, "TFLOP/s")
    add("fixed_spread", "hbm", n*(16+19*4) + 4.0*G + 4.0*G, 1e9, pk["hbm_gbs"], "GB/s")          # + grid clear
    add("ind_spread", "hbm", n_f*(n*(16+12) + 4.0*G + 4.0*G), 1e9, pk["hbm_gbs"], "GB/s")
    add("fft", "hbm", (n_f+1)*(8.0*G + 8.0*G + 8.0*G), 1e9, pk["hbm_gbs"], "GB/s")             # R2C + convolution + C2R
    add("fixed_gather", "hbm", 4.0*G + n*(16+35*4), 1e9, pk["hbm_gbs"], "GB/s")
    add("ind_gather", "hbm", n_f*(4.0*G + n*(16+12)) + n*35*4, 1e9, pk["hbm_gbs"], "GB/s")
    return out


def run_reference(args, rank, real_stdout):
    """The reference's own CPU implementation (Reference platform, compiled unmodified into oracle/_ref) on a
    bounded sample of the workload: the 996-water box the synthetic boxes are tiled from."""
    if rank != 0:
        return
    from oracle.pyoracle import Oracle
    from mpidopenmmplugin_b200.workloads import water_box
    wl = args.workload or ("96k" if args.gpus == 1 else "1m")
    full_n = 2988*int(np.prod(WORKLOADS[wl]["tiles"]))
    s = water_box((1, 1, 1), polarization=0, epsilon=1e-5)
    o = Oracle(s)
    steps = max(1, min(args.steps, 20))
    for _ in range(min(args.warmup, 1)):
        o.execute()
    t0 = time.perf_counter()
    for _ in range(steps):
        e, f = o.execute()
    ms = (time.perf_counter() - t0)*1e3/steps
    scale = (full_n/2988.0)**2         # the Reference platform is O(N^2): no neighbour list (MPIDReferenceForce.cpp:919-933)
    ms_full = ms*scale
    value = NS_PER_DAY_PER_MS/ms_full
    sample = "N=2988 (996-water box, same density/parameters), %d evaluations at %.1f ms; extrapolated x(N/2988)^2=%.0f to N=%d because the Reference platform visits all pairs" % (steps, ms, scale, full_n)
    line = dict(metric="ns/day (force-evaluation limited, 2 fs) of one MPIDForce eval, waterbox PME + mutual induced", value=value, unit="ns/day",
                impl="reference", n_gpus=args.gpus, steps=steps, warmup=min(args.warmup, 1), ms_per_step=ms_full, higher_is_better=True,
                scaling="strong" if args.gpus > 1 else "weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOADS[wl]["name"], polarization="Mutual eps=1e-5 (DIIS)", sample=sample),
                cpu_baseline=dict(value=value, unit="ns/day", cores=1, kind="reference", sample=sample, sample_ms_per_eval=ms),
                e2e=dict(value=value, unit="ns/day", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line, real_stdout)


def emit(line, real_stdout):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


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
    ap.add_argument("--precision", default="mixed", choices=["mixed", "double"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from mpidopenmmplugin_b200.workloads import water_box, make_kernel
    from mpidopenmmplugin_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MPIDB200 engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = args.workload or ("96k" if world == 1 else "1m")
    s = water_box(WORKLOADS[wl]["tiles"], polarization=0, epsilon=1e-5)
    n = s.n
    G = float(np.prod(s.grid))
    k = make_kernel(s, precision=args.precision, device=local_rank)
    if world > 1:
        if rank == 0:
            uid = torch.tensor(list(MPIDB200Kernel.ncclUniqueId()), dtype=torch.uint8, device="cuda")
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        k.commInit(rank, world, bytes(uid.cpu().tolist()))
    # One explicit (non-default) stream for everything: the engine runs its main work on it (mpidb200_set_stream), so
    # the L2 flush, the zeroing of the force buffer, the timing events and the evaluation are ordered on the device.
    # (Handle 0, torch's legacy default stream, would make the engine fall back to its private stream.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    k.setStream(stream.cuda_stream)
    pos_d = torch.tensor(s.pos, dtype=torch.float64, device="cuda").contiguous()
    f_d = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        f_d.zero_()
        k.execute_device(pos_d.data_ptr(), True, True, f_d.data_ptr())
    # ---- timed region: device-resident ----------------------------------------------------------------
    # The engine's per-stage timers are CUDA events recorded around every stage on three streams; they cost ~15 % at
    # this size, so the headline loop runs with them off and a second loop of the same steps (reported separately as
    # ms_per_step_with_stage_timers) provides the per-stage times the rooflines are computed from.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    energy = 0.0
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                    # L2 flush between timed iterations (not timed)
        f_d.zero_()
        ev[i][0].record(stream)
        energy = k.execute_device(pos_d.data_ptr(), True, True, f_d.data_ptr())
        ev[i][1].record(stream)
        launches += k.getStats()["launches"]
    barrier()
    t_wall = (time.perf_counter() - t_wall0)*1e3
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)/args.steps
    # ---- same steps again with the stage timers on: per-stage CUDA-event times for the rooflines --------
    k.setProfiling(True)
    stage_sum = {}
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.zero_()
        f_d.zero_()
        ev2[i][0].record(stream)
        k.execute_device(pos_d.data_ptr(), True, True, f_d.data_ptr())
        ev2[i][1].record(stream)
        for kk, v in k.getStats()["stage_ms"].items():
            stage_sum[kk] = stage_sum.get(kk, 0.0) + v
    barrier()
    prof_ms = sum(a.elapsed_time(b) for a, b in ev2)/args.steps
    k.setProfiling(False)
    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------
    f_h = np.zeros((n, 3))
    for _ in range(2):
        k.execute(s.pos, True, True, f_h)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        f_h[:] = 0.0
        e_h = k.execute(s.pos, True, True, f_h)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0)*1e3/args.steps
    barrier()
    sampler.stop_flag = True
    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    stats = k.getStats()
    if rank == 0:
        pk = peaks()
        stage_avg = {kk: v/args.steps for kk, v in stage_sum.items()}
        n_f = stats["iterations"] + 1
        pc = stats["pair_classes"]
        roofs = stage_rooflines(dict(stage_ms=stage_avg), n/world, G, float(stats["pairs"]), n_f, pk, es_pairs=float(pc["full_full"] + pc["full_charge"]))   # rank 0's share of the work over rank 0's stage times
        dominant = max(roofs.items(), key=lambda kv: kv[1]["ms"])[0] if roofs else None
        roof = dict(roofs[dominant]) if dominant else None
        if roof:
            roof["kernel"] = dominant
            # DRAM traffic of the dominant stage's kernels from the committed `ncu --set full` capture of this workload
            # (profiles/traffic.json: bytes per evaluation), null when no capture exists for it
            roof["traffic"] = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath) and world == 1:
                t = json.load(open(tpath)).get(wl, {}).get(dominant)
                if t:
                    roof["traffic"] = t["bytes"]
                    roof["traffic_source"] = t["source"]
            roof["peak_source"] = pk["source"] if roof["bound"] == "hbm" else "nominal FP32 FMA peak 148 SM x 128 lanes x 2 x %.0f MHz (pair kernels are FMA-pipe bound, not HBM or tensor bound)" % pk["sm_max_mhz"]
            roof["algorithmic_work"] = "SURVEY.md 8(d): 2240 flop per in-cutoff pair (electrostatics; pairs of the timed kernels only: full-full + full-charge), 430 (fixed field), 150 per field evaluation (induced field)"
            roof["pairs"] = dict(stats["pair_classes"], total=int(stats["pairs"]))
            # the contract's enum is hbm | tensor; the dominant stage here is FP32-FMA bound, so the longest HBM-class stage
            # (spread / FFT / gather) is given next to it with the same fields
            hbm = {kk: v for kk, v in roofs.items() if v["bound"] == "hbm"}
            if hbm:
                hk = max(hbm, key=lambda kk: hbm[kk]["ms"])
                hroof = dict(hbm[hk], kernel=hk, traffic=None, peak_source=pk["source"])
                if os.path.exists(tpath) and world == 1:
                    t = json.load(open(tpath)).get(wl, {}).get(hk)
                    if t:
                        hroof["traffic"] = t["bytes"]
                        hroof["traffic_source"] = t["source"]
                roof["hbm_dominant"] = hroof
        line = dict(metric="ns/day (force-evaluation limited, 2 fs) of one MPIDForce eval, waterbox PME + mutual induced",
                    value=NS_PER_DAY_PER_MS/dev_ms, unit="ns/day", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=dev_ms, higher_is_better=True, scaling="strong" if world > 1 else "weak", vs_baseline=None,
                    dtype="f32 pair/grid math, f64 accumulation" if args.precision == "mixed" else "f64", data="synthetic",
                    config=dict(workload=WORKLOADS[wl]["name"], polarization="Mutual eps=1e-5 (DIIS, %d field evaluations)" % n_f,
                                cutoff_nm=s.cutoff, ewald_alpha=s.alpha, l2="256 MB buffer written between timed iterations",
                                parallelism=("atom-block rows x%d: NCCL all-reduce of the partial induced field every solver iteration and of forces/torques once; "
                                             "reciprocal pass on a second communicator / stream: %s" % (world, (
                                                 "slab decomposition (reduce-scatter, 2-D FFT on own x planes, all-to-all, x FFT + influence function on own ky rows, "
                                                 "all-to-all back, all-gather)" if sharding.uses_slab_fft(world, s.grid) else
                                                 "all-reduce of the charge grid, FFT replicated"))) if world > 1 else "1 GPU",
                                note=("N>1 runs the 1,024,884-atom box of BASELINE.json config 5 (strong scaling of ONE system); the N=1 default runs the "
                                      "95,616-atom box of config 4, so values at N=1 and N>1 are different workloads -- run `--gpus 1 --workload 1m` for the "
                                      "single-GPU point of the same system") if world > 1 else "N=1 default = BASELINE.json config 4 (96k atoms, 1 B200)"),
                    e2e=dict(value=NS_PER_DAY_PER_MS/e2e_ms, unit="ns/day", ms_per_step=e2e_ms, h2d_bytes_per_step=24*n, d2h_bytes_per_step=24*n + 8),
                    gpu_launches=int(launches), energy_kj_mol=energy, wall_ms_per_step=t_wall/args.steps, ms_per_step_with_stage_timers=prof_ms,
                    clocks=sampler.summary(), roofline=roof, roofline_stages=roofs, stage_ms=stage_avg)
        if world == 1 and not args.no_cpu_baseline:
            from oracle.pyoracle import Oracle
            sb = water_box((1, 1, 1), polarization=0, epsilon=1e-5)
            o = Oracle(sb)
            reps = 4
            t0 = time.perf_counter()
            for _ in range(reps):
                o.execute()
            ms = (time.perf_counter() - t0)*1e3/reps
            scale = (n/2988.0)**2
            line["cpu_baseline"] = dict(value=NS_PER_DAY_PER_MS/(ms*scale), unit="ns/day", cores=1, kind="reference", sample_ms_per_eval=ms,
                                        sample="oracle/_ref (reference Reference-platform code, unmodified) on N=2988 (996-water box), %d evaluations at %.0f ms; x(N/2988)^2=%.0f extrapolation to N=%d (O(N^2) pair loops)" % (reps, ms, scale, n))
        emit(line, real_stdout)
    k.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
