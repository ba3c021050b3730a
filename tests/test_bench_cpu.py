"""bench.py on CPU: the reference arm's contract line (run for real on the 996-water box), what the other ranks of a
torchrun launch do in that arm, and that the roofline fractions of the committed bench line can be recomputed from the
work counts and kernel times the same line carries."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from _common import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _run_reference_arm(extra_env=None, args=()):
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "996", "--steps", "2", "--warmup", "1",
           "--no-stock-sample", "--ref-threads", "2"] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run_reference_arm()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # ONE JSON line on stdout, whatever the libraries print
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ns/day" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"] == bench.METRIC and d["steps"] == 2 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="ns/day", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert abs(d["value"] - bench.NS_PER_DAY_PER_MS/d["ms_per_step"]) < 1e-9*d["value"]
    assert "N=2988" in d["config"]["workload"] and "Mutual" in d["config"]["polarization"]
    # the energy is that of the reference's own code on this box (golden of the 996-water Mutual configuration)
    assert np.isfinite(d["energy_kj_mol"]) and d["energy_kj_mol"] < 0


def test_reference_arm_other_ranks_exit_without_work():
    r = _run_reference_arm(extra_env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, args=("--gpus", "2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and r.stdout.strip() == "" and "no CPU fallback" in r.stderr


def test_trajectory_is_seeded_and_rigid_per_molecule():
    from mpidopenmmplugin_b200.workloads import water_box
    s = water_box((1, 1, 1))
    a, b = bench.trajectory_shifts(s, 4), bench.trajectory_shifts(s, 4)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not a[0].any()
    assert np.array_equal(a[3], 3*a[1])                                    # ballistic: displacement grows linearly
    assert np.array_equal(a[1][0::3], a[1][1::3]) and np.array_equal(a[1][0::3], a[1][2::3])   # the three sites of a water move together
    assert 0.5*bench.STEP_SIGMA_NM < a[1].std() < 1.5*bench.STEP_SIGMA_NM


COMMITTED = os.path.join(ROOT, "profiles", "r02u_bench_96k.json")


@pytest.mark.skipif(not os.path.exists(COMMITTED), reason="committed bench line not present")
def test_committed_roofline_fractions_can_be_recomputed():
    d = json.loads(open(COMMITTED).read().strip().splitlines()[-1])
    w, ku, rk = d["work_counts"], d["kernel_us_per_evaluation"], d["roofline_kernels"]
    fp32 = d["fp32_peak_measured_tflops"]

    def us(sub, excl=()):
        return sum(v["us"] for k, v in ku.items() if sub in k and not any(x in k for x in excl))

    n_f = rk["k_induced_field"]["work_per_evaluation"]/(w["pol_pol"]*150.0)
    assert abs(n_f - round(n_f)) < 1e-9 and round(n_f) == d["solver_field_evaluations"][-1]
    mine = {
        # flops per unit: DESIGN.md section 4 (2240 per full x full pair, 430/2 per directed permanent-field
        # evaluation, 150 per polarizable pair and field evaluation)
        "k_electrostatics": w["full_full"]*2240.0/(us("k_electrostatics<", ("special",))*1e-6)/1e12/fp32,
        "k_fixed_field": w["fixed_field_directed"]*215.0/(us("k_fixed_field")*1e-6)/1e12/fp32,
        "k_induced_field": round(n_f)*w["pol_pol"]*150.0/(us("k_induced_field")*1e-6)/1e12/fp32,
    }
    for k, frac in mine.items():
        assert abs(frac - rk[k]["frac"]) < 0.01*frac, (k, frac, rk[k]["frac"])
    # ... and they are the numbers the round-1 verdict recomputed by hand, not round 1's 0.71 / 0.55 / 0.32
    assert abs(mine["k_electrostatics"] - 0.37) < 0.04 and abs(mine["k_fixed_field"] - 0.20) < 0.03 and abs(mine["k_induced_field"] - 0.10) < 0.015
    # the headline kernel is the one with the most device time, over all kernels
    dom = max(rk, key=lambda k: rk[k]["us_per_evaluation"])
    assert d["roofline"]["kernel"] == dom
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"]/d["roofline"]["peak"]) < 1e-12
    # same function, same inputs -> same table (bench.kernel_rooflines is what wrote the line)
    kprof = {k: (v["launches"], v["us"]) for k, v in ku.items()}
    G = 128*128*64
    again = bench.kernel_rooflines(kprof, w, 95616, w["polarizable_sites"], 95616, G, int(round(n_f)), rk["k_filter_list"]["peak"], fp32)
    for k in ("k_electrostatics", "k_fixed_field", "k_induced_field", "k_fft2_planes_forward", "k_spread_induced", "k_gather_field"):
        assert abs(again[k]["frac"] - rk[k]["frac"]) < 1e-9, k
