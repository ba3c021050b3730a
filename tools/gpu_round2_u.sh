# round 2, run u: final single-GPU state -- full parity suite, bench lines (96k aniso / iso / double, 1M), reference arm, smoke
rm -f gpurun_out/parity_achieved.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02u_tests.log
cp gpurun_out/parity_achieved.jsonl gpurun_out/r02u_parity_achieved.jsonl 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench_96k.json 2> gpurun_out/r02u_bench_96k.err
timeout 600 python bench.py --steps 20 --warmup 5 --variant iso --no-cpu-baseline > gpurun_out/r02u_bench_96k_iso.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 --precision double --no-cpu-baseline > gpurun_out/r02u_bench_96k_double.json 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 5 --workload 1m --no-cpu-baseline > gpurun_out/r02u_bench_1m_n1.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02u_reference_arm.json 2> gpurun_out/r02u_reference_arm.err
cat gpurun_out/r02u_tests.log gpurun_out/r02u_smoke.log
python - <<'PY'
import json
for f in ('bench_96k','bench_96k_iso','bench_96k_double','bench_1m_n1'):
    d=json.load(open('gpurun_out/r02u_%s.json'%f)); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['neighbour_list']['builds'], d['neighbour_list']['reuses'])
d=json.load(open('gpurun_out/r02u_reference_arm.json')); print('reference', d['ms_per_step'], d['steps'], d['wall_s'])
PY
