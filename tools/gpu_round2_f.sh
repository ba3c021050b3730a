# round 2, run f: lean filter kernel
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "list_reuse or pair_list or speculative" 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_96k.json 2> gpurun_out/r02f_bench_96k.err
head -c 300 gpurun_out/r02f_bench_96k.json; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench_96k.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['neighbour_list']['builds'], d['neighbour_list']['reuses'])
for k,v in list(d['kernel_us_per_evaluation'].items())[:8]: print(k, v)
print(d['stage_ms_coresident_intervals'])
PY
