# round 2, run q: early permanent-field kernel from the candidate list (side stream) + list bookkeeping off the main stream
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -k "golden or waterbox_996 or list_reuse or axis_types or jittered or speculative or anisotropic" 2>&1 | tail -4
for ef in 1 0; do
  MPIDB200_EARLY_FIXED=$ef timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02q_bench_96k_early$ef.json 2> gpurun_out/r02q_bench_96k_early$ef.err
done
MPIDB200_EARLY_FIXED=1 timeout 600 python bench.py --steps 10 --warmup 5 --workload 1m --no-cpu-baseline > gpurun_out/r02q_bench_1m_early1.json 2>/dev/null
python - <<'PY'
import json
for f in ('96k_early1','96k_early0','1m_early1'):
    try:
        d=json.load(open('gpurun_out/r02q_bench_%s.json'%f)); print(f, round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), {k:round(v,3) for k,v in d['stage_ms_coresident_intervals'].items() if k in ('nlist','fixed_real','solver')})
    except Exception as e: print(f, 'failed', e)
PY
