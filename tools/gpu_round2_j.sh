# round 2, run j: cluster/DSMEM plane kernels -- correctness and timing against the other generations
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "hand_written" 2>&1 | tail -6
for mode in default cluster; do
  if [ $mode = default ]; then unset MPIDB200_FFT; else export MPIDB200_FFT=$mode; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02j_bench_96k_$mode.json 2> gpurun_out/r02j_bench_96k_$mode.err
  timeout 600 python bench.py --steps 10 --warmup 5 --workload 1m --no-cpu-baseline > gpurun_out/r02j_bench_1m_$mode.json 2> gpurun_out/r02j_bench_1m_$mode.err
done
python - <<'PY'
import json
for wl in ('96k','1m'):
    for mode in ('default','cluster'):
        try:
            d=json.load(open('gpurun_out/r02j_bench_%s_%s.json'%(wl,mode)))
            k=d['kernel_us_per_evaluation']
            print(wl, mode, round(d['ms_per_step'],4), {n:(round(v['us']/v['launches'],2)) for n,v in k.items() if 'fft2' in n})
        except Exception as e: print(wl, mode, 'failed', e)
PY
