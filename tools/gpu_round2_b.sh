# round 2, run b: list reuse -- parity suite + bench
rm -f gpurun_out/parity_achieved.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02b_tests.log
cp gpurun_out/parity_achieved.jsonl gpurun_out/r02b_parity_achieved.jsonl 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_96k.json 2> gpurun_out/r02b_bench_96k.err
MPIDB200_NO_LIST_REUSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_96k_noreuse.json 2> gpurun_out/r02b_bench_96k_noreuse.err
timeout 600 python bench.py --steps 10 --warmup 5 --workload 1m --no-cpu-baseline > gpurun_out/r02b_bench_1m_n1.json 2> gpurun_out/r02b_bench_1m_n1.err
tail -5 gpurun_out/r02b_tests.log; head -c 600 gpurun_out/r02b_bench_96k.json; tail -3 gpurun_out/r02b_bench_96k.err; head -c 400 gpurun_out/r02b_bench_1m_n1.json
