# round 2, run i: full single-GPU parity suite after the multi-rank refactors
rm -f gpurun_out/parity_achieved.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02i_tests.log
cp gpurun_out/parity_achieved.jsonl gpurun_out/r02i_parity_achieved.jsonl 2>/dev/null
cat gpurun_out/r02i_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench_96k.json 2> gpurun_out/r02i_bench_96k.err
head -c 300 gpurun_out/r02i_bench_96k.json; tail -2 gpurun_out/r02i_bench_96k.err
