# round 2, run a: parity suite, bench line, reference arm, launch list (one GPU)
nproc > gpurun_out/r02a_nproc.txt
rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02a_tests.log
cp gpurun_out/parity_achieved.jsonl gpurun_out/r02a_parity_achieved.jsonl 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_96k.json 2> gpurun_out/r02a_bench_96k.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_reference_arm.json 2> gpurun_out/r02a_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/r02a_ncu_launches.log 2>&1
tail -5 gpurun_out/r02a_tests.log; head -c 1500 gpurun_out/r02a_bench_96k.json; tail -3 gpurun_out/r02a_bench_96k.err
