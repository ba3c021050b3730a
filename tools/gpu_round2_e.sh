# round 2, run e: hand-written transforms incl. 224 family; ncu of the filter kernel; 1M single-GPU bench
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "hand_written or fused_reciprocal or list_reuse" 2>&1 | tail -15 > gpurun_out/r02e_tests.log
cat gpurun_out/r02e_tests.log
timeout 600 python bench.py --steps 10 --warmup 5 --workload 1m --no-cpu-baseline > gpurun_out/r02e_bench_1m_n1.json 2> gpurun_out/r02e_bench_1m_n1.err
head -c 500 gpurun_out/r02e_bench_1m_n1.json; tail -3 gpurun_out/r02e_bench_1m_n1.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_96k.json 2> gpurun_out/r02e_bench_96k.err
head -c 400 gpurun_out/r02e_bench_96k.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_filter_list|k_regather" -s 6 -c 3 -o gpurun_out/r02e_full_filter python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/r02e_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
