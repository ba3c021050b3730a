python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/f_tests.log
python bench.py > gpurun_out/f_bench_96k.json 2> gpurun_out/f_bench_96k.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_reference_arm.json 2> gpurun_out/f_reference_arm.err
python bench.py --gpus 1 --workload 1m --no-cpu-baseline > gpurun_out/f_bench_1m_n1.json 2> gpurun_out/f_bench_1m_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:"k_neighbor_list_cell|k_fixed_field|k_charge_site_pairs|k_electrostatics|k_simple_pairs|k_special_electrostatics|k_reciprocal_terms|k_lab_frame|k_half_compact|k_special_field_finish|k_spline_weights|k_sorted_sites" -s 48 -c 12 -o gpurun_out/f_full_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_a.log 2>&1
ncu --set full --clock-control none -k regex:"k_spread|k_gather|k_fft2|k_induced_field|k_special_field<|k_diis" -s 240 -c 14 -o gpurun_out/f_full_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_b.log 2>&1
MPIDB200_TRACE=gpurun_out/f_trace_96k.csv python tools/trace_run.py 96k > gpurun_out/f_trace_96k.log 2>&1
cat gpurun_out/f_tests.log; ls -la gpurun_out | grep " f_"
