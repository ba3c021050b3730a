# Round 2 -- the command list behind profiles/r02*_ (one B200, gpurun).  Outputs go to gpurun_out/, the summaries that are
# meant to be judged are copied to profiles/ by hand (tools/summarize_launches.py, tools/summarize_ncu_full.py).
TAG=${TAG:-r02n}
rm -f gpurun_out/parity_achieved.jsonl
if [ -z "$SKIP_TESTS" ]; then
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
cp gpurun_out/parity_achieved.jsonl gpurun_out/${TAG}_parity_achieved.jsonl 2>/dev/null
fi
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_96k.json 2> gpurun_out/${TAG}_bench_96k.err
timeout 600 python bench.py --steps 20 --warmup 5 --variant iso --no-cpu-baseline > gpurun_out/${TAG}_bench_96k_iso.json 2> gpurun_out/${TAG}_bench_96k_iso.err
timeout 600 python bench.py --steps 20 --warmup 5 --precision double --no-cpu-baseline > gpurun_out/${TAG}_bench_96k_double.json 2> gpurun_out/${TAG}_bench_96k_double.err
timeout 600 python bench.py --steps 10 --warmup 5 --workload 1m --no-cpu-baseline > gpurun_out/${TAG}_bench_1m_n1.json 2> gpurun_out/${TAG}_bench_1m_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_filter_list|k_fixed_field|k_charge_site_pairs|k_electrostatics|k_simple_pairs|k_special_electrostatics|k_reciprocal_terms|k_lab_frame|k_half_compact|k_special_field_finish|k_regather" -s 30 -c 12 -o gpurun_out/${TAG}_full_a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/${TAG}_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_spread|k_gather|k_fft2|k_induced_field|k_special_field<|k_diis" -s 200 -c 14 -o gpurun_out/${TAG}_full_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/${TAG}_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_filter_list|k_induced_field|k_fft2|k_spread|k_gather|k_electrostatics|k_charge_site" -s 120 -c 10 -o gpurun_out/${TAG}_full_1m python bench.py --steps 1 --warmup 3 --workload 1m --no-cpu-baseline --no-kernel-profile > gpurun_out/${TAG}_ncu_1m.log 2>&1
for r in full_a full_b full_1m; do ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_$r.raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_$r.ncu-rep; done
cat gpurun_out/${TAG}_tests.log 2>/dev/null; head -c 400 gpurun_out/${TAG}_bench_96k.json; echo; head -c 300 gpurun_out/${TAG}_reference_arm.json; echo; ls -la gpurun_out | grep ${TAG}
