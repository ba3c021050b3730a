# round 2, run d: ncu --set full of the list-filter kernel (and the per-iteration kernels for reference)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_filter_list|k_regather" -s 6 -c 4 -o gpurun_out/r02d_full_filter python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/r02d_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
