timeout 900 ncu --section SourceCounters --section SpeedOfLight --section WarpStateStats --section InstructionStats --import-source on --clock-control none -k regex:"k_filter_list" -s 3 -c 1 -o gpurun_out/r02g_filter python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile > gpurun_out/r02g_ncu.log 2>&1
ls -la gpurun_out/r02g_filter.ncu-rep
