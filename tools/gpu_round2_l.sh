# round 2, run l: CudaContext entry test + compute-sanitizer records
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cuda_context" 2>&1 | tail -4
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r02l_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/r02l_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_run.py > gpurun_out/r02l_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/r02l_racecheck.log
