timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cuda_context" 2>&1 | tail -12
grep cuda-context gpurun_out/parity_achieved.jsonl
