#!/bin/bash
# run w (2 GPUs): partitioned host I/O -- correctness at 24k atoms, the N>1 bench line on the 96k box, timing on the 1M box
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 70 $TR --master-port 29611 tools/io_partition_check.py 2x2x2 5 > gpurun_out/r02w_io_24k.jsonl 2> gpurun_out/r02w_io_24k.err; echo "io24k rc=$?"
cat gpurun_out/r02w_io_24k.jsonl | cut -c1-600
timeout 110 $TR --master-port 29612 bench.py --gpus 2 --workload 96k --steps 5 --warmup 3 --no-kernel-profile > gpurun_out/r02w_bench_96k_n2.json 2> gpurun_out/r02w_bench_96k_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02w_bench_96k_n2.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], json.dumps(d['e2e'])[:900])
except Exception as ex:
    print('bench line unreadable', ex)
PY
tail -3 gpurun_out/r02w_bench_96k_n2.err
timeout 75 $TR --master-port 29613 tools/io_partition_check.py 7x7x7 3 > gpurun_out/r02w_io_1m.jsonl 2> gpurun_out/r02w_io_1m.err; echo "io1m rc=$?"
cat gpurun_out/r02w_io_1m.jsonl | cut -c1-600
