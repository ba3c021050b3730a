N=${N:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02t_bench_1m_n$N.json 2> gpurun_out/r02t_bench_1m_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02t_bench_1m_n$N.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['single_gpu_same_workload'])
print(d['stage_ms_coresident_intervals'])
PY
