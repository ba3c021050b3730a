# round 2, run t (8 GPUs): 1M bench at 8 ranks, peer-to-peer all-to-all on
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02t_bench_1m_n8.json 2> gpurun_out/r02t_bench_1m_n8.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02t_bench_1m_n8.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['single_gpu_same_workload'])
for k,v in list(d['kernel_us_per_evaluation'].items())[:16]: print("   %-58s %5.1f x %7.1f"%(k,v['launches'],v['us']))
print(d['stage_ms_coresident_intervals'])
PY
tail -3 gpurun_out/r02t_bench_1m_n8.err
