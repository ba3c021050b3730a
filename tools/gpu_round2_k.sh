# round 2, run k (8 GPUs): 1M bench at 8 ranks after the owner-computes refactor (+ 8-rank parity incl. oracle fixtures)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02k_bench_1m_n8.json 2> gpurun_out/r02k_bench_1m_n8.err
echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/multirank_check.py 4x4x2 > gpurun_out/r02k_multirank_n8.jsonl 2> gpurun_out/r02k_multirank_n8.err
echo "multirank rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench_1m_n8.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['single_gpu_same_workload'])
for k,v in list(d['kernel_us_per_evaluation'].items())[:30]: print("   %-58s %5.1f x %7.1f"%(k,v['launches'],v['us']))
print(d['stage_ms_coresident_intervals'])
PY
tail -3 gpurun_out/r02k_bench_1m_n8.err
