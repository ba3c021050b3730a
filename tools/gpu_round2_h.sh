# round 2, run h (8 GPUs): sharded parity incl. oracle fixtures, 1M bench at 8 ranks
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/multirank_check.py 4x4x2 > gpurun_out/r02h_multirank_n8.jsonl 2> gpurun_out/r02h_multirank_n8.err
echo "multirank rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02h_bench_1m_n8.json 2> gpurun_out/r02h_bench_1m_n8.err
echo "bench rc=$?"
cut -c 1-400 gpurun_out/r02h_multirank_n8.jsonl; tail -3 gpurun_out/r02h_multirank_n8.err; head -c 300 gpurun_out/r02h_bench_1m_n8.json; tail -3 gpurun_out/r02h_bench_1m_n8.err
