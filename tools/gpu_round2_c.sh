# round 2, run c (2 GPUs): sharded parity for every reciprocal strategy + 1M bench at 2 ranks
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/multirank_check.py 2x2x2 > gpurun_out/r02c_multirank_n2.jsonl 2> gpurun_out/r02c_multirank_n2.err
echo "multirank rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r02c_bench_1m_n2.json 2> gpurun_out/r02c_bench_1m_n2.err
echo "bench rc=$?"
cat gpurun_out/r02c_multirank_n2.jsonl; tail -5 gpurun_out/r02c_multirank_n2.err; head -c 700 gpurun_out/r02c_bench_1m_n2.json; tail -5 gpurun_out/r02c_bench_1m_n2.err
