# round 2, run v (4 GPUs): final code, every reciprocal strategy x polarization type against one GPU and the oracle fixtures
timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 tools/multirank_check.py 4x4x2 > gpurun_out/r02v_multirank_n4.jsonl 2> gpurun_out/r02v_multirank_n4.err
echo "multirank rc=$?"
cut -c 1-120 gpurun_out/r02v_multirank_n4.jsonl; tail -2 gpurun_out/r02v_multirank_n4.err
