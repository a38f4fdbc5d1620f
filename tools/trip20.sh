#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check_gpu.py > $O/dist_check_s3.log 2>&1; grep -E "dist_check|Error|error" $O/dist_check_s3.log | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_s3k_2gpu.json 2> $O/bench_s3k_2gpu.err; echo "2gpu: $(cut -c1-400 $O/bench_s3k_2gpu.json)"; tail -3 $O/bench_s3k_2gpu.err
