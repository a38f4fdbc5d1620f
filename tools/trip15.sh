#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "bn or model or traj or step" 2>&1 | tail -40 > $O/pytest_s3f.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3f.log | head -20
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3f.json 2> $O/bench_s3f.err; echo "raw: $(cut -c60-130 $O/bench_s3f.json)"
RSS_BN_RAW=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3f_noraw.json 2>> $O/bench_s3f.err; echo "noraw: $(cut -c60-130 $O/bench_s3f_noraw.json)"
timeout 240 python tools/timeline.py s3f > $O/timeline_s3f.log 2>&1; grep "kernels in step" $O/timeline_s3f.log
