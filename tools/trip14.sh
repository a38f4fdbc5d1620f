#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_s3e.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3e.log | head -20
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3e.json 2> $O/bench_s3e.err; echo "default: $(cut -c60-130 $O/bench_s3e.json)"
RSS_BN_RAW=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3e_noraw.json 2>> $O/bench_s3e.err; echo "noraw: $(cut -c60-130 $O/bench_s3e_noraw.json)"
timeout 240 python tools/timeline.py s3e > $O/timeline_s3e.log 2>&1; grep "kernels in step" $O/timeline_s3e.log
