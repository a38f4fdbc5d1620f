#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > $O/pytest_s3b.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3b.log | head
RSS_BRANCH_STREAMS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s3b_bs1.json 2> $O/bench_s3b.err; echo "bs1: $(cut -c60-130 $O/bench_s3b_bs1.json)"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s3b_bs2.json 2>> $O/bench_s3b.err; echo "bs2: $(cut -c60-130 $O/bench_s3b_bs2.json)"
timeout 240 python tools/timeline.py s3b > $O/timeline_s3b.log 2>&1; tail -4 $O/timeline_s3b.log | cut -c1-200
