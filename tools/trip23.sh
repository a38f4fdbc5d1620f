#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_s3m.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3m.log | head
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3m.json 2> $O/bench_s3m.err; echo "head: $(cut -c60-130 $O/bench_s3m.json)"
timeout 240 python tools/timeline.py s3m > $O/timeline_s3m.log 2>&1; grep "kernels in step" $O/timeline_s3m.log
