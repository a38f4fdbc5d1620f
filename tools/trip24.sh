#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_s3n.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3n.log | head
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3n.json 2> $O/bench_s3n.err; echo "idx32: $(cut -c60-130 $O/bench_s3n.json)"
timeout 120 python tools/timeline.py s3n > $O/timeline_s3n.log 2>&1; grep "kernels in step" $O/timeline_s3n.log
