#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_s3i.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3i.log | head -20
timeout 240 python tools/timeline.py s3i > $O/timeline_s3i.log 2>&1; grep "kernels in step" $O/timeline_s3i.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3i_$name.json 2>> $O/bench_s3i.err; echo "$name: $(cut -c60-130 $O/bench_s3i_$name.json)"; }
run wg4 RSS_WGRAD_STREAMS=4
run bs1 RSS_BRANCH_STREAMS=1
