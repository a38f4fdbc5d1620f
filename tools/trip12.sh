#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_s3c.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3c.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s3c.json 2> $O/bench_s3c.err; echo "default: $(cut -c60-130 $O/bench_s3c.json)"
RSS_IGEMM_MM=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s3c_mm1.json 2>> $O/bench_s3c.err; echo "mm1: $(cut -c60-130 $O/bench_s3c_mm1.json)"
RSS_BN_KEEP_DZ=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s3c_nodz.json 2>> $O/bench_s3c.err; echo "nodz: $(cut -c60-130 $O/bench_s3c_nodz.json)"
timeout 240 python tools/timeline.py s3c > $O/timeline_s3c.log 2>&1; grep "kernels in step" $O/timeline_s3c.log
tail -5 $O/bench_s3c.err
