#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3j_$name.json 2>> $O/bench_s3j.err; echo "$name: $(cut -c60-130 $O/bench_s3j_$name.json)"; }
run prio X=1
run noprio RSS_PRIORITY=0
timeout 240 python tools/timeline.py s3j > $O/timeline_s3j.log 2>&1; grep "kernels in step" $O/timeline_s3j.log
tail -3 $O/bench_s3j.err
