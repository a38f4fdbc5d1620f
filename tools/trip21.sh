#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { name=$1; shift; env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3l_$name.json 2>> $O/bench_s3l.err; echo "$name: $(cut -c60-130 $O/bench_s3l_$name.json)"; }
run base X=1
run after RSS_WGRAD_AFTER_DGRAD=1
run cb RSS_CUDNN_BENCHMARK=1
run cb_after RSS_CUDNN_BENCHMARK=1 RSS_WGRAD_AFTER_DGRAD=1
tail -3 $O/bench_s3l.err
