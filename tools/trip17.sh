#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3h_$name.json 2>> $O/bench_s3h.err; echo "$name: $(cut -c60-130 $O/bench_s3h_$name.json)"; }
run base X=1
run apply4 RSS_BN_APPLY_BPSM=4
run apply16 RSS_BN_APPLY_BPSM=16
run wg2 RSS_WGRAD_STREAMS=2
run wg3 RSS_WGRAD_STREAMS=3
