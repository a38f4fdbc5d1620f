#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for v in 4 2 1 8; do
RSS_BN_TICKET_BPSM=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_s3g_bpsm$v.json 2>> $O/bench_s3g.err; echo "bpsm$v: $(cut -c60-130 $O/bench_s3g_bpsm$v.json)"
done
