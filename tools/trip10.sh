#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "conv_cf or wgrad or igemm" 2>&1 | tail -40 > $O/t10_unit.log
grep -E "passed|failed|FAILED|Error|assert " $O/t10_unit.log | head -20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad" --csv --log-file $O/wg_micro_launches.csv python tools/wgrad_microbench.py > $O/wg_micro.log 2>&1
for d in 0 1 2 4 3 7; do
  RSS_CF_DBG=$d timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_cf" --csv --log-file $O/cf_dbg_$d.csv python tools/cf_dbg.py > $O/cf_dbg.log 2>&1
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s2j.json 2> $O/bench_s2j.err; cut -c1-400 $O/bench_s2j.json | tr ',' '\n' | grep -E "value|frac|ms_per_launch" | head
