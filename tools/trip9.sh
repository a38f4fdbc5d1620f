#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "conv_cf or wgrad" 2>&1 | tail -40 > $O/t9_unit.log
grep -E "passed|failed|FAILED|Error|assert " $O/t9_unit.log | head -20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_cf|fprop" --csv --log-file $O/cf_micro_launches.csv python tools/cf_microbench.py > $O/cf_micro.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad" --csv --log-file $O/wg_micro_launches.csv python tools/wgrad_microbench.py > $O/wg_micro.log 2>&1
for cfg in "0 0" "1 0" "0 1" "1 1"; do set -- $cfg
  RSS_CONV_CF=$1 RSS_WGRAD_TC=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s2i_cf$1_tc$2.json 2>> $O/bench_s2i.err; echo "cf=$1 tc=$2: $(cut -c60-130 $O/bench_s2i_cf$1_tc$2.json)"
done
RSS_CONV_CF=1 RSS_WGRAD_TC=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_s2i.log; grep -E "passed|failed|FAILED" $O/pytest_s2i.log | head
RSS_CONV_CF=1 RSS_WGRAD_TC=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2i.csv python tools/profile_step.py > $O/prof_s2i.log 2>&1; tail -2 $O/prof_s2i.log
