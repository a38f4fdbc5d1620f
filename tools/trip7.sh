#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "conv_cf or wgrad" 2>&1 | tail -40 > $O/t7_unit.log
grep -E "passed|failed|FAILED|Error|assert " $O/t7_unit.log | head -20
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
for k,v in d.items():
    if k.startswith('wgrad_tc'): print(k, v)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_cf|fprop" --csv --log-file $O/cf_micro_launches.csv python tools/cf_microbench.py > $O/cf_micro.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"wgrad" --csv --log-file $O/wg_micro_launches.csv python tools/wgrad_microbench.py > $O/wg_micro.log 2>&1
RSS_CONV_CF=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_s2g.log; grep -E "passed|failed|FAILED" $O/pytest_s2g.log | head
RSS_CONV_CF=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s2g_cf.json 2> $O/bench_s2g.err; cut -c1-200 $O/bench_s2g_cf.json
RSS_CONV_CF=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s2g_nocf.json 2>> $O/bench_s2g.err; cut -c1-200 $O/bench_s2g_nocf.json
RSS_CONV_CF=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2g.csv python tools/profile_step.py > $O/prof_s2g.log 2>&1; tail -2 $O/prof_s2g.log
