#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "conv_cf" 2>&1 | tail -40 > $O/cf_default.log
grep -E "passed|failed|FAILED|Error|assert " $O/cf_default.log | head -20
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
for k,v in d.items():
    if k.startswith('cf_'): print(k, {a: float('%.3g' % b) for a, b in v.items()})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_cf|fprop|bn_stats|bn_act" --csv --log-file $O/cf_micro_launches.csv python tools/cf_microbench.py > $O/cf_micro.log 2>&1
RSS_CONV_CF=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_s2f.log; grep -E "passed|failed|FAILED" $O/pytest_s2f.log | head
RSS_CONV_CF=1 timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_s2f.json 2> $O/bench_s2f.err; cut -c1-330 $O/bench_s2f.json
timeout 600 ncu --set full --clock-control none -k regex:conv_cf_kernel -c 6 -o $O/cf_probe2 -f python tools/cf_microbench.py > $O/cf_probe2.log 2>&1
ncu -i $O/cf_probe2.ncu-rep --page raw --csv > $O/cf_probe2_raw.csv 2>/dev/null
ncu -i $O/cf_probe2.ncu-rep --page source --csv --kernel-name regex:conv_cf --launch-skip 0 --launch-count 1 > $O/cf_probe2_src.csv 2>/dev/null
RSS_CONV_CF=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2f.csv python tools/profile_step.py > $O/prof_s2f.log 2>&1; tail -2 $O/prof_s2f.log
