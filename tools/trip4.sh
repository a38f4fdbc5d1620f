#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "conv_cf" 2>&1 | tail -40 > $O/cf_default.log
grep -E "passed|failed|FAILED|Error|assert " $O/cf_default.log | head -20
if grep -q "failed" $O/cf_default.log; then
  RSS_CF_DESC_SWAP=1 timeout 300 python -m pytest tests -m gpu -q -k "conv_cf" 2>&1 | tail -40 > $O/cf_swap.log
  echo "--- swapped:"; grep -E "passed|failed|FAILED|Error|assert " $O/cf_swap.log | head -20
fi
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_report.json'))
for k,v in d.items():
    if k.startswith('cf_'): print(k, v)
PY
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_s2d.log; grep -E "passed|failed|FAILED" $O/pytest_s2d.log | head
timeout 300 python tools/cf_microbench.py 2>&1 | tail -9
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_s2d.json 2> $O/bench_s2d.err; cut -c1-330 $O/bench_s2d.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2d.csv python tools/profile_step.py > $O/prof_s2d.log 2>&1; tail -2 $O/prof_s2d.log
