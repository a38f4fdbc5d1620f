#!/bin/bash
# GPU trip 3: validate conv_cf (both descriptor conventions), full tests, microbenches, bench, ncu of wgrad/cf
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "conv_cf" 2>&1 | tail -30 > $O/cf_default.log
tail -4 $O/cf_default.log
if ! grep -q " passed" $O/cf_default.log || grep -q "failed" $O/cf_default.log; then
  RSS_CF_DESC_SWAP=1 timeout 300 python -m pytest tests -m gpu -x -q -k "conv_cf" 2>&1 | tail -30 > $O/cf_swap.log
  echo "--- swapped:"; tail -4 $O/cf_swap.log
fi
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_s2c.log; tail -6 $O/pytest_s2c.log
timeout 300 python tools/cf_microbench.py 2>&1 | tail -9
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_s2c.json 2> $O/bench_s2c.err; cut -c1-330 $O/bench_s2c.json
timeout 600 ncu --set full --clock-control none -k regex:"conv_wgrad_kernel|conv_cf_kernel" -c 14 -o $O/wg_cf_probe -f python tools/wgrad_microbench.py > $O/wg_probe.log 2>&1
ncu -i $O/wg_cf_probe.ncu-rep --page raw --csv > $O/wg_cf_probe_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2c.csv python tools/profile_step.py > $O/prof_s2c.log 2>&1; tail -2 $O/prof_s2c.log
