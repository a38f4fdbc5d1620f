#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest_s2h.log; grep -E "passed|failed|FAILED" $O/pytest_s2h.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_s2h.json 2> $O/bench_s2h.err; cut -c1-200 $O/bench_s2h.json
RSS_BN_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s2h_nofuse.json 2>> $O/bench_s2h.err; cut -c1-200 $O/bench_s2h_nofuse.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s2h.csv python tools/profile_step.py > $O/prof_s2h.log 2>&1; tail -2 $O/prof_s2h.log
