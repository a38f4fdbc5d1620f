#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/pytest_final.log; grep -E "passed|failed|FAILED|Error" $O/pytest_final.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; tail -2 $O/smoke_final.log
timeout 600 python bench.py > $O/bench_final.json 2> $O/bench_final.err; cut -c1-160 $O/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_final_ref.json 2>> $O/bench_final.err; cut -c1-200 $O/bench_final_ref.json
