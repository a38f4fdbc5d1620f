#!/bin/bash
# Final evidence run of a round on ONE B200 (gpurun): GPU test suite, the full bench line, a CUPTI timeline of the replayed step and
# `ncu --set full` pages of the kernels written in the second half of round 2.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 240 python -m pytest tests -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/final_gpu_tests.log
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d.get("e2e"), d.get("roofline"), d.get("clocks"))
PY
timeout 120 python tools/timeline.py final3 > gpurun_out/timeline_final3.log 2>&1; echo "timeline rc=$?"
python tools/timeline_summary.py gpurun_out/timeline_final3.csv > gpurun_out/timeline_final3_summary.txt 2>&1; head -4 gpurun_out/timeline_final3_summary.txt
timeout 200 ncu --profile-from-start off --set full --import-source on --clock-control none \
  -k regex:"stem_conv|neck_gather_bwd|seg_loss_fwd|accum_list|shadow_cl" -c 8 -f -o gpurun_out/ncu_r2b python tools/profile_step.py > gpurun_out/ncu_r2b.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/ncu_r2b.ncu-rep --page raw --csv > gpurun_out/ncu_r2b_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/ncu_r2b_raw.csv > gpurun_out/ncu_r2b_summary.txt 2>&1; cat gpurun_out/ncu_r2b_summary.txt | cut -c1-170
rm -f gpurun_out/ncu_r2b.ncu-rep
python tools/stem_bench.py > gpurun_out/stem_bench.jsonl 2>/dev/null; echo "stem bench rc=$?"
