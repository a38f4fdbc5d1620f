#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "loss or head or model or step or traj" 2>&1 | tail -30 > $O/pytest_s3d.log; grep -E "passed|failed|FAILED|Error" $O/pytest_s3d.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_s3d.json 2> $O/bench_s3d.err; echo "default: $(cut -c60-130 $O/bench_s3d.json)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_s3d.csv python tools/profile_step.py > $O/prof_s3d.log 2>&1; tail -1 $O/prof_s3d.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 9 -c 2 -o $O/igemm_s3d -f python tools/igemm_probe.py > $O/igemm_s3d.log 2>&1
ncu -i $O/igemm_s3d.ncu-rep --page raw --csv > $O/igemm_s3d_raw.csv 2>/dev/null
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"win_attn_bwd_tc|win_attn_fwd_tc|bn_bwd_reduce_kernel|bn_stats_fused|seg_loss_fwd" -c 12 -o $O/step_s3d -f python tools/profile_step.py > $O/step_s3d.log 2>&1
ncu -i $O/step_s3d.ncu-rep --page raw --csv > $O/step_s3d_raw.csv 2>/dev/null
ls -la $O/*.ncu-rep | tail -3
