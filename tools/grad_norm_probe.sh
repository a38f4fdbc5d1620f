#!/bin/bash
# total-gradient-norm deviation of the cfg2 bf16 graph step vs the reference's fp32 run, under a few switches (one process each)
run() { name=$1; shift; env "$@" timeout 120 python -m pytest tests/test_gpu_parity.py -q -k test_cfg2_bf16_graph_step > /dev/null 2>&1; python - <<PY
import json
r = json.load(open("gpurun_out/parity_report.json"))["cfg2_bf16_graph"]
print("$name", "grad_norm_rel", r["grad_norm_rel"], "loss_rel", r["loss_rel"], "probs", r["probs_max_abs"], {k: round(v, 3) for k, v in r.items() if k.startswith("update")})
PY
}
run default A=1
run link0 RSS_RES_LINK=0
run oldgrids RSS_BN_TICKET_BPSM=2 RSS_BN_APPLY_BPSM=8 RSS_BN_BIG_BPSM=8
run default2 A=1
