"""TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs /root/reference):  python -m oracle.gen_bench_loss

Step-0 training loss of the UNMODIFIED reference (fp32, CPU) on exactly the batches bench.py feeds rank r of the BASELINE cfg2 /
cfg3 run (16 tiles of 512x512, images seed 7+r, labels seed 1+r, weights seed 2333) -> tests/golden/bench_cfg2_loss.json.
bench.py asserts the loss of its first step against these numbers (a benched step that computes something else is not a bench)."""
import json
import os

import torch

from oracle import rssformer_ref as R
from oracle.ref_shim import build_reference_model

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_cfg2_loss.json")


def main(ranks=8, B=16, S=512):
    sd = R.synth_state_dict(2333)
    model = build_reference_model()
    out = {"batch": B, "size": S, "weights_seed": 2333, "loss_fp32_step0": {}}
    for r in range(ranks):
        model.load_state_dict(sd)
        model.train()
        img, lbl = R.synth_batch(B, S, seed_img=7 + r, seed_lbl=1 + r)
        with torch.no_grad():
            loss = sum(model(img, {"cls": lbl}).values())
        out["loss_fp32_step0"][str(r)] = float(loss.item())
        print(r, out["loss_fp32_step0"][str(r)], flush=True)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
