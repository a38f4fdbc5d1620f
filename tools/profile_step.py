"""One training step between cudaProfilerStart/Stop, for `ncu --profile-from-start off ...` (see profiles/README.md).
Numbers printed by a run under ncu are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from oracle import rssformer_ref as R  # noqa: E402  (deterministic synthetic weights / batch only)

B = int(os.environ.get("RSS_B", "16"))
S = int(os.environ.get("RSS_S", "512"))
model = P.build_rssformer(compute_dtype=torch.bfloat16)
model.load_state_dict(R.synth_state_dict(2333))
model.train()
opt = P.FlatSGD(model)
img, lbl = R.synth_batch(B, S)
img, lbl = img.cuda(), lbl.cuda()
for _ in range(3):
    P.train_step(model, opt, img, lbl)
torch.cuda.synchronize()
torch.cuda.profiler.start()
P.train_step(model, opt, img, lbl)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step; C-ABI kernel launches so far:", P.ops.COUNTERS["launches"])
