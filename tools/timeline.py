"""Kernel timeline of graph-replayed training steps (CUPTI through torch.profiler; nsys is not in the image).

Writes gpurun_out/timeline_<tag>.csv: step-relative start (us), duration (us), stream, kernel name for every kernel of
ONE replayed step, plus a per-stream summary on stdout.  Used to find what is on the critical path of the multi-stream
step (branch streams + weight-gradient stream); numbers under the profiler are not bench values."""
import gzip
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from oracle import rssformer_ref as R  # noqa: E402  (deterministic synthetic weights / batch only)

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
B = int(os.environ.get("RSS_B", "16"))
S = int(os.environ.get("RSS_S", "512"))
model = P.build_rssformer(compute_dtype=torch.bfloat16)
model.load_state_dict(R.synth_state_dict(2333))
model.train()
opt = P.FlatSGD(model)
img, lbl = R.synth_batch(B, S)
img, lbl = img.cuda(), lbl.cuda()
g = P.GraphedTrainStep(model, opt, img, lbl, warmup=2)
for _ in range(3):
    g()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        g()
    torch.cuda.synchronize()
out = os.path.join("gpurun_out", "timeline_%s.json" % tag)
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
os.remove(out)
ev.sort(key=lambda e: e["ts"])
# the second replay starts at the second optimiser kernel's successor: split after the FIRST sgd_step kernel
cut = next(i for i, e in enumerate(ev) if "sgd_step" in e["name"]) + 1
while cut < len(ev) and "sgd_step" not in ev[cut]["name"] and ev[cut]["dur"] < 4 and ev[cut]["ts"] - ev[cut - 1]["ts"] < 20:
    cut += 1                                   # the few tiny kernels that trail the optimiser (BN counters)
step = ev[cut:]
t0 = step[0]["ts"]
with open(os.path.join("gpurun_out", "timeline_%s.csv" % tag), "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in step:
        f.write("%.2f,%.2f,%s,%s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream", -1), e["name"].replace(",", ";")[:90]))
end = max(e["ts"] + e["dur"] for e in step) - t0
print("kernels in step: %d   span %.1f us   sum of durations %.1f us" % (len(step), end, sum(e["dur"] for e in step)))
per = {}
for e in step:
    s = e["args"].get("stream", -1)
    a = per.setdefault(s, [0, 0.0])
    a[0] += 1; a[1] += e["dur"]
for s, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print("stream %s: %5d kernels, busy %.1f us" % (s, n, t))
