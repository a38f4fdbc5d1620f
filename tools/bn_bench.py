"""Per-call times of the BatchNorm(+ReLU)(+residual) forward and backward on the HRNet branch geometries: split protocol (bn.cu:
statistics + apply kernels) vs one-launch cluster kernels (bn_cluster.cu).  Graph-replayed launches, L2 flushed per replay.
Profiling aid, not a bench:   python tools/bn_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from representationlearning_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
N = 8


def timeit(fn, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / N)
    ts.sort()
    return ts[len(ts) // 2]


out = {}
CL = torch.channels_last
for B, C, H in ((16, 64, 64), (16, 128, 32), (16, 256, 16), (16, 64, 32), (16, 128, 16)):
    for res in (False, True):
        x = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=CL)
        r = torch.randn_like(x) if res else None
        dy = torch.randn_like(x)
        bn = P.FusedBNAct(C, 1).cuda().train()
        for p in bn.parameters():
            p.requires_grad_(False)
        row = {"MB": x.numel() * 2 / 1e6}
        for proto in ("split", "cluster"):
            ops.BN_CLUSTER["on"] = proto == "cluster"
            xi = x.clone().requires_grad_(True)
            ri = r.clone().requires_grad_(True) if res else None
            row[proto + "_fwd_us"] = timeit(lambda: bn(x, r))

            def both():
                y = bn(xi, ri)
                torch.autograd.grad(y, [xi] + ([ri] if res else []), dy)
            row[proto + "_bwd_us"] = timeit(both) - row[proto + "_fwd_us"]
        tag = "B%d_C%d_H%d%s" % (B, C, H, "_res" if res else "")
        out[tag] = row
        print(tag, " ".join("%s=%.1f" % kv for kv in row.items()), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bn_bench.json", "w"), indent=1)
