"""Per-shape timing of the tcgen05 igemm conv vs ATen/cuDNN (fwd and dgrad), B=16 RSSFormer layer shapes.
Writes gpurun_out/conv_microbench.json; used to choose the per-shape engine policy in conv.py."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from representationlearning_b200 import conv, ops  # noqa: E402

B = 16
SHAPES = [  # (H, Cin, Cout, [(k, dil)...], tag)
    (128, 32, 32, [(3, 1)], "branch0 3x3"), (64, 64, 64, [(3, 1)], "branch1 3x3"), (32, 128, 128, [(3, 1)], "branch2 3x3"),
    (16, 256, 256, [(3, 1)], "branch3 3x3"), (128, 128, 128, [(1, 1), (3, 6), (3, 12)], "ffn conv19"),
    (128, 32, 128, [(1, 1)], "ffn fc1"), (128, 128, 32, [(1, 1)], "ffn fc2"), (128, 480, 480, [(1, 1)], "neck 1x1"),
    (128, 64, 64, [(3, 1)], "layer1 3x3"), (128, 256, 64, [(1, 1)], "layer1 1x1 a"), (128, 64, 256, [(1, 1)], "layer1 1x1 b"),
    (64, 64, 32, [(1, 1)], "fuse 1x1 64->32"), (32, 128, 32, [(1, 1)], "fuse 1x1 128->32"), (16, 256, 32, [(1, 1)], "fuse 1x1 256->32"),
    (128, 256, 32, [(3, 1)], "transition1.0"),
]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = []
for H, Cin, Cout, srcs, tag in SHAPES:
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ws = [torch.nn.Parameter(torch.randn(Cout, Cin, k, k, device="cuda") * 0.05) for k, d in srcs]
    dy = torch.randn(B, Cout, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    convs = [(w, None, k, d) for w, (k, d) in zip(ws, srcs)]
    res = dict(tag=tag, H=H, Cin=Cin, Cout=Cout, taps=sum(k * k for k, d in srcs),
               gflop=2.0 * B * H * H * Cin * Cout * sum(k * k for k, d in srcs) / 1e9)
    for eng in ("igemm", "lib"):
        conv.ENGINE["igemm"] = eng == "igemm"
        with torch.no_grad():
            res[eng + "_fwd_us"] = timeit(lambda: conv.conv_sum(x, convs))
        y = conv.conv_sum(x, convs)
        for w in ws:
            w.requires_grad_(False)      # time the data gradient alone
        y = conv.conv_sum(x, convs)
        res[eng + "_dgrad_us"] = timeit(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
        for w in ws:
            w.requires_grad_(True)
    res["igemm_fwd_tflops"] = res["gflop"] / res["igemm_fwd_us"] / 1e3
    out.append(res)
    print("%-20s H=%3d %3d->%3d taps %2d | fwd igemm %7.1f us  lib %7.1f us | dgrad igemm %7.1f  lib %7.1f | igemm %.0f TF"
          % (tag, H, Cin, Cout, res["taps"], res["igemm_fwd_us"], res["lib_fwd_us"], res["igemm_dgrad_us"], res["lib_dgrad_us"],
             res["igemm_fwd_tflops"]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/conv_microbench.json", "w"), indent=1)
