"""Per-shape timing of the fused tcgen05 conv (csrc/conv_cf.cu, with the BN-statistics epilogue) vs ATen/cuDNN conv (+ the separate
statistics kernel it replaces), forward and data gradient, B=16 RSSFormer layer shapes.  Writes gpurun_out/cf_microbench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from representationlearning_b200 import conv  # noqa: E402

B = 16
SHAPES = [  # (H, Cin, Cout, k, tag)
    (128, 32, 32, 3, "branch0 3x3"), (64, 64, 64, 3, "branch1 3x3"), (128, 64, 64, 3, "layer1 3x3"),
    (128, 64, 64, 1, "layer1 1x1 64->64"), (64, 64, 32, 1, "fuse 1x1 64->32"), (32, 128, 32, 1, "fuse 1x1 128->32"),
    (32, 128, 128, 3, "branch2 3x3 (dgrad only shape)"),
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


lib = P._lib.load()
out = []
for H, Cin, Cout, k, tag in SHAPES:
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    wl = w.bfloat16().contiguous(memory_format=torch.channels_last)
    stats = Cout in (32, 64)
    bn = P.FusedBNAct(Cout, 1).cuda().train()
    packed, _, nt, tdy, tdx, keep = conv._pack([w], [None], [k], [1], Cout, Cin, False, x.device)
    st = bn.stats_args() if stats else None
    r = dict(tag=tag, H=H, Cin=Cin, Cout=Cout, k=k, mbytes=(x.numel() + B * Cout * H * H) * 2 / 1e6)
    if lib.rss_conv_cf_supported(B, H, H, Cin, Cout, k, int(stats)):
        r["cf_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, st))
        r["cf_gbs"] = r["mbytes"] / r["cf_us"] * 1e3
        aff = torch.zeros(4, Cin, device="cuda"); aff[2] = 1.0
        r["cf_xform_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, aff, True, st))
    y = torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1)
    r["lib_us"] = timeit(lambda: torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1))
    if stats:
        r["lib_plus_stats_us"] = timeit(lambda: bn(torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1)))
    out.append(r)
    print(" | ".join("%s=%s" % (kk, ("%.1f" % v) if isinstance(v, float) else v) for kk, v in r.items()))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cf_microbench.json", "w"), indent=1)
