"""Per-shape timing of the hand-written weight-gradient kernel (csrc/conv_wgrad.cu) vs ATen/cuDNN's wgrad, B=16 RSSFormer
layer shapes.  Writes gpurun_out/wgrad_microbench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402

B = 16
SHAPES = [  # (Hi, Cin, Cout, k, stride, tag)
    (128, 32, 32, 3, 1, "branch0 3x3"), (64, 64, 64, 3, 1, "branch1 3x3"), (32, 128, 128, 3, 1, "branch2 3x3"),
    (16, 256, 256, 3, 1, "branch3 3x3"), (128, 64, 64, 3, 1, "layer1 3x3"), (128, 256, 32, 3, 1, "transition1.0"),
    (256, 64, 64, 3, 2, "stem conv2 s2"), (128, 32, 64, 3, 2, "fuse down 32->64 s2"), (64, 64, 128, 3, 2, "fuse down 64->128 s2"),
    (128, 32, 128, 1, 1, "ffn fc1"), (128, 128, 32, 1, 1, "ffn fc2"), (64, 64, 32, 1, 1, "fuse 1x1 64->32"),
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
for Hi, Cin, Cout, k, s, tag in SHAPES:
    pad = k // 2
    Ho = (Hi + 2 * pad - k) // s + 1
    x = torch.randn(B, Cin, Hi, Hi, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(B, Cout, Ho, Ho, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    dw = torch.zeros(Cout, Cin, k, k, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def own():
        rc = lib.rss_conv_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, Hi, Hi, Cin, Ho, Ho, Cout, k, s, pad, 1, st)
        assert rc == 0, rc

    def ref():
        g = torch.ops.aten.convolution_backward(dy, x, w, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [False, True, False])[1]
        dw.add_(g)

    def own_tc():
        rc = lib.rss_conv_wgrad_tc(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, Hi, Hi, Cin, Cout, k, None, None, 0, st)
        assert rc == 0, rc

    r = dict(tag=tag, Hi=Hi, Cin=Cin, Cout=Cout, k=k, stride=s, gflop=2.0 * B * Ho * Ho * Cin * Cout * k * k / 1e9,
             mbytes=(x.numel() + dy.numel()) * 2 / 1e6, own_us=timeit(own), lib_us=timeit(ref))
    if s == 1 and lib.rss_conv_wgrad_tc_supported(B, Hi, Hi, Cin, Cout, k):
        r["tc_us"] = timeit(own_tc)
    r["own_gbs"] = r["mbytes"] / r["own_us"] * 1e3 / 1e3
    out.append(r)
    print("%-22s Hi=%3d %3d->%3d k%d s%d | mma.sync %7.1f us  tcgen05 %7.1f us  lib+add %7.1f us  (CPU-launch-bound below ~60 us)" %
          (tag, Hi, Cin, Cout, k, s, r["own_us"], r.get("tc_us", float("nan")), r["lib_us"]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/wgrad_microbench.json", "w"), indent=1)
