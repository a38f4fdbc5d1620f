"""Per-kernel time of the fused stem kernels (csrc/stem.cu) at the benched geometry next to the library path they replace
(cast + permute + ATen convolution / convolution_backward): CUPTI kernel durations (torch.profiler), L2 flushed between calls.
One JSON line per variant on stdout."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from representationlearning_b200 import ops  # noqa: E402

B, S = int(os.environ.get("RSS_B", "16")), int(os.environ.get("RSS_S", "512"))
dev = "cuda"
x = torch.randn(B, 3, S, S, device=dev)
w = (torch.randn(64, 3, 3, 3, device=dev) * 0.2).requires_grad_(True)
scratch = torch.zeros(130, device=dev)
shift = torch.zeros(64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

y = ops.StemConv.apply(x, w, None)
dy = torch.randn_like(y)
out_bytes = y.numel() * 2
in_bytes = x.numel() * 4


def fwd_plain():
    ops.StemConv.apply(x, w, None)


def fwd_stats():
    ops.StemConv.apply(x, w, (scratch, shift))


def wgrad():
    w.grad = None
    yy = ops.StemConv.apply(x, w, None)
    yy.backward(dy)


def lib_fwd():
    xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    torch.nn.functional.conv2d(xb, w.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last), stride=2, padding=1)


xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
wb = w.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def lib_wgrad():
    torch.ops.aten.convolution_backward(dy, xb, wb, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False])


# event brackets include the host latency of the first launch (~30 us of Python per call): the per-kernel numbers below come from
# CUPTI (torch.profiler), which times the kernels themselves
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def kernels(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
    out = {}
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)
        if t and e.count and "fill" not in e.key.lower() and "memset" not in e.key.lower():
            out[e.key[:70]] = round(t / e.count, 1)
    return out


alg = in_bytes + out_bytes
for name, fn in (("stem_conv_fwd", fwd_plain), ("stem_conv_fwd + bn raw sums", fwd_stats), ("stem fwd + wgrad", wgrad),
                 ("library forward: cast + permute + conv", lib_fwd), ("library wgrad", lib_wgrad)):
    ks = kernels(fn)
    print(json.dumps({"variant": name, "kernel_us": ks, "sum_us": round(sum(ks.values()), 1), "algorithmic_MB_per_kernel": round(alg / 1e6, 1),
                      "B": B, "S": S}))
