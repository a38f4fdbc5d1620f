"""Per-role clock64() timeline of CTA 0 of conv_cf_kernel (RSS_CF_TRACE_PTR) + event timings of the fused-conv variants.
Writes gpurun_out/cf_trace.json.  Profiling aid, not a bench."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from representationlearning_b200 import conv  # noqa: E402

B = int(os.environ.get("B", "16"))
out = {}


def timeit(fn, reps=30):
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


for H, C, k in ((128, 32, 3), (64, 64, 3), (128, 64, 3), (128, 32, 1)):
    x = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(C, C, k, k, device="cuda") * 0.05
    wl = w.bfloat16().contiguous(memory_format=torch.channels_last)
    bn = P.FusedBNAct(C, 1).cuda().train()
    packed, _, nt, tdy, tdx, keep = conv._pack([w], [None], [k], [1], C, C, False, x.device)
    aff = torch.zeros(4, C, device="cuda"); aff[2] = 1.0
    tag = "H%d_C%d_k%d" % (H, C, k)
    r = {}
    for name, in_aff, st in (("plain", None, None), ("stats", None, bn.stats_args()), ("xform_stats", aff, bn.stats_args())):
        for dbg in ("0", "4", "7"):
            os.environ["RSS_CF_DBG"] = dbg
            trace = torch.zeros(4 * 16 * 8, dtype=torch.int64, device="cuda")
            os.environ.pop("RSS_CF_TRACE_PTR", None)
            us = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, C, C, in_aff, in_aff is not None, st))
            os.environ["RSS_CF_TRACE_PTR"] = hex(trace.data_ptr())
            conv._cf_launch(x, packed, nt, tdy, tdx, C, C, in_aff, in_aff is not None, st)
            torch.cuda.synchronize()
            os.environ.pop("RSS_CF_TRACE_PTR", None)
            t = trace.cpu().view(4, 16, 8)
            t0 = int(t[t > 0].min()) if (t > 0).any() else 0
            rel = torch.where(t > 0, t - t0, torch.full_like(t, -1))
            r["%s_dbg%s" % (name, dbg)] = dict(us=us, trace=rel.tolist())
            print(tag, name, "dbg", dbg, "%.1f us" % us, flush=True)
    os.environ["RSS_CF_DBG"] = "0"
    r["lib_us"] = timeit(lambda: torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1))
    r["lib_plus_bn_us"] = timeit(lambda: bn(torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1)))
    print(tag, "lib %.1f us, lib+bn %.1f us" % (r["lib_us"], r["lib_plus_bn_us"]), flush=True)
    if C == 32 and k == 3:
        os.environ["RSS_CF_MMA"] = "1"
        try:
            y0 = torch.ops.aten.convolution(x, wl, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1)
            y1, _ = conv._cf_launch(x, packed, nt, tdy, tdx, C, C, None, False, None)
            torch.cuda.synchronize()
            r["c32_err"] = float((y1.float() - y0.float()).abs().max() / y0.float().abs().max())
            r["c32_plain_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, C, C, None, False, None))
            r["c32_xform_stats_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, C, C, aff, True, bn.stats_args()))
            print(tag, "c32 err %.2e plain %.1f us xform+stats %.1f us" % (r["c32_err"], r["c32_plain_us"], r["c32_xform_stats_us"]), flush=True)
        except Exception as e:  # noqa: BLE001
            r["c32_error"] = repr(e)
            print("c32 failed", e, flush=True)
        os.environ.pop("RSS_CF_MMA", None)
    out[tag] = r
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cf_trace.json", "w"))
