"""Event timings of the fused tcgen05 conv (csrc/conv_cf.cu) on the benched layer geometries next to the library conv (+ the
separate BatchNorm kernels it replaces).  L2 is flushed between repetitions.  Writes gpurun_out/cf_bench.json.
Profiling aid, not a bench:   python tools/cf_bench.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from representationlearning_b200 import conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


NCU = os.environ.get("CF_BENCH_NCU", "0") != "0"      # under ncu: one launch per variant, durations come from the profiler
N_IN_GRAPH = 8


def timeit(fn, reps=10):
    """per-launch time of `fn` replayed from a CUDA graph of N_IN_GRAPH back-to-back launches (host launch latency would otherwise
    dominate a 10-us kernel); L2 flushed before each replay, so the first launch of a replay reads HBM and the rest see what the
    previous launch left in L2 -- close to the step, where the producer has just written the input"""
    if NCU:
        fn()
        torch.cuda.synchronize()
        return 0.0
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N_IN_GRAPH):
                fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / N_IN_GRAPH)
    ts.sort()
    return ts[len(ts) // 2]


CL = torch.channels_last
NSHAPES = int(os.environ.get("CF_BENCH_SHAPES", "99"))
for H, Cin, Cout, k in ((128, 32, 32, 3), (64, 64, 64, 3), (128, 64, 64, 1), (128, 32, 128, 1), (128, 128, 32, 1),
                        (64, 64, 32, 1))[:NSHAPES]:
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=CL)
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    wl = w.bfloat16().contiguous(memory_format=CL)
    bn = P.FusedBNAct(Cout, 1).cuda().train()
    packed, _, nt, tdy, tdx, keep = conv._pack([w], [None], [k], [1], Cout, Cin, False, x.device)
    aff_in = torch.zeros(4, Cin, device="cuda"); aff_in[2] = 1.0
    tag = "H%d_%dto%d_k%d" % (H, Cin, Cout, k)
    r = {"MB": (x.numel() + B * Cout * H * H) * 2 / 1e6}
    lib = P._lib.load()
    r["plain_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, None))
    r["xform_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, aff_in, True, None))
    if lib.rss_conv_cf_supported(B, H, H, Cin, Cout, k, 1):
        r["stats_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, bn.stats_args()))
        r["xform_stats_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, aff_in, True, bn.stats_args()))
    if lib.rss_conv_cf_supported(B, H, H, Cin, Cout, k, 2):
        z = torch.randn(B, Cout, H, H, device="cuda").bfloat16().contiguous(memory_format=CL)
        o = torch.relu(z)
        aff = torch.zeros(4, Cout, device="cuda"); aff[1] = 1.0; aff[2] = 1.0
        r["bnred_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, None,
                                                       bnred=(z, None, aff, True, bn._scratch)))
        r["bnred_out_add_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, None, add=z,
                                                               bnred=(z, o, aff, True, bn._scratch)))
        r["add_us"] = timeit(lambda: conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, None, add=z))
    r["lib_us"] = timeit(lambda: torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1))
    r["lib_plus_bn_us"] = timeit(lambda: bn(torch.ops.aten.convolution(x, wl, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1)))
    r["copy_us"] = timeit(lambda: x.clone())
    print(tag, " ".join("%s=%.1f" % (a, b) for a, b in r.items()), flush=True)
    out[tag] = r
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cf_bench.json", "w"), indent=1)
