"""Per-role clock64() timeline of CTA 0 of conv_cf_kernel (RSS_CF_TRACE_PTR): rows = first 16 tiles of that CTA, columns = the
events marked CF_TRACE in csrc/conv_cf.cu (cycles since the first event).  Writes gpurun_out/cf_trace.json.  Profiling aid."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P  # noqa: E402
from representationlearning_b200 import conv  # noqa: E402

B = int(os.environ.get("B", "16"))
CL = torch.channels_last
out = {}
for H, Cin, Cout, k in ((128, 32, 32, 3), (64, 64, 64, 3), (128, 64, 64, 1)):
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=CL)
    w = torch.randn(Cout, Cin, k, k, device="cuda") * 0.05
    bn = P.FusedBNAct(Cout, 1).cuda().train()
    packed, _, nt, tdy, tdx, keep = conv._pack([w], [None], [k], [1], Cout, Cin, False, x.device)
    aff = torch.zeros(4, Cin, device="cuda"); aff[2] = 1.0
    z = torch.randn(B, Cout, H, H, device="cuda").bfloat16().contiguous(memory_format=CL)
    aff_o = torch.zeros(4, Cout, device="cuda"); aff_o[1] = 1.0; aff_o[2] = 1.0
    variants = {"plain": dict(in_aff=None, stats=None), "xform": dict(in_aff=aff, stats=None),
                "stats": dict(in_aff=None, stats=bn.stats_args())}
    if k == 3:
        variants["bnred"] = dict(in_aff=None, stats=None, bnred=(z, None, aff_o, True, bn._scratch))
    for name, v in variants.items():
        def run():
            return conv._cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, v["in_aff"], v["in_aff"] is not None, v["stats"], bnred=v.get("bnred"))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        trace = torch.zeros(4 * 16 * 8, dtype=torch.int64, device="cuda")
        os.environ["RSS_CF_TRACE_PTR"] = hex(trace.data_ptr())
        run()
        torch.cuda.synchronize()
        os.environ.pop("RSS_CF_TRACE_PTR", None)
        t = trace.cpu().view(4, 16, 8)
        t0 = int(t[t > 0].min()) if (t > 0).any() else 0
        rel = torch.where(t > 0, t - t0, torch.full_like(t, -1))
        tag = "H%d_%dto%d_k%d_%s" % (H, Cin, Cout, k, name)
        out[tag] = rel.tolist()
        print(tag)
        for role, rn in enumerate(("tma", "mma", "xform", "epi")):
            for i in range(16):
                row = rel[role][i].tolist()
                if any(a >= 0 for a in row):
                    print("  %-5s %2d %s" % (rn, i, row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cf_trace.json", "w"))
