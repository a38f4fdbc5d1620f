"""conv_cf attribution run: branch-0 / branch-1 shapes, 10 launches each (use under ncu --metrics gpu__time_duration.sum)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import representationlearning_b200 as P
from representationlearning_b200 import conv
B = 16
for H, C in ((128, 32), (64, 64)):
    x = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w = torch.randn(C, C, 3, 3, device="cuda") * 0.05
    bn = P.FusedBNAct(C, 1).cuda().train()
    packed, _, nt, tdy, tdx, keep = conv._pack([w], [None], [3], [1], C, C, False, x.device)
    for _ in range(10):
        conv._cf_launch(x, packed, nt, tdy, tdx, C, C, None, False, bn.stats_args())
    torch.cuda.synchronize()
