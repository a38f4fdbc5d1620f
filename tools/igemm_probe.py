"""Launches the tcgen05 igemm conv on a few RSSFormer layer shapes (for `ncu -k regex:conv_igemm --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from representationlearning_b200 import conv  # noqa: E402

B = 16
SHAPES = [(128, 32, 32, [(3, 1)]), (64, 64, 64, [(3, 1)]), (128, 64, 64, [(3, 1)]), (128, 128, 128, [(1, 1), (3, 6), (3, 12)])]
conv.ENGINE.update(igemm=True, igemm_single=True)
for H, Cin, Cout, srcs in SHAPES:
    x = torch.randn(B, Cin, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    ws = [torch.nn.Parameter(torch.randn(Cout, Cin, k, k, device="cuda") * 0.05) for k, d in srcs]
    with torch.no_grad():
        for _ in range(3):
            y = conv.conv_sum(x, [(w, None, k, d) for w, (k, d) in zip(ws, srcs)])
    torch.cuda.synchronize()
print("done")
