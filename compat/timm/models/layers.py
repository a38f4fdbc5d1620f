import torch.nn as nn


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


trunc_normal_ = nn.init.trunc_normal_
DropPath = nn.Identity
