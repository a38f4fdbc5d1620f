"""stand-in for timm: only models.layers.to_2tuple / trunc_normal_ (multihead_isa_attention.py:12)"""
from . import models     # noqa: F401
