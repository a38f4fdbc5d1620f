"""Host-side mirror of the reference's nn.Module surface for the RSSFormer hot path.

Same class names, constructor arguments, forward signatures and state_dict keys as the reference
(SURVEY.md §8(b)); the arithmetic is done by the sm_100a kernels behind include/rss_b200.h.
Reference files are cited per class (paths relative to RSSFormer-TIP2023/module/baseline/).

Activations flow NHWC (torch channels_last) in bf16 or fp32; parameters stay fp32.
Convolutions go through `conv.conv2d` (hand-written tcgen05 implicit GEMM where enabled,
library convolution otherwise — see DESIGN.md §kernels for which layer uses which).
"""
import math

import torch
import torch.nn as nn

from . import _lib, ops
from .conv import conv2d, conv_sum, conv_sum_stats_ok, prepack_sum

BN_MOMENTUM = 0.1        # _hrnet_rssformer.py:27
CL = torch.channels_last


class FusedBNAct(nn.Module):
    """BatchNorm2d / SyncBatchNorm + activation (+ residual add) in one pass.
    Parameter/buffer names are nn.BatchNorm2d's, so checkpoints load both ways."""

    def __init__(self, num_features, act=_lib.ACT_NONE, momentum=BN_MOMENTUM, eps=1e-5, sync=False):
        super().__init__()
        self.num_features, self.act, self.momentum, self.eps, self.sync = num_features, act, momentum, eps, sync
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        # private scratch of the BN kernels ([0:2] barrier/ticket counters, [2:] per-channel accumulators; kernels leave it zeroed).
        # One per layer: layers may run concurrently on different streams.  Not part of the state_dict.
        self.register_buffer("_scratch", torch.zeros(2 + 2 * num_features), persistent=False)

    defer_counter = False      # set per INSTANCE by trainer.FlatSGD, which bumps the counters of the modules it owns with one foreach op per step

    def forward(self, x, residual=None, pre_bias=None, aff=None, link=None):
        """pre_bias (training mode only): bias of the conv that produced x, left out of x because BatchNorm cancels it.
        aff: statistics already produced by the conv kernel's epilogue (conv.conv_bn_stats)."""
        if self.training and not self.defer_counter:
            self.num_batches_tracked += 1
        return ops.BNAct.apply(x, residual, self.weight, self.bias, self.running_mean, self.running_var,
                               self.training, self.momentum, self.eps, self.act, True if self.sync else None, self._scratch,
                               pre_bias, aff, link)

    def stats_args(self):
        """what conv.conv_bn_stats needs to produce this layer's batch statistics in the conv epilogue (None: not applicable --
        eval mode, or SyncBN under a process group, where the statistics are exchanged between ranks)"""
        if not self.training or (self.sync and ops._world(True) > 1):
            return None
        return (self.weight, self.bias, self.running_mean, self.running_var, self.momentum, self.eps, self._scratch)

    def after_conv(self, x, weight, bias, **kw):
        """conv(x, weight, bias) -> this BN(+act).  In training mode the conv bias only shifts the batch mean, which the
        normalisation subtracts again: the bias add (one full pass over the conv output) is skipped and the bias is handed to
        the statistics kernel for the running mean; its gradient is identically zero."""
        if self.training and bias is not None:
            return self(conv2d(x, weight, None, **kw), pre_bias=bias)
        return self(conv2d(x, weight, bias, bias_grad=not self.training, **kw))

    def extra_repr(self):
        return "%d, act=%d, sync=%s" % (self.num_features, self.act, self.sync)


class SpatialAttention(nn.Module):
    """Parameter holder for modules/multihead_isa_pool_attention.py:101-115 (7x7 conv 2->1, no bias)."""

    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in (3, 7), "kernel size must be 3 or 7"
        if kernel_size != 7:
            raise NotImplementedError("the RSSFormer path only builds SpatialAttention(7) (pool:143-144)")
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=3, bias=False)

    def forward(self, x):
        """(B,32,H,W) -> (B,1,H,W) = sigmoid(conv1([mean_c(x), max_c(x)])), pool:110-115.  Inside the model this runs fused with
        the level soft-max and the window attention (ops.WindowAttention); called on its own it is forward-only."""
        if torch.is_grad_enabled() and (x.requires_grad or self.conv1.weight.requires_grad):
            with torch.no_grad():
                out = ops.spatial_attention(x, self.conv1.weight)
            return out.to(x.dtype)        # no graph: the differentiable path is the fused block (ops.WindowAttention)
        return ops.spatial_attention(x, self.conv1.weight).to(x.dtype)


class Mhca(nn.Module):
    """modules/DAL.py:676-1030.  forward(query,key,value) takes sequence-first (L, Bw, C) windows."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, bias=True, add_bias_kv=False, add_zero_attn=False,
                 kdim=None, vdim=None):
        super().__init__()
        if dropout != 0.0 or not bias or add_bias_kv or add_zero_attn or kdim not in (None, embed_dim) \
                or vdim not in (None, embed_dim):
            raise NotImplementedError("RSSFormer builds Mhca(embed_dim, num_heads, dropout=0.0) only (pool:140)")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)

    def proj_params(self):
        return (self.q_proj.weight, self.q_proj.bias, self.k_proj.weight, self.k_proj.bias,
                self.v_proj.weight, self.v_proj.bias, self.out_proj.weight, self.out_proj.bias)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=False, attn_mask=None, residual_attn=None):
        """DAL.py:726-735 / 873-1020 on sequence-first windows: query, key, value (L=49, Bw, C) -> (49, Bw, C).
        softmax(q k^T / sqrt(d)) v per head, times sigmoid(avg+max pool of q^T k) per channel, then out_proj -- the same window
        kernels as the fused block (differentiable), fed with the windows laid side by side as one 7 x 7*Bw image.  Only the call
        pattern RSSFormer uses is supported: key is value, no masks, no attention weights returned."""
        if key is not value or key_padding_mask is not None or attn_mask is not None or residual_attn is not None or need_weights:
            raise NotImplementedError("RSSFormer calls Mhca(x_windows, y_windows, y_windows) without masks (pool:185)")
        L, Bw, C = query.shape
        if L != 49 or key.shape != query.shape:
            raise NotImplementedError("7x7 windows only (L == 49), same shape for query and key")

        def as_image(t):            # token r*7+c of window w -> pixel (r, 7*w + c)
            return t.reshape(7, 7, Bw, C).permute(2, 0, 1, 3).reshape(1, Bw, 7, 7, C).permute(0, 4, 2, 1, 3).reshape(1, C, 7, Bw * 7)
        xi, yi = as_image(query), as_image(key)
        out = ops.WindowAttention.apply(xi, yi, 0.0, False, None, None, None, None, None, None, *self.proj_params())
        return out.reshape(C, 7, Bw, 7).permute(1, 3, 2, 0).reshape(49, Bw, C)


class InterlacedPoolAttention2(nn.Module):
    """modules/multihead_isa_pool_attention.py:117-188.  `rpe` is accepted and ignored exactly as in
    the reference (no relative-position parameter exists there either, pool:129-145)."""

    def __init__(self, embed_dim, num_heads, window_size=7, rpe=True, **kwargs):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.with_rpe = embed_dim, num_heads, window_size, rpe
        self.attn = Mhca(embed_dim, num_heads, **kwargs)
        self.atrous_block1 = SpatialAttention(7)
        self.atrous_block2 = SpatialAttention(7)
        self.weight_levels = nn.Conv2d(2, 2, kernel_size=1, stride=1, padding=0)

    def gate_params(self):
        return (self.atrous_block1.conv1.weight, self.atrous_block2.conv1.weight,
                self.weight_levels.weight, self.weight_levels.bias)

    def forward(self, x, y, H, W, **kwargs):
        """x, y: (B, N, C) token-major (already normalised) -> (B, N, C).  (pool:148-188)"""
        B, N, C = x.shape
        assert N == H * W
        xi = x.reshape(B, H, W, C).permute(0, 3, 1, 2)        # NHWC memory viewed as NCHW: zero-copy
        yi = y.reshape(B, H, W, C).permute(0, 3, 1, 2)
        out = ops.WindowAttention.apply(xi, yi, 0.0, False, None, None, *self.gate_params(), *self.attn.proj_params())
        return out.permute(0, 2, 3, 1).reshape(B, N, C)


class MlpDWBN(nn.Module):
    """modules/ffn_block.py:207-270 (token branch; the 4-D branch is dead code in the reference, :272-284).
    fc1(1x1) -> SyncBN -> GELU -> [dw(1x1) + dw6(3x3,d6) + dw12(3x3,d12)] -> SyncBN -> GELU -> fc2(1x1) -> SyncBN -> GELU."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, dw_act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU or dw_act_layer is not nn.GELU:
            raise NotImplementedError("RSSFormer builds MlpDWBN with nn.GELU (MTFM.py:92-99)")
        self.fc1 = nn.Conv2d(in_features, hidden_features, kernel_size=1)
        self.norm1 = FusedBNAct(hidden_features, _lib.ACT_GELU, sync=True)
        self.dw = nn.Conv2d(hidden_features, hidden_features, 1, 1)
        self.dw6 = nn.Conv2d(hidden_features, hidden_features, 3, 1, padding=6, dilation=6)
        self.dw12 = nn.Conv2d(hidden_features, hidden_features, 3, 1, padding=12, dilation=12)
        self.norm2 = FusedBNAct(hidden_features, _lib.ACT_GELU, sync=True)
        self.fc2 = nn.Conv2d(hidden_features, out_features, kernel_size=1)
        self.norm3 = FusedBNAct(out_features, _lib.ACT_GELU, sync=True)

    def _dw_convs(self):
        return [(self.dw.weight, self.dw.bias, 1, 1), (self.dw6.weight, self.dw6.bias, 3, 6), (self.dw12.weight, self.dw12.bias, 3, 12)]

    def prepack(self, x):
        """weight operands of the dw + dw6 + dw12 GEMM (and of its data gradient) packed on a side stream; x: the block input
        (B,C,H,W) -- only its batch / spatial shape, dtype and device matter.  None when the igemm kernel will not run."""
        B, _, H, W = x.shape
        like = torch.empty((B, 0, H, W), device=x.device, dtype=x.dtype)
        return prepack_sum(like, self._dw_convs(), want_bwd=self.training and torch.is_grad_enabled())

    def forward_nchw(self, x, prepacked=None):
        bg = not self.training       # every bias here feeds a training-mode BN: its gradient is identically zero
        x = self.norm1.after_conv(x, self.fc1.weight, self.fc1.bias)
        convs = self._dw_convs()
        n2 = self.norm2
        if conv_sum_stats_ok(x, n2.num_features) and ops.bn_accepts_raw_sums(x, n2.training, True if n2.sync else None, n2._scratch,
                                                                             n2.num_features):
            # norm2's statistics come out of the GEMM's epilogue: the 67 MB sum is written once and read once (by the apply pass)
            c = conv_sum(x, convs, bias_grad=bg, stats=(n2._scratch, n2.running_mean), prepacked=prepacked)
            x = n2(c, aff=ops.RAW_SUMS)
        else:
            x = n2(conv_sum(x, convs, bias_grad=bg, prepacked=prepacked))
        return self.norm3.after_conv(x, self.fc2.weight, self.fc2.bias)

    def forward(self, x, H, W):
        if x.dim() != 3:
            raise RuntimeError("Unsupported input shape: {}".format(x.shape))   # ffn_block.py:286-287
        B, N, C = x.shape
        if N != H * W:
            raise NotImplementedError("class-token inputs (N == H*W+1) are not on the RSSFormer path")
        y = self.forward_nchw(x.reshape(B, H, W, C).permute(0, 3, 1, 2))
        return y.permute(0, 2, 3, 1).reshape(B, N, C)


class GeneralTransformerBlock(nn.Module):
    """modules/MTFM.py:48-113.  forward(x, y): x=`low` (queries + residual stream), y=high-res branch (K/V only)."""
    expansion = 1

    def __init__(self, inplanes, planes, num_heads, window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=None):
        super().__init__()
        if drop_path != 0.0 or attn_drop != 0.0 or drop != 0.0:
            raise NotImplementedError("RSSFormer builds the block with all drop rates 0 (_hrnet_rssformer.py:308)")
        self.dim, self.out_dim, self.num_heads, self.window_size, self.mlp_ratio = inplanes, planes, num_heads, window_size, mlp_ratio
        self.attn = InterlacedPoolAttention2(self.dim, num_heads=num_heads, window_size=window_size, rpe=True, dropout=attn_drop)
        self.norm1 = nn.LayerNorm(self.dim, eps=1e-6)
        self.norm2 = nn.LayerNorm(self.out_dim, eps=1e-6)
        self.mlp = MlpDWBN(in_features=self.dim, hidden_features=int(self.dim * mlp_ratio), out_features=self.out_dim,
                           act_layer=act_layer, dw_act_layer=act_layer, drop=drop)

    def forward(self, x, y, mask=None, relu=False):
        """relu=True additionally applies the ReLU that HighResolutionModule puts on the block output
        (_hrnet_rssformer.py:435), fused with the residual add."""
        pre = self.mlp.prepack(x) if x.is_cuda else None          # the FFN GEMM's weight packs overlap the attention half
        # attention half: LN1(x), LN1(y), gate, window attention, + x  — one fused region
        t = ops.WindowAttention.apply(x, y, self.norm1.eps, True, self.norm1.weight, self.norm1.bias,
                                      *self.attn.gate_params(), *self.attn.attn.proj_params())
        # FFN half: LN2 -> MlpDWBN -> + t (-> ReLU)
        u = ops.LayerNormNHWC.apply(t, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        return ops.fuse_sum([t, self.mlp.forward_nchw(u, prepacked=pre)], [0, 0], relu)

    def extra_repr(self):
        return "num_heads={}, window_size={}, mlp_ratio={}".format(self.num_heads, self.window_size, self.mlp_ratio)


class SimpleFusion8(nn.Module):
    """hrnet_aux.py:42-68: bilinear(align_corners=True) up-sampling of the 3 coarse maps + concat + 1x1 conv + BN + ReLU."""

    def __init__(self, in_channels):
        super().__init__()
        self.fuse_conv = nn.Sequential(nn.Conv2d(in_channels, in_channels, 1), FusedBNAct(in_channels, _lib.ACT_RELU),
                                       nn.Identity())          # index 2 was nn.ReLU(True): fused into index 1

    def forward(self, feat_list):
        x0 = feat_list[0]
        cat = ops.NeckGather.apply(*feat_list)
        x = self.fuse_conv[1].after_conv(cat, self.fuse_conv[0].weight, self.fuse_conv[0].bias)
        return x, x0
