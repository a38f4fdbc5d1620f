"""HRNetV2-W32 backbone with one transformer block per multi-branch module, NHWC / sm_100a.

Mirror of RSSFormer-TIP2023/module/baseline/base_hrnet/_hrnet_rssformer.py (class names, constructor
signatures and state_dict keys preserved; file:line cited per class).  Every conv -> BN -> ReLU
(+ residual) chain runs as  conv kernel -> fused BN-statistics -> one fused apply pass.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, conv as convmod, ops
from .conv import conv2d, conv_bn_stats
from .modules import FusedBNAct, GeneralTransformerBlock, BN_MOMENTUM

# _hrnet_rssformer.py:97-125 (only the widths RSSFormer's configs use are listed: hrnetw32.py:9, hrnetw40.py)
MODEL_EXTRA = {
    "hrnetv2_w32": dict(
        stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=(4,), num_channels=(64,), fuse_method="SUM"),
        stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=(4, 4), num_channels=(32, 64), fuse_method="SUM"),
        stage3=dict(num_modules=4, num_branches=3, block="BASIC", num_blocks=(4, 4, 4), num_channels=(32, 64, 128), fuse_method="SUM"),
        stage4=dict(num_modules=3, num_branches=4, block="BASIC", num_blocks=(4, 4, 4, 4), num_channels=(32, 64, 128, 256), fuse_method="SUM")),
}


# RSS_BRANCH_STREAMS: 0 = everything on one stream, 1 = branches forked/joined inside each module, 2 (default) = data-flow schedule
# (HighResolutionModule.flow)
BRANCH_STREAMS = {"on": os.environ.get("RSS_BRANCH_STREAMS", "2") != "0", "flow": os.environ.get("RSS_BRANCH_STREAMS", "2") == "2"}
_SIDE = {}


# RSS_PRIORITY=1: the forward/backward chains (this module's streams, trainer.GraphedTrainStep's capture stream) run at HIGH stream
# priority and the weight-gradient side streams (conv.py) at the default one.  Motivation: the long library wgrad kernels occupy the
# SMs whenever a chain kernel becomes ready (10-25 us start gaps on almost every kernel of the backward chain, tools/timeline.py).
# Measured on the B=16 step: 479 img/s with priorities vs 490 without -- the weight gradients pile up behind the chain and the
# step ends with a serial tail -- so it is OFF by default.
CHAIN_PRIORITY = -1 if os.environ.get("RSS_PRIORITY", "0") != "0" else 0
FUSE_ORDER_DESC = os.environ.get("RSS_FUSE_ORDER", "1") != "0"      # see HighResolutionModule.flow
FUSE_CHAIN_STREAMS = os.environ.get("RSS_FUSE_CHAIN_STREAMS", "0") != "0"
# RSS_FUSE_ROW_STREAMS (default on): fuse row i >= 1 (its stride-2 / 1x1 chains and its sum) runs on a stream of its own, R[i],
# instead of on branch i's stream S[i].  Forward it is the same DAG.  Backward it removes a FALSE dependency that autograd's stream
# contract creates: gradients for a tensor with several consumers are accumulated on the stream of the tensor's PRODUCER node, at
# the moment the engine hands them over (issue order).  out[i] (branch i's output, produced on S[i]) is consumed by every row, so
# the accumulation "d out[3] += f_13^T(..)" -- and with it a wait for row 1's whole backward chain -- was queued on S[3] BEFORE the
# engine issued row 3's own backward nodes on that same stream: the rows' backward chains ran one after the other (row 1, row 2,
# row 3) although they are independent, and stream 0, which needs f_30^T from the LAST of them, idled ~450 us per stage-4 module
# (profiles/timeline_r2_final_526_summary.txt: "879:287 add", "873:307 add").  With R[i] only true dependencies are queued in front
# of a row's backward.
FUSE_ROW_STREAMS = os.environ.get("RSS_FUSE_ROW_STREAMS", "1") != "0"
_ROW = {}


def _side_streams(dev, n):
    lst = _SIDE.setdefault(dev, [])
    while len(lst) < n:
        lst.append(torch.cuda.Stream(dev, priority=CHAIN_PRIORITY))
    return lst


def _row_streams(dev, n):
    lst = _ROW.setdefault(dev, [])
    while len(lst) < n:
        lst.append(torch.cuda.Stream(dev, priority=CHAIN_PRIORITY))
    return lst


def _conv(cin, cout, k, stride=1, padding=0):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=padding, bias=False)


def _run(conv, bn, x, residual=None, link=None):
    """conv (no bias) -> fused BN(+act)(+residual); the batch statistics come out of the conv kernel's epilogue when the
    geometry is one csrc/conv_cf.cu covers.  link: residual-gradient hand-off of a Bottleneck (conv._ConvLib): the call WITHOUT a
    residual is the 1x1 convolution that consumes it, the call WITH a residual the BatchNorm that deposits it."""
    y, aff = conv_bn_stats(x, conv.weight, conv.stride[0], conv.padding[0], conv.dilation[0], bn.stats_args(),
                           link=link if residual is None else None)
    return bn(y, residual, aff=aff, link=link if residual is not None else None)


# RSS_RES_LINK=0: autograd sums the two gradients of a Bottleneck's input (1x1 conv + identity residual) with an elementwise kernel
RES_LINK = os.environ.get("RSS_RES_LINK", "1") != "0"


# RSS_BLOCK_FUSED (default on): stride-1 BasicBlocks without downsample whose geometry csrc/conv_cf.cu instantiates (C -> C 3x3, C in
# RSS_BLOCK_FUSED_C, default 32 = HRNet branch 0, the critical stream) run as ONE autograd node built from the fused tcgen05 conv:
#   forward   conv1 (+ bn1 statistics in the epilogue) -> bn1+ReLU -> conv2 (+ bn2 statistics) -> bn2 + residual + ReLU
#   backward  bn2 backward (reduce, apply) -> conv2 data gradient whose epilogue masks with bn1's ReLU and reduces bn1's backward sums
#             -> bn1 backward apply -> conv1 data gradient whose epilogue adds the residual-path gradient; weight gradients on the
#             side streams.  7 chain kernels -> 5, no library conv, no autograd accumulation kernel.
BLOCK_FUSED = {"on": os.environ.get("RSS_BLOCK_FUSED", "1") != "0",
               "channels": tuple(int(c) for c in os.environ.get("RSS_BLOCK_FUSED_C", "32").split(",") if c),
               # xform: conv2 applies bn1+ReLU to its staged input tile instead of reading a materialised activation
               "xform": os.environ.get("RSS_BLOCK_XFORM", "0") != "0",
               # chain: the last kernel of block k+1's backward (conv1 data gradient + residual gradient) also masks with block k's
               # output ReLU and reduces block k's bn2 backward sums, so block k starts with its apply pass
               "chain": os.environ.get("RSS_BLOCK_CHAIN", "0") != "0"}
# Both variants are bit-checked (test_fused_block_chain_vs_unfused) and OFF by default: measured on the B=16 step (gpurun 2026-10-17,
# profiles/ab_block_variants_r2.txt) 506.3 img/s with both, 506.5 xform only, 507.6 chain only, 510.0 with neither -- the in-place
# transform pass and the extra epilogue operands cost what the removed bn_act / bn_bwd_reduce launches saved.


class _BasicBlockFn(torch.autograd.Function):
    """relu(bn2(conv2(relu(bn1(conv1(x))))) + x) for a training-mode block owned by trainer.FlatSGD (parameter gradients are
    accumulated straight into the flat gradient buffer, so the parameters are not autograd inputs).  _hrnet_rssformer.py:230-246"""

    @staticmethod
    def forward(ctx, x, blk):
        lib = _lib.load()
        x = ops.nhwc(x)
        B, C, H, W = x.shape
        rows, dt, st = B * H * W, _lib.RSS_BF16, ops._st()
        w1, n1, dy1, dx1, ws1, k1 = convmod.cf_weight(blk.conv1.weight, False)
        z1, aff1 = convmod._cf_launch(x, w1, n1, dy1, dx1, C, C, None, False, blk.bn1.stats_args(), wstrides=ws1)
        w2, n2, dy2, dx2, ws2, k2 = convmod.cf_weight(blk.conv2.weight, False)
        if BLOCK_FUSED["xform"]:
            a1 = None                                  # relu(bn1(z1)) exists only inside conv2's staged tiles
            z2, aff2 = convmod._cf_launch(z1, w2, n2, dy2, dx2, C, C, aff1, True, blk.bn2.stats_args(), wstrides=ws2)
        else:
            a1 = torch.empty_like(x, memory_format=ops.CL)
            ops.account("bn", z1, a1)
            ops.check(lib.rss_bn_act_fwd(z1.data_ptr(), None, a1.data_ptr(), aff1[2].data_ptr(), aff1[3].data_ptr(), rows, C, _lib.ACT_RELU,
                                         dt, st), "rss_bn_act_fwd")
            z2, aff2 = convmod._cf_launch(a1, w2, n2, dy2, dx2, C, C, None, False, blk.bn2.stats_args(), wstrides=ws2)
        out = torch.empty_like(x, memory_format=ops.CL)
        ops.account("bn", z2, x, out)
        ops.check(lib.rss_bn_act_fwd(z2.data_ptr(), x.data_ptr(), out.data_ptr(), aff2[2].data_ptr(), aff2[3].data_ptr(), rows, C,
                                     _lib.ACT_RELU, dt, st), "rss_bn_act_fwd")
        for bn in (blk.bn1, blk.bn2):
            if not bn.defer_counter:
                bn.num_batches_tracked += 1
        ctx.save_for_backward(x, z1, a1, z2, out, aff1, aff2)
        ctx.blk = blk
        # hand-off for the chained backward: the producer of x (the previous fused block of this branch) left its bn2 context on the
        # tensor; this block's last backward kernel will do that block's bn2 masking + reduction and leave the sums in `box`
        ctx.prev = getattr(x, "_rss_bn2_ctx", None) if BLOCK_FUSED["chain"] else None
        ctx.box = {}
        if BLOCK_FUSED["chain"]:
            out._rss_bn2_ctx = (z2, aff2, blk.bn2._scratch, ctx.box)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, z1, a1, z2, out, aff1, aff2 = ctx.saved_tensors
        blk = ctx.blk
        dout = ops.nhwc(dout)
        if dout.dtype != x.dtype:
            dout = dout.to(x.dtype)
        B, C, H, W = x.shape
        rows, dt, st = B * H * W, _lib.RSS_BF16, ops._st()
        p = ops._p
        relu = _lib.ACT_RELU
        # bn2 backward (residual layer: the ReLU mask comes from the stored block output)
        sc2 = blk.bn2._scratch
        sums2 = torch.empty(2 * C, device=x.device, dtype=torch.float32)
        ops.account("bn", z2, out, dout, z2, out, dout, x, x, z1, x, x)      # bn2 reduce + apply (2 outputs), bn1 apply (z1, g1 -> dz1)
        if "sums" in ctx.box:          # the next block's last backward kernel already masked dout and reduced (sum g, sum g*xhat)
            sums2 = ctx.box.pop("sums")
        else:
            ops.check(lib.rss_bn_bwd_reduce_ws(p(z2), p(out), p(dout), p(aff2[2]), p(aff2[3]), p(aff2[0]), p(aff2[1]), p(sums2), p(sc2[2:]),
                                               p(sc2), None, rows, C, relu, dt, st), "rss_bn_bwd_reduce")
        dz2 = torch.empty_like(x, memory_format=ops.CL)
        dres = torch.empty_like(x, memory_format=ops.CL)
        ops.check(lib.rss_bn_bwd_apply(p(z2), p(out), p(dout), p(aff2[2]), p(aff2[3]), p(aff2[0]), p(aff2[1]), p(sums2), 1.0 / rows,
                                       p(dz2), p(dres), rows, C, relu, dt, p(sums2), p(blk.bn2.weight.grad), p(blk.bn2.bias.grad), st),
                  "rss_bn_bwd_apply")
        # (weight gradients are issued AFTER the data-gradient kernel they could compete with: the side stream then waits for it)
        # conv2 data gradient; its epilogue applies bn1's ReLU mask and reduces bn1's backward sums
        w2, n2, dy2, dx2, ws2, k2 = convmod.cf_weight(blk.conv2.weight, True)
        g1, sums1 = convmod._cf_launch(dz2, w2, n2, dy2, dx2, C, C, None, False, None, bnred=(z1, None, aff1, True, blk.bn1._scratch),
                                       wstrides=ws2)
        make_a1 = None
        if a1 is None:                 # re-create relu(bn1(z1)) on the weight-gradient stream, off the critical chain
            def make_a1():
                t = torch.empty_like(z1, memory_format=ops.CL)
                ops.check(lib.rss_bn_act_fwd(p(z1), None, p(t), p(aff1[2]), p(aff1[3]), rows, C, relu, dt, ops._st()), "rss_bn_act_fwd")
                return t
        convmod._wgrad(dz2, z1 if a1 is None else a1, convmod.lowp_cl(blk.conv2.weight, x.dtype), blk.conv2.weight, None, False, 1, 1, 1,
                       blk.conv2.weight.dtype, make_x=make_a1)
        dz1 = torch.empty_like(x, memory_format=ops.CL)
        ops.check(lib.rss_bn_bwd_apply(p(z1), None, p(g1), p(aff1[2]), p(aff1[3]), p(aff1[0]), p(aff1[1]), p(sums1), 1.0 / rows,
                                       p(dz1), None, rows, C, relu, dt, p(sums1), p(blk.bn1.weight.grad), p(blk.bn1.bias.grad), st),
                  "rss_bn_bwd_apply")
        # conv1 data gradient + the residual-path gradient in the epilogue
        w1, n1, dy1, dx1, ws1, k1 = convmod.cf_weight(blk.conv1.weight, True)
        if ctx.prev is not None:       # x is the previous fused block's output: its bn2 backward reduction rides in this epilogue
            pz2, paff2, pscratch, pbox = ctx.prev
            dx, psums = convmod._cf_launch(dz1, w1, n1, dy1, dx1, C, C, None, False, None, add=dres,
                                           bnred=(pz2, x, paff2, True, pscratch), wstrides=ws1)
            pbox["sums"] = psums
        else:
            dx, _ = convmod._cf_launch(dz1, w1, n1, dy1, dx1, C, C, None, False, None, add=dres, wstrides=ws1)
        convmod._wgrad(dz1, x, convmod.lowp_cl(blk.conv1.weight, x.dtype), blk.conv1.weight, None, False, 1, 1, 1, blk.conv1.weight.dtype)
        return dx, None


def _block_fused_ok(blk, x):
    if not (BLOCK_FUSED["on"] and blk.training and blk.downsample is None and blk.stride == 1 and x.is_cuda
            and x.dtype == torch.bfloat16 and torch.is_grad_enabled() and x.requires_grad):
        return False
    C = x.shape[1]
    if C not in BLOCK_FUSED["channels"] or C not in convmod.CF_SQUARE_3X3:
        return False
    for bn in (blk.bn1, blk.bn2):
        if not bn.training or bn.stats_args() is None:
            return False
    for prm in (blk.conv1.weight, blk.conv2.weight, blk.bn1.weight, blk.bn1.bias, blk.bn2.weight, blk.bn2.bias):
        if not getattr(prm, "_rss_flat", False) or ops.grad_sink(prm) is None:      # direct accumulation needs trainer.FlatSGD
            return False
    B, _, H, W = x.shape
    lib = _lib.load()
    return bool(lib.rss_conv_cf_supported(B, H, W, C, C, 3, _lib.CF_STATS)) and bool(lib.rss_conv_cf_supported(B, H, W, C, C, 3, _lib.CF_BNRED))


class BasicBlock(nn.Module):
    """_hrnet_rssformer.py:216-246"""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride, 1)
        self.bn1 = FusedBNAct(planes, _lib.ACT_RELU)
        self.conv2 = _conv(planes, planes, 3, 1, 1)
        self.bn2 = FusedBNAct(planes, _lib.ACT_RELU)          # relu(bn2(.) + residual)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        if _block_fused_ok(self, x):
            return _BasicBlockFn.apply(x, self)
        residual = x if self.downsample is None else _run(self.downsample[0], self.downsample[1], x)
        out = _run(self.conv1, self.bn1, x)
        return _run(self.conv2, self.bn2, out, residual)


class Bottleneck(nn.Module):
    """_hrnet_rssformer.py:249-287"""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 1)
        self.bn1 = FusedBNAct(planes, _lib.ACT_RELU)
        self.conv2 = _conv(planes, planes, 3, stride, 1)
        self.bn2 = FusedBNAct(planes, _lib.ACT_RELU)
        self.conv3 = _conv(planes, planes * self.expansion, 1)
        self.bn3 = FusedBNAct(planes * self.expansion, _lib.ACT_RELU)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        residual = x if self.downsample is None else _run(self.downsample[0], self.downsample[1], x)
        # identity residual: x feeds conv1 and the add behind bn3 -- their two gradients are merged by conv1's data-gradient GEMM
        link = {} if (RES_LINK and self.downsample is None and x.is_cuda and torch.is_grad_enabled() and x.requires_grad) else None
        out = _run(self.conv1, self.bn1, x, link=link)
        out = _run(self.conv2, self.bn2, out)
        return _run(self.conv3, self.bn3, out, residual, link=link)


blocks_dict = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}


def _downsample(cin, cout, stride):
    return nn.Sequential(_conv(cin, cout, 1, stride), FusedBNAct(cout, _lib.ACT_NONE))


class HighResolutionModule(nn.Module):
    """_hrnet_rssformer.py:290-437"""

    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method, multi_scale_output=True):
        super().__init__()
        if not (num_branches == len(num_blocks) == len(num_channels) == len(num_inchannels)):
            raise ValueError("NUM_BRANCHES({}) <> NUM_BLOCKS/NUM_CHANNELS/NUM_INCHANNELS".format(num_branches))
        self.num_inchannels = num_inchannels
        self.fuse_method = fuse_method
        self.num_branches = num_branches
        self.multi_scale_output = multi_scale_output
        self.branches = nn.ModuleList([self._make_one_branch(i, blocks, num_blocks, num_channels) for i in range(num_branches)])
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(False)
        self.transformer = GeneralTransformerBlock(num_channels[0], planes=num_channels[0], num_heads=2)   # :308

    def _make_one_branch(self, i, block, num_blocks, num_channels, stride=1):
        downsample = None
        if stride != 1 or self.num_inchannels[i] != num_channels[i] * block.expansion:
            downsample = _downsample(self.num_inchannels[i], num_channels[i] * block.expansion, stride)
        layers = [block(self.num_inchannels[i], num_channels[i], stride, downsample)]
        self.num_inchannels[i] = num_channels[i] * block.expansion
        layers += [block(self.num_inchannels[i], num_channels[i]) for _ in range(1, num_blocks[i])]
        return nn.Sequential(*layers)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        nb, ch = self.num_branches, self.num_inchannels
        fuse_layers = []
        for i in range(nb if self.multi_scale_output else 1):
            row = []
            for j in range(nb):
                if j > i:       # 1x1 conv + BN + nearest up-sampling (:372-380)
                    row.append(nn.Sequential(_conv(ch[j], ch[i], 1), FusedBNAct(ch[i], _lib.ACT_NONE),
                                             nn.Upsample(scale_factor=2 ** (j - i), mode="nearest")))
                elif j == i:
                    row.append(None)
                else:           # chain of stride-2 3x3 convs, ReLU on all but the last (:384-402)
                    chain = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        cout = ch[i] if last else ch[j]
                        seq = [_conv(ch[j], cout, 3, 2, 1), FusedBNAct(cout, _lib.ACT_NONE if last else _lib.ACT_RELU)]
                        if not last:
                            seq.append(nn.Identity())       # was nn.ReLU(False): fused into the BN pass
                        chain.append(nn.Sequential(*seq))
                    row.append(nn.Sequential(*chain))
            fuse_layers.append(nn.ModuleList(row))
        return nn.ModuleList(fuse_layers)

    def get_num_inchannels(self):
        return self.num_inchannels

    def _fuse(self, i, j, x):
        """fuse branch j into resolution i: returns (term, log2 of the nearest up-sampling still to apply)"""
        layer = self.fuse_layers[i][j]
        if j > i:
            return _run(layer[0], layer[1], x), j - i          # the x2^(j-i) nearest up-sampling is fused into the sum kernel
        for seq in layer:
            x = _run(seq[0], seq[1], x)
        return x, 0

    def _run_branches(self, x):
        """RSS_BRANCH_STREAMS=1 schedule: the branches of a module run on side streams and are joined before the fuse step."""
        if not BRANCH_STREAMS["on"] or not x[0].is_cuda:
            return [self.branches[i](x[i]) for i in range(self.num_branches)]
        dev = x[0].device
        cur = torch.cuda.current_stream(dev)
        side = _side_streams(dev, self.num_branches - 1)
        out = [None] * self.num_branches
        for i in range(1, self.num_branches):
            s = side[i - 1]
            s.wait_stream(cur)
            x[i].record_stream(s)
            with torch.cuda.stream(s):
                out[i] = self.branches[i](x[i])
        out[0] = self.branches[0](x[0])
        for i in range(1, self.num_branches):
            cur.wait_stream(side[i - 1])
            out[i].record_stream(cur)
        return out

    def forward(self, x):
        if self.num_branches == 1:
            return [self.branches[0](x[0])]
        if _dataflow(x[0]):
            return flow_join(self.flow(x))
        x = self._run_branches(x)
        x_fuse = []
        for i in range(len(self.fuse_layers)):
            terms, ks = [], []
            for j in range(0 if i > 0 else 1, self.num_branches):
                t, k = (x[j], 0) if i == j else self._fuse(i, j, x[j])
                terms.append(t); ks.append(k)
            if i == 0:
                low = ops.fuse_sum(terms, ks, relu=False)
                x_fuse.append(self.transformer(low, x[0], relu=True))   # (:430-431,435) x[0] enters only as keys/values
            else:
                x_fuse.append(ops.fuse_sum(terms, ks, relu=True))      # (:433,435)
        return x_fuse

    def flow(self, x):
        """Data-flow schedule of the module (default on CUDA): resolution i lives on stream i (0 = the caller's stream) ACROSS
        modules.  The only cross-stream edges are the ones the reference's data flow has -- every fuse row reads every branch
        output -- expressed as events recorded right after the producing kernel, so nothing waits for more than it needs:
          * branch i, then the terms it contributes to row 0 (1x1 conv + BN at its own resolution), on stream i;
          * fuse rows i >= 1 (stride-2 conv chains + sum) on stream i, issued BEFORE the transformer so they do not wait for it;
          * row-0 sum + the transformer block (the long pole: ~0.6 ms forward, ~1.2 ms backward) on stream 0.
        The coarse resolutions of the NEXT module start as soon as their own row is done, i.e. they overlap this module's
        transformer instead of queueing behind it, and autograd replays the same structure in the backward pass.  Inside the
        captured CUDA graph the streams become parallel sub-graphs.  Returns tensors still living on their streams
        (flow_join() before handing them to stream-unaware code)."""
        nb = self.num_branches
        dev = x[0].device
        cur = torch.cuda.current_stream(dev)
        S = [cur] + _side_streams(dev, nb - 1)[:nb - 1]
        out = [None] * nb
        for i in range(nb - 1, -1, -1):                    # coarse branches first: their launches are queued before the long one
            with torch.cuda.stream(S[i]):
                out[i] = _mark(self.branches[i](_bring(x[i], S[i], cur)), S[i])
        terms0 = []
        for j in range(1, nb):
            with torch.cuda.stream(S[j]):
                t, k = self._fuse(0, j, out[j])
                terms0.append((_mark(t, S[j]), k))
        rows = len(self.fuse_layers)
        x_fuse = [None] * rows
        R = ([cur] + _row_streams(dev, rows - 1)[:rows - 1]) if FUSE_ROW_STREAMS else S
        for i in range(rows - 1, 0, -1):
            with torch.cuda.stream(R[i]):
                # Issued finest source LAST: autograd replays a stream's nodes in reverse, so the stride-2 chain coming from branch 0
                # (f_i0, the longest: i convolutions) runs FIRST in the backward pass.  Its result is the gradient stream 0's chain
                # waits for; issued first (j ascending) it sat behind the chains of the other sources on this stream -- on the
                # coarsest stream 6 conv+BN pairs instead of 3, a 214 us stall of the critical chain per stage-4 module
                # (profiles/timeline_r2_final_526_summary.txt).  The sum keeps its term order.  RSS_FUSE_ORDER=0: ascending.
                tk = [None] * nb
                order = range(nb - 1, -1, -1) if FUSE_ORDER_DESC else range(nb)
                for j in order:
                    if i == j:
                        tk[j] = (_bring(out[j], R[i], cur), 0)
                    elif FUSE_CHAIN_STREAMS and j < i:
                        # RSS_FUSE_CHAIN_STREAMS=1: every stride-2 chain f_ij is its own branch of the graph (own stream, joined into
                        # row i's stream by the sum), so that in the backward pass f_i0 does not queue behind its siblings
                        sc = _side_streams(dev, nb - 1 + 8)[nb - 1 + (i * (i - 1)) // 2 + j]
                        with torch.cuda.stream(sc):
                            t, k = self._fuse(i, j, _bring(out[j], sc, cur))
                            _mark(t, sc)
                        tk[j] = (_bring(t, R[i], cur), k)
                    else:
                        tk[j] = self._fuse(i, j, _bring(out[j], R[i], cur))
                terms, ks = [t for t, _ in tk], [k for _, k in tk]
                x_fuse[i] = _mark(ops.fuse_sum(terms, ks, relu=True), R[i])      # (:433,435)
        low = ops.fuse_sum([_bring(t, cur, cur) for t, _ in terms0], [k for _, k in terms0], relu=False)
        x_fuse[0] = _mark(self.transformer(low, out[0], relu=True), cur)         # (:430-431,435)
        return x_fuse


def _dataflow(t):
    return BRANCH_STREAMS["on"] and BRANCH_STREAMS["flow"] and t.is_cuda


def _mark(t, s):
    """remember the stream a tensor was produced on and an event right after its producer"""
    ev = torch.cuda.Event()
    ev.record(s)
    t._rss_home, t._rss_ev = s, ev
    return t


def _bring(t, s, origin):
    """make stream `s` safe to read `t`: wait for t's producer only (not for whatever was queued on its stream afterwards).
    A tensor without a mark was produced on `origin` (the stream of the stream-unaware caller)."""
    h = getattr(t, "_rss_home", None)
    if h is None:
        if origin != s:
            s.wait_stream(origin)
            t.record_stream(s)
    elif h != s:
        s.wait_event(t._rss_ev)
        t.record_stream(s)
    return t


def flow_join(tensors):
    """hand data-flow tensors to the current stream"""
    cur = torch.cuda.current_stream(tensors[0].device)
    for t in tensors:
        _bring(t, cur, cur)
        if hasattr(t, "_rss_home"):
            del t._rss_home, t._rss_ev
    return tensors


class HighResolutionNet(nn.Module):
    """_hrnet_rssformer.py:446-650"""

    def __init__(self, extra, norm_eval=True, zero_init_residual=False, frozen_stages=-1):
        super().__init__()
        self.norm_eval, self.frozen_stages, self.zero_init_residual, self.extra = norm_eval, frozen_stages, zero_init_residual, extra
        self.conv1 = _conv(3, 64, 3, 2, 1)
        self.bn1 = FusedBNAct(64, _lib.ACT_RELU)
        self.conv2 = _conv(64, 64, 3, 2, 1)
        self.bn2 = FusedBNAct(64, _lib.ACT_RELU)
        self.relu = nn.ReLU(inplace=True)

        self.stage1_cfg = extra["stage1"]
        block = blocks_dict[self.stage1_cfg["block"]]
        nch = self.stage1_cfg["num_channels"][0]
        self.layer1 = self._make_layer(block, 64, nch, self.stage1_cfg["num_blocks"][0])
        pre = [nch * block.expansion]
        for s in (2, 3, 4):
            cfg = extra["stage%d" % s]
            setattr(self, "stage%d_cfg" % s, cfg)
            block = blocks_dict[cfg["block"]]
            nch = [c * block.expansion for c in cfg["num_channels"]]
            setattr(self, "transition%d" % (s - 1), self._make_transition_layer(pre, nch))
            stage, pre = self._make_stage(cfg, nch)
            setattr(self, "stage%d" % s, stage)
        if frozen_stages >= 0:
            raise NotImplementedError("frozen_stages >= 0 is not used by the RSSFormer config (hrnetw32.py:12)")

    def _make_transition_layer(self, pre, cur):
        layers = []
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    layers.append(nn.Sequential(_conv(pre[i], cur[i], 3, 1, 1), FusedBNAct(cur[i], _lib.ACT_RELU), nn.Identity()))
                else:
                    layers.append(None)
            else:
                chain = []
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    chain.append(nn.Sequential(_conv(cin, cout, 3, 2, 1), FusedBNAct(cout, _lib.ACT_RELU), nn.Identity()))
                layers.append(nn.Sequential(*chain))
        return nn.ModuleList(layers)

    def _make_layer(self, block, inplanes, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:
            downsample = _downsample(inplanes, planes * block.expansion, stride)
        layers = [block(inplanes, planes, stride, downsample)]
        layers += [block(planes * block.expansion, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _make_stage(self, cfg, num_inchannels, multi_scale_output=True):
        modules = []
        for i in range(cfg["num_modules"]):
            reset = not (not multi_scale_output and i == cfg["num_modules"] - 1)
            modules.append(HighResolutionModule(cfg["num_branches"], blocks_dict[cfg["block"]], cfg["num_blocks"],
                                                num_inchannels, list(cfg["num_channels"]), cfg["fuse_method"], reset))
            num_inchannels = modules[-1].get_num_inchannels()
        return nn.Sequential(*modules), num_inchannels

    @staticmethod
    def _transition(layer, x):
        if layer[0].__class__ is nn.Sequential:        # chain of stride-2 conv+BN+ReLU
            for seq in layer:
                x = _run(seq[0], seq[1], x)
            return x
        return _run(layer[0], layer[1], x)

    def _stage(self, stage, x_list):
        if not _dataflow(x_list[0]):
            return stage(x_list)
        for m in stage:                                  # tensors stay on their resolution's stream from module to module
            x_list = m.flow(x_list) if m.num_branches > 1 else m(x_list)
        return x_list

    def _next_inputs(self, transitions, y_list, n):
        """inputs of the next stage: existing resolutions pass through (or get their 3x3 conv), the new coarsest one is a
        stride-2 chain from the last output; in the data-flow schedule each lands on its resolution's stream"""
        if not _dataflow(y_list[0]):
            return [y_list[i] if transitions[i] is None else self._transition(transitions[i], y_list[min(i, len(y_list) - 1)])
                    for i in range(n)]
        dev = y_list[0].device
        cur = torch.cuda.current_stream(dev)
        S = [cur] + _side_streams(dev, n - 1)[:n - 1]
        x_list = []
        for i in range(n):
            if transitions[i] is None:
                x_list.append(y_list[i])
                continue
            src = y_list[min(i, len(y_list) - 1)]
            with torch.cuda.stream(S[i]):
                x_list.append(_mark(self._transition(transitions[i], _bring(src, S[i], cur)), S[i]))
        return x_list

    stem_dtype = None      # set by model.HRNetFusion: activation dtype of the model when forward() is handed the raw image batch

    def _stem1(self, x):
        """conv1 + bn1 + ReLU (:467-470, 531-533).  On a bf16 model the planar image batch, exactly as the reference model receives
        it, goes through csrc/stem.cu: cast, NCHW -> NHWC, the 3 -> 64 stride-2 convolution and bn1's raw sums in ONE launch (the
        library path needs a cast, a permute, a 3 -> 8 channel padding kernel and a legacy implicit GEMM: 224 us vs the 184 MB the
        layer has to move).  Anything else (fp32 strict-parity model, channels-last input, RSS_STEM=0) takes the library conv."""
        dt = self.stem_dtype
        if dt is torch.bfloat16 and ops.stem_conv_ok(x, self.conv1.weight):
            bn = self.bn1
            raw = ops.STEM["stats"] and ops.bn_accepts_raw_sums(x, bn.training, True if bn.sync else None, bn._scratch, bn.num_features)
            y = ops.StemConv.apply(x, self.conv1.weight, (bn._scratch, bn.running_mean) if raw else None)
            return bn(y, aff=ops.RAW_SUMS if raw else None)
        if dt is not None and x.dtype != dt:
            x = x.to(dt)
        return _run(self.conv1, self.bn1, ops.nhwc(x))

    def forward(self, x):
        x = self._stem1(x)
        x = _run(self.conv2, self.bn2, x)
        x = self.layer1(x)
        x_list = self._next_inputs(self.transition1, [x], self.stage2_cfg["num_branches"])
        y_list = self._stage(self.stage2, x_list)
        x_list = self._next_inputs(self.transition2, y_list, self.stage3_cfg["num_branches"])
        y_list = self._stage(self.stage3, x_list)
        x_list = self._next_inputs(self.transition3, y_list, self.stage4_cfg["num_branches"])
        # trainer.FlatSGD (world > 1) sets _stage4_hook: called once per backward pass when the gradients of ALL stage-4 inputs exist,
        # i.e. when every backward node of stage 4, the neck and the head has been issued -- their gradients (a contiguous tail of
        # the flat buffer) can be all-reduced while stages 3..1 are still differentiating
        fn = getattr(self, "_stage4_hook", None)
        if fn is not None and torch.is_grad_enabled():
            ts = [t for t in x_list if t.requires_grad]
            left = {"n": len(ts)}

            def _cb(g):
                left["n"] -= 1
                if left["n"] == 0:
                    fn()
                return g
            for t in ts:
                t.register_hook(_cb)
        y_list = self._stage(self.stage4, x_list)
        return flow_join(y_list) if _dataflow(y_list[0]) else y_list

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, FusedBNAct) and not m.sync:
                    m.eval()
        return self


def _factory(name):
    def build(pretrained=False, weight_path=None, norm_eval=False, frozen_stages=-1):
        model = HighResolutionNet(MODEL_EXTRA[name], norm_eval, zero_init_residual=False, frozen_stages=frozen_stages)
        if pretrained:
            if weight_path is None:
                raise FileNotFoundError("no network here: pass weight_path for pretrained=True (_hrnet_rssformer.py:669-675)")
            model.load_state_dict(torch.load(weight_path, map_location="cpu"), strict=False)
        return model
    build.__name__ = name
    return build


hrnetv2_w32 = _factory("hrnetv2_w32")
