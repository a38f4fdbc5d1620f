"""torch.autograd.Function wrappers around the C-ABI kernels (include/rss_b200.h).

PyTorch is plumbing here: it owns device memory and streams and records the tape; every forward and
backward below is a call into librss_b200.so on `torch.cuda.current_stream()`.  Tensors cross the
boundary as raw NHWC device pointers (channels_last memory format on the torch side).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import AttnGrads, AttnParams
from ._lib import check as _check

CL = torch.channels_last
# kernels launched per C-ABI call (for bench.py's `gpu_launches`; counted from the csrc/*.cu launch sites)
KERNELS_PER_CALL = {
    "rss_layernorm_fwd": 1, "rss_layernorm_bwd": 1, "rss_attn_fwd": 5, "rss_spatial_attention_fwd": 2, "rss_bilinear_resize": 1, "rss_confusion_matrix": 1, "rss_accum_bf16_list": 1, "rss_attn_bwd": 7, "rss_bn_stats": 1, "rss_bn_combine": 1,
    "rss_bn_finalize": 1, "rss_bn_eval_affine": 1, "rss_bn_act_fwd": 1, "rss_bn_bwd_reduce": 1, "rss_bn_bwd_apply": 1, "rss_sync_bn_finalize": 1, "rss_sync_allreduce_small": 1, "rss_bn_stats_raw": 1,
    "rss_neck_gather_fwd": 1, "rss_neck_gather_bwd": 1, "rss_head_fwd": 1, "rss_head_bwd": 1, "rss_head_probs": 1,
    "rss_headaux_fwd": 2, "rss_seg_loss_fwd": 2, "rss_seg_loss_bwd": 1, "rss_grad_sumsq": 1, "rss_sgd_step": 1,
    "rss_conv_igemm": 1, "rss_conv_pack_weights": 1, "rss_conv_wgrad": 1, "rss_conv_wgrad_tc": 1, "rss_conv_cf": 1, "rss_shadow_t_refresh": 1, "rss_shadow_cl_refresh": 1, "rss_fuse_sum_fwd": 1, "rss_fuse_sum_bwd": 1,
}
COUNTERS = {"launches": 0, "calls": 0}
# algorithmic bytes (compulsory reads + writes of the tensors a call touches) per kernel family, filled only while bench.py asks
ACCOUNT = {}
ACCOUNT_ON = [False]


def account(family, *tensors):
    if ACCOUNT_ON[0]:
        ACCOUNT[family] = ACCOUNT.get(family, 0.0) + sum(t.numel() * t.element_size() for t in tensors if t is not None)
TIMED = {}            # op name -> list of (start_event, end_event), filled only while bench.py enables it
TIMED_OPS = set()


def check(rc, what):
    COUNTERS["launches"] += KERNELS_PER_CALL.get(what, 1)
    COUNTERS["calls"] += 1
    _check(rc, what)


class timed:
    """records CUDA events around a C-ABI call on the launching stream when bench.py asks for `name`"""

    def __init__(self, name):
        self.on = name in TIMED_OPS
        self.name = name

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if self.on:
            self.e1.record()
            TIMED.setdefault(self.name, []).append((self.e0, self.e1))


def _dt(t):
    if t.dtype == torch.float32:
        return _lib.RSS_F32
    if t.dtype == torch.bfloat16:
        return _lib.RSS_BF16
    raise _lib.RssError("unsupported activation dtype %s (float32 or bfloat16)" % t.dtype)


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t):
    """fp32 contiguous view of a parameter (parameters are fp32 masters; a cast is a bug upstream)."""
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.detach().float().contiguous()
    return t


def nhwc(t):
    """Return `t` (logical NCHW) laid out NHWC in memory."""
    return t.contiguous(memory_format=CL)


def grad_sink(p):
    """fp32 gradient buffer of a parameter that kernels may accumulate into directly (bypassing one AccumulateGrad
    add kernel per parameter per step: ~1.4 k launches).  None -> return the gradient through autograd as usual.
    Only parameters owned by trainer.FlatSGD take the direct path: for anyone else (a plain torch optimiser, DDP's reducer
    hooks, gradient accumulation with zero_grad(set_to_none=False)) AccumulateGrad and its hooks must run."""
    if p is None or not isinstance(p, torch.nn.Parameter) or not getattr(p, "_rss_flat", False):
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous():
        return None
    return g


import os
# (a one-launch BatchNorm with a device-wide spin barrier between the statistics and the apply phase was measured at 382 vs 418 img/s
#  -- the barrier costs more than the launch it saves -- and removed from the library after round 2)
BN_KEEP_DZ = {"on": os.environ.get("RSS_BN_KEEP_DZ", "1") != "0"}
# (round 2 also tried one-launch BatchNorm with 8-CTA thread-block clusters per 16-channel slice exchanging partial sums through
#  distributed shared memory -- no atomics, no tickets: correct, but 37-52 us vs 14-15 us for the two split kernels on the 8.4 MB
#  branch-1 tensors and no better on 2-4 MB ones (32-byte row slices at a 128-512 byte stride from 32-128 CTAs are L2-latency
#  bound); removed again, numbers in profiles/bn_cluster_vs_split_r2.txt, code in the history at commit "Cluster/DSMEM ...")
# "raw" BatchNorm protocol (csrc/bn.cu BnFin): the statistics / reduce kernels only add their sums into the layer's scratch and
# the apply kernels finalise -- fewer dependent global round trips per layer
# Measured on the B=16 step (gpurun 2026-10-17, timeline_s3f): statistics -3.3 us and reduce -2.1 us per layer, but the "last block
# finished" ticket costs more on the 1184-block apply grids (+6.9 / +3.2 us), 453 vs 458 img/s -> OFF by default.
BN_RAW = {"on": os.environ.get("RSS_BN_RAW", "0") != "0"}


# SyncBatchNorm statistics exchange over NVLink peer memory (csrc/sync_exchange.cu) instead of one NCCL collective per layer and pass.
# Needs torch.distributed._symmetric_memory on the default process group; anything else (sub-groups, rendezvous failure,
# RSS_SYNC_SYMM=0) keeps the NCCL path.
SYNC_SYMM = {"on": os.environ.get("RSS_SYNC_SYMM", "1") != "0", "obj": None, "failed": False}


class _SyncExchange:
    CHANNELS = 512                          # one per SyncBN layer (assigned in first-use order = program order, identical on all ranks)

    def __init__(self, dev):
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.chan_bytes = (lib.rss_sync_exchange_bytes(self.world) + 255) // 256 * 256
        if self.chan_bytes == 0:
            raise RuntimeError("world size not supported by the exchange kernel")
        self.buf = symm_mem.empty(self.CHANNELS * self.chan_bytes // 4, dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.bases = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self.counters = torch.zeros(self.CHANNELS, dtype=torch.int32, device=dev)
        self.channels = {}
        torch.cuda.synchronize(dev)
        dist.barrier()                      # every rank's buffer is zeroed and mapped before the first exchange

    def channel(self, scratch):
        """(byte offset, counter pointer) of the channel of the layer owning `scratch`; None when the channels are used up"""
        k = scratch.data_ptr()
        c = self.channels.get(k)
        if c is None:
            if len(self.channels) >= self.CHANNELS:
                return None
            c = self.channels[k] = len(self.channels)
        return c * self.chan_bytes, self.counters[c:c + 1].data_ptr()


def sync_exchange(group, dev, C):
    """the symmetric-memory exchange for SyncBN on the DEFAULT group, or None (then the NCCL collectives are used).  The first call
    is collective (rendezvous + barrier): it happens in the eager warm-up step, at the same layer on every rank."""
    if not SYNC_SYMM["on"] or SYNC_SYMM["failed"] or group is not True or dev.type != "cuda" or 2 * C + 1 > 2 * 512 + 7:
        return None
    if SYNC_SYMM["obj"] is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        try:
            SYNC_SYMM["obj"] = _SyncExchange(dev)
        except Exception as e:  # noqa: BLE001  (no peer mapping on this box / torch build: fall back to NCCL, on every rank alike)
            SYNC_SYMM["failed"] = True
            import sys
            sys.stderr.write("representationlearning_b200: symmetric-memory SyncBN exchange unavailable (%r), using NCCL\n" % (e,))
            return None
    return SYNC_SYMM["obj"]


def _world(group):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group if group is not True else None)


# ----------------------------------------------------------------------------------------------
# LayerNorm over channels of NHWC tokens
# ----------------------------------------------------------------------------------------------
class LayerNormNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _lib.require_device()
        lib = _lib.load()
        x = nhwc(x)
        B, C, H, W = x.shape
        rows = B * H * W
        y = torch.empty_like(x, memory_format=CL)
        stats = torch.empty(2, rows, device=x.device, dtype=torch.float32)
        g, b = _f32(gamma), _f32(beta)
        check(lib.rss_layernorm_fwd(_p(x), _p(y), _p(stats[0]), _p(stats[1]), _p(g), _p(b), eps, rows, C, _dt(x), _st()),
              "rss_layernorm_fwd")
        ctx.save_for_backward(x, stats, g)
        ctx.refs = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, stats, g = ctx.saved_tensors
        dy = nhwc(dy)
        B, C, H, W = x.shape
        rows = B * H * W
        dx = torch.empty_like(x, memory_format=CL)
        sg, sb = grad_sink(ctx.refs[0]), grad_sink(ctx.refs[1])
        direct = sg is not None and sb is not None
        dgb = None if direct else torch.zeros(2, C, device=x.device, dtype=torch.float32)
        check(lib.rss_layernorm_bwd(_p(dy), _p(x), _p(stats[0]), _p(stats[1]), _p(g), None, _p(dx),
                                    _p(sg if direct else dgb[0]), _p(sb if direct else dgb[1]), rows, C, _dt(x), _st()),
              "rss_layernorm_bwd")
        return (dx, None, None, None) if direct else (dx, dgb[0], dgb[1], None)


# ----------------------------------------------------------------------------------------------
# gate + window attention region
# ----------------------------------------------------------------------------------------------
_ATTN_NAMES = ("ln_w", "ln_b", "sa1_w", "sa2_w", "lvl_w", "lvl_b", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")


class WindowAttention(torch.autograd.Function):
    """out = [x +] Attn([LN1](x), [LN1](y)); tensors are (B,C,H,W) channels_last.
    params: 14 tensors in _ATTN_NAMES order; ln_w/ln_b may be None (no norm1); sa1_w..lvl_b may be None (no saliency gate)."""

    @staticmethod
    def forward(ctx, x, y, eps, residual, *params):
        _lib.require_device()
        lib = _lib.load()
        x, y = nhwc(x), nhwc(y)
        B, C, H, W = x.shape
        HW = H * W
        has_ln = params[0] is not None
        ps = [None if p is None else _f32(p) for p in params]
        ap = AttnParams()
        for n, p in zip(_ATTN_NAMES, ps):
            setattr(ap, n, _p(p))
        ap.ln_eps, ap.C, ap.num_heads, ap.window = float(eps), C, 2, 7
        dev = x.device
        ln_stats = torch.empty(4, B * HW, device=dev, dtype=torch.float32) if has_ln else None
        pooled = torch.empty(B, 4, HW, device=dev, dtype=torch.float32)
        amax = torch.empty(B, 2, HW, device=dev, dtype=torch.uint8)
        smap = torch.empty(B, 2, HW, device=dev, dtype=torch.float32)
        no_gate = params[2] is None                      # Mhca.forward alone: constant gate of ones, no gradient into a gate branch
        gmap = torch.ones(B, 2, HW, device=dev, dtype=torch.float32) if no_gate else torch.empty(B, 2, HW, device=dev, dtype=torch.float32)
        out = torch.empty_like(x, memory_format=CL)
        flags = (0 if residual else 1) | (4 if no_gate else 0)
        account("attn_fwd", x, y, out)
        with timed("rss_attn_fwd"):
            check(lib.rss_attn_fwd(_p(x), _p(y), ctypes.byref(ap), B, H, W, _dt(x), flags, _p(ln_stats),
                                   _p(pooled), _p(amax), _p(smap), _p(gmap), _p(out), _st()), "rss_attn_fwd")
        ctx.save_for_backward(x, y, ln_stats, pooled, amax, smap, gmap, *ps)
        ctx.eps, ctx.flags, ctx.refs = float(eps), flags, params
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, y, ln_stats, pooled, amax, smap, gmap = ctx.saved_tensors[:7]
        ps = ctx.saved_tensors[7:]
        dout = nhwc(dout)
        B, C, H, W = x.shape
        ap = AttnParams()
        for n, p in zip(_ATTN_NAMES, ps):
            setattr(ap, n, _p(p))
        ap.ln_eps, ap.C, ap.num_heads, ap.window = ctx.eps, C, 2, 7
        sinks = [grad_sink(r) for r in ctx.refs]
        direct = all(s is not None or p is None for s, p in zip(sinks, ps))
        grads = sinks if direct else [None if p is None else torch.zeros_like(p) for p in ps]
        ag = AttnGrads()
        for n, g in zip(_ATTN_NAMES, grads):
            setattr(ag, n, _p(g))
        wsb = lib.rss_attn_bwd_workspace_bytes(B, H, W, _dt(x))
        ws = torch.empty(wsb, device=x.device, dtype=torch.uint8)
        dx = torch.empty_like(x, memory_format=CL)
        dy = torch.empty_like(x, memory_format=CL)
        account("attn_bwd", dout, x, y, dx, dy)
        with timed("rss_attn_bwd"):
            check(lib.rss_attn_bwd(_p(dout), _p(x), _p(y), ctypes.byref(ap), B, H, W, _dt(x), ctx.flags,
                                   _p(ln_stats), _p(pooled), _p(amax), _p(smap), _p(gmap), _p(ws), wsb, _p(dx), _p(dy),
                                   ctypes.byref(ag), _st()), "rss_attn_bwd")
        return (dx, dy, None, None) + ((None,) * len(ps) if direct else tuple(grads))


# ----------------------------------------------------------------------------------------------
# BatchNorm(+SyncBN) + activation (+ residual)
# ----------------------------------------------------------------------------------------------
# BNAct(aff=RAW_SUMS): the layer's scratch already holds sum (x-K), sum (x-K)^2 with K = the running mean, produced by the epilogue of
# the convolution that wrote x (rss_conv_igemm_stats): the statistics pass over x is skipped and the apply kernel (one rank) or the
# peer-memory exchange kernel (SyncBN under a process group) finalises them.  Callers check bn_accepts_raw_sums() first.
RAW_SUMS = "raw-sums"


def bn_accepts_raw_sums(x, training, group, scratch, C):
    if not training or scratch is None or scratch.numel() < 2 + 2 * C or not x.is_cuda:
        return False
    world = _world(group)
    if world == 1:
        return True
    ex = sync_exchange(group, x.device, C)
    return ex is not None and ex.channel(scratch) is not None


class BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, training, momentum, eps, act, group, scratch=None,
                pre_bias=None, aff=None, link=None):
        """aff: (4,C) mean/invstd/scale/shift already produced (and running statistics already updated) by the convolution
        kernel's epilogue (rss_conv_cf): the statistics pass is skipped.
        pre_bias: bias of the producing conv, NOT added to x (training mode only: a per-channel shift cancels in the
        normalisation, so only the running mean has to see it; saves one full-tensor pass per biased conv)"""
        _lib.require_device()
        lib = _lib.load()
        x = nhwc(x)
        if residual is not None:
            residual = nhwc(residual)
            if residual.dtype != x.dtype:
                residual = residual.to(x.dtype)
        B, C, H, W = x.shape
        rows = B * H * W
        dev, dt, st = x.device, _dt(x), _st()
        g, b = _f32(gamma), _f32(beta)
        raw_ready = isinstance(aff, str) and aff == RAW_SUMS
        have_aff = aff is not None and not raw_ready
        if raw_ready and not bn_accepts_raw_sums(x, training, group, scratch, C):
            raise _lib.RssError("BNAct(aff=RAW_SUMS): this layer / process group cannot finalise raw sums")
        if not have_aff:
            aff = torch.empty(4, C, device=dev, dtype=torch.float32)      # mean, invstd, scale, shift
        world = _world(group) if training else 1
        if have_aff and (not training or world != 1):
            raise _lib.RssError("a precomputed BatchNorm affine is only valid for single-rank training-mode statistics")
        if pre_bias is not None:
            if not training:
                raise _lib.RssError("pre_bias folding is only valid for training-mode BatchNorm")
            pre_bias = _f32(pre_bias)
        y = torch.empty_like(x, memory_format=CL)
        account("bn", None if (have_aff or raw_ready or not training) else x, x, residual, y)     # statistics pass + apply pass
        raw = False
        if have_aff:
            pass
        elif training and world == 1:
            if scratch is None or scratch.numel() < 2 + 2 * C:      # [0] last-block ticket, [2:] accumulators; kernel leaves zeros
                scratch = torch.zeros(2 + 2 * C, device=dev, dtype=torch.float32)
            raw = BN_RAW["on"] or raw_ready
            if raw:
                if not raw_ready:
                    check(lib.rss_bn_stats_raw(_p(x), _p(scratch[2:]), rows, C, dt, _p(running_mean), _p(pre_bias), st), "rss_bn_stats_raw")
                check(lib.rss_bn_act_fwd_raw(_p(x), _p(residual), _p(y), _p(scratch[2:]), _p(scratch), rows, C, act, dt, _p(g), _p(b),
                                             _p(running_mean), _p(running_var), momentum, eps, _p(aff[0]), _p(aff[1]), _p(aff[2]),
                                             _p(aff[3]), _p(pre_bias), st), "rss_bn_act_fwd_raw")
        if raw:
            pass
        elif have_aff:
            pass
        elif training and world == 1:
            check(lib.rss_bn_stats_fused(_p(x), _p(scratch[2:]), _p(scratch), rows, C, dt, _p(g), _p(b),
                                         _p(running_mean), _p(running_var), momentum, eps, _p(aff[0]), _p(aff[1]), _p(aff[2]),
                                         _p(aff[3]), _p(pre_bias), st), "rss_bn_stats_fused")
        elif (training and world > 1 and scratch is not None and scratch.numel() >= 2 + 2 * C and sync_exchange(group, dev, C) is not None
              and sync_exchange(group, dev, C).channel(scratch) is not None):
            # SyncBN: raw local sums -> one-shot exchange over NVLink peer memory, finalised by the same one-CTA kernel
            ex = sync_exchange(group, dev, C)
            off, cnt = ex.channel(scratch)
            if not raw_ready:
                check(lib.rss_bn_stats_raw(_p(x), _p(scratch[2:]), rows, C, dt, _p(running_mean), _p(pre_bias), st), "rss_bn_stats_raw")
            check(lib.rss_sync_bn_finalize(_p(ex.bases), off, ex.rank, ex.world, cnt, _p(scratch[2:]), C, rows, _p(g), _p(b),
                                           _p(running_mean), _p(running_var), momentum, eps, _p(aff[0]), _p(aff[1]), _p(aff[2]),
                                           _p(aff[3]), _p(pre_bias), st), "rss_sync_bn_finalize")
        elif training:
            nparts = lib.rss_bn_stats_nparts(rows, C)
            part = torch.empty(nparts * C * 2 + nparts, device=dev, dtype=torch.float32)
            cnt = part[nparts * C * 2:]
            check(lib.rss_bn_stats(_p(x), _p(part), _p(cnt), rows, C, dt, st), "rss_bn_stats")
            stat = torch.empty(C * 2 + 1, device=dev, dtype=torch.float32)
            check(lib.rss_bn_combine(_p(part), _p(cnt), nparts, C, _p(stat), _p(stat[C * 2:]), st), "rss_bn_combine")
            if world > 1:     # SyncBN: exchange (mean, M2, count) per rank, Chan-combine again
                gathered = torch.empty(world, C * 2 + 1, device=dev, dtype=torch.float32)
                dist.all_gather_into_tensor(gathered, stat, group=None if group is True else group)
                parts = gathered[:, :C * 2].contiguous()
                cnts = gathered[:, C * 2].contiguous()
                check(lib.rss_bn_combine(_p(parts), _p(cnts), world, C, _p(stat), _p(stat[C * 2:]), st), "rss_bn_combine")
            check(lib.rss_bn_finalize(_p(stat), _p(stat[C * 2:]), _p(g), _p(b), _p(running_mean), _p(running_var),
                                      momentum, eps, C, _p(aff[0]), _p(aff[1]), _p(aff[2]), _p(aff[3]), _p(pre_bias), st), "rss_bn_finalize")
        else:
            check(lib.rss_bn_eval_affine(_p(g), _p(b), _p(running_mean), _p(running_var), eps, C,
                                         _p(aff[0]), _p(aff[1]), _p(aff[2]), _p(aff[3]), st), "rss_bn_eval_affine")
        if not raw:
            check(lib.rss_bn_act_fwd(_p(x), _p(residual), _p(y), _p(aff[2]), _p(aff[3]), rows, C, act, dt, st), "rss_bn_act_fwd")
        ctx.scratch = scratch
        # link: see conv._ConvLib -- the residual gradient is handed to the 1x1 convolution that shares the residual's source tensor
        ctx.link = link if (link is not None and link.get("armed") and residual is not None) else None
        ctx.save_for_backward(x, y if (act == _lib.ACT_RELU and residual is not None) else None, aff)
        ctx.act, ctx.training, ctx.group, ctx.world, ctx.has_res = act, training, group, world, residual is not None
        ctx.refs = (gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        out = BNAct._backward(ctx, dy) + (None,)
        if ctx.link is not None and out[1] is not None:
            ctx.link["dres"] = out[1]
            out = (out[0], None) + out[2:]
        return out

    @staticmethod
    def _backward(ctx, dy):
        lib = _lib.load()
        x, y, aff = ctx.saved_tensors
        dy = nhwc(dy)
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        B, C, H, W = x.shape
        rows = B * H * W
        dt, st = _dt(x), _st()
        sg, sb = grad_sink(ctx.refs[0]), grad_sink(ctx.refs[1])
        direct = sg is not None and sb is not None
        dx = torch.empty_like(x, memory_format=CL)
        dres = torch.empty_like(x, memory_format=CL) if ctx.has_res else None
        account("bn", x, dy, y, x, dy, y, dx, dres)                                   # reduce pass + apply pass
        sums = torch.empty(2 * C, device=x.device, dtype=torch.float32)
        sc = ctx.scratch        # the layer's persistent zeroed scratch (forward statistics kernel): no memset node per launch
        have_sc = sc is not None and sc.numel() >= 2 + 2 * C
        # GELU layers: the reduce pass keeps dz = dy*gelu'(z) so the apply pass does not pay for the derivative a second time
        dz = torch.empty_like(x, memory_format=CL) if (ctx.act == _lib.ACT_GELU and BN_KEEP_DZ["on"]) else None
        if BN_RAW["on"] and have_sc and dz is None and ctx.training and ctx.world == 1:
            # raw protocol: totals stay in the scratch, the apply kernel reads them there and its last block does the bookkeeping
            check(lib.rss_bn_bwd_reduce_ws(_p(x), _p(y), _p(dy), _p(aff[2]), _p(aff[3]), _p(aff[0]), _p(aff[1]), None,
                                           _p(sc[2:]), None, None, rows, C, ctx.act, dt, st), "rss_bn_bwd_reduce")
            check(lib.rss_bn_bwd_apply_raw(_p(x), _p(y), _p(dy), _p(aff[2]), _p(aff[3]), _p(aff[0]), _p(aff[1]), _p(sc[2:]), _p(sc),
                                           1.0 / rows, _p(dx), _p(dres), rows, C, ctx.act, dt, None if direct else _p(sums),
                                           _p(sg) if direct else None, _p(sb) if direct else None, st), "rss_bn_bwd_apply")
            if direct:
                return (dx, dres) + (None,) * 12
            return (dx, dres, sums[C:], sums[:C]) + (None,) * 10
        ex = sync_exchange(ctx.group, x.device, C) if (ctx.training and ctx.world > 1 and have_sc) else None
        if ex is not None and ex.channel(sc) is None:
            ex = None
        if ex is not None:          # SyncBN: local sums stay in the scratch (no ticket), exchanged over NVLink peer memory
            check(lib.rss_bn_bwd_reduce_ws(_p(x), _p(y), _p(dy), _p(aff[2]), _p(aff[3]), _p(aff[0]), _p(aff[1]), None,
                                           _p(sc[2:]), None, _p(dz), rows, C, ctx.act, dt, st), "rss_bn_bwd_reduce")
            local = torch.empty(2 * C, device=x.device, dtype=torch.float32)
            off, cnt = ex.channel(sc)
            check(lib.rss_sync_allreduce_small(_p(ex.bases), off, ex.rank, ex.world, cnt, _p(sc[2:]), 2 * C, _p(sums), _p(local), st),
                  "rss_sync_allreduce_small")
        else:
            check(lib.rss_bn_bwd_reduce_ws(_p(x), _p(y), _p(dy), _p(aff[2]), _p(aff[3]), _p(aff[0]), _p(aff[1]), _p(sums),
                                           _p(sc[2:]) if have_sc else None, _p(sc) if have_sc else None, _p(dz),
                                           rows, C, ctx.act, dt, st), "rss_bn_bwd_reduce")
            local = sums
        if ctx.training:
            red = sums
            if ctx.world > 1 and ex is None:       # SyncBN: dx needs the global sums, the parameter gradients the local ones (DDP averages them)
                local = sums.clone()
                dist.all_reduce(red, group=None if ctx.group is True else ctx.group)
            inv_count = 1.0 / (rows * ctx.world)
        else:                       # eval-mode BN is a fixed affine map: no batch-statistic terms
            red = torch.zeros_like(sums)
            inv_count = 0.0
        if dz is not None:
            check(lib.rss_bn_bwd_apply_dz(_p(x), _p(dz), _p(aff[2]), _p(aff[0]), _p(aff[1]), _p(red), inv_count, _p(dx), rows, C, dt,
                                          _p(local), _p(sg) if direct else None, _p(sb) if direct else None, st), "rss_bn_bwd_apply_dz")
        else:
            check(lib.rss_bn_bwd_apply(_p(x), _p(y), _p(dy), _p(aff[2]), _p(aff[3]), _p(aff[0]), _p(aff[1]), _p(red), inv_count,
                                       _p(dx), _p(dres), rows, C, ctx.act, dt, _p(local), _p(sg) if direct else None,
                                       _p(sb) if direct else None, st), "rss_bn_bwd_apply")
        if direct:
            return (dx, dres) + (None,) * 12
        return (dx, dres, local[C:], local[:C]) + (None,) * 10


# ----------------------------------------------------------------------------------------------
# multi-resolution fuse sum (+ReLU)
# ----------------------------------------------------------------------------------------------
class FuseSum(torch.autograd.Function):
    """out = [relu](sum_j nearest_up_{2^k_j}(term_j)); term j has spatial size (H >> k_j, W >> k_j)."""

    @staticmethod
    def forward(ctx, relu, ks, *terms):
        _lib.require_device()
        lib = _lib.load()
        terms = [nhwc(t) for t in terms]
        dt0 = terms[0].dtype
        terms = [t if t.dtype == dt0 else t.to(dt0) for t in terms]
        n = len(terms)
        B, C, h0, w0 = terms[0].shape
        H, W = h0 << ks[0], w0 << ks[0]
        out = torch.empty((B, C, H, W), device=terms[0].device, dtype=dt0, memory_format=CL)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in terms])
        kk = (ctypes.c_int * n)(*ks)
        check(lib.rss_fuse_sum_fwd(ptrs, kk, n, _p(out), B, H, W, C, int(relu), _dt(out), _st()), "rss_fuse_sum_fwd")
        ctx.save_for_backward(out if relu else None)
        ctx.cfg = (relu, ks, [tuple(t.shape) for t in terms])
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        (out,) = ctx.saved_tensors
        relu, ks, shapes = ctx.cfg
        dout = nhwc(dout)
        B, C, H, W = dout.shape
        grads = []
        shared = None
        for j, (k, shp) in enumerate(zip(ks, shapes)):
            if not ctx.needs_input_grad[2 + j]:
                grads.append(None)
                continue
            if k == 0 and not relu:
                grads.append(dout)               # identity term of an un-activated sum: the gradient passes through
                continue
            if k == 0 and shared is not None:
                grads.append(shared)             # all full-resolution terms share the same masked gradient
                continue
            d = torch.empty(shp, device=dout.device, dtype=dout.dtype, memory_format=CL)
            check(lib.rss_fuse_sum_bwd(_p(dout), _p(out), _p(d), k, B, H, W, C, int(relu), _dt(dout), _st()), "rss_fuse_sum_bwd")
            if k == 0:
                shared = d
            grads.append(d)
        return (None, None) + tuple(grads)


def spatial_attention(x, conv_w):
    """SpatialAttention.forward alone: sigmoid(conv7x7([mean_c(x), max_c(x)])) -> (B,1,H,W) fp32.  Forward only."""
    _lib.require_device()
    lib = _lib.load()
    x = x.detach().contiguous()                          # NCHW-contiguous: the layout the reference module sees
    B, C, H, W = x.shape
    if C != 32:
        raise _lib.RssError("the RSSFormer gate kernels are built for 32 channels (got %d)" % C)
    HW = H * W
    dev = x.device
    out = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32)
    pooled = torch.empty(B, 4, HW, device=dev, dtype=torch.float32)
    amax = torch.empty(B, 2, HW, device=dev, dtype=torch.uint8)
    maps = torch.empty(2, B, 2, HW, device=dev, dtype=torch.float32)
    check(lib.rss_spatial_attention_fwd(_p(x), _p(_f32(conv_w)), _p(out), _p(pooled), _p(amax), _p(maps), B, H, W, _dt(x), _st()),
          "rss_spatial_attention_fwd")
    return out


def fuse_sum(terms, ks, relu):
    return FuseSum.apply(bool(relu), tuple(int(k) for k in ks), *terms)


# ----------------------------------------------------------------------------------------------
# neck gather (3x bilinear up + concat), head, headaux, loss
# ----------------------------------------------------------------------------------------------
def _geom(feats):
    n = len(feats)
    C = (ctypes.c_int * 4)(*[f.shape[1] for f in feats])
    h = (ctypes.c_int * 4)(*[f.shape[2] for f in feats])
    w = (ctypes.c_int * 4)(*[f.shape[3] for f in feats])
    assert n == 4
    return C, h, w


class NeckGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f0, f1, f2, f3):
        _lib.require_device()
        lib = _lib.load()
        feats = [nhwc(f) for f in (f0, f1, f2, f3)]
        feats = [f if f.dtype == feats[0].dtype else f.to(feats[0].dtype) for f in feats]
        B = feats[0].shape[0]
        C, h, w = _geom(feats)
        out = torch.empty((B, sum(f.shape[1] for f in feats), feats[0].shape[2], feats[0].shape[3]),
                          device=f0.device, dtype=feats[0].dtype, memory_format=CL)
        check(lib.rss_neck_gather_fwd(*[_p(f) for f in feats], _p(out), B, C, h, w, _dt(out), _st()), "rss_neck_gather_fwd")
        ctx.shapes = [tuple(f.shape) for f in feats]
        return out

    @staticmethod
    def backward(ctx, dcat):
        lib = _lib.load()
        dcat = nhwc(dcat)
        ds = [torch.empty(s, device=dcat.device, dtype=dcat.dtype, memory_format=CL) for s in ctx.shapes]
        C, h, w = _geom(ds)
        check(lib.rss_neck_gather_bwd(_p(dcat), *[_p(d) for d in ds], ctx.shapes[0][0], C, h, w, _dt(dcat), _st()),
              "rss_neck_gather_bwd")
        return tuple(ds)


# HRNet stem conv1 from the planar image batch (csrc/stem.cu).  RSS_STEM=0: library path (cast + permute + padded legacy conv);
# RSS_STEM_STATS=0: bn1's statistics from the separate pass instead of the conv epilogue.
STEM = {"on": os.environ.get("RSS_STEM", "1") != "0", "stats": os.environ.get("RSS_STEM_STATS", "1") != "0"}


def stem_conv_ok(x, weight):
    """True when StemConv covers this call: planar (B,3,H,W) fp32/bf16 CUDA input that needs no gradient, (64,3,3,3) weight."""
    return (STEM["on"] and x.is_cuda and x.dim() == 4 and x.shape[1] == 3 and x.dtype in (torch.float32, torch.bfloat16)
            and x.is_contiguous() and not x.requires_grad and tuple(weight.shape) == (64, 3, 3, 3))


class StemConv(torch.autograd.Function):
    """_hrnet_rssformer.py:467 conv1 = Conv2d(3,64,3,stride 2,padding 1,bias=False) on the image batch as the reference model
    receives it ((B,3,H,W) planar, fp32 or bf16) -> (B,64,Ho,Wo) bf16 channels_last.  stats = (scratch, running_mean) of the
    BatchNorm that follows: its raw sums come out of the same launch (call the BatchNorm with aff=RAW_SUMS)."""

    @staticmethod
    def forward(ctx, x, weight, stats):
        _lib.require_device()
        lib = _lib.load()
        if not stem_conv_ok(x, weight):
            raise _lib.RssError("StemConv: planar (B,3,H,W) fp32/bf16 input without gradient and a (64,3,3,3) weight expected")
        B, _, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((B, 64, Ho, Wo), device=x.device, dtype=torch.bfloat16, memory_format=CL)
        acc = shift = None
        if stats is not None:
            scratch, shift = stats
            if scratch.numel() < 2 + 2 * 64:
                raise _lib.RssError("StemConv: BatchNorm scratch too small for the statistics epilogue")
            acc = scratch[2:]
        account("stem", x, y)
        check(lib.rss_stem_conv_fwd(_p(x), _p(_f32(weight)), _p(y), B, H, W, _dt(x), _p(acc), _p(shift), _st()), "rss_stem_conv_fwd")
        ctx.save_for_backward(x)
        ctx.refs = (weight,)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, = ctx.saved_tensors
        weight = ctx.refs[0]
        if not ctx.needs_input_grad[1]:
            return None, None, None
        dy = nhwc(dy)
        if dy.dtype != torch.bfloat16:
            dy = dy.to(torch.bfloat16)
        B, _, H, W = x.shape
        sink = grad_sink(weight)
        dw = sink if sink is not None else torch.zeros(64, 3, 3, 3, device=x.device, dtype=torch.float32)
        account("stem", x, dy)
        check(lib.rss_stem_conv_wgrad(_p(x), _p(dy), _p(dw), B, H, W, _dt(x), _st()), "rss_stem_conv_wgrad")
        return None, (None if sink is not None else dw.to(weight.dtype)), None


class HeadConv(torch.autograd.Function):
    """(B,C,h,w) channels_last -> low-resolution logits (B,h,w,8) fp32 (class 7 = padding)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_device()
        lib = _lib.load()
        x = nhwc(x)
        B, C, h, w = x.shape
        wt = _f32(weight).reshape(7, C)
        logits = torch.empty(B, h, w, 8, device=x.device, dtype=torch.float32)
        check(lib.rss_head_fwd(_p(x), _p(wt), _p(_f32(bias)), _p(logits), B * h * w, C, _dt(x), _st()), "rss_head_fwd")
        ctx.save_for_backward(x, wt)
        ctx.wshape = tuple(weight.shape)
        ctx.refs = (weight, bias)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        lib = _lib.load()
        x, wt = ctx.saved_tensors
        B, C, h, w = x.shape
        dlogits = dlogits.contiguous()
        dx = torch.empty_like(x, memory_format=CL)
        sw, sb = grad_sink(ctx.refs[0]), grad_sink(ctx.refs[1])
        direct = sw is not None and sb is not None
        dw = sw if direct else torch.zeros(7, C, device=x.device, dtype=torch.float32)
        db = sb if direct else torch.zeros(7, device=x.device, dtype=torch.float32)
        check(lib.rss_head_bwd(_p(x), _p(dlogits), _p(wt), _p(dx), _p(dw), _p(db), B * h * w, C, _dt(x), _st()), "rss_head_bwd")
        return (dx, None, None) if direct else (dx, dw.reshape(ctx.wshape), db)


def head_probs(logits_lr, scale, want_argmax=False):
    """eval output: softmax over classes of the x`scale` bilinear (align_corners=True) up-sampling; NCHW fp32."""
    lib = _lib.load()
    B, h, w, _ = logits_lr.shape
    probs = torch.empty(B, 7, h * scale, w * scale, device=logits_lr.device, dtype=torch.float32)
    am = torch.empty(B, h * scale, w * scale, device=logits_lr.device, dtype=torch.uint8) if want_argmax else None
    check(lib.rss_head_probs(_p(logits_lr), _p(probs), _p(am), B, h, w, scale, _st()), "rss_head_probs")
    return (probs, am) if want_argmax else probs


def headaux(f0, weight, bias):
    """AdaptiveAvgPool2d(1) + Linear(C,7); no gradient (the reference consumes it under no_grad)."""
    _lib.require_device()
    lib = _lib.load()
    f0 = nhwc(f0.detach())
    B, C, H, W = f0.shape
    ws = torch.empty(B * C, device=f0.device, dtype=torch.float32)
    scores = torch.empty(B, 7, device=f0.device, dtype=torch.float32)
    check(lib.rss_headaux_fwd(_p(f0), _p(_f32(weight)), _p(_f32(bias)), _p(ws), _p(scores), B, H * W, C, _dt(f0), _st()),
          "rss_headaux_fwd")
    return scores


class SegLoss(torch.autograd.Function):
    """fc_loss of SegmentationLossaux on LOW-resolution logits (the x`scale` up-sampling is fused)."""

    @staticmethod
    def forward(ctx, logits_lr, labels, aux_scores, scale, ignore_index):
        lib = _lib.load()
        B, h, w, _ = logits_lr.shape
        logits_lr = logits_lr.contiguous()
        labels = labels.contiguous()
        if labels.dtype != torch.int64:
            labels = labels.long()
        assert labels.shape == (B, h * scale, w * scale), (labels.shape, (B, h * scale, w * scale))
        dev = logits_lr.device
        acc = torch.empty(lib.rss_seg_loss_acc_floats(B), device=dev, dtype=torch.float32)
        gdir = torch.empty(B, h, w, 8, device=dev, dtype=torch.float32)
        out4 = torch.empty(4, device=dev, dtype=torch.float32)
        check(lib.rss_seg_loss_fwd(_p(logits_lr), _p(labels), _p(aux_scores.contiguous()), _p(acc), _p(gdir), _p(out4),
                                   B, h, w, scale, ignore_index, _st()), "rss_seg_loss_fwd")
        ctx.save_for_backward(gdir, out4)
        return out4[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        lib = _lib.load()
        gdir, out4 = ctx.saved_tensors
        B, h, w, _ = gdir.shape
        up = dloss.detach().float().reshape(1).contiguous()
        dl = torch.empty_like(gdir)
        check(lib.rss_seg_loss_bwd(_p(gdir), _p(out4), _p(up), _p(dl), B, h, w, _st()), "rss_seg_loss_bwd")
        return dl, None, None, None, None


# ----------------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------------
def grad_sumsq(flat_grad, grad_scale, out):
    check(_lib.load().rss_grad_sumsq(_p(flat_grad), flat_grad.numel(), grad_scale, _p(out), _st()), "rss_grad_sumsq")


def sgd_step(flat_p, flat_g, flat_m, sumsq, grad_scale, max_norm, lr_dev, momentum, weight_decay, zero_grad, shadow=None):
    check(_lib.load().rss_sgd_step(_p(flat_p), _p(flat_g), _p(flat_m), flat_p.numel(), _p(sumsq), grad_scale, max_norm, _p(lr_dev),
                                   momentum, weight_decay, int(zero_grad), _p(shadow), _st()), "rss_sgd_step")


def shadow_cl_refresh(flat_p, shadow_cl, table, row_start, n_entries, max_row_floats):
    check(_lib.load().rss_shadow_cl_refresh(_p(flat_p), _p(shadow_cl), _p(table), _p(row_start), n_entries, max_row_floats, _st()),
          "rss_shadow_cl_refresh")


def shadow_t_refresh(flat_p, shadow_t, table, n_entries):
    check(_lib.load().rss_shadow_t_refresh(_p(flat_p), _p(shadow_t), _p(table), n_entries, _st()), "rss_shadow_t_refresh")
