"""Convolution dispatch for the RSSFormer path (NHWC activations, fp32 master weights).

Two engines, chosen per layer shape (DESIGN.md lists which layer uses which):
  * `igemm`  — the hand-written tcgen05/TMA implicit-GEMM kernels of csrc/conv_igemm.cu (bf16 operands,
               fp32 TMEM accumulation) through the C ABI;
  * `lib`    — ATen's convolution (cuDNN) for the shapes the igemm kernels do not cover yet.
Both take the low-precision copy of the weight from the optimiser's bf16 shadow buffer when one is
registered (no per-call cast kernels), and return fp32 weight gradients to the master parameter.
"""
import ctypes
import os

import torch

from . import _lib, ops

CL = torch.channels_last
# per-shape engine policy from tools/conv_microbench.py (profiles/conv_microbench_r1.json): the igemm kernel runs the
# FFN's fused 17-tap conv (forward + data gradient); single convs stay on the library until the halo-reuse version lands.
ENGINE = {"igemm": True, "igemm_single": False}

# Weight gradients are needed only by the optimiser, data gradients by the next backward node: wgrad kernels go to a side
# stream so they overlap the (latency-bound) rest of the backward chain; trainer joins the stream before the all-reduce.
WGRAD = {"async": os.environ.get("RSS_WGRAD_STREAM", "1") != "0", "streams": {}, "used": set()}


# side streams used round-robin (measured on the B=16 step: 1 -> 464, 2 -> 493, 3 -> 492 img/s: one stream serialised the weight
# gradients into the longest chain of the backward pass)
WGRAD["n"] = max(1, int(os.environ.get("RSS_WGRAD_STREAMS", "2")))
# RSS_WGRAD_AFTER_DGRAD=1: issue the data gradient (on the critical chain) before forking the weight gradient, so the side stream
# waits for it instead of competing with it for SMs
# (round 2, with the fused branch-0 blocks: 513.4 img/s with it vs 507-510 without -> on by default)
WGRAD["after_dgrad"] = os.environ.get("RSS_WGRAD_AFTER_DGRAD", "1") != "0"


def wgrad_stream(dev):
    lst = WGRAD["streams"].get(dev)
    if lst is None:
        lst = WGRAD["streams"][dev] = [torch.cuda.Stream(dev) for _ in range(WGRAD["n"])]
        WGRAD["rr"] = 0
    WGRAD["rr"] = (WGRAD["rr"] + 1) % len(lst)
    return lst[WGRAD["rr"]]


# bf16 weight gradients of the library kernels waiting for their fp32 accumulation into the flat gradient buffer: one multi-tensor
# launch at join time (rss_accum_bf16_list) instead of one torch elementwise kernel per convolution (RSS_WGRAD_BATCH=0: per conv)
PENDING = {"on": os.environ.get("RSS_WGRAD_BATCH", "1") != "0", "items": [], "tables": {}}


def _flush_pending():
    items = PENDING["items"]
    if not items:
        return
    PENDING["items"] = []
    dev = items[0][0].device
    cur = torch.cuda.current_stream(dev)
    key = tuple((dw.data_ptr(), sink.data_ptr(), dw.numel(), cin, kk) for sink, dw, cin, kk in items)
    ent = PENDING["tables"].get(key)
    if ent is None:
        if len(PENDING["tables"]) > 8:
            PENDING["tables"].clear()
        rows, starts, c = [], [0], 0
        for src, dst, n, cin, kk in key:
            rows.append([src, dst, n, cin, kk])
            c += int(_lib.load().rss_accum_chunks(n, cin, kk))
            starts.append(c)
        host = (torch.tensor(rows, dtype=torch.int64).pin_memory(), torch.tensor(starts, dtype=torch.int64).pin_memory())
        ent = PENDING["tables"][key] = (host, torch.empty_like(host[0], device=dev), torch.empty_like(host[1], device=dev), c)
    host, table, starts, total = ent
    table.copy_(host[0], non_blocking=True)               # (captured as memcpy nodes: the pinned tables outlive the graph)
    starts.copy_(host[1], non_blocking=True)
    for it in items:
        it[1].record_stream(cur)
    ops.check(_lib.load().rss_accum_bf16_list(table.data_ptr(), starts.data_ptr(), len(items), total, ops._st()), "rss_accum_bf16_list")


def join_wgrad(dev=None):
    """make the current stream wait for every outstanding weight-gradient kernel (call before reading gradients)"""
    for d, lst in WGRAD["streams"].items():
        if (dev is None or d == dev) and d in WGRAD["used"]:
            for s in lst:
                torch.cuda.current_stream(d).wait_stream(s)
    WGRAD["used"].clear()
    _flush_pending()


ENGINE["wgrad"] = os.environ.get("RSS_WGRAD_KERNEL", "1") != "0"     # hand-written split-K weight gradient (csrc/conv_wgrad.cu)


# tcgen05 split-K weight gradient (csrc/conv_wgrad_tc.cu): bit-correct, but every M128 x N<=128 x K16 MMA costs ~350-400 cycles of
# operand fetch (profiles/wgrad_microbench_r1.json), which makes it slower than both other engines on these layers: off by default
ENGINE["wgrad_tc"] = os.environ.get("RSS_WGRAD_TC", "0") != "0"
# per-shape policy for the mma.sync kernel from the same microbench (ncu kernel durations, B=16): it beats the library's wgrad +
# the separate fp32 accumulate on the 32->32 3x3 layers of branch 0 (31.7 vs 47.6+8 us) and on the small 1x1 fuse convs
# (10.8 vs 25.6 us); elsewhere the re-reads of X/dY per 32x32 channel block lose against the library kernel
ENGINE["wgrad_policy"] = os.environ.get("RSS_WGRAD_POLICY", "measured")


def _tc_wgrad_ok(x, weight, stride, padding, dilation):
    Cout, Cin, k, _ = weight.shape
    if not ENGINE["wgrad_tc"] or stride != 1 or dilation != 1 or padding != k // 2:
        return False
    B, _, H, W = x.shape
    return bool(_lib.load().rss_conv_wgrad_tc_supported(B, H, W, Cin, Cout, k))


def _own_wgrad_ok(dy, x, weight, want_b, stride, padding, dilation):
    if not ENGINE["wgrad"] or want_b or x.dtype != torch.bfloat16 or not x.is_cuda:
        return False
    Cout, Cin, k, _ = weight.shape
    if _tc_wgrad_ok(x, weight, stride, padding, dilation):
        return True
    if ENGINE["wgrad_policy"] == "measured":
        if not ((k == 3 and stride == 1 and Cin == 32 and Cout == 32) or (k == 1 and Cin * Cout <= 64 * 32)):
            return False
    return bool(_lib.load().rss_conv_wgrad_supported(Cin, Cout, k, stride, padding, dilation))


def _wgrad(dy, x, w_lp, weight, bias, want_b, stride, padding, dilation, wdtype, make_x=None):
    """make_x: optional callable run on the weight-gradient stream right before the kernel, returning the activation `x` (used by
    the fused BasicBlock, which never materialises relu(bn1(.)) in the forward pass and re-creates it here, off the critical chain)
    weight (+bias) gradient, accumulated into the flat fp32 grad buffer when there is one; returns (dw, db) to hand back
    through autograd (None when already accumulated).  Engine: csrc/conv_wgrad.cu (split-K mma kernel, float atomics straight
    into the fp32 gradient) for the HRNet-family shapes, the library's wgrad otherwise."""
    own = _own_wgrad_ok(dy, x, weight, want_b, stride, padding, dilation)

    def run_own(sink):
        B, Cin, Hi, Wi = x.shape
        _, Cout, Ho, Wo = dy.shape
        if _tc_wgrad_ok(x, weight, stride, padding, dilation):
            ops.check(_lib.load().rss_conv_wgrad_tc(x.data_ptr(), dy.data_ptr(), sink.data_ptr(), B, Hi, Wi, Cin, Cout, weight.shape[2],
                                                    None, None, 0, ops._st()), "rss_conv_wgrad_tc")
            return
        ops.check(_lib.load().rss_conv_wgrad(x.data_ptr(), dy.data_ptr(), sink.data_ptr(), B, Hi, Wi, Cin, Ho, Wo, Cout,
                                             weight.shape[2], stride, padding, dilation, ops._st()), "rss_conv_wgrad")

    def run_lib():
        _, dw, db = torch.ops.aten.convolution_backward(dy, x, w_lp, [w_lp.shape[0]] if want_b else None, [stride, stride],
                                                        [padding, padding], [dilation, dilation], False, [0, 0], 1,
                                                        [False, True, want_b])
        sw = ops.grad_sink(weight)
        if sw is not None:
            kk = dw.shape[2] * dw.shape[3]
            if PENDING["on"] and dw.dtype == torch.bfloat16 and sw.is_contiguous() and (dw.is_contiguous() or dw.is_contiguous(memory_format=CL)):
                # the library returns k x k gradients in (Cout,kh,kw,Cin) memory order; the sink is in parameter order
                PENDING["items"].append((sw, dw, dw.shape[1], 0 if (kk == 1 or dw.is_contiguous()) else kk))
            else:
                sw.add_(dw)
            dw = None
        else:
            dw = dw.to(wdtype)
        if db is not None:
            sb = ops.grad_sink(bias)
            if sb is not None:
                sb.add_(db)
                db = None
            else:
                db = db.to(wdtype)
        return dw, db

    def run():
        nonlocal x
        if make_x is not None:
            x = make_x()
        if not own:
            return run_lib()
        sink = ops.grad_sink(weight)
        if sink is not None:
            run_own(sink)
            return None, None
        dw = torch.zeros(weight.shape, device=x.device, dtype=torch.float32)
        run_own(dw)
        return dw.to(wdtype), None

    direct = ops.grad_sink(weight) is not None and (not want_b or ops.grad_sink(bias) is not None)
    # only parameters owned by trainer.FlatSGD go asynchronous (its step joins the stream); anyone else reads .grad right away
    if not (WGRAD["async"] and direct and dy.is_cuda and getattr(weight, "_rss_flat", False)):
        return run()
    dev = dy.device
    cur, side = torch.cuda.current_stream(dev), wgrad_stream(dev)
    side.wait_stream(cur)
    dy.record_stream(side); x.record_stream(side); w_lp.record_stream(side)
    with torch.cuda.stream(side):
        run()
    WGRAD["used"].add(dev)
    return None, None


def register_shadow(param, view):
    """attach the bf16 view kept fresh by the fused optimiser step (trainer.py) to its fp32 master parameter"""
    param._rss_shadow = view


def register_shadow_cl(param, view):
    """channels-last bf16 view ((Cout,Cin,kh,kw) logical shape, (Cout,kh,kw,Cin) memory) refreshed once per step by trainer.py"""
    param._rss_shadow_cl = view


def _lowp(w, dtype, channels_last=False):
    if w is None or w.dtype == dtype:
        return w
    if channels_last:
        s = getattr(w, "_rss_shadow_cl", None)
        if s is not None and s.dtype == dtype and s.shape == w.shape:
            return s
    s = getattr(w, "_rss_shadow", None)
    if s is not None and s.dtype == dtype and s.numel() == w.numel():
        return s.view(w.shape)
    return w.detach().to(dtype)


def lowp_cl(weight, dtype):
    """channels-last low-precision copy of a conv weight for the library kernels (the per-step shadow when there is one)"""
    return _lowp(weight, dtype, channels_last=True).contiguous(memory_format=CL)


class _ConvLib(torch.autograd.Function):
    """link (optional dict): residual-gradient hand-off inside a Bottleneck.  The block input x feeds this 1x1 convolution AND the
    residual add behind bn3; autograd would sum the two gradients with one elementwise kernel over the 134 MB tensor (62 us x 3 on
    the one-stream tail of the step).  Instead the BatchNorm backward deposits its residual gradient in link['dres'] (returning
    None to autograd) and this node's data gradient is ONE GEMM with beta = 1 into that buffer."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, bias_grad, link=None):
        x = ops.nhwc(x)
        if link is not None:
            if weight.shape[2] == 1 and weight.shape[3] == 1 and stride == 1 and padding == 0 and x.is_cuda and ctx.needs_input_grad[0]:
                link["armed"] = True
            else:
                link = None
        ctx.link = link
        w = _lowp(weight, x.dtype, channels_last=True).contiguous(memory_format=CL)     # no-op on the per-step channels-last shadow
        b = _lowp(bias, x.dtype)
        y = torch.ops.aten.convolution(x, w, b, [stride, stride], [padding, padding], [dilation, dilation], False, [0, 0], 1)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, dilation, bias is not None and bias_grad, weight.dtype)
        ctx.refs = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding, dilation, has_bias, wdtype = ctx.cfg
        dy = ops.nhwc(dy)
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dx = dw = db = None
        first = ctx.needs_input_grad[1] and not WGRAD["after_dgrad"]
        if first:
            dw, db = _wgrad(dy, x, w, ctx.refs[0], ctx.refs[1], has_bias and ctx.needs_input_grad[2], stride, padding, dilation, wdtype)
        if ctx.needs_input_grad[0]:
            acc = ctx.link.pop("dres", None) if ctx.link is not None else None
            if acc is not None and acc.dtype == dy.dtype and acc.shape == x.shape and acc.is_contiguous(memory_format=CL):
                Cout, Cin = w.shape[0], w.shape[1]
                acc.permute(0, 2, 3, 1).reshape(-1, Cin).addmm_(dy.permute(0, 2, 3, 1).reshape(-1, Cout), w.reshape(Cout, Cin))
                dx = acc
            else:
                dx = torch.ops.aten.convolution_backward(dy, x, w, None, [stride, stride], [padding, padding], [dilation, dilation],
                                                         False, [0, 0], 1, [True, False, False])[0]
                if acc is not None:
                    dx = dx + acc.to(dx.dtype)
        if ctx.needs_input_grad[1] and not first:
            dw, db = _wgrad(dy, x, w, ctx.refs[0], ctx.refs[1], has_bias and ctx.needs_input_grad[2], stride, padding, dilation, wdtype)
        return dx, dw, db, None, None, None, None, None


def _igemm_ok(x, Cout):
    if not ENGINE["igemm"] or x.dtype != torch.bfloat16 or not x.is_cuda:
        return False
    B, Cin, H, W = x.shape
    return bool(_lib.load().rss_conv_igemm_supported(B, H, W, Cin, Cout))


def _pack(weights, biases, ksizes, dils, Cout, Cin, transpose, dev):
    """-> (packed bf16 buffer, bias_sum or None, n_taps, dy[], dx[]) through rss_conv_pack_weights"""
    lib = _lib.load()
    n = len(weights)
    ws = [ops._f32(w) for w in weights]
    bs = [None if b is None else ops._f32(b) for b in biases]
    wptr = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
    has_b = any(b is not None for b in bs) and not transpose
    bptr = (ctypes.c_void_p * n)(*[None if b is None else b.data_ptr() for b in bs]) if has_b else None
    ks = (ctypes.c_int * n)(*ksizes)
    ds = (ctypes.c_int * n)(*dils)
    packed = torch.empty(lib.rss_conv_packed_bytes(n, ks, Cout, Cin) // 2, device=dev, dtype=torch.bfloat16)
    bias_sum = torch.empty(Cout, device=dev, dtype=torch.float32) if has_b else None
    nt = ctypes.c_int(0)
    dy = (ctypes.c_int * 32)()
    dx = (ctypes.c_int * 32)()
    ops.check(lib.rss_conv_pack_weights(wptr, bptr, ks, ds, n, Cout, Cin, int(transpose), packed.data_ptr(),
                                        None if bias_sum is None else bias_sum.data_ptr(), ctypes.byref(nt), dy, dx, ops._st()),
              "rss_conv_pack_weights")
    return packed, bias_sum, nt.value, dy, dx, (ws, bs)


def prepack_sum(x_like, convs, want_bwd=True):
    """Both weight operands of conv_sum's igemm launch (forward pack; transposed pack for the data gradient) produced NOW on a side
    stream, for a convolution that runs later on the current stream: the two rss_conv_pack_weights launches (~10 us each) leave the
    critical chain (a transformer block calls this before its attention half).  -> opaque dict for conv_sum(..., prepacked=) or None
    when the igemm kernel does not cover the geometry.  x_like: a tensor with the conv input's batch / spatial shape and dtype."""
    Cout, Cin = convs[0][0].shape[0], convs[0][0].shape[1]
    B, _, H, W = x_like.shape
    if not (ENGINE["igemm"] and x_like.dtype == torch.bfloat16 and x_like.is_cuda and os.environ.get("RSS_IGEMM_PREPACK", "1") != "0"
            and bool(_lib.load().rss_conv_igemm_supported(B, H, W, Cin, Cout))):
        return None
    dev = x_like.device
    weights, biases = [c[0] for c in convs], [c[1] for c in convs]
    ksizes, dils = [c[2] for c in convs], [c[3] for c in convs]
    cur, side = torch.cuda.current_stream(dev), wgrad_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        fwd = _pack(weights, biases, ksizes, dils, Cout, Cin, False, dev)
        bwd = _pack(weights, [None] * len(convs), ksizes, dils, Cout, Cin, True, dev) if want_bwd else None
        ev = torch.cuda.Event()
        ev.record(side)
    for pk in (fwd, bwd):
        if pk is not None:
            pk[0].record_stream(cur)
            if pk[1] is not None:
                pk[1].record_stream(cur)
    WGRAD["used"].add(dev)
    return {"fwd": fwd, "bwd": bwd, "event": ev, "ids": tuple(id(w) for w in weights)}


class _ConvIgemm(torch.autograd.Function):
    """sum_s conv2d(x, w_s, b_s, stride 1, padding = dil_s*(k_s//2), dilation dil_s) as ONE tcgen05 implicit GEMM.
    Data gradient: same kernel on the transposed pack.  Weight gradient: library kernel per source (for now)."""

    @staticmethod
    def forward(ctx, x, cfg, *wb):
        ksizes, dils, bias_grad = cfg[:3]
        stats = cfg[3] if len(cfg) > 3 else None         # (scratch, running_mean): BatchNorm raw sums from the GEMM's epilogue
        pre = cfg[4] if len(cfg) > 4 else None           # prepack_sum(): operands packed earlier on a side stream
        n = len(ksizes)
        weights, biases = wb[:n], wb[n:]
        lib = _lib.load()
        x = ops.nhwc(x)
        B, Cin, H, W = x.shape
        Cout = weights[0].shape[0]
        if pre is not None and pre["ids"] == tuple(id(w) for w in weights):
            torch.cuda.current_stream(x.device).wait_event(pre["event"])
            packed, bias_sum, nt, dy, dx, keep = pre["fwd"]
            ctx.pre_bwd = pre["bwd"]
        else:
            packed, bias_sum, nt, dy, dx, keep = _pack(weights, biases, ksizes, dils, Cout, Cin, False, x.device)
            ctx.pre_bwd = None
        y = torch.empty((B, Cout, H, W), device=x.device, dtype=x.dtype, memory_format=CL)
        with ops.timed("rss_conv_igemm"):
            if stats is not None:
                scratch, shift = stats
                ops.check(lib.rss_conv_igemm_stats(x.data_ptr(), packed.data_ptr(), None if bias_sum is None else bias_sum.data_ptr(),
                                                   y.data_ptr(), B, H, W, Cin, Cout, nt, dy, dx, scratch[2:].data_ptr(), ops._p(shift),
                                                   None, ops._st()), "rss_conv_igemm_stats")
            else:
                ops.check(lib.rss_conv_igemm(x.data_ptr(), packed.data_ptr(), None if bias_sum is None else bias_sum.data_ptr(),
                                             y.data_ptr(), B, H, W, Cin, Cout, nt, dy, dx, ops._st()), "rss_conv_igemm")
        ctx.save_for_backward(x)
        ctx.cfg, ctx.refs, ctx.n = cfg[:3], wb, n
        return y

    @staticmethod
    def backward(ctx, dy_):
        (x,) = ctx.saved_tensors
        ksizes, dils, bias_grad = ctx.cfg
        n = ctx.n
        weights, biases = ctx.refs[:n], ctx.refs[n:]
        lib = _lib.load()
        dy_ = ops.nhwc(dy_)
        if dy_.dtype != x.dtype:
            dy_ = dy_.to(x.dtype)
        B, Cin, H, W = x.shape
        Cout = weights[0].shape[0]
        gw, gb = [], []
        for s in range(n):                         # weight gradients first: they start on the side stream while dgrad runs here
            k, d = ksizes[s], dils[s]
            if not ctx.needs_input_grad[2 + s]:
                gw.append(None); gb.append(None)
                continue
            w_lp = _lowp(weights[s], x.dtype, channels_last=True).contiguous(memory_format=CL)
            want_b = biases[s] is not None and bias_grad
            dw, db = _wgrad(dy_, x, w_lp, weights[s], biases[s], want_b, 1, d * (k // 2), d, weights[s].dtype)
            gw.append(dw); gb.append(db)
        dx = None
        if ctx.needs_input_grad[0]:
            packed, _, nt, tdy, tdx, keep = ctx.pre_bwd if ctx.pre_bwd is not None else _pack(weights, [None] * n, ksizes, dils, Cout, Cin,
                                                                                            True, x.device)
            dx = torch.empty_like(x, memory_format=CL)
            with ops.timed("rss_conv_igemm"):
                ops.check(lib.rss_conv_igemm(dy_.data_ptr(), packed.data_ptr(), None, dx.data_ptr(), B, H, W, Cout, Cin, nt, tdy, tdx,
                                             ops._st()), "rss_conv_igemm")
        return (dx, None) + tuple(gw) + tuple(gb)


ENGINE["cf"] = os.environ.get("RSS_CONV_CF", "0") != "0"      # fused tcgen05 conv (+BN statistics epilogue): csrc/conv_cf.cu
# (bit-correct, but its cp.async producer is still slower than the library conv + separate statistics kernel: off by default)


def _cf_ok(x, Cout, k, mode):
    """mode: False/True (plain / statistics epilogue, as conv_bn_stats passes it) or one of _lib.CF_*"""
    if not ENGINE["cf"] or x.dtype != torch.bfloat16 or not x.is_cuda:
        return False
    B, Cin, H, W = x.shape
    return bool(_lib.load().rss_conv_cf_supported(B, H, W, Cin, Cout, k, int(mode)))


def _cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, in_aff, in_relu, stats, add=None, bnred=None, wstrides=(0, 0)):
    """one rss_conv_cf launch -> (y, extra).
    in_aff: (4,Cin) mean/invstd/scale/shift of the BatchNorm whose (+ReLU) output this conv consumes (applied on load) or None
    stats = (gamma, beta, running_mean, running_var, momentum, eps, scratch): statistics epilogue, extra = (4,Cout) affine
    add: tensor of y's shape added in the epilogue (residual / residual-path gradient)
    bnred = (z, out_or_None, aff(4,Cout), relu, scratch): BatchNorm-backward reduction epilogue, y = masked gradient,
            extra = sums (2*Cout)"""
    lib = _lib.load()
    B, _, H, W = x.shape
    y = torch.empty((B, Cout, H, W), device=x.device, dtype=x.dtype, memory_format=CL)
    extra = None
    ep = None
    keep = []
    if stats is not None or add is not None or bnred is not None:
        ep = _lib.ConvCfEpilogue()
        ep.mode = _lib.CF_PLAIN
        if add is not None:
            add = ops.nhwc(add)
            assert add.shape == y.shape and add.dtype == y.dtype
            ep.add = add.data_ptr()
        if stats is not None:
            gamma, beta, rm, rv, mom, eps, scratch = stats
            extra = torch.empty(4, Cout, device=x.device, dtype=torch.float32)
            g, b = ops._f32(gamma), ops._f32(beta)
            keep += [g, b]
            ep.mode = _lib.CF_STATS
            ep.accum, ep.ticket = scratch[2:].data_ptr(), scratch.data_ptr()
            ep.gamma, ep.beta, ep.running_mean, ep.running_var = g.data_ptr(), b.data_ptr(), ops._p(rm), ops._p(rv)
            ep.momentum, ep.eps = float(mom), float(eps)
            ep.mean_out, ep.invstd_out, ep.scale_out, ep.shift_out = (extra[i].data_ptr() for i in range(4))
        elif bnred is not None:
            z, out, aff, relu, scratch = bnred
            assert z.shape == y.shape and z.dtype == y.dtype and z.is_contiguous(memory_format=CL)
            extra = torch.empty(2 * Cout, device=x.device, dtype=torch.float32)
            ep.mode = _lib.CF_BNRED
            ep.accum, ep.ticket = scratch[2:].data_ptr(), scratch.data_ptr()
            ep.bn_z, ep.bn_out = z.data_ptr(), ops._p(out)
            ep.bn_mean, ep.bn_invstd, ep.bn_scale, ep.bn_shift = (aff[i].data_ptr() for i in range(4))
            ep.bn_relu = int(relu)
            ep.sums_out = extra.data_ptr()
    isc = ish = None
    if in_aff is not None:
        isc, ish = in_aff[2].data_ptr(), in_aff[3].data_ptr()
    ops.account("cf", x, y, add, None if bnred is None else bnred[0], None if bnred is None else bnred[1])
    with ops.timed("rss_conv_cf"):
        ops.check(lib.rss_conv_cf(x.data_ptr(), packed.data_ptr(), y.data_ptr(), B, H, W, Cin, Cout, nt, tdy, tdx,
                                  int(wstrides[0]), int(wstrides[1]), isc, ish,
                                  int(in_relu), None if ep is None else ctypes.byref(ep), ops._st()), "rss_conv_cf")
    return y, extra


CF_SQUARE_3X3 = (32, 64)         # channel counts of the C -> C 3x3 layers conv_cf.cu instantiates with all three epilogues
_TAPS3 = {}


def _taps3(sign):
    """ctypes (dy[], dx[]) of the 3x3 taps in natural order t = ky*3 + kx; sign = -1: the data-gradient offsets"""
    if sign not in _TAPS3:
        _TAPS3[sign] = ((ctypes.c_int * 9)(*[sign * (t // 3 - 1) for t in range(9)]), (ctypes.c_int * 9)(*[sign * (t % 3 - 1) for t in range(9)]))
    return _TAPS3[sign]


def cf_weight(weight, transpose):
    """weight operand of rss_conv_cf for a (C,C,3,3) parameter -> (tensor, n_taps, dy[], dx[], (row_stride, tap_stride), keepalive).
    With trainer.FlatSGD the per-step shadows are used in place (channels-last copy for the forward operand, transposed copy for
    the data gradient); otherwise the weight is packed on the fly."""
    Cout, Cin, k, _ = weight.shape
    if k == 3:
        if not transpose:
            s = getattr(weight, "_rss_shadow_cl", None)
            if s is not None and s.dtype == torch.bfloat16:       # memory [Cout][tap][Cin]
                return s, 9, *_taps3(1), (9 * Cin, Cin), None
        else:
            s = getattr(weight, "_rss_shadow_t", None)
            if s is not None:                                     # memory [Cin][tap][Cout]: N = Cin rows, K = Cout contiguous
                return s, 9, *_taps3(-1), (9 * Cout, Cout), None
    if transpose:
        packed, _, nt, tdy, tdx, keep = _pack([weight], [None], [k], [1], Cout, Cin, True, weight.device)
    else:
        packed, _, nt, tdy, tdx, keep = _pack([weight], [None], [k], [1], Cout, Cin, False, weight.device)
    return packed, nt, tdy, tdx, (0, 0), keep


class _ConvCF(torch.autograd.Function):
    """stride-1 k x k (k in {1,3}) convolution through csrc/conv_cf.cu; with `stats` the kernel's epilogue also produces the
    training-mode BatchNorm affine of the output (returned as a non-differentiable (4,Cout) tensor for ops.BNAct).
    Backward: data gradient through the same kernel (transposed pack) when the swapped geometry is supported, weight gradient
    through csrc/conv_wgrad.cu."""

    @staticmethod
    def forward(ctx, x, weight, stats):
        x = ops.nhwc(x)
        B, Cin, H, W = x.shape
        Cout, _, k, _ = weight.shape
        packed, _, nt, tdy, tdx, keep = _pack([weight], [None], [k], [1], Cout, Cin, False, x.device)
        y, aff = _cf_launch(x, packed, nt, tdy, tdx, Cin, Cout, None, False, stats)
        ctx.save_for_backward(x)
        ctx.k, ctx.ref = k, weight
        if aff is None:
            return y
        ctx.mark_non_differentiable(aff)
        return y, aff

    @staticmethod
    def backward(ctx, dy, *unused):
        (x,) = ctx.saved_tensors
        weight, k = ctx.ref, ctx.k
        dy = ops.nhwc(dy)
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        B, Cin, H, W = x.shape
        Cout = weight.shape[0]
        dw = None
        w_lp = None
        if ctx.needs_input_grad[1]:
            if not _own_wgrad_ok(dy, x, weight, False, 1, k // 2, 1):
                w_lp = _lowp(weight, x.dtype, channels_last=True).contiguous(memory_format=CL)
            dw, _ = _wgrad(dy, x, w_lp if w_lp is not None else x, weight, None, False, 1, k // 2, 1, weight.dtype)
        dx = None
        if ctx.needs_input_grad[0]:
            if _cf_ok(dy, Cin, k, False):
                packed, _, nt, tdy, tdx, keep = _pack([weight], [None], [k], [1], Cout, Cin, True, x.device)
                dx, _ = _cf_launch(dy, packed, nt, tdy, tdx, Cout, Cin, None, False, None)
            else:
                w_lp = w_lp if w_lp is not None else _lowp(weight, x.dtype, channels_last=True).contiguous(memory_format=CL)
                dx = torch.ops.aten.convolution_backward(dy, x, w_lp, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1,
                                                         [True, False, False])[0]
        return dx, dw, None


def conv_bn_stats(x, weight, stride, padding, dilation, bn_stats, link=None):
    """conv (no bias) whose epilogue also yields the BatchNorm statistics of its output.  bn_stats = (gamma, beta, running_mean,
    running_var, momentum, eps, scratch) or None.  Returns (y, aff) with aff None when the statistics still have to be computed by
    the caller (library conv, unsupported geometry)."""
    k = weight.shape[2]
    if stride == 1 and dilation == 1 and padding == k // 2 and k in (1, 3) and _cf_ok(x, weight.shape[0], k, bn_stats is not None):
        out = _ConvCF.apply(x, weight, bn_stats)
        return out if bn_stats is not None else (out, None)
    return _ConvLib.apply(x, weight, None, stride, padding, dilation, False, link), None


# BatchNorm statistics of the FFN's norm2 from the epilogue of the dw + dw6 + dw12 GEMM (rss_conv_igemm_stats); RSS_IGEMM_STATS=0: the
# separate statistics pass over the 67 MB tensor
ENGINE["igemm_stats"] = os.environ.get("RSS_IGEMM_STATS", "1") != "0"


def conv_sum_stats_ok(x, Cout):
    """True when conv_sum(x, ..., stats=...) would run the igemm kernel with the statistics epilogue"""
    return ENGINE["igemm_stats"] and Cout <= 128 and _igemm_ok(x, Cout)


# (fc1 -> norm1 and fc2 -> norm3 through the same kernel -- a (Cout,Cin,1,1) bf16 weight IS its [tap][Cout][Cin] operand -- with the
#  statistics epilogue were measured at 521.2 vs 525.2 img/s with cuBLAS + the separate statistics pass: K = 32 / N = 32 GEMMs leave the
#  128-row tcgen05 tile store-bound, 40-48 us vs 18-22 us for nvjet; not kept.  rss_conv_igemm_stats keeps the stat_shift_sub operand
#  that path needed: K = running_mean - bias for a convolution whose bias is left to the BatchNorm.)


def conv_sum(x, convs, bias_grad=True, stats=None, prepacked=None):
    """sum of parallel stride-1 'same' convolutions of the same input (the FFN's dw + dw6 + dw12): one igemm launch when the
    geometry is supported, else the library convs added up.  convs: list of (weight, bias, ksize, dilation).
    stats = (scratch, running_mean) of the BatchNorm that consumes the sum: its raw sums are produced by the GEMM's epilogue
    (only valid when conv_sum_stats_ok(x, Cout); the BatchNorm is then called with aff=ops.RAW_SUMS)."""
    Cout = convs[0][0].shape[0]
    if stats is not None and not conv_sum_stats_ok(x, Cout):
        raise _lib.RssError("conv_sum: statistics epilogue requested for a geometry the igemm kernel does not cover")
    if _igemm_ok(x, Cout):
        cfg = (tuple(c[2] for c in convs), tuple(c[3] for c in convs), bias_grad)
        if stats is not None or prepacked is not None:
            cfg = cfg + (stats, prepacked)
        return _ConvIgemm.apply(x, cfg, *[c[0] for c in convs], *[c[1] for c in convs])
    out = None
    for w, b, k, d in convs:
        t = _ConvLib.apply(x, w, b, 1, d * (k // 2), d, bias_grad)
        out = t if out is None else out + t
    return out


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, bias_grad=True):
    """bias_grad=False: the caller guarantees d loss/d bias == 0 (conv feeding a training-mode BatchNorm: the batch-mean
    subtraction cancels any per-channel shift; the reference computes ~1e-18 round-off there), so the (B*H*W)-long
    reduction is skipped."""
    k = weight.shape[2]
    if ENGINE.get("igemm_single", False) and stride == 1 and padding == dilation * (k // 2) and k in (1, 3) \
            and _igemm_ok(x, weight.shape[0]):
        return _ConvIgemm.apply(x, ((k,), (dilation,), bias_grad), weight, bias)
    return _ConvLib.apply(x, weight, bias, stride, padding, dilation, bias_grad)
