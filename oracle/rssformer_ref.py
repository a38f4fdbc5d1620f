"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the RSSFormer training hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this file.  The product (`representationlearning_b200`) never does.

This is a from-scratch functional restatement in plain PyTorch of the algorithm in
`/root/reference/RSSFormer-TIP2023`; every function cites the reference
file:line it follows (paths relative to `RSSFormer-TIP2023/module/baseline/` unless absolute).
It works on a flat `state_dict` with the reference's own key names, in explicit-index form
(flat-buffer saliency gate, explicit window gather, closed-form loss) so that it doubles as the
executable spec for the CUDA kernels.

PARITY PIN: the reference holds no tests/golden vectors for this path,
so the pin is the reference itself run in the authoring container: `oracle/gen_golden.py`
imports the unmodified reference through `oracle/ref_shim.py`, checks this file against it in
fp64 (forward, loss and every parameter gradient; max |diff| recorded in
`tests/golden/PIN_REPORT.json`) and writes the fixtures in `tests/golden/`.
"""
import math
import torch
import torch.nn.functional as F

WINDOW = 7          # MTFM.py:56
NUM_HEADS = 2       # _hrnet_rssformer.py:308
LN_EPS = 1e-6       # MTFM.py:64
BN_EPS = 1e-5       # nn.BatchNorm2d / nn.SyncBatchNorm default
BN_MOMENTUM = 0.1   # _hrnet_rssformer.py:27, torch default for the FFN SyncBN
NUM_CLASSES = 7

# hrnetv2_w32 table, restated from base_hrnet/_hrnet_rssformer.py:97-125
HRNET_W32 = dict(
    stage2=dict(num_modules=1, channels=(32, 64)),
    stage3=dict(num_modules=4, channels=(32, 64, 128)),
    stage4=dict(num_modules=3, channels=(32, 64, 128, 256)),
)


# ------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------
class Ctx:
    """Carries the state_dict, train/eval flag and collects updated BN running statistics."""

    def __init__(self, sd, training):
        self.sd = sd
        self.training = training
        self.new_stats = {}

    def __getitem__(self, k):
        return self.sd[k]


def _bn(ctx, x, prefix, eps=BN_EPS, momentum=BN_MOMENTUM):
    """BatchNorm2d / SyncBatchNorm (single process) — train: batch statistics over (B,H,W),
    running stats updated with momentum 0.1 and the unbiased variance; eval: running stats.
    (_hrnet_rssformer.py:222,225 ...; modules/ffn_block.py:222,231,234)"""
    w, b = ctx[prefix + ".weight"], ctx[prefix + ".bias"]
    rm, rv = ctx[prefix + ".running_mean"], ctx[prefix + ".running_var"]
    if not ctx.training:
        return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)
    dims = (0, 2, 3)
    n = x.numel() // x.shape[1]
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    with torch.no_grad():
        ctx.new_stats[prefix + ".running_mean"] = (1 - momentum) * rm + momentum * mean.detach()
        ctx.new_stats[prefix + ".running_var"] = (1 - momentum) * rv + momentum * var.detach() * (n / max(n - 1, 1))
        ctx.new_stats[prefix + ".num_batches_tracked"] = ctx[prefix + ".num_batches_tracked"] + 1
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var + eps)[None, :, None, None]
    return xh * w[None, :, None, None] + b[None, :, None, None]


def _conv(ctx, x, prefix, stride=1, padding=0, dilation=1):
    return F.conv2d(x, ctx[prefix + ".weight"], ctx.sd.get(prefix + ".bias"), stride, padding, dilation)


def gelu(x):
    """exact-erf GELU (nn.GELU() default; MTFM.py:63, ffn_block.py:213-214)"""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def layer_norm(t, w, b, eps=LN_EPS):
    """LayerNorm over the channel (last) axis of tokens (B,N,C) (MTFM.py:64,78-79,107,109)."""
    mu = t.mean(-1, keepdim=True)
    var = ((t - mu) ** 2).mean(-1, keepdim=True)
    return (t - mu) * torch.rsqrt(var + eps) * w + b


# ------------------------------------------------------------------------------------------
# a2: saliency gate — InterlacedPoolAttention2.forward, multihead_isa_pool_attention.py:148-167
# ------------------------------------------------------------------------------------------
def saliency_gate(xn, yn, w_sa1, w_sa2, w_lvl, b_lvl, H, W):
    """xn, yn: (B,N,C) token-major *contiguous* buffers (the LN outputs).

    The reference does `x.view(B,C,H,W)` on that buffer (pool:150-151): a re-interpretation,
    not a transpose.  In flat form, with f = index into the per-image (N*C) buffer:
        map entry j (0<=j<HW) pools the C flat elements {k*HW + j : k<C}          (pool:110-112)
        s_x = sigmoid(conv7x7([mean,max](x)))   s_y likewise with atrous_block2    (pool:113-115,156-157)
        (l0,l1) = weight_levels([s_x,s_y]) ; (g0,g1) = softmax                     (pool:161-163)
        x_flat[f] *= g0[f mod HW] ; y_flat[f] *= g1[f mod HW]                      (pool:164-167)
    Returns gated (B,N,C) tensors and the gate maps (B,2,H,W)."""
    B, N, C = xn.shape
    HW = H * W
    xf = xn.reshape(B, C, HW)          # flat f = k*HW + j  ->  [k][j]
    yf = yn.reshape(B, C, HW)

    def smap(f, w):
        pooled = torch.stack([f.mean(1), f.max(1).values], 1).reshape(B, 2, H, W)
        return torch.sigmoid(F.conv2d(pooled, w, None, 1, 3))

    s = torch.cat([smap(xf, w_sa1), smap(yf, w_sa2)], 1)            # (B,2,H,W)
    g = torch.softmax(F.conv2d(s, w_lvl, b_lvl), dim=1)             # (B,2,H,W)
    xg = (xf * g[:, 0].reshape(B, 1, HW)).reshape(B, N, C)
    yg = (yf * g[:, 1].reshape(B, 1, HW)).reshape(B, N, C)
    return xg, yg, g


# ------------------------------------------------------------------------------------------
# a3: centre zero-pad + 7x7 window gather — multihead_isa_attention.py:373-426
# ------------------------------------------------------------------------------------------
def pad_amounts(H, W, ws=WINDOW):
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    return Hp, Wp, (Hp - H) // 2, (Wp - W) // 2      # pad//2 before, the rest after (isa:377-381)


def window_gather(t, H, W, ws=WINDOW):
    """(B,H*W,C) -> (nWin, ws*ws, C); window index = (b*qh + i)*qw + j (isa:402-413).
    Pad tokens are exact zeros and are NOT masked: they act as real keys/values (no mask is built)."""
    B, N, C = t.shape
    Hp, Wp, ph, pw = pad_amounts(H, W, ws)
    t = F.pad(t.reshape(B, H, W, C), (0, 0, pw, Wp - W - pw, ph, Hp - H - ph))
    qh, qw = Hp // ws, Wp // ws
    t = t.reshape(B, qh, ws, qw, ws, C).permute(0, 1, 3, 2, 4, 5)
    return t.reshape(B * qh * qw, ws * ws, C)


def window_scatter(tw, B, H, W, ws=WINDOW):
    """inverse of window_gather followed by the crop (isa:384-390,415-426)."""
    C = tw.shape[-1]
    Hp, Wp, ph, pw = pad_amounts(H, W, ws)
    qh, qw = Hp // ws, Wp // ws
    t = tw.reshape(B, qh, qw, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    return t[:, ph:ph + H, pw:pw + W, :].reshape(B, H * W, C)


# ------------------------------------------------------------------------------------------
# a4: Mhca — modules/DAL.py:785-1030 (the live branch only: no masks, dropout 0)
# ------------------------------------------------------------------------------------------
def mhca(xw, yw, wq, bq, wk, bk, wv, bv, wo, bo, num_heads=NUM_HEADS):
    """xw (queries), yw (keys=values): (nWin, L, C).
    q=(Wq x+bq)*hd^-0.5 (DAL:873); k,v (DAL:874-875); per head S=q k^T (DAL:959), P=softmax (DAL:996),
    S2=q^T k (hd x hd), gate=sigmoid(mean(S2)+max(S2)) (DAL:1003-1010), O=(P v)*gate (DAL:1012-1013),
    heads merged, out_proj (DAL:1017-1020)."""
    nW, L, C = xw.shape
    hd = C // num_heads
    q = (xw @ wq.t() + bq) * (float(hd) ** -0.5)
    k = yw @ wk.t() + bk
    v = yw @ wv.t() + bv
    q = q.reshape(nW, L, num_heads, hd).transpose(1, 2)      # (nW, h, L, hd)
    k = k.reshape(nW, L, num_heads, hd).transpose(1, 2)
    v = v.reshape(nW, L, num_heads, hd).transpose(1, 2)
    P = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    S2 = q.transpose(-1, -2) @ k                              # (nW, h, hd, hd)
    gate = torch.sigmoid(S2.mean((-1, -2)) + S2.amax((-1, -2)))
    O = (P @ v) * gate[..., None, None]
    O = O.transpose(1, 2).reshape(nW, L, C)
    return O @ wo.t() + bo


def window_attention(ctx, p, xn, yn, H, W):
    """a2+a3+a4 chained: InterlacedPoolAttention2.forward (pool:148-188)."""
    B = xn.shape[0]
    xg, yg, _ = saliency_gate(xn, yn, ctx[p + "atrous_block1.conv1.weight"], ctx[p + "atrous_block2.conv1.weight"],
                              ctx[p + "weight_levels.weight"], ctx[p + "weight_levels.bias"], H, W)
    a = p + "attn."
    ow = mhca(window_gather(xg, H, W), window_gather(yg, H, W),
              ctx[a + "q_proj.weight"], ctx[a + "q_proj.bias"], ctx[a + "k_proj.weight"], ctx[a + "k_proj.bias"],
              ctx[a + "v_proj.weight"], ctx[a + "v_proj.bias"], ctx[a + "out_proj.weight"], ctx[a + "out_proj.bias"])
    return window_scatter(ow, B, H, W)


# ------------------------------------------------------------------------------------------
# a5: MlpDWBN (token branch) — modules/ffn_block.py:237-270
# ------------------------------------------------------------------------------------------
def ffn(ctx, p, t, H, W):
    B, N, C = t.shape
    x = t.permute(0, 2, 1).reshape(B, C, H, W)
    x = gelu(_bn(ctx, _conv(ctx, x, p + "fc1"), p + "norm1"))                      # ffn:246-248
    x = (_conv(ctx, x, p + "dw") + _conv(ctx, x, p + "dw6", 1, 6, 6)) + _conv(ctx, x, p + "dw12", 1, 12, 12)  # ffn:250-257
    x = gelu(_bn(ctx, x, p + "norm2"))                                             # ffn:258-259
    x = gelu(_bn(ctx, _conv(ctx, x, p + "fc2"), p + "norm3"))                      # ffn:261-263
    return x.reshape(B, C, N).permute(0, 2, 1)


# ------------------------------------------------------------------------------------------
# a1: GeneralTransformerBlock.forward — modules/MTFM.py:101-113
# ------------------------------------------------------------------------------------------
def transformer_block(ctx, p, x, y):
    """x = `low` (queries + residual stream), y = high-res branch (keys/values only). NCHW in/out."""
    B, C, H, W = x.shape
    t = x.reshape(B, C, H * W).permute(0, 2, 1)
    u = y.reshape(B, C, H * W).permute(0, 2, 1)
    n1w, n1b = ctx[p + "norm1.weight"], ctx[p + "norm1.bias"]
    xn = layer_norm(t, n1w, n1b).contiguous()         # the same norm1 serves both inputs (MTFM:107)
    yn = layer_norm(u, n1w, n1b).contiguous()
    t = t + window_attention(ctx, p + "attn.", xn, yn, H, W)
    t = t + ffn(ctx, p + "mlp.", layer_norm(t, ctx[p + "norm2.weight"], ctx[p + "norm2.bias"]), H, W)
    return t.permute(0, 2, 1).reshape(B, C, H, W)


# ------------------------------------------------------------------------------------------
# a7: residual blocks, stem, transitions — _hrnet_rssformer.py:216-287, 461-466, 512-546
# ------------------------------------------------------------------------------------------
def basic_block(ctx, p, x):
    o = F.relu(_bn(ctx, _conv(ctx, x, p + "conv1", 1, 1), p + "bn1"))
    o = _bn(ctx, _conv(ctx, o, p + "conv2", 1, 1), p + "bn2")
    return F.relu(o + x)


def bottleneck(ctx, p, x):
    o = F.relu(_bn(ctx, _conv(ctx, x, p + "conv1"), p + "bn1"))
    o = F.relu(_bn(ctx, _conv(ctx, o, p + "conv2", 1, 1), p + "bn2"))
    o = _bn(ctx, _conv(ctx, o, p + "conv3"), p + "bn3")
    r = x
    if (p + "downsample.0.weight") in ctx.sd:
        r = _bn(ctx, _conv(ctx, x, p + "downsample.0"), p + "downsample.1")
    return F.relu(o + r)


# ------------------------------------------------------------------------------------------
# a6: HighResolutionModule.forward — _hrnet_rssformer.py:410-437 (+ fuse layers :361-405)
# ------------------------------------------------------------------------------------------
def hr_module(ctx, p, xs):
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for blk in range(4):
            xs[i] = basic_block(ctx, "%sbranches.%d.%d." % (p, i, blk), xs[i])
    outs = []
    for i in range(nb):
        low = 0
        for j in range(1, nb):
            if j == i:
                term = xs[j]
            elif j > i:      # 1x1 conv + BN + nearest upsample x2^(j-i)   (:372-380)
                f = "%sfuse_layers.%d.%d." % (p, i, j)
                term = _bn(ctx, _conv(ctx, xs[j], f + "0"), f + "1")
                term = F.interpolate(term, scale_factor=2 ** (j - i), mode="nearest")
            else:            # chain of (i-j) stride-2 3x3 convs, ReLU on all but the last (:384-402)
                term = xs[j]
                for k in range(i - j):
                    f = "%sfuse_layers.%d.%d.%d." % (p, i, j, k)
                    term = _bn(ctx, _conv(ctx, term, f + "0", 2, 1), f + "1")
                    if k != i - j - 1:
                        term = F.relu(term)
            low = low + term
        if i == 0:
            y = transformer_block(ctx, p + "transformer.", low, xs[0])     # (:430-431) no "+ x[0]" residual
        else:
            y = xs[0]
            for k in range(i):
                f = "%sfuse_layers.%d.0.%d." % (p, i, k)
                y = _bn(ctx, _conv(ctx, y, f + "0", 2, 1), f + "1")
                if k != i - 1:
                    y = F.relu(y)
            y = y + low
        outs.append(F.relu(y))
    return outs


def hrnet_backbone(ctx, img, p="backbone.hrnet."):
    """HighResolutionNet.forward — _hrnet_rssformer.py:605-640."""
    x = F.relu(_bn(ctx, _conv(ctx, img, p + "conv1", 2, 1), p + "bn1"))
    x = F.relu(_bn(ctx, _conv(ctx, x, p + "conv2", 2, 1), p + "bn2"))
    for i in range(4):
        x = bottleneck(ctx, "%slayer1.%d." % (p, i), x)

    def cbr(pref, t, stride):
        return F.relu(_bn(ctx, _conv(ctx, t, pref + "0", stride, 1), pref + "1"))

    ys = [cbr(p + "transition1.0.", x, 1), cbr(p + "transition1.1.0.", x, 2)]
    ys = hr_module(ctx, p + "stage2.0.", ys)
    ys = ys + [cbr(p + "transition2.2.0.", ys[-1], 2)]
    for m in range(HRNET_W32["stage3"]["num_modules"]):
        ys = hr_module(ctx, "%sstage3.%d." % (p, m), ys)
    ys = ys + [cbr(p + "transition3.3.0.", ys[-1], 2)]
    for m in range(HRNET_W32["stage4"]["num_modules"]):
        ys = hr_module(ctx, "%sstage4.%d." % (p, m), ys)
    return ys


# ------------------------------------------------------------------------------------------
# a8 neck, a9 head, a10 aux head — hrnet_aux.py:42-68, 78-81, 86-87, 99-101
# ------------------------------------------------------------------------------------------
def neck(ctx, feats):
    x0 = feats[0]
    size = x0.shape[2:]
    ups = [x0] + [F.interpolate(f, size=size, mode="bilinear", align_corners=True) for f in feats[1:]]
    x = torch.cat(ups, 1)
    x = F.relu(_bn(ctx, _conv(ctx, x, "neck.fuse_conv.0"), "neck.fuse_conv.1"))
    return x, x0


def head(ctx, x, scale=4):
    x = _conv(ctx, x, "head.0")
    return F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=True)


def headaux(ctx, f0):
    return f0.mean((2, 3)) @ ctx["headaux.0.weight"].t() + ctx["headaux.0.bias"]


# ------------------------------------------------------------------------------------------
# a11: loss — module/CGFL.py:201-227, 72-101; losses/auxloss.py:257-305
# ------------------------------------------------------------------------------------------
def fg_presence_scores(labels, aux_scores):
    """l1_b of auxloss.py:276-292.  multi-hot over 7 classes built from unique(fg-mask) per image:
    index 0 set iff the image has any background-or-ignored pixel, index 1 iff any foreground
    (label>0) pixel (CGFL:210-213 + auxloss:276-281)."""
    B = labels.shape[0]
    fg = labels > 0
    onehot = torch.zeros_like(aux_scores)
    onehot[:, 0] = (~fg).reshape(B, -1).any(1).to(aux_scores.dtype)
    onehot[:, 1] = fg.reshape(B, -1).any(1).to(aux_scores.dtype)
    return (1.0 / (1.0 + torch.exp((aux_scores - onehot).abs()))).sum(1) / (2 * B)


def segmentation_loss(logits, labels, aux_scores, ignore_index=-1):
    """{'fc_loss': CE_mean * sum_pix (1-p_t)(1-l1_b/7) / (n_valid + B)} (CGFL:74-99).
    The modulating sum runs over ALL pixels; ignored pixels are gathered at class 0 (CGFL:90-96).
    Gradient flows only through CE (the factor is computed under no_grad)."""
    B = logits.shape[0]
    ce = F.cross_entropy(logits, labels, ignore_index=ignore_index)
    with torch.no_grad():
        l1 = fg_presence_scores(labels, aux_scores)
        p = torch.softmax(logits, 1)
        valid = labels != ignore_index
        tgt = torch.where(valid, labels, torch.zeros_like(labels))
        pt = p.gather(1, tgt[:, None]).squeeze(1)
        mod = ((1.0 - pt) * (1.0 - l1 / 7.0)[:, None, None]).sum()
        factor = mod / (valid.sum() + B)
    return ce * factor


# ------------------------------------------------------------------------------------------
# HRNetFusion.forward — hrnet_aux.py:89-110
# ------------------------------------------------------------------------------------------
def model_forward(sd, img, labels=None, training=True):
    """train: returns ({'fc_loss': scalar}, new_running_stats); eval: ((B,7,H,W) softmax, {})."""
    ctx = Ctx(sd, training)
    feats = hrnet_backbone(ctx, img)
    fused, f0 = neck(ctx, feats)
    aux = headaux(ctx, f0)
    logits = head(ctx, fused)
    if training:
        return {"fc_loss": segmentation_loss(logits, labels.long(), aux)}, ctx.new_stats
    return torch.softmax(logits, 1), ctx.new_stats


# ------------------------------------------------------------------------------------------
# a12: optimiser step semantics — configs/base/loveda.py:68-99 (engine is `ever`, unpinned)
# ------------------------------------------------------------------------------------------
def poly_lr(it, base_lr=0.01, power=0.9, max_iters=30000):
    return base_lr * (1.0 - it / max_iters) ** power


def sgd_step(params, grads, momenta, lr, momentum=0.9, weight_decay=1e-4, max_norm=35.0):
    """clip_grad_norm_(35, L2) over the params that received a grad, then torch.optim.SGD semantics
    (g += wd*p ; buf = m*buf + g (buf=g on first step) ; p -= lr*buf).  In-place on params/momenta.
    Returns the pre-clip total norm."""
    live = [g for g in grads if g is not None]
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in live)).to(live[0].dtype)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    with torch.no_grad():
        for i, (p, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            g = g * coef + weight_decay * p
            if momenta[i] is None:
                momenta[i] = g.clone()
            else:
                momenta[i].mul_(momentum).add_(g)
            p.add_(momenta[i], alpha=-lr)
    return total


# ------------------------------------------------------------------------------------------
# state_dict spec + deterministic synthetic weights (no checkpoint/network offline)
# ------------------------------------------------------------------------------------------
def state_dict_spec():
    """Ordered {key: shape} of the reference `HRNetFusion(hrnetv2_w32)` state_dict, rebuilt from the
    constructor logic (_hrnet_rssformer.py:446-603, hrnet_aux.py:72-87). Pinned against the real
    reference's key set by gen_golden.py / tests."""
    spec = {}

    def conv(p, co, ci, k, bias=False):
        spec[p + ".weight"] = (co, ci, k, k)
        if bias:
            spec[p + ".bias"] = (co,)

    def bn(p, c):
        spec[p + ".weight"] = (c,)
        spec[p + ".bias"] = (c,)
        spec[p + ".running_mean"] = (c,)
        spec[p + ".running_var"] = (c,)
        spec[p + ".num_batches_tracked"] = ()

    def lin(p, co, ci):
        spec[p + ".weight"] = (co, ci)
        spec[p + ".bias"] = (co,)

    h = "backbone.hrnet."
    conv(h + "conv1", 64, 3, 3); bn(h + "bn1", 64)
    conv(h + "conv2", 64, 64, 3); bn(h + "bn2", 64)
    for i in range(4):
        p = "%slayer1.%d." % (h, i)
        cin = 64 if i == 0 else 256
        conv(p + "conv1", 64, cin, 1); bn(p + "bn1", 64)
        conv(p + "conv2", 64, 64, 3); bn(p + "bn2", 64)
        conv(p + "conv3", 256, 64, 1); bn(p + "bn3", 256)
        if i == 0:
            conv(p + "downsample.0", 256, 64, 1); bn(p + "downsample.1", 256)
    conv(h + "transition1.0.0", 32, 256, 3); bn(h + "transition1.0.1", 32)
    conv(h + "transition1.1.0.0", 64, 256, 3); bn(h + "transition1.1.0.1", 64)

    def block(p, C=32, hidden=128):
        a = p + "attn."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(a + "attn." + n, C, C)
        spec[a + "atrous_block1.conv1.weight"] = (1, 2, 7, 7)
        spec[a + "atrous_block2.conv1.weight"] = (1, 2, 7, 7)
        spec[a + "weight_levels.weight"] = (2, 2, 1, 1)
        spec[a + "weight_levels.bias"] = (2,)
        for n in ("norm1", "norm2"):
            spec[p + n + ".weight"] = (C,)
            spec[p + n + ".bias"] = (C,)
        m = p + "mlp."
        conv(m + "fc1", hidden, C, 1, True); bn(m + "norm1", hidden)
        conv(m + "dw", hidden, hidden, 1, True)
        conv(m + "dw6", hidden, hidden, 3, True)
        conv(m + "dw12", hidden, hidden, 3, True)
        bn(m + "norm2", hidden)
        conv(m + "fc2", C, hidden, 1, True); bn(m + "norm3", C)

    def stage(name, cfg):
        ch = cfg["channels"]
        nb = len(ch)
        for m in range(cfg["num_modules"]):
            p = "%s%s.%d." % (h, name, m)
            for i in range(nb):
                for b in range(4):
                    q = "%sbranches.%d.%d." % (p, i, b)
                    conv(q + "conv1", ch[i], ch[i], 3); bn(q + "bn1", ch[i])
                    conv(q + "conv2", ch[i], ch[i], 3); bn(q + "bn2", ch[i])
            for i in range(nb):
                for j in range(nb):
                    if j > i:
                        q = "%sfuse_layers.%d.%d." % (p, i, j)
                        conv(q + "0", ch[i], ch[j], 1); bn(q + "1", ch[i])
                    elif j < i:
                        for k in range(i - j):
                            q = "%sfuse_layers.%d.%d.%d." % (p, i, j, k)
                            co = ch[i] if k == i - j - 1 else ch[j]
                            conv(q + "0", co, ch[j], 3); bn(q + "1", co)
            block(p + "transformer.")

    stage("stage2", HRNET_W32["stage2"])
    conv(h + "transition2.2.0.0", 128, 64, 3); bn(h + "transition2.2.0.1", 128)
    stage("stage3", HRNET_W32["stage3"])
    conv(h + "transition3.3.0.0", 256, 128, 3); bn(h + "transition3.3.0.1", 256)
    stage("stage4", HRNET_W32["stage4"])
    conv("neck.fuse_conv.0", 480, 480, 1, True); bn("neck.fuse_conv.1", 480)
    conv("head.0", NUM_CLASSES, 480, 1, True)
    lin("headaux.0", NUM_CLASSES, 32)
    return spec


def synth_state_dict(seed=2333, dtype=torch.float32, spec=None):
    """Deterministic weights that depend only on (seed, key, shape): every key gets its own
    generator, so the reference model, this oracle and the CUDA product can be loaded with
    bit-identical parameters on any machine.  Scales follow PyTorch's default inits
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear) — the reference applies no custom init
    (_hrnet_rssformer.py:187-207 are never called).  BN/LN affine and running stats are perturbed
    away from (1,0,0,1) so that parity tests exercise them."""
    import zlib
    spec = spec or state_dict_spec()
    sd = {}
    for k, shp in spec.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = (0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif k.endswith("running_var"):
            sd[k] = (1.0 + 0.2 * torch.rand(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif len(shp) == 1 and k.endswith(".weight"):          # BN / LN gamma
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        elif len(shp) == 1:                                      # biases, BN/LN beta
            sd[k] = (0.1 * torch.randn(shp, generator=g, dtype=torch.float64)).to(dtype)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)
            sd[k] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return sd


def synth_batch(B, S, seed_img=7, seed_lbl=1, dtype=torch.float32):
    """SURVEY §8(d) synthetic tiles: images randn(B,3,S,S) seed 7, labels randint(-1,7) seed 1."""
    gi = torch.Generator().manual_seed(seed_img)
    gl = torch.Generator().manual_seed(seed_lbl)
    img = torch.randn(B, 3, S, S, generator=gi, dtype=torch.float32).to(dtype)
    lbl = torch.randint(-1, 7, (B, S, S), generator=gl, dtype=torch.int64)
    return img, lbl
