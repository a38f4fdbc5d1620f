"""GPU parity of the fused stem convolution (csrc/stem.cu) through the C ABI: Conv2d(3,64,3,stride 2,padding 1,bias=False) of
_hrnet_rssformer.py:467 read straight from the planar image batch, its BatchNorm raw sums, its weight gradient, and the model-level
switch (RSS_STEM) -- compared with a plain PyTorch fp32 convolution of the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, H, W): even / odd sizes, rows shorter and longer than one 128-pixel tile, a ragged second tile
SHAPES = [(2, 64, 64), (1, 37, 53), (1, 20, 600), (3, 9, 258), (1, 512, 512)]


def _ref(x, w):
    xr = x.to(torch.bfloat16).double()
    wr = w.to(torch.bfloat16).double()
    return F.conv2d(xr, wr, stride=2, padding=1)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
def test_stem_forward_and_raw_sums(shape, in_dtype, report):
    from representationlearning_b200 import ops
    B, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H + W)
    x = (torch.randn(B, 3, H, W, device="cuda", generator=g) * 1.7 + 0.3).to(in_dtype)
    w = torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2
    scratch = torch.zeros(2 + 128, device="cuda")
    shift = torch.randn(64, device="cuda", generator=g) * 0.1
    y = ops.StemConv.apply(x, w, (scratch, shift))
    torch.cuda.synchronize()
    ref = _ref(x, w)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    err = (y.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    report["stem/fwd_%dx%dx%d_%s" % (B, H, W, str(in_dtype).split(".")[-1])] = {"max_abs_err": err, "max_abs": scale}
    assert err <= 2.0 ** -8 * scale + 1e-6, (err, scale)       # one bf16 rounding of an fp32-accumulated sum
    # raw sums of the ROUNDED outputs, shifted by K
    d = y.double() - shift.double().view(1, 64, 1, 1)
    s1, s2 = d.sum((0, 2, 3)), (d * d).sum((0, 2, 3))
    e1 = (scratch[2:66].double() - s1).abs()
    assert bool((e1 <= 1e-5 * d.abs().sum((0, 2, 3)) + 1e-3).all()), e1.max()      # fp32 partial sums + float atomics
    assert torch.allclose(scratch[66:130].double(), s2, rtol=2e-4, atol=1e-3), (scratch[66:130].double() - s2).abs().max()
    assert scratch[:2].abs().sum().item() == 0.0                 # ticket words untouched
    # without statistics the output is bit-identical
    y2 = ops.StemConv.apply(x, w, None)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("shape", SHAPES)
def test_stem_weight_gradient(shape, report):
    from representationlearning_b200 import ops
    B, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(7 + H * W)
    x = torch.randn(B, 3, H, W, device="cuda", generator=g)
    w = (torch.randn(64, 3, 3, 3, device="cuda", generator=g) * 0.2).requires_grad_(True)
    y = ops.StemConv.apply(x, w, None)
    dy = torch.randn(y.shape, device="cuda", generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y.backward(dy)
    torch.cuda.synchronize()
    wr = w.detach().to(torch.bfloat16).double().requires_grad_(True)
    F.conv2d(x.to(torch.bfloat16).double(), wr, stride=2, padding=1).backward(dy.double())
    err = (w.grad.double() - wr.grad).norm().item() / max(wr.grad.norm().item(), 1e-30)
    report["stem/wgrad_%dx%dx%d" % shape] = {"rel_l2": err}
    assert err < 1e-4, err                                       # fp32 accumulation of exact bf16 x bf16 products


def test_stem_rejects_what_it_does_not_cover():
    from representationlearning_b200 import _lib, ops
    w = torch.randn(64, 3, 3, 3, device="cuda")
    x = torch.randn(1, 3, 16, 16, device="cuda")
    assert ops.stem_conv_ok(x, w)
    assert not ops.stem_conv_ok(x.contiguous(memory_format=torch.channels_last), w)
    assert not ops.stem_conv_ok(x.clone().requires_grad_(True), w)
    assert not ops.stem_conv_ok(torch.randn(1, 4, 16, 16, device="cuda"), w)
    with pytest.raises(_lib.RssError):
        ops.StemConv.apply(torch.randn(1, 4, 16, 16, device="cuda"), w, None)


def test_model_stem_switch_equivalence(monkeypatch, report):
    """whole bf16 model, training mode: the fused stem (with and without the statistics epilogue) against the library stem --
    same loss, same bn1 running statistics, same conv1 gradient up to bf16 rounding noise"""
    import representationlearning_b200 as P
    from representationlearning_b200 import ops
    from oracle import rssformer_ref as R
    img, lbl = R.synth_batch(2, 64)
    img, lbl = img.cuda(), lbl.cuda()
    out = {}
    for mode in ("lib", "fused", "fused_nostats"):
        monkeypatch.setitem(ops.STEM, "on", mode != "lib")
        monkeypatch.setitem(ops.STEM, "stats", mode == "fused")
        m = P.build_rssformer(compute_dtype=torch.bfloat16)
        m.load_state_dict(R.synth_state_dict(2333))
        m.train()
        loss = sum(m(img, {"cls": lbl}).values())
        loss.backward()
        torch.cuda.synchronize()
        hr = m.backbone.hrnet
        out[mode] = (loss.item(), hr.bn1.running_mean.clone(), hr.bn1.running_var.clone(), hr.conv1.weight.grad.clone())
    for mode in ("fused", "fused_nostats"):
        a, b = out[mode], out["lib"]
        report["stem/model_%s" % mode] = {"loss": a[0], "loss_lib": b[0],
                                          "rm_err": (a[1] - b[1]).abs().max().item(), "rv_err": (a[2] - b[2]).abs().max().item()}
        # bn1's statistics see only the stem: tight.  The loss of a 64x64 tile passes through BatchNorms over 8 samples at the coarse
        # resolutions (ill-conditioned), where one flipped bf16 rounding moves it by ~1e-2; the benched geometry
        # is pinned by bench.py's loss check and test_cfg2_* instead.
        assert torch.allclose(a[1], b[1], rtol=0, atol=2e-4) and torch.allclose(a[2], b[2], rtol=2e-3, atol=1e-5)
        assert abs(a[0] - b[0]) <= 3e-2 * abs(b[0]), (a[0], b[0])
        assert torch.isfinite(a[3]).all() and a[3].abs().max().item() > 0
        cos = torch.nn.functional.cosine_similarity(a[3].flatten(), b[3].flatten(), dim=0).item()
        report["stem/model_%s" % mode]["conv1_grad_cosine"] = cos
