"""TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs /root/reference):

    python -m oracle.gen_golden

1. imports the UNMODIFIED reference through oracle/ref_shim.py,
2. pins oracle/rssformer_ref.py against it in fp64 (forward, loss, every parameter gradient, BN
   running statistics) and writes tests/golden/PIN_REPORT.json,
3. writes golden input/output vectors produced BY THE REFERENCE ITSELF (fp64 run, stored fp32)
   to tests/golden/*.npz.  Inputs and weights are regenerated from seeds
   (oracle.rssformer_ref.synth_state_dict / synth_batch), only outputs are stored.
"""
import io
import json
import os
import contextlib

import numpy as np
import torch

from oracle import rssformer_ref as R
from oracle.ref_shim import build_reference_model, load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
BLK = "backbone.hrnet.stage2.0.transformer."


def _np(t):
    return t.detach().to(torch.float32).cpu().numpy()


def block_inputs(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    y = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    dout = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    return x, y, dout


def loss_inputs(B, S, seed, case):
    g = torch.Generator().manual_seed(seed)
    logits = 2.0 * torch.randn(B, 7, S, S, generator=g, dtype=torch.float64)
    labels = torch.randint(-1, 7, (B, S, S), generator=g, dtype=torch.int64)
    aux = torch.randn(B, 7, generator=g, dtype=torch.float64)
    if case == "edge":            # image 0 all background, image 1 all ignored, image 2 a single fg class
        labels[0] = 0
        labels[1] = -1
        labels[2] = 3
    return logits, labels, aux


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = load_reference()
    report = {}
    sd64 = R.synth_state_dict(2333, torch.float64)

    # ---------------- state_dict keys -----------------------------------------------------
    model = build_reference_model().double()
    ref_sd = model.state_dict()
    spec = R.state_dict_spec()
    assert list(spec) == list(ref_sd), "state_dict key order differs from the reference"
    for k in spec:
        assert tuple(ref_sd[k].shape) == tuple(spec[k]), k
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump({k: list(v) for k, v in spec.items()}, f, indent=0)
    report["state_dict_keys"] = len(spec)
    report["num_parameters"] = int(sum(p.numel() for p in model.parameters()))
    model.load_state_dict(sd64)

    # ---------------- transformer block (a1-a5), reference run -------------------------------
    blk_sd = {k[len(BLK):]: v for k, v in sd64.items() if k.startswith(BLK)}
    for (B, H, W, seed) in [(2, 15, 15, 11), (2, 16, 16, 12), (1, 14, 21, 13), (1, 28, 28, 14)]:
        with contextlib.redirect_stdout(io.StringIO()):
            blk = ns.MTFM.GeneralTransformerBlock(32, 32, 2).double()
        blk.load_state_dict(blk_sd)
        blk.train()
        x, y, dout = block_inputs(B, 32, H, W, seed)
        x.requires_grad_(True); y.requires_grad_(True)
        out = blk(x, y)
        out.backward(dout)
        # oracle restatement on the same inputs
        sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
               for k, v in sd64.items() if k.startswith(BLK)}
        xo = x.detach().clone().requires_grad_(True); yo = y.detach().clone().requires_grad_(True)
        ctx = R.Ctx(sdg, True)
        oo = R.transformer_block(ctx, BLK, xo, yo)
        oo.backward(dout)
        d = dict(fwd=(out - oo).abs().max().item(), dx=(x.grad - xo.grad).abs().max().item(),
                 dy=(y.grad - yo.grad).abs().max().item())
        gsc = max(p.grad.abs().max().item() for p in blk.parameters())
        d["dparam_over_maxgrad"] = max((p.grad - sdg[BLK + k].grad).abs().max().item() for k, p in blk.named_parameters()) / gsc
        report["block_B%d_H%d_W%d" % (B, H, W)] = d
        arrs = dict(out=_np(out), dx=_np(x.grad), dy=_np(y.grad))
        for k, p in blk.named_parameters():
            arrs["grad." + k] = _np(p.grad)
        for k, v in blk.state_dict().items():
            if "running" in k:
                arrs["stat." + k] = _np(v)
        np.savez_compressed(os.path.join(OUT, "block_B%d_H%d_W%d_seed%d.npz" % (B, H, W, seed)), **arrs)

    # ---------------- loss (a11) incl. edge cases, reference run ----------------------------
    with contextlib.redirect_stdout(io.StringIO()):
        loss_mod = ns.CGFL.SegmentationLossaux(dict(ignore_index=-1, ce=dict()))
    for case, B, S, seed in [("rand", 2, 32, 21), ("edge", 4, 16, 22)]:
        logits, labels, aux = loss_inputs(B, S, seed, case)
        logits.requires_grad_(True)
        l = loss_mod(logits, labels, aux)["fc_loss"]
        l.backward()
        lo = logits.detach().clone().requires_grad_(True)
        l2 = R.segmentation_loss(lo, labels, aux)
        l2.backward()
        report["loss_" + case] = dict(loss=abs(l.item() - l2.item()), dlogits=(logits.grad - lo.grad).abs().max().item())
        np.savez_compressed(os.path.join(OUT, "loss_%s_B%d_S%d_seed%d.npz" % (case, B, S, seed)),
                            loss=np.float64(l.item()), dlogits=_np(logits.grad))

    # ---------------- head up-sampling + loss from LOW-resolution logits (hrnet_aux.py:80 + CGFL), reference run
    up = torch.nn.UpsamplingBilinear2d(scale_factor=4.0)
    for case, B, h, seed in [("rand", 2, 16, 31), ("edge", 4, 8, 32)]:
        g = torch.Generator().manual_seed(seed)
        lr = 2.0 * torch.randn(B, 7, h, h, generator=g, dtype=torch.float64)
        labels = torch.randint(-1, 7, (B, 4 * h, 4 * h), generator=g, dtype=torch.int64)
        aux = torch.randn(B, 7, generator=g, dtype=torch.float64)
        if case == "edge":
            labels[0] = 0; labels[1] = -1; labels[2] = 3
        lr.requires_grad_(True)
        l = loss_mod(up(lr), labels, aux)["fc_loss"]
        l.backward()
        np.savez_compressed(os.path.join(OUT, "headloss_%s_B%d_h%d_seed%d.npz" % (case, B, h, seed)),
                            loss=np.float64(l.item()), dlogits_lr=_np(lr.grad))

    # ---------------- full model, S=64 B=2 (train + eval), reference run ----------------------
    img, lbl = R.synth_batch(2, 64, dtype=torch.float64)
    model.eval()
    with torch.no_grad():
        probs = model(img)
        po, _ = R.model_forward(sd64, img, training=False)
    report["model_S64_eval"] = (probs - po).abs().max().item()
    model.train()
    loss = sum(model(img, {"cls": lbl}).values())
    loss.backward()
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd64.items()}
    lo, stats = R.model_forward(sdg, img, lbl, training=True)
    lo["fc_loss"].backward()
    gmax = max(p.grad.abs().max().item() for p in model.parameters() if p.grad is not None)
    gd = 0.0
    no_grad = []
    grad_norms = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            no_grad.append(k)
            continue
        gd = max(gd, (p.grad - sdg[k].grad).abs().max().item())
        grad_norms[k] = float(p.grad.norm().item())
    new_sd = model.state_dict()
    sdiff = max((new_sd[k].double() - v.double()).abs().max().item() for k, v in stats.items())
    report["model_S64_train"] = dict(loss=abs(loss.item() - lo["fc_loss"].item()), dparam_over_maxgrad=gd / gmax,
                                     running_stats=sdiff, params_without_grad=no_grad)
    arrs = dict(probs=_np(probs), loss=np.float64(loss.item()))
    keep = ["backbone.hrnet.conv1.weight", "head.0.weight", "neck.fuse_conv.0.weight",
            BLK + "attn.attn.q_proj.weight", BLK + "attn.attn.k_proj.bias", BLK + "attn.atrous_block1.conv1.weight",
            BLK + "attn.weight_levels.weight", BLK + "norm1.weight", BLK + "mlp.dw6.weight", BLK + "mlp.norm2.weight",
            "backbone.hrnet.stage4.2.transformer.mlp.fc2.weight", "backbone.hrnet.stage4.2.branches.0.3.conv2.weight"]
    named = dict(model.named_parameters())
    for k in keep:
        arrs["grad." + k] = _np(named[k].grad)
    # rounding envelope of the reference itself: its fp32 run vs its fp64 run on the same inputs
    m32 = build_reference_model()
    m32.load_state_dict(R.synth_state_dict(2333, torch.float32))
    m32.train()
    sum(m32(img.float(), {"cls": lbl}).values()).backward()
    n32 = dict(m32.named_parameters())
    for k in keep:
        arrs["fp32dev." + k] = np.float64(((n32[k].grad.double() - named[k].grad).abs().max() / named[k].grad.abs().max()).item())
    gmx = max(grad_norms.values())
    arrs["fp32dev_gradnorm_worst"] = np.float64(max(abs(n32[k].grad.norm().item() - n) / max(n, 1e-3 * gmx)
                                                    for k, n in grad_norms.items() if n > 1e-12))
    report["model_S64_fp32_reference_envelope"] = {k: float(arrs["fp32dev." + k]) for k in keep}
    report["model_S64_fp32_reference_envelope"]["gradnorm_worst"] = float(arrs["fp32dev_gradnorm_worst"])
    arrs["stat.backbone.hrnet.bn1.running_mean"] = _np(new_sd["backbone.hrnet.bn1.running_mean"])
    arrs["stat." + BLK + "mlp.norm2.running_var"] = _np(new_sd[BLK + "mlp.norm2.running_var"])
    np.savez_compressed(os.path.join(OUT, "model_S64_B2.npz"), **arrs)
    with open(os.path.join(OUT, "model_S64_B2_gradnorms.json"), "w") as f:
        json.dump(grad_norms, f, indent=0)

    # ---------------- cfg1: one 512x512 tile, B=1 (fp32 reference run, as train.py would) -------
    model32 = build_reference_model()
    model32.load_state_dict(R.synth_state_dict(2333, torch.float32))
    img, lbl = R.synth_batch(1, 512)
    model32.eval()
    with torch.no_grad():
        p512 = model32(img)
    top2 = p512.topk(2, dim=1).values
    model32.train()
    l512 = sum(model32(img, {"cls": lbl}).values())
    np.savez_compressed(os.path.join(OUT, "model_S512_B1.npz"),
                        argmax=p512.argmax(1).to(torch.uint8).numpy(),
                        top2gap=(top2[:, 0] - top2[:, 1]).to(torch.float16).numpy(),
                        probs_strided=p512[:, :, ::8, ::8].numpy(), loss=np.float64(l512.item()))
    with torch.no_grad():
        po, _ = R.model_forward(R.synth_state_dict(2333, torch.float32), img, training=False)
    report["model_S512_eval_fp32"] = dict(max_abs=(p512 - po).abs().max().item(),
                                          argmax_agree=float((p512.argmax(1) == po.argmax(1)).float().mean().item()))

    with open(os.path.join(OUT, "PIN_REPORT.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
