"""TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs /root/reference):

    python -m oracle.gen_golden_cfg

Golden vectors at the BENCHED geometries, produced by the UNMODIFIED reference (oracle/ref_shim.py) on CPU:

  tests/golden/model_S512_B2_cfg2.npz   BASELINE cfg2 geometry (512x512 tiles, B reduced 16 -> 2 so that the fixture stays small and
                                        the CPU run stays in seconds): eval probabilities (strided) / argmax / top-2 gap, training
                                        loss, total gradient norm, 12 parameter gradients and the first SGD update of two of them
  tests/golden/model_S1024_B1_cfg5.npz  BASELINE cfg5 geometry (1024x1024, branch 0 = 256x256 -> padded to 259 for the 7x7 windows)

Both hold, next to the fp32 results, the reference's OWN bf16-autocast deviation from them per tensor ("envelope"): the GPU
tests hold the bf16 kernels to <= 2x that envelope instead of a hand-picked tolerance.
Inputs and weights are regenerated from seeds (oracle.rssformer_ref.synth_state_dict / synth_batch); only outputs are stored.
"""
import json
import os

import numpy as np
import torch

from oracle import rssformer_ref as R
from oracle.ref_shim import build_reference_model

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
BLK = "backbone.hrnet.stage2.0.transformer."
KEEP = ["backbone.hrnet.conv1.weight", "head.0.weight", "neck.fuse_conv.0.weight",
        BLK + "attn.attn.q_proj.weight", BLK + "attn.attn.k_proj.bias", BLK + "attn.atrous_block1.conv1.weight",
        BLK + "attn.weight_levels.weight", BLK + "norm1.weight", BLK + "mlp.dw6.weight", BLK + "mlp.norm2.weight",
        "backbone.hrnet.stage4.2.transformer.mlp.fc2.weight", "backbone.hrnet.stage4.2.branches.0.3.conv2.weight",
        "backbone.hrnet.stage3.1.branches.0.1.conv1.weight", "backbone.hrnet.stage3.1.branches.0.1.bn2.weight",
        "backbone.hrnet.stage3.1.branches.1.2.conv2.weight"]
HP = dict(lr=0.01, momentum=0.9, weight_decay=1e-4, max_norm=35.0)       # configs/base/loveda.py:68-93


def _np(t):
    return t.detach().to(torch.float32).cpu().numpy()


def _l2(a, b):
    return float(((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item())


def _mx(a, b):
    return float(((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item())


def run(model, img, lbl, autocast):
    """eval probabilities, then one training forward/backward -> (probs, loss, {name: grad})"""
    img = img.to(next(model.parameters()).dtype)
    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if autocast else torch.autocast("cpu", enabled=False)
    model.eval()
    with torch.no_grad(), ctx:
        probs = model(img).float()
    model.train()
    for p in model.parameters():
        p.grad = None
    with ctx:
        loss = sum(model(img, {"cls": lbl}).values())
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return probs, float(loss.item()), grads


def make(S, B, tag, with_grads):
    torch.manual_seed(0)
    sd = R.synth_state_dict(2333, torch.float32)
    img, lbl = R.synth_batch(B, S)
    model = build_reference_model()
    model.load_state_dict(sd)
    p32, l32, g32 = run(model, img, lbl, False)
    model.load_state_dict(sd)                       # the training pass moved the BN running statistics
    p16, l16, g16 = run(model, img, lbl, True)
    top2 = p32.topk(2, dim=1).values
    arrs = dict(argmax=p32.argmax(1).to(torch.uint8).numpy(), top2gap=(top2[:, 0] - top2[:, 1]).to(torch.float16).numpy(),
                probs_strided=p32[:, :, ::8, ::8].numpy(), loss=np.float64(l32))
    env = dict(loss_rel=abs(l16 - l32) / abs(l32), probs_max=float((p16 - p32).abs().max().item()),
               argmax_agree=float((p16.argmax(1) == p32.argmax(1)).float().mean().item()))
    if with_grads:
        # fp64 run of the reference = the truth the gradients are stored from; its fp32 run's deviation = the fp32 envelope
        m64 = build_reference_model().double()
        m64.load_state_dict(R.synth_state_dict(2333, torch.float64))
        _, l64, g64 = run(m64, img, lbl, False)
        env["fp32_loss_rel"] = abs(l32 - l64) / abs(l64)
        for k in KEEP:
            env["fp32_grad_l2." + k] = _l2(g32[k], g64[k])
            arrs["grad64." + k] = _np(g64[k])
        del m64
        gn = torch.sqrt(sum(g.double().pow(2).sum() for g in g32.values()))
        gn16 = torch.sqrt(sum(g.double().pow(2).sum() for g in g16.values()))
        arrs["grad_norm"] = np.float64(gn.item())
        env["grad_norm_rel"] = abs(gn16.item() - gn.item()) / gn.item()
        clip = min(1.0, HP["max_norm"] / (gn.item() + 1e-6))
        for k in KEEP:
            arrs["grad." + k] = _np(g32[k])
            env["grad_l2." + k] = _l2(g16[k], g32[k])
            env["grad_max." + k] = _mx(g16[k], g32[k])
        # first SGD step (momentum buffer starts at 0): p1 = p0 - lr * (clip * g + wd * p0)
        for k in KEEP[:2]:
            arrs["update." + k] = _np(-HP["lr"] * (clip * g32[k] + HP["weight_decay"] * sd[k]))
    arrs["envelope_json"] = np.frombuffer(json.dumps(env).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **arrs)
    print(tag, json.dumps(env, indent=1))
    return env


def main():
    os.makedirs(OUT, exist_ok=True)
    rep = {"cfg2_S512_B2": make(512, 2, "model_S512_B2_cfg2", True),
           "cfg5_S1024_B1": make(1024, 1, "model_S1024_B1_cfg5", False)}
    with open(os.path.join(OUT, "ENVELOPE_REPORT.json"), "w") as f:
        json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main()
