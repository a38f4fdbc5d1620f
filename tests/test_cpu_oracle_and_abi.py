"""CPU suite (`-m "not gpu"`): oracle vs committed golden vectors, host logic, C-ABI symbol export."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import rssformer_ref as R

BLK = "backbone.hrnet.stage2.0.transformer."


def _rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def test_pin_report_is_tight():
    rep = json.load(open(os.path.join(GOLDEN, "PIN_REPORT.json")))
    assert rep["state_dict_keys"] == 2169 and rep["num_parameters"] == 32142270
    for k, v in rep.items():
        if k.startswith("block_"):
            assert max(v.values()) < 1e-11, (k, v)
    assert rep["model_S64_eval"] < 1e-12
    assert rep["model_S64_train"]["loss"] < 1e-12 and rep["model_S64_train"]["dparam_over_maxgrad"] < 1e-9
    assert rep["model_S64_train"]["params_without_grad"] == ["headaux.0.weight", "headaux.0.bias"]
    assert rep["model_S512_eval_fp32"]["argmax_agree"] == 1.0


def test_state_dict_spec_matches_reference_keys():
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    mine = R.state_dict_spec()
    assert list(mine) == list(spec)
    assert all(list(mine[k]) == spec[k] for k in spec)


def test_product_module_tree_has_reference_state_dict():
    import representationlearning_b200 as P
    m = P.HRNetFusion(P.RSSFORMER_CONFIG)
    spec = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    sd = m.state_dict()
    assert list(sd) == list(spec)
    assert all(list(sd[k].shape) == spec[k] for k in spec)
    m.load_state_dict(R.synth_state_dict())          # round trip with reference-shaped checkpoints
    blk = P.GeneralTransformerBlock(32, 32, 2)
    assert sorted(blk.state_dict()) == sorted(k[len(BLK):] for k in spec if k.startswith(BLK))


@pytest.mark.parametrize("name", ["block_B2_H15_W15_seed11", "block_B2_H16_W16_seed12", "block_B1_H14_W21_seed13",
                                  "block_B1_H28_W28_seed14"])
def test_oracle_block_vs_reference_golden(name):
    from oracle.gen_golden import block_inputs
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, H, W, seed = [int(x) for x in re.match(r"block_B(\d+)_H(\d+)_W(\d+)_seed(\d+)", name).groups()]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
          for k, v in R.synth_state_dict(2333, torch.float64).items() if k.startswith(BLK)}
    x, y, dout = block_inputs(B, 32, H, W, seed)
    x.requires_grad_(True); y.requires_grad_(True)
    ctx = R.Ctx(sd, True)
    out = R.transformer_block(ctx, BLK, x, y)
    out.backward(dout)
    assert _rel(out, g["out"]) < 1e-6 and _rel(x.grad, g["dx"]) < 1e-6 and _rel(y.grad, g["dy"]) < 1e-5
    gmax = max(np.abs(g[k]).max() for k in g.files if k.startswith("grad."))
    for k in g.files:
        if k.startswith("grad."):
            d = (sd[BLK + k[5:]].grad - torch.as_tensor(g[k], dtype=torch.float64)).abs().max().item()
            assert d / gmax < 1e-6, k
        if k.startswith("stat."):
            assert _rel(ctx.new_stats[BLK + k[5:]], g[k]) < 1e-6, k


@pytest.mark.parametrize("name,case,B,S,seed", [("loss_rand_B2_S32_seed21", "rand", 2, 32, 21), ("loss_edge_B4_S16_seed22", "edge", 4, 16, 22)])
def test_oracle_loss_vs_reference_golden(name, case, B, S, seed):
    from oracle.gen_golden import loss_inputs
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    logits, labels, aux = loss_inputs(B, S, seed, case)
    logits.requires_grad_(True)
    l = R.segmentation_loss(logits, labels, aux)
    l.backward()
    assert abs(l.item() - float(g["loss"])) < 1e-9
    assert _rel(logits.grad, g["dlogits"]) < 1e-6


def test_oracle_model_vs_reference_golden_S64():
    g = np.load(os.path.join(GOLDEN, "model_S64_B2.npz"))
    sd = R.synth_state_dict(2333, torch.float64)
    img, lbl = R.synth_batch(2, 64, dtype=torch.float64)
    with torch.no_grad():
        probs, _ = R.model_forward(sd, img, training=False)
    assert _rel(probs, g["probs"]) < 1e-6
    sdg = {k: (v.clone().requires_grad_(True) if k in ("head.0.weight", BLK + "norm1.weight") else v) for k, v in sd.items()}
    out, stats = R.model_forward(sdg, img, lbl, training=True)
    out["fc_loss"].backward()
    assert abs(out["fc_loss"].item() - float(g["loss"])) < 1e-9
    assert _rel(sdg["head.0.weight"].grad, g["grad.head.0.weight"]) < 1e-6
    assert _rel(sdg[BLK + "norm1.weight"].grad, g["grad." + BLK + "norm1.weight"]) < 1e-5
    assert _rel(stats["backbone.hrnet.bn1.running_mean"], g["stat.backbone.hrnet.bn1.running_mean"]) < 1e-6


def test_oracle_sgd_matches_torch_optim():
    torch.manual_seed(0)
    ps = [torch.randn(5, 3), torch.randn(7)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.SGD(ref, lr=0.01, momentum=0.9, weight_decay=1e-4)
    mom = [None, None]
    for it in range(3):
        gs = [torch.randn_like(p) * 30 for p in ps]
        for r, g_ in zip(ref, gs):
            r.grad = g_.clone()
        torch.nn.utils.clip_grad_norm_(ref, 35.0, 2)
        for gq in opt.param_groups:
            gq["lr"] = R.poly_lr(it)
        opt.step()
        R.sgd_step(ps, gs, mom, R.poly_lr(it))
        for a, b in zip(ps, ref):
            assert torch.allclose(a, b.detach(), atol=1e-6)


def test_poly_lr_host_logic():
    from representationlearning_b200.trainer import poly_lr
    assert poly_lr(0) == 0.01
    assert abs(poly_lr(15000) - 0.01 * 0.5 ** 0.9) < 1e-12
    assert poly_lr(30000) == 0.0 and poly_lr(40000) == 0.0
    assert abs(poly_lr(123) - R.poly_lr(123)) < 1e-15


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "rss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rss_[a-z0-9_]+)\s*\(", src)))


def test_c_abi_library_exports_every_declared_symbol():
    from representationlearning_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "declared in include/rss_b200.h but not exported: " + s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes prototypes and header disagree"
    assert lib.rss_version() >= 100
    # exported dynamic symbols carry no torch / C++ types (plain C ABI)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert all(s in exported for s in syms)


def test_product_fails_loudly_without_gpu():
    import representationlearning_b200 as P
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = P.HRNetFusion(P.RSSFORMER_CONFIG)
    with pytest.raises(P._lib.RssError):
        m(torch.randn(1, 3, 64, 64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "representationlearning_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_channels_last_weight_shadow_view_is_used_without_a_copy():
    """conv._lowp(channels_last=True) hands the library convolution the per-step (Cout,kh,kw,Cin) bf16 shadow as a
    channels_last-strided (Cout,Cin,kh,kw) view: `.contiguous(memory_format=channels_last)` must be a no-op on it."""
    from representationlearning_b200 import conv
    w = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
    store = torch.empty(8 * 3 * 3 * 4, dtype=torch.bfloat16)
    view = store.view(8, 3, 3, 4).permute(0, 3, 1, 2)
    view.copy_(w.detach())                                   # what rss_shadow_cl_refresh writes: (co, kh, kw, ci) memory order
    conv.register_shadow_cl(w, view)
    got = conv._lowp(w, torch.bfloat16, channels_last=True)
    assert got.shape == w.shape and got.data_ptr() == store.data_ptr()
    assert got.contiguous(memory_format=torch.channels_last).data_ptr() == store.data_ptr()
    assert torch.equal(got.float(), w.detach().bfloat16().float())
    assert store.view(8, 3, 3, 4)[2, 1, 2, 3] == w.detach()[2, 3, 1, 2].bfloat16()
    # without a registered shadow the plain cast path is taken
    w2 = torch.nn.Parameter(torch.randn(8, 4, 3, 3))
    assert conv._lowp(w2, torch.bfloat16, channels_last=True).dtype == torch.bfloat16


def test_stream_schedule_switches_are_inert_on_cpu():
    from representationlearning_b200 import hrnet
    assert not hrnet._dataflow(torch.zeros(1))               # the data-flow schedule only engages for CUDA tensors
    t = [torch.zeros(2), torch.zeros(2)]
    assert not hasattr(t[0], "_rss_home")


def test_accum_chunk_rule_is_the_one_definition():
    """rss_accum_chunks (host function of the library) is the only definition of how rss_accum_bf16_list splits a table entry:
    4096-element chunks for same-order entries, whole (Cin*kk)-element rows for k x k entries whose row fits the staging buffer."""
    from representationlearning_b200 import _lib
    lib = _lib.load()
    for n in (1, 5, 4096, 4097, 70000):
        assert lib.rss_accum_chunks(n, 1, 0) == -(-n // 4096)
    for cout, cin, kk in ((32, 32, 9), (256, 256, 9), (480, 40, 9), (1, 2, 49), (64, 3, 9), (5, 24, 9), (3, 600, 9)):
        L, n = cin * kk, cout * cin * kk
        want = -(-cout // max(1, 4096 // L)) if L <= 4096 else -(-n // 4096)
        assert lib.rss_accum_chunks(n, cin, kk) == want, (cout, cin, kk)


def test_stem_entry_points_reject_bad_geometry_before_touching_the_device():
    """shape / dtype errors of the stem kernels are reported by the host side (no launch, so this runs without a GPU)"""
    from representationlearning_b200 import _lib
    lib = _lib.load()
    assert lib.rss_stem_conv_fwd(None, None, None, 0, 8, 8, 0, None, None, None) == -1          # RSS_ERR_SHAPE
    assert lib.rss_stem_conv_wgrad(None, None, None, 2, 0, 8, 0, None) == -1
    assert lib.rss_stem_conv_fwd(None, None, ctypes.c_void_p(8), 1, 8, 8, 0, None, None, None) == -1   # y not 16-byte aligned
