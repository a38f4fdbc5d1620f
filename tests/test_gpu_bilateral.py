"""GPU parity of the permutohedral bilateral filter (csrc/bilateral.cu) through the C ABI: BIT-IDENTICAL to the golden vectors the
reference produced, to the C restatement at the reference's working sizes, to the compiled reference when oracle/_ref holds it;
the host entry point with the reference's argument list; DenseEnergyLoss forward/backward (SCD-AAAI2023/utils/losses.py:52-120)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import bilateral as B

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _dev_filter(img, seg, srgb, sxy, want_m=False):
    from representationlearning_b200 import scd
    m = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = scd.bilateral_filter(torch.from_numpy(img).cuda(), torch.from_numpy(seg).cuda(), srgb, sxy, lattice_points=m)
    torch.cuda.synchronize()
    return (out.cpu().numpy(), int(m.item())) if want_m else out.cpu().numpy()


def test_device_filter_bit_identical_to_reference_golden(report):
    z = np.load(os.path.join(GOLDEN, "bilateral_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    pin = json.load(open(os.path.join(GOLDEN, "BILATERAL_PIN.json")))
    for n in names:
        img, seg, out, sg = z[n + "/img"], z[n + "/seg"], z[n + "/out"], z[n + "/sigma"]
        got, m = _dev_filter(img, seg, float(sg[0]), float(sg[1]), want_m=True)
        diff = float(np.abs(got - out).max())
        report["bilateral/" + n] = {"max_abs_diff": diff, "lattice_points": m}
        assert m == sum(pin[n]["lattice_points"]), (n, m)
        assert np.array_equal(_bits(got), _bits(out)), (n, diff)


@pytest.mark.parametrize("name", ["scd_voc_2x21x160x160", "cfg4_4x21x224x224", "noise_2x4x96x101"])
def test_device_filter_at_working_sizes(name, report):
    v = json.load(open(os.path.join(GOLDEN, "BILATERAL_PIN.json")))[name]
    N, K, H, W = v["shape"]
    img, seg = B.synth(N, K, H, W, seed=v["seed"], kind=v["kind"])
    got, m = _dev_filter(img, seg, v["sigma_rgb"], v["sigma_xy"], want_m=True)
    ora = B.oracle_filter(img, seg, v["sigma_rgb"], v["sigma_xy"])
    report["bilateral/" + name] = {"max_abs_diff": float(np.abs(got - ora).max()), "lattice_points": m}
    assert m == sum(v["lattice_points"])
    assert int(np.bitwise_xor.reduce(_bits(got).ravel())) == v["ref_crc"]          # checksum of the REFERENCE's output
    assert np.array_equal(_bits(got), _bits(ora))
    if B.have_reference():
        assert np.array_equal(_bits(got), _bits(B.reference_filter(img, seg, v["sigma_rgb"], v["sigma_xy"])))


def test_device_filter_properties_full_size():
    """size-independent properties at BASELINE cfg4's geometry (448 crop at scale 0.5, 8 images): run-to-run bit reproducibility
    (no float atomics), exact power-of-two scaling, plane independence"""
    img, seg = B.synth(8, 21, 224, 224, seed=31)
    a = _dev_filter(img, seg, 15.0, 50.0)
    b = _dev_filter(img, seg, 15.0, 50.0)
    assert np.array_equal(_bits(a), _bits(b))
    c = _dev_filter(img, 0.5 * seg, 15.0, 50.0)
    assert np.array_equal(_bits(0.5 * a), _bits(c))
    d = _dev_filter(img, np.ascontiguousarray(seg[:, 3:9]), 15.0, 50.0)
    assert np.array_equal(_bits(d), _bits(a[:, 3:9]))


def test_host_entry_point_with_the_reference_argument_list():
    from representationlearning_b200 import scd
    N, K, H, W = 2, 5, 33, 29
    img, seg = B.synth(N, K, H, W, seed=17)
    images, ins = img.flatten(), seg.flatten()
    AS = np.zeros(ins.shape, dtype=np.float32)
    scd.bilateralfilter_batch(images, ins, AS, N, K, H, W, 15.0, 25.0)             # the call of utils/losses.py:70
    ora = B.oracle_filter(img, seg, 15.0, 25.0)
    assert np.array_equal(_bits(AS), _bits(ora.ravel()))
    with pytest.raises(Exception):
        scd.bilateralfilter_batch(images[:-1].copy(), ins, AS, N, K, H, W, 15.0, 25.0)   # length check


def test_dense_energy_loss_forward_backward(report):
    """DenseEnergyLoss against a restatement of utils/losses.py:52-120 that filters with the CPU oracle"""
    from representationlearning_b200 import scd
    import torch.nn.functional as F
    torch.manual_seed(3)
    N, K, S = 2, 21, 64
    img_np, _ = B.synth(N, K, S, S, seed=9)
    images = torch.from_numpy(img_np)
    logits = torch.randn(N, K, S, S)
    rois = (torch.rand(N, S, S) > 0.1).float()
    label = torch.randint(0, K, (N, 1, S, S)).float()
    label[torch.rand(N, 1, S, S) > 0.8] = 255.0
    weight, srgb, sxy, sf = 1e-7, 15.0, 100.0, 0.5

    def reference_loss(seg):
        si = F.interpolate(images, scale_factor=sf)
        ss = F.interpolate(seg, scale_factor=sf, mode="bilinear", align_corners=False)
        sr = F.interpolate(rois.unsqueeze(1), scale_factor=sf).squeeze(1)
        sl = F.interpolate(label, scale_factor=sf, mode="nearest")
        unl = (sl.long() == 255).squeeze(1)
        gate = sr.clone() - ss.max(dim=1)[0]
        gate[unl] = 1
        gate[gate < 0] = 0
        masked = (ss * sr.unsqueeze(1)).detach()
        AS = torch.from_numpy(B.oracle_filter(si.numpy(), masked.numpy(), srgb, sxy * sf)) * gate.unsqueeze(1)
        loss = -(masked.double() * AS.double()).sum() / N
        grad_scaled = -2.0 * AS / N * sr.unsqueeze(1)                                # what the reference's backward returns
        return weight * loss, AS, grad_scaled, ss

    seg_ref = torch.softmax(logits, 1).requires_grad_(True)
    l_ref, AS_ref, g_scaled, ss = reference_loss(seg_ref.detach())
    ss_leaf = F.interpolate(seg_ref, scale_factor=sf, mode="bilinear", align_corners=False)
    ss_leaf.backward(weight * g_scaled)                                             # chain through the interpolation
    seg = torch.softmax(logits, 1).cuda().requires_grad_(True)
    layer = scd.DenseEnergyLoss(weight=weight, sigma_rgb=srgb, sigma_xy=sxy, scale_factor=sf)
    loss = layer(images.cuda(), seg, rois.cuda(), label.cuda())
    loss.backward()
    e_loss = abs(loss.item() - float(l_ref)) / abs(float(l_ref))
    e_grad = ((seg.grad.cpu() - seg_ref.grad).abs().max() / seg_ref.grad.abs().max()).item()
    report["bilateral/dense_energy_loss"] = {"loss_rel": e_loss, "grad_rel": e_grad}
    assert loss.shape == (1,) and e_loss < 1e-5 and e_grad < 1e-5, (e_loss, e_grad)
