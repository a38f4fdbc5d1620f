"""CPU tests of the bilateral-filter row: the C restatement against the golden vectors the REFERENCE produced,
against the compiled reference itself when it is present, and domain properties of the algorithm.  No GPU, no product code."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import bilateral as B


def _cases():
    z = np.load(os.path.join(GOLDEN, "bilateral_cases.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_oracle_bit_identical_to_reference_golden():
    z, names = _cases()
    assert len(names) >= 10
    for n in names:
        img, seg, out, sg = z[n + "/img"], z[n + "/seg"], z[n + "/out"], z[n + "/sigma"]
        got = B.oracle_filter(img, seg, float(sg[0]), float(sg[1]))
        assert np.array_equal(_bits(got), _bits(out)), n


def test_pin_report_says_bit_identical_everywhere():
    pin = json.load(open(os.path.join(GOLDEN, "BILATERAL_PIN.json")))
    assert len(pin) >= 13 and all(v["oracle_bit_identical_to_reference"] and v["max_abs_diff"] == 0.0 for v in pin.values())
    # every residue of H*W modulo 4 is covered (the SSE padding-lane quirk)
    assert {(v["shape"][2] * v["shape"][3]) % 4 for v in pin.values()} == {0, 1, 2, 3}


def test_oracle_lattice_sizes_and_checksums_match_pin():
    pin = json.load(open(os.path.join(GOLDEN, "BILATERAL_PIN.json")))
    for name in ("scd_voc_2x21x160x160", "noise_2x4x96x101"):
        v = pin[name]
        N, K, H, W = v["shape"]
        img, seg = B.synth(N, K, H, W, seed=v["seed"], kind=v["kind"])
        out, m = B.oracle_filter(img, seg, v["sigma_rgb"], v["sigma_xy"], want_lattice=True)
        assert m.tolist() == v["lattice_points"]
        assert int(np.bitwise_xor.reduce(_bits(out).ravel())) == v["ref_crc"], name


@pytest.mark.skipif(not B.have_reference(), reason="oracle/_ref/libbilateralfilter_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("shape,kind,sig", [((3, 5, 37, 41), "natural", (15.0, 50.0)), ((1, 2, 64, 64), "noise", (15.0, 50.0)),
                                            ((2, 7, 50, 30), "natural", (4.0, 3.0)), ((1, 4, 48, 48), "flat", (15.0, 50.0))])
def test_oracle_vs_compiled_reference(shape, kind, sig):
    img, seg = B.synth(*shape, seed=21, kind=kind)
    ref = B.reference_filter(img, seg, *sig)
    got = B.oracle_filter(img, seg, *sig)
    assert np.array_equal(_bits(got), _bits(ref))


def test_filter_properties():
    """scaling by a power of two is exact; planes are independent; a flat guide image with flat input gives a flat interior"""
    img, seg = B.synth(1, 4, 24, 28, seed=5)
    a = B.oracle_filter(img, seg, 15.0, 20.0)
    b = B.oracle_filter(img, 4.0 * seg, 15.0, 20.0)
    assert np.array_equal(_bits(4.0 * a), _bits(b))
    c = B.oracle_filter(img, seg[:, 1:3], 15.0, 20.0)
    assert np.array_equal(_bits(c), _bits(a[:, 1:3]))
    # superposition holds to rounding
    d = B.oracle_filter(img, seg[:, :1] + seg[:, 1:2], 15.0, 20.0)
    assert np.abs(d[:, 0] - (a[:, 0] + a[:, 1])).max() <= 1e-5 * np.abs(d).max()
    assert np.all(a >= 0.0)                      # non-negative weights on non-negative input


def test_abi_declares_the_reference_entry_point():
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "rss_b200.h")).read()
    assert "bilateralfilter.hpp:12" in hdr and "rss_bilateralfilter_batch_host" in hdr
    from representationlearning_b200 import _lib
    args = _lib.SIGNATURES["rss_bilateralfilter_batch_host"][1]
    assert len(args) == 12                       # the reference's argument list verbatim
