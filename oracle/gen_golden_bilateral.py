"""TEST INFRASTRUCTURE: golden vectors of the bilateral-filter row, produced by the REFERENCE ITSELF
(oracle/_ref/libbilateralfilter_ref.so = SCD-AAAI2023/wrapper/bilateralfilter/{bilateralfilter,permutohedral}.cpp compiled by
oracle/bilateral.py) on seeded inputs, and the pin of the C restatement against it.  Run in the authoring container:
    python -m oracle.gen_golden_bilateral
writes tests/golden/bilateral_cases.npz (inputs + reference outputs, small shapes) and tests/golden/BILATERAL_PIN.json
(bit-identity of oracle/bilateral_oracle.c with the reference on those and on larger shapes, lattice sizes)."""
import json
import os

import numpy as np

from oracle import bilateral as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (name, N, K, H, W, kind, sigma_rgb, sigma_xy, seed); H*W % 4 in {0, 1, 2, 3} all present (padding-lane quirk)
SMALL = [
    ("natural_16x16", 2, 3, 16, 16, "natural", 15.0, 50.0, 1),
    ("natural_13x11_mod3", 1, 2, 13, 11, "natural", 15.0, 5.0, 2),
    ("noise_9x7_mod3", 1, 2, 9, 7, "noise", 15.0, 3.0, 3),
    ("noise_5x5_mod1", 2, 2, 5, 5, "noise", 8.0, 2.0, 4),
    ("natural_6x7_mod2", 1, 3, 6, 7, "natural", 15.0, 4.0, 5),
    ("flat_32x32", 1, 3, 32, 32, "flat", 15.0, 50.0, 6),
    ("fine_31x33_mod3", 1, 2, 31, 33, "natural", 3.0, 1.0, 7),
    ("scd_40x40_K21", 2, 21, 40, 40, "natural", 15.0, 50.0, 8),
    ("wide_K33_24x20", 1, 33, 24, 20, "natural", 15.0, 10.0, 9),
    ("one_pixel", 1, 1, 1, 1, "noise", 15.0, 50.0, 10),
]
# pinned but not stored (inputs regenerate from the seed): the reference's working size, BASELINE cfg4's 448 crop at scale 0.5, a huge lattice
LARGE = [
    ("scd_voc_2x21x160x160", 2, 21, 160, 160, "natural", 15.0, 50.0, 11),
    ("cfg4_4x21x224x224", 4, 21, 224, 224, "natural", 15.0, 50.0, 12),
    ("noise_2x4x96x101", 2, 4, 96, 101, "noise", 15.0, 50.0, 13),
]


def main():
    B.build(verbose=True)
    assert B.have_reference(), "needs /root/reference (authoring container)"
    store, pin = {}, {}
    for name, N, K, H, W, kind, srgb, sxy, seed in SMALL + LARGE:
        img, seg = B.synth(N, K, H, W, seed=seed, kind=kind)
        ref = B.reference_filter(img, seg, srgb, sxy)
        ora, m = B.oracle_filter(img, seg, srgb, sxy, want_lattice=True)
        same = bool(np.array_equal(ref.view(np.uint32), ora.view(np.uint32)))
        pin[name] = {"shape": [N, K, H, W], "kind": kind, "sigma_rgb": srgb, "sigma_xy": sxy, "seed": seed, "lattice_points": m.tolist(),
                     "oracle_bit_identical_to_reference": same, "max_abs_diff": float(np.abs(ref - ora).max()),
                     "ref_checksum": float(ref.astype(np.float64).sum()), "ref_crc": int(np.bitwise_xor.reduce(ref.view(np.uint32).ravel()))}
        assert same, name
        if (name, N, K, H, W, kind, srgb, sxy, seed) in SMALL:
            store[name + "/img"], store[name + "/seg"], store[name + "/out"] = img, seg, ref
            store[name + "/sigma"] = np.array([srgb, sxy], np.float32)
        print(name, "M", m.tolist(), "bit-identical", same)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bilateral_cases.npz"), **store)
    with open(os.path.join(ROOT, "tests", "golden", "BILATERAL_PIN.json"), "w") as f:
        json.dump(pin, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
