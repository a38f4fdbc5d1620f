"""Writes a synthetic LoveDA-shaped directory tree (BASELINE cfg1: "one synthetic 512x512 tile"): ./LoveDA/{Train,Val}/{Urban,Rural}/
{images_png,masks_png}/<i>.png with 1024x1024 RGB tiles (LoveDA's tile size; the training transform crops 512x512) and masks in
0..7 (0 = no-data -> ignore_index -1 after data/loveda.py:84's "-1").   python tools/make_synth_loveda.py DIR [n_per_split] [size]"""
import os
import sys

import numpy as np
from PIL import Image


def make(root, n=1, size=1024, seed=7):
    rng = np.random.default_rng(seed)
    for split in ("Train", "Val"):
        for dom in ("Urban", "Rural"):
            for sub in ("images_png", "masks_png"):
                os.makedirs(os.path.join(root, "LoveDA", split, dom, sub), exist_ok=True)
            for i in range(n):
                img = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
                blocks = rng.integers(0, 8, (size // 64, size // 64), dtype=np.uint8)
                mask = np.kron(blocks, np.ones((64, 64), dtype=np.uint8))
                Image.fromarray(img).save(os.path.join(root, "LoveDA", split, dom, "images_png", "%d.png" % i))
                Image.fromarray(mask).save(os.path.join(root, "LoveDA", split, dom, "masks_png", "%d.png" % i))


if __name__ == "__main__":
    make(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1, int(sys.argv[3]) if len(sys.argv) > 3 else 1024)
