#!/usr/bin/env python
"""Regenerates tests/golden/ref_renders_top160.npz: the top 160 rows of the two 800 x 450 PNG renders the reference
repository commits at its root (image.png: Book-1 final scene with moving spheres; image2.png: the canonical Book-1
final scene).  They are OUTPUTS OF THE REFERENCE PROGRAM ITSELF — the only ones there are (no Zig toolchain here, no
headless mode in the reference).  The small spheres of those scenes were placed by an unseeded generator, but above
the horizon the frame shows only things the reference fixes: the sky gradient (src/camera.zig:204-206), the
silhouettes of the three big spheres (src/main.zig:303-309) and the sky reflected by the metal one — which is what
tests/test_reference_renders.py holds the oracle and the CUDA path against.
Needs /root/reference and PIL (this container only); the fixture it writes travels with the repo."""
import os

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
ROWS = 160
out = {}
for key, name in (("image_png", "image.png"), ("image2_png", "image2.png")):
    im = np.asarray(Image.open(os.path.join(REF, name)).convert("RGB"))
    assert im.shape == (450, 800, 3), im.shape
    out[key] = np.ascontiguousarray(im[:ROWS])
dst = os.path.join(ROOT, "tests", "golden", "ref_renders_top160.npz")
np.savez_compressed(dst, **out)
print(f"{dst}: {os.path.getsize(dst)} bytes")
