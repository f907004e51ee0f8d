#!/usr/bin/env python
"""Regenerates tests/golden/earthmap_rgb.npz: content/earthmap.jpg decoded by the reference's own vendored
stb_image v2.28 (the decoder zstbi wraps), so the texels the ImageTexture path reads are bit-identical to the
reference's.  Needs /root/reference (this container only); the fixture it writes travels with the repo."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
out_dir = os.path.join(ROOT, "oracle", "_ref")
os.makedirs(out_dir, exist_ok=True)
exe = os.path.join(out_dir, "ref_stb_decode")
subprocess.run(["gcc", "-O2", "-I", os.path.join(REF, "libs/zstbi/libs/stbi"), "-o", exe,
                os.path.join(ROOT, "oracle", "ref_stb_decode.c"), "-lm"], check=True)
raw = os.path.join(out_dir, "earthmap.rgba")
subprocess.run([exe, os.path.join(REF, "content/earthmap.jpg"), raw], check=True)
with open(raw, "rb") as f:
    w, h, c = map(int, f.readline().split())
    data = np.frombuffer(f.read(), dtype=np.uint8).reshape(h, w, 4)
assert (data[:, :, 3] == 255).all()
dst = os.path.join(ROOT, "tests", "golden", "earthmap_rgb.npz")
np.savez_compressed(dst, rgb=np.ascontiguousarray(data[:, :, :3]))
print(f"{w}x{h} ({c} source channels) -> {dst} ({os.path.getsize(dst)} bytes)")
