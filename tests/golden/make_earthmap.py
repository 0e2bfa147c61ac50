"""Decodes the reference's only data fixture, /root/reference/earthmap.jpg (1024x512 RGB JPEG, used by the commented
earth() / final_scene(): raytracer/src/main.rs:281,601), ONCE with PIL (libjpeg) and stores the raw RGB8 texels as
tests/golden/earthmap_rgb8.npz, because /root/reference does not exist on the GPU box.  The reference decodes with the
`image ^0.23` crate; decoder differences are <= 1 LSB per texel.

    python tests/golden/make_earthmap.py [/root/reference/earthmap.jpg]
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/earthmap.jpg"
img = np.asarray(Image.open(src).convert("RGB"), dtype=np.uint8)
assert img.shape == (512, 1024, 3), img.shape
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "earthmap_rgb8.npz")
np.savez_compressed(out, rgb=img, source_sha256=hashlib.sha256(open(src, "rb").read()).hexdigest())
print(out, os.path.getsize(out), "bytes; texel sha256", hashlib.sha256(img.tobytes()).hexdigest())
