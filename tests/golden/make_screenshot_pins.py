"""Builds tests/golden/readme_screenshots.npz from the reference's README screenshots (docs~/*.jpg).

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_screenshot_pins.py

The reference has no tests and no golden vectors; its README examples, however, are screenshots of the Unity editor
that show BOTH the rendered heightmap and the inspector panel with every parameter of the run (README.md:25-40,
docs~/0.jpg ... 6.jpg).  Each preview is cropped and resampled to 250 x 250 RGB here; tests/test_screenshot_pins.py
correlates the oracle's output for the panel's parameters with it.  This pins the oracle against the reference's own
OUTPUT at image precision (basis function, fBm law, position mapping, stage order); it is not a bit-level pin.

  file    panel (FBMSource: noise type, resolution, hurst, octaves, xpos, zpos, noise size)         preview box (l, t, r, b)
  3.jpg   Simplex 1000 0.422 13 0 424 1757, no filter enabled                                      (23, 40, 1118, 1135)
  4.jpg   same + Kernel Filter Gauss 5 x17                                                         (22, 27, 1117, 1122)
  5.jpg   same + Gauss 5 x17 + Erosion Filter x5 + Flow Map Filter x5 (blue = flow, red/green = terrain)   (25, 29, 1120, 1125)
  0.jpg   Cellular 1000 hurst 1 13 0 0 1757, no filter enabled                                     (24, 51, 1120, 1147)
The preview plane is mirrored left-right with respect to the array (Unity plane UVs).
"""
import os

import numpy as np
from PIL import Image

REF = "/root/reference/docs~"
BOXES = {"3": (23, 40, 1118, 1135), "4": (22, 27, 1117, 1122), "5": (25, 29, 1120, 1125), "0": (24, 51, 1120, 1147)}
S = 250

out = {}
for name, box in BOXES.items():
    im = Image.open(os.path.join(REF, name + ".jpg")).convert("RGB").crop(box).resize((S, S), Image.BILINEAR)
    out["shot" + name] = np.asarray(im, dtype=np.uint8)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "readme_screenshots.npz"), **out)
print({k: v.shape for k, v in out.items()})
