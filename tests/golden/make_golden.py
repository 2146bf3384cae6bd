"""Generates the regression fixtures in this directory from the CPU oracle (list mode = the reference's live path).

The reference itself cannot be run here (Rust, no toolchain), so these are NOT reference outputs: they pin the
oracle against accidental drift and give the GPU tests fixed vectors to be compared with.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

CASES = [("random_spheres", 40, 20, 8, 50), ("two_perlin_spheres", 40, 20, 4, 50), ("small", 40, 20, 8, 10),
         ("smallpt", 32, 32, 16, 10), ("random", 40, 20, 8, 50)]

only = set(sys.argv[1:])
for preset, w, h, s, d in CASES:
    if only and preset not in only:
        continue
    img, rays = orc.Scene(preset, w, h).update(s, d, mode=orc.HIT_LIST)
    name = "%s_%dx%d_s%d_d%d.npz" % (preset, w, h, s, d)
    np.savez_compressed(os.path.join(HERE, name), preset=preset, width=w, height=h, samples=s, max_depth=d, rays=rays, image=img)
    print(name, rays, img.mean(axis=(0, 1)))
