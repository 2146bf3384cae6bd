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



def earth_picture(w=48, h=24, seed=7):
    """The `earth` preset needs media/earthmap.jpg, which the reference tree does not ship: a synthetic RGB8 picture
    (stored inside the fixture) stands in for what RgbImage::open would have decoded."""
    r = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 120 * np.sin(x * 0.31 + 1.0), 127 + 120 * np.cos(y * 0.23), 127 + 120 * np.sin((x + y) * 0.11)], axis=2)
    return np.clip(base + r.integers(-6, 7, (h, w, 3)), 0, 255).astype(np.uint8)


only = set(sys.argv[1:])
if not only or "earth" in only:
    # Image textures get their (u, v) only in the SoA epilogue (spheres_soa.rs:141); the live list path passes 0, 0
    # (sphere.rs:44-45) and would paint one texel, so this fixture is an SoA-mode render
    w, h, s, d = 40, 20, 8, 50
    pic = earth_picture()
    img, rays = orc.Scene("earth", w, h, image=pic).update(s, d, mode=orc.HIT_SOA_SCALAR)
    name = "earth_%dx%d_s%d_d%d.npz" % (w, h, s, d)
    np.savez_compressed(os.path.join(HERE, name), preset="earth", width=w, height=h, samples=s, max_depth=d, rays=rays, image=img,
                        picture=pic, mode=orc.HIT_SOA_SCALAR)
    print(name, rays, img.mean(axis=(0, 1)))
for preset, w, h, s, d in CASES:
    if only and preset not in only:
        continue
    img, rays = orc.Scene(preset, w, h).update(s, d, mode=orc.HIT_LIST)
    name = "%s_%dx%d_s%d_d%d.npz" % (preset, w, h, s, d)
    np.savez_compressed(os.path.join(HERE, name), preset=preset, width=w, height=h, samples=s, max_depth=d, rays=rays, image=img)
    print(name, rays, img.mean(axis=(0, 1)))
