"""Exact mismatch counts of the GPU images against the oracle's SoA mode for the parity configs (development aid: the numbers
pinned in tests/test_gpu_parity.py come from here).  python tools/parity_report.py"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
import pathtrace_rs_b200 as pt
SOA_ITER = orc.HIT_SOA_SCALAR | 0x100
for preset, w, h, spp, depth in [("small", 100, 50, 16, 10), ("smallpt", 64, 64, 16, 10), ("final", 64, 32, 4, 10), ("small", 37, 23, 5, 0),
                                 ("random_spheres", 200, 100, 100, 50), ("random", 160, 80, 32, 50), ("two_perlin_spheres", 192, 108, 16, 50),
                                 ("stress100k", 64, 36, 2, 50), ("random_spheres", 96, 54, 21, 50)]:
    params = pt.Params(w, h, spp, depth)
    pr = pt.Preset(preset, params).create_scene(0)
    img, rays = pr.update(params)
    ref, ref_rays = orc.Scene(preset, w, h).update(spp, depth, mode=SOA_ITER)
    diff = np.abs(img - ref)
    npx = int(np.any(img != ref, axis=2).sum())
    print(f"{preset:20s} {w}x{h} spp{spp} d{depth}: rays gpu {rays} oracle {ref_rays} (delta {rays-ref_rays}); pixels differing {npx} of {w*h}; max abs diff {diff.max():.3e}; pixels > 1e-5: {int(np.any(diff>1e-5,axis=2).sum())}; > 1e-6: {int(np.any(diff>1e-6,axis=2).sum())}", flush=True)
