"""PT_PERLIN_SELECT validation: the select form of the Perlin trilinear weights (default) against the product form
(lib/noselect, built with -DPT_PERLIN_SELECT=0): bit-identical images, and what it buys.  python tools/select_check.py"""
import os, sys, subprocess
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
code = r'''
import sys, os, hashlib
sys.path.insert(0, %r)
from pathtrace_rs_b200 import ffi
if sys.argv[1]: ffi.LIB_DIR = os.path.join(ffi.LIB_DIR, sys.argv[1])
import pathtrace_rs_b200 as pt
for w, h, spp in ((480, 270, 32), (1920, 1080, 64)):
    params = pt.Params(w, h, spp, 50)
    pr = pt.Preset("two_perlin_spheres", params).create_scene(0)
    best = 1e9
    for i in range(3):
        img, rays = pr.update()
        best = min(best, pr.stats().kernel_ms)
    print("%%-10s two_perlin_spheres %%dx%%d spp%%d: %%.2f ms %%.1f Mrays/s rays %%d sha %%s" %% (sys.argv[1] or "select", w, h, spp, best, rays / 1e3 / best, rays, hashlib.sha256(img.tobytes()).hexdigest()[:16]))
''' % ROOT
for v in ("", "noselect"):
    subprocess.run([sys.executable, "-c", code, v])
