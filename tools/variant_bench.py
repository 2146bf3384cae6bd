"""Times kernel variants built into pathtrace_rs_b200/lib/<variant>/ (development aid)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pathtrace_rs_b200 import ffi
variant = sys.argv[1] if len(sys.argv) > 1 else ""
if variant:
    ffi.LIB_DIR = os.path.join(ffi.LIB_DIR, variant)
import pathtrace_rs_b200 as pt
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
for preset, w, h, s in (("random_spheres", 1200, 800, spp),):
    params = pt.Params(w, h, s, 50)
    pr = pt.Preset(preset, params).create_scene(0)
    best = 1e9
    for i in range(3):
        img, rays = pr.update()
        best = min(best, pr.stats().kernel_ms)
    n = len(pr)
    print(f"variant '{variant}' {preset} {w}x{h} spp{s}: kernel {best:.2f} ms {rays/1e6/(best*1e-3):.1f} Mrays/s {rays*16*n/(best*1e-3)/74.45e12*100:.1f}% of FP32 peak  mean {img.mean():.6f}")
