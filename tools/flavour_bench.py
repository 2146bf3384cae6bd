"""Times one resident-kernel flavour of one library variant: python tools/flavour_bench.py <variant|''> <flavour> [spp] [preset] [w] [h]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pathtrace_rs_b200 import ffi
variant = sys.argv[1] if len(sys.argv) > 1 else ""
if variant:
    ffi.LIB_DIR = os.path.join(ffi.LIB_DIR, variant)
import pathtrace_rs_b200 as pt
rk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 256
preset = sys.argv[4] if len(sys.argv) > 4 else "random_spheres"
w = int(sys.argv[5]) if len(sys.argv) > 5 else 1200
h = int(sys.argv[6]) if len(sys.argv) > 6 else 800
params = pt.Params(w, h, spp, 50)
pr = pt.Preset(preset, params).create_scene(0, pt.PtOptions(resident_kernel=rk))
best = 1e9
for i in range(3):
    img, rays = pr.update()
    best = min(best, pr.stats().kernel_ms)
n, st = len(pr), pr.stats()
print(f"variant '{variant}' flavour {rk} {preset} n={n} {w}x{h} spp{spp}: kernel {best:.2f} ms {rays/1e6/(best*1e-3):.1f} Mrays/s {rays*16*n/(best*1e-3)/74.45e12*100:.1f}% of FP32 peak  mean {img.mean():.6f} rays {rays} grid {st.grid_ctas}x{st.cta_threads} smem {st.smem_bytes} lane_eff {rays/32/max(1,st.warp_sweeps):.3f}", flush=True)
