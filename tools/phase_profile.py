"""Phase split of the megakernel (refill / sweep / shade) from the -DPT_PROFILE build.  Development aid.
   make -C pathtrace_rs_b200 lib/libptgpu_prof.so && python tools/phase_profile.py"""
import ctypes as C, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pathtrace_rs_b200 import ffi
# swap in the profile build before anything loads the product library
prof = os.path.join(ffi.LIB_DIR, "prof")
ffi.LIB_DIR = prof
import pathtrace_rs_b200 as pt
import numpy as np

def run(preset, w, h, spp, depth):
    params = pt.Params(w, h, spp, depth)
    pr = pt.Preset(preset, params).create_scene(0)
    L = pt.libptgpu()
    buf = (C.c_ulonglong * 8)()
    L.pt_profile_read(buf)
    img, rays = pr.update()
    st = pr.stats()
    L.pt_profile_read(buf)
    refill, sweep, shade, trips, lanes, total = [int(buf[i]) for i in range(6)]
    tot = refill + sweep + shade
    print(f"{preset} {w}x{h} spp{spp}: kernel {st.kernel_ms:.2f} ms rays {rays} | warp trips {trips} lane efficiency {lanes / (32.0 * trips):.3f} | "
          f"phase share of warp-clocks: refill {refill / tot:.3f} sweep {sweep / tot:.3f} shade {shade / tot:.3f} | clk/trip {tot / trips:.0f} (sweep {sweep / trips:.0f})")

run("random_spheres", 1200, 800, 64, 50)
run("random_spheres", 1200, 800, 256, 50)
run("two_perlin_spheres", 960, 540, 64, 50)
