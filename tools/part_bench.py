"""Kernel time of one row-tile part of a preset (development aid): python tools/part_bench.py preset w h spp part_count [chunk]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import pathtrace_rs_b200 as pt
from pathtrace_rs_b200 import ffi
preset, w, h, spp, parts = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
params = pt.Params(w, h, spp, 50)
pr = pt.Preset(preset, params).create_scene(0)
for i in range(2):
    img, rays = pr.update(params, part=ffi.PtPartition(4, 0, parts, 0) if parts > 1 else None)
    st = pr.stats()
    print(f"{preset} {w}x{h} spp{spp} part 0/{parts} env {os.environ.get('PTGPU_CHUNK_SAMPLES')}: kernel {st.kernel_ms:.1f} ms rays {rays} {rays/1e3/st.kernel_ms:.2f} Mrays/s grid {st.grid_ctas}", flush=True)
