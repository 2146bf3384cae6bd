"""One small render per kernel flavour, for compute-sanitizer (tools/gpu_sanitize.sh).  python tools/sanitize_case.py <case>"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import pathtrace_rs_b200 as pt
O = pt.PtOptions
CASES = {  # name: (preset, w, h, spp, options, devices)
    "regroup": ("random_spheres", 64, 32, 4, O(resident_kernel=4), 0),
    "regroup_motion": ("random", 64, 32, 4, O(resident_kernel=4), 0),
    "mma": ("random_spheres", 64, 32, 4, None, 0),  # the default: regroup kernel with the tensor-path pre-filter
    "mma_motion": ("random", 64, 32, 4, O(resident_kernel=5), 0),
    "mma_chunked": ("random_spheres", 200, 120, 16, O(resident_kernel=5, chunk_samples=4), 0),
    "noise": ("two_perlin_spheres", 64, 32, 4, None, 0),
    "image": ("earth", 64, 32, 4, None, 0),
    "streamed": ("random_spheres", 64, 32, 2, O(force_stream_tile_blocks=16), 0),
    "streamed_motion": ("random", 64, 32, 2, O(force_stream_tile_blocks=16), 0),
    "chunked": ("random_spheres", 200, 120, 16, O(chunk_samples=4), 0),
    "pair_const": ("random_spheres", 64, 32, 4, O(resident_kernel=2), 0),
    "pair_lds": ("random", 64, 32, 4, O(resident_kernel=3), 0),
    "wave": ("random_spheres", 96, 48, 4, O(resident_kernel=1), 0),
    "wave_motion_chunked": ("random", 640, 400, 16, O(resident_kernel=1, chunk_samples=8), 0),
    "multi": ("random_spheres", 64, 37, 4, O(tile_rows=3), [0, 0, 0]),
}
name = sys.argv[1]
if name == "image":
    pt.write_ppm("/tmp/earth.ppm", (np.arange(32 * 16 * 3) % 251).astype(np.uint8).reshape(16, 32, 3))
    os.environ["PATHTRACE_EARTHMAP"] = "/tmp/earth.ppm"
if name == "debug_hits":
    pr = pt.Preset("random", pt.Params(64, 32, 1, 5)).create_scene(0)
    rng = np.random.default_rng(0)
    d = rng.normal(size=(5000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.hstack([np.tile([13.0, 2.0, 3.0], (5000, 1)), -np.abs(d)]).astype(np.float32)
    for opt in (None, O(resident_kernel=4), O(force_stream_tile_blocks=16), O(resident_kernel=2)):
        p2 = pt.Preset("random", pt.Params(64, 32, 1, 5)).create_scene(0, opt)
        for mode in (0, 1):
            idx, t = p2.debug_hits(rays, times=rng.random(5000).astype(np.float32), mode=mode)
    print("case debug_hits ok", int((idx >= 0).sum()))
    sys.exit(0)
preset, w, h, spp, opt, dev = CASES[name]
params = pt.Params(w, h, spp, 50)
pr = pt.Preset(preset, params).create_scene(dev, opt)
img, rays = pr.update(params)
img2 = img.copy()
pr.update(params, frame_num=1, buffer=img2)
print("case %s ok: rays %d mean %.6f" % (name, rays, img.mean()))
