"""Which kernel should render a mid-size scene (2 000 - 13 000 spheres: fits in shared memory for the FP32 regroup kernel, too
large for the tensor-path regroup kernel at 2 CTAs/SM)?  python tools/midsize_bench.py <n_spheres> [spp]
Renders n Lambertian spheres on a ground sphere through the raw C ABI with (a) the FP32 regroup kernel, (b) the L2-streamed
kernel with the tensor-path pre-filter (forced tiles), (c) the library's automatic choice."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from pathtrace_rs_b200 import ffi
if len(sys.argv) > 3 and sys.argv[3]:
    ffi.LIB_DIR = os.path.join(ffi.LIB_DIR, sys.argv[3])  # a tools/build_variant.sh build
import pathtrace_rs_b200 as pt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w, h = 1200, 800
rng = np.random.default_rng(5)
side = int(np.ceil(np.sqrt(n - 1)))
gx, gz = np.meshgrid(np.arange(side) - side / 2, np.arange(side) - side / 2)
cr = np.zeros((n, 4), np.float32)
cr[0] = [0, -1000, 0, 1000]
cr[1:, 0] = (gx.ravel()[: n - 1] + 0.9 * rng.random(n - 1)) * 0.5
cr[1:, 1] = 0.1
cr[1:, 2] = (gz.ravel()[: n - 1] + 0.9 * rng.random(n - 1)) * 0.5
cr[1:, 3] = 0.1
L = ffi.libptgpu()
cols = [np.ascontiguousarray(cr[:, i]) for i in range(4)]
mats = (ffi.PtMaterial * n)(); texs = (ffi.PtTexture * n)()
for i in range(n):
    mats[i].kind, mats[i].texture = 0, i
    texs[i].kind, texs[i].odd, texs[i].even = 0, -1, -1
    texs[i].color[:] = [0.5, 0.5, 0.5]
midx = np.arange(n, dtype=np.int32)
d = ffi.PtSceneDesc(); d.struct_size, d.n_spheres = C.sizeof(ffi.PtSceneDesc), n
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
d.centre_x, d.centre_y, d.centre_z, d.radius = fp(cols[0]), fp(cols[1]), fp(cols[2]), fp(cols[3])
d.material_index = midx.ctypes.data_as(C.POINTER(C.c_int32)); d.n_materials = d.n_textures = n; d.materials, d.textures = mats, texs
cam = pt.Preset("random_spheres", pt.Params(w, h, 1, 50)).camera
p = pt.Params(w, h, spp, 50).to_ffi()
ref = None
for name, opt in (("FP32 regroup (resident_kernel 4)", pt.PtOptions(resident_kernel=4)), ("streamed, tensor path (256-block tiles)", pt.PtOptions(force_stream_tile_blocks=256)),
                  ("streamed, tensor path (128-block tiles)", pt.PtOptions(force_stream_tile_blocks=128)),
                  ("automatic", None)):
    scene = C.c_void_p(); dev = (C.c_int * 1)(0)
    ffi.check(L.pt_scene_create_multi(C.byref(d), dev, 1, C.byref(opt) if opt is not None else None, C.byref(scene)))
    buf = np.zeros((h, w, 3), np.float32); rays = C.c_uint64(0); best = 1e9
    for i in range(3):
        ffi.check(L.pt_render(scene, C.byref(p), C.byref(cam), 0, buf.ctypes.data_as(C.c_void_p), C.byref(rays)))
        st = ffi.PtRenderStats(); ffi.check(L.pt_scene_stats(scene, C.byref(st))); best = min(best, st.kernel_ms)
    same = "" if ref is None else (" image == first: %s" % np.array_equal(buf, ref))
    if ref is None: ref = buf.copy()
    print(f"n={n} {sys.argv[3] if len(sys.argv) > 3 else ''} {name}: kernel {best:.2f} ms {rays.value/1e3/best:.1f} Mrays/s resident={st.resident} grid {st.grid_ctas}x{st.cta_threads} smem {st.smem_bytes}{same}", flush=True)
    L.pt_scene_destroy(scene)
