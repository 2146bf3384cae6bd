"""ctypes binding of the CPU oracle (oracle/pt_oracle.cpp).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "_build", "liboracle.so")

HIT_LIST, HIT_SOA_SCALAR, HIT_SOA_AVX2 = 0, 1, 2
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT = 0, 1, 2, 3
TEX_CONSTANT, TEX_CHECKER, TEX_NOISE, TEX_IMAGE = 0, 1, 2, 3


class OrcParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32),
                ("max_depth", C.c_uint32), ("random_seed", C.c_uint32), ("use_bvh", C.c_uint32)]


def build(force=False):
    src = os.path.join(_ORACLE_DIR, "pt_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_scene_build.restype = C.c_void_p
        L.orc_scene_build.argtypes = [C.c_char_p, C.POINTER(OrcParams)]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_scene_custom.restype = C.c_void_p
        L.orc_scene_custom.argtypes = [C.c_int32] + [C.c_void_p] * 6
        L.orc_update.restype = C.c_uint64
        L.orc_update.argtypes = [C.c_void_p, C.POINTER(OrcParams), C.c_uint32, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_uint32, C.c_uint32]
        for name in ("orc_scene_counts", "orc_scene_spheres", "orc_scene_materials", "orc_scene_textures",
                     "orc_scene_perlin", "orc_scene_camera", "orc_scene_sky", "orc_scene_motion"):
            getattr(L, name).argtypes = [C.c_void_p] + [C.c_void_p] * {"orc_scene_counts": 3, "orc_scene_camera": 1, "orc_scene_motion": 1}.get(name, 2)
        L.orc_rng_seed.argtypes = [C.c_uint64, C.c_void_p]
        L.orc_rng_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_rng_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_sincos.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_turb.restype = C.c_float
        L.orc_turb.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.orc_noise.restype = C.c_float
        L.orc_noise.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.orc_tex_value.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_srgb.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_set_earth_image.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_scene_add_image.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_scene_add_image_texture.argtypes = [C.c_void_p, C.c_int32]
        L.orc_scene_add_checker_texture.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_scene_set_sphere_texture.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_scene_image_count.argtypes = [C.c_void_p]
        L.orc_scene_image.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_image_value.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p]
        L.orc_sphere_uv.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.orc_hit.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_bench_fixture_ray.argtypes = [C.POINTER(OrcParams), C.c_void_p]
        L.orc_record_rays.restype = C.c_int64
        L.orc_record_rays.argtypes = [C.c_void_p, C.POINTER(OrcParams), C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32]
        L.orc_hit_times.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_next_f32_after_random_spheres.restype = C.c_float
        L.orc_hw_threads.restype = C.c_int32
        L.orc_has_avx2.restype = C.c_int32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def params(width, height, samples, max_depth):
    return OrcParams(width, height, samples, max_depth, 0, 0)


class Scene:
    """A preset built by the oracle's restatement of presets.rs (scene rng seed 0)."""

    def __init__(self, preset, width, height, samples=1, max_depth=50, custom=None, image=None):
        """image: uint8 [h, w, 3] (row 0 = top) — what RgbImage::open would decode for the `earth` preset."""
        self.preset = preset
        if image is not None:
            image = np.ascontiguousarray(image, np.uint8)
            lib().orc_set_earth_image(image.shape[1], image.shape[0], _p(image))
        self.p = params(width, height, samples, max_depth)
        if custom is not None:
            # custom = dict(centre_radius[n,4], kind[n], params5[n,5], motion[n,6] or None, cam15[15], sky[3] or None)
            self._keep = {k: (None if v is None else np.ascontiguousarray(v, np.int32 if k == "kind" else np.float32)) for k, v in custom.items()}
            k = self._keep
            ptr = lambda a: None if a is None else _p(a)
            self.h = lib().orc_scene_custom(len(k["kind"]), ptr(k["centre_radius"]), ptr(k["kind"]), ptr(k["params5"]), ptr(k.get("motion")),
                                            ptr(k["cam15"]), ptr(k.get("sky")))
            return
        self.h = lib().orc_scene_build(preset.encode(), C.byref(self.p))
        if not self.h:
            raise ValueError("unrecognised preset " + preset)

    def close(self):
        if self.h:
            lib().orc_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Image textures (texture.rs:6-37) on a custom scene ----
    def add_image(self, image):
        image = np.ascontiguousarray(image, np.uint8)
        return lib().orc_scene_add_image(self.h, image.shape[1], image.shape[0], _p(image))

    def add_image_texture(self, image_index):
        return lib().orc_scene_add_image_texture(self.h, image_index)

    def add_checker_texture(self, odd, even):
        return lib().orc_scene_add_checker_texture(self.h, odd, even)

    def set_sphere_texture(self, sphere, tex):
        lib().orc_scene_set_sphere_texture(self.h, sphere, tex)

    def images(self):
        out = []
        for i in range(lib().orc_scene_image_count(self.h)):
            w, h = C.c_uint32(), C.c_uint32()
            lib().orc_scene_image(self.h, i, C.byref(w), C.byref(h), None)
            px = np.zeros((h.value, w.value, 3), np.uint8)
            lib().orc_scene_image(self.h, i, C.byref(w), C.byref(h), _p(px))
            out.append(px)
        return out

    def image_value(self, image_index, u, v):
        out = np.zeros(3, np.float32)
        lib().orc_image_value(self.h, image_index, u, v, _p(out))
        return out

    def counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().orc_scene_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def flat(self):
        """dict of numpy arrays in the same logical layout as the product's PtSceneDesc."""
        ns, nm, nt = self.counts()
        cr = np.zeros((ns, 4), np.float32)
        mat = np.zeros(ns, np.int32)
        lib().orc_scene_spheres(self.h, _p(cr), _p(mat))
        mk = np.zeros((nm, 2), np.int32)
        mf = np.zeros((nm, 5), np.float32)
        lib().orc_scene_materials(self.h, _p(mk), _p(mf))
        tk = np.zeros((nt, 3), np.int32)
        tf = np.zeros((nt, 4), np.float32)
        lib().orc_scene_textures(self.h, _p(tk), _p(tf))
        rv = np.zeros((256, 3), np.float32)
        perm = np.zeros((3, 256), np.uint32)
        lib().orc_scene_perlin(self.h, _p(rv), _p(perm))
        cam = np.zeros(24, np.float32)
        lib().orc_scene_camera(self.h, _p(cam))
        hs = C.c_int32()
        sky = np.zeros(3, np.float32)
        lib().orc_scene_sky(self.h, C.byref(hs), _p(sky))
        motion = np.zeros((ns, 6), np.float32)  # centre1 (3), time0, time1, moving flag
        lib().orc_scene_motion(self.h, _p(motion))
        return dict(centre_radius=cr, sphere_material=mat, motion=motion, mat_kind_tex=mk, mat_albedo_fuzz_ref=mf,
                    tex_kind_odd_even=tk, tex_color_scale=tf, randvec=rv, perm=perm, camera=cam,
                    has_sky=int(hs.value), sky=sky)

    def update(self, samples, max_depth, frame_num=0, buffer=None, mode=HIT_LIST, nthreads=0, rows=None):
        """Scene::update restated; returns (buffer[h,w,3] bottom-up, ray_count)."""
        w, h = self.p.width, self.p.height
        p = params(w, h, samples, max_depth)
        if buffer is None:
            buffer = np.zeros((h, w, 3), np.float32)
        assert buffer.dtype == np.float32 and buffer.flags.c_contiguous
        if nthreads <= 0:
            nthreads = os.cpu_count() or 1
        r0, r1 = rows if rows is not None else (0, h)
        rays = lib().orc_update(self.h, C.byref(p), frame_num, _p(buffer), mode, nthreads, r0, r1)
        return buffer, int(rays)

    def hit(self, rays6, mode=HIT_LIST, times=None, nthreads=0):
        """Nearest hit of explicit rays, t in (0.001, f32::MAX): (index in the sphere list or -1, t)."""
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        idx = np.zeros(len(rays6), np.int32)
        t = np.zeros(len(rays6), np.float32)
        tm = None if times is None else np.ascontiguousarray(times, np.float32)
        lib().orc_hit_times(self.h, mode, _p(rays6), None if tm is None else _p(tm), len(rays6), _p(idx), _p(t), nthreads or (os.cpu_count() or 1))
        return idx, t

    def record_rays(self, samples, max_depth, cap, pixels=None, frame_num=0, nthreads=0):
        """Every ray `Scene::update` hands to the hit test for pixels [pixels[0], pixels[1]) (default: the whole image), in
        trace order, with the rays' times: (rays6 [n, 6], times [n])."""
        p = params(self.p.width, self.p.height, samples, max_depth)
        p0, p1 = pixels if pixels is not None else (0, self.p.width * self.p.height)
        rays = np.zeros((cap, 6), np.float32)
        times = np.zeros(cap, np.float32)
        n = lib().orc_record_rays(self.h, C.byref(p), frame_num, p0, p1, _p(rays), _p(times), cap, nthreads or (os.cpu_count() or 1))
        return rays[:n], times[:n]

    def turb(self, x, y, z):
        return lib().orc_turb(self.h, x, y, z)

    def noise(self, x, y, z):
        return lib().orc_noise(self.h, x, y, z)

    def tex_value(self, tex, x, y, z):
        out = np.zeros(3, np.float32)
        lib().orc_tex_value(self.h, tex, x, y, z, _p(out))
        return out


def rng_seed(seed):
    s = np.zeros(4, np.uint64)
    lib().orc_rng_seed(C.c_uint64(seed), _p(s))
    return s


def rng_u64(state, n):
    out = np.zeros(n, np.uint64)
    lib().orc_rng_u64(_p(state), _p(out), n)
    return out


def rng_f32(state, n):
    out = np.zeros(n, np.float32)
    lib().orc_rng_f32(_p(state), _p(out), n)
    return out


def sincos(x):
    x = np.ascontiguousarray(x, np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    lib().orc_sincos(_p(x), _p(s), _p(c), x.size)
    return s, c


def srgb(rgb):
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
    out = np.zeros((len(rgb), 3), np.uint8)
    lib().orc_srgb(_p(rgb), _p(out), len(rgb))
    return out


def sphere_uv(n):
    """material.rs:41-49 get_sphere_uv for one normal."""
    out = np.zeros(2, np.float32)
    lib().orc_sphere_uv(float(n[0]), float(n[1]), float(n[2]), _p(out))
    return out


def bench_fixture_ray():
    p = params(200, 100, 10, 10)
    r = np.zeros(6, np.float32)
    lib().orc_bench_fixture_ray(C.byref(p), _p(r))
    return r
