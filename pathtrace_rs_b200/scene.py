"""Python face of the host mirror: Params + Preset (Storage + hitables + Camera + Scene of one preset).

Mirrors the call sequence of ``render_offline`` (src/offline.rs:16-29):
``Preset(name, params)`` = new_rng + Storage::new + presets::from_name; ``.create_scene()`` = params.new_scene
(GPU upload); ``.update(...)`` = ``scene.update(&params, &camera, frame_num, &mut buffer) -> ray_count``.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import ffi


@dataclass
class Params:  # src/params.rs:11-18; defaults main.rs:78-85
    width: int = 1280
    height: int = 720
    samples: int = 4
    max_depth: int = 10
    random_seed: bool = False
    use_bvh: bool = False
    seed_salt: int = 0

    def to_pth(self):
        return ffi.PthParams(self.width, self.height, self.samples, self.max_depth, int(self.random_seed), int(self.use_bvh),
                             self.seed_salt)

    def to_ffi(self):
        p = ffi.PtParams()
        p.width, p.height, p.samples, p.max_depth = self.width, self.height, self.samples, self.max_depth
        p.random_seed, p.use_bvh, p.seed_salt = int(self.random_seed), int(self.use_bvh), self.seed_salt
        return p


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Preset:
    def __init__(self, name, params):
        self.name = name
        self.params = params
        self._L = ffi.libpthost()
        pp = params.to_pth()
        self._h = self._L.pth_preset_build(name.encode(), C.byref(pp))
        if not self._h:
            raise ValueError(self._L.pth_last_error().decode())  # "unrecognised preset" (offline.rs:21)
        self._has_scene = False

    def close(self):
        if getattr(self, "_h", None):
            self._L.pth_preset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self._L.pth_preset_len(self._h))

    @property
    def camera(self):
        cam = ffi.PtCamera()
        self._L.pth_preset_camera(self._h, C.byref(cam))
        return cam

    def flat(self):
        """Host-side scene dump (no GPU needed): per-sphere centre/radius, material kind, colour/fuzz/ref_idx."""
        n = len(self)
        cr = np.zeros((n, 4), np.float32)
        kind = np.zeros(n, np.int32)
        p5 = np.zeros((n, 5), np.float32)
        self._L.pth_preset_spheres(self._h, _vp(cr), _vp(kind), _vp(p5))
        motion = np.zeros((n, 6), np.float32)  # centre1 (3), time0, time1, moving flag (moving_sphere.rs:16-26)
        self._L.pth_preset_motion(self._h, _vp(motion))
        per = ffi.PtPerlin()
        self._L.pth_preset_perlin(self._h, C.byref(per))
        sky = np.zeros(3, np.float32)
        has_sky = self._L.pth_preset_sky(self._h, _vp(sky))
        cam = self.camera
        cam24 = np.frombuffer(bytes(cam), dtype=np.float32).copy()
        return dict(centre_radius=cr, kind=kind, params5=p5, motion=motion,
                    randvec=np.ctypeslib.as_array(per.randvec).reshape(256, 3).copy(),
                    perm=np.stack([np.ctypeslib.as_array(per.perm_x), np.ctypeslib.as_array(per.perm_y),
                                   np.ctypeslib.as_array(per.perm_z)]).copy(),
                    camera=cam24, has_sky=int(has_sky), sky=sky,
                    next_f32=float(self._L.pth_preset_next_f32(self._h)))

    def create_scene(self, device=0, options=None):
        """params.new_scene on one GPU (`device` an int) or replicated on several (`device` a list: the one `update` call
        then splits the image over them by interleaved row tiles inside the library).  `options`: ffi.PtOptions or None."""
        devices = [device] if isinstance(device, int) else list(device)
        arr = (C.c_int32 * len(devices))(*devices)
        opt_ref = C.byref(options) if options is not None else None
        if self._L.pth_scene_create_multi(self._h, arr, len(devices), opt_ref) != 0:
            raise RuntimeError(self._L.pth_last_error().decode())
        self._has_scene = True
        self.devices = devices
        return self

    @property
    def scene_handle(self):
        """The PtScene* the host mirror created (for direct C-ABI calls such as pt_render_device)."""
        h = self._L.pth_scene_handle(self._h)
        if not h:
            raise RuntimeError("scene not created: call create_scene() on a machine with a B200")
        return C.c_void_p(h)

    def update(self, params=None, frame_num=0, buffer=None, part=None):
        """Scene::update through the host mirror and pt_render (host buffer in, host buffer out)."""
        params = params or self.params
        if buffer is None:
            buffer = np.zeros((params.height, params.width, 3), np.float32)
        assert buffer.dtype == np.float32 and buffer.flags.c_contiguous and buffer.size == params.width * params.height * 3
        rays = C.c_uint64(0)
        pp = params.to_pth()
        part_ref = C.byref(part) if part is not None else None
        if self._L.pth_scene_update(self._h, C.byref(pp), frame_num, part_ref, _vp(buffer), C.byref(rays)) != 0:
            raise RuntimeError(self._L.pth_last_error().decode())
        return buffer, int(rays.value)

    def update_progressive(self, params, frame_num, want_rgb=False, want_rgb8=False):
        """pt_render_progressive: the windowed worker loop (glium_window.rs:96-131) with the accumulation buffer resident on
        the device.  Returns (rgb or None, rgb8 or None, ray_count)."""
        L = ffi.libptgpu()
        p = params.to_ffi()
        cam = self.camera
        rgb = np.zeros((params.height, params.width, 3), np.float32) if want_rgb else None
        rgb8 = np.zeros((params.height, params.width, 3), np.uint8) if want_rgb8 else None
        rays = C.c_uint64(0)
        ffi.check(L.pt_render_progressive(self.scene_handle, C.byref(p), C.byref(cam), frame_num, _vp(rgb) if want_rgb else None,
                                          _vp(rgb8) if want_rgb8 else None, C.byref(rays)))
        return rgb, rgb8, int(rays.value)

    def update_device(self, params, frame_num, d_rgb_ptr, d_rays_ptr, stream_ptr=0, part=None):
        """pt_render_device: device-resident buffers (torch tensors' data_ptr()), asynchronous on `stream_ptr`."""
        L = ffi.libptgpu()
        p = params.to_ffi()
        cam = self.camera
        part_ref = C.byref(part) if part is not None else None
        ffi.check(L.pt_render_device(self.scene_handle, C.byref(p), C.byref(cam), frame_num, part_ref, C.c_void_p(d_rgb_ptr),
                                     C.c_void_p(d_rays_ptr), C.c_void_p(stream_ptr)))

    def stats(self):
        st = ffi.PtRenderStats()
        ffi.check(ffi.libptgpu().pt_scene_stats(self.scene_handle, C.byref(st)))
        return st

    def device_stats(self):
        """Per-GPU statistics of the last render (one PtRenderStats per device of the scene)."""
        L = ffi.libptgpu()
        out = []
        for i in range(L.pt_scene_device_count(self.scene_handle)):
            st = ffi.PtRenderStats()
            ffi.check(L.pt_scene_device_stats(self.scene_handle, i, C.byref(st)))
            out.append(st)
        return out

    def debug_hits(self, rays6, times=None, mode=0, want_flagged=False):
        """pt_debug_hits: nearest hit of caller-supplied unit-direction rays through the render kernel's own sweep
        (mode 0) or through the exact test on every sphere (mode 1).  Returns (idx, t[, flagged])."""
        rays6 = np.ascontiguousarray(rays6, np.float32).reshape(-1, 6)
        n = rays6.shape[0]
        idx = np.full(n, -2, np.int32)
        t = np.zeros(n, np.float32)
        flagged = np.zeros(n, np.uint32) if want_flagged else None
        tm = np.ascontiguousarray(times, np.float32) if times is not None else None
        ffi.check(ffi.libptgpu().pt_debug_hits(self.scene_handle, _vp(rays6), _vp(tm) if tm is not None else None, n, mode, _vp(idx), _vp(t),
                                              _vp(flagged) if want_flagged else None))
        return (idx, t, flagged) if want_flagged else (idx, t)

    def srgb8(self, rgb):
        h, w = rgb.shape[:2]
        out = np.zeros((h, w, 3), np.uint8)
        rgb = np.ascontiguousarray(rgb, np.float32)
        ffi.check(ffi.libptgpu().pt_srgb8(self.scene_handle, _vp(rgb), w, h, _vp(out)))
        return out


def device_info(device=0):
    info = ffi.PtDeviceInfo()
    ffi.check(ffi.libptgpu().pt_device_info(device, C.byref(info)))
    return info


def probe_fp32_peak(device=0):
    v = C.c_double(0)
    ffi.check(ffi.libptgpu().pt_probe_fp32_peak(device, C.byref(v)))
    return v.value


def image_open(path):
    """RgbImage::open (src/texture.rs:14-25) through the host mirror: uint8 [height, width, 3], row 0 = top."""
    L = ffi.libpthost()
    w, h = C.c_uint32(), C.c_uint32()
    if L.pth_image_open(os.fsencode(path), C.byref(w), C.byref(h), None, 0) != 0:
        raise RuntimeError(L.pth_last_error().decode())
    px = np.zeros((h.value, w.value, 3), np.uint8)
    if L.pth_image_open(os.fsencode(path), C.byref(w), C.byref(h), _vp(px), px.nbytes) != 0:
        raise RuntimeError(L.pth_last_error().decode())
    return px


def write_ppm(path, rgb8):
    """Binary PPM (P6) writer — the one image format RgbImage::open decodes in this mirror."""
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (rgb8.shape[1], rgb8.shape[0]))
        f.write(rgb8.tobytes())


def render_offline(preset, params, output_png="", device=0, frames=1):
    """offline::render_offline (src/offline.rs:16-60). `device`: an int or a list of GPUs; `frames` > 1 runs the
    progressive loop headless.  Returns (seconds, ray_count)."""
    L = ffi.libpthost()
    secs, rays = C.c_double(0), C.c_uint64(0)
    pp = params.to_pth()
    devices = [device] if isinstance(device, int) else list(device)
    arr = (C.c_int32 * len(devices))(*devices)
    if L.pth_render_offline_multi(preset.encode(), C.byref(pp), output_png.encode(), arr, len(devices), frames, C.byref(secs), C.byref(rays)) != 0:
        raise RuntimeError(L.pth_last_error().decode())
    return secs.value, int(rays.value)
