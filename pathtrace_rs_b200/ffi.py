"""ctypes declarations for include/ptgpu.h (libptgpu.so) and the host mirror's C interface (libpthost.so).

Loading fails loudly when a library is missing: there is no fallback of any kind.
"""
import ctypes as C
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# PTGPU_LIB_DIR: development hook — load kernel variants built by tools/build_variant.sh (lib/<variant>/) instead of lib/
LIB_DIR = os.environ.get("PTGPU_LIB_DIR") or os.path.join(_PKG, "lib")
HEADER = os.path.join(_ROOT, "include", "ptgpu.h")

PT_OK, PT_ERR_INVALID, PT_ERR_UNSUPPORTED, PT_ERR_NO_DEVICE, PT_ERR_CUDA, PT_ERR_TOO_LARGE = range(6)
PT_TEX_CONSTANT, PT_TEX_CHECKER, PT_TEX_NOISE, PT_TEX_IMAGE = range(4)
PT_MAT_LAMBERTIAN, PT_MAT_METAL, PT_MAT_DIELECTRIC, PT_MAT_DIFFUSE_LIGHT = range(4)


class PtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ptgpu error %d: %s" % (code, msg))
        self.code = code


class PtParams(C.Structure):  # src/params.rs:11-18
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32), ("max_depth", C.c_uint32),
                ("random_seed", C.c_uint8), ("use_bvh", C.c_uint8), ("_pad", C.c_uint8 * 6), ("seed_salt", C.c_uint64)]


class PtCamera(C.Structure):  # src/camera.rs:8-19
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("time0", C.c_float), ("time1", C.c_float), ("lens_radius", C.c_float)]


class PtTexture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("color", C.c_float * 3), ("odd", C.c_int32), ("even", C.c_int32),
                ("scale", C.c_float), ("image", C.c_int32)]


class PtImage(C.Structure):  # src/texture.rs:6-10
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("data", C.POINTER(C.c_uint8))]


class PtMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture", C.c_int32), ("albedo", C.c_float * 3), ("fuzz", C.c_float),
                ("ref_idx", C.c_float), ("_pad", C.c_int32)]


class PtPerlin(C.Structure):
    _fields_ = [("randvec", (C.c_float * 3) * 256), ("perm_x", C.c_uint32 * 256), ("perm_y", C.c_uint32 * 256),
                ("perm_z", C.c_uint32 * 256)]


class PtMotion(C.Structure):  # src/collision/moving_sphere.rs:7-26
    _fields_ = [("centre1", C.c_float * 3), ("time0", C.c_float), ("time1", C.c_float), ("moving", C.c_uint32)]


class PtSceneDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_spheres", C.c_uint32),
                ("centre_x", C.POINTER(C.c_float)), ("centre_y", C.POINTER(C.c_float)), ("centre_z", C.POINTER(C.c_float)),
                ("radius", C.POINTER(C.c_float)), ("material_index", C.POINTER(C.c_int32)),
                ("n_materials", C.c_uint32), ("n_textures", C.c_uint32),
                ("materials", C.POINTER(PtMaterial)), ("textures", C.POINTER(PtTexture)), ("perlin", C.POINTER(PtPerlin)),
                ("has_sky", C.c_uint32), ("sky", C.c_float * 3), ("motion", C.POINTER(PtMotion)),
                ("n_images", C.c_uint32), ("_pad", C.c_uint32), ("images", C.POINTER(PtImage))]


class PtPartition(C.Structure):
    _fields_ = [("tile_rows", C.c_uint32), ("part_index", C.c_uint32), ("part_count", C.c_uint32), ("_pad", C.c_uint32)]


class PtOptions(C.Structure):  # explicit launch options (ABI v4; replaces the environment hooks of v3)
    _fields_ = [("struct_size", C.c_uint32), ("force_stream_tile_blocks", C.c_int32), ("stream_ctas", C.c_int32),
                ("chunk_samples", C.c_int32), ("spatial_order", C.c_int32), ("tile_rows", C.c_uint32), ("resident_kernel", C.c_uint32)]

    def __init__(self, force_stream_tile_blocks=0, stream_ctas=0, chunk_samples=0, spatial_order=-1, tile_rows=0, resident_kernel=0):
        super().__init__(C.sizeof(PtOptions), force_stream_tile_blocks, stream_ctas, chunk_samples, spatial_order, tile_rows, resident_kernel)


class PtDeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_clock_khz", C.c_int32), ("fp32_fma_peak_flops", C.c_double), ("global_mem_bytes", C.c_uint64)]


class PtRenderStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("ray_count", C.c_uint64),
                ("n_spheres", C.c_uint32), ("kernel_launches", C.c_uint32), ("grid_ctas", C.c_uint32),
                ("cta_threads", C.c_uint32), ("smem_bytes", C.c_uint32), ("resident", C.c_uint32), ("warp_sweeps", C.c_uint64)]


class PthParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32), ("max_depth", C.c_uint32),
                ("random_seed", C.c_uint32), ("use_bvh", C.c_uint32), ("seed_salt", C.c_uint64)]


def abi_symbols():
    """Every function name include/ptgpu.h declares (used by the symbol-export test)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


_ptgpu = None
_pthost = None


def _load(name):
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise ImportError("%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C pathtrace_rs_b200` "
                          "(there is no fallback implementation)" % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def libptgpu():
    global _ptgpu
    if _ptgpu is None:
        L = _load("libptgpu.so")
        vp = C.c_void_p
        L.pt_abi_version.restype = C.c_int
        L.pt_last_error.restype = C.c_char_p
        L.pt_abi_struct_size.restype = C.c_uint32
        L.pt_abi_struct_size.argtypes = [C.c_int]
        L.pt_partition_rows.restype = C.c_uint32
        L.pt_partition_rows.argtypes = [C.POINTER(PtPartition), C.c_uint32, vp, C.c_uint32]
        L.pt_device_count.restype = C.c_int
        L.pt_device_info.argtypes = [C.c_int, C.POINTER(PtDeviceInfo)]
        L.pt_scene_create.argtypes = [C.POINTER(PtSceneDesc), C.c_int, C.POINTER(vp)]
        L.pt_scene_create_multi.argtypes = [C.POINTER(PtSceneDesc), C.POINTER(C.c_int), C.c_uint32, C.POINTER(PtOptions), C.POINTER(vp)]
        L.pt_scene_device_count.restype = C.c_uint32
        L.pt_scene_device_count.argtypes = [vp]
        L.pt_scene_device_stats.argtypes = [vp, C.c_uint32, C.POINTER(PtRenderStats)]
        L.pt_host_register.argtypes = [vp, C.c_uint64]
        L.pt_host_unregister.argtypes = [vp]
        L.pt_debug_hits.argtypes = [vp, vp, vp, C.c_uint32, C.c_int32, vp, vp, vp]
        L.pt_scene_destroy.argtypes = [vp]
        L.pt_scene_destroy.restype = None
        L.pt_render.argtypes = [vp, C.POINTER(PtParams), C.POINTER(PtCamera), C.c_uint32, vp, C.POINTER(C.c_uint64)]
        L.pt_render_part.argtypes = [vp, C.POINTER(PtParams), C.POINTER(PtCamera), C.c_uint32, C.POINTER(PtPartition), vp,
                                     C.POINTER(C.c_uint64)]
        L.pt_render_progressive.argtypes = [vp, C.POINTER(PtParams), C.POINTER(PtCamera), C.c_uint32, vp, vp, C.POINTER(C.c_uint64)]
        L.pt_render_device.argtypes = [vp, C.POINTER(PtParams), C.POINTER(PtCamera), C.c_uint32, C.POINTER(PtPartition), vp, vp, vp]
        L.pt_srgb8.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp]
        L.pt_srgb8_device.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp]
        L.pt_scene_stats.argtypes = [vp, C.POINTER(PtRenderStats)]
        L.pt_probe_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.pt_scene_storage_order.restype = C.c_uint32
        L.pt_scene_storage_order.argtypes = [C.POINTER(PtSceneDesc), C.POINTER(PtOptions), vp, C.c_uint32]
        L.pt_scene_mma_operand.restype = C.c_uint32
        L.pt_scene_mma_operand.argtypes = [C.POINTER(PtSceneDesc), C.POINTER(PtOptions), vp, C.c_uint32, vp, vp]
        _ptgpu = L
    return _ptgpu


def libpthost():
    global _pthost
    if _pthost is None:
        libptgpu()  # dependency, same directory (rpath $ORIGIN)
        L = _load("libpthost.so")
        vp = C.c_void_p
        L.pth_last_error.restype = C.c_char_p
        L.pth_preset_build.restype = vp
        L.pth_preset_build.argtypes = [C.c_char_p, C.POINTER(PthParams)]
        L.pth_preset_free.argtypes = [vp]
        L.pth_preset_free.restype = None
        L.pth_preset_len.argtypes = [vp]
        L.pth_preset_len.restype = C.c_uint32
        L.pth_preset_camera.argtypes = [vp, C.POINTER(PtCamera)]
        L.pth_preset_next_f32.argtypes = [vp]
        L.pth_preset_next_f32.restype = C.c_float
        L.pth_preset_spheres.argtypes = [vp, vp, vp, vp]
        L.pth_preset_motion.argtypes = [vp, vp]
        L.pth_preset_perlin.argtypes = [vp, C.POINTER(PtPerlin)]
        L.pth_preset_sky.argtypes = [vp, vp]
        L.pth_scene_create.argtypes = [vp, C.c_int32]
        L.pth_scene_create_multi.argtypes = [vp, C.POINTER(C.c_int32), C.c_uint32, C.POINTER(PtOptions)]
        L.pth_render_offline_multi.argtypes = [C.c_char_p, C.POINTER(PthParams), C.c_char_p, C.POINTER(C.c_int32), C.c_uint32, C.c_uint32,
                                               C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.pth_scene_handle.argtypes = [vp]
        L.pth_scene_handle.restype = vp
        L.pth_scene_update.argtypes = [vp, C.POINTER(PthParams), C.c_uint32, C.POINTER(PtPartition), vp, C.POINTER(C.c_uint64)]
        L.pth_render_offline.argtypes = [C.c_char_p, C.POINTER(PthParams), C.c_char_p, C.c_int32, C.POINTER(C.c_double),
                                         C.POINTER(C.c_uint64)]
        L.pth_write_png.argtypes = [C.c_char_p, vp, C.c_uint32, C.c_uint32]
        L.pth_image_open.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp, C.c_uint64]
        _pthost = L
    return _pthost


def check(rc):
    if rc != PT_OK:
        raise PtError(rc, libptgpu().pt_last_error().decode(errors="replace"))
