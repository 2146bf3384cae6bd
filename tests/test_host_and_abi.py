"""CPU-side checks of the product: the C ABI library loads and exports what include/ptgpu.h declares, the ctypes
struct layouts match the compiled ones, the host mirror builds the same scenes as the oracle bit for bit, and the
error behaviour without a GPU is loud (no fallback)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import orc
import pathtrace_rs_b200 as pt
from pathtrace_rs_b200 import ffi, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRESETS = ["random", "random_spheres", "small", "two_perlin_spheres", "smallpt", "final", "stress100k"]


def _no_gpu():
    return pt.libptgpu().pt_device_count() == 0


def test_abi_symbols_exported():
    L = pt.libptgpu()
    names = pt.abi_symbols()
    assert len(names) >= 15 and "pt_render" in names and "pt_scene_create" in names
    for n in names:
        assert hasattr(L, n), n
    assert L.pt_abi_version() == 4
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ffi.LIB_DIR, "libptgpu.so")], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_ctypes_layouts_match_compiled_structs():
    L = pt.libptgpu()
    structs = [ffi.PtParams, ffi.PtCamera, ffi.PtTexture, ffi.PtMaterial, ffi.PtPerlin, ffi.PtSceneDesc, ffi.PtPartition,
               ffi.PtDeviceInfo, ffi.PtRenderStats, ffi.PtMotion, ffi.PtImage, ffi.PtOptions]
    for i, s in enumerate(structs):
        assert L.pt_abi_struct_size(i) == C.sizeof(s), s.__name__
    assert L.pt_abi_struct_size(99) == 0
    assert C.sizeof(ffi.PtCamera) == 24 * 4


def test_library_is_sm100a_cuda_with_packed_fp32_and_tma():
    """The shipped kernel is real sm_100a SASS: packed FP32 (FFMA2) in the sweep, bulk TMA copy + mbarrier."""
    so = os.path.join(ffi.LIB_DIR, "libptgpu.so")
    try:
        sass = subprocess.check_output(["cuobjdump", "-sass", so], text=True, stderr=subprocess.STDOUT)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    for mnemonic in ("FFMA2", "UBLKCP", "SYNCS"):
        assert mnemonic in sass, mnemonic
    import re
    kernels = re.split(r"Function : ", sass)
    pick = lambda prefix: [k for k in kernels if k.startswith(prefix)]
    # Resident kernel, parameter-image flavour (every preset of the reference): the sphere pairs are UNIFORM operands —
    # LDCU.64 from the kernel parameters (constant bank 0, uniform index) feeding FFMA2 Rscalar(ray) * URpair(spheres) +
    # Rpair.  ptxas only keeps this shape while the sweep's loop counter is provably uniform; if it ever falls back to LDC
    # into vector registers the loop runs 1.7x slower (tools/probe_sweep2.cu), so the build is checked here, not on the GPU.
    for name in ("_ZN2pt22pt_megakernel_residentILb1ELb0E", "_ZN2pt22pt_megakernel_residentILb1ELb1E",
                 "_ZN2pt18pt_megakernel_waveILb0E", "_ZN2pt18pt_megakernel_waveILb1E"):
        body = pick(name)
        assert body, name
        assert len(re.findall(r"LDCU\.64 UR\d+, c\[0x0\]\[UR\d+", body[0])) >= 24, "uniform sphere loads lost"
        assert len(re.findall(r"FFMA2 R\d+, R\d+(?:\.reuse)?\.F32, UR\d+\.F32x2\.HI_LO, ", body[0])) >= 96, "FFMA2 with uniform sphere pairs lost"
    # LDS flavours (larger resident scenes, streamed scenes): broadcast LDS.128 of the pre-filter image (not generic loads,
    # not local memory) feeding FFMA2 Rpair(spheres) * Rscalar(ray) + Rpair.
    for name in ("_ZN2pt21pt_megakernel_regroupILb0ELb0E", "_ZN2pt21pt_megakernel_regroupILb1ELb0E", "_ZN2pt22pt_megakernel_residentILb0E",
                 "_ZN2pt22pt_megakernel_streamedILb0ELb0E", "_ZN2pt22pt_megakernel_streamedILb1ELb0E"):  # mangled prefixes
        body = pick(name)
        assert body, name
        assert len(re.findall(r"LDS\.128", body[0])) >= 8, name + ": sweep loads are not LDS.128"
        assert len(re.findall(r"FFMA2 R\d+, R\d+(?:\.reuse)?\.F32x2\.HI_LO, R\d+(?:\.reuse)?\.F32, ", body[0])) >= 20, name + ": packed sweep lost"
    # Tensor-path flavour of the regroup kernel (the default for the RTIOW scenes, pt_sweep_mma.cuh): 8 HMMA.16816.F32 with a zero
    # accumulator per loop step (the measured issue rate depends on it: tools/probe_mma_mix.cu), fed by LDS.128 fragments, and
    # the packed -(A*A) - B that turns the two dot products into the sign the candidate test reads.
    for name in ("_ZN2pt21pt_megakernel_regroupILb0ELb1E", "_ZN2pt21pt_megakernel_regroupILb1ELb1E", "_ZN2pt21pt_debug_hits_regroupILb0ELb1E",
                 "_ZN2pt22pt_megakernel_streamedILb0ELb1E", "_ZN2pt22pt_megakernel_streamedILb1ELb1E"):
        body = pick(name)
        assert body, name
        assert len(re.findall(r"HMMA\.16816\.F32 R\d+, R\d+(?:\.reuse)?, R\d+(?:\.reuse)?, RZ", body[0])) == 8, name + ": the 8 MMAs of a loop step"
        assert len(re.findall(r"FFMA2 R\d+, -R\d+\.F32x2\.HI_LO, R\d+\.F32x2\.HI_LO, -R\d+\.F32x2\.HI_LO", body[0])) >= 8, name + ": packed L' lost"
        assert "LDS.128" in body[0]


# ---- the tensor-path pre-filter, evaluated on the CPU from the operand the library builds ---------------------------------
def _mma_operand(cr, options=None, motion=None):
    """pt_scene_mma_operand (host only) for Lambertian spheres `cr` [n, 4] (+ optional MovingSphere records [n, 6]):
    (stored, rows [stored, 16] as f32, scale[7], order[n])."""
    n = len(cr)
    cols = [np.ascontiguousarray(cr[:, i], np.float32) for i in range(4)]
    mats = (ffi.PtMaterial * n)()
    texs = (ffi.PtTexture * n)()
    for i in range(n):
        mats[i].kind, mats[i].texture = 0, i
        texs[i].kind, texs[i].odd, texs[i].even = 0, -1, -1
    midx = np.arange(n, dtype=np.int32)
    d = ffi.PtSceneDesc()
    d.struct_size, d.n_spheres = C.sizeof(ffi.PtSceneDesc), n
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    d.centre_x, d.centre_y, d.centre_z, d.radius = fp(cols[0]), fp(cols[1]), fp(cols[2]), fp(cols[3])
    d.material_index = midx.ctypes.data_as(C.POINTER(C.c_int32))
    d.n_materials = d.n_textures = n
    d.materials, d.textures = mats, texs
    if motion is not None:
        mot = (ffi.PtMotion * n)()
        for i in range(n):
            mot[i].centre1[:] = [float(x) for x in motion[i, :3]]
            mot[i].time0, mot[i].time1, mot[i].moving = float(motion[i, 3]), float(motion[i, 4]), int(motion[i, 5])
        d.motion = mot
    cap = (n + 15) // 16 * 16
    rows = np.zeros((cap, 16), np.uint16)
    scale = np.zeros(7, np.float32)
    order = np.zeros(n, np.uint32)
    stored = pt.libptgpu().pt_scene_mma_operand(C.byref(d), C.byref(options) if options is not None else None, rows.ctypes.data_as(C.c_void_p), cap,
                                                scale.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p))
    return int(stored), rows.view(np.float16).astype(np.float32), scale, order


def _mma_filter_margin(cr, rays, options=None, motion=None, times=None):
    """The tensor-path pre-filter (pt_sweep_mma.cuh) restated in numpy on the library's own sphere operand: the ray operand in
    f32 as mma_ray_operand builds it, f16 hi/lo split, exact products, and a PESSIMISTIC accumulation: every dot product is
    moved against the candidate by 2^-20 of the sum of its |terms| (twice what tools/probe_mma.cu measures for the tensor
    core's f32 accumulate).  Returns (number of (ray, sphere) pairs the reference's exact f32 expression accepts, the smallest
    worst-case L' among them divided by sigma^2 (|c|^2 + r^2 + |o|^2), flagged pairs per ray)."""
    stored, rows, (sigma, s, inv_s, max_o2, tx, ty, tz), order = _mma_operand(cr, options, motion)
    assert stored >= len(cr) and stored % 16 == 0
    t = np.array([tx, ty, tz], np.float32)
    S = rows[: len(cr)].astype(np.float64)                     # stored order
    c = cr[order].astype(np.float32)
    f32 = np.float32
    slack = f32(2.0 ** -15)
    keep = (((rays[:, :3] - t).astype(np.float64)) ** 2).sum(1) <= max_o2  # the kernel's own guard: rays beyond the extent bypass stage 1
    rays = rays[keep]
    if times is not None:
        times = times[keep]
    o_abs, d = rays[:, :3].astype(f32), rays[:, 3:].astype(f32)
    o = o_abs - t                                              # the filter works relative to the scene's offset (exact, see MmaScale)
    assert np.array_equal(o.astype(np.float64), o_abs.astype(np.float64) - t.astype(np.float64))
    nod = -((o[:, 0] * d[:, 0] + o[:, 1] * d[:, 1]) + o[:, 2] * d[:, 2])
    oo = ((o[:, 0] * o[:, 0] + o[:, 1] * o[:, 1]) + o[:, 2] * o[:, 2]) * (f32(1.0) - slack)
    zero = np.zeros_like(nod)
    ra = np.stack([d[:, 0], d[:, 1], d[:, 2], f32(sigma) * nod * f32(inv_s), zero], 1)
    rb = np.stack([f32(2.0) * f32(sigma) * o[:, 0], f32(2.0) * f32(sigma) * o[:, 1], f32(2.0) * f32(sigma) * o[:, 2],
                   -(f32(sigma) * f32(sigma)) * oo * f32(inv_s), np.full_like(nod, s)], 1).astype(f32)

    def operand(r):  # [R_hi(5) | R_lo(5) | R_hi(5) | 0] as exact f64 values of the f16 pieces
        hi = r.astype(np.float16)
        lo = (r - hi.astype(f32)).astype(np.float16)
        return np.concatenate([hi, lo, hi, np.zeros((len(r), 1), np.float16)], 1).astype(np.float64)
    RA, RB = operand(ra), operand(rb)
    A = RA @ S.T                                                # [rays, spheres]: exact sums of exact f16 products
    B = RB @ S.T
    eA = 2.0 ** -20 * (np.abs(RA) @ np.abs(S).T)
    eB = 2.0 ** -20 * (np.abs(RB) @ np.abs(S).T)
    worst = np.maximum(np.abs(A) - eA, 0.0) ** 2 + (B - eB)
    worst -= 2.0 ** -23 * np.maximum(A * A, np.abs(B))          # the packed FMA that forms L'
    # the reference's exact expression, unfused f32 (spheres_soa.rs:116-121)
    co = c[None, :, :3] - o_abs[:, None, :]
    nb = (co[..., 0] * d[:, None, 0] + co[..., 1] * d[:, None, 1]) + co[..., 2] * d[:, None, 2]
    cc = ((co[..., 0] * co[..., 0] + co[..., 1] * co[..., 1]) + co[..., 2] * co[..., 2]) - c[None, :, 3] * c[None, :, 3]
    hit = (nb * nb - cc) > f32(0.0)
    if motion is not None:  # MovingSphere::ray_hit at the ray's time (moving_sphere.rs:28-31,38-48), unfused f32, for the spheres that move
        mo = motion[order].astype(f32)
        moving = mo[:, 5] != 0
        with np.errstate(divide="ignore", invalid="ignore"):  # static spheres have no interval: masked below
            inv_dt = (f32(1.0) / (mo[:, 4] - mo[:, 3])).astype(f32)                                   # moving_sphere.rs:24
            tt = ((times.astype(f32)[:, None] - mo[None, :, 3]) * inv_dt[None, :]).astype(f32)        # :29
            ct = (c[None, :, :3] + tt[..., None] * (mo[None, :, :3] - c[None, :, :3])).astype(f32)    # centre_start + s * centre_delta
            ct = np.where(moving[None, :, None], ct, c[None, :, :3])
        oc = o_abs[:, None, :] - ct
        a_ = (d[:, None, 0] * d[:, None, 0] + d[:, None, 1] * d[:, None, 1]) + d[:, None, 2] * d[:, None, 2]
        b_ = (oc[..., 0] * d[:, None, 0] + oc[..., 1] * d[:, None, 1]) + oc[..., 2] * d[:, None, 2]
        c_ = ((oc[..., 0] * oc[..., 0] + oc[..., 1] * oc[..., 1]) + oc[..., 2] * oc[..., 2]) - c[None, :, 3] * c[None, :, 3]
        hit = np.where(moving[None, :], (b_ * b_ - a_ * c_) > f32(0.0), hit)
    c64 = c.astype(np.float64)
    scale = float(sigma) ** 2 * (((c64[None, :, :3] - t.astype(np.float64)) ** 2).sum(2) + c64[None, :, 3] ** 2 + (o.astype(np.float64) ** 2).sum(1)[:, None])
    margin = (worst / scale)[hit]
    return int(hit.sum()), float(margin.min()), float((worst > 0).sum() / len(rays))


def _tangent_rays(cr, rng, per_sphere):
    """Lines that graze every sphere within a few 2^-22 steps of its radius, from near and far, plus rays leaving a surface
    along its tangent plane (what a conservative filter can get wrong)."""
    out = []
    for cx, cy, cz, r in np.asarray(cr, np.float64):
        r = abs(r)
        u = rng.normal(size=(per_sphere, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
        v = np.cross(u, rng.normal(size=(per_sphere, 3))); v /= np.linalg.norm(v, axis=1, keepdims=True)
        eps = rng.integers(-8, 9, size=(per_sphere, 1)) * 2.0 ** -22
        closest = np.array([cx, cy, cz]) + v * r * (1.0 + eps)
        dist = rng.choice([0.5, 3.0, 40.0], size=(per_sphere, 1)) * max(r, 1e-3)
        out.append(np.hstack([closest - u * dist, u]))
        nrm = rng.normal(size=(per_sphere, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        tang = np.cross(nrm, rng.normal(size=(per_sphere, 3))); tang /= np.linalg.norm(tang, axis=1, keepdims=True)
        dd = tang + rng.choice([0.0, 1e-7, -1e-7, 1e-3, -1e-3], size=(per_sphere, 1)) * nrm
        out.append(np.hstack([np.array([cx, cy, cz]) + nrm * r, dd / np.linalg.norm(dd, axis=1, keepdims=True)]))
    rays = np.vstack(out).astype(np.float32)
    rays[:, 3:] /= np.linalg.norm(rays[:, 3:].astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    return rays


def test_tensor_path_prefilter_is_conservative_in_a_cpu_restatement():
    """Stage 1 of the sweep on the tensor path must flag every sphere the reference's exact expression accepts.  The GPU tests
    check that per ray on hardware (tests/test_gpu_hits.py); this is the same statement without a GPU, from the sphere operand
    the library builds (pt_scene_mma_operand) and with the tensor core's accumulation replaced by a worst case."""
    sc = orc.Scene("random_spheres", 96, 48)
    cr = sc.flat()["centre_radius"]
    rays, _ = sc.record_rays(2, 50, 6000)                      # rays of real paths: camera rays and bounces off surfaces
    rng = np.random.default_rng(17)
    extent2 = _mma_operand(cr)[2][3]
    assert 1.5e7 < extent2 < 1.7e7                             # twice the scene's reach (2 x 2000), squared
    for batch in (rays[:4000], _tangent_rays(cr[rng.choice(len(cr), 60, replace=False)], rng, 24), _tangent_rays(cr[:1], rng, 800)):
        hits, margin, flagged = _mma_filter_margin(cr, batch)
        assert hits > 0.2 * len(batch)
        assert margin > 0.0, "an accepted hit is not a candidate under the worst-case accumulation: margin %g" % margin
        assert flagged < 6.0                                   # and the filter still filters: candidates per ray (488 spheres)
    # the same scene far from the origin (|c|, |o| ~ 60 000 on two axes against radii of 0.2): the filter works relative to the
    # scene's offset, which it may only do where the subtraction is exact
    far = cr.copy()
    far[:, 0] += np.float32(40000.0)
    far[:, 2] -= np.float32(50000.0)
    shifted = rays[:3000].copy()
    shifted[:, 0] += np.float32(40000.0)
    shifted[:, 2] -= np.float32(50000.0)
    scale_far = _mma_operand(far)[2]
    assert abs(scale_far[4] - 40000.0) < 1200 and scale_far[5] == 0.0 and abs(scale_far[6] + 50000.0) < 1200
    hits, margin, _ = _mma_filter_margin(far, shifted)
    assert hits > 1000 and margin > 0.0
    # ... and a scene only moderately off-centre keeps t = 0 (the subtraction would round) and a larger extent instead
    near = cr.copy()
    near[:, 0] += np.float32(3000.0)
    assert _mma_operand(near)[2][4] == 0.0
    # Hitable::MovingSphere enters the filter as the static sphere that bounds its sweep: every sphere the reference's moving test
    # accepts at the ray's time must be a candidate (393 moving spheres of the `random` preset, times over the whole shutter)
    scm = orc.Scene("random", 96, 48)
    fm = scm.flat()
    rays_m, times_m = scm.record_rays(2, 50, 5000)
    assert (fm["motion"][:, 5] != 0).sum() == 393 and times_m.min() >= 0.0 and times_m.max() <= 1.0
    for tm in (times_m[:3000], np.zeros(3000, np.float32), np.ones(3000, np.float32)):
        hits, margin, flagged = _mma_filter_margin(fm["centre_radius"], rays_m[:3000], motion=fm["motion"], times=tm)
        assert hits > 1500 and margin > 0.0 and flagged < 12.0
    # scenes the tensor path is not offered to: fewer than 128 spheres; spheres tiny against the scene's extent
    assert _mma_operand(cr[:100])[0] == 0
    tiny = cr.copy()
    tiny[1:, 3] = 1.0e-4
    assert _mma_operand(tiny)[0] == 0


@pytest.mark.parametrize("preset", PRESETS)
def test_host_mirror_builds_the_oracles_scene(preset):
    w, h = 120, 80
    host = pt.Preset(preset, pt.Params(w, h, 1, 10)).flat()
    o = orc.Scene(preset, w, h).flat()
    assert np.array_equal(host["centre_radius"], o["centre_radius"])
    assert np.array_equal(host["motion"], o["motion"])  # MovingSphere end points / times / flags (moving_sphere.rs:16-26)
    assert (host["motion"][:, 5].sum() > 0) == (preset == "random")
    assert np.array_equal(host["camera"], o["camera"])
    assert np.array_equal(host["randvec"], o["randvec"]) and np.array_equal(host["perm"], o["perm"])
    assert host["has_sky"] == o["has_sky"] and np.array_equal(host["sky"], o["sky"])
    okind = o["mat_kind_tex"][o["sphere_material"], 0] if len(o["sphere_material"]) else np.zeros(0, np.int32)
    assert np.array_equal(host["kind"], okind)
    # colour: constant-texture colour for Lambertian/DiffuseLight, albedo for Metal; fuzz; ref_idx
    for i in range(0, len(okind), max(1, len(okind) // 500)):
        m = o["sphere_material"][i]
        k, t = o["mat_kind_tex"][m]
        exp = o["mat_albedo_fuzz_ref"][m].copy()
        if k in (orc.MAT_LAMBERTIAN, orc.MAT_DIFFUSE_LIGHT) and o["tex_kind_odd_even"][t, 0] == orc.TEX_CONSTANT:
            exp[:3] = o["tex_color_scale"][t, :3]
        if k in (orc.MAT_LAMBERTIAN, orc.MAT_DIFFUSE_LIGHT) and o["tex_kind_odd_even"][t, 0] != orc.TEX_CONSTANT:
            continue
        assert np.array_equal(host["params5"][i], exp), (i, k)


def test_earth_preset_and_rgb_image_open(tmp_path, monkeypatch):
    """presets.rs:555-594 + texture.rs:14-25 through the host mirror: same sphere/camera as the oracle's preset, pixels
    decoded from a binary PPM (the mirror's one decoder), loud errors for a missing file or another format."""
    r = np.random.default_rng(11)
    im = r.integers(0, 256, (6, 10, 3)).astype(np.uint8)
    path = tmp_path / "earthmap.ppm"
    pt.write_ppm(path, im)
    np.testing.assert_array_equal(pt.image_open(path), im)
    with open(tmp_path / "commented.ppm", "wb") as f:  # header comments and arbitrary whitespace are legal PPM
        f.write(b"P6 # made by hand\n10\t6\n# maxval next\n255\n" + im.tobytes())
    np.testing.assert_array_equal(pt.image_open(tmp_path / "commented.ppm"), im)
    monkeypatch.setenv("PATHTRACE_EARTHMAP", str(path))
    host = pt.Preset("earth", pt.Params(120, 80, 1, 10)).flat()
    o = orc.Scene("earth", 120, 80, image=im).flat()
    assert np.array_equal(host["centre_radius"], o["centre_radius"]) and np.array_equal(host["camera"], o["camera"])
    assert host["kind"].tolist() == [orc.MAT_LAMBERTIAN] and host["has_sky"] == o["has_sky"] == 0
    monkeypatch.setenv("PATHTRACE_EARTHMAP", str(tmp_path / "missing.jpg"))
    with pytest.raises(ValueError, match="cannot open"):  # image::open(path).unwrap() in the reference
        pt.Preset("earth", pt.Params(8, 8, 1, 1))
    (tmp_path / "not_ppm.jpg").write_bytes(b"\xff\xd8\xff\xe0 not really a jpeg")
    with pytest.raises(RuntimeError, match="binary PPM"):
        pt.image_open(tmp_path / "not_ppm.jpg")
    (tmp_path / "short.ppm").write_bytes(b"P6\n4 4\n255\n" + bytes(10))
    with pytest.raises(RuntimeError, match="truncated"):
        pt.image_open(tmp_path / "short.ppm")


def _desc_from_flat(cr, motion=None):
    """A PtSceneDesc good enough for the host-only entry points (sphere arrays + optional motion)."""
    n = len(cr)
    cols = [np.ascontiguousarray(cr[:, i], np.float32) for i in range(4)]
    d = ffi.PtSceneDesc()
    d.struct_size, d.n_spheres = C.sizeof(ffi.PtSceneDesc), n
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    d.centre_x, d.centre_y, d.centre_z, d.radius = fp(cols[0]), fp(cols[1]), fp(cols[2]), fp(cols[3])
    keep = [cols]
    if motion is not None:
        mot = (ffi.PtMotion * n)()
        for i in range(n):
            mot[i].centre1[:] = [float(x) for x in motion[i, :3]]
            mot[i].time0, mot[i].time1, mot[i].moving = float(motion[i, 3]), float(motion[i, 4]), int(motion[i, 5])
        d.motion = mot
        keep.append(mot)
    return d, keep


def _storage_order(cr, motion=None, options=None):
    d, keep = _desc_from_flat(cr, motion)
    out = np.zeros(max(len(cr), 1), np.uint32)
    mode = pt.libptgpu().pt_scene_storage_order(C.byref(d), C.byref(options) if options is not None else None,
                                                out.ctypes.data_as(C.c_void_p), len(cr))
    return int(mode), out[: len(cr)]


def test_storage_order_is_a_permutation_with_large_spheres_first():
    """pt_scene_storage_order (host only): what pt_scene_create does to the sphere list of a resident scene — large spheres
    first in list order, the rest along a Morton curve, exact duplicates in list order (equal-t ties stay with the first)."""
    cr = orc.Scene("random_spheres", 200, 100).flat()["centre_radius"]
    n = len(cr)
    mode, order = _storage_order(cr)
    assert mode == 2 and sorted(order.tolist()) == list(range(n))
    assert order[:4].tolist() == [0, n - 3, n - 2, n - 1]  # ground + the three unit spheres (presets.rs:195-213), in list order
    # independent restatement of the Morton part
    c = cr[:, :3].astype(np.float64)
    lo = np.array([np.sort(c[:, a])[int(0.02 * (n - 1))] for a in range(3)])
    hi = np.array([np.sort(c[:, a])[int(0.98 * (n - 1))] for a in range(3)])
    hi = np.where(hi > lo, hi, lo + 1)
    q = (np.clip((c - lo) / (hi - lo), 0, 1) * 1023).astype(np.uint32)

    def spread(v):
        v = v & 0x3ff; v = (v | (v << 16)) & 0x030000ff; v = (v | (v << 8)) & 0x0300f00f
        v = (v | (v << 4)) & 0x030c30c3; v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    small = [i for i in range(n) if i not in (0, n - 3, n - 2, n - 1)]
    assert order[4:].tolist() == sorted(small, key=lambda i: (int(code[i]), i))
    # compactness is the point: a 16-sphere group spans far less ground than a strip of the list
    extent = lambda o: np.mean([np.ptp(c[o[g:g + 16], 0]) + np.ptp(c[o[g:g + 16], 2]) for g in range(16, n - 15, 16)])
    assert extent(order) < 0.6 * extent(np.arange(n))
    # duplicates keep their list order; explicit PtOptions and the small-scene / streamed-scene rules
    dup = np.vstack([cr, cr[7:8], cr[0:1]])
    _, o2 = _storage_order(dup)
    pos = {int(v): j for j, v in enumerate(o2)}
    assert pos[7] < pos[n] and pos[0] < pos[n + 1]
    assert _storage_order(cr[:40])[0] == 0 and _storage_order(cr[:40])[1].tolist() == list(range(40))
    listed = _storage_order(cr, options=pt.PtOptions(spatial_order=0))
    assert listed[0] == 0 and listed[1].tolist() == list(range(n))
    assert _storage_order(cr, options=pt.PtOptions(spatial_order=1))[0] == 1
    assert _storage_order(cr, options=pt.PtOptions(force_stream_tile_blocks=16))[0] == 0  # streamed scenes keep the caller's order
    big = orc.Scene("stress100k", 64, 36).flat()["centre_radius"]
    assert _storage_order(big)[0] == 0
    # moving spheres are placed by the middle of their path
    fr = orc.Scene("random", 200, 100).flat()
    m, o3 = _storage_order(fr["centre_radius"], fr["motion"])
    assert m == 2 and sorted(o3.tolist()) == list(range(len(o3)))


def test_host_rng_continues_like_the_reference():
    p = pt.Preset("random_spheres", pt.Params(200, 100, 10, 10))
    assert abs(p.flat()["next_f32"] - 0.17442238) < 1e-8


def test_unrecognised_preset():
    with pytest.raises(ValueError, match="unrecognised preset"):
        pt.Preset("cornell_smoke", pt.Params(8, 8, 1, 1))


def test_no_gpu_means_loud_failure_not_fallback():
    if not _no_gpu():
        pytest.skip("a GPU is present")
    p = pt.Preset("small", pt.Params(8, 8, 1, 1))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        p.create_scene(0)
    with pytest.raises(RuntimeError, match="scene not created"):
        p.update()
    info = ffi.PtDeviceInfo()
    assert pt.libptgpu().pt_device_info(0, C.byref(info)) == ffi.PT_ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pathtrace_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "lib":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "pt_oracle" not in text and "import orc" not in text, os.path.join(dirpath, f)
    out = subprocess.check_output(["ldd", os.path.join(ffi.LIB_DIR, "libpthost.so")], text=True)
    assert "oracle" not in out


@pytest.mark.parametrize("height,tile,count", [(30, 4, 3), (100, 4, 8), (2160, 4, 8), (7, 4, 4), (5, 8, 2), (1, 1, 3)])
def test_partition_rows_cover_the_image_once(height, tile, count):
    seen = np.zeros(height, np.int32)
    sizes = []
    for idx in range(count):
        rows = parallel.owned_rows(ffi.PtPartition(tile, idx, count, 0), height)
        assert np.all(np.diff(rows.astype(np.int64)) > 0)
        assert np.all((rows // tile) % count == idx)
        seen[rows] += 1
        sizes.append(len(rows))
    assert np.all(seen == 1)
    assert max(sizes) - min(sizes) <= tile
    whole = parallel.owned_rows(ffi.PtPartition(0, 0, 0, 0), height)  # defaults: tile 4, one part
    assert np.array_equal(whole, np.arange(height))


def test_png_writer_roundtrip(tmp_path):
    import zlib
    rgb = (np.arange(5 * 7 * 3) % 251).astype(np.uint8).reshape(5, 7, 3)
    path = str(tmp_path / "t.png")
    assert pt.libpthost().pth_write_png(path.encode(), rgb.ctypes.data_as(C.c_void_p), 7, 5) == 0
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        n = int.from_bytes(data[pos:pos + 4], "big")
        typ = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == int.from_bytes(data[pos + 8 + n:pos + 12 + n], "big")
        if typ == b"IHDR":
            ihdr = body
        if typ == b"IDAT":
            idat += body
        pos += 12 + n
    assert int.from_bytes(ihdr[:4], "big") == 7 and int.from_bytes(ihdr[4:8], "big") == 5 and ihdr[8:10] == b"\x08\x02"
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(5, 1 + 7 * 3)
    assert np.all(raw[:, 0] == 0) and np.array_equal(raw[:, 1:].reshape(5, 7, 3), rgb)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract: ONE JSON line on stdout.  The reference arm runs without a GPU (the CPU restatement on the host
    cores); whatever a library prints goes to stderr (bench.claim_stdout), the line carries the keys the driver reads."""
    import json
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    code = ("import os, sys; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'cfg1', '--steps', '1', '--warmup', '0'];"
            "sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.write(1, b'a library banner\\n'); sys.exit(bench.main())" % root)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1 and "a library banner" in p.stderr
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cfg1" in line["config"]["workload"]
