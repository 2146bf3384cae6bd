"""Per-ray parity of the sphere sweep: `pt_debug_hits` runs the render kernels' own two-stage
sweep — packed conservative pre-filter over all spheres, exact re-test of the flagged ones — for caller-supplied rays, and
must return bit for bit the nearest hit (index in the caller's list, t) of

  * the oracle's `SpheresSoA::hit_scalar` restatement (src/collision/spheres_soa.rs:105-155), and
  * the reference's exact expression evaluated on EVERY sphere on the device (mode 1: no pre-filter at all).

This is the test the image-level suites cannot give: a pre-filter that drops a sphere the exact expression accepts is a
silently wrong pixel.  Rays: the reference's own bench fixture (src/bench.rs:17-26, src/collision/spheres_soa.rs:464-485),
more than 10^7 rays recorded from real paths of BASELINE configs 1 and 5, and adversarial sets aimed at the filter's
slack (grazing the r = 1000 ground, origins on surfaces, tangent rays, scenes and cameras translated out to 10^6, a hollow
shell, moving spheres at the ends of the shutter).  Run on a B200: pytest -m gpu.
"""
import ctypes as C

import numpy as np
import pytest

import orc
import pathtrace_rs_b200 as pt
from pathtrace_rs_b200 import ffi

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(3.4028234663852886e38)


def _check(pr, sc, rays, times=None, oracle=True, oracle_mode=orc.HIT_SOA_SCALAR):
    """mode 0 (shipped sweep) == mode 1 (exact on every sphere) == oracle, bit for bit; returns (idx, t, flagged)."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    idx, t, flagged = pr.debug_hits(rays, times, mode=0, want_flagged=True)
    idx1, t1, tested = pr.debug_hits(rays, times, mode=1, want_flagged=True)
    assert np.array_equal(idx, idx1), "pre-filter dropped or mis-ranked a hit: %d rays differ" % int((idx != idx1).sum())
    assert np.array_equal(t.view(np.uint32), t1.view(np.uint32))
    assert (t[idx < 0] == FLT_MAX).all() and (t[idx >= 0] > np.float32(0.001)).all()
    assert (flagged <= tested).all() and (flagged[idx >= 0] >= 1).all()  # a hit sphere was a candidate
    if oracle:
        oi, ot = sc.hit(rays, oracle_mode, times=times)
        assert np.array_equal(idx, oi), "%d of %d rays pick another sphere than the oracle" % (int((idx != oi).sum()), len(rays))
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    return idx, t, flagged


def _unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)


def test_bench_fixture_ray():
    """src/bench.rs:17-26: `random_spheres`, the camera's centre ray; src/collision/spheres_soa.rs:464-485 times exactly this."""
    pr = pt.Preset("random_spheres", pt.Params(200, 100, 10, 10)).create_scene(0)
    sc = orc.Scene("random_spheres", 200, 100)
    ray = orc.bench_fixture_ray()
    idx, t, _ = _check(pr, sc, ray[None, :])
    li, lt = sc.hit(ray[None, :], orc.HIT_LIST)  # the live HitableList path agrees on this ray too
    assert idx[0] == li[0] and idx[0] >= 0 and abs(float(t[0]) - float(lt[0])) <= 1e-5 * float(lt[0])


def test_ten_million_recorded_rays_of_cfg1():
    """Every ray `Scene::update` traces for BASELINE config 1 at 200 spp (primary rays, every bounce, depth 50): > 10^7
    rays, each through the shipped sweep, the device's exact-on-all test and the oracle."""
    w, h = 200, 100
    pr = pt.Preset("random_spheres", pt.Params(w, h, 1, 50)).create_scene(0)
    assert pr.stats().resident in (0, 1, 2, 3)
    sc = orc.Scene("random_spheres", w, h)
    rays, _ = sc.record_rays(200, 50, 12_000_000)
    assert len(rays) > 10_000_000
    idx, t, flagged = _check(pr, sc, rays)
    assert 0.5 < (idx >= 0).mean() < 0.8
    # what the filter costs: candidates per ray handed to the exact test (488 spheres swept)
    assert 1.0 < flagged.mean() < 4.0, flagged.mean()


def test_recorded_rays_through_every_kernel_flavour():
    """The same rays through every sweep the library has: the default resident kernel (pre-filter on the tensor path), the
    same kernel with the packed-FP32 pre-filter, the L2-streamed kernel forced onto the 488-sphere scene (8 tiles, the last
    one ragged), every storage order, and the two-rays-per-lane sweeps with the sphere pairs as uniform operands
    (kernel-parameter image) and as LDS.128 operands."""
    w, h = 96, 48
    sc = orc.Scene("random_spheres", w, h)
    rays, _ = sc.record_rays(16, 50, 400_000)
    base = None
    for opt in (None, pt.PtOptions(resident_kernel=5), pt.PtOptions(resident_kernel=4), pt.PtOptions(force_stream_tile_blocks=16),
                pt.PtOptions(force_stream_tile_blocks=16, resident_kernel=4), pt.PtOptions(force_stream_tile_blocks=4),
                pt.PtOptions(spatial_order=0), pt.PtOptions(spatial_order=1), pt.PtOptions(resident_kernel=5, spatial_order=0),
                pt.PtOptions(resident_kernel=2), pt.PtOptions(resident_kernel=3), pt.PtOptions(resident_kernel=1)):
        pr = pt.Preset("random_spheres", pt.Params(w, h, 1, 50)).create_scene(0, opt)
        idx, t, _ = _check(pr, sc, rays)
        if base is not None:
            assert np.array_equal(idx, base[0]) and np.array_equal(t, base[1])
        base = (idx, t)


def test_recorded_rays_of_cfg5_stress100k():
    """BASELINE config 5 (99 860 spheres, streamed through L2 in TMA tiles): rays of real paths, both pre-filters."""
    w, h = 48, 27
    sc = orc.Scene("stress100k", w, h)
    rays, _ = sc.record_rays(2, 50, 60_000, nthreads=8)
    assert len(rays) > 3000
    for opt in (None, pt.PtOptions(resident_kernel=4)):  # streamed kernel with the tensor-path / the packed-FP32 pre-filter
        pr = pt.Preset("stress100k", pt.Params(w, h, 1, 50)).create_scene(0, opt)
        assert pr.stats().n_spheres in (0, 99860)
        idx, _, flagged = _check(pr, sc, rays)
        assert (idx >= 0).mean() > 0.5 and flagged.mean() < 40


def _custom_scene(cr, options=None):
    """Lambertian spheres from an [n, 4] centre/radius array through the raw C ABI; returns (scene handle wrapper, oracle)."""
    cr = np.ascontiguousarray(cr, np.float32)
    n = len(cr)
    custom = dict(centre_radius=cr, kind=np.zeros(n, np.int32), params5=np.full((n, 5), 0.5, np.float32), motion=None,
                  cam15=np.array([13, 2, 3, 0, 0, 0, 0, 1, 0, 20.0, 2.0, 0.0, 10.0, 0.0, 1.0], np.float32), sky=None)
    sc = orc.Scene("custom", 64, 32, custom=custom)
    L = ffi.libptgpu()
    cols = [np.ascontiguousarray(cr[:, i]) for i in range(4)]
    mats = (ffi.PtMaterial * n)()
    texs = (ffi.PtTexture * n)()
    for i in range(n):
        mats[i].kind, mats[i].texture = 0, i
        texs[i].kind, texs[i].odd, texs[i].even = 0, -1, -1
        texs[i].color[:] = [0.5, 0.5, 0.5]
    midx = np.arange(n, dtype=np.int32)
    d = ffi.PtSceneDesc()
    d.struct_size, d.n_spheres = C.sizeof(ffi.PtSceneDesc), n
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    d.centre_x, d.centre_y, d.centre_z, d.radius = fp(cols[0]), fp(cols[1]), fp(cols[2]), fp(cols[3])
    d.material_index = midx.ctypes.data_as(C.POINTER(C.c_int32))
    d.n_materials = d.n_textures = n
    d.materials, d.textures = mats, texs
    scene = C.c_void_p()
    dev = (C.c_int * 1)(0)
    ffi.check(L.pt_scene_create_multi(C.byref(d), dev, 1, C.byref(options) if options is not None else None, C.byref(scene)))

    class H:  # just enough of Preset for _check
        scene_handle = scene

        def debug_hits(self, rays6, times=None, mode=0, want_flagged=False):
            return pt.Preset.debug_hits(self, rays6, times, mode, want_flagged)

        def close(self):
            L.pt_scene_destroy(scene)
    return H(), sc


def _adversarial_rays(cr, rng, n_per_sphere=64):
    """Rays aimed at what a conservative filter can get wrong: tangent lines of every sphere (offset from the silhouette by
    -8 .. +8 ulp-scale steps of the radius), rays starting ON a surface (bounce rays) along and near the tangent plane,
    rays from far away and from inside."""
    cr = np.asarray(cr, np.float64)
    out = []
    for c, r in zip(cr[:, :3], np.abs(cr[:, 3])):
        u = rng.normal(size=(n_per_sphere, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)               # direction of travel
        v = np.cross(u, rng.normal(size=(n_per_sphere, 3)))
        v /= np.linalg.norm(v, axis=1, keepdims=True)               # perpendicular: where the line passes the centre
        eps = rng.integers(-8, 9, size=(n_per_sphere, 1)) * 2.0 ** -22
        closest = c + v * r * (1.0 + eps)                            # grazing: perpendicular distance = r (1 + eps)
        dist = rng.choice([0.5, 3.0, 40.0, 3000.0], size=(n_per_sphere, 1)) * max(r, 1e-3)
        out.append(np.hstack([closest - u * dist, u]))
        # origin on the surface, direction in / just above / just below the tangent plane
        nrm = rng.normal(size=(n_per_sphere, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        tang = np.cross(nrm, rng.normal(size=(n_per_sphere, 3)))
        tang /= np.linalg.norm(tang, axis=1, keepdims=True)
        tilt = rng.choice([0.0, 1e-7, -1e-7, 1e-3, -1e-3, 0.5, -0.5], size=(n_per_sphere, 1))
        d = tang + tilt * nrm
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        out.append(np.hstack([c + nrm * r, d]))
    rays = np.vstack(out).astype(np.float32)
    rays[:, 3:] = _unit(rays[:, 3:])  # unit in f32, as every direction on the path (camera.rs:65, material.rs:63,82,112,120)
    return rays


def test_adversarial_rays_on_the_rtiow_scene():
    """Tangent / surface-origin / far rays against every sphere of `random_spheres` (ground r = 1000, glass, metal)."""
    pr = pt.Preset("random_spheres", pt.Params(64, 32, 1, 10)).create_scene(0)
    sc = orc.Scene("random_spheres", 64, 32)
    rays = _adversarial_rays(sc.flat()["centre_radius"], np.random.default_rng(5), 96)
    idx, _, _ = _check(pr, sc, rays)
    assert 0.3 < (idx >= 0).mean() < 0.98
    # grazing the ground itself: rays skimming the r = 1000 sphere at heights from -1e-3 to +1e-3 all around the scene
    r = np.random.default_rng(6)
    ang = r.uniform(0, 2 * np.pi, 200_000)
    height = r.choice([0.0, 1e-6, -1e-6, 1e-4, -1e-4, 1e-3, -1e-3, 0.2, 0.4], size=len(ang))
    o = np.stack([30 * np.cos(ang), height + 1e-3 * r.normal(size=len(ang)), 30 * np.sin(ang)], 1)
    tgt = np.stack([r.uniform(-11, 11, len(ang)), height, r.uniform(-11, 11, len(ang))], 1)
    rays = np.hstack([o, _unit(tgt - o)]).astype(np.float32)
    _check(pr, sc, rays)


def test_hollow_shell_and_small_preset():
    """`small` (presets.rs:217-269): a glass sphere with a NEGATIVE-radius shell inside it (presets.rs:265)."""
    pr = pt.Preset("small", pt.Params(64, 32, 1, 10)).create_scene(0)
    sc = orc.Scene("small", 64, 32)
    cr = sc.flat()["centre_radius"]
    assert (cr[:, 3] < 0).any()
    rec, _ = sc.record_rays(64, 10, 300_000)
    _check(pr, sc, rec)
    _check(pr, sc, _adversarial_rays(cr, np.random.default_rng(7), 4096))


@pytest.mark.parametrize("shift", [0.0, 1e2, 1e3, 1e4, 1e5, 1e6])
def test_scene_and_rays_translated_far_from_the_origin(shift):
    """The expanded discriminant cancels |c|^2 against |o|^2: translate a cluster of spheres AND the rays by `shift` along a
    diagonal.  At 10^6 one f32 ulp of a coordinate is 0.06 — the geometry itself is coarse — but filter and exact test see
    the same inputs, so the sweep must still return exactly what the exact expression accepts.  Runs the parameter-image
    kernel, the LDS kernel (3 000 spheres) and the streamed kernel."""
    rng = np.random.default_rng(int(shift) + 1)
    off = np.array([shift, 0.5 * shift, -0.75 * shift])
    for n, options in ((300, None), (3000, None), (300, pt.PtOptions(force_stream_tile_blocks=8))):
        cr = np.zeros((n, 4))
        cr[:, :3] = rng.uniform(-20, 20, (n, 3)) + off
        cr[:, 3] = rng.choice([0.05, 0.3, 1.0, 5.0, -0.7], size=n)
        cr[0] = [off[0], off[1] - 1000.0, off[2], 1000.0]  # a ground sphere under the cluster
        h, sc = _custom_scene(cr.astype(np.float32), options)
        try:
            cr32 = sc.flat()["centre_radius"]
            rays = _adversarial_rays(cr32, rng, 24 if n <= 300 else 4)
            # plus camera-like rays from outside the cluster
            o = rng.uniform(-60, 60, (50_000, 3)) + off
            tgt = rng.uniform(-20, 20, (50_000, 3)) + off
            rays = np.vstack([rays, np.hstack([o, _unit(tgt - o)]).astype(np.float32)])
            idx, _, _ = _check(h, sc, rays)
            assert (idx >= 0).mean() > 0.2
        finally:
            h.close()


def test_origin_far_from_a_scene_at_the_origin():
    """|o| swept to 10^6 against spheres near the origin, and the reverse (huge |c|, small |o|)."""
    rng = np.random.default_rng(9)
    cr = np.zeros((200, 4))
    cr[:, :3] = rng.uniform(-10, 10, (200, 3))
    cr[:, 3] = rng.uniform(0.1, 2.0, 200)
    far = np.zeros((40, 4))
    far[:, :3] = _unit(rng.normal(size=(40, 3))) * (10.0 ** rng.uniform(3, 6, (40, 1)))
    far[:, 3] = np.linalg.norm(far[:, :3], axis=1) * rng.uniform(0.01, 0.5, 40)
    h, sc = _custom_scene(np.vstack([cr, far]).astype(np.float32))
    try:
        rays = []
        for mag in (1e1, 1e2, 1e3, 1e4, 1e5, 1e6):
            o = _unit(rng.normal(size=(20_000, 3))).astype(np.float64) * mag
            tgt = rng.uniform(-12, 12, (20_000, 3))
            rays.append(np.hstack([o, _unit(tgt - o)]))
        rays = np.vstack(rays).astype(np.float32)
        idx, _, _ = _check(h, sc, rays)
        assert (idx >= 0).mean() > 0.3
        _check(h, sc, _adversarial_rays(sc.flat()["centre_radius"], rng, 64))
    finally:
        h.close()


def test_moving_spheres_at_the_ends_of_the_shutter():
    """Preset `random` (presets.rs:150-172, moving_sphere.rs:28-73): the pre-filter sees the static sphere that bounds a
    MovingSphere's whole sweep; rays at time0, time1 and in between, tangent to the sphere where it IS at that time."""
    pr = pt.Preset("random", pt.Params(64, 32, 1, 10)).create_scene(0)
    sc = orc.Scene("random", 64, 32)
    fl = sc.flat()
    cr, mo = fl["centre_radius"].astype(np.float64), fl["motion"].astype(np.float64)
    moving = mo[:, 5] > 0
    assert moving.sum() == 393
    rng = np.random.default_rng(12)
    rays, times = [], []
    for tm in (0.0, 1.0, 0.5, 0.999999, 1e-7):
        s = (tm - mo[:, 3]) / np.where(moving, mo[:, 4] - mo[:, 3], 1.0)
        at = cr.copy()
        at[moving, :3] = cr[moving, :3] + s[moving, None] * (mo[moving, :3] - cr[moving, :3])
        a = _adversarial_rays(at, rng, 24)
        rays.append(a)
        times.append(np.full(len(a), tm, np.float32))
    rays, times = np.vstack(rays), np.concatenate(times)
    # the oracle's hybrid mode = what the kernel computes (static spheres in the SoA form, moving ones in the live form)
    idx, _, _ = _check(pr, sc, rays, times)
    assert moving[idx[idx >= 0]].mean() > 0.3
    rec, rt = sc.record_rays(32, 50, 400_000)
    _check(pr, sc, rec, rt)
    # streamed kernel, same rays
    pr2 = pt.Preset("random", pt.Params(64, 32, 1, 10)).create_scene(0, pt.PtOptions(force_stream_tile_blocks=16))
    _check(pr2, sc, rays, times)


def test_duplicate_and_concentric_spheres_tie_rule():
    """Equal t: the first sphere in the caller's list wins (strict `<`, spheres_soa.rs:126 / hitable_list.rs:49-54), whatever
    order the device stores the spheres in."""
    rng = np.random.default_rng(3)
    base = np.zeros((120, 4))
    base[:, :3] = rng.uniform(-8, 8, (120, 3))
    base[:, 3] = rng.uniform(0.2, 1.5, 120)
    cr = np.vstack([base, base[::-1], base[:40]])  # every sphere two or three times, in shuffled positions
    h, sc = _custom_scene(cr.astype(np.float32))
    try:
        o = _unit(rng.normal(size=(100_000, 3))).astype(np.float64) * 30
        tgt = rng.uniform(-8, 8, (100_000, 3))
        rays = np.hstack([o, _unit(tgt - o)]).astype(np.float32)
        idx, _, _ = _check(h, sc, rays)
        assert (idx[idx >= 0] < 120).all()  # always the first copy
    finally:
        h.close()


def _kernel_of(h, origin=(0.0, 0.0, 0.0)):
    """One 8x8 render through the raw ABI, camera at `origin` -> PtRenderStats.resident: 0 streamed / 1 resident with the
    packed-FP32 pre-filter, 2 resident / 3 streamed with the pre-filter on the tensor path (pt_sweep_mma.cuh)."""
    L = ffi.libptgpu()
    p = pt.Params(8, 8, 1, 2).to_ffi()
    cam = ffi.PtCamera()
    cam.origin[:] = [float(x) for x in origin]
    cam.lower_left_corner[:] = [float(x) - 1.0 for x in origin]
    cam.horizontal[:] = [2.0, 0.0, 0.0]
    cam.vertical[:] = [0.0, 2.0, 0.0]
    cam.u[:] = [1.0, 0.0, 0.0]
    cam.v[:] = [0.0, 1.0, 0.0]
    cam.time1 = 1.0
    buf = np.zeros((8, 8, 3), np.float32)
    rays = C.c_uint64(0)
    ffi.check(L.pt_render(h.scene_handle, C.byref(p), C.byref(cam), 0, buf.ctypes.data_as(C.c_void_p), C.byref(rays)))
    st = ffi.PtRenderStats()
    ffi.check(L.pt_scene_stats(h.scene_handle, C.byref(st)))
    return int(st.resident)


def _renders_on_the_tensor_path(h, origin=(0.0, 0.0, 0.0)):
    return _kernel_of(h, origin) == 2


def test_mid_size_scene_streams_on_the_tensor_path():
    """3 000 spheres: the FP32 pre-filter image still fits in shared memory, the tensor-path image no longer does at two CTAs
    per SM.  The automatic choice streams such a scene through L2 with the tensor-path pre-filter (1.5-1.75x faster than the
    resident FP32 kernel, tools/midsize_bench.py); resident_kernel = 4 keeps it resident.  Same hits either way — also in the
    spatial storage order the scene was given as a resident candidate."""
    rng = np.random.default_rng(21)
    cr = np.hstack([rng.uniform(-30, 30, (3000, 1)), rng.uniform(0.1, 0.4, (3000, 1)), rng.uniform(-30, 30, (3000, 1)), rng.uniform(0.1, 0.4, (3000, 1))])
    cr[0] = [0.0, -1000.0, 0.0, 1000.0]
    o = np.hstack([rng.uniform(-30, 30, (80_000, 1)), rng.uniform(0.5, 6.0, (80_000, 1)), rng.uniform(-30, 30, (80_000, 1))])
    tgt = np.hstack([rng.uniform(-30, 30, (80_000, 1)), rng.uniform(0.0, 0.4, (80_000, 1)), rng.uniform(-30, 30, (80_000, 1))])
    rays = np.hstack([o, _unit(tgt - o)]).astype(np.float32)
    rays[:, 3:] = _unit(rays[:, 3:])
    base = None
    for opt, want in ((None, 3), (pt.PtOptions(resident_kernel=4), 1), (pt.PtOptions(spatial_order=0), 3)):
        h, sc = _custom_scene(cr.astype(np.float32), opt)
        try:
            assert _kernel_of(h) == want
            idx, t, _ = _check(h, sc, rays)
            if base is not None:
                assert np.array_equal(idx, base[0]) and np.array_equal(t, base[1])
            base = (idx, t)
        finally:
            h.close()


def test_tensor_path_queue_overflow_and_rays_outside_its_domain():
    """What the tensor-path pre-filter does when it cannot do its job: (1) 400 nested shells around the origin — every ray
    flags every 16-sphere step, the finder lanes' 12-entry queues fill up and the quad falls back to the exact test on every
    sphere from the first dropped step on; (2) rays whose origin lies beyond the extent the f16 operands are scaled for
    (twice the scene's reach), with non-finite-free but non-unit directions among them — they bypass stage 1 altogether.
    Both must still give the exact sweep's (index, t) bit for bit."""
    rng = np.random.default_rng(11)
    cr = np.zeros((400, 4))
    cr[:, :3] = rng.normal(size=(400, 3)) * 0.05
    cr[:, 3] = np.linspace(5.0, 45.0, 400)
    h, sc = _custom_scene(cr.astype(np.float32), pt.PtOptions(resident_kernel=5))
    try:
        assert _renders_on_the_tensor_path(h)
        o = rng.uniform(-1.0, 1.0, (60_000, 3))
        inside = np.hstack([o, _unit(rng.normal(size=(60_000, 3)))])                       # inside every shell: 400 candidates per ray
        far_o = _unit(rng.normal(size=(60_000, 3))) * rng.choice([150.0, 1.0e3, 1.0e5], size=(60_000, 1))
        tgt = rng.uniform(-40.0, 40.0, (60_000, 3))
        far = np.hstack([far_o, _unit(tgt - far_o)])                                      # beyond the extent (2 x 45 = 90)
        rays = np.vstack([inside, far]).astype(np.float32)
        rays[:, 3:] = _unit(rays[:, 3:])
        idx, _, flagged = _check(h, sc, rays)
        assert (flagged[:60_000] == 400).all() and (idx[:60_000] >= 0).all()
        assert (flagged[60_000:] == 400).all()   # out of domain: the exact test on everything
    finally:
        h.close()


def test_tensor_path_is_chosen_only_where_it_can_pay():
    """resident_kernel = 5 (or automatic) needs at least 128 spheres and spheres that are not tiny against the scene's extent
    (the f16 subnormal range costs an absolute slack); otherwise the scene keeps the packed-FP32 pre-filter.  Either way the
    hits are the exact sweep's."""
    rng = np.random.default_rng(12)
    few = np.hstack([rng.uniform(-6, 6, (100, 3)), rng.uniform(0.2, 1.0, (100, 1))])
    tiny = np.hstack([rng.uniform(-6, 6, (300, 3)), np.full((300, 1), 1.0e-3)])
    tiny[0] = [0.0, -1.0e5, 0.0, 1.0e5]                                                   # a huge ground sets the extent
    normal = np.hstack([rng.uniform(-6, 6, (300, 3)), rng.uniform(0.2, 1.0, (300, 1))])
    # far from the origin the filter works relative to the scene's offset (exact subtraction, MmaScale): still the tensor path
    away = normal.copy()
    away[:, 0] += 1.0e5
    away[:, 2] -= 2.0e5
    for cr, want, shift in ((few, False, np.zeros(3)), (tiny, False, np.zeros(3)), (normal, True, np.zeros(3)), (away, True, np.array([1.0e5, 0.0, -2.0e5]))):
        h, sc = _custom_scene(cr.astype(np.float32), pt.PtOptions(resident_kernel=5))
        try:
            assert _renders_on_the_tensor_path(h, shift if want else (0.0, 0.0, 0.0)) == want
            o = _unit(rng.normal(size=(50_000, 3))) * 14.0 + shift   # inside the operands' extent (twice the reach of the spheres, ~20)
            tgt = rng.uniform(-6, 6, (50_000, 3)) + shift
            rays = np.hstack([o, _unit(tgt - o)]).astype(np.float32)
            rays[:, 3:] = _unit(rays[:, 3:])
            idx, _, flagged = _check(h, sc, rays)
            assert (idx >= 0).mean() > 0.3 and flagged.mean() < 30   # the filter filters (300 spheres)
        finally:
            h.close()


def test_debug_hits_rejects_bad_arguments():
    pr = pt.Preset("small", pt.Params(16, 8, 1, 1)).create_scene(0)
    L = ffi.libptgpu()
    rays = np.zeros((4, 6), np.float32)
    idx, t = np.zeros(4, np.int32), np.zeros(4, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.pt_debug_hits(pr.scene_handle, vp(rays), None, 4, 7, vp(idx), vp(t), None) == ffi.PT_ERR_INVALID
    assert L.pt_debug_hits(pr.scene_handle, None, None, 4, 0, vp(idx), vp(t), None) == ffi.PT_ERR_INVALID
    assert L.pt_debug_hits(pr.scene_handle, vp(rays), None, 0, 0, vp(idx), vp(t), None) == ffi.PT_OK
