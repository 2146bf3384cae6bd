"""How often does the sweep's pre-filter fire?  Replays real rays (recorded from the oracle) against the sphere set
and counts candidate events per lane and per warp under different pre-filters.  Development aid for the sweep design."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import orc

W, H, SPP, DEPTH = 1200, 800, 16, 50
sc = orc.Scene("random_spheres", W, H)
L = orc.lib()
L.orc_trace_pixel_rays.restype = C.c_int64
L.orc_trace_pixel_rays.argtypes = [C.c_void_p, C.POINTER(orc.OrcParams), C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64]
p = orc.params(W, H, SPP, DEPTH)
f = sc.flat()
c = f["centre_radius"][:, :3].astype(np.float32); r2 = (f["centre_radius"][:, 3] ** 2).astype(np.float32)
rng = np.random.default_rng(0)
tot = {}
nwarps = 0
for trial in range(40):
    y = int(rng.integers(0, H)); x0 = int(rng.integers(0, W - 32))
    lanes = []
    for l in range(32):
        buf = np.zeros((SPP * (DEPTH + 1), 6), np.float32)
        n = L.orc_trace_pixel_rays(sc.h, C.byref(p), x0 + l, y, buf.ctypes.data_as(C.c_void_p), len(buf))
        lanes.append(buf[:n])
    iters = min(len(a) for a in lanes)
    for k in range(iters):
        rays = np.stack([a[k] for a in lanes])  # 32 x 6: what the warp sweeps in trip k
        o, d = rays[:, None, :3], rays[:, None, 3:]
        co = c[None] - o
        nb = (co * d).sum(-1); cc = (co * co).sum(-1) - r2[None]
        disc = nb * nb - cc
        pos = disc > 0
        sq = np.sqrt(np.where(pos, disc, 0)); t0 = nb - sq; t1 = nb + sq
        t = np.where(t0 > 1e-3, t0, t1); valid = pos & (t > 1e-3)
        filt = {
            "disc>0": pos,
            "disc>0 & (nb>0|inside)": pos & ((nb > 0) | (cc < 0)),
            "valid (t>tmin)": valid,
        }
        # running-min updates in index order (what the exact path accepts)
        tt = np.where(valid, t, np.inf)
        runmin = np.minimum.accumulate(tt, axis=1)
        upd = valid & (tt <= runmin) & (np.concatenate([np.full((32, 1), np.inf), runmin[:, :-1]], axis=1) > tt)
        filt["accepted updates"] = upd
        for name, m in filt.items():
            lane_events = m.sum()
            pair = m[:, : (m.shape[1] // 2) * 2].reshape(32, -1, 2)
            warp_pair_branches = pair.any(axis=(0, 2)).sum()           # outer `if (a || b)` taken by the warp
            warp_half_exec = m.any(axis=0).sum()                       # distinct spheres with >= 1 lane -> inner blocks executed
            e = tot.setdefault(name, [0, 0, 0])
            e[0] += lane_events; e[1] += warp_pair_branches; e[2] += warp_half_exec
        nwarps += 1
print("warp-trips:", nwarps)
for name, e in tot.items():
    print("%-28s lane events/lane-sweep %.2f | warp: pair-branches taken %.1f / 244, exact blocks executed %.1f per sweep" % (name, e[0] / nwarps / 32, e[1] / nwarps, e[2] / nwarps))
