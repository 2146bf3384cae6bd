"""Exploratory GPU-vs-oracle comparison (development aid; the real checks live in tests/)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import pathtrace_rs_b200 as pt
import orc

def run(preset, w, h, spp, depth, compare=True):
    p = pt.Params(w, h, spp, depth)
    pr = pt.Preset(preset, p).create_scene(0)
    t = time.time(); img, rays = pr.update(); dt = time.time() - t
    st = pr.stats()
    n = len(pr)
    print(f"{preset} {w}x{h} spp{spp} d{depth}: gpu {dt*1e3:.1f} ms wall, kernel {st.kernel_ms:.2f} ms, rays {rays} ({rays/(w*h*spp):.3f}/sample) "
          f"{rays/1e6/(st.kernel_ms*1e-3):.1f} Mrays/s  {rays*16*n/(st.kernel_ms*1e-3)/1e12:.2f} TFLOP/s grid {st.grid_ctas} smem {st.smem_bytes} resident {st.resident}", flush=True)
    if compare:
        sc = orc.Scene(preset, w, h)
        t = time.time(); ref, rrays = sc.update(spp, depth, mode=orc.HIT_SOA_SCALAR | 0x100); dt = time.time() - t
        d = np.abs(img - ref)
        print(f"   oracle(soa,iter) {dt:.2f}s rays {rrays}; rays equal {rays == rrays}; max abs diff {d.max():.3e}; exact pixels {np.mean(np.all(img == ref, axis=2))*100:.2f}% ; "
              f"within 1e-5: {np.mean(np.all(d < 1e-5, axis=2))*100:.2f}% ; mean gpu {img.mean(axis=(0,1))} ref {ref.mean(axis=(0,1))}", flush=True)
    return img

if __name__ == "__main__":
    info = pt.device_info(0)
    print(info.name.decode(), info.sm_count, info.sm_clock_khz, info.fp32_fma_peak_flops / 1e12, "TF nominal; probe", pt.probe_fp32_peak(0) / 1e12)
    run("final", 64, 32, 4, 10)
    run("small", 100, 50, 16, 10)
    run("random_spheres", 200, 100, 10, 10)
    run("random_spheres", 200, 100, 100, 50)
    run("two_perlin_spheres", 200, 100, 16, 50)
    run("smallpt", 100, 100, 16, 10)
    run("random_spheres", 1200, 800, 16, 50, compare=False)
    run("random_spheres", 1200, 800, 64, 50, compare=False)
    run("stress100k", 64, 36, 2, 50, compare=True)
