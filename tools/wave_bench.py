import sys, os
sys.path.insert(0, os.getcwd())
import pathtrace_rs_b200 as pt
for rk, name in ((4, "regroup"), (1, "wave"), (2, "pair-const"), (3, "pair-lds")):
    for preset, w, h, spp in (("random_spheres", 1200, 800, 256), ("random", 1200, 800, 128), ("two_perlin_spheres", 1920, 1080, 64)):
        params = pt.Params(w, h, spp, 50)
        pr = pt.Preset(preset, params).create_scene(0, pt.PtOptions(resident_kernel=rk))
        best = 1e9
        for i in range(3):
            img, rays = pr.update()
            best = min(best, pr.stats().kernel_ms)
        n = len(pr)
        print(f"{name:10s} {preset} n={n} {w}x{h} spp{spp}: kernel {best:.2f} ms {rays/1e6/(best*1e-3):.1f} Mrays/s {rays*16*n/(best*1e-3)/74.45e12*100:.1f}% of FP32 peak  mean {img.mean():.6f} rays {rays} grid {pr.stats().grid_ctas}x{pr.stats().cta_threads} smem {pr.stats().smem_bytes} lane_eff {rays/32/max(1,pr.stats().warp_sweeps):.3f}", flush=True)
