import sys, os
sys.path.insert(0, os.getcwd())
import pathtrace_rs_b200 as pt
rk = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
preset = sys.argv[3] if len(sys.argv) > 3 else "random_spheres"
w, h = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (1200, 800)
params = pt.Params(w, h, spp, 50)
pr = pt.Preset(preset, params).create_scene(0, pt.PtOptions(resident_kernel=rk))
img, rays = pr.update()
st = pr.stats()
print(f"kernel {rk} {preset} {w}x{h} spp{spp}: {st.kernel_ms:.2f} ms {rays/1e3/st.kernel_ms:.1f} Mrays/s lane_eff {rays/32/max(1,st.warp_sweeps):.3f}")
