import sys, os
sys.path.insert(0, os.getcwd())
import pathtrace_rs_b200 as pt
params = pt.Params(160, 80, 8, 50)
pr = pt.Preset("random_spheres", params).create_scene(0, pt.PtOptions(resident_kernel=1))
img, rays = pr.update()
print("ok", rays, img.mean(), pr.stats().kernel_ms)
