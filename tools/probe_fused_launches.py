"""Developer probe: fused vs unfused launch structure."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import scenes
sid, integ, spp, depth = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
w, h = 64, 48
sc = tp.Scene(device=0); cam = tp.Camera(45.0, w, h)
scenes.load_scene(sid, sc, cam); sc.build(cam)
def render(fused):
    sc.ctx.set_option("fused_launches", fused)
    img = tp.RendererImage(w, h, tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=0), max_depth=depth))
    img.render("sobol")
    return img.accumulators.copy(), dict(img.stats), img
a, sa, _ = render(int(sys.argv[5]) if len(sys.argv) > 5 else 3)
b, sb, img = render(0)
print("fused nonfinite", (~np.isfinite(a)).sum(), "plain nonfinite", (~np.isfinite(b)).sum(), sa["closest_rays"], sb["closest_rays"], sa["shadow_rays"], sb["shadow_rays"])
bad = np.argwhere(~np.isfinite(a).all(2) | (np.abs(a - b).sum(2) > 0))
print("differing pixels", len(bad), bad[:10].tolist())
for y, x in bad[:3]:
    xy = np.array([[x, y]] * spp, dtype=np.uint32); si = np.arange(spp, dtype=np.uint32)
    sc.ctx.set_option("fused_launches", 1); g1 = img.path_samples("sobol", xy, si)
    sc.ctx.set_option("fused_launches", 0); g0 = img.path_samples("sobol", xy, si)
    print((x, y), "fused", g1[:4].tolist(), "plain", g0[:4].tolist(), "film fused", a[y, x], "plain", b[y, x])
