"""Developer probe: image-level GPU-vs-oracle error by max_depth.  usage: probe_error_by_depth.py scene integrator sampler spp"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import capi, scenes
from oracle import oracle
std, tab = capi.load_tables()
sid, integ, smp, spp = (int(sys.argv[1]) if sys.argv[1].isdigit() else sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4])
w, h = 64, 48
sc = tp.Scene(device=0); cam = tp.Camera(45.0, w, h)
scenes.load_scene(sid, sc, cam); sc.build(cam)
osc = oracle.scene_from_description(sc.desc, cam.position, std, tab)
for depth in (1, 2, 3, 4, 6, 8, 12, 16):
    img = tp.RendererImage(w, h, tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=0), max_depth=depth))
    img.render(smp)
    acc, _, st = osc.render(osc.params(w, h, spp, integ, smp, cam, max_depth=depth))
    g, o = img.accumulators / spp, acc / spp
    d = np.abs(g - o).sum(2)
    ys, xs = np.nonzero(d > 1e-4 * (np.abs(o).sum(2) + 1e-3))
    print(f"depth {depth}: MRE {np.abs(g - o).mean() / np.abs(o).mean():.3e} signed {(g - o).mean() / np.abs(o).mean():+.3e} rays {img.stats['closest_rays']} {st['closest_rays']} "
          f"shadow {img.stats['shadow_rays']} {st['shadow_rays']} differing px {len(ys)} bbox x {xs.min() if len(xs) else -1}-{xs.max() if len(xs) else -1} y {ys.min() if len(ys) else -1}-{ys.max() if len(ys) else -1}", flush=True)
