"""Developer probe: per-bounce rays of ONE path, GPU (option debug_path_log, stderr) next to the oracle's recorded rays.
usage: probe_one_path.py scene integrator sampler spp x y sample"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import capi, scenes
from oracle import oracle
std, tab = capi.load_tables()
sid, integ, smp, spp, x, y, s = (int(sys.argv[1]) if sys.argv[1].isdigit() else sys.argv[1]), sys.argv[2], sys.argv[3], *[int(a) for a in sys.argv[4:8]]
w, h = 64, 48
sc = tp.Scene(device=0); cam = tp.Camera(45.0, w, h)
scenes.load_scene(sid, sc, cam); sc.build(cam)
osc = oracle.scene_from_description(sc.desc, cam.position, std, tab)
closest, shadow = osc.record_rays(osc.params(w, h, spp, integ, smp, cam, window=(x, y, x + 1, y + 1)))
starts = np.nonzero((np.abs(closest[:, :3]) < 1e-4).all(1))[0].tolist() + [len(closest)]
rays = closest[starts[s]:starts[s + 1]]
np.set_printoptions(precision=9, linewidth=200)
print("oracle closest rays of the path:")
for i, r in enumerate(rays):
    hit, _, _ = osc.trace(np.concatenate([r, [np.finfo(np.float32).max]])[None].astype(np.float32))
    print(f"  {i}: o {r[:3]} d {r[3:]} -> prim {hit[0, 0]} tri {hit[0, 1]} t {hit[0, 2:3].view(np.float32)[0]:.9g}")
img = tp.RendererImage(w, h, tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=0)))
sc.ctx.set_option("debug_path_log", 1)
g = img.path_samples(smp, np.array([[x, y]], dtype=np.uint32), np.array([s], dtype=np.uint32))
o = osc.path_samples(osc.params(w, h, spp, integ, smp, cam), np.array([[x, y]], dtype=np.uint32), np.array([s], dtype=np.uint32))
print("gpu", g, "oracle", o)
