"""Developer probe (not a test): locate the pixels / paths where a GPU frame differs from the oracle's.  Run under gpurun.
usage: probe_frame_diff.py scene integrator sampler [spp]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import capi, scenes
from oracle import oracle

std, tab = capi.load_tables()
sid, integ, smp = (int(sys.argv[1]) if sys.argv[1].isdigit() else sys.argv[1]), sys.argv[2], sys.argv[3]
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 32
w, h = 64, 48
sc = tp.Scene(device=0); cam = tp.Camera(45.0, w, h)
scenes.load_scene(sid, sc, cam); sc.build(cam)
osc = oracle.scene_from_description(sc.desc, cam.position, std, tab)
img = tp.RendererImage(w, h, tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=0)))
img.render(smp)
acc, _, st = osc.render(osc.params(w, h, spp, integ, smp, cam))
g, o = img.accumulators / spp, acc / spp
d = np.abs(g - o).sum(2)
print("MRE", np.abs(g - o).mean() / np.abs(o).mean(), "rays", img.stats["closest_rays"], st["closest_rays"], "pixels differing >1e-5:", (d > 1e-5).sum())
order = np.argsort(d.ravel())[::-1][:8]
for idx in order:
    y, x = divmod(int(idx), w)
    xy = np.array([[x, y]] * spp, dtype=np.uint32); si = np.arange(spp, dtype=np.uint32)
    gs = img.path_samples(smp, xy, si); os_ = osc.path_samples(osc.params(w, h, spp, integ, smp, cam), xy, si)
    bad = np.nonzero(np.abs(gs - os_).max(1) > 1e-6 * (np.abs(os_).max(1) + 1e-6))[0]
    print(f"pixel ({x},{y}) diff {d[y, x]:.4e}: gpu {g[y, x]} oracle {o[y, x]}; differing samples {bad.tolist()}")
    for s in bad[:4]:
        print("   s", s, "gpu", gs[s], "oracle", os_[s])
# depth scan of the worst paths: first max_depth at which the two sides part
print("--- depth scan")
for idx in order[:6]:
    y, x = divmod(int(idx), w)
    xy = np.array([[x, y]] * spp, dtype=np.uint32); si = np.arange(spp, dtype=np.uint32)
    gs = img.path_samples(smp, xy, si); os_ = osc.path_samples(osc.params(w, h, spp, integ, smp, cam), xy, si)
    bad = np.nonzero(np.abs(gs - os_).max(1) > 1e-6 * (np.abs(os_).max(1) + 1e-6))[0]
    if not len(bad):
        continue
    s = int(bad[0])
    for depth in range(0, 17):
        im2 = tp.RendererImage(w, h, tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=0), max_depth=depth))
        g1 = im2.path_samples(smp, np.array([[x, y]], dtype=np.uint32), np.array([s], dtype=np.uint32))[0]
        o1 = osc.path_samples(osc.params(w, h, spp, integ, smp, cam, max_depth=depth), np.array([[x, y]], dtype=np.uint32), np.array([s], dtype=np.uint32))[0]
        print(f"  pixel ({x},{y}) s {s} depth {depth}: gpu {g1} oracle {o1} {'DIFF' if np.abs(g1 - o1).max() > 1e-6 * (np.abs(o1).max() + 1e-6) else ''}")
