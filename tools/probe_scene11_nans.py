"""Developer probe: NaN pixels of scene 11 (MIS), GPU vs oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import capi, scenes
from oracle import oracle
std, tab = capi.load_tables()
w, h, spp = 64, 48, 32
sc = tp.Scene(device=0); cam = tp.Camera(45.0, w, h)
scenes.load_scene(11, sc, cam); sc.build(cam)
osc = oracle.scene_from_description(sc.desc, cam.position, std, tab)
for smp in ("sobol", "random"):
    img = tp.RendererImage(w, h, tp.RENDERERS["mis"](tp.RendererArgs((w, h), spp, sc, cam, seed=0)))
    img.render(smp)
    acc, _, st = osc.render(osc.params(w, h, spp, "mis", smp, cam))
    gn, on = ~np.isfinite(img.accumulators).all(2), ~np.isfinite(acc).all(2)
    print(smp, "gpu nan px", gn.sum(), "oracle nan px", on.sum(), "both", (gn & on).sum(), "gpu only", (gn & ~on).sum(), "oracle only", (~gn & on).sum(), "rays", img.stats["closest_rays"], st["closest_rays"])
    # per-path view on oracle-only pixels
    ys, xs = np.nonzero(~gn & on)
    for y, x in list(zip(ys, xs))[:3]:
        xy = np.array([[x, y]] * spp, dtype=np.uint32); si = np.arange(spp, dtype=np.uint32)
        g = img.path_samples(smp, xy, si); o = osc.path_samples(osc.params(w, h, spp, "mis", smp, cam), xy, si)
        bad = np.nonzero(~np.isfinite(o).all(1))[0]
        print("  pixel", x, y, "oracle NaN samples", bad.tolist(), "gpu values there", g[bad].tolist())
