"""Developer probe: frame means of pt / nee / mis (unbiasedness check) on the GPU.  usage: gpu_means.py spp scene..."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import scenes
spp = int(sys.argv[1])
w, h = 200, 150
for sid in [int(a) for a in sys.argv[2:]]:
    sc = tp.Scene(device=0); cam = tp.Camera(45.0, w, h)
    scenes.load_scene(sid, sc, cam); sc.build(cam)
    for smp in ("random", "sobol"):
        for integ in ("pt", "nee", "mis"):
            img = tp.RendererImage(w, h, tp.RENDERERS[integ](tp.RendererArgs((w, h), spp, sc, cam, seed=1)))
            img.render(smp)
            a = img.accumulators.astype(np.float64) / spp
            print(sid, smp, integ, a.mean(axis=(0, 1)), "max", a.max(), flush=True)
