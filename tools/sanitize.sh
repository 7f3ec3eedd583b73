# Developer tool (run under gpurun): compute-sanitizer memcheck / racecheck over one small frame of every kind and a device-built soup
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import scenes, assets
for sid, integ, smp in ((19, "mis", "sobol"), (3, "nee", "random"), (17, "pt", "sobol"), ("lights", "mis", "sobol"), (8, "mis", "sobol")):
    sc, cam = tp.Scene(device=0), tp.Camera(45.0, 48, 36)
    scenes.load_scene(sid, sc, cam); sc.build(cam)
    img = tp.RendererImage(48, 36, tp.RENDERERS[integ](tp.RendererArgs((48, 36), 6, sc, cam))).render(smp)
    print(sid, integ, smp, float(img.accumulators.mean()), img.stats["closest_rays"], flush=True)
    xy = np.array([[0, 0], [47, 35], [10, 20]], np.uint32)
    print(img.path_samples(smp, xy, np.array([0, 5, 3], np.uint32)).sum(), flush=True)
mesh = assets.triangle_soup(5000, 3)
sc = tp.Scene(device=0); sc.build_soup(mesh.positions[mesh.indices.reshape(-1)].reshape(-1, 3, 3))
rays = np.zeros((2000, 7), np.float32); rays[:, 2] = 3; rays[:, 3:6] = np.random.default_rng(0).normal(size=(2000, 3)) * 0.2 + [0, 0, -1]; rays[:, 6] = 3e38
print("soup hits", int((sc.trace(rays)[:, 0] >= 0).sum()), int(sc.trace(rays, any_hit=True)[:, 0].sum()), flush=True)
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
tail -12 gpurun_out/sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
tail -8 gpurun_out/sanitize_racecheck.log
