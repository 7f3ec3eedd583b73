import sys, time, numpy as np
sys.path.insert(0, ".")
import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import assets
import torch
mesh = assets.triangle_soup(200000, 3)
sc = tp.Scene(device=0); sc.build_soup(mesh.positions[mesh.indices.reshape(-1)].reshape(-1, 3, 3))
rng = np.random.default_rng(0)
for n in (1000, 100000, 1000000, 4000000):
    rays = np.zeros((n, 7), np.float32); rays[:, 2] = 3; rays[:, 3:6] = rng.normal(size=(n, 3)) * 0.2 + [0, 0, -1]; rays[:, 6] = 3e38
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h = sc.trace(rays); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(n, rep, round(dt * 1e3, 2), "ms", round(n / dt / 1e6, 2), "Mrays/s", flush=True)
