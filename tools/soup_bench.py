#!/usr/bin/env python3
"""BASELINE.json configs[4]: synthetic triangle-soup traversal micro-benchmark (run under gpurun).

N triangles (centres uniform in [-1,1]^3, edges uniform in [-l,l]^3, l = 0.5 N^(-1/3)); COHERENT rays = 1920x1080 pinhole
primaries from (0,0,3) toward the origin, fov 45; INCOHERENT rays = cosine-hemisphere directions about the geometric normal
at each primary hit.  The soup BVH comes from the host binned-SAH builder (the reference's builder is O(N^2): soups are outside
topology parity).  Times tcpt_trace_device (device-resident rays and hit records) with CUDA events; prints one JSON line per N."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import toy_cpu_pathtracing_b200 as tp  # noqa: E402
from toy_cpu_pathtracing_b200 import assets, scenes  # noqa: E402

W, H = 1920, 1080
STREAM = None


def primaries():
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    s = np.float32(np.tan(np.deg2rad(45.0) / 2))
    dx = (2 * (x + 0.5) / W - 1) * (W / H) * s
    dy = (1 - 2 * (y + 0.5) / H) * s
    d = np.stack([dx, dy, -np.ones_like(dx)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.zeros_like(d)  # Render space: the camera sits at the origin
    return o.astype(np.float32), d.astype(np.float32)


def pack(o, d, tmax):
    n = len(o)
    r = np.zeros((2 * n, 4), dtype=np.float32)
    r[:n, :3], r[:n, 3], r[n:, :3] = o, tmax, d
    return torch.from_numpy(r).cuda()


def trace(ctx, rays, n, any_hit=False, reps=5):
    hits = torch.empty(n * 24, dtype=torch.uint8, device="cuda")
    s = STREAM   # a real (non-default) stream: libtcpt launches on the handle it is given, so the events bracket the kernel
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        ctx.check(ctx.lib.tcpt_trace_device(ctx.handle, C.c_void_p(rays.data_ptr()), n, int(any_hit), C.c_void_p(hits.data_ptr()), C.c_void_p(s.cuda_stream)))
        e1.record(s)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, hits


def main():
    global STREAM
    STREAM = torch.cuda.Stream()
    for n_tri in [int(a) for a in sys.argv[1:]] or [1_000_000]:
        scene = tp.Scene(device=0)
        scene.ctx.set_option("binned_builder", 1)
        cam = tp.Camera(45.0, W, H)
        t0 = time.time(); scenes.load_soup(scene, cam, n_tri); scene.build(cam); build_s = time.time() - t0
        ctx = scene.ctx
        o, d = primaries()
        n = len(o)
        rays = pack(o, d, np.finfo(np.float32).max)
        ms_c, hits = trace(ctx, rays, n)
        h = hits.cpu().numpy()
        t = h[: n * 16].view(np.float32).reshape(n, 4)[:, 0]
        prim = h[n * 16:].view(np.int32).reshape(n, 2)[:, 0]
        hit = prim >= 0
        # incoherent: cosine-weighted direction about +-z-ish random frame at each hit point (geometry normal not needed for a throughput test)
        rng = np.random.default_rng(1)
        u1, u2 = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
        r, ph = np.sqrt(u1), 2 * np.pi * u2
        loc = np.stack([r * np.cos(ph), r * np.sin(ph), np.sqrt(1 - u1)], -1)
        axis = rng.normal(size=(n, 3)).astype(np.float32); axis /= np.linalg.norm(axis, axis=1, keepdims=True)
        tng = np.cross(axis, np.roll(axis, 1, axis=1)); tng /= np.linalg.norm(tng, axis=1, keepdims=True)
        btn = np.cross(axis, tng)
        d2 = (tng * loc[:, :1] + btn * loc[:, 1:2] + axis * loc[:, 2:3]).astype(np.float32)
        o2 = (o + d * np.where(hit, t, 1.0)[:, None] + d2 * 1e-4).astype(np.float32)
        sel = np.nonzero(hit)[0]
        m = len(sel)
        rays2 = pack(o2[sel], d2[sel], np.finfo(np.float32).max)
        ms_i, _ = trace(ctx, rays2, m)
        ms_s, _ = trace(ctx, rays2, m, any_hit=True)
        ctx.set_option("count_tests", 1)
        trace(ctx, rays2, m, reps=1)
        st = ctx.stats()
        ctx.set_option("count_tests", 0)
        print(json.dumps({"triangles": n_tri, "build_s": round(build_s, 2), "bvh_depth": st["max_bvh_depth"], "primary_rays": n, "primary_hit_rate": float(hit.mean()),
                          "coherent_Mrays_s": n / ms_c / 1e3, "incoherent_closest_Mrays_s": m / ms_i / 1e3, "incoherent_anyhit_Mrays_s": m / ms_s / 1e3,
                          "incoherent_box_tests_per_ray": st["box_tests"] / m, "incoherent_tri_tests_per_ray": st["tri_tests"] / m}))
        del scene


if __name__ == "__main__":
    main()
