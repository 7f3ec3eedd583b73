#!/usr/bin/env python3
"""What would re-sorting the bounce rays buy the traversal kernel?  (VERDICT r01 task 3: "re-measure a ray re-sort for bounce >= 1".)

Upper bound, without building the sort: the camera rays of the bench frame (scene 19, 3840 x 2160, one per pixel) are traced through
tcpt_trace_device, a cosine-weighted bounce ray is spawned at every hit on the floor (the Lambert bucket: 3/4 of the bounce-1 rays of
the benchmarked step), and the SAME bounce rays are traced again in four orders:
  pixel    the order of the camera rays (what the wavefront queue holds, up to the local scrambling of the shading launch)
  shuffled pixel order scrambled inside windows of 4096 rays (the scrambling a bucketed shading launch adds)
  octant   sorted by direction octant, then by the Morton code of the origin on a 1024^2 floor grid
  cell     sorted by the Morton code of the origin alone
A sort pays if  (t_pixel - t_sorted) * (rays of a step / rays here)  exceeds its own cost: key build + 64-bit radix sort + gather of
2 x 16 B per ray, about 20 ps per ray on this GPU.  Run under gpurun:  python tools/probe_ray_sort.py
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def part1by1(v):
    v = v.astype(np.uint64) & 0xffff
    v = (v | (v << 8)) & 0x00ff00ff
    v = (v | (v << 4)) & 0x0f0f0f0f
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def main():
    import torch
    wl = bench.WORKLOADS["scene19_4k"]
    W, H = wl["width"], wl["height"]
    tp, scene, cam = bench.describe_scene(wl, device=0)
    scene.build(cam)
    ctx = scene.ctx
    dev = "cuda:0"
    stream = torch.cuda.Stream(device=0)
    torch.cuda.set_stream(stream)
    FMAX = np.finfo(np.float32).max

    # camera rays in Render space (camera at the origin), pixel centres: Camera::generate_ray (camera.rs:51-81)
    f = cam.direction / np.linalg.norm(cam.direction)
    up = cam.up / np.linalg.norm(cam.up)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    scale, aspect = np.tan(np.deg2rad(cam.fov) / 2), W / H
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    dx = (2 * (x + 0.5) / W - 1) * aspect * scale
    dy = (1 - 2 * (y + 0.5) / H) * scale
    rd = np.stack([dx, dy, -np.ones_like(dx)], -1).reshape(-1, 3)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    d = (rd[:, :1] * s + rd[:, 1:2] * u + rd[:, 2:3] * (-f)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = (d * 1e-5).astype(np.float32)
    n = len(o)

    def pack(o, d):
        m = len(o)
        r = np.zeros((2 * m, 4), dtype=np.float32)
        r[:m, :3], r[:m, 3], r[m:, :3] = o, FMAX, d
        return torch.from_numpy(r).to(dev)

    def trace_dev(rays, m, any_hit, hits):
        ctx.check(ctx.lib.tcpt_trace_device(ctx.handle, C.c_void_p(rays.data_ptr()), m, int(any_hit), C.c_void_p(hits.data_ptr()), C.c_void_p(stream.cuda_stream)))

    def timed(rays, m, any_hit, hits, steps=5, warmup=2):
        for _ in range(warmup):
            trace_dev(rays, m, any_hit, hits)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            trace_dev(rays, m, any_hit, hits)
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    hits = torch.empty(n * 24, dtype=torch.uint8, device=dev)
    rays_c = pack(o, d)
    trace_dev(rays_c, n, False, hits); torch.cuda.synchronize()
    h = hits.cpu().numpy()
    t = h[: n * 16].view(np.float32).reshape(n, 4)[:, 0]
    prim = h[n * 16:].view(np.int32).reshape(n, 2)[:, 0]
    ms_cam = timed(rays_c, n, False, hits)
    # the floor = the primitive most camera rays hit
    vals, cnt = np.unique(prim[prim >= 0], return_counts=True)
    floor = vals[np.argmax(cnt)]
    sel = np.nonzero(prim == floor)[0]
    rng = np.random.default_rng(7)
    sel = np.tile(sel, 4)          # four bounce rays per floor pixel, sample-major like the wavefront's slots (slot = sample * n_pix + pixel)
    m = len(sel)
    u1, u2 = rng.random(m, dtype=np.float32), rng.random(m, dtype=np.float32)
    r, ph = np.sqrt(u1), 2 * np.pi * u2
    # the floor's normal in Render space is the world's +y (the camera transform is a translation + look-at; Render space keeps world axes)
    d2 = np.stack([r * np.cos(ph), np.sqrt(1 - u1), r * np.sin(ph)], -1).astype(np.float32)
    p = o[sel] + d[sel] * t[sel, None]
    o2 = (p + np.array([0, 1e-5, 0], np.float32) + d2 * 1e-5).astype(np.float32)

    def order_keys():
        lo, hi = p[:, [0, 2]].min(0), p[:, [0, 2]].max(0)
        g = np.clip(((p[:, [0, 2]] - lo) / (hi - lo) * 1023).astype(np.int64), 0, 1023)
        cell = (part1by1(g[:, 0]) | (part1by1(g[:, 1]) << np.uint64(1))).astype(np.uint64)
        octant = ((d2[:, 0] < 0).astype(np.uint64) | ((d2[:, 2] < 0).astype(np.uint64) << np.uint64(1)))
        # finer direction classes: 4 azimuth quadrants x 4 elevation bands
        band = np.minimum((d2[:, 1] * 4).astype(np.uint64), 3)
        return cell, (octant << np.uint64(2) | band)

    cell, dirclass = order_keys()
    win = 4096
    shuf = np.arange(m)
    for a in range(0, m, win):
        rng.shuffle(shuf[a:a + win])
    orders = {
        "pixel": np.arange(m),
        "shuffled_4096": shuf,
        "dirclass_then_cell": np.argsort((dirclass << np.uint64(20)) | cell, kind="stable"),
        "cell": np.argsort(cell, kind="stable"),
        "cell_then_dirclass": np.argsort(((cell >> np.uint64(6)) << np.uint64(4)) | dirclass, kind="stable"),
    }
    out = {"camera_rays": n, "camera_ms": ms_cam, "camera_mrays_per_s": n / ms_cam / 1e3, "bounce_rays": m, "floor_primitive": int(floor), "orders": {}}
    hits2 = torch.empty(m * 24, dtype=torch.uint8, device=dev)
    rays = pack(o2, d2); trace_dev(rays, m, False, hits2); torch.cuda.synchronize()
    out["bounce_miss_fraction"] = float((hits2.cpu().numpy()[m * 16:].view(np.int32).reshape(m, 2)[:, 0] < 0).mean())
    ctx.set_option("count_tests", 0)
    for name, idx in orders.items():
        rays = pack(o2[idx], d2[idx])
        ms_closest = timed(rays, m, False, hits2)
        ms_any = timed(rays, m, True, hits2)
        out["orders"][name] = {"closest_ms": ms_closest, "closest_mrays_per_s": m / ms_closest / 1e3, "anyhit_ms": ms_any, "anyhit_mrays_per_s": m / ms_any / 1e3}
        print(name, out["orders"][name], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
