"""BASELINE.json configs[4]: the device-built soup BVH (tcpt_scene_build_soup, csrc/lbvh.cuh: Morton order, radix tree, 4-wide collapse) under
the same traversal kernels.  A different tree must give the same hits: closest hits are decided by (t, ...) and a soup has no ties, any-hit
by existence.  Checked against the reference-topology scene of the same triangles (host builder, bit-exact with the oracle elsewhere) and
against a float64 brute force over every triangle."""
import numpy as np
import pytest

import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import assets, capi, scenes

pytestmark = pytest.mark.gpu
FMAX = np.finfo(np.float32).max
f32 = np.float32


def soup_triangles(n, seed=42):
    mesh = assets.triangle_soup(n, seed)
    return mesh, mesh.positions[mesh.indices.reshape(-1)].reshape(-1, 3, 3)


def rays_towards_the_soup(n_rays, seed=0):
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n_rays, 3)).astype(f32)
    o = (o / np.linalg.norm(o, axis=1, keepdims=True) * f32(3.0)).astype(f32)
    target = rng.uniform(-1, 1, size=(n_rays, 3)).astype(f32)
    d = target - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    return np.concatenate([o, d, np.full((n_rays, 1), FMAX, f32)], 1)


@pytest.mark.parametrize("n_tri,leaf", [(1, 2), (3, 2), (4, 4), (5, 4), (5, 1), (37, 2), (37, 16), (20000, 1), (20000, 4), (300000, 2)])
def test_device_built_soup_gives_the_hits_of_the_reference_topology(n_tri, leaf):
    mesh, tris = soup_triangles(n_tri)
    dev = tp.Scene(device=0)
    dev.ctx.set_option("soup_leaf", leaf)
    dev.build_soup(tris)
    info = dev.soup_build_info()
    assert info["records"] >= 2 and info["levels"] >= 1 and info["build_ms"] > 0
    host = tp.Scene(device=0)
    cam = tp.Camera(45.0, 8, 8)          # camera at the origin: Render space = world space
    host.create_primitive(tp.CreatePrimitiveDesc.GeometryPrimitive(host.load_obj(mesh), scenes._lambert(0.5, 0.5, 0.5), tp.Transform.identity()))
    host.build(cam)
    rays = rays_towards_the_soup(40000 if n_tri > 100 else 4000)
    if n_tri <= 100:      # aim at the few triangles there are
        c = tris.mean(axis=1)[np.arange(len(rays)) % n_tri] + np.random.default_rng(1).normal(scale=0.02, size=(len(rays), 3)).astype(f32)
        d = c - rays[:, :3]
        rays[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    a, b = dev.trace(rays), host.trace(rays)
    assert (a[:, 0] >= 0).mean() > 0.05
    assert np.array_equal(a, b)                                  # primitive, triangle, t bits, barycentric bits
    sa, sb = dev.trace(rays, any_hit=True), host.trace(rays, any_hit=True)
    assert np.array_equal(sa[:, 0], sb[:, 0])
    assert np.array_equal(sa[:, 0] != 0, a[:, 0] >= 0)


def test_device_built_soup_against_float64_brute_force():
    n_tri = 5000
    _, tris = soup_triangles(n_tri, seed=7)
    dev = tp.Scene(device=0)
    dev.build_soup(tris)
    rays = rays_towards_the_soup(3000, seed=3)
    hits = dev.trace(rays)
    # Moeller-Trumbore in float64 over every triangle
    o, d = rays[:, None, :3].astype(np.float64), rays[:, None, 3:6].astype(np.float64)
    p0, e1, e2 = tris[None, :, 0].astype(np.float64), (tris[:, 1] - tris[:, 0])[None].astype(np.float64), (tris[:, 2] - tris[:, 0])[None].astype(np.float64)
    best_t, best_i = np.full(len(rays), np.inf), np.full(len(rays), -1)
    for lo in range(0, len(rays), 500):
        sl = slice(lo, lo + 500)
        pv = np.cross(d[sl], e2)
        det = (e1 * pv).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tv = o[sl] - p0
            u = (tv * pv).sum(-1) * inv
            qv = np.cross(tv, e1)
            v = (d[sl] * qv).sum(-1) * inv
            t = (e2 * qv).sum(-1) * inv
        ok = (np.abs(det) > 1e-14) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 1e-7)
        t = np.where(ok, t, np.inf)
        best_i[sl] = np.where(np.isfinite(t.min(1)), t.argmin(1), -1)
        best_t[sl] = t.min(1)
    hit = best_i >= 0
    t_gpu = hits[:, 2].copy().view(f32)
    agree = (hits[:, 0] >= 0) == hit
    assert agree.mean() > 0.999                                   # edge-grazing rays may differ between f32 watertight and f64 Moeller-Trumbore
    both = hit & (hits[:, 0] >= 0)
    assert (hits[both, 1] == best_i[both]).mean() > 0.999
    assert np.abs(t_gpu[both] - best_t[both]).max() < 1e-4


def test_a_traversal_only_scene_refuses_to_render():
    _, tris = soup_triangles(100)
    sc = tp.Scene(device=0)
    sc.build_soup(tris)
    cam = tp.Camera(45.0, 8, 8)
    with pytest.raises(capi.TcptError, match="traversal data only"):
        tp.RendererImage(8, 8, tp.SrgbRendererMis(tp.RendererArgs((8, 8), 1, sc, cam))).render("sobol")
