"""Procedural stand-in assets.

Every mesh / texture / HDRI the reference's scenes load (`renderer/assets/*.obj`, `*/BaseColor.png`, `sky/*.exr`) is a
git-LFS pointer stub in the checkout, so the config scenes are rebuilt on deterministic procedural stand-ins
with the same structure: closed meshes with per-vertex normals and UVs, RGB8 / gray8 textures, an equirectangular f32 HDRI.
All generators are pure functions of their arguments (fixed seeds), so the oracle and the GPU path see identical bytes.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

f32 = np.float32


@dataclass
class MeshData:
    """In-memory equivalent of what `TriangleMesh::load_obj` holds after tobj (geometry/impls/triangle_mesh.rs:141-180)."""
    positions: np.ndarray  # (V,3) f32
    normals: np.ndarray    # (V,3) f32 (need not be normalised; the loader normalises)
    uvs: np.ndarray | None  # (V,2) f32 or None
    indices: np.ndarray    # (T,3) u32
    single: bool = False   # geometry of a SingleTrianglePrimitive (primitive/impls/single_triangle.rs): three inline vertices
    tangent_tri: np.ndarray | None = None  # (T,) u32: triangle whose load-time tangent each triangle receives (multi-model OBJ files, csrc/host_obj.h); None = its own

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=f32)
        self.normals = np.ascontiguousarray(self.normals, dtype=f32)
        if self.uvs is not None:
            self.uvs = np.ascontiguousarray(self.uvs, dtype=f32)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.uint32)
        if self.tangent_tri is not None:
            self.tangent_tri = np.ascontiguousarray(self.tangent_tri, dtype=np.uint32)

    @property
    def n_tris(self) -> int:
        return int(self.indices.shape[0])


def quad(p0, p1, p2, p3, normal, uv_scale=1.0) -> MeshData:
    """Two triangles (p0,p1,p2),(p0,p2,p3) sharing one normal."""
    pos = np.array([p0, p1, p2, p3], dtype=f32)
    nrm = np.tile(np.array(normal, dtype=f32), (4, 1))
    uvs = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=f32) * f32(uv_scale)
    idx = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    return MeshData(pos, nrm, uvs, idx)


def box(lo, hi, rot_y_deg=0.0, with_uv=True) -> MeshData:
    """Axis-aligned box [lo,hi] with per-face normals (24 vertices, 12 triangles), optionally rotated about Y around its centre."""
    lo = np.array(lo, dtype=np.float64)
    hi = np.array(hi, dtype=np.float64)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]])
    faces = [([4, 5, 6, 7], [0, 0, 1]), ([1, 0, 3, 2], [0, 0, -1]), ([5, 1, 2, 6], [1, 0, 0]),
             ([0, 4, 7, 3], [-1, 0, 0]), ([7, 6, 2, 3], [0, 1, 0]), ([0, 1, 5, 4], [0, -1, 0])]
    pos, nrm, uvs, idx = [], [], [], []
    for k, (vs, n) in enumerate(faces):
        for j, v in enumerate(vs):
            pos.append(c[v])
            nrm.append(n)
            uvs.append([[0, 0], [1, 0], [1, 1], [0, 1]][j])
        idx += [[4 * k, 4 * k + 1, 4 * k + 2], [4 * k, 4 * k + 2, 4 * k + 3]]
    pos = np.array(pos)
    nrm = np.array(nrm, dtype=np.float64)
    if rot_y_deg:
        a = np.deg2rad(rot_y_deg)
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        ctr = (lo + hi) / 2
        pos = (pos - ctr) @ R.T + ctr
        nrm = nrm @ R.T
    return MeshData(pos, nrm, np.array(uvs, dtype=f32) if with_uv else None, np.array(idx))


def _grid_indices(nu, nv, wrap_u):
    """Triangulate an (nv+1) x (nu or nu+1) vertex grid."""
    cols = nu if wrap_u else nu + 1
    idx = []
    for j in range(nv):
        for i in range(nu):
            i1 = (i + 1) % cols if wrap_u else i + 1
            a, b = j * cols + i, j * cols + i1
            c, d = (j + 1) * cols + i1, (j + 1) * cols + i
            idx.append([a, b, c])
            idx.append([a, c, d])
    return np.array(idx, dtype=np.uint32)


def blob(center=(0.0, 1.0, 0.0), radius=1.0, nu=50, nv=50, bump=0.18, seed=0) -> MeshData:
    """'bunny' stand-in: a smooth lumpy closed surface (displaced UV sphere) with analytic-ish normals and UVs.
    nu*nv*2 triangles (50x50 -> 5000, the order of the 530 kB bunny.obj)."""
    rng = np.random.default_rng(seed)
    ks = rng.integers(1, 4, size=(6, 2))
    ph = rng.uniform(0, 2 * np.pi, size=6)
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    U, V = np.meshgrid(u, v)
    theta, phi = V * np.pi, U * 2 * np.pi

    def rad(th, p):
        r = np.ones_like(th)
        for (ka, kb), p0 in zip(ks, ph):
            r = r + bump / 3 * np.sin(ka * p + p0) * np.sin(kb * th) * np.sin(th)
        return r * radius

    def pos(th, p):
        r = rad(th, p)
        return np.stack([r * np.sin(th) * np.cos(p), r * np.cos(th), r * np.sin(th) * np.sin(p)], -1)

    P = pos(theta, phi)
    e = 1e-4
    dth = (pos(theta + e, phi) - pos(theta - e, phi)) / (2 * e)
    dph = (pos(theta, phi + e) - pos(theta, phi - e)) / (2 * e)
    N = np.cross(dph, dth)
    ln = np.linalg.norm(N, axis=-1, keepdims=True)
    radial = P / np.maximum(np.linalg.norm(P, axis=-1, keepdims=True), 1e-12)
    N = np.where(ln > 1e-6, N / np.maximum(ln, 1e-12), radial)
    N = np.where((N * radial).sum(-1, keepdims=True) < 0, -N, N)
    P = P + np.array(center)
    idx = _grid_indices(nu, nv, wrap_u=False)
    # drop the degenerate triangles at the poles (zero area -> the reference's intersect_triangle rejects them anyway, but a
    # zero-area leaf bound would make SAH costs NaN only if a whole node had zero area; keep the mesh clean)
    Pf = P.reshape(-1, 3)
    tri = Pf[idx]
    area = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=-1)
    idx = idx[area > 1e-12]
    return MeshData(Pf, N.reshape(-1, 3), np.stack([U, V], -1).reshape(-1, 2), idx)


def knot(scale=0.45, center=(0.0, 0.0, 0.0), nu=250, nv=40, tube=0.22, p=2, q=3) -> MeshData:
    """'dragon' stand-in: a (p,q) torus-knot tube resting on y = 0, ~unit size.  nu*nv*2 triangles (250x40 -> 20 000)."""
    t = np.linspace(0.0, 2 * np.pi, nu, endpoint=False)
    s = np.linspace(0.0, 2 * np.pi, nv, endpoint=False)

    def curve(tt):
        r = np.cos(q * tt) + 2.0
        return np.stack([r * np.cos(p * tt), -np.sin(q * tt), r * np.sin(p * tt)], -1)

    C = curve(t)
    e = 1e-4
    T = curve(t + e) - curve(t - e)
    T /= np.linalg.norm(T, axis=-1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    B = np.cross(T, up)
    B /= np.linalg.norm(B, axis=-1, keepdims=True)
    Nn = np.cross(B, T)
    cs, sn = np.cos(s), np.sin(s)
    ring = Nn[:, None, :] * cs[None, :, None] + B[:, None, :] * sn[None, :, None]  # (nu,nv,3)
    P = C[:, None, :] + tube * 3.0 * ring
    P = P * scale
    P[..., 1] -= P[..., 1].min()
    P = P + np.array(center)
    U, V = np.meshgrid(np.arange(nu) / nu * 8.0, np.arange(nv) / nv, indexing="ij")
    idx = []
    for i in range(nu):
        i1 = (i + 1) % nu
        for j in range(nv):
            j1 = (j + 1) % nv
            a, b, c, d = i * nv + j, i1 * nv + j, i1 * nv + j1, i * nv + j1
            idx.append([a, b, c])
            idx.append([a, c, d])
    return MeshData(P.reshape(-1, 3), ring.reshape(-1, 3), np.stack([U, V], -1).reshape(-1, 2), np.array(idx))


def triangle_soup(n: int, seed=42) -> MeshData:
    """centres uniform in [-1,1]^3, edge vectors uniform in [-l,l]^3 with l = 0.5 n^(-1/3)."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(-1, 1, size=(n, 3))
    ell = 0.5 * n ** (-1.0 / 3.0)
    e1 = rng.uniform(-ell, ell, size=(n, 3))
    e2 = rng.uniform(-ell, ell, size=(n, 3))
    pos = np.stack([ctr, ctr + e1, ctr + e2], 1).reshape(-1, 3)
    nrm = np.cross(e1, e2)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=-1, keepdims=True), 1e-20)
    nrm = np.repeat(nrm, 3, axis=0)
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    return MeshData(pos, nrm, None, idx)


# ------------------------------------------------------------------ textures
def _value_noise(res, cells, rng):
    g = rng.uniform(0, 1, size=(cells + 1, cells + 1))
    g[-1, :] = g[0, :]
    g[:, -1] = g[:, 0]
    x = np.linspace(0, cells, res, endpoint=False)
    xi = x.astype(int)
    xf = x - xi
    w = xf * xf * (3 - 2 * xf)
    a = g[np.ix_(xi, xi)]
    b = g[np.ix_(xi, xi + 1)]
    c = g[np.ix_(xi + 1, xi)]
    d = g[np.ix_(xi + 1, xi + 1)]
    wx, wy = w[None, :], w[:, None]
    return (a * (1 - wx) + b * wx) * (1 - wy) + (c * (1 - wx) + d * wx) * wy


def fbm(res, seed, octaves=4, base_cells=4):
    rng = np.random.default_rng(seed)
    out = np.zeros((res, res))
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        out += amp * _value_noise(res, base_cells * 2 ** o, rng)
        tot += amp
        amp *= 0.5
    return out / tot


def base_color_texture(res=1024, seed=0) -> np.ndarray:
    """RGB8, sRGB-encoded (stand-in for */BaseColor.png)."""
    n1, n2, n3 = fbm(res, seed), fbm(res, seed + 1), fbm(res, seed + 2)
    yy, xx = np.mgrid[0:res, 0:res] / res
    stripes = 0.5 + 0.5 * np.sin(2 * np.pi * (6 * xx + 2 * n1))
    r = 0.25 + 0.65 * stripes * n2
    g = 0.20 + 0.55 * (1 - stripes) * n3 + 0.15 * n1
    b = 0.15 + 0.50 * n1 * (0.5 + 0.5 * np.cos(2 * np.pi * 3 * yy))
    img = np.clip(np.stack([r, g, b], -1), 0, 1)
    return np.ascontiguousarray((img * 255 + 0.5).astype(np.uint8))


def normal_texture(res=1024, seed=10, strength=2.0) -> np.ndarray:
    """RGB8 tangent-space normal map derived from an fbm height field (stand-in for */Normal.png)."""
    h = fbm(res, seed, octaves=5, base_cells=8)
    dx = (np.roll(h, -1, 1) - np.roll(h, 1, 1)) * res / 2 * strength / 64
    dy = (np.roll(h, -1, 0) - np.roll(h, 1, 0)) * res / 2 * strength / 64
    n = np.stack([-dx, -dy, np.ones_like(h)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    return np.ascontiguousarray(((n * 0.5 + 0.5) * 255 + 0.5).astype(np.uint8))


def gray_texture(res=1024, seed=20, lo=0.0, hi=1.0, threshold=None) -> np.ndarray:
    """gray8 (stand-in for Metallic.png / Roughness.png)."""
    h = fbm(res, seed, octaves=4, base_cells=4)
    h = (h - h.min()) / (h.max() - h.min())
    if threshold is not None:
        h = (h > threshold).astype(np.float64)
    v = lo + (hi - lo) * h
    return np.ascontiguousarray((np.clip(v, 0, 1) * 255 + 0.5).astype(np.uint8))


def sky_hdri(w=1024, h=512, sun_dir=(0.4, 0.7, 0.3), sun_power=60.0) -> np.ndarray:
    """Equirectangular linear-RGB f32 sky: horizon-to-zenith gradient, ground, and a soft sun lobe (stand-in for the 1k EXR)."""
    v = (np.arange(h) + 0.5) / h
    u = (np.arange(w) + 0.5) / w
    theta, phi = np.meshgrid(v * np.pi, u * 2 * np.pi, indexing="ij")
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], -1)
    s = np.array(sun_dir, dtype=np.float64)
    s /= np.linalg.norm(s)
    up = np.clip(d[..., 1], 0, 1)
    sky = np.stack([0.35 + 0.25 * (1 - up), 0.50 + 0.25 * (1 - up), 0.85 + 0.10 * (1 - up)], -1) * (0.6 + 0.6 * up[..., None])
    ground = np.array([0.18, 0.16, 0.14])
    img = np.where(d[..., 1:2] >= 0, sky, ground)
    cosang = np.clip((d * s).sum(-1), -1, 1)
    lobe = np.exp((cosang - 1) / 0.002) * sun_power + np.exp((cosang - 1) / 0.05) * 1.5
    img = img + lobe[..., None] * np.array([1.0, 0.93, 0.82])
    return np.ascontiguousarray(img.astype(f32))
