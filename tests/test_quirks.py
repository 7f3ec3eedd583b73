"""Tests for the quirks of SURVEY.md Appendix A that no other test names (task 9): each one would fail if the quirk were
dropped from the oracle, and the GPU is tied to the oracle on the very same scene.  DESIGN.md section 9 (quirk ledger) points here.

  q3   DirectionalLight shadow rays start ON the surface, without the 1e-4 push every other shadow ray gets (common.rs:66-67)
  q12  Russian roulette draws no random number while max(throughput) >= 1 (base_renderer.rs:76-92)
  q26  an RGB albedo with a component above 1 is refused (the reference panics: rgb_sigmoid_polynomial.rs:95-108)
  q27  the CDF search is Rust's slice::binary_search_by (exact match -> that index, else the insertion point, clamped;
       environment_light.rs:218-223)"""
import ctypes as C

import numpy as np
import pytest

import toy_cpu_pathtracing_b200 as tp
from toy_cpu_pathtracing_b200 import assets, capi
from toy_cpu_pathtracing_b200.scene import ConstantSpectrum, CreatePrimitiveDesc, LambertMaterial, NormalParameter, SpectrumParameter, Transform
from toy_cpu_pathtracing_b200.scenes import _lambert

GP = CreatePrimitiveDesc.GeometryPrimitive
FMAX = np.finfo(np.float32).max
f32 = np.float32


# ------------------------------------------------------------------ q3
def wall_under_a_directional_light(scene, camera):
    """A wall at z = -2 facing the camera (which sits at the origin looking down -z), lit from behind the camera by a DirectionalLight
    (it shines along local +z = the direction TOWARDS the light is +z)."""
    wall = assets.quad((-3, -3, -2), (3, -3, -2), (3, 3, -2), (-3, 3, -2), (0, 0, 1))
    scene.create_primitive(GP(scene.load_obj(wall), _lambert(0.5, 0.5, 0.5), Transform.identity()))
    scene.create_primitive(CreatePrimitiveDesc.DirectionalLightPrimitive(2.0, ConstantSpectrum(1.0), Transform.identity()))


def split_paths(closest):
    """Recorded closest-hit rays of a single-threaded oracle render -> list of per-path arrays (a camera ray starts 1e-5 from the origin)."""
    starts = np.nonzero(np.linalg.norm(closest[:, :3], axis=1) < 1e-4)[0]
    return [closest[a:b] for a, b in zip(starts, list(starts[1:]) + [len(closest)])]


def test_q3_directional_shadow_rays_start_on_the_surface(bundle_factory):
    w, h = 24, 18
    b = bundle_factory(wall_under_a_directional_light, w, h, require_gpu=False)
    p = b.oparams("nee", "sobol", 1, max_depth=1, threads=1)
    closest, shadow = b.oracle.record_rays(p)
    paths = split_paths(closest)
    assert len(paths) == w * h and len(shadow) == w * h          # every camera ray hits the wall; one light -> one shadow ray per vertex
    cam = np.concatenate([np.stack([q[0] for q in paths]), np.full((w * h, 1), FMAX, f32)], 1)
    hits, _, _ = b.oracle.trace(cam)
    assert (hits[:, 0] == 0).all()
    # Triangle::intersect's hit position: the barycentric blend of the vertices, left to right in f32 (ray.rs:161-165); the instance
    # transform is the identity and the camera sits at the world origin, so render space = world space
    mesh = b.scene.desc.meshes[0]
    tri = mesh.indices[hits[:, 1]]
    bary = hits[:, 3:6].copy().view(f32)
    v = mesh.positions
    pos = (v[tri[:, 0]] * bary[:, 0:1] + v[tri[:, 1]] * bary[:, 1:2]) + v[tri[:, 2]] * bary[:, 2:3]
    assert pos.dtype == f32
    # q3: origin == hit position to the bit (no move_forward), direction towards the light, unbounded
    assert np.array_equal(shadow[:, :3].view(np.uint32), pos.view(np.uint32))
    assert np.allclose(shadow[:, 3:6], [0, 0, 1], atol=1e-7) and (shadow[:, 6] == FMAX).all()
    # every OTHER shadow ray is pushed 1e-4 along its direction: with that rule these origins would sit at z = -2 + 1e-4
    assert np.abs(shadow[:, 2] + 2.0).max() < 1e-6


@pytest.mark.gpu
def test_q3_gpu_follows_the_oracle_on_the_same_scene(bundle_factory):
    w, h, spp = 24, 18, 16
    b = bundle_factory(wall_under_a_directional_light, w, h)
    for integ in ("nee", "mis"):
        g = b.image(integ, spp).render("sobol").accumulators
        o, _, _ = b.oracle.render(b.oparams(integ, "sobol", spp))
        assert o.mean() > 0.05 and np.abs(g - o).mean() / np.abs(o).mean() < 1e-4
        # a lit wall: the unoffset ray does not shadow the surface it starts on (t_hit < delta_t rejects the self hit, ray.rs:137-158)
        assert (g[..., 1] > 0).mean() > 0.99


# ------------------------------------------------------------------ q12
def two_planes(albedo):
    def load(scene, camera):
        mat = LambertMaterial.new(SpectrumParameter.constant(ConstantSpectrum(albedo)), NormalParameter.none())
        floor = assets.quad((-4, 0, 4), (4, 0, 4), (4, 0, -4), (-4, 0, -4), (0, 1, 0))
        ceil_ = assets.quad((-4, 2, -4), (4, 2, -4), (4, 2, 4), (-4, 2, 4), (0, -1, 0))
        scene.create_primitive(GP(scene.load_obj(floor), mat, Transform.identity()))
        scene.create_primitive(GP(scene.load_obj(ceil_), mat, Transform.identity()))
        scene.create_primitive(CreatePrimitiveDesc.EnvironmentLightPrimitive(1.0, np.full((4, 8, 3), 0.5, f32), Transform.identity()))
        camera.set_look_to((0.0, 1.0, 0.0), (0.0, -0.6, -0.8), (0.0, 1.0, 0.0))
    load.__name__ = f"two_planes_{albedo}"
    return load


LOADERS = {a: two_planes(a) for a in (2.0, 0.5)}


def predicted_u(scene, w, h, px, py, n_bounces, rr_draws):
    """ux of the cosine-hemisphere sample of bounce k = 1.. under the pt integrator: dimensions are wavelength (1), pixel (2), then per
    bounce the lobe selector (1, drawn or skipped: the dimension counts either way) and the direction (2), and -- only if `rr_draws` --
    one Russian-roulette number after each bounce."""
    kinds = [1, 2]
    for _ in range(n_bounces):
        kinds += [1, 2] + ([1] if rr_draws else [])
    vals = scene.sampler_stream("sobol", 1, w, h, 0, px, py, 0, kinds)
    per = 4 if rr_draws else 3
    ux = [vals[3 + per * k + 1] for k in range(n_bounces)]
    rr = [vals[3 + per * k + 3] for k in range(n_bounces)] if rr_draws else None
    return np.array(ux, f32), rr


@pytest.mark.parametrize("albedo", [2.0, 0.5])
def test_q12_russian_roulette_draws_no_sample_at_full_throughput(bundle_factory, albedo):
    """Two facing Lambert planes.  albedo 2 (a ConstantSpectrum): throughput doubles per bounce, max >= 1, RR must neither kill nor consume a
    dimension, so bounce k samples its direction from dimensions 3k+1, 3k+2.  albedo 0.5: max < 1 from the first bounce on, RR draws one
    number per bounce (dimensions shift by one per bounce) and kills when u >= p.  The cosine of a Lambert sample to the surface normal is
    sqrt(1 - ux), which the recorded extension rays show as |d.y|."""
    w, h = 12, 9
    b = bundle_factory(LOADERS[albedo], w, h, require_gpu=False)
    closest, _ = b.oracle.record_rays(b.oparams("pt", "sobol", 1, threads=1))
    paths = split_paths(closest)
    assert len(paths) == w * h
    checked = wrong_model_agrees = 0
    for i, path in enumerate(paths):
        px, py = i % w, i // w
        n_ext = len(path) - 1                       # extension rays actually traced
        if n_ext < 2:
            continue
        k = min(n_ext, 4)
        got = 1.0 - path[1:1 + k, 4].astype(np.float64) ** 2          # 1 - d.y^2 = ux
        ux_free, _ = predicted_u(b.oracle, w, h, px, py, k, rr_draws=False)
        ux_draw, rr = predicted_u(b.oracle, w, h, px, py, k, rr_draws=True)
        right, wrong = (ux_free, ux_draw) if albedo >= 1.0 else (ux_draw, ux_free)
        assert np.abs(got - right).max() < 2e-5, (px, py, got, right)
        wrong_model_agrees += int(np.abs(got[1:] - wrong[1:]).max() < 2e-5)
        checked += 1
    assert checked >= 30 and wrong_model_agrees <= 1
    if albedo < 1.0:
        # the kills, path by path: RR runs after the extension ray of a bounce has hit a surface (base_renderer.rs:76-92), with p = 0.5
        last = np.stack([np.concatenate([q[-1], [FMAX]]) for q in paths]).astype(f32)
        hits, _, _ = b.oracle.trace(last)
        n_killed = 0
        for i, path in enumerate(paths):
            n_ext = len(path) - 1
            _, rr = predicted_u(b.oracle, w, h, i % w, i // w, n_ext + 1, rr_draws=True)
            if any(abs(r - 0.5) < 1e-5 for r in rr[:n_ext]):
                continue
            assert all(r < 0.5 for r in rr[:n_ext - 1]), (i, rr)              # survived every earlier roulette
            if hits[i, 0] >= 0 and n_ext < 16:
                assert rr[n_ext - 1] >= 0.5, (i, rr)                          # ... and the path ended because this one killed it
                n_killed += 1
        assert n_killed > 20
    else:
        # nothing is ever killed: a path only ends because its last ray escaped (or at max_depth)
        last = np.stack([np.concatenate([q[-1], [FMAX]]) for q in paths if len(q) < 17]).astype(f32)
        hits, _, _ = b.oracle.trace(last)
        assert (hits[:, 0] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("albedo", [2.0, 0.5])
def test_q12_gpu_follows_the_oracle_on_the_same_scene(bundle_factory, albedo):
    w, h, spp = 12, 9, 16
    b = bundle_factory(LOADERS[albedo], w, h)
    xy = np.stack(np.meshgrid(np.arange(w), np.arange(h)), -1).reshape(-1, 2).astype(np.uint32)
    for s in (0, 7):
        si = np.full(len(xy), s, np.uint32)
        g = b.image("pt", spp).path_samples("sobol", xy, si)
        o = b.oracle.path_samples(b.oparams("pt", "sobol", spp), xy, si)
        assert np.isfinite(o).all() and o.max() > 0
        # a shifted sampler dimension sends the path somewhere else: it would show as O(1) differences on most paths
        assert (np.abs(g - o).max(1) <= 1e-5 * (1.0 + np.abs(o).max(1))).mean() >= 0.97


# ------------------------------------------------------------------ q26
def test_q26_rgb_above_one_is_refused(tables):
    from oracle import oracle
    ctx = capi.Context(0, require_gpu=False)
    scene = tp.Scene(context=ctx)
    with pytest.raises(capi.TcptError, match="exceeds 1"):
        scene.add_material(scene.desc.material_desc(_lambert(1.2, 0.2, 0.2)))
    assert scene.add_material(scene.desc.material_desc(_lambert(1.0, 0.2, 0.2))) == 0          # exactly 1 is fine (grey 1 -> +inf coefficient)
    with pytest.raises(capi.TcptError):
        scene.rgb_to_coeffs([1.5, 0.2, 0.2], gamma_encoded=False)
    osc = oracle.OracleScene(tables[0], tables[1])
    with pytest.raises(ValueError, match="panics"):
        osc.rgb_to_coeffs([1.5, 0.2, 0.2], gamma_encoded=False)
    cs, _ = osc.rgb_to_coeffs([1.0, 1.0, 1.0], gamma_encoded=False)
    assert cs[0] == 0 and cs[1] == 0 and np.isposinf(cs[2])


# ------------------------------------------------------------------ q27
def rust_binary_search_clamped(cdf, u):
    """slice::binary_search_by(|p| p.partial_cmp(&u).unwrap()) of Rust >= 1.82 (the reference builds with the current stable toolchain,
    flake.nix: fenix.stable; edition 2024), then `Ok(i) => i, Err(i) => i.min(len - 1)` (environment_light.rs:218-223)."""
    size, base = len(cdf), 0
    while size > 1:
        half = size // 2
        mid = base + half
        if not (cdf[mid] > u):        # cmp != Greater
            base = mid
        size -= half
    if cdf[base] == u:
        return base
    return min(base + (1 if cdf[base] < u else 0), len(cdf) - 1)


def cdf_cases():
    rng = np.random.default_rng(11)
    cases = []
    for n in (1, 2, 3, 8, 33, 512):
        wgt = rng.random(n).astype(f32)
        wgt[rng.random(n) < 0.3] = 0.0                                  # zero-weight texels: plateaus (duplicate CDF values)
        if not wgt.any():
            wgt[0] = 1.0
        run = np.cumsum(wgt, dtype=f32)
        cdf = (run / run[-1]).astype(f32)
        u = np.concatenate([cdf, np.nextafter(cdf, f32(0)), np.nextafter(cdf, f32(2)), rng.random(200).astype(f32), [f32(0), f32(1), f32(0.5), f32(0.25)]]).astype(f32)
        cases.append((cdf, u[(u >= 0) & (u <= 1)]))
    eq = np.array([0.25, 0.5, 0.75, 1.0], f32)                          # four equal texels: Sobol points hit the entries exactly
    cases.append((eq, np.array([0, 0.25, 0.5, 0.75, 1.0, 0.1, 0.3, 0.6, 0.9], f32)))
    return cases


def test_q27_oracle_cdf_search_is_rusts_binary_search():
    from oracle import oracle
    for cdf, u in cdf_cases():
        want = np.array([rust_binary_search_clamped(cdf, x) for x in u], np.uint32)
        assert np.array_equal(oracle.cdf_search(cdf, u), want)
    # the exact-match branch decides: u == cdf[i] returns i, not the insertion point i + 1
    eq = np.array([0.25, 0.5, 0.75, 1.0], f32)
    assert oracle.cdf_search(eq, np.array([0.5, np.nextafter(f32(0.5), f32(1))], f32)).tolist() == [1, 2]


@pytest.mark.gpu
@pytest.mark.parametrize("cells_per_entry", [1, 4])
def test_q27_device_guide_table_search_is_rusts_binary_search(cells_per_entry):
    ctx = capi.Context(0)
    for cdf, u in cdf_cases():
        g = 1
        while g < len(cdf):
            g *= 2
        out = np.zeros(len(u), np.uint32)
        ctx.check(ctx.lib.tcpt_cdf_search(ctx.handle, capi.as_ptr(cdf, C.c_float), len(cdf), g * cells_per_entry, capi.as_ptr(u, C.c_float), len(u), capi.as_ptr(out, C.c_uint32)))
        want = np.array([rust_binary_search_clamped(cdf, x) for x in u], np.uint32)
        # on a plateau the reference may return any of the equal entries' indices depending on its probing sequence; the device's answer must be
        # the reference's wherever the entry is unique, and an entry with the same CDF value (same texel weight 0 -> same pdf 0) otherwise
        same = out == want
        assert (cdf[out[~same]] == cdf[want[~same]]).all() if (~same).any() else True
        unique = np.array([(cdf == cdf[k]).sum() == 1 for k in want])
        assert same[unique].all()
