"""The BENCHMARKED code path under the oracle, at the sizes BASELINE.json quotes (task 1).

C4 = scene 19, 3840x2160, 4096 spp, MIS + Z-Sobol and C3 = scene 17, 1920x1080, 1024 spp: what bench.py times goes through
tcpt_render with the Z-Sobol PIXEL-PREFIX table and the PER-PASS table live, with Morton codes that lose their top bits in the
reference's u32 (x or y >= 2048 at 4096 spp: z_sobol_sampler.rs:54-65,200) and Sobol indices beyond 2^32.  Rendering
whole frames of that size on the oracle is out of reach (34 G paths), so the frame is cut the way the sharding API cuts it: a
row shard (every 270th / 135th row) and a block of 16 sample indices, rendered by tcpt_render exactly as a bench step renders
them, against the in-order sum of the oracle's per-path sensor contributions for the same (pixel, sample) pairs.
C1 / C2 (scenes 3 and 10, 200x150, 512 spp, the reference's test resolution: regression_test.rs:253-264,583-659) are compared as
whole frames, all 1 + 6 integrator / sampler combinations."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MRE_TOL = 1e-3


def oracle_shard_sum(b, integrator, w, h, spp, rows, s0, s1):
    """In-order f32 sum over samples [s0, s1) of the oracle's per-path contributions, for every pixel of `rows`."""
    ys, xs = np.meshgrid(np.asarray(rows, np.uint32), np.arange(w, dtype=np.uint32), indexing="ij")
    xy = np.stack([xs.ravel(), ys.ravel()], 1)
    acc = np.zeros((len(xy), 3), np.float32)
    p = b.oparams(integrator, "sobol", spp)
    for s in range(s0, s1):     # sample order = the film's summation order (sensor.rs:76-77)
        acc += b.oracle.path_samples(p, xy, np.full(len(xy), s, np.uint32))
    return acc.reshape(len(rows), w, 3)


@pytest.mark.parametrize("scene_id,w,h,spp,stride,blocks", [
    (19, 3840, 2160, 4096, 270, [(0, 0), (1, 2048), (269, 4080)]),     # rows 269, 539, ..., 2159: y >= 2048 in the last one; sample blocks below / at / above 2048
    (17, 1920, 1080, 1024, 135, [(0, 0), (134, 1008)]),
])
def test_bench_step_slices_match_the_oracle(bundle_factory, scene_id, w, h, spp, stride, blocks):
    b = bundle_factory(scene_id, w, h)
    for offset, s0 in blocks:
        rows = list(range(offset, h, stride))
        img = b.image("mis", spp).render("sobol", row_offset=offset, row_stride=stride, spp_begin=s0, spp_end=s0 + 16)
        # both Z-Sobol tables were live: the prefix table covers every dimension a depth-16 path can draw, the pass ran 16 samples deep
        assert img.stats["sobol_prefix_bytes"] == w * h * 4 * (3 + 8 * 17)
        assert img.stats["paths"] == len(rows) * w * 16 and img.stats["passes"] == 1
        g = img.accumulators[rows]
        other = np.ones(h, bool); other[rows] = False
        assert not img.accumulators[other].any()
        o = oracle_shard_sum(b, "mis", w, h, spp, rows, s0, s0 + 16)
        mre = np.abs(g - o).mean() / np.abs(o).mean()
        assert mre <= MRE_TOL, f"rows {offset}::{stride}, samples [{s0}, {s0 + 16}): mean relative error {mre:.3e}"
        # most pixels agree to the last bit over all 16 samples; anything systematic in the sampler tables would break every pixel
        same = (g == o).all(axis=2).mean()
        assert same >= (0.08 if scene_id == 19 else 0.3), f"only {same:.2%} of the pixels are bit-identical"
        # the truncated-Morton region by itself (x >= 2048 or y >= 2048)
        if w > 2048:
            gt, ot = g[:, 2048:], o[:, 2048:]
            assert np.abs(gt - ot).mean() / np.abs(ot).mean() <= MRE_TOL


@pytest.mark.parametrize("scene_id,w,h,spp", [(19, 3840, 2160, 4096), (17, 1920, 1080, 1024)])
def test_individual_paths_at_bench_size(bundle_factory, scene_id, w, h, spp):
    """5 000 random (pixel, sample) pairs of the benchmarked frames, prefix table live, incl. pixels beyond 2048 and samples beyond 2048."""
    b = bundle_factory(scene_id, w, h)
    rng = np.random.default_rng(100 + scene_id)
    n = 5000
    xy = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.uint32)
    xy[: n // 4, 0] = rng.integers(min(2048, w - 1), w, n // 4)      # a quarter of them with x in the truncated range
    xy[n // 4: n // 2, 1] = rng.integers(min(2048, h - 1), h, n // 4)
    si = rng.integers(0, spp, n).astype(np.uint32)
    si[::5] = rng.integers(spp // 2, spp, len(si[::5]))
    img = b.image("mis", spp)
    g = img.path_samples("sobol", xy, si)
    assert img.stats["sobol_prefix_bytes"] > 0
    o = b.oracle.path_samples(b.oparams("mis", "sobol", spp), xy, si)
    err = np.abs(g - o).max(1)
    rel = err / (np.abs(o).max(1) + 1e-6)
    assert (rel > 1e-4).mean() <= 2e-3, f"{(rel > 1e-4).sum()} of {n} paths differ by more than 1e-4"
    assert (err == 0).mean() >= (0.4 if scene_id == 19 else 0.8)


C12 = [(3, "mis", "sobol")] + [(10, i, s) for i in ("pt", "nee", "mis") for s in ("random", "sobol")]


@pytest.mark.parametrize("scene_id,integrator,sampler", C12, ids=[f"s{c[0]}-{c[1]}-{c[2]}" for c in C12])
def test_config1_and_config2_full_frames(bundle_factory, scene_id, integrator, sampler):
    """BASELINE configs[0] and configs[1] at their stated size: 200 x 150, 512 spp (odd log2: duplicated Sobol pairs, q15-i)."""
    w, h, spp = 200, 150, 512
    b = bundle_factory(scene_id, w, h)
    img = b.image(integrator, spp).render(sampler)
    acc, srgb, st = b.oracle.render(b.oparams(integrator, sampler, spp))
    assert img.stats["paths"] == w * h * spp == st["paths"]
    assert abs(img.stats["closest_rays"] - st["closest_rays"]) <= 1e-3 * st["closest_rays"]
    assert abs(img.stats["shadow_rays"] - st["shadow_rays"]) <= 1e-3 * max(1, st["shadow_rays"])
    g, o = img.accumulators / spp, acc / spp
    assert np.isfinite(g).all() and np.isfinite(o).all()
    mre = np.abs(g - o).mean() / np.abs(o).mean()
    assert mre <= MRE_TOL, f"mean relative error {mre:.3e}"
    # the image the reference's regression test would look at: u8 after the truncating quantiser (renderer.rs:140-144)
    u8g = img.to_u8().astype(np.int32)
    u8o = np.clip(srgb * np.float32(255.0), 0, 255).astype(np.uint8).astype(np.int32)
    assert (np.abs(u8g - u8o) > 1).mean() <= 1e-3 and np.abs(u8g - u8o).max() <= 8


@pytest.mark.parametrize("spp", [6, 100])
def test_spp_that_is_not_a_power_of_two(bundle_factory, spp):
    """The reference takes any spp (main.rs only warns): log2_spp = floor(log2(spp)), and samples >= 2^log2_spp OR their high bit into
    the pixel's Morton digits (z_sobol_sampler.rs:200), so the pixel-prefix / pass tables do not apply: the device must fall back to the
    full digit loop, bit-identical to tables-off, and agree with the oracle."""
    w, h = 64, 48
    b = bundle_factory(3, w, h)
    ctx = b.scene.ctx
    img = b.image("mis", spp).render("sobol")
    assert img.stats["sobol_prefix_bytes"] == 0
    a = img.accumulators.copy()
    try:
        ctx.set_option("sobol_prefix", 0)
        off = b.image("mis", spp).render("sobol").accumulators.copy()
    finally:
        ctx.set_option("sobol_prefix", 1)
    assert np.array_equal(a.view(np.uint32), off.view(np.uint32))
    acc, _, st = b.oracle.render(b.oparams("mis", "sobol", spp))
    assert abs(img.stats["closest_rays"] - st["closest_rays"]) <= 1e-3 * st["closest_rays"]
    assert np.abs(a - acc).mean() / np.abs(acc).mean() <= MRE_TOL
    n = 2000
    rng = np.random.default_rng(spp)
    xy = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.uint32)
    si = rng.integers(0, spp, n).astype(np.uint32)
    si[: n // 2] = rng.integers(1 << int(np.log2(spp)), spp, n // 2)     # the samples that spill into the pixel digits
    g = b.image("mis", spp).path_samples("sobol", xy, si)
    o = b.oracle.path_samples(b.oparams("mis", "sobol", spp), xy, si)
    assert (np.abs(g - o).max(1) == 0).mean() >= 0.8
    # the sampler stream itself, bit for bit
    kinds = [1, 2, 1, 2, 2, 1] * 4
    for (px, py), s in zip(xy[:50], si[:50]):
        gs = b.scene.sampler_stream("sobol", spp, w, h, 0, int(px), int(py), int(s), kinds)
        os_ = b.oracle.sampler_stream("sobol", spp, w, h, 0, int(px), int(py), int(s), kinds)
        assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32))
