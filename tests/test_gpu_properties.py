"""Size-independent properties of the GPU path at the configs' full sizes, plus the reference's own cross-integrator test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_run_to_run_bitwise_reproducible(bundle_factory):
    b = bundle_factory(3, 200, 150)
    a = b.image("mis", 64).render("sobol").accumulators.copy()
    c = b.image("mis", 64).render("sobol").accumulators.copy()
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


def test_pass_tiling_does_not_change_the_film(bundle_factory):
    """The film kernel adds each pixel's samples in sample order, so the slot budget (how many passes a frame takes) must not
    change a single bit."""
    b = bundle_factory(3, 200, 150)
    full = b.image("mis", 64).render("sobol").accumulators.copy()
    tiled = b.image("mis", 64).render("sobol", max_slots=50000).accumulators.copy()   # 30 000 px -> 1 sample per pass, 64 passes
    tiny = b.image("mis", 64).render("sobol", max_slots=7777).accumulators.copy()     # pixel chunks smaller than the frame
    assert np.array_equal(full.view(np.uint32), tiled.view(np.uint32))
    assert np.array_equal(full.view(np.uint32), tiny.view(np.uint32))


@pytest.mark.parametrize("spp", [8, 64])   # odd and even log2(spp): the sample digits sit at odd / even bit offsets
def test_sobol_prefix_table_is_bit_transparent(bundle_factory, spp):
    """The Z-Sobol pixel-prefix table only caches the pixel digits of get_sample_index: table on, table off and a table that
    covers only the first dimensions (the rest computed in full) must give the same film to the bit."""
    b = bundle_factory(3, 200, 150)
    ctx = b.scene.ctx
    try:
        on = b.image("mis", spp).render("sobol")
        assert on.stats["sobol_prefix_bytes"] == 200 * 150 * 4 * (3 + 8 * 17)
        a = on.accumulators.copy()
        ctx.set_option("sobol_prefix", 0)
        off = b.image("mis", spp).render("sobol")
        assert off.stats["sobol_prefix_bytes"] == 0
        ctx.set_option("sobol_prefix", 1)
        ctx.set_option("sobol_prefix_mb", 1)          # 1 MiB / (30 000 px x 4 B) = 8 dimensions
        b.image("mis", spp * 2).render("sobol")       # another spp in between: the cached table must be rebuilt
        part = b.image("mis", spp).render("sobol")
        assert part.stats["sobol_prefix_bytes"] == 200 * 150 * 4 * 8
        assert np.array_equal(a.view(np.uint32), off.accumulators.view(np.uint32))
        assert np.array_equal(a.view(np.uint32), part.accumulators.view(np.uint32))
    finally:
        ctx.set_option("sobol_prefix", 1)
        ctx.set_option("sobol_prefix_mb", 8192)


def test_sobol_hash_table_is_bit_transparent(bundle_factory):
    """The per-dimension Owen-scramble seeds hash(dimension, seed) come from a table built when the seed changes: table on, table off and a
    change of seed in between must give the same films to the bit."""
    b = bundle_factory(19, 200, 150)
    ctx = b.scene.ctx
    try:
        on = [b.image("mis", 16, seed=s).render("sobol").accumulators.copy() for s in (0, 7, 0)]
        ctx.set_option("sobol_hash", 0)
        off = [b.image("mis", 16, seed=s).render("sobol").accumulators.copy() for s in (0, 7)]
    finally:
        ctx.set_option("sobol_hash", 1)
    assert np.array_equal(on[0].view(np.uint32), off[0].view(np.uint32))
    assert np.array_equal(on[1].view(np.uint32), off[1].view(np.uint32))
    assert np.array_equal(on[0].view(np.uint32), on[2].view(np.uint32))
    assert not np.array_equal(on[0].view(np.uint32), on[1].view(np.uint32))


@pytest.mark.parametrize("scene_id,integrator", [(19, "mis"), (10, "nee"), (8, "pt")])
def test_fused_launches_do_not_change_the_film(bundle_factory, scene_id, integrator):
    """Two launches per bounce (k_trace_fused, k_shade_all) against one launch per queue and per shading bucket: the same
    work in another launch structure, so film and ray counts must be identical to the bit."""
    b = bundle_factory(scene_id, 200, 150)
    ctx = b.scene.ctx
    try:
        fused = b.image(integrator, 16).render("sobol")
        a, sa = fused.accumulators.copy(), dict(fused.stats)
        ctx.set_option("fused_launches", 0)
        plain = b.image(integrator, 16).render("sobol")
    finally:
        ctx.set_option("fused_launches", 1)
    assert np.array_equal(a.view(np.uint32), plain.accumulators.view(np.uint32))
    assert (sa["closest_rays"], sa["shadow_rays"], sa["paths"]) == (plain.stats["closest_rays"], plain.stats["shadow_rays"], plain.stats["paths"])
    assert sa["kernel_launches"] < plain.stats["kernel_launches"] / 2


@pytest.mark.parametrize("scene_id", [19, 3, 2])   # one environment light / one emissive mesh / one point light
def test_single_light_shortcut_is_bit_transparent(bundle_factory, scene_id):
    """With one light of strictly positive power its selection probability is w / w = 1 for every wavelength set; skipping the
    per-vertex power table must not change a bit."""
    b = bundle_factory(scene_id, 200, 150)
    ctx = b.scene.ctx
    a = b.image("mis", 16).render("sobol").accumulators.copy()
    try:
        ctx.set_option("light_shortcut", 0)
        b.scene.build(b.camera)                      # the flag is decided when the scene is uploaded
        c = b.image("mis", 16).render("sobol").accumulators.copy()
    finally:
        ctx.set_option("light_shortcut", 1)
        b.scene.build(b.camera)
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


@pytest.mark.parametrize("scene_id,integrator", [(19, "mis"), (19, "nee"), (17, "mis")])   # scenes lit by an environment map
def test_environment_nee_table_is_bit_transparent(bundle_factory, scene_id, integrator):
    """Light sampling draws a TEXEL of the environment map; the direction through its centre, the pdf of that direction and the
    illuminant spectrum looked up there (environment_light.rs:326-350, :234-259, :304-316) only depend on the texel, so they are
    evaluated once per texel when the scene is uploaded (k_env_nee_table, same device code) and read back per sample: not a bit may change."""
    b = bundle_factory(scene_id, 200, 150)
    ctx = b.scene.ctx
    a = b.image(integrator, 16).render("sobol").accumulators.copy()
    try:
        ctx.set_option("env_nee_table", 0)
        b.scene.build(b.camera)                      # the table is built when the scene is uploaded
        c = b.image(integrator, 16).render("sobol").accumulators.copy()
    finally:
        ctx.set_option("env_nee_table", 1)
        b.scene.build(b.camera)
    assert a.any() and np.array_equal(a.view(np.uint32), c.view(np.uint32))


@pytest.mark.parametrize("integrator", ["mis", "pt"])
def test_illuminant_half_shortcut_is_bit_transparent(bundle_factory, integrator):
    """The colour of an RgbIlluminantSpectrum is divided by twice its largest component (rgb_illuminant_spectrum.rs:27-40), so that component
    is exactly 0.5: its sRGB decoding and the z-node interval of every environment lookup are constants, evaluated once by the device code
    the lookups run (k_illum_half).  With the shortcut, without it, and without the per-texel table (the shortcut then also serves light
    sampling): the same film to the bit."""
    b = bundle_factory(19, 200, 150)
    ctx = b.scene.ctx
    a = b.image(integrator, 16).render("sobol").accumulators.copy()
    try:
        ctx.set_option("illum_half", 0)
        b.scene.build(b.camera)                      # the constants ride in the uploaded scene
        c = b.image(integrator, 16).render("sobol").accumulators.copy()
        ctx.set_option("env_nee_table", 0)
        b.scene.build(b.camera)
        d = b.image(integrator, 16).render("sobol").accumulators.copy()
        ctx.set_option("illum_half", 1)
        b.scene.build(b.camera)
        e = b.image(integrator, 16).render("sobol").accumulators.copy()
    finally:
        ctx.set_option("illum_half", 1)
        ctx.set_option("env_nee_table", 1)
        b.scene.build(b.camera)
    assert a.any()
    for other in (c, d, e):
        assert np.array_equal(a.view(np.uint32), other.view(np.uint32))


@pytest.mark.parametrize("scene_id,spp", [(3, 64), (19, 256), (10, 16)])
def test_sobol_pass_table_is_bit_transparent(bundle_factory, scene_id, spp):
    """The per-pass table caches the permuted sample digits that all samples of a pixel share inside one pass (and the permutation row
    of the first digit that varies): on, off, covering only a few dimensions, with passes of 4 / 16 / odd numbers of samples, with
    row shards and with sample ranges that are not aligned to their length -- always the same film to the bit."""
    b = bundle_factory(scene_id, 200, 150)
    ctx = b.scene.ctx
    n = 200 * 150
    try:
        ctx.set_option("sobol_pass", 0)
        ref = b.image("mis", spp).render("sobol").accumulators.copy()
        ctx.set_option("sobol_pass", 1)
        for dims, slots in ((40, 0), (40, 16 * n), (5, 4 * n), (200, 7 * n), (40, 16 * 7777)):
            ctx.set_option("sobol_pass_dims", dims)
            img = b.image("mis", spp).render("sobol", max_slots=slots)
            assert np.array_equal(ref.view(np.uint32), img.accumulators.view(np.uint32)), (dims, slots)
        ctx.set_option("sobol_pass_dims", 11)
        parts = [b.image("mis", spp).render("sobol", row_offset=r, row_stride=3).accumulators.copy() for r in range(3)]
        assert np.array_equal(sum(parts).view(np.uint32), ref.view(np.uint32))
        # sample ranges [3, 3 + 9) and [12, spp): the digits that vary inside a range depend on where it starts
        ctx.set_option("sobol_pass", 0)
        a = [b.image("mis", spp).render("sobol", spp_begin=s0, spp_end=s1).accumulators.copy() for s0, s1 in ((3, 12), (12, spp))]
        ctx.set_option("sobol_pass", 1)
        c = [b.image("mis", spp).render("sobol", spp_begin=s0, spp_end=s1).accumulators.copy() for s0, s1 in ((3, 12), (12, spp))]
        for x, y in zip(a, c):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    finally:
        ctx.set_option("sobol_pass", 1)
        ctx.set_option("sobol_pass_dims", 11)


def test_row_shards_sum_bitwise_to_the_full_frame(bundle_factory):
    """Tile (row-interleaved) sharding: each pixel is rendered entirely by one shard; the others hold exact zeros."""
    b = bundle_factory(10, 200, 150)
    full = b.image("nee", 32).render("sobol").accumulators.copy()
    for world in (2, 3, 8):
        parts = [b.image("nee", 32).render("sobol", row_offset=r, row_stride=world).accumulators.copy() for r in range(world)]
        for r, p in enumerate(parts):
            mask = np.ones(150, bool); mask[r::world] = False
            assert not p[mask].any()
        assert np.array_equal(sum(parts).view(np.uint32), full.view(np.uint32))


def test_spp_shards_sum_to_the_full_frame(bundle_factory):
    """Sample-range sharding: the same samples, per-pixel sums re-associated across shards."""
    b = bundle_factory(3, 200, 150)
    full = b.image("mis", 64).render("sobol").accumulators
    parts = sum(b.image("mis", 64).render("sobol", spp_begin=s, spp_end=s + 16).accumulators.astype(np.float64) for s in range(0, 64, 16))
    assert np.abs(parts - full).max() <= 1e-4 * max(1.0, np.abs(full).max())


def test_full_size_config1_counts_and_energy(bundle_factory):
    """BASELINE.json configs[0] at full size (scene 3, MIS + Sobol, 200x150x512): 15.36 M paths, every one accounted for, a
    finite film, and the duplicated Sobol pairs of odd log2(spp) (Appendix A q15-i) visible as pairwise-equal path samples."""
    b = bundle_factory(3, 200, 150)
    img = b.image("mis", 512).render("sobol")
    assert img.stats["paths"] == 200 * 150 * 512
    assert img.stats["closest_rays"] > img.stats["paths"] and img.stats["shadow_rays"] > 0
    assert np.isfinite(img.accumulators).all() and (img.pixels >= 0).all() and (img.pixels <= 1).all()
    xy = np.array([[50, 60]] * 8, dtype=np.uint32)
    si = np.array([0, 1, 2, 3, 100, 101, 510, 511], dtype=np.uint32)
    s = b.image("mis", 512).path_samples("sobol", xy, si)
    assert np.array_equal(s[0::2], s[1::2])


def linearise_gamma22(img):
    return np.power(np.clip(img, 0, 1).astype(np.float64), 2.2)


def median3(img):
    pad = np.pad(img, ((1, 1), (1, 1), (0, 0)), mode="edge")
    stack = np.stack([pad[dy:dy + img.shape[0], dx:dx + img.shape[1]] for dy in range(3) for dx in range(3)], 0)
    return np.median(stack, 0)


@pytest.mark.parametrize("scene_id", [3, 7])
def test_cross_integrator_consistency_like_the_reference(bundle_factory, scene_id):
    """renderer/tests/renderer_consistency_test.rs:319-353: pt vs nee and pt vs mis, random sampler, 2048 spp, 200x150, u8
    images, 3x3 median, RMSE in gamma-2.2-linearised space <= 0.013."""
    b = bundle_factory(scene_id, 200, 150)
    imgs = {}
    for integ in ("pt", "nee", "mis"):
        u8 = b.image(integ, 2048).render("random").to_u8()
        imgs[integ] = linearise_gamma22(median3(u8.astype(np.float64)) / 255.0)
    for other in ("nee", "mis"):
        rmse = np.sqrt(np.mean((imgs["pt"] - imgs[other]) ** 2))
        assert rmse <= 0.013, f"pt vs {other}: RMSE {rmse:.4f}"


def test_cross_integrator_means_with_dispersive_glass(bundle_factory):
    """Scene 8 (SF11 glass: caustics, hero-wavelength collapse after `terminate_secondary`) is too noisy in pt for the image-level
    RMSE bound above at 2048 spp, but the three estimators still integrate the same image: per-channel frame means agree."""
    b = bundle_factory(8, 200, 150)
    means = {i: b.image(i, 4096).render("random").accumulators.astype(np.float64).mean(axis=(0, 1)) / 4096 for i in ("pt", "nee", "mis")}
    assert np.allclose(means["nee"], means["mis"], rtol=0.015), means
    assert np.allclose(means["pt"], means["mis"], rtol=0.015), means
