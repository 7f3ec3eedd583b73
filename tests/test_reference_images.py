"""renderer/tests/regression_test.rs on the GPU path: the 42 reference-image comparisons, runnable the moment the LFS payloads exist."""
import numpy as np
import pytest

from toy_cpu_pathtracing_b200 import regression


def test_rmse_matches_the_reference_definition():
    a = np.zeros((2, 2, 3), dtype=np.uint8); b = a.copy()
    assert regression.calculate_rmse(a, b) == 0.0
    b[0, 0, 0] = 255      # one of 12 channel values differs by linear 1.0
    assert abs(regression.calculate_rmse(a, b) - (1.0 / 12.0) ** 0.5) < 1e-12
    b[:] = 10             # below the sRGB knee: linear = c / 12.92
    assert abs(regression.calculate_rmse(a, b) - (10 / 255) / 12.92) < 1e-12


def test_the_case_table_is_the_references():
    cases = regression.REGRESSION_CASES
    assert len(cases) == 42 and len({c[4] for c in cases}) == 42
    assert sorted({c[0] for c in cases}) == [0, 3, 6, 7, 8, 9, 10]
    assert (8, "mis", "sobol", 2048, "reference_scene8_mis_sobol.png", 0.08) in cases and (0, "pt", "random", 512, "reference_pt_random.png", 0.05) in cases


@pytest.mark.gpu
@pytest.mark.parametrize("case", regression.REGRESSION_CASES, ids=[c[4][:-4] for c in regression.REGRESSION_CASES])
def test_reference_image(bundle_factory, case):
    scene_id, integrator, sampler, spp, name, max_rmse = case
    ref = regression.reference_image(name)
    from toy_cpu_pathtracing_b200 import scenes
    if ref is None or scenes.real_asset_path("bunny") is None:
        pytest.skip("test_references / renderer/assets payloads are git-LFS stubs here: set TCPT_REFERENCE_DIR and TCPT_ASSET_DIR")
    b = bundle_factory(scene_id, 200, 150)
    img = b.image(integrator, spp).render(sampler).to_u8()
    rmse = regression.calculate_rmse(img, ref)
    assert rmse <= max_rmse, f"RMSE {rmse:.6f} exceeds {max_rmse} for {name}"


DRY_SPP = 32


@pytest.mark.gpu
@pytest.mark.parametrize("case", regression.REGRESSION_CASES, ids=["dry_" + c[4][:-4] for c in regression.REGRESSION_CASES])
def test_regression_harness_dry_run(bundle_factory, tmp_path, case):
    """The 42 cases of regression_test.rs executed end to end while the real reference PNGs are LFS stubs: the stand-in "reference" of a
    case is the ORACLE's render of the same scene / integrator / sampler at the test resolution, quantised and written with the
    reference's own rule ((v * 255) as u8, renderer.rs:140-144); the GPU render goes through RendererImage.save, both files are read
    back and compared with calculate_rmse against the case's own threshold.  (DRY_SPP samples instead of 512-2048: the oracle renders
    42 frames here; the full-spp frames of scenes 3 and 10 are compared in test_gpu_full_size.py.)"""
    import cv2
    scene_id, integrator, sampler, _spp, name, max_rmse = case
    b = bundle_factory(scene_id, 200, 150)
    _, srgb, _ = b.oracle.render(b.oparams(integrator, sampler, DRY_SPP))
    ref_u8 = np.clip(np.nan_to_num(srgb * np.float32(255.0), nan=0.0), 0, 255).astype(np.uint8)
    ref_path = tmp_path / name
    assert cv2.imwrite(str(ref_path), ref_u8[..., ::-1])
    rmse = regression.run_render_and_compare(b.image(integrator, DRY_SPP), sampler, str(tmp_path / ("out_" + name)), str(ref_path), max_rmse)
    assert rmse <= 0.02, f"GPU and oracle frames of the same samples should be far inside the noise threshold, got {rmse}"
    assert not (tmp_path / ("out_" + name)).exists()      # the harness removes its output


@pytest.mark.gpu
def test_regression_harness_detects_a_wrong_image(bundle_factory, tmp_path):
    """Negative control: a reference rendered with another scene must trip the threshold."""
    import cv2
    b, other = bundle_factory(3, 200, 150), bundle_factory(10, 200, 150)
    _, srgb, _ = other.oracle.render(other.oparams("mis", "sobol", 8))
    ref_path = tmp_path / "wrong.png"
    cv2.imwrite(str(ref_path), np.clip(srgb * np.float32(255.0), 0, 255).astype(np.uint8)[..., ::-1])
    with pytest.raises(AssertionError, match="exceeds threshold"):
        regression.run_render_and_compare(b.image("mis", 8), "sobol", str(tmp_path / "out.png"), str(ref_path), 0.05)
