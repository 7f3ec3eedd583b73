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
