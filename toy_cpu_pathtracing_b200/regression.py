"""The reference's image-regression contract (renderer/tests/regression_test.rs): 42 renders at 200x150 compared with
test_references/*.png by RMSE in linearised sRGB.  The PNG payloads (and the assets the scenes load) are git-LFS objects that are absent
from the checkout this backend was built against, so the comparison only becomes meaningful once TCPT_REFERENCE_DIR (the
test_references/ directory) and TCPT_ASSET_DIR (renderer/assets/) point at real payloads; tests/test_reference_images.py skips until then."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np

# (scene, renderer, sampler, spp, reference file, max RMSE): regression_test.rs:110-659
REGRESSION_CASES = [
    (0, 'pt', 'random', 512, 'reference_pt_random.png', 0.05),
    (0, 'pt', 'sobol', 512, 'reference_pt_sobol.png', 0.05),
    (0, 'nee', 'random', 512, 'reference_nee_random.png', 0.05),
    (0, 'nee', 'sobol', 512, 'reference_nee_sobol.png', 0.05),
    (0, 'mis', 'random', 512, 'reference_mis_random.png', 0.05),
    (0, 'mis', 'sobol', 512, 'reference_mis_sobol.png', 0.05),
    (3, 'pt', 'random', 512, 'reference_scene3_pt_random.png', 0.05),
    (3, 'pt', 'sobol', 512, 'reference_scene3_pt_sobol.png', 0.05),
    (3, 'nee', 'random', 512, 'reference_scene3_nee_random.png', 0.05),
    (3, 'nee', 'sobol', 512, 'reference_scene3_nee_sobol.png', 0.05),
    (3, 'mis', 'random', 512, 'reference_scene3_mis_random.png', 0.05),
    (3, 'mis', 'sobol', 512, 'reference_scene3_mis_sobol.png', 0.05),
    (6, 'pt', 'random', 512, 'reference_scene6_pt_random.png', 0.05),
    (6, 'pt', 'sobol', 512, 'reference_scene6_pt_sobol.png', 0.05),
    (6, 'nee', 'random', 512, 'reference_scene6_nee_random.png', 0.05),
    (6, 'nee', 'sobol', 512, 'reference_scene6_nee_sobol.png', 0.05),
    (6, 'mis', 'random', 512, 'reference_scene6_mis_random.png', 0.05),
    (6, 'mis', 'sobol', 512, 'reference_scene6_mis_sobol.png', 0.05),
    (7, 'pt', 'random', 1024, 'reference_scene7_pt_random.png', 0.05),
    (7, 'pt', 'sobol', 1024, 'reference_scene7_pt_sobol.png', 0.05),
    (7, 'nee', 'random', 512, 'reference_scene7_nee_random.png', 0.05),
    (7, 'nee', 'sobol', 512, 'reference_scene7_nee_sobol.png', 0.05),
    (7, 'mis', 'random', 512, 'reference_scene7_mis_random.png', 0.05),
    (7, 'mis', 'sobol', 512, 'reference_scene7_mis_sobol.png', 0.05),
    (8, 'pt', 'random', 2048, 'reference_scene8_pt_random.png', 0.075),
    (8, 'pt', 'sobol', 2048, 'reference_scene8_pt_sobol.png', 0.085),
    (8, 'nee', 'random', 2048, 'reference_scene8_nee_random.png', 0.075),
    (8, 'nee', 'sobol', 2048, 'reference_scene8_nee_sobol.png', 0.085),
    (8, 'mis', 'random', 2048, 'reference_scene8_mis_random.png', 0.075),
    (8, 'mis', 'sobol', 2048, 'reference_scene8_mis_sobol.png', 0.08),
    (9, 'pt', 'random', 512, 'reference_scene9_pt_random.png', 0.06),
    (9, 'pt', 'sobol', 512, 'reference_scene9_pt_sobol.png', 0.06),
    (9, 'nee', 'random', 512, 'reference_scene9_nee_random.png', 0.06),
    (9, 'nee', 'sobol', 512, 'reference_scene9_nee_sobol.png', 0.06),
    (9, 'mis', 'random', 512, 'reference_scene9_mis_random.png', 0.06),
    (9, 'mis', 'sobol', 512, 'reference_scene9_mis_sobol.png', 0.06),
    (10, 'pt', 'random', 512, 'reference_scene10_pt_random.png', 0.06),
    (10, 'pt', 'sobol', 512, 'reference_scene10_pt_sobol.png', 0.06),
    (10, 'nee', 'random', 512, 'reference_scene10_nee_random.png', 0.06),
    (10, 'nee', 'sobol', 512, 'reference_scene10_nee_sobol.png', 0.06),
    (10, 'mis', 'random', 512, 'reference_scene10_mis_random.png', 0.06),
    (10, 'mis', 'sobol', 512, 'reference_scene10_mis_sobol.png', 0.06),
]


def srgb_to_linear(c: np.ndarray) -> np.ndarray:       # regression_test.rs:6-12
    c = np.asarray(c, dtype=np.float64)
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def calculate_rmse(img1_u8: np.ndarray, img2_u8: np.ndarray) -> float:   # regression_test.rs:14-38
    if img1_u8.shape != img2_u8.shape:
        raise ValueError("image dimensions differ")
    d = srgb_to_linear(img1_u8.astype(np.float64) / 255.0) - srgb_to_linear(img2_u8.astype(np.float64) / 255.0)
    return float(np.sqrt(np.mean(d * d)))


def open_rgb8(path):
    """image::open(path).to_rgb8() of the harness (regression_test.rs:77-83): decode, then the image crate's conversion (libtcpt)."""
    from .scene import convert_image, decode_image
    return convert_image(decode_image(path), "rgb8")


def run_render_and_compare(image, sampler, output_file, reference_file, max_rmse, **render_kw):
    """run_render_and_compare (regression_test.rs:41-107) with the renderer in-process: render -> RendererImage.save (the truncating PNG
    writer of renderer.rs:136-149) -> open both files -> calculate_rmse -> threshold.  Returns the RMSE; raises AssertionError like the
    reference's assert!, and removes the rendered file afterwards (regression_test.rs:103-106)."""
    image.render(sampler, **render_kw)
    image.save(output_file)
    try:
        rmse = calculate_rmse(open_rgb8(output_file), open_rgb8(reference_file))
    finally:
        try:
            os.remove(output_file)
        except OSError:
            pass
    assert rmse <= max_rmse, f"RMSE {rmse:.6f} exceeds threshold {max_rmse:.6f} for {output_file} vs {reference_file}"
    return rmse


def reference_image(name: str):
    """The decoded RGB8 reference PNG, or None when its payload is not available (LFS pointer stub / no TCPT_REFERENCE_DIR)."""
    root = os.environ.get("TCPT_REFERENCE_DIR")
    if not root:
        return None
    p = Path(root) / name
    if not p.is_file() or p.read_bytes()[:8] != b"\x89PNG\r\n\x1a\n":
        return None
    import cv2
    img = cv2.imread(str(p), cv2.IMREAD_COLOR)
    return None if img is None else np.ascontiguousarray(img[..., ::-1])
