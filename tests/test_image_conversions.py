"""The `image` crate conversions behind the reference's texture loaders: scene/src/texture/loader.rs:43-87
(`to_rgb8`, `to_luma8`) and primitive/impls/environment_light.rs:36-37 (`to_rgb32f`).  libtcpt's tcpt_image_convert (csrc/host_image.h) is
checked against the rules written out again in numpy below, on raw arrays and on PNG files decoded from disk.  CPU-only tests."""
import numpy as np
import pytest

from toy_cpu_pathtracing_b200.scene import FloatTexture, RgbTexture, convert_image, decode_image


# ---- the rules, restated (image 0.25.6 color.rs / traits.rs): oracle side of this test
def depth_to_u8(a):
    if a.dtype == np.uint8:
        return a
    if a.dtype == np.uint16:
        return ((a.astype(np.uint32) + 128) // 257).astype(np.uint8)
    v = np.clip(a.astype(np.float32), np.float32(0), np.float32(1)) * np.float32(255)
    return np.floor(v + np.float32(0.5)).astype(np.uint8)          # f32::round = half away from zero; v >= 0 here


def depth_to_f32(a):
    if a.dtype == np.uint8:
        return a.astype(np.float32) / np.float32(255)
    if a.dtype == np.uint16:
        return a.astype(np.float32) / np.float32(65535)
    return a.astype(np.float32)


def luma_in_source_depth(rgb):
    if rgb.dtype == np.float32:
        r, g, b = (rgb[..., k].astype(np.float64) for k in range(3))
        return ((2126.0 * r + 7152.0 * g + 722.0 * b) / 10000.0).astype(np.float32)
    r, g, b = (rgb[..., k].astype(np.uint64) for k in range(3))
    return ((2126 * r + 7152 * g + 722 * b) // 10000).astype(rgb.dtype)


def expected(a, kind):
    ch = 1 if a.ndim == 2 else a.shape[2]
    colour = ch >= 3
    if kind == "luma8":
        return depth_to_u8(luma_in_source_depth(a[..., :3]) if colour else (a if a.ndim == 2 else a[..., 0]))
    conv = depth_to_u8 if kind == "rgb8" else depth_to_f32
    if colour:
        return conv(a[..., :3])
    g = conv(a if a.ndim == 2 else a[..., 0])
    return np.stack([g, g, g], -1)


def random_image(rng, ch, dtype, h=37, w=23):
    shape = (h, w) if ch == 1 else (h, w, ch)
    if dtype == np.float32:
        a = rng.normal(0.5, 0.6, size=shape).astype(np.float32)     # values below 0 and above 1 included
        a.flat[:4] = [0.5 / 255, 1.5 / 255, 254.5 / 255, 2.5 / 255]  # .5 cases of the rounding
        return a
    a = rng.integers(0, np.iinfo(dtype).max + 1, size=shape).astype(dtype)
    a.flat[:6] = np.array([0, 1, np.iinfo(dtype).max, 128, 129, 385 % (np.iinfo(dtype).max + 1)], dtype=dtype)
    return a


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("ch", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["rgb8", "luma8", "rgb32f"])
def test_conversions_follow_the_image_crate_rules(ch, dtype, kind):
    a = random_image(np.random.default_rng(ch * 10 + np.dtype(dtype).itemsize), ch, dtype)
    got, want = convert_image(a, kind), expected(a, kind)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got, want)


def test_known_values():
    # u16 -> u8 rounds to nearest (65535 / 257 = 255 exactly): 128 -> 0 (256 / 257), 129 -> 1, 65535 -> 255, 32896 -> 128
    a = np.array([[128, 129, 65535, 32896, 32767]], np.uint16)
    assert convert_image(a, "luma8").tolist() == [[0, 1, 255, 128, 127]]
    # integer luma: weights 2126 / 7152 / 722 over 10000, truncated -- pure green 255 is 182, not round(0.7152 * 255) = 182.4 -> 182; pure red 54; pure blue 18
    px = np.array([[[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255], [1, 1, 1], [10, 200, 30]]], np.uint8)
    assert convert_image(px, "luma8").tolist() == [[54, 182, 18, 255, 1, (2126 * 10 + 7152 * 200 + 722 * 30) // 10000]]
    # alpha is dropped, never pre-multiplied
    rgba = np.array([[[200, 100, 50, 0], [200, 100, 50, 255]]], np.uint8)
    assert convert_image(rgba, "rgb8").tolist() == [[[200, 100, 50], [200, 100, 50]]]
    # 16-bit colour to luma8: the luma is formed at 16 bits and reduced afterwards
    c16 = np.array([[[65535, 0, 0], [300, 40000, 123]]], np.uint16)
    l16 = [(2126 * 65535) // 10000, (2126 * 300 + 7152 * 40000 + 722 * 123) // 10000]
    assert convert_image(c16, "luma8").tolist() == [[(v + 128) // 257 for v in l16]]
    assert np.array_equal(convert_image(np.array([[0, 255]], np.uint8), "rgb32f"), np.array([[[0, 0, 0], [1, 1, 1]]], np.float32))


def test_png_files_through_the_texture_loaders(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    rgb8, rgb16, g8, g16, rgba8 = (random_image(rng, c, d, 16, 20) for c, d in ((3, np.uint8), (3, np.uint16), (1, np.uint8), (1, np.uint16), (4, np.uint8)))
    files = {"rgb8": rgb8, "rgb16": rgb16, "g8": g8, "g16": g16, "rgba8": rgba8}
    for name, a in files.items():
        bgr = a if a.ndim == 2 else a[..., [2, 1, 0] + ([3] if a.shape[2] == 4 else [])]
        assert cv2.imwrite(str(tmp_path / f"{name}.png"), bgr)
    for name, a in files.items():
        dec = decode_image(tmp_path / f"{name}.png")
        assert dec.dtype == a.dtype and np.array_equal(dec, a), name            # the decoder hands back the file's own depth and channels
        assert np.array_equal(RgbTexture.load_srgb(tmp_path / f"{name}.png").data, expected(a, "rgb8")), name
        assert np.array_equal(FloatTexture.load(tmp_path / f"{name}.png").data, expected(a, "luma8")), name
    assert RgbTexture.load_srgb(tmp_path / "g8.png").data.shape == (16, 20, 3)
    assert FloatTexture.load(tmp_path / "rgb16.png").data.shape == (16, 20)
