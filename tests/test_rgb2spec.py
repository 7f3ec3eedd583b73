"""The rgb -> spectrum path end to end, against colour science instead of against the oracle: look a colour up in the coefficient
table (libtcpt's host lookup, spectrum/src/rgb_sigmoid_polynomial.rs:87-155), evaluate the sigmoid polynomial, integrate it against
the CIE observer under D65 and convert back to sRGB -- the procedure of the reference's rgb_to_spec/tests/test.rs:156-183.  A wrong
table index, a wrong axis order or a bad refit of the (LFS-absent) table shows up as a colour error."""
import struct

import numpy as np
import pytest


def std_tables(tables):
    std = tables[0]
    f = np.frombuffer(std[8 + 104 * 4: 8 + 104 * 4 + 4 * 470 * 4], dtype="<f4").reshape(4, 470)
    return f[0].astype(np.float64), f[1].astype(np.float64), f[2].astype(np.float64), f[3].astype(np.float64)


XYZ_TO_SRGB = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]])


def round_trip(scene, tables, rgb_linear):
    cx, cy, cz, d65 = std_tables(tables)
    cs, _ = scene.rgb_to_coeffs(np.asarray(rgb_linear, dtype=np.float32), gamma_encoded=False)
    t = np.arange(470) / 470.0
    s = 1.0 / (1.0 + np.exp(-(cs[0] * t * t + cs[1] * t + cs[2])))
    xyz = np.array([(s * cx * d65).sum(), (s * cy * d65).sum(), (s * cz * d65).sum()])
    return XYZ_TO_SRGB @ xyz


def test_white_point_and_normalisation(bundle_factory, tables):
    cx, cy, cz, d65 = std_tables(tables)
    assert abs((cy * d65).sum() - 1.0) < 1e-4                      # D65 is normalised to Y = 1 (presets.rs)
    white = XYZ_TO_SRGB @ np.array([(cx * d65).sum(), (cy * d65).sum(), (cz * d65).sum()])
    assert np.allclose(white, [1, 1, 1], atol=5e-3)                 # a unit reflector under D65 is sRGB white


def test_colours_survive_the_spectral_round_trip(bundle_factory, tables):
    b = bundle_factory(3, 24, 18, require_gpu=False)
    rng = np.random.default_rng(3)
    cols = np.concatenate([rng.uniform(0.02, 0.98, (300, 3)),
                           [[0.8, 0.8, 0.8], [0.9, 0.05, 0.05], [0.05, 0.9, 0.05], [0.05, 0.05, 0.9], [0.5, 0.5, 0.5], [0.2, 0.6, 0.9], [0.95, 0.9, 0.1]]])
    err = np.array([np.abs(round_trip(b.scene, tables, c) - c).max() for c in cols])
    # smooth three-parameter spectra reproduce in-gamut colours to well under a percent; saturated corners are the worst case
    assert np.median(err) < 3e-3 and np.quantile(err, 0.95) < 1.5e-2 and err.max() < 4e-2, (np.median(err), err.max())


def test_grey_axis_is_flat(bundle_factory, tables):
    """r = g = b takes the reference's closed-form branch (rgb_sigmoid_polynomial.rs:95-108): a constant spectrum of that value."""
    b = bundle_factory(3, 24, 18, require_gpu=False)
    for v in (0.1, 0.5, 0.8):
        cs, _ = b.scene.rgb_to_coeffs(np.array([v, v, v], dtype=np.float32), gamma_encoded=False)
        assert cs[0] == 0 and cs[1] == 0 and abs(1 / (1 + np.exp(-cs[2])) - v) < 1e-6
        assert np.allclose(round_trip(b.scene, tables, [v, v, v]), [v, v, v], atol=5e-3)
