#!/usr/bin/env python3
"""Extract the STANDARD numeric tables the hot path reads from the reference checkout.

These are published public data sets (pbrt-v4 Sobol generator matrices, CIE 1931 2-degree colour
matching functions, CIE D65), not code authored by the reference.  They are written once into
`toy_cpu_pathtracing_b200/data/std_tables.bin` (committed) so nothing reads /root/reference at run time.

Sources (file:line in /root/reference):
  renderer/src/sampler/sobol_matrices.rs:7     SOBOL_MATRICES_32, first 2 x 52 words (only dims 0,1 are used:
                                               renderer/src/sampler/z_sobol_sampler.rs:208,221,225)
  spectrum/src/presets.rs:469,944,1419,1894    CIE_X / CIE_Y / CIE_Z / CIE_LAMBDA (471 entries, 360..830 nm)
  spectrum/src/presets.rs:2081                 CIE_ILLUM_D6500 interleaved (lambda, value)

The dense 470-entry tables are produced with the reference's own float32 arithmetic:
  DenselySampledSpectrum::from(PiecewiseLinearSpectrum)          spectrum/src/spectrum/densely_sampled_spectrum.rs:37-49
  PiecewiseLinearSpectrum::value                                 spectrum/src/spectrum/piecewise_linear_spectrum.rs:67-80
  PiecewiseLinearSpectrum::from_interleaved(normalized=true)     spectrum/src/spectrum/piecewise_linear_spectrum.rs:34-64
  inner_product                                                  spectrum/src/spectrum.rs:67-79

  spectrum/src/presets.rs:2365-2978            metal eta/k and glass eta tables (pbrt-v4 data), interleaved (lambda, value);
                                               densely resampled like CachedSpectrum::init does (presets.rs:129-200)

File layout (little endian):
  magic 'TCPTSTD2' | u32 sobol[104] | f32 cie_x[470] | f32 cie_y[470] | f32 cie_z[470] | f32 d65[470]
  | u32 n_presets | f32 preset[n_presets][470]      preset order = PRESETS below (include/tcpt.h TCPT_PRESET_*)
"""
import re
import struct
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "toy_cpu_pathtracing_b200" / "data" / "std_tables.bin"
f32 = np.float32


def rust_array(text: str, name: str):
    m = re.search(r"(?:const|static)\s+" + name + r"\s*:[^=]*=\s*&?\[(.*?)\];", text, re.S)
    assert m, name
    body = re.sub(r"//[^\n]*", "", m.group(1))
    toks = [t.strip().replace("_", "") for t in body.split(",") if t.strip()]
    return toks


def pwl_value(lams, vals, lam):
    """PiecewiseLinearSpectrum::value in float32."""
    if lam < lams[0] or lam > lams[-1]:
        return f32(0.0)
    i = 0
    while i < len(lams) - 1 and lams[i + 1] < lam:
        i += 1
    t = f32(f32(lam - lams[i]) / f32(lams[i + 1] - lams[i]))
    return f32(f32(vals[i] * f32(f32(1.0) - t)) + f32(vals[i + 1] * t))


def dense_from_pwl(lams, vals):
    return np.array([pwl_value(lams, vals, f32(360.0) + f32(i)) for i in range(470)], dtype=f32)


PRESETS = ["AU_ETA", "AU_K", "AG_ETA", "AG_K", "CU_ETA", "CU_K", "AL_ETA", "AL_K", "CU_ZN_ETA", "CU_ZN_K",
           "GLASS_BK7_ETA", "GLASS_BAF10_ETA", "GLASS_FK51A_ETA", "GLASS_LASF9_ETA", "GLASS_SF5_ETA", "GLASS_SF10_ETA", "GLASS_SF11_ETA"]


def main():
    sob = rust_array((REF / "renderer/src/sampler/sobol_matrices.rs").read_text(), "SOBOL_MATRICES_32")
    sobol = np.array([int(t, 16) for t in sob[:104]], dtype=np.uint32)

    presets = (REF / "spectrum/src/presets.rs").read_text()
    lam = np.array([f32(t) for t in rust_array(presets, "CIE_LAMBDA")], dtype=f32)
    assert len(lam) == 471
    cie = {}
    for k in ("CIE_X", "CIE_Y", "CIE_Z"):
        v = np.array([f32(t) for t in rust_array(presets, k)], dtype=f32)
        assert len(v) == 471
        cie[k] = dense_from_pwl(lam, v)

    d65_raw = np.array([f32(t) for t in rust_array(presets, "CIE_ILLUM_D6500")], dtype=f32)
    dl, dv = d65_raw[0::2].copy(), d65_raw[1::2].copy()
    # from_interleaved(normalized = true): y_self = inner_product(spec, Y); values / y_self
    y_self = f32(0.0)
    for i in range(470):
        l = f32(360.0) + f32(i)
        y_self = f32(y_self + f32(pwl_value(dl, dv, l) * cie["CIE_Y"][i]))
    d65 = np.array([f32(pwl_value(dl, dv, f32(360.0) + f32(i)) / y_self) for i in range(470)], dtype=f32)

    OUT.parent.mkdir(parents=True, exist_ok=True)
    presets_dense = []
    for name in PRESETS:
        raw = np.array([f32(t) for t in rust_array(presets, name)], dtype=f32)
        presets_dense.append(dense_from_pwl(raw[0::2].copy(), raw[1::2].copy()))

    with open(OUT, "wb") as fh:
        fh.write(b"TCPTSTD2")
        fh.write(sobol.astype("<u4").tobytes())
        for k in ("CIE_X", "CIE_Y", "CIE_Z"):
            fh.write(cie[k].astype("<f4").tobytes())
        fh.write(d65.astype("<f4").tobytes())
        fh.write(struct.pack("<I", len(presets_dense)))
        for t in presets_dense:
            fh.write(t.astype("<f4").tobytes())
    print("wrote", OUT, OUT.stat().st_size, "bytes; y_self =", y_self, "sum(Y) =", cie["CIE_Y"].sum())


if __name__ == "__main__":
    main()
