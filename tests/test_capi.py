"""The C-ABI library loads, exports every symbol include/*.h declares, and fails loudly (never falls back) without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    names = set()
    for h in (ROOT / "include").glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names |= set(re.findall(r"\b(tcpt_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(built):
    from toy_cpu_pathtracing_b200 import capi
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libtcpt.so does not export {s}"
    assert sorted(capi.EXPORTED_SYMBOLS) == syms, "capi.EXPORTED_SYMBOLS is out of sync with include/*.h"


def test_static_library_is_built(built):
    from toy_cpu_pathtracing_b200 import capi
    assert (capi.LIB_PATH.parent / "libtcpt.a").exists()


def test_struct_layouts_match_the_header(built):
    from toy_cpu_pathtracing_b200 import capi
    assert C.sizeof(capi.SpectrumParam) == 20 and C.sizeof(capi.FloatParam) == 16 and C.sizeof(capi.NormalParam) == 8
    assert C.sizeof(capi.MaterialDesc) == 4 + 20 + 16 + 8 + 4 + 4 + 16 * 6 + 20
    assert C.sizeof(capi.RenderParams) == 4 * 9 + 36 + 4 * 5


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="GPU present")
def test_no_cpu_fallback_without_gpu(built):
    """On a box without a GPU the context reports TCPT_ERR_CUDA and every computing entry point refuses to run."""
    from toy_cpu_pathtracing_b200 import capi
    with pytest.raises(capi.TcptError) as e:
        capi.Context(0, require_gpu=True)
    assert e.value.code == capi.TCPT_ERR_CUDA
    ctx = capi.Context(0, require_gpu=False)
    assert not ctx.has_gpu
    p = capi.RenderParams()
    p.width, p.height, p.spp = 4, 4, 1
    out = np.zeros(48, dtype=np.float32)
    rc = ctx.lib.tcpt_render(ctx.handle, C.byref(p), capi.as_ptr(out, C.c_float), None)
    assert rc == capi.TCPT_ERR_CUDA and not out.any()
    rays = np.zeros((1, 7), dtype=np.float32); hits = np.zeros((1, 6), dtype=np.int32)
    assert ctx.lib.tcpt_trace(ctx.handle, capi.as_ptr(rays, C.c_float), 1, 0, capi.as_ptr(hits, C.c_int32)) == capi.TCPT_ERR_CUDA


def test_argument_errors_are_codes_not_crashes(built):
    from toy_cpu_pathtracing_b200 import capi
    ctx = capi.Context(0, require_gpu=False)
    lib, h = ctx.lib, ctx.handle
    assert lib.tcpt_scene_add_mesh(h, None, None, None, 0, None, 0) == capi.TCPT_ERR_INVALID
    pos = np.zeros((3, 3), dtype=np.float32); idx = np.array([[0, 1, 7]], dtype=np.uint32)
    assert lib.tcpt_scene_add_mesh(h, capi.as_ptr(pos, C.c_float), capi.as_ptr(pos, C.c_float), None, 3, capi.as_ptr(idx, C.c_uint32), 1) == capi.TCPT_ERR_INVALID
    assert b"index out of range" in lib.tcpt_last_error(h)
    assert lib.tcpt_scene_add_primitive(h, 5, 5, capi.as_ptr(np.eye(4, dtype=np.float32), C.c_float)) == capi.TCPT_ERR_INVALID
    assert lib.tcpt_set_option(h, b"no_such_option", 1) == capi.TCPT_ERR_INVALID
    cam = np.zeros(3, dtype=np.float32)
    lib.tcpt_scene_clear(h)
    assert lib.tcpt_scene_build(h, capi.as_ptr(cam, C.c_float)) == capi.TCPT_ERR_INVALID  # empty scene


def test_rust_binding_source_is_in_sync_with_the_header(built):
    """integration/rust/gpu_backend/src/ffi.rs cannot be compiled here (no Rust toolchain): at least keep its `extern "C"` block and
    `#[repr(C)]` structs field-for-field in step with include/tcpt.h, through the ctypes structs the tests above tie to the header."""
    from toy_cpu_pathtracing_b200 import capi
    text = (ROOT / "integration" / "rust" / "gpu_backend" / "src" / "ffi.rs").read_text()
    fns = set(re.findall(r"pub fn (tcpt_[a-z0-9_]+)\(", text))
    declared = set(declared_symbols())
    assert fns <= declared, fns - declared
    product = {s for s in declared if not re.search(r"trace|sampler_stream|path_samples|get_bvh|get_wide_bvh|build_bvh|rgb_to_coeffs|mesh_tangents|upload_flat|cdf_search", s)}
    assert product <= fns, f"ffi.rs misses {product - fns}"        # everything but the test / introspection probes is bound

    def rust_fields(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, text, flags=re.S).group(1)
        return [(f, t) for f, t in re.findall(r"pub (\w+): ([^,]+),", body)]

    ctype_of = {"i32": C.c_int32, "u32": C.c_uint32, "f32": C.c_float, "u64": C.c_uint64, "f64": C.c_double, "[f32; 3]": C.c_float * 3,
                "TcptSpectrumParam": capi.SpectrumParam, "TcptFloatParam": capi.FloatParam, "TcptNormalParam": capi.NormalParam}
    for rust, py in (("TcptSpectrumParam", capi.SpectrumParam), ("TcptFloatParam", capi.FloatParam), ("TcptNormalParam", capi.NormalParam),
                     ("TcptMaterialDesc", capi.MaterialDesc), ("TcptRenderParams", capi.RenderParams), ("TcptStats", capi.Stats)):
        rf = rust_fields(rust)
        assert [f if f != "ty" else "type" for f, _ in rf] == [f for f, _ in py._fields_], rust
        assert [ctype_of[t.strip()] for _, t in rf] == [t for _, t in py._fields_], rust
