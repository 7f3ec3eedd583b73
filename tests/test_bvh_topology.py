"""BVH topology must be bit-exact (BASELINE.json north_star): the O(N log N)-per-level host builder of libtcpt against the
oracle's LITERAL restatement of the reference's O(N^2) full-sweep SAH builder (scene/src/bvh.rs:92-295), node by node in the
reference's own flattened order: kind, second_offset / item_count / item, and the bit patterns of every box."""
import ctypes as C

import numpy as np
import pytest


def host_build_boxes(boxes):
    from toy_cpu_pathtracing_b200 import capi
    lib = capi.load_library()
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    n = lib.tcpt_build_bvh_boxes(capi.as_ptr(b, C.c_float), len(b), None, 0)
    out = np.zeros((n, 8), dtype=np.uint32)
    lib.tcpt_build_bvh_boxes(capi.as_ptr(b, C.c_float), len(b), capi.as_ptr(out, C.c_uint32), n)
    return out


def random_boxes(rng, n, kind):
    if kind == "uniform":
        c = rng.uniform(-1, 1, (n, 3)); e = rng.uniform(0.001, 0.2, (n, 3))
    elif kind == "clustered":
        c = rng.normal(0, 0.05, (n, 3)) + rng.integers(-2, 3, (n, 1)); e = rng.uniform(0.0, 0.05, (n, 3))
    elif kind == "ties":          # many identical centres on every axis: exercises the stable sort and the strict '<' tie rules
        c = rng.integers(-2, 3, (n, 3)).astype(np.float64) * 0.5; e = np.full((n, 3), 0.25)
    elif kind == "flat":          # zero extent on one axis (axis-aligned quads)
        c = rng.uniform(-1, 1, (n, 3)); e = rng.uniform(0.01, 0.2, (n, 3)); e[:, 1] = 0
    else:
        raise KeyError(kind)
    return np.concatenate([c - e, c + e], 1).astype(np.float32)


@pytest.mark.parametrize("kind", ["uniform", "clustered", "ties", "flat"])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 64, 301])
def test_builders_agree_on_random_boxes(built, kind, n):
    from oracle import oracle
    rng = np.random.default_rng(n * 7 + len(kind))
    boxes = random_boxes(rng, n, kind)
    literal = oracle.build_bvh_boxes(boxes, literal=True)
    fast = oracle.build_bvh_boxes(boxes, literal=False)
    host = host_build_boxes(boxes)
    assert np.array_equal(literal, fast), "oracle prefix/suffix sweep differs from the literal O(N^2) sweep"
    assert np.array_equal(literal, host), "libtcpt host builder differs from the reference-literal builder"
    # structural sanity of the reference layout: every item appears exactly once
    items = literal[literal[:, 0] == 2, 1]
    assert sorted(items.tolist()) == list(range(n))


def test_single_item_is_a_leaf(built):
    from oracle import oracle
    b = np.array([[0, 0, 0, 1, 1, 1]], dtype=np.float32)
    ref = oracle.build_bvh_boxes(b, literal=True)
    assert ref[:, 0].tolist() == [1, 2] and ref[0, 1] == 1
    assert np.array_equal(ref, host_build_boxes(b))


@pytest.mark.parametrize("scene_id", [1, 3, 7, 17, 19])
def test_scene_bvhs_match_literal_reference_build(bundle_factory, tables, scene_id):
    """TLAS (incl. the rotate-scale-translate instance of scene 17, the three instances of scene 19, the four scaled instances of scene 7
    and the vertex-tight box of scene 1's SingleTriangle next to its rotated mesh instance) and every BLAS."""
    from oracle import oracle
    b = bundle_factory(scene_id, 64, 48, require_gpu=False)
    lit = oracle.scene_from_description(b.scene.desc, b.camera.position, tables[0], tables[1], literal_build=len(b.scene.desc.meshes[0].indices) <= 6000)
    n_geom = len(b.scene.desc.meshes)
    for which in [-1] + list(range(n_geom)):
        if which >= 0 and len(b.scene.desc.meshes[which].indices) > 6000:
            ref = b.oracle.get_bvh(which)      # big meshes: the oracle's fast sweep (proved equal to the literal one above)
        else:
            ref = lit.get_bvh(which) if which < 0 or len(b.scene.desc.meshes[which].indices) <= 6000 else b.oracle.get_bvh(which)
        got = b.scene.get_bvh(which)
        assert got.shape == ref.shape and np.array_equal(got, ref), f"scene {scene_id} bvh {which}"


def test_single_triangle_tangent_and_delta_light_errors(bundle_factory):
    """SingleTriangle geometry keeps the reference's unfallbacked per-hit tangent (single_triangle.rs:118-124); delta lights validate
    their arguments with codes."""
    import ctypes as C
    from toy_cpu_pathtracing_b200 import capi
    b = bundle_factory(1, 64, 48, require_gpu=False)
    single = [i for i, m in enumerate(b.scene.desc.meshes) if m.single]
    assert len(single) == 1
    t_h, t_o = b.scene.mesh_tangents(single[0]), b.oracle.mesh_tangents(single[0])
    assert t_h.shape == (1, 3) and np.array_equal(t_h.view(np.uint32), t_o.view(np.uint32)) and np.allclose(t_h[0], [1, 0, 0])
    ctx = b.scene.ctx
    spec = capi.SpectrumParam(capi.SPEC_D65, (C.c_float * 3)(0, 0, 0), -1)
    eye = np.eye(4, dtype=np.float32)
    assert ctx.lib.tcpt_scene_add_delta_light(ctx.handle, 9, 1.0, C.byref(spec), 0.0, 0.0, capi.as_ptr(eye, C.c_float)) == capi.TCPT_ERR_INVALID
    spec.kind = capi.SPEC_TEXTURE_SRGB
    assert ctx.lib.tcpt_scene_add_delta_light(ctx.handle, capi.LIGHT_POINT, 1.0, C.byref(spec), 0.0, 0.0, capi.as_ptr(eye, C.c_float)) == capi.TCPT_ERR_INVALID
    assert ctx.lib.tcpt_scene_add_single_triangle(ctx.handle, None, None, None) == capi.TCPT_ERR_INVALID


def test_load_time_tangents_and_table_indices_match(bundle_factory):
    b = bundle_factory(3, 64, 48, require_gpu=False)
    assert np.array_equal(b.scene.mesh_tangents(0).view(np.uint32), b.oracle.mesh_tangents(0).view(np.uint32))
    rng = np.random.default_rng(5)
    cols = np.concatenate([rng.uniform(0, 1, (200, 3)), [[0.8, 0.8, 0.8], [0, 0, 0], [1, 1, 1], [1, 0, 0], [0, 1, 0.999], [0.7, 0.8, 1.0], [0.4, 0.9, 1.0]]]).astype(np.float32)
    for gamma in (True, False):
        for c in cols:
            cs_h, ix_h = b.scene.rgb_to_coeffs(c, gamma)
            cs_o, ix_o = b.oracle.rgb_to_coeffs(c, gamma)
            assert ix_h.tolist() == ix_o.tolist()
            assert cs_h.view(np.uint32).tolist() == cs_o.view(np.uint32).tolist()
