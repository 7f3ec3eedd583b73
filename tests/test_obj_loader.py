"""Scene::load_obj on real files: libtcpt's loader (csrc/host_obj.h, product) against hand-written OBJ files
whose expected arrays are written out by hand from tobj 4.0.3's `single_index + triangulate` rules, and against the oracle's pure-Python
restatement (oracle/obj_oracle.py) on generated files.  Reference: scene/src/geometry/impls/triangle_mesh.rs:141-243.  Triangle and vertex
ORDER is what is tested: it feeds the stable sorts of the SAH builder, so BVH-topology parity on an asset depends on it.  CPU-only tests."""
import ctypes as C

import numpy as np
import pytest

import toy_cpu_pathtracing_b200 as tp
from oracle import obj_oracle
from toy_cpu_pathtracing_b200 import capi
from toy_cpu_pathtracing_b200.scene import load_obj


def raw_load(path):
    """The loader's arrays before the Scene-level consistency checks."""
    lib = capi.load_library()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.tcpt_obj_load(str(path).encode(), C.byref(h), err, len(err))
    if rc != 0:
        raise ValueError(err.value.decode())
    counts = (C.c_uint32 * 5)()
    lib.tcpt_obj_counts(h, counts)
    nv, nn, nt, ntri, nm = (int(c) for c in counts)
    pos, nrm, uvs = np.zeros((nv, 3), np.float32), np.zeros((nn, 3), np.float32), np.zeros((nt, 2), np.float32)
    idx, ttri = np.zeros((ntri, 3), np.uint32), np.zeros(ntri, np.uint32)
    lib.tcpt_obj_copy(h, capi.as_ptr(pos, C.c_float), capi.as_ptr(nrm, C.c_float), capi.as_ptr(uvs, C.c_float), capi.as_ptr(idx, C.c_uint32), capi.as_ptr(ttri, C.c_uint32))
    lib.tcpt_obj_free(h)
    return pos, nrm, uvs, idx, ttri, nm


def write(tmp_path, text, name="m.obj"):
    p = tmp_path / name
    p.write_text(text)
    return p


def same_as_oracle(path):
    got, want = raw_load(path), obj_oracle.load_obj(path)
    for g, w, what in zip(got[:5], want[:5], ("positions", "normals", "uvs", "indices", "tangent_tri")):
        assert g.shape == w.shape and np.array_equal(g.view(np.uint32) if g.dtype == np.float32 else g, w.view(np.uint32) if w.dtype == np.float32 else w), what
    assert got[5] == want[5]
    return got


QUAD_SHARED = """# two triangles sharing an edge, every vertex referenced as v/vt/vn
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
f 1/1/1 2/2/1 3/3/1
f 1/1/1 3/3/1 4/4/1
"""


def test_shared_triples_are_one_vertex_in_first_use_order(tmp_path):
    pos, nrm, uvs, idx, ttri, nm = same_as_oracle(write(tmp_path, QUAD_SHARED))
    assert idx.tolist() == [[0, 1, 2], [0, 2, 3]]
    assert pos.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]]
    assert uvs.tolist() == [[0, 0], [1, 0], [1, 1], [0, 1]] and nrm.tolist() == [[0, 0, 1]] * 4
    assert ttri.tolist() == [0, 1] and nm == 1


def test_first_use_order_is_not_file_order(tmp_path):
    text = "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 5 5 5\nvn 0 0 1\nf 3//1 1//1 2//1\n"      # v 4 is never used; vertex 3 comes first
    pos, nrm, uvs, idx, _, _ = same_as_oracle(write(tmp_path, text))
    assert pos.tolist() == [[0, 1, 0], [0, 0, 0], [1, 0, 0]] and idx.tolist() == [[0, 1, 2]] and len(uvs) == 0 and len(nrm) == 3


def test_same_position_with_another_texcoord_or_normal_is_another_vertex(tmp_path):
    text = ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvt 0.5 0.5\nvn 0 0 1\nvn 0 1 0\n"
            "f 1/1/1 2/2/1 3/3/1\n"
            "f 1/4/1 2/2/1 3/3/2\n")       # v1 with vt 4 -> new vertex; v2/vt2/vn1 shared; v3 with vn 2 -> new vertex
    pos, nrm, uvs, idx, _, _ = same_as_oracle(write(tmp_path, text))
    assert idx.tolist() == [[0, 1, 2], [3, 1, 4]]
    assert pos.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 0], [0, 1, 0]]
    assert uvs[3].tolist() == [0.5, 0.5] and nrm[4].tolist() == [0, 1, 0]


def test_quads_and_polygons_are_fans_in_face_order(tmp_path):
    text = ("v 0 0 0\nv 1 0 0\nv 2 1 0\nv 1 2 0\nv 0 1 0\nv 9 9 9\nvn 0 0 1\n"
            "f 1//1 2//1 3//1 4//1 5//1\n"       # pentagon: (a,b,c) (a,c,d) (a,d,e)
            "f 6//1 1//1 2//1 3//1\n")           # quad: (a,b,c) (a,c,d)
    _, _, _, idx, _, _ = same_as_oracle(write(tmp_path, text))
    assert idx.tolist() == [[0, 1, 2], [0, 2, 3], [0, 3, 4], [5, 0, 1], [5, 1, 2]]


def test_negative_indices_count_back_from_what_has_been_read_so_far(tmp_path):
    text = ("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf -3//-1 -2//-1 -1//-1\n"
            "v 0 0 1\nv 1 0 1\nv 0 1 1\nvn 0 1 0\nf -3//-1 -2//-1 -1//-1\n")
    pos, nrm, _, idx, _, _ = same_as_oracle(write(tmp_path, text))
    assert idx.tolist() == [[0, 1, 2], [3, 4, 5]] and pos[3].tolist() == [0, 0, 1] and nrm[5].tolist() == [0, 1, 0]


def test_points_lines_comments_and_unknown_statements(tmp_path):
    text = ("#comment without a blank\n# comment\n\ns off\nv 0 0 0 1 0 0\nv 1 0 0\nv 0 1 0\nvt 0.25 0.75 0\nvn 0 0 1\n"
            "f 1//1\nf 1//1 2//1\nl 1 2\n"          # a point and two lines: dropped
            "l 1/1/1 2/1/1 3/1/1\n"                 # an `l` statement with three vertices goes through the face path (tobj)
            "f 1/1/1 2/1/1 3/1/1\r\n")              # CRLF line ending
    pos, _, uvs, idx, _, _ = same_as_oracle(write(tmp_path, text))
    assert idx.tolist() == [[0, 1, 2], [0, 1, 2]] and len(pos) == 3 and uvs.tolist() == [[0.25, 0.75]] * 3


def test_models_are_concatenated_without_a_vertex_offset(tmp_path):
    """Two `o` groups: per-model vertex numbering, and the reference's `indices.extend(mesh.indices)` adds no offset (triangle_mesh.rs:177),
    so the second model's triangles address the first model's vertices; the tangent loop re-runs over all triangles after each model."""
    text = ("o first\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvt 1 1\nvn 0 0 1\n"
            "f 1/1/1 2/2/1 3/3/1\nf 2/2/1 4/4/1 3/3/1\n"
            "o second\nv 0 0 1\nv 1 0 1\nv 0 1 1\n"
            "f 5/1/1 6/2/1 7/3/1\n")
    pos, nrm, uvs, idx, ttri, nm = same_as_oracle(write(tmp_path, text))
    assert nm == 2 and len(pos) == 7
    assert idx.tolist() == [[0, 1, 2], [1, 3, 2], [0, 1, 2]]          # third triangle: local numbering of model 2, not 4 5 6
    assert ttri.tolist() == [0, 1, 0]                                  # pushes: [0, 1] after model 1, [0, 1, 2] after model 2 -> tangents[2] = T(0)
    mesh = load_obj(write(tmp_path, text, "again.obj"))
    assert mesh.tangent_tri is not None and mesh.tangent_tri.tolist() == [0, 1, 0]


def test_group_statement_before_any_face_opens_no_model(tmp_path):
    text = "g a\ng b\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\ng trailing\n"
    *_, idx, ttri, nm = same_as_oracle(write(tmp_path, text))
    assert idx.tolist() == [[0, 1, 2]] and nm == 2        # the model closed by `g trailing`, then the (empty) one closed at end of file


def test_usemtl_splits_only_for_materials_a_loadable_mtl_defines(tmp_path):
    body = ("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\n"
            "usemtl red\nf 1//1 2//1 3//1\nusemtl blue\nf 2//1 4//1 3//1\n")
    # no mtllib (or an unreadable one): every usemtl maps to "no material", nothing changes, one model, shared vertices
    *_, idx, _, nm = same_as_oracle(write(tmp_path, "mtllib missing.mtl\n" + body, "a.obj"))
    assert nm == 1 and idx.tolist() == [[0, 1, 2], [1, 3, 2]]
    # with a material library that defines both: the change of material closes the model; vertices are numbered per model
    (tmp_path / "two.mtl").write_text("newmtl red\nKd 1 0 0\nnewmtl blue\nKd 0 0 1\n")
    *_, idx, _, nm = same_as_oracle(write(tmp_path, "mtllib two.mtl\n" + body, "b.obj"))
    assert nm == 2 and idx.tolist() == [[0, 1, 2], [0, 1, 2]]


@pytest.mark.parametrize("tok,bits", [("0.1", 0x3DCCCCCD), ("-0", 0x80000000), ("1e-3", 0x3A83126F), ("+.5", 0x3F000000), ("1.", 0x3F800000),
                                      ("16777217", 0x4B800000), ("3.4028235e38", 0x7F7FFFFF), ("3.4028236e38", 0x7F800000), ("1e39", 0x7F800000), ("7.038531e-26", 0x15AE43FD),
                                      ("1.00000017881393432617187500001", 0x3F800002)])
def test_decimal_tokens_are_rounded_to_binary32_once(tmp_path, tok, bits):
    text = f"v {tok} 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n"
    pos, *_ = same_as_oracle(write(tmp_path, text))
    assert int(pos[0, 0].view(np.uint32)) == bits


@pytest.mark.parametrize("text", ["v 0 0\nf 1 1 1\n", "v 0 0 0\nvt 1\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 x\n", "v 0x10 0 0\n",
                                  "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/1/1/1 2 3\n"])
def test_malformed_files_are_refused_by_both(tmp_path, text):
    p = write(tmp_path, text)
    with pytest.raises(ValueError):
        raw_load(p)
    with pytest.raises((ValueError, IndexError)):
        obj_oracle.load_obj(p)


def test_scene_level_rules(tmp_path):
    # the reference indexes normals[i] for every vertex: a file without vn cannot be shaded (triangle_mesh.rs:57-61 panics)
    with pytest.raises(ValueError, match="vn"):
        load_obj(write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "nonormal.obj"))
    with pytest.raises(ValueError, match="texcoords"):
        load_obj(write(tmp_path, "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvn 0 0 1\nf 1/1/1 2//1 3//1\n", "partuv.obj"))
    ctx = capi.Context(0, require_gpu=False)
    assert ctx.lib.tcpt_scene_load_obj(ctx.handle, str(tmp_path / "nonormal.obj").encode()) == capi.TCPT_ERR_INVALID and "vn" in ctx.last_error()
    assert ctx.lib.tcpt_scene_load_obj(ctx.handle, str(tmp_path / "nope.obj").encode()) == capi.TCPT_ERR_INVALID


def random_obj(rng, n_v=60, n_f=80, groups=False):
    lines = []
    for _ in range(n_v):
        lines.append("v " + " ".join(f"{x:.6g}" for x in rng.normal(size=3)))
    n_t, n_n = 25, 12
    for _ in range(n_t):
        lines.append("vt " + " ".join(f"{x:.5f}" for x in rng.random(2)))
    for _ in range(n_n):
        lines.append("vn " + " ".join(f"{x:.4f}" for x in rng.normal(size=3)))
    for k in range(n_f):
        if groups and k and k % 23 == 0:
            lines.append(f"g part{k}")
        arity = int(rng.choice([3, 3, 3, 4, 4, 5, 6, 2, 1]))
        verts = []
        for v in rng.choice(n_v, size=arity, replace=False):
            t, n = int(rng.integers(0, 6)), int(rng.integers(0, n_n))        # few texcoords / normals per position: many repeated triples
            vi = int(v) + 1 if rng.random() < 0.7 else int(v) - n_v
            verts.append(f"{vi}/{(v + t) % n_t + 1}/{n + 1}")
        lines.append("f " + " ".join(verts))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed,groups", [(0, False), (1, False), (2, True), (3, True)])
def test_generated_files_match_the_oracle_and_build_the_same_bvh(tmp_path, seed, groups):
    rng = np.random.default_rng(seed)
    path = write(tmp_path, random_obj(rng, groups=groups))
    pos, nrm, uvs, idx, ttri, nm = same_as_oracle(path)
    assert (nm > 1) == groups and len(idx) > 80
    # Scene::load_obj -> the same BLAS, node for node, as handing the oracle's arrays to add_mesh; tangents incl. the multi-model rule
    from oracle import oracle
    ctx = capi.Context(0, require_gpu=False)
    g = ctx.lib.tcpt_scene_load_obj(ctx.handle, str(path).encode())
    assert g == 0, ctx.last_error()
    scene = tp.Scene(context=ctx)
    std, tab = capi.load_tables()
    osc = oracle.OracleScene(std, tab, faithful=True, literal_build=False)
    o_pos, o_nrm, o_uv, o_idx, o_ttri, _ = obj_oracle.load_obj(path)
    og = osc.add_mesh(o_pos, o_nrm, o_uv, o_idx)
    osc.set_tangent_source(og, o_ttri)
    from toy_cpu_pathtracing_b200.scene import SceneDescription
    from toy_cpu_pathtracing_b200.scenes import _lambert
    mat = SceneDescription().material_desc(_lambert(0.5, 0.5, 0.5))
    eye = np.eye(4, dtype=np.float32).T.copy().ravel()
    ctx.check(ctx.lib.tcpt_scene_add_material(ctx.handle, C.byref(mat)))
    ctx.check(ctx.lib.tcpt_scene_add_primitive(ctx.handle, 0, 0, capi.as_ptr(eye, C.c_float)))
    ctx.check(ctx.lib.tcpt_scene_build(ctx.handle, capi.as_ptr(np.zeros(3, np.float32), C.c_float)), allow_no_gpu=True)
    osc.add_material(mat); osc.add_primitive(0, 0, eye); osc.build(np.zeros(3, np.float32))
    assert np.array_equal(scene.get_bvh(0), osc.get_bvh(0))
    gt, ot = scene.mesh_tangents(0), osc.mesh_tangents(0)
    assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32))
