"""ORACLE (test infrastructure, not product code): pure-Python restatement of what TriangleMesh::load_obj holds after loading an OBJ file
(/root/reference/scene/src/geometry/impls/triangle_mesh.rs:141-243).

The parser behind it is the crates.io dependency tobj 4.0.3 (Cargo.lock), absent from the reference tree; its published algorithm is
restated here from its documented behaviour (`LoadOptions { single_index: true, triangulate: true, ignore_points: true, ignore_lines:
true }`): statements v / vt / vn / f / l / o / g / mtllib / usemtl, `v/vt/vn` vertex references with negative (relative) indices, one
output vertex per distinct (v, vt, vn) triple numbered in first-use order within a model, triangle / quad / fan triangulation, a model
boundary at every o / g (and at a usemtl that switches between materials a loadable .mtl defines) once faces have been read, and one
final model at end of file.  PARITY UNPINNED against tobj itself (no Rust toolchain here): tests/test_obj_loader.py pins the rules with
hand-written files whose expected arrays are written out by hand.

Everything is plain lists; decimal tokens are rounded to binary32 exactly once (Rust's f32::from_str is correctly rounded)."""
from __future__ import annotations

import os

import numpy as np

MISSING = None


def _f32(tok: str) -> float:
    """f32::from_str: correctly rounded decimal -> binary32 (a detour through binary64 would round twice)."""
    if any(c in tok for c in "xXpP_") or tok != tok.strip():
        raise ValueError(tok)
    d = float(tok)
    if d != d or d in (float("inf"), float("-inf")):
        return d
    from fractions import Fraction
    exact = Fraction(tok)
    fmax = np.finfo(np.float32).max
    near = np.float32(min(max(d, -float(fmax)), float(fmax)))
    best = None
    with np.errstate(over="ignore"):
        cands = (np.nextafter(near, np.float32(-np.inf)), near, np.nextafter(near, np.float32(np.inf)))
    for c in cands:
        if not np.isfinite(c):
            continue
        err = abs(Fraction(float(c)) - exact)
        even = (int(np.array(c, np.float32).view(np.uint32)) & 1) == 0
        if best is None or err < best[0] or (err == best[0] and even):
            best = (err, float(c))
    big = Fraction(float(np.finfo(np.float32).max)) + Fraction(2) ** 103       # halfway to 2^128: beyond it the result is infinite
    if abs(exact) >= big:
        return float("inf") if exact > 0 else float("-inf")
    return best[1]


def _floats(tokens, n):
    vals = [_f32(t) for t in tokens[:n]]
    if len(vals) != n:
        raise ValueError("too few numbers")
    return vals


def _vertex(tok, n_pos, n_tex, n_nrm):
    ref = [MISSING, MISSING, MISSING]
    for field, part in enumerate(tok.split("/")):
        if part == "":
            continue
        digits = part[1:] if part[0] in "+-" else part
        if field > 2 or not digits.isdigit() or not digits.isascii():
            raise ValueError(f"bad face vertex {tok!r}")
        x = int(part)
        size = (n_pos, n_tex, n_nrm)[field]
        ref[field] = size + x if x < 0 else (x - 1 if x > 0 else MISSING)
    return tuple(ref)


def _export(pos, tex, nrm, faces):
    """One tobj model: arrays in first-use order of the (v, vt, vn) triples of THIS model."""
    seen, out_pos, out_tex, out_nrm, out_idx = {}, [], [], [], []

    def add(v):
        if v in seen:
            out_idx.append(seen[v])
            return
        if v[0] is MISSING or not (0 <= v[0] < len(pos)):
            raise ValueError("face vertex out of bounds")
        out_pos.append(pos[v[0]])
        if tex and v[1] is not MISSING:
            out_tex.append(tex[v[1]])
        if nrm and v[2] is not MISSING:
            out_nrm.append(nrm[v[2]])
        seen[v] = len(seen)
        out_idx.append(seen[v])

    for f in faces:
        if len(f) in (1, 2):
            continue                                  # points and lines are ignored
        if not f:
            raise ValueError("invalid polygon")
        for k in range(1, len(f) - 1):                # (a, b, c), (a, c, d), ...: the fan; a quad gives (a,b,c)(a,c,d)
            add(f[0]); add(f[k]); add(f[k + 1])
    return out_pos, out_tex, out_nrm, out_idx


def _mtl_names(path):
    names = []
    try:
        for line in open(path):
            w = line.split()
            if w and w[0] == "newmtl":
                name = line.rstrip("\r\n")[6:].strip()
                if not name:
                    return None
                names.append(name)
    except OSError:
        return None
    return names


def load_obj(path):
    """(positions (V,3), normals (Vn,3), uvs (Vt,2), indices (T,3), tangent_tri (T,), n_models) as the reference holds them."""
    pos, tex, nrm, faces, models = [], [], [], [], []
    materials, current = {}, None
    for line in open(path):
        line = line.rstrip("\r\n")
        w = line.split()
        if not w or w[0] == "#":
            continue
        key = w[0]
        if key == "v":
            pos.append(tuple(_floats(w[1:], 3)))
        elif key == "vt":
            tex.append(tuple(_floats(w[1:], 2)))
        elif key == "vn":
            nrm.append(tuple(_floats(w[1:], 3)))
        elif key in ("f", "l"):
            faces.append([_vertex(t, len(pos), len(tex), len(nrm)) for t in w[1:]])
        elif key in ("o", "g"):
            if faces:
                models.append(_export(pos, tex, nrm, faces))
                faces = []
        elif key == "mtllib":
            name = line.split(" ", 1)[1].strip() if " " in line else ""
            got = _mtl_names(os.path.join(os.path.dirname(str(path)), name)) if name else None
            for n in got or []:
                materials.setdefault(n, len(materials))
        elif key == "usemtl":
            name = line[7:].strip()
            if name:
                new = materials.get(name)
                if new != current and faces:
                    models.append(_export(pos, tex, nrm, faces))
                    faces = []
                current = new
    models.append(_export(pos, tex, nrm, faces))

    # triangle_mesh.rs:163-226: extend, extend, extend, indices.extend (no vertex offset), then the tangent loop over ALL triangles so far
    P, T, N, I, pushed = [], [], [], [], []
    for mp, mt, mn, mi in models:
        P += mp; T += mt; N += mn; I += mi
        if T:
            pushed += list(range(len(I) // 3))
    n_tri = len(I) // 3
    tangent_tri = np.array(pushed[:n_tri] if T else list(range(n_tri)), dtype=np.uint32)
    f32 = np.float32
    return (np.array(P, dtype=f32).reshape(-1, 3), np.array(N, dtype=f32).reshape(-1, 3), np.array(T, dtype=f32).reshape(-1, 2),
            np.array(I, dtype=np.uint32).reshape(-1, 3), tangent_tri, len(models))
