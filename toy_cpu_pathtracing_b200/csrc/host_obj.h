// Wavefront OBJ ingestion for Scene::load_obj (product host code, C++ stand-in for the Rust host side).
//
// The reference's TriangleMesh::load_obj (/root/reference/scene/src/geometry/impls/triangle_mesh.rs:141-243) is
//     tobj::load_obj(path, LoadOptions { single_index: true, triangulate: true, ignore_points: true, ignore_lines: true })
// followed by a concatenation of the returned models.  tobj is a crates.io dependency (Cargo.lock: tobj 4.0.3) whose source is not
// vendored in the reference tree; what follows restates its published algorithm (load_obj_buf, parse_face, VertexIndices::parse,
// export_faces, add_vertex).  What matters downstream is the ORDER of the output arrays: triangle order feeds the stable sorts
// of the SAH builder (scene/src/bvh.rs:92-295), so BVH-topology parity on a real asset needs the same vertex and face order.
//
//   * statements: v (3 floats, optional colour ignored), vt (first 2 floats), vn (3 floats), f and l (both go through parse_face:
//     an `l` statement with 3+ vertices becomes a face, as in tobj), o / g (close the current model if it has faces), mtllib, usemtl
//     (a change of material closes the current model, but only for materials a loadable .mtl file defines); everything else ignored
//   * face vertices v, v/vt, v//vn, v/vt/vn; negative indices count back from the arrays as they stand when the face is read;
//     0 or an empty field = missing
//   * single_index: a vertex is the (v, vt, vn) triple; triples are numbered in first-use order PER MODEL (a fresh map per model)
//   * triangulate: 3 -> (a,b,c); 4 -> (a,b,c)(a,c,d); n -> fan (a, v[k], v[k+1]); points and lines (1 or 2 vertices) dropped
//   * the last model is always closed at end of file, even when it is empty
// and the reference's own concatenation (triangle_mesh.rs:163-182), reproduced with its two quirks:
//   * `indices.extend(mesh.indices)` adds NO vertex offset, so the indices of a second model address the first model's vertices
//   * the tangent loop runs over ALL triangles gathered so far after every model and pushes again, so with several models
//     `tangents[j]` of a later triangle j is the tangent of an earlier triangle: `tangent_tri[j]` names the triangle whose
//     load-time tangent triangle j ends up with (identity for one-model files)
#pragma once
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace tcpt {

struct ObjData {
    std::vector<float> positions, normals, texcoords;   // concatenated over the models, 3 / 3 / 2 floats per vertex
    std::vector<uint32_t> indices;                       // 3 per triangle, as the reference holds them (no per-model offset)
    std::vector<uint32_t> tangent_tri;                   // per triangle: the triangle whose load-time tangent it receives; empty = identity
    uint32_t n_models = 0;
};

namespace obj_detail {
constexpr size_t MISSING = (size_t)-1;
struct VI { size_t v, vt, vn; bool operator<(const VI& o) const { return std::tie(v, vt, vn) < std::tie(o.v, o.vt, o.vn); } };

inline std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (unsigned char)s[i] <= ' ') ++i;
        size_t j = i;
        while (j < s.size() && (unsigned char)s[j] > ' ') ++j;
        if (j > i) out.emplace_back(s, i, j - i);
        i = j;
    }
    return out;
}
// f32::from_str: decimal / exponent notation, inf, nan; the whole token must parse (strtof also takes hex floats: refused here)
inline bool parse_f32(const std::string& t, float* out) {
    if (t.empty() || t.find_first_of("xXpP") != std::string::npos) return false;
    char* end = nullptr;
    errno = 0;
    const float v = std::strtof(t.c_str(), &end);
    if (end != t.c_str() + t.size()) return false;
    *out = v;
    return true;
}
// parse_floatn: up to n tokens; all of them must parse and exactly n must be found
inline bool parse_floatn(const std::vector<std::string>& w, size_t first, std::vector<float>& vals, size_t n) {
    const size_t sz = vals.size();
    for (size_t k = first; k < w.size() && k < first + n; ++k) {
        float v;
        if (!parse_f32(w[k], &v)) return false;
        vals.push_back(v);
    }
    return sz + n == vals.size();
}
inline bool parse_isize(const std::string& t, long long* out) {
    if (t.empty()) return false;
    size_t i = (t[0] == '+' || t[0] == '-') ? 1 : 0;
    if (i == t.size()) return false;
    for (size_t k = i; k < t.size(); ++k) if (t[k] < '0' || t[k] > '9') return false;
    errno = 0;
    *out = std::strtoll(t.c_str(), nullptr, 10);
    return errno == 0;
}
// VertexIndices::parse
inline bool parse_vertex(const std::string& tok, size_t pos_sz, size_t tex_sz, size_t norm_sz, VI* out) {
    size_t idx[3] = {MISSING, MISSING, MISSING};
    size_t field = 0, start = 0;
    for (;;) {
        const size_t slash = tok.find('/', start);
        const std::string part = tok.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
        if (!part.empty()) {
            long long x;
            if (!parse_isize(part, &x)) return false;
            if (field > 2) return false;
            const size_t sz = field == 0 ? pos_sz : field == 1 ? tex_sz : norm_sz;
            idx[field] = x < 0 ? (size_t)((long long)sz + x) : (size_t)(x - 1);   // 0 -> usize::MAX = missing
        }
        if (slash == std::string::npos) break;
        start = slash + 1; ++field;
    }
    *out = VI{idx[0], idx[1], idx[2]};
    return true;
}

struct Model { std::vector<float> positions, normals, texcoords; std::vector<uint32_t> indices; };

// export_faces + add_vertex (single_index, triangulate, ignore_points, ignore_lines)
inline bool export_faces(const std::vector<float>& pos, const std::vector<float>& tex, const std::vector<float>& nrm, const std::vector<std::vector<VI>>& faces, Model& m, std::string& err) {
    std::map<VI, uint32_t> index_map;
    auto add_vertex = [&](const VI& v) -> bool {
        auto it = index_map.find(v);
        if (it != index_map.end()) { m.indices.push_back(it->second); return true; }
        if (v.v == MISSING || v.v * 3 + 2 >= pos.size()) { err = "face vertex out of bounds"; return false; }
        m.positions.insert(m.positions.end(), {pos[v.v * 3], pos[v.v * 3 + 1], pos[v.v * 3 + 2]});
        if (!tex.empty() && v.vt != MISSING) {
            if (v.vt * 2 + 1 >= tex.size()) { err = "face texcoord out of bounds"; return false; }
            m.texcoords.insert(m.texcoords.end(), {tex[v.vt * 2], tex[v.vt * 2 + 1]});
        }
        if (!nrm.empty() && v.vn != MISSING) {
            if (v.vn * 3 + 2 >= nrm.size()) { err = "face normal out of bounds"; return false; }
            m.normals.insert(m.normals.end(), {nrm[v.vn * 3], nrm[v.vn * 3 + 1], nrm[v.vn * 3 + 2]});
        }
        const uint32_t next = (uint32_t)index_map.size();
        m.indices.push_back(next);
        index_map.emplace(v, next);
        return true;
    };
    for (const auto& f : faces) {
        if (f.size() == 1 || f.size() == 2) continue;                       // ignore_points / ignore_lines
        if (f.empty()) { err = "invalid polygon"; return false; }
        if (f.size() == 3) { if (!add_vertex(f[0]) || !add_vertex(f[1]) || !add_vertex(f[2])) return false; continue; }
        if (f.size() == 4) {
            if (!add_vertex(f[0]) || !add_vertex(f[1]) || !add_vertex(f[2]) || !add_vertex(f[0]) || !add_vertex(f[2]) || !add_vertex(f[3])) return false;
            continue;
        }
        size_t b = 1;
        for (size_t c = 2; c < f.size(); ++c) { if (!add_vertex(f[0]) || !add_vertex(f[b]) || !add_vertex(f[c])) return false; b = c; }
    }
    return true;
}

// names a .mtl file defines, in order (load_mtl_buf: `newmtl <name>`); false = the library could not be loaded (tobj then maps no material)
inline bool mtl_names(const std::string& path, std::vector<std::string>& names) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const auto w = split_ws(line);
        if (w.empty() || w[0] != "newmtl") continue;
        std::string name = line.size() > 6 ? line.substr(6) : "";
        const size_t a = name.find_first_not_of(" \t"), b = name.find_last_not_of(" \t");
        name = a == std::string::npos ? "" : name.substr(a, b - a + 1);
        if (name.empty()) return false;
        names.push_back(name);
    }
    return true;
}
}  // namespace obj_detail

// TriangleMesh::load_obj up to the point where the arrays are complete (tangents and bounds are computed by HostScene::add_mesh)
inline bool load_obj_file(const std::string& path, ObjData& out, std::string& err) {
    using namespace obj_detail;
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    const size_t slash = path.find_last_of("/\\");
    const std::string dir = slash == std::string::npos ? "" : path.substr(0, slash + 1);
    std::vector<float> pos, tex, nrm, colour;
    std::vector<std::vector<VI>> faces;
    std::vector<Model> models;
    std::map<std::string, int> mat_map;
    int mat_id = -1;
    auto close_model = [&]() -> bool {
        Model m;
        if (!export_faces(pos, tex, nrm, faces, m, err)) return false;
        models.push_back(std::move(m));
        faces.clear();
        return true;
    };
    std::string line;
    size_t line_no = 0;
    while (std::getline(in, line)) {
        ++line_no;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const auto w = split_ws(line);
        if (w.empty() || w[0] == "#") continue;
        const std::string& k = w[0];
        auto bad = [&](const char* what) { err = path + ":" + std::to_string(line_no) + ": " + what; return false; };
        if (k == "v") {
            if (!parse_floatn(w, 1, pos, 3)) return bad("position parse error");
            parse_floatn(w, 4, colour, 3);
        } else if (k == "vt") {
            if (!parse_floatn(w, 1, tex, 2)) return bad("texcoord parse error");
        } else if (k == "vn") {
            if (!parse_floatn(w, 1, nrm, 3)) return bad("normal parse error");
        } else if (k == "f" || k == "l") {
            std::vector<VI> f;
            for (size_t i = 1; i < w.size(); ++i) {
                VI v;
                if (!parse_vertex(w[i], pos.size() / 3, tex.size() / 2, nrm.size() / 3, &v)) return bad("face parse error");
                f.push_back(v);
            }
            faces.push_back(std::move(f));
        } else if (k == "o" || k == "g") {
            if (!faces.empty() && !close_model()) return false;
        } else if (k == "mtllib") {
            // the file name may hold spaces: everything after the first blank, trimmed (tobj: line.split_once(' '))
            const size_t sp = line.find(' ');
            std::string file = sp == std::string::npos ? "" : line.substr(sp + 1);
            const size_t a = file.find_first_not_of(" \t"), b = file.find_last_not_of(" \t");
            file = a == std::string::npos ? "" : file.substr(a, b - a + 1);
            std::vector<std::string> names;
            if (!file.empty() && mtl_names(dir + file, names))
                for (const auto& n : names) { const int id = (int)mat_map.size(); mat_map.emplace(n, id); }
        } else if (k == "usemtl") {
            std::string name = line.size() > 7 ? line.substr(7) : "";
            const size_t a = name.find_first_not_of(" \t"), b = name.find_last_not_of(" \t");
            name = a == std::string::npos ? "" : name.substr(a, b - a + 1);
            if (!name.empty()) {
                const auto it = mat_map.find(name);
                const int new_mat = it == mat_map.end() ? -1 : it->second;
                if (mat_id != new_mat && !faces.empty() && !close_model()) return false;
                mat_id = new_mat;
            }
        }
    }
    if (!close_model()) return false;

    // the reference's concatenation (triangle_mesh.rs:163-226)
    out = ObjData();
    out.n_models = (uint32_t)models.size();
    std::vector<uint32_t> pushed;   // the triangle each pushed tangent belongs to
    for (const Model& m : models) {
        out.positions.insert(out.positions.end(), m.positions.begin(), m.positions.end());
        out.normals.insert(out.normals.end(), m.normals.begin(), m.normals.end());
        out.texcoords.insert(out.texcoords.end(), m.texcoords.begin(), m.texcoords.end());
        out.indices.insert(out.indices.end(), m.indices.begin(), m.indices.end());   // no vertex offset
        if (!out.texcoords.empty()) for (uint32_t t = 0; t < out.indices.size() / 3; ++t) pushed.push_back(t);
    }
    const size_t ntri = out.indices.size() / 3;
    bool identity = true;
    for (size_t t = 0; t < ntri && t < pushed.size(); ++t) identity = identity && pushed[t] == t;
    if (!identity) out.tangent_tri.assign(pushed.begin(), pushed.begin() + ntri);
    return true;
}

}  // namespace tcpt
