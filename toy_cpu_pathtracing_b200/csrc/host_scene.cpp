// Host scene preparation (product code, plain C++; compiled with -ffp-contract=off).  See host_scene.h.
#include "host_scene.h"

#include <functional>

#include <algorithm>
#include <cmath>
#include <cstring>

namespace tcpt {

namespace {
// color/src/eotf.rs:63-71
inline float srgb_to_linear(float c) { return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }

// load-time tangent for one triangle (geometry/impls/triangle_mesh.rs:181-225)
V3 tangent_from_normal(V3 n) {  // Normal::generate_tangent (math/src/normal.rs:55-66)
    V3 cand = std::fabs(n.x) > 0.999f ? v3(0, 1, 0) : v3(1, 0, 0);
    float pm = dot3(n, cand);
    return unit(sub(cand, scale(n, pm)));
}
V3 fallback_tangent(V3 e1, V3 e2) {
    V3 cp = cross3(e1, e2);
    if (dot3(cp, cp) < 1e-12f) return v3(1, 0, 0);
    V3 n = unit(unit(cp));  // .normalize().to_normal() normalises twice
    return tangent_from_normal(n);
}
}  // namespace

void HostScene::clear() {
    meshes.clear(); textures.clear(); materials.clear(); primitives.clear(); envs.clear();
    tlas = BuiltBvh(); tlas_prims.clear(); error.clear();
}

int HostScene::add_mesh(const float* pos, const float* nrm, const float* uv, int nverts, const uint32_t* idx, int ntris) {
    if (!pos || !nrm || !idx || nverts <= 0 || ntris <= 0) { error = "add_mesh: empty or null mesh (OBJ files must carry vn normals)"; return TCPT_ERR_INVALID; }
    for (int i = 0; i < ntris * 3; ++i) if (idx[i] >= (uint32_t)nverts) { error = "add_mesh: index out of range"; return TCPT_ERR_INVALID; }
    HostMesh m;
    m.positions.resize(nverts); m.normals.resize(nverts);
    for (int i = 0; i < nverts; ++i) {
        m.positions[i] = v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        m.normals[i] = unit(v3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));  // Normal::new normalises (triangle_mesh.rs:168-172)
    }
    if (uv) { m.uvs.resize(nverts); for (int i = 0; i < nverts; ++i) m.uvs[i] = V2{uv[2 * i], uv[2 * i + 1]}; }
    m.indices.assign(idx, idx + (size_t)ntris * 3);
    if (uv) {
        m.tangents.resize(ntris);
        for (int t = 0; t < ntris; ++t) {
            const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
            V3 e1 = sub(m.positions[i1], m.positions[i0]), e2 = sub(m.positions[i2], m.positions[i0]);
            float du1 = m.uvs[i1].x - m.uvs[i0].x, dv1 = m.uvs[i1].y - m.uvs[i0].y;
            float du2 = m.uvs[i2].x - m.uvs[i0].x, dv2 = m.uvs[i2].y - m.uvs[i0].y;
            float denom = du1 * dv2 - dv1 * du2;
            float r = 1.0f / denom;
            V3 tg = scale(sub(scale(e1, dv2), scale(e2, dv1)), r);  // r * (edge1*duv2.y - edge2*duv1.y): f32 * Vector3
            // NOTE f32*Vector3 multiplies each component as r * c; multiplication commutes, so scale() is the same value
            if (std::fabs(denom) < 1e-6f) tg = fallback_tangent(e1, e2);
            else {
                tg = unit(tg);
                if (any_nan(tg)) tg = fallback_tangent(e1, e2);
            }
            m.tangents[t] = tg;
        }
    }
    const float inf = INFINITY;
    m.bounds = Box{{inf, inf, inf}, {-inf, -inf, -inf}};
    for (const V3& p : m.positions) { m.bounds.lo = min3(m.bounds.lo, p); m.bounds.hi = max3(m.bounds.hi, p); }
    meshes.push_back(std::move(m));
    return (int)meshes.size() - 1;
}

void build_cdf_guide(const float* cdf, uint32_t n, uint32_t G, uint32_t* out) {
    uint32_t k = 0;
    for (uint32_t j = 0; j <= G; ++j) {
        const float edge = (float)j / (float)G;
        while (k < n && cdf[k] <= edge) ++k;
        out[j] = k;
    }
}

int HostScene::set_tangent_source(int mesh, const uint32_t* tri, int ntris) {
    if (mesh < 0 || mesh >= (int)meshes.size() || !tri) { error = "set_tangent_source: bad geometry"; return TCPT_ERR_INVALID; }
    HostMesh& m = meshes[mesh];
    if ((size_t)ntris * 3 != m.indices.size()) { error = "set_tangent_source: triangle count mismatch"; return TCPT_ERR_INVALID; }
    if (m.tangents.empty()) return TCPT_OK;   // no uvs: the reference holds no tangents either
    for (int t = 0; t < ntris; ++t) if (tri[t] >= (uint32_t)ntris) { error = "set_tangent_source: triangle out of range"; return TCPT_ERR_INVALID; }
    std::vector<V3> moved(ntris);
    for (int t = 0; t < ntris; ++t) moved[t] = m.tangents[tri[t]];
    m.tangents.swap(moved);
    m.built = false;
    return TCPT_OK;
}

int HostScene::add_single_triangle(const float pos[9], const float nrm[9], const float uv[6]) {
    if (!pos || !nrm || !uv) { error = "add_single_triangle: null argument"; return TCPT_ERR_INVALID; }
    const uint32_t idx[3] = {0, 1, 2};
    const int g = add_mesh(pos, nrm, uv, 3, idx, 1);
    if (g < 0) return g;
    HostMesh& m = meshes[g];
    m.single = true;
    // SingleTriangle::intersect computes the tangent per hit, without the mesh loader's fallbacks (single_triangle.rs:118-124)
    const V3 e1 = sub(m.positions[1], m.positions[0]), e2 = sub(m.positions[2], m.positions[0]);
    const float du1 = m.uvs[1].x - m.uvs[0].x, dv1 = m.uvs[1].y - m.uvs[0].y, du2 = m.uvs[2].x - m.uvs[0].x, dv2 = m.uvs[2].y - m.uvs[0].y;
    const float r = 1.0f / (du1 * dv2 - dv1 * du2);
    m.tangents[0] = unit(scale(sub(scale(e1, dv2), scale(e2, dv1)), r));
    return g;
}

int HostScene::add_texture(const uint8_t* data, uint32_t w, uint32_t h, uint32_t channels) {
    if (!data || w == 0 || h == 0 || (channels != 1 && channels != 3)) { error = "add_texture: bad arguments"; return TCPT_ERR_INVALID; }
    HostTexture t; t.w = w; t.h = h; t.channels = channels;
    t.data.assign(data, data + (size_t)w * h * channels);
    textures.push_back(std::move(t));
    return (int)textures.size() - 1;
}

// RgbToSpectrumTable::get (spectrum/src/rgb_sigmoid_polynomial.rs:87-155)
bool HostScene::rgb_to_coeffs(const float rgb_in[3], bool gamma_encoded, float cs[3], int32_t index[4]) const {
    if (!tables.set) return false;
    float rgb[3];
    for (int k = 0; k < 3; ++k) {
        float c = gamma_encoded ? srgb_to_linear(rgb_in[k]) : rgb_in[k];
        rgb[k] = c > 0.0f ? c : 0.0f;  // max(0) with NaN -> 0
    }
    float mx = sel_max(rgb[0], sel_max(rgb[1], rgb[2]));
    if (mx > 1.0f) return false;  // the reference panics here
    if (index) { index[0] = -1; index[1] = index[2] = index[3] = 0; }
    if (rgb[0] == rgb[1] && rgb[1] == rgb[2]) { cs[0] = 0; cs[1] = 0; cs[2] = std::log(rgb[0] / (1.0f - rgb[0])); return true; }
    int m = 0;
    { float best = rgb[0]; if (rgb[1] > best) { best = rgb[1]; m = 1; } if (rgb[2] > best) m = 2; }
    const float* zn = tables.rgb2spec.data();
    const float* tab = zn + 64;
    const float z = rgb[m];
    const float x = rgb[(m + 1) % 3] * (64.0f - 1.0f) / z;
    const float y = rgb[(m + 2) % 3] * (64.0f - 1.0f) / z;
    uint32_t xi = x > 0.0f ? (uint32_t)x : 0; if (xi > 62) xi = 62;
    uint32_t yi = y > 0.0f ? (uint32_t)y : 0; if (yi > 62) yi = 62;
    uint32_t zi = 62;
    for (uint32_t i = 0; i <= 62; ++i) if (zn[i + 1] > z) { zi = i; break; }
    const float dx = x - (float)xi, dy = y - (float)yi, dz = (z - zn[zi]) / (zn[zi + 1] - zn[zi]);
    if (index) { index[0] = m; index[1] = (int32_t)zi; index[2] = (int32_t)yi; index[3] = (int32_t)xi; }
    for (int i = 0; i < 3; ++i) {
        auto co = [&](uint32_t a, uint32_t b, uint32_t c) { return tab[((((size_t)m * 64 + (zi + c)) * 64 + (yi + b)) * 64 + (xi + a)) * 3 + i]; };
        auto lerp = [](float a, float b, float t) { return a + (b - a) * t; };
        cs[i] = lerp(lerp(lerp(co(0, 0, 0), co(1, 0, 0), dx), lerp(co(0, 1, 0), co(1, 1, 0), dx), dy),
                     lerp(lerp(co(0, 0, 1), co(1, 0, 1), dx), lerp(co(0, 1, 1), co(1, 1, 1), dx), dy), dz);
    }
    return true;
}

tcpt_flat_spectrum HostScene::resolve_spectrum(const tcpt_spectrum_param& p) const {
    tcpt_flat_spectrum s{};
    s.texture = -1; s.scale = 1.0f;
    switch (p.kind) {
        case TCPT_SPEC_CONSTANT: s.kind = 0; s.c[0] = p.value[0]; break;
        case TCPT_SPEC_RGB_ALBEDO_SRGB:
        case TCPT_SPEC_RGB_ALBEDO_LINEAR:
            s.kind = 1;
            if (!rgb_to_coeffs(p.value, p.kind == TCPT_SPEC_RGB_ALBEDO_SRGB, s.c, nullptr)) { s.c[0] = s.c[1] = 0; s.c[2] = INFINITY; }
            break;
        case TCPT_SPEC_D65: s.kind = 3; break;
        case TCPT_SPEC_TEXTURE_SRGB: s.kind = 4; s.texture = p.texture; break;
        case TCPT_SPEC_PRESET: s.kind = 5; s.texture = p.texture; break;
        default: s.kind = 0; break;
    }
    return s;
}

// RgbIlluminantSpectrum::<ColorSrgb>::new (spectrum/src/spectrum/rgb_illuminant_spectrum.rs:27-40)
tcpt_flat_spectrum HostScene::illuminant_from_rgb(const float rgb[3]) const {
    tcpt_flat_spectrum s{};
    s.kind = 2; s.texture = -1;
    float mx = sel_max(rgb[0], sel_max(rgb[1], rgb[2]));
    s.scale = 2.0f * mx;
    float scaled[3] = {rgb[0] / s.scale, rgb[1] / s.scale, rgb[2] / s.scale};
    if (!rgb_to_coeffs(scaled, true, s.c, nullptr)) { s.c[0] = s.c[1] = 0; s.c[2] = INFINITY; }
    return s;
}

int HostScene::add_material(const tcpt_material_desc& d) {
    if (!tables.set) { error = "add_material: call tcpt_set_tables first"; return TCPT_ERR_INVALID; }
    auto cf = [](const tcpt_float_param& p) { return tcpt_flat_float{p.kind == 1, p.value, p.texture, p.gamma_corrected}; };
    auto tex_ok = [&](int t) { return t >= 0 && t < (int)textures.size(); };
    if ((d.color.kind == TCPT_SPEC_TEXTURE_SRGB && !tex_ok(d.color.texture)) || (d.normal.texture >= 0 && !tex_ok(d.normal.texture))) { error = "add_material: texture index out of range"; return TCPT_ERR_INVALID; }
    auto preset_ok = [&](const tcpt_spectrum_param& p) { return p.kind != TCPT_SPEC_PRESET || (p.texture >= 0 && (size_t)p.texture * 470 < tables.presets.size()); };
    if (!preset_ok(d.color) || !preset_ok(d.coat_tint)) { error = "add_material: spectrum preset id out of range (std_tables blob without presets?)"; return TCPT_ERR_INVALID; }
    if (d.type < TCPT_MAT_LAMBERT || d.type > TCPT_MAT_GLASS) { error = "add_material: unknown material type"; return TCPT_ERR_INVALID; }
    if ((d.type == TCPT_MAT_METAL && (d.color.kind != TCPT_SPEC_PRESET || d.coat_tint.kind != TCPT_SPEC_PRESET)) || (d.type == TCPT_MAT_GLASS && d.color.kind != TCPT_SPEC_PRESET)) {
        error = "add_material: metal needs eta and k presets (color, coat_tint), glass an eta preset (color)"; return TCPT_ERR_INVALID;
    }
    // RgbSigmoidPolynomial::from_rgb panics when a (linearised) component exceeds 1 (rgb_sigmoid_polynomial.rs:95-108): refuse the material
    auto rgb_ok = [&](const tcpt_spectrum_param& p) {
        if (p.kind != TCPT_SPEC_RGB_ALBEDO_SRGB && p.kind != TCPT_SPEC_RGB_ALBEDO_LINEAR) return true;
        float cs[3];
        return rgb_to_coeffs(p.value, p.kind == TCPT_SPEC_RGB_ALBEDO_SRGB, cs, nullptr);
    };
    if (!rgb_ok(d.color) || !rgb_ok(d.coat_tint)) { error = "add_material: an RGB albedo component exceeds 1 after linearisation (the reference panics: rgb_sigmoid_polynomial.rs:95-108)"; return TCPT_ERR_INVALID; }
    tcpt_flat_material m{};
    m.type = d.type;
    m.color = resolve_spectrum(d.color);
    m.coat_tint = resolve_spectrum(d.coat_tint);
    m.intensity = cf(d.intensity); m.roughness = cf(d.roughness); m.metallic = cf(d.metallic); m.ior = cf(d.ior);
    m.coat_ior = cf(d.coat_ior); m.coat_roughness = cf(d.coat_roughness); m.coat_thickness = cf(d.coat_thickness);
    m.normal_texture = d.normal.texture; m.normal_flip_y = d.normal.flip_y;
    m.eta = d.eta; m.thin_surface = d.thin_surface;
    materials.push_back(m);
    return (int)materials.size() - 1;
}

int HostScene::add_primitive(int geometry, int material, const float l2w[16]) {
    if (geometry < 0 || geometry >= (int)meshes.size() || material < 0 || material >= (int)materials.size()) { error = "add_primitive: bad index"; return TCPT_ERR_INVALID; }
    HostPrimitive p;
    p.geometry = geometry; p.material = material;
    std::memcpy(p.local_to_world.m, l2w, 64);
    p.kind = materials[material].type == TCPT_MAT_EMISSIVE ? 1 : 0;
    if (p.kind == 1) {  // EmissiveTriangleMesh::new (primitive/impls/emissive_triangle_mesh.rs:28-69): world-space triangle areas
        const HostMesh& m = meshes[geometry];
        const size_t nt = m.indices.size() / 3;
        p.area_list.resize(nt); p.area_table.resize(nt);
        for (size_t t = 0; t < nt; ++t) {
            V3 a = m4_point(p.local_to_world, m.positions[m.indices[3 * t]]);
            V3 b = m4_point(p.local_to_world, m.positions[m.indices[3 * t + 1]]);
            V3 c = m4_point(p.local_to_world, m.positions[m.indices[3 * t + 2]]);
            p.area_list[t] = len3(cross3(sub(a, b), sub(a, c))) * 0.5f;
        }
        float run = 0.0f;
        for (size_t t = 0; t < nt; ++t) { run += p.area_list[t]; p.area_table[t] = run; }
        p.area_sum = run;
        for (float& v : p.area_table) v /= run;
    }
    primitives.push_back(std::move(p));
    return (int)primitives.size() - 1;
}

int HostScene::add_env_light(float intensity, const float* rgb, uint32_t w, uint32_t h, const float l2w[16]) {
    if (!tables.set) { error = "add_env_light: call tcpt_set_tables first"; return TCPT_ERR_INVALID; }
    if (!rgb || w < 2 || h < 2) { error = "add_env_light: bad image"; return TCPT_ERR_INVALID; }
    HostEnv e; e.intensity = intensity; e.w = w; e.h = h;
    e.data.assign(rgb, rgb + (size_t)w * h * 3);
    // mean colour -> integrated illuminant spectrum (environment_light.rs:49-66)
    float tot[3] = {0, 0, 0};
    for (size_t i = 0; i < (size_t)w * h; ++i) { tot[0] += e.data[3 * i]; tot[1] += e.data[3 * i + 1]; tot[2] += e.data[3 * i + 2]; }
    const float pc = (float)(w * h);
    for (float& t : tot) t /= pc;
    e.integrated = illuminant_from_rgb(tot);
    // luminance * sin(theta) two-level CDF (environment_light.rs:165-215)
    const float pi = 3.14159265358979323846f;
    std::vector<float> row_w(h);
    e.conditional.assign((size_t)w * h, 0.0f);
    for (uint32_t y = 0; y < h; ++y) {
        const float theta = (((float)y + 0.5f) / (float)h) * pi;
        const float jw = sel_max(std::sin(theta), 1e-8f);
        float run = 0.0f;
        for (uint32_t x = 0; x < w; ++x) {
            const float* px = &e.data[((size_t)y * w + x) * 3];
            const float lum = 0.299f * px[0] + 0.587f * px[1] + 0.114f * px[2];
            run += lum * jw;
            e.conditional[(size_t)y * w + x] = run;
        }
        row_w[y] = run;
        if (run > 0.0f) for (uint32_t x = 0; x < w; ++x) e.conditional[(size_t)y * w + x] /= run;
    }
    float total = 0.0f;
    for (float r : row_w) total += r;
    e.total_weight = total;
    e.marginal.resize(h);
    float run = 0.0f;
    for (uint32_t y = 0; y < h; ++y) { run += row_w[y]; e.marginal[y] = total > 0.0f ? run / total : (float)(y + 1) / (float)h; }
    // guide tables: guide[j] = #{i : cdf[i] <= j/G} for j = 0..G (cdf is non-decreasing, j/G is exact in f32)
    auto pow2_at_least = [](uint32_t n) { uint32_t g = 1; while (g < n) g <<= 1; return g; };
    auto build_guide = build_cdf_guide;
    // four guide cells per CDF entry while the conditional guides stay below 64 MB: a warp waits for its longest scan, and the scans are
    // long exactly where the map is dark (many entries per cell); the search result does not depend on G
    const uint32_t fine = (size_t)h * (4 * (size_t)pow2_at_least(w) + 1) * 4 <= ((size_t)64 << 20) ? 4u : 1u;
    e.guide_h = fine * pow2_at_least(h); e.guide_w = fine * pow2_at_least(w);
    e.marginal_guide.resize(e.guide_h + 1);
    build_guide(e.marginal.data(), h, e.guide_h, e.marginal_guide.data());
    e.conditional_guide.resize((size_t)h * (e.guide_w + 1));
    for (uint32_t y = 0; y < h; ++y) build_guide(&e.conditional[(size_t)y * w], w, e.guide_w, &e.conditional_guide[(size_t)y * (e.guide_w + 1)]);
    envs.push_back(std::move(e));
    HostPrimitive p; p.kind = 2; p.env = (int)envs.size() - 1;
    std::memcpy(p.local_to_world.m, l2w, 64);
    primitives.push_back(std::move(p));
    return (int)primitives.size() - 1;
}

int HostScene::add_delta_light(int kind, float intensity, const tcpt_spectrum_param& spectrum, float angle_inner, float angle_outer, const float l2w[16]) {
    if (!tables.set) { error = "add_delta_light: call tcpt_set_tables first"; return TCPT_ERR_INVALID; }
    if (kind < TCPT_LIGHT_POINT || kind > TCPT_LIGHT_DIRECTIONAL) { error = "add_delta_light: unknown light kind"; return TCPT_ERR_INVALID; }
    if (spectrum.kind == TCPT_SPEC_TEXTURE_SRGB) { error = "add_delta_light: a light spectrum cannot be a texture"; return TCPT_ERR_INVALID; }
    if (spectrum.kind == TCPT_SPEC_PRESET && (spectrum.texture < 0 || (size_t)spectrum.texture * 470 >= tables.presets.size())) { error = "add_delta_light: spectrum preset id out of range"; return TCPT_ERR_INVALID; }
    if (spectrum.kind == TCPT_SPEC_RGB_ALBEDO_SRGB || spectrum.kind == TCPT_SPEC_RGB_ALBEDO_LINEAR) {
        float cs[3];
        if (!rgb_to_coeffs(spectrum.value, spectrum.kind == TCPT_SPEC_RGB_ALBEDO_SRGB, cs, nullptr)) { error = "add_delta_light: an RGB component exceeds 1 after linearisation (the reference panics: rgb_sigmoid_polynomial.rs:95-108)"; return TCPT_ERR_INVALID; }
    }
    HostPrimitive p; p.kind = kind; p.light_intensity = intensity; p.angle_inner = angle_inner; p.angle_outer = angle_outer;
    p.light_spectrum = resolve_spectrum(spectrum);
    std::memcpy(p.local_to_world.m, l2w, 64);
    primitives.push_back(std::move(p));
    return (int)primitives.size() - 1;
}

// Convert the reference-order binary tree (pre-order inner / leaf records) into the device's 4-wide records (include/tcpt_flat.h).
// A wide node starts from a binary node's two children and repeatedly replaces its largest-area inner child by that child's two
// children, in place, until it holds four children or only leaves (the children keep the reference's DFS order).  Every reference
// LEAF survives with its own box bits and item order; reference inner boxes that end up interior to a wide node are not stored:
// a ray that passes a leaf box passes all of its ancestors' boxes (exact min / max merges, monotone slab test; dtraverse.cuh).
// `slot_base` = absolute index of the BVH's first item slot; node indices written into the records are absolute (`out` index).
// Returns the number of records appended; *max_stack receives the deepest the traversal stack can get inside this BVH.
static uint32_t put_nodes(const BuiltBvh& b, std::vector<tcpt_bvh_node>& out, uint32_t slot_base, uint32_t* max_stack) {
    const size_t base = out.size();
    const float inf = INFINITY;
    auto blank = [&]() {
        tcpt_bvh_node n;
        for (int k = 0; k < 4; ++k) {
            n.q[0 + k] = inf; n.q[4 + k] = -inf; n.q[8 + k] = inf; n.q[12 + k] = -inf; n.q[16 + k] = inf; n.q[20 + k] = -inf;
            const uint32_t none = 0xffffffffu, zero = 0u;
            std::memcpy(&n.q[24 + k], &none, 4); std::memcpy(&n.q[28 + k], &zero, 4);
        }
        return n;
    };
    auto set_child = [&](size_t rec, int k, const Box& box, uint32_t entry, uint32_t cnt) {
        float* q = out[rec].q;
        q[0 + k] = box.lo.x; q[4 + k] = box.hi.x; q[8 + k] = box.lo.y; q[12 + k] = box.hi.y; q[16 + k] = box.lo.z; q[20 + k] = box.hi.z;
        std::memcpy(&q[24 + k], &entry, 4); std::memcpy(&q[28 + k], &cnt, 4);
    };
    *max_stack = 0;
    if (b.nodes.empty()) { out.push_back(blank()); return 1; }
    // a range of leaf items as a child: one entry for up to 16 items; a longer reference leaf becomes a wide node of up to four
    // sub-ranges that all carry the LEAF's box (the same test decides all of them, as it decides the whole leaf in the reference)
    struct Pending { size_t rec; int slot; uint32_t node; };            // binary node `node` to be written as child `slot` of record `rec`
    std::vector<Pending> todo;
    // returns the stack depth the subtree can reach
    std::function<uint32_t(size_t, int, const Box&, uint32_t, uint32_t)> put_range = [&](size_t rec, int k, const Box& box, uint32_t first, uint32_t cnt) -> uint32_t {
        if (cnt <= 16) { set_child(rec, k, box, 0x80000000u | ((cnt - 1u) << 27) | (slot_base + first), cnt); return 0; }
        const size_t sub = out.size();
        out.push_back(blank());
        set_child(rec, k, box, (uint32_t)sub, 0);
        uint32_t parts = (cnt + 15u) / 16u; if (parts > 4u) parts = 4u;
        uint32_t depth = 0, done = 0;
        for (uint32_t i = 0; i < parts; ++i) {
            const uint32_t take = (cnt - done) / (parts - i);
            depth = std::max(depth, put_range(sub, (int)i, box, first + done, take));
            done += take;
        }
        return depth + parts - 1u;
    };
    std::function<uint32_t(uint32_t)> put_wide = [&](uint32_t root) -> uint32_t {   // binary inner node (or a leaf root) -> one wide record; pre-order
        const size_t rec = out.size();
        out.push_back(blank());
        std::vector<uint32_t> kids;
        const BuildNode& rn = b.nodes[root];
        if (rn.count) kids.push_back(root);
        else { kids.push_back(root + 1); kids.push_back(rn.second); }
        while (kids.size() < 4) {
            int best = -1; float best_area = -1.0f;
            for (size_t i = 0; i < kids.size(); ++i) {
                const BuildNode& c = b.nodes[kids[i]];
                if (c.count == 0 && c.box.half_area2() > best_area) { best_area = c.box.half_area2(); best = (int)i; }
            }
            if (best < 0) break;
            const uint32_t c = kids[best];
            kids[best] = c + 1;
            kids.insert(kids.begin() + best + 1, b.nodes[c].second);
        }
        uint32_t depth = 0;
        for (size_t i = 0; i < kids.size(); ++i) {
            const BuildNode& c = b.nodes[kids[i]];
            uint32_t d;
            if (c.count) d = put_range(rec, (int)i, c.box, c.first_item, c.count);
            else { const size_t child_rec = out.size(); d = put_wide(kids[i]); set_child(rec, (int)i, c.box, (uint32_t)child_rec, 0); }
            depth = std::max(depth, d);
        }
        return depth + (uint32_t)kids.size() - 1u;   // the siblings of the child being walked wait on the stack
    };
    *max_stack = put_wide(0);
    return (uint32_t)(out.size() - base);
}

// first slot of the leaf that owns each item slot
static std::vector<uint32_t> leaf_first_of_slots(const BuiltBvh& b) {
    std::vector<uint32_t> r(b.items.size(), 0);
    for (const BuildNode& n : b.nodes) for (uint32_t j = 0; j < n.count; ++j) r[n.first_item + j] = n.first_item;
    return r;
}

int HostScene::build(const float cam_pos[3], FlatStorage& S) {
    S = FlatStorage();
    if (primitives.empty()) { error = "build: scene has no primitives"; return TCPT_ERR_INVALID; }
    const M4 world_to_render = m4_translation(v3(-cam_pos[0], -cam_pos[1], -cam_pos[2]));  // camera.rs:84-86

    // BLAS per geometry that is referenced (PrimitiveBvh::build, primitive/bvh.rs:116-136)
    for (const HostPrimitive& p : primitives) {
        if (p.geometry < 0) continue;
        HostMesh& m = meshes[p.geometry];
        if (m.built) continue;
        const size_t nt = m.indices.size() / 3;
        std::vector<Box> ib(nt);
        for (size_t t = 0; t < nt; ++t) {
            const V3 a = m.positions[m.indices[3 * t]], b = m.positions[m.indices[3 * t + 1]], c = m.positions[m.indices[3 * t + 2]];
            ib[t] = Box{min3(min3(a, b), c), max3(max3(a, b), c)};
        }
        m.bvh = use_binned_builder ? BinnedBuilder(ib).build() : SahBuilder(ib).build();
        m.built = true;
    }

    // primitives + TLAS over geometry primitives in creation order (primitive/bvh.rs:74-80)
    tlas_prims.clear();
    std::vector<Box> tb;
    std::vector<M4> l2r(primitives.size());
    for (size_t i = 0; i < primitives.size(); ++i) {
        l2r[i] = m4_mul(world_to_render, primitives[i].local_to_world);
        if (primitives[i].geometry < 0) continue;
        tlas_prims.push_back((int)i);
        const HostMesh& gm = meshes[primitives[i].geometry];
        if (gm.single) {  // SingleTriangle::bounds: box of the three transformed vertices (single_triangle.rs:86-94)
            const float inf = INFINITY;
            Box b{{inf, inf, inf}, {-inf, -inf, -inf}};
            for (int k = 0; k < 3; ++k) { const V3 q = m4_point(l2r[i], gm.positions[k]); b.lo = min3(b.lo, q); b.hi = max3(b.hi, q); }
            tb.push_back(b);
        } else tb.push_back(transform_box(l2r[i], gm.bounds));
    }
    if (tb.empty()) { error = "build: no geometry primitives"; return TCPT_ERR_INVALID; }
    tlas = SahBuilder(tb).build();

    uint32_t tlas_stack = 0;
    put_nodes(tlas, S.nodes, 0, &tlas_stack);
    {
        const std::vector<uint32_t> lf = leaf_first_of_slots(tlas);
        for (size_t k = 0; k < tlas.items.size(); ++k) { S.tlas_items.push_back(tlas_prims[tlas.items[k]]); S.tlas_items.push_back((int32_t)lf[k]); }
    }
    const uint32_t tlas_nodes = (uint32_t)S.nodes.size();

    uint32_t deepest = 0;
    geom_flat.assign(meshes.size(), -1);
    for (size_t g = 0; g < meshes.size(); ++g) {
        const HostMesh& m = meshes[g];
        if (!m.built) continue;
        tcpt_flat_geometry fg{};
        fg.node_base = (uint32_t)S.nodes.size();
        fg.slot_base = (uint32_t)(S.tri_verts.size() / 12); fg.tri_count = (uint32_t)(m.indices.size() / 3);
        fg.vertex_base = (uint32_t)(S.positions.size() / 3); fg.index_base = (uint32_t)(S.indices.size() / 3);
        fg.tangent_base = (uint32_t)(S.tangents.size() / 3); fg.has_uv = m.uvs.empty() ? 0 : 1; fg.single = m.single ? 1 : 0;
        uint32_t blas_stack = 0;
        fg.node_count = put_nodes(m.bvh, S.nodes, fg.slot_base, &blas_stack);
        const std::vector<uint32_t> lf = leaf_first_of_slots(m.bvh);
        for (size_t slot = 0; slot < m.bvh.items.size(); ++slot) {
            const uint32_t tri = m.bvh.items[slot];
            V3 p[3] = {m.positions[m.indices[3 * tri]], m.positions[m.indices[3 * tri + 1]], m.positions[m.indices[3 * tri + 2]]};
            V3 n = cross3(sub(p[1], p[0]), sub(p[2], p[0]));
            uint32_t degenerate = dot3(n, n) == 0.0f ? 1u : 0u;  // math/src/ray.rs:50-57, decided once on the host with the same arithmetic
            for (int k = 0; k < 3; ++k) {
                float w; uint32_t bits = k == 0 ? tri : (k == 1 ? degenerate : lf[slot]);
                std::memcpy(&w, &bits, 4);
                S.tri_verts.insert(S.tri_verts.end(), {p[k].x, p[k].y, p[k].z, w});
            }
        }
        for (const V3& v : m.positions) S.positions.insert(S.positions.end(), {v.x, v.y, v.z});
        for (const V3& v : m.normals) S.normals.insert(S.normals.end(), {v.x, v.y, v.z});
        if (m.uvs.empty()) S.uvs.resize(S.uvs.size() + 2 * m.positions.size(), 0.0f);
        else for (const V2& v : m.uvs) S.uvs.insert(S.uvs.end(), {v.x, v.y});
        S.indices.insert(S.indices.end(), m.indices.begin(), m.indices.end());
        if (m.tangents.empty()) S.tangents.resize(S.tangents.size() + m.indices.size(), 0.0f);
        else for (const V3& v : m.tangents) S.tangents.insert(S.tangents.end(), {v.x, v.y, v.z});
        geom_flat[g] = (int)S.geometries.size();
        S.geometries.push_back(fg);
        deepest = std::max(deepest, blas_stack);
    }

    // light list in primitive order (light_sampler.rs:168-187)
    for (size_t i = 0; i < primitives.size(); ++i) if (primitives[i].kind != 0) S.light_list.push_back((int)i);
    // DirectionalLight::preprocess (directional_light.rs:85-89): bounding sphere of the scene bounds = root box of the TLAS
    // (primitive/bvh.rs:139-141, math/src/bounds.rs:62-77)
    float dir_area = 0.0f;
    {
        const Box sb = tlas.nodes[0].box;
        const V3 center = scale(add(sb.lo, sb.hi), 0.5f);
        const float radius = len3(sub(center, sb.hi));  // center.distance(max)
        dir_area = 3.14159265358979323846f * radius * radius;
    }
    if (S.light_list.size() > TCPT_MAX_LIGHTS) { error = "build: too many lights"; return TCPT_ERR_LIMIT; }

    for (size_t i = 0; i < primitives.size(); ++i) {
        const HostPrimitive& p = primitives[i];
        tcpt_flat_primitive fp{};
        const M4 inv = m4_inverse(l2r[i]);
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) { fp.l2r[c * 3 + r] = l2r[i].at(c, r); fp.r2l[c * 3 + r] = inv.at(c, r); }
        fp.geometry = p.geometry >= 0 ? geom_flat[p.geometry] : -1;
        fp.material = p.material; fp.kind = p.kind; fp.env = p.env;
        fp.identity = m4_is_identity(l2r[i]) && m4_is_identity(inv);
        fp.light_index = -1;
        for (size_t k = 0; k < S.light_list.size(); ++k) if (S.light_list[k] == (int)i) fp.light_index = (int)k;
        fp.light_intensity = p.light_intensity; fp.angle_inner = p.angle_inner; fp.angle_outer = p.angle_outer; fp.light_spectrum = p.light_spectrum;
        fp.dir_area = p.kind == TCPT_LIGHT_DIRECTIONAL ? dir_area : 0.0f;
        fp.area_base = (uint32_t)S.area_list.size(); fp.area_sum = p.area_sum;
        S.area_list.insert(S.area_list.end(), p.area_list.begin(), p.area_list.end());
        S.area_table.insert(S.area_table.end(), p.area_table.begin(), p.area_table.end());
        S.primitives.push_back(fp);
    }
    S.materials = materials;
    for (const HostTexture& t : textures) {
        tcpt_flat_texture ft{};
        ft.offset = S.texture_bytes.size(); ft.width = t.w; ft.height = t.h; ft.channels = t.channels;
        S.texture_bytes.insert(S.texture_bytes.end(), t.data.begin(), t.data.end());
        while (S.texture_bytes.size() % 16) S.texture_bytes.push_back(0);
        S.textures.push_back(ft);
    }
    for (size_t i = 0; i < envs.size(); ++i) {
        const HostEnv& e = envs[i];
        tcpt_flat_env fe{};
        fe.intensity = e.intensity; fe.width = e.w; fe.height = e.h; fe.total_weight = e.total_weight; fe.integrated = e.integrated;
        fe.data_offset = S.env_floats.size(); S.env_floats.insert(S.env_floats.end(), e.data.begin(), e.data.end());
        fe.marginal_offset = S.env_floats.size(); S.env_floats.insert(S.env_floats.end(), e.marginal.begin(), e.marginal.end());
        fe.conditional_offset = S.env_floats.size(); S.env_floats.insert(S.env_floats.end(), e.conditional.begin(), e.conditional.end());
        fe.guide_h = e.guide_h; fe.guide_w = e.guide_w;
        fe.marginal_guide_offset = S.env_guides.size(); S.env_guides.insert(S.env_guides.end(), e.marginal_guide.begin(), e.marginal_guide.end());
        fe.conditional_guide_offset = S.env_guides.size(); S.env_guides.insert(S.env_guides.end(), e.conditional_guide.begin(), e.conditional_guide.end());
        fe.primitive = -1;
        for (size_t k = 0; k < primitives.size(); ++k) if (primitives[k].env == (int)i) fe.primitive = (int)k;
        S.envs.push_back(fe);
    }

    tcpt_flat_scene& v = S.view;
    v.bvh_nodes = S.nodes.data(); v.n_bvh_nodes = S.nodes.size(); v.tlas_node_count = tlas_nodes;
    v.tlas_items = S.tlas_items.data(); v.n_tlas_items = (uint32_t)(S.tlas_items.size() / 2);
    v.tri_verts = S.tri_verts.data(); v.n_tri_slots = S.tri_verts.size() / 12;
    v.positions = S.positions.data(); v.normals = S.normals.data(); v.uvs = S.uvs.data(); v.n_vertices = S.positions.size() / 3;
    v.indices = S.indices.data(); v.n_triangles = S.indices.size() / 3; v.tangents = S.tangents.data();
    v.geometries = S.geometries.data(); v.n_geometries = (uint32_t)S.geometries.size();
    v.primitives = S.primitives.data(); v.n_primitives = (uint32_t)S.primitives.size();
    v.materials = S.materials.data(); v.n_materials = (uint32_t)S.materials.size();
    v.textures = S.textures.data(); v.n_textures = (uint32_t)S.textures.size();
    v.texture_bytes = S.texture_bytes.data(); v.n_texture_bytes = S.texture_bytes.size();
    v.area_list = S.area_list.data(); v.area_table = S.area_table.data(); v.n_area = S.area_list.size();
    v.light_list = S.light_list.data(); v.n_lights = (uint32_t)S.light_list.size();
    v.envs = S.envs.data(); v.n_envs = (uint32_t)S.envs.size();
    v.env_floats = S.env_floats.data(); v.n_env_floats = S.env_floats.size();
    v.env_guides = S.env_guides.data(); v.n_env_guides = S.env_guides.size();
    // worst-case height of the traversal stack: the siblings waiting along the deepest TLAS path, the rest of a TLAS leaf's items
    // (one range entry), then the same inside the deepest BLAS (put_nodes walks every root-to-leaf path)
    v.max_bvh_depth = tlas_stack + 1 + deepest;
    if (v.max_bvh_depth >= TCPT_TRAVERSAL_STACK) { error = "build: BVH deeper than the traversal stack"; return TCPT_ERR_LIMIT; }
    if (v.n_tri_slots >= 0x07ffffffu || v.n_tlas_items >= 0x07ffffffu) { error = "build: more item slots than a stack entry can address (2^27)"; return TCPT_ERR_LIMIT; }
    return TCPT_OK;
}

// reference-order dump (bvh.rs:234-295): node-only pre-order + inline item records
int HostScene::dump_built(const BuiltBvh& b, uint32_t* out, int max_nodes) {
    const size_t n = b.nodes.size();
    std::vector<uint32_t> ref_index(n + 1);
    uint32_t items_before = 0;
    for (size_t i = 0; i < n; ++i) { ref_index[i] = (uint32_t)i + items_before; items_before += b.nodes[i].count; }
    ref_index[n] = (uint32_t)n + items_before;
    int k = 0;
    auto put = [&](uint32_t kind, uint32_t value, const Box* box) {
        if (k < max_nodes) {
            uint32_t* o = out + 8 * (size_t)k;
            o[0] = kind; o[1] = value;
            float f[6] = {0, 0, 0, 0, 0, 0};
            if (box) { f[0] = box->lo.x; f[1] = box->lo.y; f[2] = box->lo.z; f[3] = box->hi.x; f[4] = box->hi.y; f[5] = box->hi.z; }
            std::memcpy(o + 2, f, 24);
        }
        ++k;
    };
    for (size_t i = 0; i < n; ++i) {
        const BuildNode& nd = b.nodes[i];
        if (nd.count == 0) put(0, ref_index[nd.second] - ref_index[i], &nd.box);
        else {
            put(1, nd.count, &nd.box);
            for (uint32_t j = 0; j < nd.count; ++j) put(2, b.items[nd.first_item + j], nullptr);
        }
    }
    return k;
}

int HostScene::dump_bvh(int which, uint32_t* out, int max_nodes) const {
    if (which < 0) return dump_built(tlas, out, max_nodes);
    if (which >= (int)meshes.size()) return TCPT_ERR_INVALID;
    return dump_built(meshes[which].bvh, out, max_nodes);
}

}  // namespace tcpt
