// ORACLE (test infrastructure).  Scene layer: triangle-mesh geometry (BLAS), primitives (TLAS), emissive mesh lights,
// environment light, power-weighted light sampler.  Restated from
//   /root/reference/scene/src/{scene,light_sampler,samples}.rs, geometry/impls/triangle_mesh.rs, primitive/bvh.rs,
//   primitive/impls/{triangle_mesh,emissive_triangle_mesh,environment_light}.rs, material/impls/emissive_material.rs.
#pragma once
#include <cmath>
#include <memory>
#include <vector>

#include "obsdf.h"
#include "obvh.h"
#include "omath.h"
#include "ospectrum.h"

namespace orc {

// geometry/impls/triangle_mesh.rs
struct Mesh {
    std::vector<Vec3> positions, normals, tangents;
    std::vector<Vec2> uvs;
    std::vector<uint32_t> indices;
    Bounds bounds;
    Bvh bvh;
    bool bvh_built = false;
    bool single = false;  // geometry of a SingleTriangle primitive (primitive/impls/single_triangle.rs): no BVH, no box tests, raw tangent

    // load_obj post-processing (triangle_mesh.rs:161-243): normalise normals, per-triangle tangents when UVs exist
    void finalize() {
        for (auto& n : normals) n = make_normal(n);  // Normal::new normalises (normal.rs:19-21)
        tangents.clear();
        if (!uvs.empty()) {
            for (size_t t = 0; t + 2 < indices.size(); t += 3) {
                Vec3 p0 = positions[indices[t]], p1 = positions[indices[t + 1]], p2 = positions[indices[t + 2]];
                Vec3 e1 = p1 - p0, e2 = p2 - p0;
                Vec2 uv0 = uvs[indices[t]], uv1 = uvs[indices[t + 1]], uv2 = uvs[indices[t + 2]];
                float d1x = uv1.x - uv0.x, d1y = uv1.y - uv0.y, d2x = uv2.x - uv0.x, d2y = uv2.y - uv0.y;
                float denom = d1x * d2y - d1y * d2x;
                float r = 1.0f / denom;
                Vec3 tangent = r * (e1 * d2y - e2 * d1y);
                auto fallback = [&]() {
                    Vec3 cp = cross(e1, e2);
                    if (length_squared(cp) < 1e-12f) return Vec3(1, 0, 0);
                    Vec3 n = make_normal(normalize(cp));
                    return generate_tangent(n);
                };
                if (single) tangent = normalize(tangent);  // single_triangle.rs:118-124: no fallbacks
                else if (std::fabs(denom) < 1e-6f) tangent = fallback();
                else {
                    tangent = normalize(tangent);
                    if (is_nan(tangent)) tangent = fallback();
                }
                tangents.push_back(tangent);
            }
        }
        const float inf = INFINITY;
        Vec3 mn(inf, inf, inf), mx(-inf, -inf, -inf);
        for (auto& p : positions) { mn = vmin(mn, p); mx = vmax(mx, p); }
        bounds = Bounds{mn, mx};
    }
    // Normal::orthogonalize_vector / generate_tangent (normal.rs:46-66)
    static Vec3 orthogonalize(Vec3 n, Vec3 v) { float pm = dot(n, v); return normalize(v - n * pm); }
    static Vec3 generate_tangent(Vec3 n) { Vec3 cand = std::fabs(n.x) > 0.999f ? Vec3(0, 1, 0) : Vec3(1, 0, 0); return orthogonalize(n, cand); }

    size_t n_tris() const { return indices.size() / 3; }
    void tri_positions(uint32_t t, Vec3 ps[3]) const { for (int k = 0; k < 3; ++k) ps[k] = positions[indices[t * 3 + k]]; }
    void build_bvh(bool literal) {
        if (bvh_built) return;
        std::vector<Bounds> ib(n_tris());
        for (uint32_t t = 0; t < n_tris(); ++t) {
            Vec3 ps[3];
            tri_positions(t, ps);
            const float inf = INFINITY;
            Vec3 mn(inf, inf, inf), mx(-inf, -inf, -inf);
            for (int k = 0; k < 3; ++k) { mn = vmin(mn, ps[k]); mx = vmax(mx, ps[k]); }
            ib[t] = Bounds{mn, mx};
        }
        bvh.build(ib, literal);
        bvh_built = true;
    }
};

struct LocalHit {
    float t_hit;
    uint32_t tri;
    Vec3 position, normal, shading_normal, tangent;
    Vec2 uv;
    float bary[3];
};

// Triangle::intersect (geometry/impls/triangle_mesh.rs:42-110), attribute part
inline void fill_local_hit(const Mesh& m, uint32_t tri, const TriangleHit& h, LocalHit* out) {
    out->t_hit = h.t_hit;
    out->tri = tri;
    out->position = h.position;
    out->normal = h.normal;
    for (int k = 0; k < 3; ++k) out->bary[k] = h.bary[k];
    uint32_t i0 = m.indices[tri * 3], i1 = m.indices[tri * 3 + 1], i2 = m.indices[tri * 3 + 2];
    out->shading_normal = make_normal(m.normals[i0] * h.bary[0] + m.normals[i1] * h.bary[1] + m.normals[i2] * h.bary[2]);
    if (m.uvs.empty()) { out->uv.x = 0; out->uv.y = 0; }
    else {
        Vec2 a = m.uvs[i0], b = m.uvs[i1], c = m.uvs[i2];
        out->uv.x = a.x * h.bary[0] + b.x * h.bary[1] + c.x * h.bary[2];
        out->uv.y = a.y * h.bary[0] + b.y * h.bary[1] + c.y * h.bary[2];
    }
    if (m.single) out->tangent = m.tangents[tri];  // SingleTriangle::intersect passes its tangent on without orthogonalising it
    else out->tangent = m.tangents.empty() ? Mesh::generate_tangent(out->shading_normal) : Mesh::orthogonalize(out->shading_normal, m.tangents[tri]);
}

struct SurfaceInteraction {
    Vec3 position, normal, shading_normal, tangent;
    Vec2 uv;
    int material = -1;
};
struct Intersection {
    float t_hit = 0;
    Vec3 wo;
    int primitive = -1;
    uint32_t tri = 0;
    float bary[3] = {0, 0, 0};
    SurfaceInteraction si;
};

// Transform::from_shading_normal_tangent (math/src/transform.rs:186-203)
inline Mat4 shading_transform(const SurfaceInteraction& si) {
    Vec3 n = normalize(si.shading_normal);
    Vec3 b = normalize(cross(normalize(n), si.tangent));
    Vec3 t = normalize(cross(b, n));
    return inverse(Mat4::from_cols3(t, b, n));
}

enum PrimitiveKind : int { PRIM_MESH = 0, PRIM_EMISSIVE_MESH = 1, PRIM_ENV_LIGHT = 2, PRIM_POINT_LIGHT = 3, PRIM_SPOT_LIGHT = 4, PRIM_DIRECTIONAL_LIGHT = 5 };

struct EnvLight {
    float intensity = 1.0f;
    uint32_t w = 0, h = 0;
    std::vector<float> data;  // h*w*3
    Spectrum integrated;
    std::vector<float> marginal_cdf;
    std::vector<float> conditional_cdf;  // h*w
    float total_weight = 0;
};

struct Primitive {
    int kind = PRIM_MESH;
    int geometry = -1;
    int material = -1;
    Mat4 local_to_world = Mat4::identity();
    Mat4 local_to_render = Mat4::identity();
    Mat4 render_to_local = Mat4::identity();  // cache of local_to_render.inverse() (same bits as the per-call inverse)
    // emissive mesh (emissive_triangle_mesh.rs:28-69)
    std::vector<float> area_list, area_table;
    float area_sum = 0;
    int env = -1;
    // delta lights (primitive/impls/{point,spot,directional}_light.rs)
    float light_intensity = 0.0f, angle_inner = 0.0f, angle_outer = 0.0f;
    Spectrum light_spectrum;
    float dir_area = 0.0f;  // DirectionalLight::preprocess: pi * r^2 of the scene's bounding sphere
    bool is_light() const { return kind != PRIM_MESH; }
    bool is_delta() const { return kind >= PRIM_POINT_LIGHT; }
};

struct LightSampler {
    std::vector<float> weights, table;
    float weight_sum = 0;
};

struct RayStats {
    uint64_t closest = 0, shadow = 0;
    TraversalCounters tc;
};

struct Scene {
    Tables T;
    std::vector<Mesh> meshes;
    std::vector<Texture> textures;
    std::vector<Material> materials;
    std::vector<Primitive> primitives;
    std::vector<EnvLight> envs;
    std::vector<int> light_list;  // LightSamplerFactory.light_list (light_sampler.rs:163-187)
    Bvh tlas;
    std::vector<int> tlas_items;  // item id -> primitive index
    bool faithful = true;         // reference-faithful cost model (per-call inverses, per-candidate attribute interpolation)
    bool literal_build = true;

    // Scene::build (scene.rs:64-76)
    void build(Vec3 cam_pos) {
        Mat4 world_to_render = Mat4::from_translation(-cam_pos);  // camera.rs:84-86
        for (auto& p : primitives) {
            p.local_to_render = mul(world_to_render, p.local_to_world);
            p.render_to_local = inverse(p.local_to_render);
        }
        for (auto& p : primitives) if (p.geometry >= 0) meshes[p.geometry].build_bvh(literal_build);
        tlas_items.clear();
        std::vector<Bounds> ib;
        for (size_t i = 0; i < primitives.size(); ++i) {
            if (primitives[i].geometry < 0) continue;
            tlas_items.push_back((int)i);
            const Mesh& gm = meshes[primitives[i].geometry];
            if (gm.single) {  // SingleTriangle::bounds (single_triangle.rs:86-94): box of the transformed vertices
                const float inf = INFINITY;
                Vec3 mn(inf, inf, inf), mx(-inf, -inf, -inf);
                for (int k = 0; k < 3; ++k) { Vec3 q = transform_point3(primitives[i].local_to_render, gm.positions[k]); mn = vmin(mn, q); mx = vmax(mx, q); }
                ib.push_back(Bounds{mn, mx});
            } else ib.push_back(transform_bounds(primitives[i].local_to_render, gm.bounds));
        }
        tlas.build(ib, literal_build);
        // LightSamplerFactory::build (light_sampler.rs:168-187): every light in primitive order, preprocess(scene_bounds) on the way
        light_list.clear();
        for (size_t i = 0; i < primitives.size(); ++i) {
            Primitive& p = primitives[i];
            if (!p.is_light()) continue;
            if (p.kind == PRIM_DIRECTIONAL_LIGHT) {  // directional_light.rs:85-89, Bounds::bounding_sphere (math/src/bounds.rs:73-77)
                Bounds sb = tlas.nodes.empty() ? Bounds{Vec3(0, 0, 0), Vec3(0, 0, 0)} : tlas.nodes[0].bounds;  // PrimitiveBvh::scene_bounds (primitive/bvh.rs:139-141)
                Vec3 center = (sb.mn + sb.mx) * 0.5f;
                float radius = length(sb.mx - center);
                p.dir_area = PI_F * radius * radius;
            }
            light_list.push_back((int)i);
        }
    }

    void init_emissive(Primitive& p) const {
        const Mesh& m = meshes[p.geometry];
        p.area_list.clear(); p.area_table.clear(); p.area_sum = 0;
        for (uint32_t t = 0; t < m.n_tris(); ++t) {
            Vec3 p0 = transform_point3(p.local_to_world, m.positions[m.indices[t * 3]]);
            Vec3 p1 = transform_point3(p.local_to_world, m.positions[m.indices[t * 3 + 1]]);
            Vec3 p2 = transform_point3(p.local_to_world, m.positions[m.indices[t * 3 + 2]]);
            Vec3 e0 = p0 - p1, e1 = p0 - p2;  // p1.vector_to(p0), p2.vector_to(p0)
            p.area_list.push_back(length(cross(e0, e1)) * 0.5f);
        }
        for (float a : p.area_list) { p.area_sum += a; p.area_table.push_back(p.area_sum); }
        for (float& a : p.area_table) a /= p.area_sum;
    }

    // primitive::TriangleMesh::intersect (primitive/impls/triangle_mesh.rs:89-118) + Transform * Intersection (primitive/bvh.rs:96-108, samples.rs:129-141)
    bool intersect_primitive(int prim_index, const Ray& ray, float t_max, Intersection* out, TraversalCounters* ctr) const {
        const Primitive& p = primitives[prim_index];
        const Mesh& m = meshes[p.geometry];
        Mat4 inv = faithful ? inverse(p.local_to_render) : p.render_to_local;
        Ray rl = transform_ray(inv, ray);
        LocalHit scratch;
        auto item_fn = [&](uint32_t tri, float tmax, float* t, uint64_t* payload) {
            Vec3 ps[3];
            m.tri_positions(tri, ps);
            TriangleHit th;
            if (!intersect_triangle(rl, tmax, ps, &th)) return false;
            if (faithful) fill_local_hit(m, tri, th, &scratch);
            *t = th.t_hit;
            *payload = tri;
            return true;
        };
        float t; uint64_t payload;
        if (m.single) {  // SingleTriangle::intersect (single_triangle.rs:100-141): the triangle test alone
            if (ctr) ctr->tri_tests++;
            if (!item_fn(0, t_max, &t, &payload)) return false;
        } else if (!m.bvh.intersect(rl, t_max, item_fn, &t, &payload, ctr)) return false;
        if (!out) return true;
        uint32_t tri = (uint32_t)payload;
        Vec3 ps[3];
        m.tri_positions(tri, ps);
        TriangleHit th;
        intersect_triangle(rl, t_max, ps, &th);
        LocalHit lh;
        fill_local_hit(m, tri, th, &lh);
        const Mat4& M = p.local_to_render;
        out->t_hit = lh.t_hit;
        out->wo = transform_vector3(M, -rl.d);
        out->primitive = prim_index;
        out->tri = tri;
        for (int k = 0; k < 3; ++k) out->bary[k] = lh.bary[k];
        out->si.position = transform_point3(M, lh.position);
        out->si.normal = transform_normal(M, lh.normal);
        out->si.shading_normal = transform_normal(M, lh.shading_normal);
        out->si.tangent = transform_vector3(M, lh.tangent);
        out->si.uv = lh.uv;
        out->si.material = p.material;
        return true;
    }
    bool intersect_p_primitive(int prim_index, const Ray& ray, float t_max, TraversalCounters* ctr) const {
        const Primitive& p = primitives[prim_index];
        const Mesh& m = meshes[p.geometry];
        Mat4 inv = faithful ? inverse(p.local_to_render) : p.render_to_local;
        Ray rl = transform_ray(inv, ray);
        auto item_fn = [&](uint32_t tri, float tmax) {
            Vec3 ps[3];
            m.tri_positions(tri, ps);
            return intersect_triangle(rl, tmax, ps, nullptr);
        };
        if (m.single) { if (ctr) ctr->tri_tests++; return item_fn(0, t_max); }
        return m.bvh.intersect_p(rl, t_max, item_fn, ctr);
    }

    // Scene::intersect (scene.rs:80-90): exhaustive closest hit through the TLAS
    bool intersect(const Ray& ray, float t_max, Intersection* out, RayStats* st) const {
        if (st) st->closest++;
        std::vector<Intersection> tmp;  // candidate storage indexed by payload
        tmp.reserve(4);
        auto item_fn = [&](uint32_t item, float tmax, float* t, uint64_t* payload) {
            Intersection is;
            if (!intersect_primitive(tlas_items[item], ray, tmax, &is, st ? &st->tc : nullptr)) return false;
            *t = is.t_hit;
            *payload = tmp.size();
            tmp.push_back(is);
            return true;
        };
        float t; uint64_t payload;
        if (!tlas.intersect(ray, t_max, item_fn, &t, &payload, st ? &st->tc : nullptr)) return false;
        *out = tmp[payload];
        return true;
    }
    // Scene::intersect_p (scene.rs:93-103)
    bool intersect_p(const Ray& ray, float t_max, RayStats* st) const {
        if (st) st->shadow++;
        auto item_fn = [&](uint32_t item, float tmax) { return intersect_p_primitive(tlas_items[item], ray, tmax, st ? &st->tc : nullptr); };
        return tlas.intersect_p(ray, t_max, item_fn, st ? &st->tc : nullptr);
    }

    // ---- emissive material (emissive_material.rs:48-80, edf.rs:20-31)
    SampledSpectrum emissive_radiance(const MaterialContext& c, const Material& m, Vec2 uv, const SampledWavelengths& wl) const {
        SampledSpectrum r = sample_spectrum_param(c, m.color, uv).sample(T, wl);
        float inten = sample_float_param(c, m.intensity, uv);
        return r * inten;
    }
    SampledSpectrum emissive_average_intensity(const MaterialContext& c, const Material& m, const SampledWavelengths& wl) const {
        Vec2 mid; mid.x = 0.5f; mid.y = 0.5f;
        SampledSpectrum r = sample_spectrum_param(c, m.color, mid).sample(T, wl);
        float inten = sample_float_param(c, m.intensity, mid);
        return r * inten;
    }

    // PrimitiveLight::phi
    SampledSpectrum phi(const MaterialContext& c, int prim_index, const SampledWavelengths& wl) const {
        const Primitive& p = primitives[prim_index];
        if (p.kind == PRIM_EMISSIVE_MESH) return emissive_average_intensity(c, materials[p.material], wl) * p.area_sum;  // emissive_triangle_mesh.rs:166-173
        if (p.kind == PRIM_POINT_LIGHT) return p.light_spectrum.sample(T, wl) * (4.0f * PI_F * p.light_intensity);  // point_light.rs:70-72: 4.0 * PI * intensity * spectrum
        if (p.kind == PRIM_SPOT_LIGHT) {  // spot_light.rs:84-95: intensity * spectrum * 2.0 * PI * (...), left to right
            float ci = std::cos(p.angle_inner), co = std::cos(p.angle_outer);
            return p.light_spectrum.sample(T, wl) * p.light_intensity * 2.0f * PI_F * ((1.0f - ci) + (ci - co) / 2.0f);
        }
        if (p.kind == PRIM_DIRECTIONAL_LIGHT) return p.light_spectrum.sample(T, wl) * (p.light_intensity * p.dir_area);  // directional_light.rs:80-83
        const EnvLight& e = envs[p.env];
        return e.intensity * e.integrated.sample(T, wl);  // environment_light.rs:299-301
    }
    // LightSamplerFactory::create (light_sampler.rs:190-220)
    LightSampler light_sampler(const MaterialContext& c, const SampledWavelengths& wl) const {
        LightSampler ls;
        for (int prim : light_list) {
            float w = phi(c, prim, wl).average();
            ls.weight_sum += w;
            ls.weights.push_back(w);
        }
        ls.table.assign(ls.weights.size(), 0.0f);
        float cum = 0.0f;
        for (size_t i = 0; i < ls.table.size(); ++i) { cum += ls.weights[i]; ls.table[i] = cum / ls.weight_sum; }
        return ls;
    }
    // LightSampler::sample_light (light_sampler.rs:26-44); returns index into light_list or -1
    int sample_light(const LightSampler& ls, float u, float* probability) const {
        if (ls.table.empty() || ls.weight_sum == 0.0f) return -1;
        for (size_t i = 0; i < ls.table.size(); ++i)
            if (u < ls.table[i]) { *probability = ls.weights[i] / ls.weight_sum; return (int)i; }
        size_t last = ls.table.size() - 1;
        *probability = ls.weights[last] / ls.weight_sum;
        return (int)last;
    }
    float light_probability(const LightSampler& ls, int prim_index) const {
        if (ls.table.empty() || ls.weight_sum == 0.0f) return 0.0f;
        for (size_t i = 0; i < light_list.size(); ++i) if (light_list[i] == prim_index) return ls.weights[i] / ls.weight_sum;
        return 0.0f;
    }
    // LightSampler::probability_infinite_light (light_sampler.rs:121-158)
    float light_probability_infinite(const LightSampler& ls, int prim_index) const {
        if (ls.table.empty() || ls.weight_sum == 0.0f) return 0.0f;
        float inf_sum = 0.0f;
        for (size_t i = 0; i < light_list.size(); ++i) if (primitives[light_list[i]].kind == PRIM_ENV_LIGHT) inf_sum += ls.weights[i];
        if (inf_sum == 0.0f) return 0.0f;
        for (size_t i = 0; i < light_list.size(); ++i)
            if (light_list[i] == prim_index) return primitives[prim_index].kind == PRIM_ENV_LIGHT ? ls.weights[i] / inf_sum : 0.0f;
        return 0.0f;
    }

    // ---- area light sampling (emissive_triangle_mesh.rs:176-309)
    struct AreaSample { SampledSpectrum radiance; float pdf, pdf_dir; Vec3 light_normal, position; };
    AreaSample sample_area_light(const MaterialContext& c, int prim_index, Vec3 shading_pos, const SampledWavelengths& wl, float s, Vec2 uv) const {
        const Primitive& p = primitives[prim_index];
        const Mesh& m = meshes[p.geometry];
        size_t index = 0;
        for (size_t i = 0; i < p.area_table.size(); ++i) if (s < p.area_table[i]) { index = i; break; }
        float b0, b1;
        if (uv.x < uv.y) { b0 = uv.x / 2.0f; b1 = uv.y - b0; } else { b1 = uv.y / 2.0f; b0 = uv.x - b1; }
        float b2 = 1.0f - b0 - b1;
        const Mat4& M = p.local_to_render;
        Vec3 p0 = transform_point3(M, m.positions[m.indices[index * 3]]);
        Vec3 p1 = transform_point3(M, m.positions[m.indices[index * 3 + 1]]);
        Vec3 p2 = transform_point3(M, m.positions[m.indices[index * 3 + 2]]);
        Vec3 pos = p0 * b0 + p1 * b1 + p2 * b2;
        Vec3 normal = make_normal(normalize(cross(p1 - p0, p2 - p0)));
        Vec2 luv; luv.x = 0; luv.y = 0;
        if (!m.uvs.empty()) {
            Vec2 a = m.uvs[m.indices[index * 3]], b = m.uvs[m.indices[index * 3 + 1]], cc = m.uvs[m.indices[index * 3 + 2]];
            luv.x = a.x * b0 + b.x * b1 + cc.x * b2;
            luv.y = a.y * b0 + b.y * b1 + cc.y * b2;
        }
        // EmissiveSingleTriangle::sample_radiance hands the two RANDOM NUMBERS on as the sample point's uv, not the interpolated
        // vertex uvs (emissive_single_triangle.rs:190-252: `uv` is the sample argument) -- visible with textured emission only.
        // Everything else of EmissiveSingleTriangle coincides with a one-triangle EmissiveTriangleMesh: area (:38-43, same bits as
        // emissive_triangle_mesh.rs:40-48 since cross(-a,-b) = cross(a,b)), pdf 1/area (:271,304), phi (:160-167).
        if (m.single) luv = uv;
        Vec3 wi = normalize(pos - shading_pos);
        AreaSample r;
        // UniformEdf: radiance is direction independent, so the light's tangent frame (:253-266) does not enter the value
        r.radiance = emissive_radiance(c, materials[p.material], luv, wl);
        r.pdf = 1.0f / p.area_sum;
        float distance = length(pos - shading_pos);
        r.pdf_dir = r.pdf * (distance * distance) / rmax(std::fabs(dot(normal, -wi)), 1e-8f);
        r.light_normal = normal;
        r.position = pos;
        return r;
    }
    // EmissiveTriangleMesh::pdf_light_sample (emissive_triangle_mesh.rs:331-353)
    float pdf_area_light_sample(int prim_index, uint32_t tri) const {
        const Primitive& p = primitives[prim_index];
        float probability = tri == 0 ? p.area_table[0] : p.area_table[tri] - p.area_table[tri - 1];
        return 1.0f / p.area_list[tri] * probability;
    }
    // Scene::pdf_light_sample (scene.rs:156-181)
    float pdf_light_sample(const LightSampler& ls, Vec3 shading_pos, const Intersection& is) const {
        const Primitive& p = primitives[is.primitive];
        if (p.kind != PRIM_EMISSIVE_MESH) return 0.0f;
        float probability = light_probability(ls, is.primitive);
        float pdf_area = pdf_area_light_sample(is.primitive, is.tri);
        Vec3 dv = shading_pos - is.si.position;
        float distance = length(dv);
        Vec3 wo = -normalize(dv);
        float pdf_dir = pdf_area * (distance * distance) / std::fabs(dot(is.si.normal, wo));
        return probability * pdf_dir;
    }

    // ---- environment light (environment_light.rs)
    void init_env(EnvLight& e) const {
        float tot[3] = {0, 0, 0};
        for (uint32_t y = 0; y < e.h; ++y) for (uint32_t x = 0; x < e.w; ++x) for (int c = 0; c < 3; ++c) tot[c] += e.data[((size_t)y * e.w + x) * 3 + c];
        float pc = (float)(e.w * e.h);
        for (int c = 0; c < 3; ++c) tot[c] /= pc;
        e.integrated = make_rgb_illuminant(T, Vec3(tot[0], tot[1], tot[2]), true);
        std::vector<float> row_weights(e.h, 0.0f);
        e.conditional_cdf.assign((size_t)e.w * e.h, 0.0f);
        for (uint32_t y = 0; y < e.h; ++y) {
            float row_sum = 0.0f;
            for (uint32_t x = 0; x < e.w; ++x) {
                float v = ((float)y + 0.5f) / (float)e.h;
                float theta = v * PI_F;
                const float* px = &e.data[((size_t)y * e.w + x) * 3];
                float lum = 0.299f * px[0] + 0.587f * px[1] + 0.114f * px[2];
                float weight = lum * rmax(std::sin(theta), 1e-8f);
                row_sum += weight;
                e.conditional_cdf[(size_t)y * e.w + x] = row_sum;
            }
            row_weights[y] = row_sum;
            if (row_sum > 0.0f) for (uint32_t x = 0; x < e.w; ++x) e.conditional_cdf[(size_t)y * e.w + x] /= row_sum;
        }
        float total = 0.0f;
        for (float r : row_weights) total += r;
        e.total_weight = total;
        e.marginal_cdf.assign(e.h, 0.0f);
        float cum = 0.0f;
        for (uint32_t y = 0; y < e.h; ++y) {
            cum += row_weights[y];
            e.marginal_cdf[y] = total > 0.0f ? cum / total : (float)(y + 1) / (float)e.h;
        }
    }
    static void direction_to_spherical(Vec3 d, float* theta, float* phi) {
        *theta = clampf(std::acos(d.y), 0.0f, PI_F);
        float p = std::atan2(d.z, d.x);
        if (p < 0.0f) p += 2.0f * PI_F;
        *phi = p;
    }
    static void env_sample_texture(const EnvLight& e, float u, float v, float out[3]) {
        u = clampf(u, 0.0f, 1.0f); v = clampf(v, 0.0f, 1.0f);
        float x = u * (float)(e.w - 1), y = v * (float)(e.h - 1);
        uint32_t x0 = f2u_sat(std::floor(x)), y0 = f2u_sat(std::floor(y));
        uint32_t x1 = std::min(x0 + 1, e.w - 1), y1 = std::min(y0 + 1, e.h - 1);
        float fx = x - (float)x0, fy = y - (float)y0;
        auto px = [&](uint32_t xx, uint32_t yy, int c) { return e.data[((size_t)yy * e.w + xx) * 3 + c]; };
        for (int c = 0; c < 3; ++c) {
            float p0 = px(x0, y0, c) * (1.0f - fx) + px(x1, y0, c) * fx;
            float p1 = px(x0, y1, c) * (1.0f - fx) + px(x1, y1, c) * fx;
            out[c] = p0 * (1.0f - fy) + p1 * fy;
        }
    }
    // Rust slice::binary_search_by(partial_cmp): the implementation details decide WHICH equal element is returned; CDFs with
    // exact duplicates only occur on zero-weight pixels, where either index yields the same direction weight 0 -> pdf 0.
    static size_t sample_from_cdf(const float* cdf, size_t n, float u) {
        // Rust 1.7x+ binary_search_by
        size_t size = n, base = 0;
        if (size == 0) return 0;
        while (size > 1) {
            size_t half = size / 2, mid = base + half;
            if (!(cdf[mid] > u)) base = mid;  // cmp != Greater -> base = mid
            size -= half;
        }
        if (cdf[base] == u) return base;
        size_t ins = base + (cdf[base] < u ? 1 : 0);
        return std::min(ins, n - 1);
    }
    float env_direction_pdf(const Primitive& p, Vec3 direction) const {
        const EnvLight& e = envs[p.env];
        if (e.total_weight <= 0.0f) return 0.0f;
        Vec3 dl = transform_vector3(faithful ? inverse(p.local_to_render) : p.render_to_local, direction);
        float theta, phi;
        direction_to_spherical(dl, &theta, &phi);
        float u = phi / (2.0f * PI_F), v = theta / PI_F;
        uint32_t x = std::min(f2u_sat(std::floor(u * (float)e.w)), e.w - 1);
        uint32_t y = std::min(f2u_sat(std::floor(v * (float)e.h)), e.h - 1);
        const float* px = &e.data[((size_t)y * e.w + x) * 3];
        float lum = 0.299f * px[0] + 0.587f * px[1] + 0.114f * px[2];
        float sin_theta = rmax(std::sin(theta), 1e-8f);
        float pdf_texture = lum * sin_theta / e.total_weight;
        float jac = (float)e.w * (float)e.h / (2.0f * PI_F * PI_F * sin_theta);
        return pdf_texture * jac;
    }
    SampledSpectrum env_direction_radiance(const Primitive& p, Vec3 dir, const SampledWavelengths& wl) const {
        const EnvLight& e = envs[p.env];
        Vec3 dl = transform_vector3(faithful ? inverse(p.local_to_render) : p.render_to_local, dir);
        float theta, phi;
        direction_to_spherical(dl, &theta, &phi);
        float rgb[3];
        env_sample_texture(e, phi / (2.0f * PI_F), theta / PI_F, rgb);
        Spectrum s = make_rgb_illuminant(T, Vec3(rgb[0], rgb[1], rgb[2]), true);
        return s.sample(T, wl) * e.intensity;
    }
    struct InfiniteSample { SampledSpectrum radiance; float pdf_dir; Vec3 wi; };
    InfiniteSample sample_infinite_light(int prim_index, const SampledWavelengths& wl, Vec2 uv) const {
        const Primitive& p = primitives[prim_index];
        const EnvLight& e = envs[p.env];
        size_t y = sample_from_cdf(e.marginal_cdf.data(), e.h, uv.x);
        size_t x = sample_from_cdf(&e.conditional_cdf[y * e.w], e.w, uv.y);
        float u = ((float)x + 0.5f) / (float)e.w, v = ((float)y + 0.5f) / (float)e.h;
        float theta = v * PI_F, phi = u * 2.0f * PI_F;
        Vec3 wl_local(std::sin(theta) * std::cos(phi), std::cos(theta), std::sin(theta) * std::sin(phi));
        InfiniteSample r;
        r.wi = transform_vector3(p.local_to_render, wl_local);
        r.pdf_dir = env_direction_pdf(p, r.wi);
        r.radiance = env_direction_radiance(p, r.wi, wl);
        return r;
    }
    // Scene::evaluate_infinite_light_radiance (scene.rs:213-231)
    SampledSpectrum evaluate_infinite_light_radiance(Vec3 dir, const SampledWavelengths& wl) const {
        SampledSpectrum tot = SampledSpectrum::zero();
        for (const auto& p : primitives) if (p.kind == PRIM_ENV_LIGHT) tot += env_direction_radiance(p, dir, wl);
        return tot;
    }
    // Scene::pdf_infinite_light_sample (scene.rs:184-210)
    float pdf_infinite_light_sample(const LightSampler& ls, Vec3 dir) const {
        float tot = 0.0f;
        for (size_t i = 0; i < primitives.size(); ++i)
            if (primitives[i].kind == PRIM_ENV_LIGHT) tot += light_probability_infinite(ls, (int)i) * env_direction_pdf(primitives[i], dir);
        return tot;
    }
};

}  // namespace orc
