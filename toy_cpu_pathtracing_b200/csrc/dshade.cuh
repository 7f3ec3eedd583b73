// Device shading: hit reconstruction, BSDFs and materials, light sampling, sensor.  One path vertex per thread.
// Follows (file:line in /root/reference):
//   hit reconstruction   math/src/ray.rs:161-175, scene/src/geometry/impls/triangle_mesh.rs:79-96, scene/src/primitive/bvh.rs:96-108
//   material helpers     scene/src/material/common.rs:7-139
//   Lambert              scene/src/material/bsdf/lambert.rs:11-120, impls/lambert_material.rs:42-179
//   Dielectric/Plastic   scene/src/material/bsdf/dielectric.rs:167-645, impls/plastic_material.rs:122-264
//   GeneralizedSchlick   scene/src/material/bsdf/generalized_schlick.rs:92-918 (ScatterMode::R, the only mode the materials use)
//   SimplePbr/Clearcoat  impls/simple_pbr_material.rs:78-537, impls/simple_pbr_clearcoat_material.rs:88-845
//   lights               scene/src/light_sampler.rs:17-220, primitive/impls/emissive_triangle_mesh.rs:166-353,
//                        primitive/impls/environment_light.rs:86-351, scene/src/scene.rs:107-231
//   sensor               renderer/src/sensor.rs:41-88
// Convention of the reference kept everywhere: every BSDF value `f` already contains the cosine |wi.z|.
#pragma once
#include "dcommon.cuh"

namespace tcpt {

enum : int { ST_DIFFUSE = 0, ST_SPECULAR_REFLECTION = 1, ST_SPECULAR_TRANSMISSION = 2, ST_GLOSSY_REFLECTION = 3, ST_GLOSSY_TRANSMISSION = 4 };

struct BsdfSample { S4 f; float3 wi; float pdf; int type; };
struct MatSample {
    S4 f; float3 wi; float pdf; int type; bool sampled;
    __device__ __forceinline__ bool is_specular() const { return type == ST_SPECULAR_REFLECTION || type == ST_SPECULAR_TRANSMISSION; }
};
__device__ __forceinline__ MatSample mat_fail() { MatSample m; m.f = s4(0.0f); m.wi = f3(0, 0, 1); m.pdf = 0.0f; m.type = ST_DIFFUSE; m.sampled = false; return m; }
__device__ __forceinline__ MatSample mat_ok(const S4& f, float3 wi, float pdf, int type) { MatSample m; m.f = f; m.wi = wi; m.pdf = pdf; m.type = type; m.sampled = true; return m; }

// ---------------------------------------------------------------- surface interaction in Render space
struct DSurface {
    float3 position, normal, shading_normal, tangent, wo;
    float2 uv;
    int material, prim;
    uint32_t tri;
};

__device__ __forceinline__ float3 orthogonalize(float3 n, float3 v) { const float pm = dot(n, v); return normalize(v - n * pm); }
__device__ __forceinline__ float3 generate_tangent(float3 n) { return orthogonalize(n, fabsf(n.x) > 0.999f ? f3(0, 1, 0) : f3(1, 0, 0)); }

__device__ __forceinline__ void reconstruct_hit(const DScene& sc, int prim, uint32_t tri, float b0, float b1, float b2, float3 ray_d, DSurface& s) {
    const tcpt_flat_primitive& P = sc.primitives[prim];
    const tcpt_flat_geometry& G = sc.geometries[P.geometry];
    const uint32_t* idx = sc.indices + 3 * ((size_t)G.index_base + tri);
    const uint32_t i0 = __ldg(idx) + G.vertex_base, i1 = __ldg(idx + 1) + G.vertex_base, i2 = __ldg(idx + 2) + G.vertex_base;
    const float* pp = sc.positions; const float* nn = sc.normals;
    const float3 p0 = f3(__ldg(pp + 3 * (size_t)i0), __ldg(pp + 3 * (size_t)i0 + 1), __ldg(pp + 3 * (size_t)i0 + 2));
    const float3 p1 = f3(__ldg(pp + 3 * (size_t)i1), __ldg(pp + 3 * (size_t)i1 + 1), __ldg(pp + 3 * (size_t)i1 + 2));
    const float3 p2 = f3(__ldg(pp + 3 * (size_t)i2), __ldg(pp + 3 * (size_t)i2 + 1), __ldg(pp + 3 * (size_t)i2 + 2));
    const float3 n0 = f3(__ldg(nn + 3 * (size_t)i0), __ldg(nn + 3 * (size_t)i0 + 1), __ldg(nn + 3 * (size_t)i0 + 2));
    const float3 n1 = f3(__ldg(nn + 3 * (size_t)i1), __ldg(nn + 3 * (size_t)i1 + 1), __ldg(nn + 3 * (size_t)i1 + 2));
    const float3 n2 = f3(__ldg(nn + 3 * (size_t)i2), __ldg(nn + 3 * (size_t)i2 + 1), __ldg(nn + 3 * (size_t)i2 + 2));
    const float3 pos_l = (p0 * b0 + p1 * b1) + p2 * b2;                      // ray.rs:161-165
    const float3 ng_l = normalize(normalize(cross(p1 - p0, p2 - p0)));       // ray.rs:168-174 (.normalize().to_normal())
    const float3 ns_l = normalize((n0 * b0 + n1 * b1) + n2 * b2);            // triangle_mesh.rs:79-83
    float3 tan_l;
    if (G.has_uv) {
        const float* uu = sc.uvs;
        const float2 a = make_float2(__ldg(uu + 2 * (size_t)i0), __ldg(uu + 2 * (size_t)i0 + 1));
        const float2 b = make_float2(__ldg(uu + 2 * (size_t)i1), __ldg(uu + 2 * (size_t)i1 + 1));
        const float2 c = make_float2(__ldg(uu + 2 * (size_t)i2), __ldg(uu + 2 * (size_t)i2 + 1));
        s.uv = make_float2((a.x * b0 + b.x * b1) + c.x * b2, (a.y * b0 + b.y * b1) + c.y * b2);
        const float* tt = sc.tangents + 3 * ((size_t)G.tangent_base + tri);
        tan_l = f3(__ldg(tt), __ldg(tt + 1), __ldg(tt + 2));
        if (!G.single) tan_l = orthogonalize(ns_l, tan_l);  // SingleTriangle hands its tangent on as computed (single_triangle.rs:118-135)
    } else {
        s.uv = make_float2(0.0f, 0.0f);
        tan_l = generate_tangent(ns_l);
    }
    // Transform * Intersection (primitive/bvh.rs:96-108, scene/src/samples.rs:129-141)
    const float3 d_l = xf_vector(P.r2l, ray_d);
    s.position = xf_point(P.l2r, pos_l);
    s.normal = xf_normal_by_inverse(P.r2l, ng_l);
    s.shading_normal = xf_normal_by_inverse(P.r2l, ns_l);
    s.tangent = xf_vector(P.l2r, tan_l);
    s.wo = xf_vector(P.l2r, -d_l);
    s.material = P.material; s.prim = prim; s.tri = tri;
}

// Transform::from_shading_normal_tangent (math/src/transform.rs:186-203): returns render->tangent and its inverse
__device__ __forceinline__ void shading_frame(const DSurface& s, M3& r2t, M3& t2r) {
    const float3 n = normalize(s.shading_normal);
    const float3 b = normalize(cross(normalize(n), s.tangent));
    const float3 t = normalize(cross(b, n));
    M3 m; m.c0 = t; m.c1 = b; m.c2 = n;
    r2t = m3_inverse(m);
    t2r = m3_inverse(r2t);
}

// ---------------------------------------------------------------- material/common.rs
__device__ __forceinline__ float cos2_theta(float3 w) { return w.z * w.z; }
__device__ __forceinline__ float tan2_theta(float3 w) { const float c2 = cos2_theta(w); return c2 == 0.0f ? TCPT_INF : (1.0f - c2) / c2; }
__device__ __forceinline__ float cos_phi(float3 w) { const float st = sqrtf(rmax(1.0f - cos2_theta(w), 0.0f)); return st == 0.0f ? 1.0f : clampf(w.x / st, -1.0f, 1.0f); }
__device__ __forceinline__ float sin_phi(float3 w) { const float st = sqrtf(rmax(1.0f - cos2_theta(w), 0.0f)); return st == 0.0f ? 0.0f : clampf(w.y / st, -1.0f, 1.0f); }
__device__ __forceinline__ bool half_vector(float3 wo, float3 wi, float3* wm) { const float3 m = wo + wi; if (length_squared(m) == 0.0f) return false; *wm = normalize(m); return true; }
__device__ __forceinline__ float3 reflect(float3 wo, float3 n) { return n * (2.0f * dot(wo, n)) - wo; }
__device__ __forceinline__ bool same_hemisphere(float3 a, float3 b) { return a.z * b.z > 0.0f; }
__device__ __forceinline__ float pow2(float x) { return x * x; }
__device__ __forceinline__ float pow6(float x) { const float x2 = x * x; const float x4 = x2 * x2; return x4 * x2; }
__device__ __forceinline__ float2 sample_uniform_disk_polar(float2 u) { const float r = sqrtf(u.x); const float th = 2.0f * TCPT_PI * u.y; return make_float2(r * cosf(th), r * sinf(th)); }
__device__ __noinline__ S4 fresnel_dielectric(float cos_theta_i, const S4& eta) {
    cos_theta_i = clampf(cos_theta_i, 0.0f, 1.0f);
    const float sin2_theta_i = 1.0f - cos_theta_i * cos_theta_i;
    const S4 sin2_t = s4(sin2_theta_i) / (eta * eta);
    const S4 cos_t = s4_sqrt(s4_clamp(s4(1.0f) - sin2_t, 0.0f, 1.0f));
    const S4 ci = s4(cos_theta_i);
    const S4 r_parl = (eta * ci - cos_t) / (eta * ci + cos_t);
    const S4 r_perp = (ci - eta * cos_t) / (ci + eta * cos_t);
    return (r_parl * r_parl + r_perp * r_perp) * 0.5f;
}
__device__ inline bool refract(float3 wi, float3 n, float eta, float3* wt) {
    const float cos_theta_i = dot(n, wi);
    const float sin2_theta_i = rmax(1.0f - cos_theta_i * cos_theta_i, 0.0f);
    const float sin2_theta_t = sin2_theta_i / (eta * eta);
    if (sin2_theta_t >= 1.0f) return false;
    const float cos_theta_t = sqrtf(rmax(1.0f - sin2_theta_t, 0.0f));
    const float3 t = (-wi) / eta + n * (cos_theta_i / eta - cos_theta_t);
    if (length_squared(t) < 1e-12f) return false;
    *wt = normalize(t);
    return true;
}

// ---------------------------------------------------------------- Trowbridge-Reitz helpers (dielectric.rs:22-120 == generalized_schlick.rs:96-190)
struct Ggx {
    float ax, ay;
    __device__ __forceinline__ bool effectively_smooth() const { return rmax(ax, ay) < 1e-3f; }
    __device__ __forceinline__ float D_i(float3 wm) const {
        const float t2 = tan2_theta(wm);
        if (!isfinite(t2)) return 0.0f;
        const float cos4 = pow2(cos2_theta(wm));
        const float e = t2 * (pow2(cos_phi(wm)) / pow2(ax) + pow2(sin_phi(wm)) / pow2(ay));
        return 1.0f / (TCPT_PI * ax * ay * cos4 * pow2(1.0f + e));
    }
    __device__ __noinline__ static float D_v(const Ggx self, float3 wm) { return self.D_i(wm); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ float D(float3 wm) const { return D_v(*this, wm); }
    __device__ __forceinline__ float lambda_i(float3 w) const {
        const float t2 = tan2_theta(w);
        if (isinf(t2)) return 0.0f;
        const float a2 = pow2(cos_phi(w) * ax) + pow2(sin_phi(w) * ay);
        return (sqrtf(1.0f + a2 * t2) - 1.0f) / 2.0f;
    }
    __device__ __noinline__ static float lambda_v(const Ggx self, float3 w) { return self.lambda_i(w); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ float lambda(float3 w) const { return lambda_v(*this, w); }
    __device__ float G1(float3 w) const { return 1.0f / (1.0f + lambda(w)); }
    __device__ float G(float3 wo, float3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    __device__ float Dvis(float3 w, float3 wm) const {
        const float c = fabsf(w.z);
        if (c == 0.0f) return 0.0f;
        return G1(w) / c * D(wm) * fabsf(dot(w, wm));
    }
    __device__ __forceinline__ float3 sample_wm_i(float3 w, float2 u) const {
        float3 wh = normalize(f3(ax * w.x, ay * w.y, w.z));
        if (wh.z < 0.0f) wh = -wh;
        const float3 t1 = wh.z < 0.99999f ? normalize(cross(f3(0, 0, 1), wh)) : f3(1, 0, 0);
        const float3 t2 = cross(wh, t1);
        const float2 p = sample_uniform_disk_polar(u);
        const float h = sqrtf(rmax(1.0f - p.x * p.x, 0.0f));
        const float lf = (1.0f + wh.z) / 2.0f;
        const float py = h * (1.0f - lf) + p.y * lf;
        const float pz = sqrtf(rmax(1.0f - p.x * p.x - py * py, 0.0f));
        const float3 nh = (t1 * p.x + t2 * py) + wh * pz;
        return normalize(f3(ax * nh.x, ay * nh.y, rmax(1e-6f, nh.z)));
    }
    __device__ __noinline__ static float3 sample_wm_v(const Ggx self, float3 w, float2 u) { return self.sample_wm_i(w, u); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ float3 sample_wm(float3 w, float2 u) const { return sample_wm_v(*this, w, u); }
};

__device__ inline bool generalized_half_vector(float3 wo, float3 wi, float eta, float3* out) {
    const float co = wo.z, ci = wi.z;
    const bool refl = ci * co > 0.0f;
    const float etap = !refl ? (co > 0.0f ? eta : 1.0f / eta) : 1.0f;
    float3 wm = wi * etap + wo;
    if (ci == 0.0f || co == 0.0f || length_squared(wm) == 0.0f) return false;
    wm = normalize(wm);
    if (wm.z < 0.0f) wm = -wm;
    if (dot(wm, wi) * ci < 0.0f || dot(wm, wo) * co < 0.0f) return false;
    *out = wm;
    return true;
}

// ---------------------------------------------------------------- bsdf/lambert.rs
__device__ __forceinline__ bool lambert_sample(const S4& albedo, float3 wo, float2 uv, BsdfSample* out) {
    const float wo_cos = wo.z;
    if (wo_cos == 0.0f) return false;
    const float r = sqrtf(uv.x), th = 2.0f * TCPT_PI * uv.y;
    float3 wi = f3(r * cosf(th), r * sinf(th), sqrtf(1.0f - uv.x));
    if (wo_cos < 0.0f) wi.z = -wi.z;
    const float wi_cos = wi.z;
    if (wi_cos == 0.0f) return false;
    if (signum(wo_cos) != signum(wi_cos)) return false;
    out->f = albedo * fabsf(wi_cos) / TCPT_PI;
    out->pdf = fabsf(wi_cos) / TCPT_PI;
    out->wi = wi;
    out->type = ST_DIFFUSE;
    return true;
}
__device__ __forceinline__ bool lambert_ok(float3 wo, float3 wi) { return !(wo.z == 0.0f || wi.z == 0.0f) && signum(wo.z) == signum(wi.z); }
__device__ __forceinline__ S4 lambert_eval(const S4& albedo, float3 wo, float3 wi) { return lambert_ok(wo, wi) ? albedo * fabsf(wi.z) / TCPT_PI : s4(0.0f); }
__device__ __forceinline__ float lambert_pdf(float3 wo, float3 wi) { return lambert_ok(wo, wi) ? fabsf(wi.z) / TCPT_PI : 0.0f; }

// ---------------------------------------------------------------- bsdf/dielectric.rs
struct Dielectric {
    S4 eta; bool entering, thin; Ggx g;
    __device__ __forceinline__ S4 eta_spectrum() const { return (thin || entering) ? eta : s4(1.0f) / eta; }
    __device__ __forceinline__ static void thin_coeffs(float fresnel, float* pr, float* pt) {
        float r = fresnel; const float t = 1.0f - r, r2 = r * r;
        r = r2 > 1.0f ? 1.0f : r + (t * t * r) / (1.0f - r2);
        *pr = r; *pt = t;
    }
    __device__ bool sample_specular(float3 wo, float ux, DWavelengths& wl, BsdfSample* out) const {
        const float wo_cos = wo.z;
        const float3 n = entering ? f3(0, 0, 1) : f3(0, 0, -1);
        const S4 es = eta_spectrum();
        const float etap = es.v[0];
        const S4 fresnel = fresnel_dielectric(fabsf(wo_cos), es);
        float pr, pt;
        if (thin) thin_coeffs(s4_avg(fresnel), &pr, &pt);
        else { pr = s4_avg(fresnel); pt = 1.0f - pr; }
        if (ux < pr / (pr + pt)) {
            if (fabsf(wo_cos) < 1e-6f) return false;
            out->f = fresnel; out->wi = f3(-wo.x, -wo.y, wo.z); out->pdf = pr / (pr + pt); out->type = ST_SPECULAR_REFLECTION;
            return true;
        }
        if (thin) {
            const float3 wi = f3(-wo.x, -wo.y, -wo.z);
            if (wi.z == 0.0f) return false;
            out->f = s4(1.0f) - fresnel; out->wi = wi; out->pdf = pt / (pr + pt); out->type = ST_SPECULAR_TRANSMISSION;
            return true;
        }
        if (!s4_is_constant(eta) && !wl.terminated) wl = wavelengths_uniform(wl.lambda[0], true);
        float3 wt;
        if (!refract(wo, n, etap, &wt)) return false;
        if (wt.z == 0.0f) return false;
        out->f = (s4(1.0f) - fresnel) / pow2(etap); out->wi = wt; out->pdf = pt / (pr + pt); out->type = ST_SPECULAR_TRANSMISSION;
        return true;
    }
    __device__ bool mf_reflection(float3 wo, float3 wm, const S4& fresnel, float prob, BsdfSample* out) const {
        const float3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return false;
        const float cd = fabsf(dot(wo, wm));
        if (cd < 1e-6f) return false;
        out->pdf = g.Dvis(wo, wm) / (4.0f * cd) * prob;
        const float d = g.D(wm), gg = g.G(wo, wi);
        // quirk: an extra |wi.z| here that evaluate() does not have (dielectric.rs:318 vs :589)
        out->f = fresnel * d * gg * fabsf(wi.z) / (4.0f * fabsf(wo.z));
        out->wi = wi; out->type = ST_GLOSSY_REFLECTION;
        return true;
    }
    __device__ bool mf_transmission(float3 wo, float3 wm, const S4& tr, float prob, float etap, BsdfSample* out) const {
        const float3 wmr = entering ? wm : -wm;
        float3 wi;
        if (!refract(wo, wmr, etap, &wi)) return false;
        if (same_hemisphere(wo, wi) || fabsf(wi.z) == 0.0f) return false;
        const float denom = pow2(dot(wi, wm) + dot(wo, wm) / etap);
        const float dwm_dwi = fabsf(dot(wi, wm)) / denom;
        out->pdf = g.Dvis(wo, wm) * dwm_dwi * prob;
        const float d = g.D(wm), gg = g.G(wo, wi);
        out->f = tr * d * gg * fabsf(dot(wi, wm)) * fabsf(dot(wo, wm)) / (denom * fabsf(wo.z) * etap * etap);
        out->wi = wi; out->type = ST_GLOSSY_TRANSMISSION;
        return true;
    }
    __device__ __forceinline__ bool sample_i(float3 wo, float2 uv, float uc, DWavelengths& wl, BsdfSample* out) const {
        if (wo.z == 0.0f) return false;
        if (g.effectively_smooth()) return sample_specular(wo, uc, wl, out);  // selector = uc (dielectric.rs:180)
        const float3 wm = g.sample_wm(wo, uv);
        const S4 es = eta_spectrum();
        const float eta_scalar = es.v[0];
        const S4 fresnel = fresnel_dielectric(fabsf(dot(wo, wm)), es);
        const float pr = s4_avg(fresnel), pt = 1.0f - pr;
        if (thin) {
            float tpr, tpt;
            thin_coeffs(s4_avg(fresnel), &tpr, &tpt);
            if (uc < tpr / (tpr + tpt)) return mf_reflection(wo, wm, fresnel, tpr / (tpr + tpt), out);
            out->f = s4(1.0f) - fresnel; out->wi = f3(-wo.x, -wo.y, -wo.z); out->pdf = tpt / (tpr + tpt); out->type = ST_GLOSSY_TRANSMISSION;
            return true;
        } else if (uc < pr / (pr + pt)) {
            return mf_reflection(wo, wm, fresnel, pr / (pr + pt), out);
        }
        if (!s4_is_constant(eta) && !wl.terminated) wl = wavelengths_uniform(wl.lambda[0], true);
        return mf_transmission(wo, wm, s4(1.0f) - fresnel, pt / (pr + pt), eta_scalar, out);
    }
    __device__ __noinline__ static bool sample_v(const Dielectric self, float3 wo, float2 uv, float uc, DWavelengths& wl, BsdfSample* out) { return self.sample_i(wo, uv, uc, wl, out); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ bool sample(float3 wo, float2 uv, float uc, DWavelengths& wl, BsdfSample* out) const { return sample_v(*this, wo, uv, uc, wl, out); }
    __device__ __forceinline__ S4 evaluate_i(float3 wo, float3 wi) const {
        if (g.effectively_smooth()) return s4(0.0f);
        const S4 es = eta_spectrum();
        const float eta_scalar = es.v[0];
        float3 wm;
        if (!generalized_half_vector(wo, wi, eta_scalar, &wm)) return s4(0.0f);
        const S4 fresnel = fresnel_dielectric(fabsf(dot(wo, wm)), es);
        const bool refl = wi.z * wo.z > 0.0f;
        const float d = g.D(wm), gg = g.G(wo, wi);
        if (refl) return fresnel * d * gg / (4.0f * fabsf(wo.z));
        const float denom = pow2(dot(wi, wm) + dot(wo, wm) / eta_scalar);
        return (s4(1.0f) - fresnel) * d * gg * fabsf(dot(wi, wm)) * fabsf(dot(wo, wm)) / (denom * fabsf(wo.z) * eta_scalar * eta_scalar);
    }
    __device__ __noinline__ static S4 evaluate_v(const Dielectric self, float3 wo, float3 wi) { return self.evaluate_i(wo, wi); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ S4 evaluate(float3 wo, float3 wi) const { return evaluate_v(*this, wo, wi); }
    __device__ __forceinline__ float pdf_i(float3 wo, float3 wi) const {
        if (g.effectively_smooth()) return 0.0f;
        const S4 es = eta_spectrum();
        const float eta_scalar = es.v[0];
        float3 wm;
        if (!generalized_half_vector(wo, wi, eta_scalar, &wm)) return 0.0f;
        const S4 fresnel = fresnel_dielectric(fabsf(dot(wo, wm)), es);
        const float pr = s4_avg(fresnel), pt = 1.0f - pr;
        const bool refl = wi.z * wo.z > 0.0f;
        if (refl) return g.Dvis(wo, wm) / (4.0f * fabsf(dot(wo, wm))) * pr / (pr + pt);
        if (thin) return pt / (pr + pt);
        const float denom = pow2(dot(wi, wm) + dot(wo, wm) / eta_scalar);
        const float dwm_dwi = fabsf(dot(wi, wm)) / denom;
        return g.Dvis(wo, wm) * dwm_dwi * pt / (pr + pt);
    }
    __device__ __noinline__ static float pdf_v(const Dielectric self, float3 wo, float3 wi) { return self.pdf_i(wo, wi); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ float pdf(float3 wo, float3 wi) const { return pdf_v(*this, wo, wi); }
};
// DielectricBsdf::new (dielectric.rs:127-148): an eta whose first lane is 0 falls back to the constant 1
__device__ __forceinline__ Dielectric make_dielectric(const S4& eta, bool entering, bool thin, float alpha) {
    Dielectric d; d.eta = eta.v[0] == 0.0f ? s4(1.0f) : eta; d.entering = entering; d.thin = thin; d.g.ax = alpha; d.g.ay = alpha; return d;
}
__device__ __forceinline__ Dielectric make_dielectric(float eta, bool entering, bool thin, float alpha) { return make_dielectric(s4(eta), entering, thin, alpha); }

// ---------------------------------------------------------------- bsdf/conductor.rs
struct Cplx { float re, im; };
__device__ __forceinline__ Cplx cadd(Cplx a, Cplx b) { return Cplx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return Cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return Cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx cscale(Cplx a, float s) { return Cplx{a.re * s, a.im * s}; }
__device__ __forceinline__ Cplx cdiv(Cplx a, Cplx b) {
    const float denom = b.re * b.re + b.im * b.im;
    if (denom == 0.0f) return Cplx{0.0f, 0.0f};
    return Cplx{(a.re * b.re + a.im * b.im) / denom, (a.im * b.re - a.re * b.im) / denom};
}
__device__ __forceinline__ Cplx csqrt(Cplx a) {  // polar form (conductor.rs:30-36)
    const float r = sqrtf(a.re * a.re + a.im * a.im);
    const float theta = atan2f(a.im, a.re);
    const float sr = sqrtf(r), ht = theta * 0.5f;
    return Cplx{sr * cosf(ht), sr * sinf(ht)};
}
__device__ __forceinline__ float cnorm(Cplx a) { return a.re * a.re + a.im * a.im; }
// fresnel_complex (conductor.rs:91-123)
__device__ __noinline__ S4 fresnel_complex(float cos_theta_i, const S4& eta, const S4& k) {
    cos_theta_i = clampf(cos_theta_i, 0.0f, 1.0f);
    S4 r;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const Cplx ce{eta.v[i], k.v[i]};
        const float sin2_i = 1.0f - cos_theta_i * cos_theta_i;
        const Cplx sin2_t = cdiv(Cplx{sin2_i, 0.0f}, cmul(ce, ce));
        const Cplx cos_t = csqrt(csub(Cplx{1.0f, 0.0f}, sin2_t));
        const Cplx r_parl = cdiv(csub(cscale(ce, cos_theta_i), cos_t), cadd(cscale(ce, cos_theta_i), cos_t));
        const Cplx r_perp = cdiv(csub(Cplx{cos_theta_i, 0.0f}, cmul(ce, cos_t)), cadd(Cplx{cos_theta_i, 0.0f}, cmul(ce, cos_t)));
        r.v[i] = (cnorm(r_parl) + cnorm(r_perp)) * 0.5f;
    }
    return r;
}
struct Conductor {  // conductor.rs:125-439
    S4 eta, k; Ggx g;
    __device__ S4 torrance_sparrow(float3 wo, float3 wi, float3 wm) const {
        const float co = fabsf(wo.z), ci = fabsf(wi.z);
        if (co == 0.0f || ci == 0.0f) return s4(0.0f);
        const S4 fr = fresnel_complex(fabsf(dot(wo, wm)), eta, k);
        const float d = g.D(wm), gg = g.G(wo, wi);
        return fr * d * gg / (4.0f * co);
    }
    __device__ float pdf_microfacet(float3 wo, float3 wi) const {
        if (!same_hemisphere(wo, wi)) return 0.0f;
        float3 wm;
        if (!half_vector(wo, wi, &wm)) return 0.0f;
        const float vis = g.Dvis(wo, wm);
        const float jac = 4.0f * fabsf(dot(wo, wm));
        if (jac == 0.0f) return 0.0f;
        return vis / jac;
    }
    __device__ bool sample(float3 wo, float2 uv, BsdfSample* out) const {
        if (wo.z == 0.0f) return false;
        if (g.effectively_smooth()) {
            const float3 wi = f3(-wo.x, -wo.y, wo.z);
            if (wi.z == 0.0f) return false;
            out->f = fresnel_complex(fabsf(wi.z), eta, k); out->wi = wi; out->pdf = 1.0f; out->type = ST_SPECULAR_REFLECTION;
            return true;
        }
        const float3 wm = g.sample_wm(wo, uv);
        const float3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return false;
        out->f = torrance_sparrow(wo, wi, wm); out->wi = wi; out->pdf = pdf_microfacet(wo, wi); out->type = ST_GLOSSY_REFLECTION;
        return true;
    }
    __device__ S4 evaluate(float3 wo, float3 wi) const {
        if (g.effectively_smooth()) return s4(0.0f);
        const float co = fabsf(wo.z), ci = fabsf(wi.z);
        if (co == 0.0f || ci == 0.0f) return s4(0.0f);
        if (!same_hemisphere(wo, wi)) return s4(0.0f);
        float3 wm;
        if (!half_vector(wo, wi, &wm)) return s4(0.0f);
        return torrance_sparrow(wo, wi, wm);
    }
    __device__ float pdf(float3 wo, float3 wi) const { return g.effectively_smooth() ? 0.0f : pdf_microfacet(wo, wi); }
};

// ---------------------------------------------------------------- bsdf/generalized_schlick.rs (ScatterMode::R)
// The materials only ever build it with r90 = 1, exponent = 5, tint = 1 (simple_pbr_material.rs:290-520,
// simple_pbr_clearcoat_material.rs:171-188, 580-829); the Lazanyi term is kept because (1 - tint) = 0 must still multiply through.
struct Schlick {
    S4 r0; Ggx g;
    // fresnel (generalized_schlick.rs:192-228) with r90 = 1, exponent = 5, tint = 1: the Lazanyi correction is
    // a = f(cos_max) * (1 - tint) / (...) = +-0 for every finite r0, and `base - 0 * x` == base, so only the Schlick term is evaluated.
    __device__ S4 fresnel_at(float cos_theta) const {
        cos_theta = clampf(cos_theta, 0.0f, 1.0f);
        const float omc = 1.0f - cos_theta;
        const float o2 = omc * omc;
        return r0 + (s4(1.0f) - r0) * (o2 * o2 * omc);  // (1 - cos)^5: three roundings, within 2 ulp of powf(omc, 5.0)
    }
    __device__ __forceinline__ bool sample_i(float3 wo, float2 uv, BsdfSample* out) const {
        if (wo.z == 0.0f) return false;
        if (g.effectively_smooth()) {
            const float3 wi = f3(-wo.x, -wo.y, wo.z);
            if (wi.z == 0.0f) return false;
            out->f = fresnel_at(fabsf(wo.z)); out->wi = wi; out->pdf = 1.0f; out->type = ST_SPECULAR_REFLECTION;
            return true;
        }
        const float3 wm = g.sample_wm(wo, uv);
        const S4 fr = fresnel_at(fabsf(dot(wo, wm)));
        const float3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return false;
        const float cd = fabsf(dot(wo, wm));
        if (cd < 1e-6f) return false;
        const float pdf = g.Dvis(wo, wm) / (4.0f * cd) * 1.0f;
        const float d = g.D(wm), gg = g.G(wo, wi);
        const float ci = fabsf(wi.z), co = fabsf(wo.z);
        if (ci == 0.0f || co == 0.0f) return false;
        out->f = fr * d * gg / (4.0f * co); out->wi = wi; out->pdf = pdf; out->type = ST_GLOSSY_REFLECTION;
        return true;
    }
    __device__ __noinline__ static bool sample_v(const Schlick self, float3 wo, float2 uv, BsdfSample* out) { return self.sample_i(wo, uv, out); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ bool sample(float3 wo, float2 uv, BsdfSample* out) const { return sample_v(*this, wo, uv, out); }
    __device__ __forceinline__ S4 evaluate_i(float3 wo, float3 wi) const {
        if (g.effectively_smooth()) return s4(0.0f);
        const float co = fabsf(wo.z), ci = fabsf(wi.z);
        if (co == 0.0f || ci == 0.0f) return s4(0.0f);
        if (!same_hemisphere(wo, wi)) return s4(0.0f);
        float3 wm;
        if (!half_vector(wo, wi, &wm)) return s4(0.0f);
        const S4 fr = fresnel_at(fabsf(dot(wo, wm)));
        const float d = g.D(wm), gg = g.G(wo, wi);
        return fr * d * gg / (4.0f * co);
    }
    __device__ __noinline__ static S4 evaluate_v(const Schlick self, float3 wo, float3 wi) { return self.evaluate_i(wo, wi); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ S4 evaluate(float3 wo, float3 wi) const { return evaluate_v(*this, wo, wi); }
    __device__ __forceinline__ float pdf_i(float3 wo, float3 wi) const {
        if (g.effectively_smooth()) return 0.0f;
        if (!same_hemisphere(wo, wi)) return 0.0f;
        float3 wm;
        if (!half_vector(wo, wi, &wm)) return 0.0f;
        const float vis = g.Dvis(wo, wm);
        const float jac = 4.0f * fabsf(dot(wo, wm));
        if (jac == 0.0f) return 0.0f;
        return vis / jac;
    }
    __device__ __noinline__ static float pdf_v(const Schlick self, float3 wo, float3 wi) { return self.pdf_i(wo, wi); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ float pdf(float3 wo, float3 wi) const { return pdf_v(*this, wo, wi); }
    // 64-sample stochastic estimate (generalized_schlick.rs:893-918); `f` already holds a cosine, reproduced as is
    __device__ __forceinline__ S4 directional_albedo_i(float3 wo, DAuxRng rng) const {
        S4 sum = s4(0.0f);
        if (g.effectively_smooth()) {
            // every one of the 64 samples is the same mirror sample (the random numbers are drawn but unused), so the term is
            // computed once and ADDED 64 times: the running sum rounds exactly like the reference's loop
            BsdfSample s;
            S4 term = s4(0.0f);
            if (sample(wo, make_float2(0.0f, 0.0f), &s)) {
                const float ci = fabsf(s.wi.z);
                if (ci > 0.0f && s.pdf > 0.0f) term = s.f * ci / s.pdf;
            }
            for (int i = 0; i < 64; ++i) sum = sum + term;
            return sum / 64.0f;
        }
        // rough coat: 64 calls of sample(wo, uv) (generalized_schlick.rs:893-918 -> :419-458).  Everything that only depends on
        // wo -- the stretched direction and its frame in sample_wm, Lambda(wo) inside G1 and G, |wo.z| -- is evaluated once;
        // D(wm), which sample() evaluates inside Dvis and again for f, is evaluated once per sample.  Same expressions in the
        // same order as Ggx::sample_wm / Dvis / G, so every term has the bits of the straightforward loop.
        if (wo.z == 0.0f) return sum / 64.0f;
        float3 wh = normalize(f3(g.ax * wo.x, g.ay * wo.y, wo.z));
        if (wh.z < 0.0f) wh = -wh;
        const float3 t1 = wh.z < 0.99999f ? normalize(cross(f3(0, 0, 1), wh)) : f3(1, 0, 0);
        const float3 t2 = cross(wh, t1);
        const float lf = (1.0f + wh.z) / 2.0f;
        const float lam_o = g.lambda(wo), g1_o = 1.0f / (1.0f + lam_o), co = fabsf(wo.z);
        for (int i = 0; i < 64; ++i) {
            rng.next();  // uc (unused by the R-mode sampler, but drawn by the reference)
            float2 uv; uv.x = rng.next(); uv.y = rng.next();
            const float2 p = sample_uniform_disk_polar(uv);
            const float h = sqrtf(rmax(1.0f - p.x * p.x, 0.0f));
            const float py = h * (1.0f - lf) + p.y * lf;
            const float pz = sqrtf(rmax(1.0f - p.x * p.x - py * py, 0.0f));
            const float3 nh = (t1 * p.x + t2 * py) + wh * pz;
            const float3 wm = normalize(f3(g.ax * nh.x, g.ay * nh.y, rmax(1e-6f, nh.z)));
            const float cd = fabsf(dot(wo, wm));
            const S4 fr = fresnel_at(cd);
            const float3 wi = reflect(wo, wm);
            if (!same_hemisphere(wo, wi)) continue;
            if (cd < 1e-6f) continue;
            const float d = g.D(wm);
            const float dvis = co == 0.0f ? 0.0f : g1_o / co * d * cd;
            const float pdf = dvis / (4.0f * cd) * 1.0f;
            const float gg = 1.0f / (1.0f + lam_o + g.lambda(wi));
            const float ci = fabsf(wi.z);
            if (ci == 0.0f || co == 0.0f) continue;
            const S4 f = fr * d * gg / (4.0f * co);
            if (ci > 0.0f && pdf > 0.0f) sum = sum + f * ci / pdf;
        }
        return sum / 64.0f;
    }
    __device__ __noinline__ static S4 directional_albedo_v(const Schlick self, float3 wo, DAuxRng rng) { return self.directional_albedo_i(wo, rng); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ S4 directional_albedo(float3 wo, DAuxRng rng) const { return directional_albedo_v(*this, wo, rng); }
};
__device__ __forceinline__ float r0_of(float ior) { const float r = (ior - 1.0f) / (ior + 1.0f); return r * r; }
__device__ __forceinline__ Schlick make_schlick(const S4& r0, float alpha) { Schlick s; s.r0 = r0; s.g.ax = alpha; s.g.ay = alpha; return s; }

// ---------------------------------------------------------------- SimplePbr (simple_pbr_material.rs:274-537 == simple_pbr_clearcoat_material.rs:540-845)
struct PbrBase {
    S4 base_color; float metallic, roughness, ior;
    // sample*: `wi` comes back in the NORMAL-MAP frame; the caller rotates it into the shading frame (pbr_sample below).  Taking the
    // matrix by reference here kept it in the caller's stack frame: 9 stores and 27+ local loads per vertex.
    __device__ MatSample sample_metallic(float alpha, float3 wo, float2 uv) const {
        BsdfSample s;
        if (!make_schlick(base_color, alpha).sample(wo, uv, &s)) return mat_fail();
        return mat_ok(s.f, s.wi, s.pdf, s.type);
    }
    __device__ MatSample sample_dielectric(float alpha, float3 wo, float uc, float2 uv) const {
        const Schlick gs = make_schlick(s4(r0_of(ior)), alpha);
        const float fresnel = s4_avg(gs.fresnel_at(fabsf(wo.z)));
        BsdfSample s;
        if (uc < fresnel) {
            if (!gs.sample(wo, uv, &s)) return mat_fail();
            return mat_ok(s.f, s.wi, s.pdf * fresnel, s.type);
        }
        if (!lambert_sample(base_color, wo, uv, &s)) return mat_fail();
        return mat_ok(s.f * (1.0f - fresnel), s.wi, s.pdf * (1.0f - fresnel), s.type);
    }
    __device__ __forceinline__ MatSample sample_nm_i(float3 wo, float uc, float2 uv) const {
        const float alpha = roughness * roughness;
        if (metallic >= 1.0f) return sample_metallic(alpha, wo, uv);
        if (metallic <= 0.0f) return sample_dielectric(alpha, wo, uc, uv);
        if (uc <= metallic) return sample_metallic(alpha, wo, uv);
        return sample_dielectric(alpha, wo, (uc - metallic) / (1.0f - metallic), uv);
    }
    __device__ __noinline__ static MatSample sample_nm_v(const PbrBase self, float3 wo, float uc, float2 uv) { return self.sample_nm_i(wo, uc, uv); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ MatSample sample_nm(float3 wo, float uc, float2 uv) const { return sample_nm_v(*this, wo, uc, uv); }
    __device__ __forceinline__ MatSample sample(float3 wo, float uc, float2 uv, const M3& from_nm) const {
        MatSample ms = sample_nm(wo, uc, uv);
        if (ms.sampled) ms.wi = m3_vector(from_nm, ms.wi);
        return ms;
    }
    __device__ S4 eval_dielectric(float alpha, float3 wo, float3 wi) const {
        const Schlick gs = make_schlick(s4(r0_of(ior)), alpha);
        const S4 direct = gs.evaluate(wo, wi);
        const float fresnel = s4_avg(gs.fresnel_at(fabsf(wo.z)));
        return direct + (1.0f - fresnel) * lambert_eval(base_color, wo, wi);
    }
    __device__ __forceinline__ S4 evaluate_i(float3 wo, float3 wi) const {
        const float alpha = roughness * roughness;
        if (metallic >= 1.0f) return make_schlick(base_color, alpha).evaluate(wo, wi);
        if (metallic <= 0.0f) return eval_dielectric(alpha, wo, wi);
        return make_schlick(base_color, alpha).evaluate(wo, wi) * metallic + eval_dielectric(alpha, wo, wi) * (1.0f - metallic);
    }
    __device__ __noinline__ static S4 evaluate_v(const PbrBase self, float3 wo, float3 wi) { return self.evaluate_i(wo, wi); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ S4 evaluate(float3 wo, float3 wi) const { return evaluate_v(*this, wo, wi); }
    __device__ float pdf_dielectric(float alpha, float3 wo, float3 wi) const {
        const Schlick gs = make_schlick(s4(r0_of(ior)), alpha);
        const float direct = gs.pdf(wo, wi);
        const float fresnel = s4_avg(gs.fresnel_at(fabsf(wo.z)));
        return fresnel * direct + (1.0f - fresnel) * lambert_pdf(wo, wi);
    }
    __device__ __forceinline__ float pdf_i(float3 wo, float3 wi) const {
        const float alpha = roughness * roughness;
        if (metallic >= 1.0f) return make_schlick(s4(1.0f), alpha).pdf(wo, wi);
        if (metallic <= 0.0f) return pdf_dielectric(alpha, wo, wi);
        return make_schlick(s4(1.0f), alpha).pdf(wo, wi) * metallic + pdf_dielectric(alpha, wo, wi) * (1.0f - metallic);
    }
    __device__ __noinline__ static float pdf_v(const PbrBase self, float3 wo, float3 wi) { return self.pdf_i(wo, wi); }   // `this` travels by value, not through the caller's stack frame
    __device__ __forceinline__ float pdf(float3 wo, float3 wi) const { return pdf_v(*this, wo, wi); }
};

// Beer-Lambert coat attenuation (simple_pbr_clearcoat_material.rs:88-107)
__device__ __noinline__ S4 coat_attenuation(const S4& tint, float thickness, float cos_theta) {
    S4 r;
    const float thickness_m = thickness * 0.001f;
    const float l = thickness_m / rmax(cos_theta, 1e-4f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float log_tint = logf(rmax(tint.v[i], 1e-10f));
        const float sigma = (-1.0f * log_tint) / 0.001f;
        r.v[i] = expf((-1.0f * sigma) * l);
    }
    return r;
}

// ---------------------------------------------------------------- materials (BsdfSurfaceMaterial::{sample,evaluate,pdf}, material/traits.rs:29-83)
struct MatParams;
struct MatCtx {
    const DScene* sc;
    uint32_t path_key, depth;  // aux RNG keying
    const MatParams* mp;        // this vertex's material parameters, looked up once (see load_mat_params)
};

__device__ __noinline__ float3 param_normal(const DScene& sc, const tcpt_flat_material& m, float2 uv) {  // normal_texture.rs:40-66
    if (m.normal_texture < 0) return normalize(f3(0, 0, 1));
    const float3 rgb = tex_rgb(sc.textures[m.normal_texture], uv);
    float x = rgb.x * 2.0f - 1.0f, y = rgb.y * 2.0f - 1.0f, z = rgb.z * 2.0f - 1.0f;
    if (m.normal_flip_y) y = -y;
    const float len = sqrtf(x * x + y * y + z * z);
    if (len > 0.0f) return normalize(f3(x / len, y / len, z / len));
    return normalize(f3(0, 0, 1));
}
// Transform::from_normal_map (math/src/transform.rs:216-244)
struct NmFrame { M3 to_nm, from_nm; bool identity; };
__device__ __noinline__ void normal_map_frame(float3 nm, M3& to_nm, M3& from_nm) {
    const float3 z = normalize(nm);
    const float3 cand = fabsf(dot(z, f3(1, 0, 0))) < 0.9f ? f3(1, 0, 0) : f3(0, 1, 0);
    const float3 x = normalize(cand - dot(z, cand) * z);
    const float3 y = normalize(cross(z, x));
    M3 m; m.c0 = x; m.c1 = y; m.c2 = z;
    to_nm = m3_inverse(m);
    from_nm = m3_inverse(to_nm);
}
// The normal-map frame of a vertex: the reference rebuilds it inside sample(), evaluate() and pdf() from the same inputs
// (e.g. lambert_material.rs:53-59, 107-113, 146-152); it is built once per vertex here.  Without a normal map the normal is
// (0,0,1) and both matrices are exactly the identity (every cofactor is 0 or 1), so the transforms are skipped.
__device__ __forceinline__ void material_frame(const DScene& sc, const tcpt_flat_material& m, float2 sp_uv, NmFrame& f) {
    f.identity = m.normal_texture < 0;
    if (!f.identity) normal_map_frame(param_normal(sc, m, sp_uv), f.to_nm, f.from_nm);
}
__device__ __forceinline__ float3 to_nm(const NmFrame& f, float3 v) { return f.identity ? v : m3_vector(f.to_nm, v); }

__device__ __forceinline__ PbrBase load_pbr(const DScene& sc, const tcpt_flat_material& m, float2 uv, const DWavelengths& wl) {
    PbrBase b;
    b.base_color = spectrum_sample(sc, param_spectrum(sc, m.color, uv), wl);
    b.metallic = param_float(sc, m.metallic, uv);
    b.roughness = param_float(sc, m.roughness, uv);
    b.ior = param_float(sc, m.ior, uv);
    return b;
}
struct Coat { float ior, roughness, thickness; S4 tint; };
__device__ __forceinline__ Coat load_coat(const DScene& sc, const tcpt_flat_material& m, float2 uv, const DWavelengths& wl) {
    Coat c;
    c.ior = param_float(sc, m.coat_ior, uv);
    c.roughness = param_float(sc, m.coat_roughness, uv);
    c.tint = spectrum_sample(sc, param_spectrum(sc, m.coat_tint, uv), wl);
    c.thickness = param_float(sc, m.coat_thickness, uv);
    return c;
}
__device__ __forceinline__ Schlick coat_bsdf(const Coat& c) { return make_schlick(s4(r0_of(c.ior)), c.roughness * c.roughness); }

// The parameters a material reads at a vertex are functions of (material, surface uv, wavelengths) only; sample(), evaluate() and
// pdf() of the reference each look them up again (texture taps, rgb -> spectrum, sigmoids).  Looked up once per vertex here and
// shared: same values, same bits.  Not for the dispersive dielectrics, whose sample() may collapse the wavelengths in between.
struct MatParams { PbrBase pbr; Coat coat; S4 k; };
template <int MT>
__device__ __forceinline__ void load_mat_params(const DScene& sc, const tcpt_flat_material& m, float2 sp_uv, const DWavelengths& wl, MatParams& p) {
    if (MT == TCPT_MAT_LAMBERT) p.pbr.base_color = spectrum_sample(sc, param_spectrum(sc, m.color, sp_uv), wl);
    if (MT == TCPT_MAT_SIMPLE_PBR || MT == TCPT_MAT_CLEARCOAT_PBR) p.pbr = load_pbr(sc, m, sp_uv, wl);
    if (MT == TCPT_MAT_CLEARCOAT_PBR) p.coat = load_coat(sc, m, sp_uv, wl);
    if (MT == TCPT_MAT_METAL) {
        p.pbr.base_color = spectrum_sample(sc, spectrum_from_flat(m.color), wl); p.k = spectrum_sample(sc, spectrum_from_flat(m.coat_tint), wl);
        p.pbr.roughness = param_float(sc, m.roughness, sp_uv);
    }
}

// `ng_t` = geometric normal in the tangent frame, `sp_uv` = surface uv
// MT = the material type as a compile-time constant: k_shade is instantiated once per shading bucket, so each instantiation
// carries only its own material's code (I-cache footprint and register pressure of the fused kernel were the first bottleneck)
template <int MT>
__device__ __forceinline__ MatSample material_sample(const MatCtx& c, const tcpt_flat_material& m, const NmFrame& fr, float uc, float2 uv, DWavelengths& wl, float3 wo, float3 ng_t, float2 sp_uv) {
    const DScene& sc = *c.sc;
    M3 from_nm;
    if (fr.identity) { from_nm.c0 = f3(1, 0, 0); from_nm.c1 = f3(0, 1, 0); from_nm.c2 = f3(0, 0, 1); } else from_nm = fr.from_nm;
    const float3 wo_nm = to_nm(fr, wo);
    switch (MT) {
        case TCPT_MAT_LAMBERT: {  // lambert_material.rs:42-97
            const S4 albedo = c.mp->pbr.base_color;
            BsdfSample s;
            if (!lambert_sample(albedo, wo_nm, uv, &s)) return mat_fail();
            const float3 wi_sh = m3_vector(from_nm, s.wi);
            if (signum(dot(ng_t, wi_sh)) != signum(dot(ng_t, wo))) return mat_fail();
            return mat_ok(s.f, wi_sh, s.pdf, s.type);
        }
        case TCPT_MAT_PLASTIC: {  // plastic_material.rs:122-187 (roughness passed unsquared as alpha)
            const float rough = param_float(sc, m.roughness, sp_uv);
            const Dielectric d = make_dielectric(m.eta, dot(ng_t, wo) > 0.0f, m.thin_surface != 0, rough);
            BsdfSample s;
            if (!d.sample(wo_nm, uv, uc, wl, &s)) return mat_fail();
            if (dot(s.wi, wo_nm) < 0.0f) s.f = s.f * spectrum_sample(sc, param_spectrum(sc, m.color, uv), wl);  // quirk: filter looked up at the RANDOM uv (:167)
            return mat_ok(s.f, m3_vector(from_nm, s.wi), s.pdf, s.type);
        }
        case TCPT_MAT_SIMPLE_PBR: return c.mp->pbr.sample(wo_nm, uc, uv, from_nm);
        case TCPT_MAT_METAL: {  // metal_material.rs:122-173: eta / k presets in `color` / `coat_tint`, alpha = roughness^2
            Conductor cb;
            cb.eta = c.mp->pbr.base_color; cb.k = c.mp->k;
            const float rough = c.mp->pbr.roughness;
            cb.g.ax = cb.g.ay = rough * rough;
            BsdfSample s;
            if (!cb.sample(wo_nm, uv, &s)) return mat_fail();
            const float3 wi_sh = m3_vector(from_nm, s.wi);
            if (signum(dot(ng_t, wi_sh)) != signum(dot(ng_t, wo))) return mat_fail();
            return mat_ok(s.f, wi_sh, s.pdf, s.type);
        }
        case TCPT_MAT_GLASS: {  // glass_material.rs:97-147: eta(lambda) preset in `color`, roughness passed unsquared as alpha
            const S4 eta = spectrum_sample(sc, spectrum_from_flat(m.color), wl);
            const float rough = param_float(sc, m.roughness, sp_uv);
            const Dielectric d = make_dielectric(eta, dot(ng_t, wo) > 0.0f, m.thin_surface != 0, rough);
            BsdfSample s;
            if (!d.sample(wo_nm, uv, uc, wl, &s)) return mat_fail();
            return mat_ok(s.f, m3_vector(from_nm, s.wi), s.pdf, s.type);
        }
        case TCPT_MAT_CLEARCOAT_PBR: {  // simple_pbr_clearcoat_material.rs:121-250
            const PbrBase& b = c.mp->pbr;
            const Coat& cp = c.mp->coat;
            if (cp.thickness <= 0.0f) return b.sample(wo_nm, uc, uv, from_nm);
            const Schlick coat = coat_bsdf(cp);
            const float fc = s4_avg(coat.directional_albedo(wo_nm, aux_rng(c.path_key, c.depth, 0)));
            if (uc < fc) {
                BsdfSample s;
                if (!coat.sample(wo_nm, uv, &s)) return mat_fail();
                return mat_ok(s.f, m3_vector(from_nm, s.wi), s.pdf * fc, s.type);
            }
            const float uc2 = (uc - fc) / (1.0f - fc);
            MatSample sub = b.sample(wo_nm, uc2, uv, from_nm);
            if (!sub.sampled) return sub;
            const S4 att = coat_attenuation(cp.tint, cp.thickness, wo_nm.z) * coat_attenuation(cp.tint, cp.thickness, sub.wi.z);
            return mat_ok(sub.f * att, sub.wi, sub.pdf * (1.0f - fc), sub.type);
        }
        default: return mat_fail();
    }
}

// evaluate() and pdf() of the same (wo, wi) pair, as the NEE helpers call them back to back (common.rs:142-158)
template <int MT>
__device__ __forceinline__ void material_eval_pdf(const MatCtx& c, const tcpt_flat_material& m, const NmFrame& fr, const DWavelengths& wl, float3 wo, float3 wi, float3 ng_t, float2 sp_uv,
                                         bool want_pdf, S4* f_out, float* pdf_out) {
    const DScene& sc = *c.sc;
    const float3 wo_nm = to_nm(fr, wo), wi_nm = to_nm(fr, wi);
    *pdf_out = 0.0f;
    switch (MT) {
        case TCPT_MAT_LAMBERT: {  // lambert_material.rs:99-170
            if (signum(dot(ng_t, wi)) != signum(dot(ng_t, wo))) { *f_out = s4(0.0f); return; }
            const S4 albedo = c.mp->pbr.base_color;
            *f_out = lambert_eval(albedo, wo_nm, wi_nm);
            if (want_pdf) *pdf_out = lambert_pdf(wo_nm, wi_nm);
            return;
        }
        case TCPT_MAT_PLASTIC: {  // plastic_material.rs:189-264
            const float rough = param_float(sc, m.roughness, sp_uv);
            const Dielectric d = make_dielectric(m.eta, dot(ng_t, wo) > 0.0f, m.thin_surface != 0, rough);
            S4 f = d.evaluate(wo_nm, wi_nm);
            if (dot(wi_nm, wo_nm) < 0.0f) f = f * spectrum_sample(sc, param_spectrum(sc, m.color, sp_uv), wl);
            *f_out = f;
            if (want_pdf) *pdf_out = d.pdf(wo_nm, wi_nm);
            return;
        }
        case TCPT_MAT_SIMPLE_PBR: {
            const PbrBase& b = c.mp->pbr;
            *f_out = b.evaluate(wo_nm, wi_nm);
            if (want_pdf) *pdf_out = b.pdf(wo_nm, wi_nm);
            return;
        }
        case TCPT_MAT_METAL: {  // metal_material.rs:175-252
            if (signum(dot(ng_t, wi)) != signum(dot(ng_t, wo))) { *f_out = s4(0.0f); return; }
            Conductor cb;
            cb.eta = c.mp->pbr.base_color; cb.k = c.mp->k;
            const float rough = c.mp->pbr.roughness;
            cb.g.ax = cb.g.ay = rough * rough;
            *f_out = cb.evaluate(wo_nm, wi_nm);
            if (want_pdf) *pdf_out = cb.pdf(wo_nm, wi_nm);
            return;
        }
        case TCPT_MAT_GLASS: {  // glass_material.rs:149-221
            const float rough = param_float(sc, m.roughness, sp_uv);
            const Dielectric d = make_dielectric(spectrum_sample(sc, spectrum_from_flat(m.color), wl), dot(ng_t, wo) > 0.0f, m.thin_surface != 0, rough);
            *f_out = d.evaluate(wo_nm, wi_nm);
            if (want_pdf) *pdf_out = d.pdf(wo_nm, wi_nm);
            return;
        }
        case TCPT_MAT_CLEARCOAT_PBR: {  // simple_pbr_clearcoat_material.rs:252-433: evaluate and pdf each draw their OWN albedo estimate
            const PbrBase& b = c.mp->pbr;
            const Coat& cp = c.mp->coat;
            if (cp.thickness <= 0.0f) { *f_out = b.evaluate(wo_nm, wi_nm); if (want_pdf) *pdf_out = b.pdf(wo_nm, wi_nm); return; }
            const Schlick coat = coat_bsdf(cp);
            const float fc = s4_avg(coat.directional_albedo(wo_nm, aux_rng(c.path_key, c.depth, 1)));
            const S4 att = coat_attenuation(cp.tint, cp.thickness, wo_nm.z) * coat_attenuation(cp.tint, cp.thickness, wi_nm.z);
            *f_out = coat.evaluate(wo_nm, wi_nm) * fc + b.evaluate(wo_nm, wi_nm) * att * (1.0f - fc);
            if (want_pdf) {
                const float fc2 = s4_avg(coat.directional_albedo(wo_nm, aux_rng(c.path_key, c.depth, 2)));
                *pdf_out = coat.pdf(wo_nm, wi_nm) * fc2 + b.pdf(wo_nm, wi_nm) * (1.0f - fc2);
            }
            return;
        }
        default: *f_out = s4(0.0f); return;
    }
}

// ---------------------------------------------------------------- lights
// emissive_material.rs:48-80 (UniformEdf: two-sided, direction independent)
__device__ __forceinline__ S4 emissive_radiance(const DScene& sc, const tcpt_flat_material& m, float2 uv, const DWavelengths& wl) {
    const S4 r = spectrum_sample(sc, param_spectrum(sc, m.color, uv), wl);
    return r * param_float(sc, m.intensity, uv);
}

struct LightTable { float w[TCPT_MAX_LIGHTS]; float sum; };
// LightSamplerFactory::create (light_sampler.rs:190-220): phi(lambda).average() per light
__device__ __noinline__ void light_table(const DScene& sc, const DWavelengths& wl, LightTable& lt) {
    lt.sum = 0.0f;
    for (uint32_t i = 0; i < sc.n_lights; ++i) {
        const tcpt_flat_primitive& P = sc.primitives[sc.light_list[i]];
        S4 phi;
        if (P.kind == 1) phi = emissive_radiance(sc, sc.materials[P.material], make_float2(0.5f, 0.5f), wl) * P.area_sum;  // emissive_triangle_mesh.rs:166-173
        else if (P.kind == 3) phi = spectrum_sample(sc, spectrum_from_flat(P.light_spectrum), wl) * (4.0f * TCPT_PI * P.light_intensity);  // point_light.rs:70-72
        else if (P.kind == 4) {  // spot_light.rs:84-95
            const float ci = cosf(P.angle_inner), co = cosf(P.angle_outer);
            phi = spectrum_sample(sc, spectrum_from_flat(P.light_spectrum), wl) * P.light_intensity * 2.0f * TCPT_PI * ((1.0f - ci) + (ci - co) / 2.0f);
        } else if (P.kind == 5) phi = spectrum_sample(sc, spectrum_from_flat(P.light_spectrum), wl) * (P.light_intensity * P.dir_area);  // directional_light.rs:80-83
        else { const DEnv& e = sc.envs[P.env]; phi = e.intensity * spectrum_sample(sc, spectrum_from_flat(e.integrated), wl); }  // environment_light.rs:299-301
        lt.w[i] = s4_avg(phi);
        lt.sum += lt.w[i];
    }
}
// LightSampler::sample_light (light_sampler.rs:26-44): linear scan of the normalised running sum
__device__ inline int sample_light(const DScene& sc, const LightTable& lt, float u, float* prob) {
    if (sc.n_lights == 0 || lt.sum == 0.0f) return -1;
    float cum = 0.0f;
    for (uint32_t i = 0; i < sc.n_lights; ++i) {
        cum += lt.w[i];
        if (u < cum / lt.sum) { *prob = lt.w[i] / lt.sum; return (int)i; }
    }
    *prob = lt.w[sc.n_lights - 1] / lt.sum;
    return (int)sc.n_lights - 1;
}

// environment_light.rs:102-116
__device__ __forceinline__ void direction_to_spherical(float3 d, float* theta, float* phi) {
    *theta = clampf(acosf(d.y), 0.0f, TCPT_PI);
    float p = atan2f(d.z, d.x);
    if (p < 0.0f) p += 2.0f * TCPT_PI;
    *phi = p;
}
__device__ inline float3 env_texel_bilinear(const DEnv& e, float u, float v) {  // environment_light.rs:124-160
    u = clampf(u, 0.0f, 1.0f); v = clampf(v, 0.0f, 1.0f);
    const float x = u * (float)(e.w - 1), y = v * (float)(e.h - 1);
    const uint32_t x0 = f2u_sat(floorf(x)), y0 = f2u_sat(floorf(y));
    const uint32_t x1 = min(x0 + 1u, e.w - 1u), y1 = min(y0 + 1u, e.h - 1u);
    const float fx = x - (float)x0, fy = y - (float)y0;
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float p00 = __ldg(e.data + ((size_t)y0 * e.w + x0) * 3 + c), p10 = __ldg(e.data + ((size_t)y0 * e.w + x1) * 3 + c);
        const float p01 = __ldg(e.data + ((size_t)y1 * e.w + x0) * 3 + c), p11 = __ldg(e.data + ((size_t)y1 * e.w + x1) * 3 + c);
        const float a = p00 * (1.0f - fx) + p10 * fx, b = p01 * (1.0f - fx) + p11 * fx;
        o[c] = a * (1.0f - fy) + b * fy;
    }
    return f3(o[0], o[1], o[2]);
}
__device__ __noinline__ float env_pdf_spherical(const DEnv& e, float theta, float phi) {  // environment_light.rs:234-259
    if (e.total_weight <= 0.0f) return 0.0f;
    const float u = phi / (2.0f * TCPT_PI), v = theta / TCPT_PI;
    const uint32_t x = min(f2u_sat(floorf(u * (float)e.w)), e.w - 1u), y = min(f2u_sat(floorf(v * (float)e.h)), e.h - 1u);
    const float* px = e.data + ((size_t)y * e.w + x) * 3;
    const float lum = 0.299f * __ldg(px) + 0.587f * __ldg(px + 1) + 0.114f * __ldg(px + 2);
    const float sin_theta = rmax(sinf(theta), 1e-8f);
    const float pdf_texture = lum * sin_theta / e.total_weight;
    const float jac = (float)e.w * (float)e.h / (2.0f * TCPT_PI * TCPT_PI * sin_theta);
    return pdf_texture * jac;
}
__device__ __noinline__ S4 env_radiance_spherical(const DScene& sc, const DEnv& e, float theta, float phi, const DWavelengths& wl) {  // environment_light.rs:304-316
    const float3 rgb = env_texel_bilinear(e, phi / (2.0f * TCPT_PI), theta / TCPT_PI);
    return spectrum_sample(sc, illuminant_from_rgb(sc, rgb), wl) * e.intensity;
}
// EnvironmentLight::sample_infinite_light (environment_light.rs:326-350) from the drawn texel (xx, yy) on: direction through the texel
// centre, the pdf of that direction (:234-259) and the illuminant spectrum of the bilinear lookup there (:304-316).  A pure function of
// the texel, so k_env_nee_table evaluates it once per texel with this very code and the shading kernels read the result (bit-identical:
// same instructions, same inputs); env_nee_texel_call is the per-sample fallback when the table is switched off.
struct EnvNee { float3 wi_r; float pdf_dir; DSpectrum spec; };
__device__ __forceinline__ EnvNee env_nee_texel(const DScene& sc, const tcpt_flat_primitive& LP, const DEnv& e, uint32_t xx, uint32_t yy) {
    EnvNee r;
    const float eu = ((float)xx + 0.5f) / (float)e.w, ev = ((float)yy + 0.5f) / (float)e.h;
    const float theta = ev * TCPT_PI, phi = eu * 2.0f * TCPT_PI;
    const float3 wl_local = f3(sinf(theta) * cosf(phi), cosf(theta), sinf(theta) * sinf(phi));
    r.wi_r = xf_vector(LP.l2r, wl_local);
    float th_l, ph_l;
    direction_to_spherical(xf_vector(LP.r2l, r.wi_r), &th_l, &ph_l);
    r.pdf_dir = env_pdf_spherical(e, th_l, ph_l);
    r.spec = illuminant_from_rgb(sc, env_texel_bilinear(e, ph_l / (2.0f * TCPT_PI), th_l / TCPT_PI));
    return r;
}
__device__ __noinline__ EnvNee env_nee_texel_call(const DScene& sc, const tcpt_flat_primitive& LP, const DEnv& e, uint32_t xx, uint32_t yy) { return env_nee_texel(sc, LP, e, xx, yy); }
__device__ __forceinline__ EnvNee env_nee_lookup(const DScene& sc, const tcpt_flat_primitive& LP, const DEnv& e, uint32_t xx, uint32_t yy) {
    if (e.nee_table == nullptr) return env_nee_texel_call(sc, LP, e, xx, yy);
    const float4* row = e.nee_table + 2 * ((size_t)yy * e.w + xx);
    const float4 a = __ldg(row), b = __ldg(row + 1);
    EnvNee r;
    r.wi_r = f3(a.x, a.y, a.z); r.pdf_dir = a.w;
    r.spec.kind = 2; r.spec.table = 0; r.spec.c[0] = b.x; r.spec.c[1] = b.y; r.spec.c[2] = b.z; r.spec.scale = b.w;
    return r;
}
// direction -> (theta, phi) is evaluated once per direction and shared by the pdf and radiance lookups (the reference
// recomputes it inside each from the same direction: identical values)
__device__ __forceinline__ float env_direction_pdf(const DScene& sc, const tcpt_flat_primitive& P, float3 dir) {
    float theta, phi;
    direction_to_spherical(xf_vector(P.r2l, dir), &theta, &phi);
    return env_pdf_spherical(sc.envs[P.env], theta, phi);
}
__device__ __forceinline__ S4 env_direction_radiance(const DScene& sc, const tcpt_flat_primitive& P, float3 dir, const DWavelengths& wl) {
    float theta, phi;
    direction_to_spherical(xf_vector(P.r2l, dir), &theta, &phi);
    return env_radiance_spherical(sc, sc.envs[P.env], theta, phi, wl);
}
// Scene::evaluate_infinite_light_radiance (scene.rs:213-231)
__device__ inline S4 scene_env_radiance(const DScene& sc, float3 dir, const DWavelengths& wl) {
    S4 tot = s4(0.0f);
    for (uint32_t i = 0; i < sc.n_envs; ++i) tot = tot + env_direction_radiance(sc, sc.primitives[sc.envs[i].primitive], dir, wl);
    return tot;
}
// Scene::pdf_infinite_light_sample (scene.rs:184-210) with LightSampler::probability_infinite_light (light_sampler.rs:121-158)
__device__ inline float scene_env_pdf(const DScene& sc, const LightTable& lt, float3 dir) {
    if (sc.n_lights == 0 || lt.sum == 0.0f) return 0.0f;
    float inf_sum = 0.0f;
    for (uint32_t i = 0; i < sc.n_lights; ++i) if (sc.primitives[sc.light_list[i]].kind == 2) inf_sum += lt.w[i];
    if (inf_sum == 0.0f) return 0.0f;
    float tot = 0.0f;
    for (uint32_t i = 0; i < sc.n_lights; ++i) {
        const tcpt_flat_primitive& P = sc.primitives[sc.light_list[i]];
        if (P.kind == 2) tot += (lt.w[i] / inf_sum) * env_direction_pdf(sc, P, dir);
    }
    return tot;
}
// radiance (scene.rs:213-231) and, for MIS, the light-sampling pdf (scene.rs:184-210) of an escaped direction in one pass over the
// environment lights, sharing the spherical coordinates; both sums run in primitive order like the reference's loops
__device__ inline void scene_env_radiance_pdf(const DScene& sc, const LightTable* lt, float3 dir, const DWavelengths& wl, S4* radiance, float* pdf) {
    S4 tot = s4(0.0f);
    float inf_sum = 0.0f, ptot = 0.0f;
    const bool want_pdf = lt != nullptr && sc.n_lights != 0 && lt->sum != 0.0f;
    if (want_pdf) for (uint32_t i = 0; i < sc.n_lights; ++i) if (sc.primitives[sc.light_list[i]].kind == 2) inf_sum += lt->w[i];
    for (uint32_t i = 0; i < sc.n_envs; ++i) {
        const DEnv& e = sc.envs[i];
        const tcpt_flat_primitive& P = sc.primitives[e.primitive];
        float theta, phi;
        direction_to_spherical(xf_vector(P.r2l, dir), &theta, &phi);
        tot = tot + env_radiance_spherical(sc, e, theta, phi, wl);
        if (want_pdf && inf_sum != 0.0f && P.light_index >= 0) ptot += (lt->w[P.light_index] / inf_sum) * env_pdf_spherical(e, theta, phi);
    }
    *radiance = tot;
    if (pdf) *pdf = ptot;
}
// Rust slice::binary_search_by(partial_cmp) then clamp (environment_light.rs:218-223).  On a non-decreasing CDF the
// reference's probing sequence ends at base = (#entries <= u) - 1 (or 0), so with k = #{i : cdf[i] <= u} the result is
//   k == 0 -> 0 ;  cdf[k-1] == u -> k-1 ;  else min(k, n-1)
// and k is found with the guide table (tcpt_flat_env) plus a bisection / scan of the few entries it leaves, instead of
// log2(n) dependent loads (profile: the 19 serial L2 round trips of the two searches were the top stall of bounce-0 shading).
__device__ __noinline__ uint32_t sample_from_cdf(const float* cdf, uint32_t n, const uint32_t* guide, uint32_t G, float u) {
    const uint32_t j = min(f2u_sat(u * (float)G), G - 1u);   // u * 2^m is exact
    uint32_t k = __ldg(guide + j);
    const uint32_t hi = __ldg(guide + j + 1);
    uint32_t top = hi;
    while (top - k > 2u) { const uint32_t mid = (k + top) >> 1; if (__ldg(cdf + mid) <= u) k = mid + 1u; else top = mid; }   // entries <= u are a prefix of [k, hi)
    while (k < top && __ldg(cdf + k) <= u) ++k;
    if (k == 0u) return 0u;
    const float c = __ldg(cdf + k - 1u);
    if (c == u) return k - 1u;
    return min(k, n - 1u);
}

// Scene::pdf_light_sample (scene.rs:156-181) for a BSDF-sampled hit on an emissive mesh
// (takes the four things it reads of the hit by value: a `const DSurface&` kept the whole surface record in the caller's stack frame)
__device__ __noinline__ float scene_pdf_light_sample(const DScene& sc, const LightTable& lt, float3 shading_pos, int hit_prim, uint32_t hit_tri, float3 hit_position, float3 hit_normal) {
    const tcpt_flat_primitive& P = sc.primitives[hit_prim];
    if (P.kind != 1) return 0.0f;
    float probability = 0.0f;
    if (sc.n_lights != 0 && lt.sum != 0.0f && P.light_index >= 0) probability = lt.w[P.light_index] / lt.sum;
    const float* table = sc.area_table + P.area_base;
    const float tri_prob = hit_tri == 0 ? __ldg(table) : __ldg(table + hit_tri) - __ldg(table + hit_tri - 1);
    const float pdf_area = 1.0f / __ldg(sc.area_list + P.area_base + hit_tri) * tri_prob;  // emissive_triangle_mesh.rs:331-353
    const float3 dv = shading_pos - hit_position;
    const float distance = length(dv);
    const float3 wo = -normalize(dv);
    const float pdf_dir = pdf_area * (distance * distance) / fabsf(dot(hit_normal, wo));
    return probability * pdf_dir;
}

// ---------------------------------------------------------------- sensor (renderer/src/sensor.rs:41-78)
// (by value: the wavelength record is a pure function of (lambda0, terminated) -- terminate_secondary, sampled_spectrum.rs:351-360, leaves
// exactly what wavelengths_uniform builds for a terminated path -- and is rebuilt here instead of being read from the caller's stack)
__device__ __noinline__ float3 sensor_rgb(const DScene& sc, float lambda0, bool terminated, const S4 s, float exposure) {
    const DWavelengths wl = wavelengths_uniform(lambda0, terminated);
    const int count = wl.terminated ? 1 : 4;
    float x = 0.0f, y = 0.0f, z = 0.0f;
    for (int k = 0; k < count; ++k) {
        uint32_t i = f2u_sat(floorf(wl.lambda[k] - 360.0f));
        if (i == 470u) i = 0u;
        const float a = s.v[k] / wl.pdf[k];
        const float nc = a / 4.0f;
        const float4 cm = cmf_at(sc, 360.0f + (float)i);
        x += nc * cm.x; y += nc * cm.y; z += nc * cm.z;
    }
    const float* m = sc.xyz_to_rgb;
    const float3 rgb = f3((m[0] * x + m[3] * y) + m[6] * z, (m[1] * x + m[4] * y) + m[7] * z, (m[2] * x + m[5] * y) + m[8] * z);
    return rgb * exposure;
}

}  // namespace tcpt
