// ORACLE (test infrastructure).  BSDFs and materials, restated from
//   /root/reference/scene/src/material/common.rs, bsdf/{lambert,dielectric,generalized_schlick}.rs,
//   impls/{lambert,plastic,simple_pbr,simple_pbr_clearcoat}_material.rs, texture/{sampler,normal_texture,float_texture,rgb_texture}.rs,
//   math/src/transform.rs:216-244 (normal-map frame).
// All `f` values include the cosine |wi.z| by this codebase's convention.
#pragma once
#include <cmath>
#include <vector>

#include "omath.h"
#include "osampler.h"
#include "ospectrum.h"

namespace orc {

constexpr float PI_F = 3.14159265358979323846f;

// ---------------------------------------------------------------- textures (texture/sampler.rs)
struct Texture {
    std::vector<uint8_t> data;
    uint32_t w = 0, h = 0, channels = 3;
};
inline float rust_fract(float x) { return x - std::trunc(x); }
inline float lerp2d(float p00, float p10, float p01, float p11, float fx, float fy) {
    float top = p00 * (1.0f - fx) + p10 * fx;
    float bottom = p01 * (1.0f - fx) + p11 * fx;
    return top * (1.0f - fy) + bottom * fy;
}
struct BilinearTaps { uint32_t x0, y0, x1, y1; float fx, fy; };
inline BilinearTaps bilinear_taps(uint32_t width, uint32_t height, Vec2 uv) {
    float u = std::fabs(rust_fract(uv.x));
    float v = 1.0f - std::fabs(rust_fract(uv.y));
    float x = u * ((float)width - 1.0f);
    float y = v * ((float)height - 1.0f);
    BilinearTaps t;
    t.x0 = f2u_sat(std::floor(x));
    t.y0 = f2u_sat(std::floor(y));
    t.x1 = std::min(t.x0 + 1, width - 1);
    t.y1 = std::min(t.y0 + 1, height - 1);
    t.fx = x - (float)t.x0;
    t.fy = y - (float)t.y0;
    return t;
}
inline void bilinear_sample_rgb(const Texture& tex, Vec2 uv, float out[3]) {
    BilinearTaps t = bilinear_taps(tex.w, tex.h, uv);
    auto px = [&](uint32_t x, uint32_t y, int c) { return (float)tex.data[((size_t)y * tex.w + x) * 3 + c] / 255.0f; };
    for (int c = 0; c < 3; ++c) out[c] = lerp2d(px(t.x0, t.y0, c), px(t.x1, t.y0, c), px(t.x0, t.y1, c), px(t.x1, t.y1, c), t.fx, t.fy);
}
inline float bilinear_sample_gray(const Texture& tex, Vec2 uv) {
    BilinearTaps t = bilinear_taps(tex.w, tex.h, uv);
    auto px = [&](uint32_t x, uint32_t y) { return (float)tex.data[(size_t)y * tex.w + x] / 255.0f; };
    return lerp2d(px(t.x0, t.y0), px(t.x1, t.y0), px(t.x0, t.y1), px(t.x1, t.y1), t.fx, t.fy);
}

// ---------------------------------------------------------------- parameters (material/parameter.rs)
struct SpectrumParam {
    bool is_texture = false;
    Spectrum spectrum;  // constant
    int texture = -1;   // RgbTexture sRGB, SpectrumType::Albedo
};
struct FloatParam {
    bool is_texture = false;
    float value = 0.0f;
    int texture = -1;
    bool gamma_corrected = false;
};
struct NormalParam {
    int texture = -1;
    bool flip_y = false;
};

enum MaterialType : int { MAT_LAMBERT = 0, MAT_EMISSIVE = 1, MAT_PLASTIC = 2, MAT_SIMPLE_PBR = 3, MAT_CLEARCOAT_PBR = 4, MAT_METAL = 5, MAT_GLASS = 6 };
enum SampleType : int { ST_DIFFUSE = 0, ST_SPECULAR_REFLECTION = 1, ST_SPECULAR_TRANSMISSION = 2, ST_GLOSSY_REFLECTION = 3, ST_GLOSSY_TRANSMISSION = 4 };

struct BsdfSample {
    SampledSpectrum f;
    Vec3 wi;
    float pdf;
    int sample_type;
};
struct MaterialSample {
    SampledSpectrum f = SampledSpectrum::zero();
    Vec3 wi = Vec3(0, 0, 1);
    float pdf = 0.0f;
    int sample_type = ST_DIFFUSE;
    bool is_sampled = false;
    // samples.rs:69-95: a failed sample reports Diffuse and therefore counts as non-specular
    bool is_specular() const { return sample_type == ST_SPECULAR_REFLECTION || sample_type == ST_SPECULAR_TRANSMISSION; }
    bool is_non_specular() const { return sample_type == ST_DIFFUSE || sample_type == ST_GLOSSY_REFLECTION || sample_type == ST_GLOSSY_TRANSMISSION; }
};
inline MaterialSample make_sample(const SampledSpectrum& f, Vec3 wi, float pdf, int st) {
    MaterialSample m; m.f = f; m.wi = wi; m.pdf = pdf; m.sample_type = st; m.is_sampled = true; return m;
}

// shading point in the VertexNormalTangent frame (SurfaceInteraction<VertexNormalTangent>): only normal + uv are read by BSDFs
struct TangentShadingPoint {
    Vec3 normal;  // geometric normal in the tangent frame
    Vec2 uv;
};

// ---------------------------------------------------------------- common.rs
inline float cos_theta(Vec3 w) { return w.z; }
inline float cos2_theta(Vec3 w) { return w.z * w.z; }
inline float abs_cos_theta(Vec3 w) { return std::fabs(w.z); }
inline float tan2_theta(Vec3 w) { float c2 = cos2_theta(w); return c2 == 0.0f ? INFINITY : (1.0f - c2) / c2; }
inline float cos_phi(Vec3 w) { float st = std::sqrt(rmax(1.0f - cos2_theta(w), 0.0f)); return st == 0.0f ? 1.0f : clampf(w.x / st, -1.0f, 1.0f); }
inline float sin_phi(Vec3 w) { float st = std::sqrt(rmax(1.0f - cos2_theta(w), 0.0f)); return st == 0.0f ? 0.0f : clampf(w.y / st, -1.0f, 1.0f); }
inline bool half_vector(Vec3 wo, Vec3 wi, Vec3* wm) { Vec3 m = wo + wi; if (length_squared(m) == 0.0f) return false; *wm = normalize(m); return true; }
inline Vec3 reflect(Vec3 wo, Vec3 n) { return n * (2.0f * dot(wo, n)) - wo; }
inline bool same_hemisphere(Vec3 a, Vec3 b) { return a.z * b.z > 0.0f; }
inline Vec2 sample_uniform_disk_polar(Vec2 u) { float r = std::sqrt(u.x); float th = 2.0f * PI_F * u.y; Vec2 o; o.x = r * std::cos(th); o.y = r * std::sin(th); return o; }
inline SampledSpectrum fresnel_dielectric(float cos_theta_i, const SampledSpectrum& eta) {
    cos_theta_i = clampf(cos_theta_i, 0.0f, 1.0f);
    float sin2_theta_i = 1.0f - cos_theta_i * cos_theta_i;
    SampledSpectrum sin2_t = SampledSpectrum::constant(sin2_theta_i) / (eta * eta);
    SampledSpectrum cos_t = (SampledSpectrum::one() - sin2_t).clamp(0.0f, 1.0f).sqrt();
    SampledSpectrum ci = SampledSpectrum::constant(cos_theta_i);
    SampledSpectrum r_parl = (eta * ci - cos_t) / (eta * ci + cos_t);
    SampledSpectrum r_perp = (ci - eta * cos_t) / (ci + eta * cos_t);
    return (r_parl * r_parl + r_perp * r_perp) * 0.5f;
}
inline bool refract(Vec3 wi, Vec3 n, float eta, Vec3* wt) {
    float cos_theta_i = dot(n, wi);
    float sin2_theta_i = rmax(1.0f - cos_theta_i * cos_theta_i, 0.0f);
    float sin2_theta_t = sin2_theta_i / (eta * eta);
    if (sin2_theta_t >= 1.0f) return false;
    float cos_theta_t = std::sqrt(rmax(1.0f - sin2_theta_t, 0.0f));
    Vec3 t = (-wi) / eta + n * (cos_theta_i / eta - cos_theta_t);
    if (length_squared(t) < 1e-12f) return false;
    *wt = normalize(t);
    return true;
}
inline float powi2(float x) { return x * x; }
inline float powi6(float x) { float x2 = x * x; float x4 = x2 * x2; return x4 * x2; }  // Rust powi(6) via llvm.powi: x^2 -> x^4 -> * x^2

// ---------------------------------------------------------------- GGX helpers shared verbatim by dielectric.rs:22-120 and generalized_schlick.rs:96-190
struct Ggx {
    float ax, ay;
    bool effectively_smooth() const { return rmax(ax, ay) < 1e-3f; }
    float D(Vec3 wm) const {
        float t2 = tan2_theta(wm);
        if (!std::isfinite(t2)) return 0.0f;
        float cos4 = powi2(cos2_theta(wm));
        float e = t2 * (powi2(cos_phi(wm)) / powi2(ax) + powi2(sin_phi(wm)) / powi2(ay));
        return 1.0f / (PI_F * ax * ay * cos4 * powi2(1.0f + e));
    }
    float lambda(Vec3 w) const {
        float t2 = tan2_theta(w);
        if (std::isinf(t2)) return 0.0f;
        float a2 = powi2(cos_phi(w) * ax) + powi2(sin_phi(w) * ay);
        return (std::sqrt(1.0f + a2 * t2) - 1.0f) / 2.0f;
    }
    float G1(Vec3 w) const { return 1.0f / (1.0f + lambda(w)); }
    float G(Vec3 wo, Vec3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    float Dvis(Vec3 w, Vec3 wm) const {
        float c = std::fabs(w.z);
        if (c == 0.0f) return 0.0f;
        return G1(w) / c * D(wm) * std::fabs(dot(w, wm));
    }
    Vec3 sample_wm(Vec3 w, Vec2 u) const {
        Vec3 wh = normalize(Vec3(ax * w.x, ay * w.y, w.z));
        if (wh.z < 0.0f) wh = -wh;
        Vec3 t1 = wh.z < 0.99999f ? normalize(cross(Vec3(0, 0, 1), wh)) : Vec3(1, 0, 0);
        Vec3 t2 = cross(wh, t1);
        Vec2 p = sample_uniform_disk_polar(u);
        float h = std::sqrt(rmax(1.0f - p.x * p.x, 0.0f));
        float lf = (1.0f + wh.z) / 2.0f;
        float py = h * (1.0f - lf) + p.y * lf;
        float pz = std::sqrt(rmax(1.0f - p.x * p.x - py * py, 0.0f));
        Vec3 nh = t1 * p.x + t2 * py + wh * pz;
        return normalize(Vec3(ax * nh.x, ay * nh.y, rmax(1e-6f, nh.z)));
    }
};

inline bool generalized_half_vector(Vec3 wo, Vec3 wi, float eta, Vec3* out) {
    float co = cos_theta(wo), ci = cos_theta(wi);
    bool refl = ci * co > 0.0f;
    float etap = !refl ? (co > 0.0f ? eta : 1.0f / eta) : 1.0f;
    Vec3 wm = wi * etap + wo;
    if (ci == 0.0f || co == 0.0f || length_squared(wm) == 0.0f) return false;
    wm = normalize(wm);
    if (wm.z < 0.0f) wm = -wm;
    if (dot(wm, wi) * ci < 0.0f || dot(wm, wo) * co < 0.0f) return false;
    *out = wm;
    return true;
}

// ---------------------------------------------------------------- bsdf/lambert.rs
struct LambertBsdf {
    SampledSpectrum albedo;
    static Vec3 sample_cosine_hemisphere(Vec2 uv) {
        float r = std::sqrt(uv.x);
        float th = 2.0f * PI_F * uv.y;
        return Vec3(r * std::cos(th), r * std::sin(th), std::sqrt(1.0f - uv.x));
    }
    bool sample(Vec3 wo, Vec2 uv, BsdfSample* out) const {
        float wo_cos = wo.z;
        if (wo_cos == 0.0f) return false;
        Vec3 wi = sample_cosine_hemisphere(uv);
        if (wo_cos < 0.0f) wi = Vec3(wi.x, wi.y, -wi.z);
        float wi_cos = wi.z;
        if (wi_cos == 0.0f) return false;
        if (signum(wo_cos) != signum(wi_cos)) return false;
        out->f = albedo * std::fabs(wi_cos) / PI_F;
        out->pdf = std::fabs(wi_cos) / PI_F;
        out->wi = wi;
        out->sample_type = ST_DIFFUSE;
        return true;
    }
    SampledSpectrum evaluate(Vec3 wo, Vec3 wi) const {
        if (wo.z == 0.0f || wi.z == 0.0f) return SampledSpectrum::zero();
        if (signum(wo.z) != signum(wi.z)) return SampledSpectrum::zero();
        return albedo * std::fabs(wi.z) / PI_F;
    }
    float pdf(Vec3 wo, Vec3 wi) const {
        if (wo.z == 0.0f || wi.z == 0.0f) return 0.0f;
        if (signum(wo.z) != signum(wi.z)) return 0.0f;
        return std::fabs(wi.z) / PI_F;
    }
};

// ---------------------------------------------------------------- bsdf/dielectric.rs
struct DielectricBsdf {
    SampledSpectrum eta;
    bool entering, thin_surface;
    Ggx g;
    DielectricBsdf(SampledSpectrum eta_, bool entering_, bool thin_, float ax, float ay) : eta(eta_), entering(entering_), thin_surface(thin_), g{ax, ay} {
        if (eta.v[0] == 0.0f) eta = SampledSpectrum::constant(1.0f);
    }
    SampledSpectrum eta_spectrum() const { return (thin_surface || entering) ? eta : SampledSpectrum::one() / eta; }
    static void thin_coeffs(float fresnel, float* pr, float* pt) {
        float r = fresnel, t = 1.0f - r, r2 = r * r;
        r = r2 > 1.0f ? 1.0f : r + (t * t * r) / (1.0f - r2);
        *pr = r; *pt = t;
    }
    bool sample(Vec3 wo, Vec2 uv, float uc, SampledWavelengths& wl, BsdfSample* out) const {
        if (wo.z == 0.0f) return false;
        if (g.effectively_smooth()) { Vec2 u2; u2.x = uc; u2.y = uv.x; return sample_specular(wo, u2, wl, out); }
        return sample_microfacet(wo, uv, uc, wl, out);
    }
    bool sample_specular(Vec3 wo, Vec2 uv, SampledWavelengths& wl, BsdfSample* out) const {
        float wo_cos = wo.z;
        Vec3 n = entering ? Vec3(0, 0, 1) : Vec3(0, 0, -1);
        SampledSpectrum es = eta_spectrum();
        float etap = es.v[0];
        SampledSpectrum fresnel = fresnel_dielectric(std::fabs(wo_cos), es);
        float pr, pt;
        if (thin_surface) thin_coeffs(fresnel.average(), &pr, &pt);
        else { pr = fresnel.average(); pt = 1.0f - pr; }
        if (uv.x < pr / (pr + pt)) {
            if (std::fabs(wo_cos) < 1e-6f) return false;
            *out = BsdfSample{fresnel, Vec3(-wo.x, -wo.y, wo.z), pr / (pr + pt), ST_SPECULAR_REFLECTION};
            return true;
        }
        if (thin_surface) {
            Vec3 wi(-wo.x, -wo.y, -wo.z);
            if (wi.z == 0.0f) return false;
            *out = BsdfSample{SampledSpectrum::one() - fresnel, wi, pt / (pr + pt), ST_SPECULAR_TRANSMISSION};
            return true;
        }
        if (!eta.is_constant()) wl.terminate_secondary();
        Vec3 wt;
        if (!refract(wo, n, etap, &wt)) return false;
        if (wt.z == 0.0f) return false;
        SampledSpectrum tr = SampledSpectrum::one() - fresnel;
        *out = BsdfSample{tr / powi2(etap), wt, pt / (pr + pt), ST_SPECULAR_TRANSMISSION};
        return true;
    }
    bool sample_microfacet(Vec3 wo, Vec2 u, float uc, SampledWavelengths& wl, BsdfSample* out) const {
        Vec3 wm = g.sample_wm(wo, u);
        SampledSpectrum es = eta_spectrum();
        float eta_scalar = es.v[0];
        SampledSpectrum fresnel = fresnel_dielectric(std::fabs(dot(wo, wm)), es);
        float pr = fresnel.average(), pt = 1.0f - pr;
        if (thin_surface) {
            float tpr, tpt;
            thin_coeffs(fresnel.average(), &tpr, &tpt);
            if (uc < tpr / (tpr + tpt)) return sample_mf_reflection(wo, wm, fresnel, tpr / (tpr + tpt), out);
            Vec3 wi(-wo.x, -wo.y, -wo.z);
            *out = BsdfSample{SampledSpectrum::one() - fresnel, wi, tpt / (tpr + tpt), ST_GLOSSY_TRANSMISSION};
            return true;
        } else if (uc < pr / (pr + pt)) {
            return sample_mf_reflection(wo, wm, fresnel, pr / (pr + pt), out);
        }
        if (!eta.is_constant()) wl.terminate_secondary();
        return sample_mf_transmission(wo, wm, SampledSpectrum::one() - fresnel, pt / (pr + pt), eta_scalar, out);
    }
    bool sample_mf_reflection(Vec3 wo, Vec3 wm, const SampledSpectrum& fresnel, float prob, BsdfSample* out) const {
        Vec3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return false;
        float cd = std::fabs(dot(wo, wm));
        if (cd < 1e-6f) return false;
        float pdf = g.Dvis(wo, wm) / (4.0f * cd) * prob;
        float d = g.D(wm), gg = g.G(wo, wi);
        // quirk: extra |wi.z| here that evaluate_microfacet does not have (dielectric.rs:318 vs :589)
        SampledSpectrum f = fresnel * d * gg * abs_cos_theta(wi) / (4.0f * abs_cos_theta(wo));
        *out = BsdfSample{f, wi, pdf, ST_GLOSSY_REFLECTION};
        return true;
    }
    bool sample_mf_transmission(Vec3 wo, Vec3 wm, const SampledSpectrum& tr, float prob, float etap, BsdfSample* out) const {
        Vec3 wmr = entering ? wm : -wm;
        Vec3 wi;
        if (!refract(wo, wmr, etap, &wi)) return false;
        if (same_hemisphere(wo, wi) || std::fabs(wi.z) == 0.0f) return false;
        float denom = powi2(dot(wi, wm) + dot(wo, wm) / etap);
        float dwm_dwi = std::fabs(dot(wi, wm)) / denom;
        float pdf = g.Dvis(wo, wm) * dwm_dwi * prob;
        float d = g.D(wm), gg = g.G(wo, wi);
        SampledSpectrum ft = tr * d * gg * std::fabs(dot(wi, wm)) * std::fabs(dot(wo, wm)) / (denom * abs_cos_theta(wo) * etap * etap);
        *out = BsdfSample{ft, wi, pdf, ST_GLOSSY_TRANSMISSION};
        return true;
    }
    SampledSpectrum evaluate(Vec3 wo, Vec3 wi) const {
        if (g.effectively_smooth()) return SampledSpectrum::zero();
        SampledSpectrum es = eta_spectrum();
        float eta_scalar = es.v[0];
        Vec3 wm;
        if (!generalized_half_vector(wo, wi, eta_scalar, &wm)) return SampledSpectrum::zero();
        SampledSpectrum fresnel = fresnel_dielectric(std::fabs(dot(wo, wm)), es);
        bool refl = cos_theta(wi) * cos_theta(wo) > 0.0f;
        float d = g.D(wm), gg = g.G(wo, wi);
        if (refl) return fresnel * d * gg / (4.0f * abs_cos_theta(wo));
        float denom = powi2(dot(wi, wm) + dot(wo, wm) / eta_scalar);
        SampledSpectrum tr = SampledSpectrum::one() - fresnel;
        return tr * d * gg * std::fabs(dot(wi, wm)) * std::fabs(dot(wo, wm)) / (denom * abs_cos_theta(wo) * eta_scalar * eta_scalar);
    }
    float pdf(Vec3 wo, Vec3 wi) const {
        if (g.effectively_smooth()) return 0.0f;
        SampledSpectrum es = eta_spectrum();
        float eta_scalar = es.v[0];
        Vec3 wm;
        if (!generalized_half_vector(wo, wi, eta_scalar, &wm)) return 0.0f;
        SampledSpectrum fresnel = fresnel_dielectric(std::fabs(dot(wo, wm)), es);
        float pr = fresnel.average(), pt = 1.0f - pr;
        bool refl = cos_theta(wi) * cos_theta(wo) > 0.0f;
        if (refl) return g.Dvis(wo, wm) / (4.0f * std::fabs(dot(wo, wm))) * pr / (pr + pt);
        if (thin_surface) return pt / (pr + pt);
        float denom = powi2(dot(wi, wm) + dot(wo, wm) / eta_scalar);
        float dwm_dwi = std::fabs(dot(wi, wm)) / denom;
        return g.Dvis(wo, wm) * dwm_dwi * pt / (pr + pt);
    }
};

// ---------------------------------------------------------------- bsdf/generalized_schlick.rs, ScatterMode::R only
// (every call site in the materials passes ScatterMode::R: simple_pbr_material.rs:290-520, simple_pbr_clearcoat_material.rs:182-829)
struct GeneralizedSchlickBsdf {
    SampledSpectrum r0, r90;
    float exponent;
    SampledSpectrum tint;
    Ggx g;
    SampledSpectrum fresnel_at(float cos_theta) const {
        cos_theta = clampf(cos_theta, 0.0f, 1.0f);
        float omc = 1.0f - cos_theta;
        const float COS_MAX = 1.0f / 7.0f;
        const float OM_COS_MAX = 1.0f - COS_MAX;
        SampledSpectrum base = r0 + (r90 - r0) * std::pow(omc, exponent);
        SampledSpectrum at_max = r0 + (r90 - r0) * std::pow(OM_COS_MAX, exponent);
        SampledSpectrum a = at_max * (SampledSpectrum::one() - tint) / (COS_MAX * powi6(OM_COS_MAX));
        SampledSpectrum laz = a * cos_theta * powi6(omc);
        return base - laz;
    }
    SampledSpectrum fresnel(Vec3 wo) const { return fresnel_at(std::fabs(wo.z)); }
    bool sample(Vec3 wo, Vec2 uv, float /*uc*/, BsdfSample* out) const {
        if (wo.z == 0.0f) return false;
        if (g.effectively_smooth()) {
            SampledSpectrum fr = fresnel_at(std::fabs(wo.z));
            Vec3 wi(-wo.x, -wo.y, wo.z);
            if (wi.z == 0.0f) return false;
            *out = BsdfSample{fr, wi, 1.0f, ST_SPECULAR_REFLECTION};
            return true;
        }
        Vec3 wm = g.sample_wm(wo, uv);
        SampledSpectrum fr = fresnel_at(std::fabs(dot(wo, wm)));
        // sample_microfacet_reflection(prob = 1.0) (generalized_schlick.rs:419-458)
        Vec3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return false;
        float cd = std::fabs(dot(wo, wm));
        if (cd < 1e-6f) return false;
        float pdf = g.Dvis(wo, wm) / (4.0f * cd) * 1.0f;
        float d = g.D(wm), gg = g.G(wo, wi);
        float ci = std::fabs(wi.z), co = std::fabs(wo.z);
        if (ci == 0.0f || co == 0.0f) return false;
        *out = BsdfSample{fr * d * gg / (4.0f * co), wi, pdf, ST_GLOSSY_REFLECTION};
        return true;
    }
    SampledSpectrum evaluate(Vec3 wo, Vec3 wi) const {
        if (g.effectively_smooth()) return SampledSpectrum::zero();
        float co = std::fabs(wo.z), ci = std::fabs(wi.z);
        if (co == 0.0f || ci == 0.0f) return SampledSpectrum::zero();
        if (!same_hemisphere(wo, wi)) return SampledSpectrum::zero();
        Vec3 wm;
        if (!half_vector(wo, wi, &wm)) return SampledSpectrum::zero();
        SampledSpectrum fr = fresnel_at(std::fabs(dot(wo, wm)));
        float d = g.D(wm), gg = g.G(wo, wi);
        return fr * d * gg / (4.0f * co);
    }
    float pdf(Vec3 wo, Vec3 wi) const {
        if (g.effectively_smooth()) return 0.0f;
        if (!same_hemisphere(wo, wi)) return 0.0f;
        Vec3 wm;
        if (!half_vector(wo, wi, &wm)) return 0.0f;
        float vis = g.Dvis(wo, wm);
        float jac = 4.0f * std::fabs(dot(wo, wm));
        if (jac == 0.0f) return 0.0f;
        return vis / jac;
    }
    // 64-sample stochastic estimate driven by an independent RNG stream (generalized_schlick.rs:893-918)
    SampledSpectrum directional_albedo(Vec3 wo, AuxRng rng) const {
        SampledSpectrum sum = SampledSpectrum::zero();
        for (int i = 0; i < 64; ++i) {
            float uc = rng.next();
            Vec2 uv;
            uv.x = rng.next();
            uv.y = rng.next();
            BsdfSample s;
            if (sample(wo, uv, uc, &s)) {
                float ci = std::fabs(s.wi.z);
                if (ci > 0.0f && s.pdf > 0.0f) sum += s.f * ci / s.pdf;
            }
        }
        return sum / 64.0f;
    }
};

// ---------------------------------------------------------------- bsdf/conductor.rs
struct Complex {
    float re, im;
    float norm() const { return re * re + im * im; }
    Complex sqrt() const {  // conductor.rs:30-36: polar form
        float r = std::sqrt(re * re + im * im);
        float theta = std::atan2(im, re);
        float sr = std::sqrt(r), ht = theta * 0.5f;
        return Complex{sr * std::cos(ht), sr * std::sin(ht)};
    }
};
inline Complex operator+(Complex a, Complex b) { return {a.re + b.re, a.im + b.im}; }
inline Complex operator-(Complex a, Complex b) { return {a.re - b.re, a.im - b.im}; }
inline Complex operator*(Complex a, Complex b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Complex operator*(Complex a, float s) { return {a.re * s, a.im * s}; }
inline Complex operator/(Complex a, Complex b) {
    float denom = b.re * b.re + b.im * b.im;
    if (denom == 0.0f) return {0.0f, 0.0f};
    return {(a.re * b.re + a.im * b.im) / denom, (a.im * b.re - a.re * b.im) / denom};
}
// fresnel_complex (conductor.rs:91-123)
inline SampledSpectrum fresnel_complex(float cos_theta_i, const SampledSpectrum& eta, const SampledSpectrum& k) {
    cos_theta_i = clampf(cos_theta_i, 0.0f, 1.0f);
    SampledSpectrum r;
    for (int i = 0; i < NS; ++i) {
        Complex ce{eta.v[i], k.v[i]};
        float sin2_i = 1.0f - cos_theta_i * cos_theta_i;
        Complex sin2_t = Complex{sin2_i, 0.0f} / (ce * ce);
        Complex cos_t = (Complex{1.0f, 0.0f} - sin2_t).sqrt();
        Complex r_parl = (ce * cos_theta_i - cos_t) / (ce * cos_theta_i + cos_t);
        Complex r_perp = (Complex{cos_theta_i, 0.0f} - ce * cos_t) / (Complex{cos_theta_i, 0.0f} + ce * cos_t);
        r.v[i] = (r_parl.norm() + r_perp.norm()) * 0.5f;
    }
    return r;
}
struct ConductorBsdf {  // conductor.rs:125-439; the GGX helpers are the same functions as dielectric.rs / generalized_schlick.rs
    SampledSpectrum eta, k;
    Ggx g;
    SampledSpectrum torrance_sparrow(Vec3 wo, Vec3 wi, Vec3 wm) const {
        float co = std::fabs(wo.z), ci = std::fabs(wi.z);
        if (co == 0.0f || ci == 0.0f) return SampledSpectrum::zero();
        SampledSpectrum fr = fresnel_complex(std::fabs(dot(wo, wm)), eta, k);
        float d = g.D(wm), gg = g.G(wo, wi);
        return fr * d * gg / (4.0f * co);
    }
    float pdf_microfacet(Vec3 wo, Vec3 wi) const {
        if (!same_hemisphere(wo, wi)) return 0.0f;
        Vec3 wm;
        if (!half_vector(wo, wi, &wm)) return 0.0f;
        float vis = g.Dvis(wo, wm);
        float jac = 4.0f * std::fabs(dot(wo, wm));
        if (jac == 0.0f) return 0.0f;
        return vis / jac;
    }
    bool sample(Vec3 wo, Vec2 uv, BsdfSample* out) const {
        if (wo.z == 0.0f) return false;
        if (g.effectively_smooth()) {
            Vec3 wi(-wo.x, -wo.y, wo.z);
            if (wi.z == 0.0f) return false;
            *out = BsdfSample{fresnel_complex(std::fabs(wi.z), eta, k), wi, 1.0f, ST_SPECULAR_REFLECTION};
            return true;
        }
        Vec3 wm = g.sample_wm(wo, uv);
        Vec3 wi = reflect(wo, wm);
        if (!same_hemisphere(wo, wi)) return false;
        *out = BsdfSample{torrance_sparrow(wo, wi, wm), wi, pdf_microfacet(wo, wi), ST_GLOSSY_REFLECTION};
        return true;
    }
    SampledSpectrum evaluate(Vec3 wo, Vec3 wi) const {
        if (g.effectively_smooth()) return SampledSpectrum::zero();
        float co = std::fabs(wo.z), ci = std::fabs(wi.z);
        if (co == 0.0f || ci == 0.0f) return SampledSpectrum::zero();
        if (!same_hemisphere(wo, wi)) return SampledSpectrum::zero();
        Vec3 wm;
        if (!half_vector(wo, wi, &wm)) return SampledSpectrum::zero();
        return torrance_sparrow(wo, wi, wm);
    }
    float pdf(Vec3 wo, Vec3 wi) const { return g.effectively_smooth() ? 0.0f : pdf_microfacet(wo, wi); }
};

// ---------------------------------------------------------------- math/src/transform.rs:216-244
struct NormalMapFrame {
    Mat4 to_nm, from_nm;
    explicit NormalMapFrame(Vec3 normal_map_normal) {
        Vec3 z = normalize(normal_map_normal);
        Vec3 cand = std::fabs(dot(z, Vec3(1, 0, 0))) < 0.9f ? Vec3(1, 0, 0) : Vec3(0, 1, 0);
        Vec3 x = normalize(cand - dot(z, cand) * z);
        Vec3 y = normalize(cross(z, x));
        Mat4 m = Mat4::from_cols3(x, y, z);
        to_nm = inverse(m);
        from_nm = inverse(to_nm);
    }
};

struct Material {
    int type = MAT_LAMBERT;
    SpectrumParam color;  // albedo / radiance / plastic colour / base colour
    FloatParam intensity;
    NormalParam normal;
    float eta = 1.5f;
    bool thin_surface = false;
    FloatParam roughness, metallic, ior, coat_ior, coat_roughness, coat_thickness;
    SpectrumParam coat_tint;
};

struct MaterialContext {
    const Tables* T;
    const std::vector<Texture>* textures;
    uint32_t aux_base = 0;  // aux streams for directional_albedo: one per (path, bounce, call site)
    uint32_t depth = 0;
};

inline Spectrum sample_spectrum_param(const MaterialContext& c, const SpectrumParam& p, Vec2 uv) {
    if (!p.is_texture) return p.spectrum;
    float rgb[3];
    bilinear_sample_rgb((*c.textures)[p.texture], uv, rgb);  // rgb_texture.rs:48-66 (sRGB-typed colour -> albedo spectrum)
    return make_rgb_albedo(*c.T, Vec3(rgb[0], rgb[1], rgb[2]), true);
}
inline float sample_float_param(const MaterialContext& c, const FloatParam& p, Vec2 uv) {
    if (!p.is_texture) return p.value;
    float v = bilinear_sample_gray((*c.textures)[p.texture], uv);
    return p.gamma_corrected ? srgb_inverse_eotf(v) : v;  // float_texture.rs:45-52
}
// NormalParameter::sample + unwrap_or(Normal::new(0,0,1)) (normal_texture.rs:40-66)
inline Vec3 sample_normal_param(const MaterialContext& c, const NormalParam& p, Vec2 uv) {
    if (p.texture < 0) return make_normal(Vec3(0, 0, 1));
    float rgb[3];
    bilinear_sample_rgb((*c.textures)[p.texture], uv, rgb);
    float x = rgb[0] * 2.0f - 1.0f, y = rgb[1] * 2.0f - 1.0f, z = rgb[2] * 2.0f - 1.0f;
    if (p.flip_y) y = -y;
    float len = std::sqrt(x * x + y * y + z * z);
    if (len > 0.0f) { x /= len; y /= len; z /= len; return make_normal(Vec3(x, y, z)); }
    return make_normal(Vec3(0, 0, 1));
}

// ----- SimplePbr helpers shared by simple_pbr_material.rs:274-537 and simple_pbr_clearcoat_material.rs:540-845 (identical code)
struct PbrBase {
    SampledSpectrum base_color;
    float metallic, roughness, ior;
    static float r0_of(float ior) { float r = (ior - 1.0f) / (ior + 1.0f); return r * r; }
    GeneralizedSchlickBsdf metal_bsdf(float alpha, bool for_pdf) const {
        return GeneralizedSchlickBsdf{for_pdf ? SampledSpectrum::constant(1.0f) : base_color, SampledSpectrum::constant(1.0f), 5.0f, SampledSpectrum::constant(1.0f), Ggx{alpha, alpha}};
    }
    GeneralizedSchlickBsdf diel_bsdf(float alpha) const {
        return GeneralizedSchlickBsdf{SampledSpectrum::constant(r0_of(ior)), SampledSpectrum::constant(1.0f), 5.0f, SampledSpectrum::constant(1.0f), Ggx{alpha, alpha}};
    }
    MaterialSample sample_metallic(float alpha, Vec3 wo, Vec2 uv, const Mat4& from_nm) const {
        BsdfSample s;
        if (!metal_bsdf(alpha, false).sample(wo, uv, 0.0f, &s)) return MaterialSample{};
        return make_sample(s.f, transform_vector3(from_nm, s.wi), s.pdf, s.sample_type);
    }
    MaterialSample sample_dielectric(float alpha, Vec3 wo, float uc, Vec2 uv, const Mat4& from_nm) const {
        GeneralizedSchlickBsdf gs = diel_bsdf(alpha);
        float fresnel = gs.fresnel(wo).average();
        BsdfSample s;
        if (uc < fresnel) {
            float uc2 = uc / fresnel;
            if (!gs.sample(wo, uv, uc2, &s)) return MaterialSample{};
            return make_sample(s.f, transform_vector3(from_nm, s.wi), s.pdf * fresnel, s.sample_type);
        }
        LambertBsdf lb{base_color};
        if (!lb.sample(wo, uv, &s)) return MaterialSample{};
        return make_sample(s.f * (1.0f - fresnel), transform_vector3(from_nm, s.wi), s.pdf * (1.0f - fresnel), s.sample_type);
    }
    MaterialSample sample(Vec3 wo, float uc, Vec2 uv, const Mat4& from_nm) const {
        float alpha = roughness * roughness;
        if (metallic >= 1.0f) return sample_metallic(alpha, wo, uv, from_nm);
        if (metallic <= 0.0f) return sample_dielectric(alpha, wo, uc, uv, from_nm);
        if (uc <= metallic) return sample_metallic(alpha, wo, uv, from_nm);
        return sample_dielectric(alpha, wo, (uc - metallic) / (1.0f - metallic), uv, from_nm);
    }
    SampledSpectrum eval_metallic(float alpha, Vec3 wo, Vec3 wi) const { return metal_bsdf(alpha, false).evaluate(wo, wi); }
    SampledSpectrum eval_dielectric(float alpha, Vec3 wo, Vec3 wi) const {
        GeneralizedSchlickBsdf gs = diel_bsdf(alpha);
        SampledSpectrum direct = gs.evaluate(wo, wi);
        float fresnel = gs.fresnel(wo).average();
        SampledSpectrum lam = LambertBsdf{base_color}.evaluate(wo, wi);
        return direct + (1.0f - fresnel) * lam;
    }
    SampledSpectrum evaluate(Vec3 wo, Vec3 wi) const {
        float alpha = roughness * roughness;
        if (metallic >= 1.0f) return eval_metallic(alpha, wo, wi);
        if (metallic <= 0.0f) return eval_dielectric(alpha, wo, wi);
        return eval_metallic(alpha, wo, wi) * metallic + eval_dielectric(alpha, wo, wi) * (1.0f - metallic);
    }
    float pdf_metallic(float alpha, Vec3 wo, Vec3 wi) const { return metal_bsdf(alpha, true).pdf(wo, wi); }
    float pdf_dielectric(float alpha, Vec3 wo, Vec3 wi) const {
        GeneralizedSchlickBsdf gs = diel_bsdf(alpha);
        float direct = gs.pdf(wo, wi);
        float fresnel = gs.fresnel(wo).average();
        float lam = LambertBsdf{base_color}.pdf(wo, wi);
        return fresnel * direct + (1.0f - fresnel) * lam;
    }
    float pdf(Vec3 wo, Vec3 wi) const {
        float alpha = roughness * roughness;
        if (metallic >= 1.0f) return pdf_metallic(alpha, wo, wi);
        if (metallic <= 0.0f) return pdf_dielectric(alpha, wo, wi);
        return pdf_metallic(alpha, wo, wi) * metallic + pdf_dielectric(alpha, wo, wi) * (1.0f - metallic);
    }
};

// Beer–Lambert coat attenuation (simple_pbr_clearcoat_material.rs:88-107)
inline SampledSpectrum coat_attenuation(const SampledSpectrum& tint, float thickness, float cos_theta) {
    SampledSpectrum log_tint = tint.log();
    SampledSpectrum sigma = (-1.0f * log_tint) / 0.001f;
    float thickness_m = thickness * 0.001f;
    float l = thickness_m / rmax(cos_theta, 1e-4f);
    return ((-1.0f * sigma) * l).exp();
}

struct ClearcoatParams {
    float ior, roughness, thickness;
    SampledSpectrum tint;
    GeneralizedSchlickBsdf bsdf() const {
        float a = roughness * roughness;
        return GeneralizedSchlickBsdf{SampledSpectrum::constant(PbrBase::r0_of(ior)), SampledSpectrum::constant(1.0f), 5.0f, SampledSpectrum::constant(1.0f), Ggx{a, a}};
    }
};

inline PbrBase load_pbr_base(const MaterialContext& c, const Material& m, Vec2 uv, const SampledWavelengths& wl) {
    PbrBase b;
    b.base_color = sample_spectrum_param(c, m.color, uv).sample(*c.T, wl);
    b.metallic = sample_float_param(c, m.metallic, uv);
    b.roughness = sample_float_param(c, m.roughness, uv);
    b.ior = sample_float_param(c, m.ior, uv);
    return b;
}
inline ClearcoatParams load_coat(const MaterialContext& c, const Material& m, Vec2 uv, const SampledWavelengths& wl) {
    ClearcoatParams p;
    p.ior = sample_float_param(c, m.coat_ior, uv);
    p.roughness = sample_float_param(c, m.coat_roughness, uv);
    p.tint = sample_spectrum_param(c, m.coat_tint, uv).sample(*c.T, wl);
    p.thickness = sample_float_param(c, m.coat_thickness, uv);
    return p;
}

// BsdfSurfaceMaterial::sample (material/traits.rs:29-47)
inline MaterialSample material_sample(const MaterialContext& c, const Material& m, float uc, Vec2 uv, SampledWavelengths& wl, Vec3 wo, const TangentShadingPoint& sp) {
    NormalMapFrame fr(sample_normal_param(c, m.normal, sp.uv));
    Vec3 wo_nm = transform_vector3(fr.to_nm, wo);
    switch (m.type) {
        case MAT_LAMBERT: {  // lambert_material.rs:42-97
            SampledSpectrum albedo = sample_spectrum_param(c, m.color, sp.uv).sample(*c.T, wl);
            BsdfSample s;
            if (!LambertBsdf{albedo}.sample(wo_nm, uv, &s)) return MaterialSample{};
            Vec3 wi_sh = transform_vector3(fr.from_nm, s.wi);
            if (signum(dot(sp.normal, wi_sh)) != signum(dot(sp.normal, wo))) return MaterialSample{};
            return make_sample(s.f, wi_sh, s.pdf, s.sample_type);
        }
        case MAT_PLASTIC: {  // plastic_material.rs:122-187
            SampledSpectrum eta = SampledSpectrum::constant(m.eta);
            float rough = sample_float_param(c, m.roughness, sp.uv);
            bool entering = dot(sp.normal, wo) > 0.0f;
            DielectricBsdf d(eta, entering, m.thin_surface, rough, rough);  // roughness passed UNSQUARED as alpha
            BsdfSample s;
            if (!d.sample(wo_nm, uv, uc, wl, &s)) return MaterialSample{};
            if (dot(s.wi, wo_nm) < 0.0f) s.f *= sample_spectrum_param(c, m.color, uv).sample(*c.T, wl);  // quirk q21: filter looked up at the RANDOM uv
            return make_sample(s.f, transform_vector3(fr.from_nm, s.wi), s.pdf, s.sample_type);
        }
        case MAT_SIMPLE_PBR: {  // simple_pbr_material.rs:78-150
            PbrBase b = load_pbr_base(c, m, sp.uv, wl);
            return b.sample(wo_nm, uc, uv, fr.from_nm);
        }
        case MAT_METAL: {  // metal_material.rs:122-173: eta/k presets in `color` / `coat_tint`, alpha = roughness^2
            SampledSpectrum eta = m.color.spectrum.sample(*c.T, wl), k = m.coat_tint.spectrum.sample(*c.T, wl);
            float rough = sample_float_param(c, m.roughness, sp.uv);
            float alpha = rough * rough;
            ConductorBsdf cb{eta, k, Ggx{alpha, alpha}};
            BsdfSample s;
            if (!cb.sample(wo_nm, uv, &s)) return MaterialSample{};
            Vec3 wi_sh = transform_vector3(fr.from_nm, s.wi);
            if (signum(dot(sp.normal, wi_sh)) != signum(dot(sp.normal, wo))) return MaterialSample{};
            return make_sample(s.f, wi_sh, s.pdf, s.sample_type);
        }
        case MAT_GLASS: {  // glass_material.rs:97-147: DielectricBsdf with the glass's eta(lambda), roughness passed unsquared
            SampledSpectrum eta = m.color.spectrum.sample(*c.T, wl);
            float rough = sample_float_param(c, m.roughness, sp.uv);
            bool entering = dot(sp.normal, wo) > 0.0f;
            DielectricBsdf d(eta, entering, m.thin_surface, rough, rough);
            BsdfSample s;
            if (!d.sample(wo_nm, uv, uc, wl, &s)) return MaterialSample{};
            return make_sample(s.f, transform_vector3(fr.from_nm, s.wi), s.pdf, s.sample_type);
        }
        case MAT_CLEARCOAT_PBR: {  // simple_pbr_clearcoat_material.rs:121-250
            PbrBase b = load_pbr_base(c, m, sp.uv, wl);
            ClearcoatParams cp = load_coat(c, m, sp.uv, wl);
            if (cp.thickness <= 0.0f) return b.sample(wo_nm, uc, uv, fr.from_nm);
            GeneralizedSchlickBsdf coat = cp.bsdf();
            float fc = coat.directional_albedo(wo_nm, aux_rng(c.aux_base, c.depth, 0)).average();
            if (uc < fc) {
                float uc2 = uc / fc;
                BsdfSample s;
                if (!coat.sample(wo_nm, uv, uc2, &s)) return MaterialSample{};
                return make_sample(s.f, transform_vector3(fr.from_nm, s.wi), s.pdf * fc, s.sample_type);
            }
            float uc2 = (uc - fc) / (1.0f - fc);
            MaterialSample sub = b.sample(wo_nm, uc2, uv, fr.from_nm);
            if (!sub.is_sampled) return sub;
            SampledSpectrum att = coat_attenuation(cp.tint, cp.thickness, wo_nm.z) * coat_attenuation(cp.tint, cp.thickness, sub.wi.z);
            return make_sample(sub.f * att, sub.wi, sub.pdf * (1.0f - fc), sub.sample_type);
        }
        default: return MaterialSample{};
    }
}

// BsdfSurfaceMaterial::evaluate (f only; the reference's `pdf: 1.0` field is never read)
// BsdfSurfaceMaterial::sample_albedo_spectrum (lambert_material.rs:172-178, plastic_material.rs:266-273, simple_pbr_material.rs:259-266,
// simple_pbr_clearcoat_material.rs:435-442, metal_material.rs:267-278, glass_material.rs:224-231)
inline SampledSpectrum material_albedo(const MaterialContext& c, const Material& m, const SampledWavelengths& wl, Vec2 uv) {
    switch (m.type) {
        case MAT_LAMBERT: case MAT_SIMPLE_PBR: case MAT_CLEARCOAT_PBR: return sample_spectrum_param(c, m.color, uv).sample(*c.T, wl);
        case MAT_METAL: return fresnel_complex(1.0f, m.color.spectrum.sample(*c.T, wl), m.coat_tint.spectrum.sample(*c.T, wl));
        case MAT_PLASTIC: case MAT_GLASS: return SampledSpectrum::constant(1.0f);
        default: return SampledSpectrum::zero();
    }
}

inline SampledSpectrum material_evaluate(const MaterialContext& c, const Material& m, const SampledWavelengths& wl, Vec3 wo, Vec3 wi, const TangentShadingPoint& sp) {
    NormalMapFrame fr(sample_normal_param(c, m.normal, sp.uv));
    Vec3 wo_nm = transform_vector3(fr.to_nm, wo), wi_nm = transform_vector3(fr.to_nm, wi);
    switch (m.type) {
        case MAT_LAMBERT: {  // lambert_material.rs:99-136
            SampledSpectrum albedo = sample_spectrum_param(c, m.color, sp.uv).sample(*c.T, wl);
            if (signum(dot(sp.normal, wi)) != signum(dot(sp.normal, wo))) return SampledSpectrum::zero();
            return LambertBsdf{albedo}.evaluate(wo_nm, wi_nm);
        }
        case MAT_PLASTIC: {  // plastic_material.rs:189-228
            float rough = sample_float_param(c, m.roughness, sp.uv);
            bool entering = dot(sp.normal, wo) > 0.0f;
            DielectricBsdf d(SampledSpectrum::constant(m.eta), entering, m.thin_surface, rough, rough);
            SampledSpectrum f = d.evaluate(wo_nm, wi_nm);
            if (dot(wi_nm, wo_nm) < 0.0f) f *= sample_spectrum_param(c, m.color, sp.uv).sample(*c.T, wl);
            return f;
        }
        case MAT_SIMPLE_PBR: return load_pbr_base(c, m, sp.uv, wl).evaluate(wo_nm, wi_nm);
        case MAT_METAL: {  // metal_material.rs:175-213
            if (signum(dot(sp.normal, wi)) != signum(dot(sp.normal, wo))) return SampledSpectrum::zero();
            float rough = sample_float_param(c, m.roughness, sp.uv);
            float alpha = rough * rough;
            return ConductorBsdf{m.color.spectrum.sample(*c.T, wl), m.coat_tint.spectrum.sample(*c.T, wl), Ggx{alpha, alpha}}.evaluate(wo_nm, wi_nm);
        }
        case MAT_GLASS: {  // glass_material.rs:149-185
            float rough = sample_float_param(c, m.roughness, sp.uv);
            bool entering = dot(sp.normal, wo) > 0.0f;
            return DielectricBsdf(m.color.spectrum.sample(*c.T, wl), entering, m.thin_surface, rough, rough).evaluate(wo_nm, wi_nm);
        }
        case MAT_CLEARCOAT_PBR: {  // simple_pbr_clearcoat_material.rs:252-353
            PbrBase b = load_pbr_base(c, m, sp.uv, wl);
            ClearcoatParams cp = load_coat(c, m, sp.uv, wl);
            if (cp.thickness <= 0.0f) return b.evaluate(wo_nm, wi_nm);
            GeneralizedSchlickBsdf coat = cp.bsdf();
            float fc = coat.directional_albedo(wo_nm, aux_rng(c.aux_base, c.depth, 1)).average();
            SampledSpectrum coat_f = coat.evaluate(wo_nm, wi_nm);
            SampledSpectrum sub_f = b.evaluate(wo_nm, wi_nm);
            SampledSpectrum att = coat_attenuation(cp.tint, cp.thickness, wo_nm.z) * coat_attenuation(cp.tint, cp.thickness, wi_nm.z);
            return coat_f * fc + sub_f * att * (1.0f - fc);
        }
        default: return SampledSpectrum::zero();
    }
}

inline float material_pdf(const MaterialContext& c, const Material& m, const SampledWavelengths& wl, Vec3 wo, Vec3 wi, const TangentShadingPoint& sp) {
    NormalMapFrame fr(sample_normal_param(c, m.normal, sp.uv));
    Vec3 wo_nm = transform_vector3(fr.to_nm, wo), wi_nm = transform_vector3(fr.to_nm, wi);
    switch (m.type) {
        case MAT_LAMBERT: {  // lambert_material.rs:138-170
            if (signum(dot(sp.normal, wi)) != signum(dot(sp.normal, wo))) return 0.0f;
            return LambertBsdf{SampledSpectrum::zero()}.pdf(wo_nm, wi_nm);
        }
        case MAT_PLASTIC: {
            float rough = sample_float_param(c, m.roughness, sp.uv);
            bool entering = dot(sp.normal, wo) > 0.0f;
            return DielectricBsdf(SampledSpectrum::constant(m.eta), entering, m.thin_surface, rough, rough).pdf(wo_nm, wi_nm);
        }
        case MAT_SIMPLE_PBR: return load_pbr_base(c, m, sp.uv, wl).pdf(wo_nm, wi_nm);
        case MAT_METAL: {  // metal_material.rs:215-252
            if (signum(dot(sp.normal, wi)) != signum(dot(sp.normal, wo))) return 0.0f;
            float rough = sample_float_param(c, m.roughness, sp.uv);
            float alpha = rough * rough;
            return ConductorBsdf{m.color.spectrum.sample(*c.T, wl), m.coat_tint.spectrum.sample(*c.T, wl), Ggx{alpha, alpha}}.pdf(wo_nm, wi_nm);
        }
        case MAT_GLASS: {  // glass_material.rs:187-221
            float rough = sample_float_param(c, m.roughness, sp.uv);
            bool entering = dot(sp.normal, wo) > 0.0f;
            return DielectricBsdf(m.color.spectrum.sample(*c.T, wl), entering, m.thin_surface, rough, rough).pdf(wo_nm, wi_nm);
        }
        case MAT_CLEARCOAT_PBR: {  // simple_pbr_clearcoat_material.rs:355-433
            PbrBase b = load_pbr_base(c, m, sp.uv, wl);
            ClearcoatParams cp = load_coat(c, m, sp.uv, wl);
            if (cp.thickness <= 0.0f) return b.pdf(wo_nm, wi_nm);
            GeneralizedSchlickBsdf coat = cp.bsdf();
            float fc = coat.directional_albedo(wo_nm, aux_rng(c.aux_base, c.depth, 2)).average();
            return coat.pdf(wo_nm, wi_nm) * fc + b.pdf(wo_nm, wi_nm) * (1.0f - fc);
        }
        default: return 0.0f;
    }
}

}  // namespace orc
