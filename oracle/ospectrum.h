// ORACLE (test infrastructure).  Spectral layer: SampledSpectrum / SampledWavelengths, dense tables, rgb->spectrum lookup,
// sRGB gamut + EOTF.  Follows /root/reference/spectrum/src/{sampled_spectrum,rgb_sigmoid_polynomial,spectrum}.rs,
// spectrum/src/spectrum/{constant,densely_sampled,rgb_albedo,rgb_illuminant}_spectrum.rs and color/src/{gamut,eotf}.rs.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "omath.h"

namespace orc {

constexpr int NS = 4;  // N_SPECTRUM_SAMPLES (sampled_spectrum.rs:9)
constexpr float LAMBDA_MIN = 360.0f, LAMBDA_MAX = 830.0f;
constexpr int N_DENSE = 470;

struct SampledSpectrum {
    float v[NS];
    static SampledSpectrum constant(float c) { SampledSpectrum s; for (int i = 0; i < NS; ++i) s.v[i] = c; return s; }
    static SampledSpectrum zero() { return constant(0.0f); }
    static SampledSpectrum one() { return constant(1.0f); }
    bool is_zero() const { for (int i = 0; i < NS; ++i) if (v[i] != 0.0f) return false; return true; }
    bool is_constant() const { for (int i = 0; i < NS; ++i) if (!(v[i] == v[0])) return false; return true; }
    // fold(NEG_INFINITY, f32::max) (sampled_spectrum.rs:251-256)
    float max_value() const { float m = -INFINITY; for (int i = 0; i < NS; ++i) m = rmax(m, v[i]); return m; }
    // iter().sum() / 4 (sampled_spectrum.rs:259-262)
    float average() const { float s = 0.0f; for (int i = 0; i < NS; ++i) s += v[i]; return s / (float)NS; }
    SampledSpectrum sqrt() const { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = std::sqrt(v[i]); return r; }
    SampledSpectrum clamp(float lo, float hi) const { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = clampf(v[i], lo, hi); return r; }
    SampledSpectrum exp() const { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = std::exp(v[i]); return r; }
    // log with max(1e-10) (sampled_spectrum.rs:223-229)
    SampledSpectrum log() const { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = std::log(rmax(v[i], 1e-10f)); return r; }
};
inline SampledSpectrum operator+(const SampledSpectrum& a, const SampledSpectrum& b) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
inline SampledSpectrum operator-(const SampledSpectrum& a, const SampledSpectrum& b) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = a.v[i] - b.v[i]; return r; }
inline SampledSpectrum operator*(const SampledSpectrum& a, const SampledSpectrum& b) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = a.v[i] * b.v[i]; return r; }
inline SampledSpectrum operator*(const SampledSpectrum& a, float s) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = a.v[i] * s; return r; }
inline SampledSpectrum operator*(float s, const SampledSpectrum& a) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = s * a.v[i]; return r; }
// division by zero yields zero (sampled_spectrum.rs:59-83)
inline SampledSpectrum operator/(const SampledSpectrum& a, float s) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = (s == 0.0f) ? 0.0f : a.v[i] / s; return r; }
inline SampledSpectrum operator/(const SampledSpectrum& a, const SampledSpectrum& b) { SampledSpectrum r; for (int i = 0; i < NS; ++i) r.v[i] = (b.v[i] == 0.0f) ? 0.0f : a.v[i] / b.v[i]; return r; }
inline SampledSpectrum& operator+=(SampledSpectrum& a, const SampledSpectrum& b) { for (int i = 0; i < NS; ++i) a.v[i] += b.v[i]; return a; }
inline SampledSpectrum& operator*=(SampledSpectrum& a, const SampledSpectrum& b) { for (int i = 0; i < NS; ++i) a.v[i] *= b.v[i]; return a; }
inline SampledSpectrum& operator*=(SampledSpectrum& a, float s) { for (int i = 0; i < NS; ++i) a.v[i] *= s; return a; }
// `/=` by scalar zero is a no-op (sampled_spectrum.rs:107-115)
inline SampledSpectrum& operator/=(SampledSpectrum& a, float s) { if (s != 0.0f) for (int i = 0; i < NS; ++i) a.v[i] /= s; return a; }

// sampled_spectrum.rs:304-366
struct SampledWavelengths {
    float lambda[NS];
    float pdf[NS];
    static SampledWavelengths new_uniform(float u) {
        SampledWavelengths r;
        for (int i = 0; i < NS; ++i) r.pdf[i] = 1.0f / (LAMBDA_MAX - LAMBDA_MIN);
        r.lambda[0] = LAMBDA_MIN + u * (LAMBDA_MAX - LAMBDA_MIN);
        float delta = (LAMBDA_MAX - LAMBDA_MIN) / (float)NS;
        for (int i = 1; i < NS; ++i) {
            r.lambda[i] = r.lambda[i - 1] + delta;
            if (r.lambda[i] >= LAMBDA_MAX) r.lambda[i] = LAMBDA_MIN + (r.lambda[i] - LAMBDA_MAX);
        }
        return r;
    }
    bool is_secondary_terminated() const { for (int i = 1; i < NS; ++i) if (pdf[i] != 0.0f) return false; return true; }
    void terminate_secondary() {
        if (is_secondary_terminated()) return;
        for (int i = 1; i < NS; ++i) pdf[i] = 0.0f;
        pdf[0] /= (float)NS;
    }
};

// sRGB EOTF (color/src/eotf.rs:51-72)
inline float srgb_inverse_eotf(float c) { return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }
inline float srgb_oetf(float c) { return c <= 0.0031308f ? 12.92f * c : 1.055f * std::pow(c, 1.0f / 2.4f) - 0.055f; }

// glam Mat3 (column major) pieces needed by color/src/gamut.rs:29-39
struct Mat3 { Vec3 c0, c1, c2; };
inline Vec3 mul(const Mat3& m, Vec3 v) { return (m.c0 * v.x + m.c1 * v.y) + m.c2 * v.z; }
inline Mat3 inverse(const Mat3& m) {
    Vec3 t0 = cross(m.c1, m.c2), t1 = cross(m.c2, m.c0), t2 = cross(m.c0, m.c1);
    float det = dot(m.c2, t2);
    float id = 1.0f / det;
    Vec3 a = t0 * id, b = t1 * id, c = t2 * id;
    return Mat3{Vec3(a.x, b.x, c.x), Vec3(a.y, b.y, c.y), Vec3(a.z, b.z, c.z)};  // transpose
}
inline Mat3 mul(const Mat3& a, const Mat3& b) { return Mat3{mul(a, b.c0), mul(a, b.c1), mul(a, b.c2)}; }
inline Vec3 xy_to_xyz(float x, float y) {
    if (y == 0.0f) return Vec3(0, 0, 0);
    return Vec3(x * 1.0f / y, 1.0f, (1.0f - x - y) * 1.0f / y);
}
// GamutSrgb::new (gamut.rs:43-69)
inline void srgb_matrices(Mat3* rgb_to_xyz, Mat3* xyz_to_rgb) {
    Vec3 r = xy_to_xyz(0.6400f, 0.3300f), g = xy_to_xyz(0.3000f, 0.6000f), b = xy_to_xyz(0.1500f, 0.0600f), w = xy_to_xyz(0.3127f, 0.3290f);
    Mat3 rgb{r, g, b};
    Vec3 c = mul(inverse(rgb), w);
    Mat3 diag{Vec3(c.x, 0, 0), Vec3(0, c.y, 0), Vec3(0, 0, c.z)};
    Mat3 m = mul(rgb, diag);
    *rgb_to_xyz = m;
    *xyz_to_rgb = inverse(m);
}

// All read-only tables of the hot path.
struct Tables {
    uint32_t sobol[104];
    float cie_x[N_DENSE], cie_y[N_DENSE], cie_z[N_DENSE], d65[N_DENSE];
    float d65_max = 0.0f;
    float z_nodes[64];
    const float* rgb2spec = nullptr;  // [3][64][64][64][3]
    std::vector<float> presets;       // n x 470: metal eta/k and glass eta tables, densely resampled (presets.rs:129-200)
    Mat3 xyz_to_rgb, rgb_to_xyz;
};

// DenselySampledSpectrum::value (densely_sampled_spectrum.rs:57-67)
inline float dense_value(const float* tab, float lambda) {
    if (!(lambda >= LAMBDA_MIN && lambda <= LAMBDA_MAX)) return 0.0f;
    uint32_t index = f2u_sat(std::floor(lambda - LAMBDA_MIN));
    return index < (uint32_t)N_DENSE ? tab[index] : 0.0f;
}

// rgb_sigmoid_polynomial.rs:19-27
inline float sigmoid(float x) { return 1.0f / (1.0f + std::exp(-x)); }
inline float parabolic(float t, const float c[3]) { return t * t * c[0] + t * c[1] + c[2]; }

struct Rgb2SpecIndex { int m, xi, yi, zi; };

// RgbToSpectrumTable::get (rgb_sigmoid_polynomial.rs:87-155).  `gamma_encoded` selects ColorSrgb (true) vs ColorSrgbLinear.
// Returns false where the reference panics (component > 1).
inline bool rgb_to_coeffs(const Tables& T, Vec3 rgb_in, bool gamma_encoded, float cs[3], Rgb2SpecIndex* idx = nullptr) {
    Vec3 rgb = rgb_in;
    if (gamma_encoded) rgb = Vec3(srgb_inverse_eotf(rgb.x), srgb_inverse_eotf(rgb.y), srgb_inverse_eotf(rgb.z));
    rgb = vmax(rgb, Vec3(0, 0, 0));
    if (max_element(rgb) > 1.0f) return false;
    if (idx) *idx = Rgb2SpecIndex{-1, 0, 0, 0};
    if (rgb.x == rgb.y && rgb.y == rgb.z) {
        cs[0] = 0.0f; cs[1] = 0.0f; cs[2] = std::log(rgb.x / (1.0f - rgb.x));
        return true;
    }
    int m = max_position(rgb);
    float z = rgb[m];
    float x = rgb[(m + 1) % 3] * (64.0f - 1.0f) / z;
    float y = rgb[(m + 2) % 3] * (64.0f - 1.0f) / z;
    uint32_t xi = f2u_sat(x); if (xi > 62) xi = 62;
    uint32_t yi = f2u_sat(y); if (yi > 62) yi = 62;
    uint32_t zi = 62;
    for (uint32_t i = 0; i <= 62; ++i) if (T.z_nodes[i + 1] > z) { zi = i; break; }
    float dx = x - (float)xi, dy = y - (float)yi;
    float dz = (z - T.z_nodes[zi]) / (T.z_nodes[zi + 1] - T.z_nodes[zi]);
    if (idx) *idx = Rgb2SpecIndex{m, (int)xi, (int)yi, (int)zi};
    auto lerp = [](float a, float b, float t) { return a + (b - a) * t; };
    for (int i = 0; i < 3; ++i) {
        auto co = [&](uint32_t ddx, uint32_t ddy, uint32_t ddz) {
            return T.rgb2spec[(((((size_t)m * 64 + (zi + ddz)) * 64 + (yi + ddy)) * 64 + (xi + ddx)) * 3) + i];
        };
        cs[i] = lerp(lerp(lerp(co(0, 0, 0), co(1, 0, 0), dx), lerp(co(0, 1, 0), co(1, 1, 0), dx), dy),
                     lerp(lerp(co(0, 0, 1), co(1, 0, 1), dx), lerp(co(0, 1, 1), co(1, 1, 1), dx), dy), dz);
    }
    return true;
}

// Tagged union over the Spectrum kinds that reach the hot path.
enum SpectrumKind : int { SPEC_CONSTANT = 0, SPEC_RGB_ALBEDO = 1, SPEC_RGB_ILLUMINANT = 2, SPEC_D65 = 3, SPEC_PRESET = 5 };
struct Spectrum {
    int kind = SPEC_CONSTANT;
    float c = 0.0f;                      // constant
    float coef[3] = {0, 0, 0};           // sigmoid polynomial
    float scale = 1.0f;                  // illuminant
    int table = 0;                       // SPEC_PRESET: index of a DenselySampledSpectrum preset (presets::au_eta() ...)
    // SpectrumTrait::value
    float value(const Tables& T, float lambda) const {
        switch (kind) {
            case SPEC_CONSTANT: return c;
            case SPEC_RGB_ALBEDO: {
                float t = (lambda - LAMBDA_MIN) / (LAMBDA_MAX - LAMBDA_MIN);
                return sigmoid(parabolic(t, coef));
            }
            case SPEC_RGB_ILLUMINANT: {
                float t = (lambda - LAMBDA_MIN) / (LAMBDA_MAX - LAMBDA_MIN);
                return scale * sigmoid(parabolic(t, coef)) * dense_value(T.d65, lambda);  // rgb_illuminant_spectrum.rs:43-45
            }
            case SPEC_PRESET: return dense_value(T.presets.data() + (size_t)table * N_DENSE, lambda);
            default: return dense_value(T.d65, lambda);
        }
    }
    // SpectrumTrait::sample (spectrum.rs:39-53)
    SampledSpectrum sample(const Tables& T, const SampledWavelengths& l) const {
        SampledSpectrum s = SampledSpectrum::zero();
        if (l.is_secondary_terminated()) { s.v[0] = value(T, l.lambda[0]); return s; }
        for (int i = 0; i < NS; ++i) s.v[i] = value(T, l.lambda[i]);
        return s;
    }
};
inline Spectrum make_constant_spectrum(float c) { Spectrum s; s.kind = SPEC_CONSTANT; s.c = c; return s; }
// RgbAlbedoSpectrum::new (rgb_albedo_spectrum.rs:24-27)
inline Spectrum make_rgb_albedo(const Tables& T, Vec3 rgb, bool gamma_encoded) {
    Spectrum s; s.kind = SPEC_RGB_ALBEDO;
    if (!rgb_to_coeffs(T, rgb, gamma_encoded, s.coef)) { s.coef[0] = s.coef[1] = 0; s.coef[2] = INFINITY; }
    return s;
}
// RgbIlluminantSpectrum::new (rgb_illuminant_spectrum.rs:27-40): scale = 2*max, colour / scale keeps its (gamma) type
inline Spectrum make_rgb_illuminant(const Tables& T, Vec3 rgb, bool gamma_encoded) {
    Spectrum s; s.kind = SPEC_RGB_ILLUMINANT;
    float mx = max_element(rgb);
    s.scale = 2.0f * mx;
    Vec3 scaled = rgb / s.scale;
    if (!rgb_to_coeffs(T, scaled, gamma_encoded, s.coef)) { s.coef[0] = s.coef[1] = 0; s.coef[2] = INFINITY; }
    return s;
}

}  // namespace orc
