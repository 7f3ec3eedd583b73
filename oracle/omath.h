// ORACLE (test infrastructure, not product code).  CPU restatement of the reference's math layer.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything in oracle/.
//
// Follows /root/reference/math/src/{vector,point,normal,bounds,ray,transform}.rs and the glam 0.30.3 scalar semantics
// they rely on (glam is a Cargo dependency absent from the checkout: Vec3 is a plain 3-float struct on x86-64,
// dot = x*x' + y*y' + z*z' left to right, normalize = v * (1/sqrt(dot)), Mat4 is column major).
// PARITY UNPINNED: the reference cannot be compiled here (no Rust toolchain) and has no unit tests below image level,
// so these functions are pinned only by line-by-line citation.  Compile with -ffp-contract=off (Rust never fuses).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

struct Vec2 {
    float x = 0, y = 0;
};

struct Vec3 {
    float x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// glam Vec3::dot: (x*x') + (y*y') + (z*z')
inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
// glam Vec3::cross
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length_squared(Vec3 a) { return dot(a, a); }
inline float length(Vec3 a) { return std::sqrt(dot(a, a)); }
// glam Vec3::normalize = self * length_recip()
inline Vec3 normalize(Vec3 a) { return a * (1.0f / length(a)); }
inline Vec3 vabs(Vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
// Rust f32::min/max ignore NaN in favour of the other operand; glam's component-wise min/max: a<b?a:b / a>b?a:b
inline float fmin_(float a, float b) { return a < b ? a : b; }
inline float fmax_(float a, float b) { return a > b ? a : b; }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }
// Rust f32::max / f32::min (std): NaN-ignoring
inline float rmax(float a, float b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a > b ? a : b)); }
inline float rmin(float a, float b) { return std::isnan(a) ? b : (std::isnan(b) ? a : (a < b ? a : b)); }
inline float max_element(Vec3 a) { return rmax(a.x, rmax(a.y, a.z)); }
// glam max_position: index of the first maximum
inline int max_position(Vec3 a) {
    float m = a.x;
    int idx = 0;
    if (a.y > m) { m = a.y; idx = 1; }
    if (a.z > m) { idx = 2; }
    return idx;
}
inline bool is_nan(Vec3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }
// Rust f32::signum: +1 for +0/positive/inf, -1 for -0/negative, NaN for NaN
inline float signum(float x) { return std::isnan(x) ? x : (std::signbit(x) ? -1.0f : 1.0f); }
// Rust f32::clamp (min <= max assumed)
inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
// Rust `as usize`/`as u32` from f32: truncates toward zero, saturates, NaN -> 0
inline uint32_t f2u_sat(float x) {
    if (!(x > 0.0f)) return 0;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)x;
}

// Normal::from = normalise on construction (math/src/normal.rs:95-101)
inline Vec3 make_normal(Vec3 v) { return normalize(v); }

struct Vec4 {
    float x = 0, y = 0, z = 0, w = 0;
};

// glam Mat4: column-major, columns c[0..3]
struct Mat4 {
    float c[4][4];  // c[col][row]
    static Mat4 identity() {
        Mat4 m;
        std::memset(&m, 0, sizeof m);
        m.c[0][0] = m.c[1][1] = m.c[2][2] = m.c[3][3] = 1.0f;
        return m;
    }
    static Mat4 from_translation(Vec3 t) {
        Mat4 m = identity();
        m.c[3][0] = t.x; m.c[3][1] = t.y; m.c[3][2] = t.z;
        return m;
    }
    static Mat4 from_cols3(Vec3 a, Vec3 b, Vec3 cc) {
        Mat4 m = identity();
        m.c[0][0] = a.x; m.c[0][1] = a.y; m.c[0][2] = a.z;
        m.c[1][0] = b.x; m.c[1][1] = b.y; m.c[1][2] = b.z;
        m.c[2][0] = cc.x; m.c[2][1] = cc.y; m.c[2][2] = cc.z;
        return m;
    }
};
// glam Mat4::mul_vec4: ((x_axis*v.x + y_axis*v.y) + z_axis*v.z) + w_axis*v.w
inline Mat4 mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r.c[j][i] = ((a.c[0][i] * b.c[j][0] + a.c[1][i] * b.c[j][1]) + a.c[2][i] * b.c[j][2]) + a.c[3][i] * b.c[j][3];
    return r;
}
// glam Mat4::transform_point3: ((x_axis*p.x + y_axis*p.y) + z_axis*p.z) + w_axis
inline Vec3 transform_point3(const Mat4& m, Vec3 p) {
    return {((m.c[0][0] * p.x + m.c[1][0] * p.y) + m.c[2][0] * p.z) + m.c[3][0],
            ((m.c[0][1] * p.x + m.c[1][1] * p.y) + m.c[2][1] * p.z) + m.c[3][1],
            ((m.c[0][2] * p.x + m.c[1][2] * p.y) + m.c[2][2] * p.z) + m.c[3][2]};
}
inline Vec3 transform_vector3(const Mat4& m, Vec3 v) {
    return {(m.c[0][0] * v.x + m.c[1][0] * v.y) + m.c[2][0] * v.z,
            (m.c[0][1] * v.x + m.c[1][1] * v.y) + m.c[2][1] * v.z,
            (m.c[0][2] * v.x + m.c[1][2] * v.y) + m.c[2][2] * v.z};
}
inline Mat4 transpose(const Mat4& m) {
    Mat4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) r.c[j][i] = m.c[i][j];
    return r;
}
// glam 0.30.3 Mat4::inverse, scalar code path (the SSE2 path is the same cofactor scheme vectorised)
inline Mat4 inverse(const Mat4& s) {
    float m00 = s.c[0][0], m01 = s.c[0][1], m02 = s.c[0][2], m03 = s.c[0][3];
    float m10 = s.c[1][0], m11 = s.c[1][1], m12 = s.c[1][2], m13 = s.c[1][3];
    float m20 = s.c[2][0], m21 = s.c[2][1], m22 = s.c[2][2], m23 = s.c[2][3];
    float m30 = s.c[3][0], m31 = s.c[3][1], m32 = s.c[3][2], m33 = s.c[3][3];
    float coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13, coef03 = m12 * m23 - m22 * m13;
    float coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13, coef07 = m11 * m23 - m21 * m13;
    float coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12, coef11 = m11 * m22 - m21 * m12;
    float coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13, coef15 = m10 * m23 - m20 * m13;
    float coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12, coef19 = m10 * m22 - m20 * m12;
    float coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11, coef23 = m10 * m21 - m20 * m11;
    float fac0[4] = {coef00, coef00, coef02, coef03}, fac1[4] = {coef04, coef04, coef06, coef07};
    float fac2[4] = {coef08, coef08, coef10, coef11}, fac3[4] = {coef12, coef12, coef14, coef15};
    float fac4[4] = {coef16, coef16, coef18, coef19}, fac5[4] = {coef20, coef20, coef22, coef23};
    float vec0[4] = {m10, m00, m00, m00}, vec1[4] = {m11, m01, m01, m01};
    float vec2[4] = {m12, m02, m02, m02}, vec3[4] = {m13, m03, m03, m03};
    const float sa[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sb[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
    Mat4 inv;
    for (int i = 0; i < 4; ++i) {
        float inv0 = (vec1[i] * fac0[i] - vec2[i] * fac1[i]) + vec3[i] * fac2[i];
        float inv1 = (vec0[i] * fac0[i] - vec2[i] * fac3[i]) + vec3[i] * fac4[i];
        float inv2 = (vec0[i] * fac1[i] - vec1[i] * fac3[i]) + vec3[i] * fac5[i];
        float inv3 = (vec0[i] * fac2[i] - vec1[i] * fac4[i]) + vec2[i] * fac5[i];
        inv.c[0][i] = inv0 * sa[i];
        inv.c[1][i] = inv1 * sb[i];
        inv.c[2][i] = inv2 * sa[i];
        inv.c[3][i] = inv3 * sb[i];
    }
    float d0 = s.c[0][0] * inv.c[0][0], d1 = s.c[0][1] * inv.c[1][0], d2 = s.c[0][2] * inv.c[2][0], d3 = s.c[0][3] * inv.c[3][0];
    float det = ((d0 + d1) + d2) + d3;
    float rcp = 1.0f / det;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) inv.c[j][i] = inv.c[j][i] * rcp;
    return inv;
}
// Transform * Normal: inverse().transpose() per call, result re-normalised (math/src/transform.rs:44-51)
inline Vec3 transform_normal(const Mat4& m, Vec3 n) { return make_normal(transform_vector3(transpose(inverse(m)), n)); }

struct Ray {
    Vec3 o, d;
    // Ray::move_forward (math/src/ray.rs:23-26)
    Ray move_forward(float dist) const { return Ray{o + d * dist, d}; }
};
inline Ray transform_ray(const Mat4& m, const Ray& r) { return Ray{transform_point3(m, r.o), transform_vector3(m, r.d)}; }

struct Bounds {
    Vec3 mn, mx;
    // Bounds::merge (math/src/bounds.rs:81-86)
    Bounds merge(const Bounds& o) const { return Bounds{vmin(mn, o.mn), vmax(mx, o.mx)}; }
    // Bounds::center (bounds.rs:59-62)
    Vec3 center() const { return (mn + mx) * 0.5f; }
    // Bounds::area (bounds.rs:66-69)
    float area() const {
        Vec3 d = mx - mn;
        return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z);
    }
    // Bounds::intersect slab test (bounds.rs:27-55); returns false on miss
    bool intersect(const Ray& ray, float t_max, Vec3 inv_dir) const {
        float t0 = 0.0f, t1 = t_max;
        for (int i = 0; i < 3; ++i) {
            float t_near = (mn[i] - ray.o[i]) * inv_dir[i];
            float t_far = (mx[i] - ray.o[i]) * inv_dir[i];
            if (t_near > t_far) { float t = t_near; t_near = t_far; t_far = t; }
            t0 = t_near > t0 ? t_near : t0;
            t1 = t_far < t1 ? t_far : t1;
            if (t0 > t1) return false;
        }
        return true;
    }
};
// Transform * Bounds: AABB of the 8 transformed corners (math/src/transform.rs:61-74, bounds.rs:90-104)
inline Bounds transform_bounds(const Mat4& m, const Bounds& b) {
    const float inf = std::numeric_limits<float>::infinity();
    Vec3 mn(inf, inf, inf), mx(-inf, -inf, -inf);
    for (int k = 0; k < 8; ++k) {
        Vec3 p((k & 1) ? b.mx.x : b.mn.x, (k & 2) ? b.mx.y : b.mn.y, (k & 4) ? b.mx.z : b.mn.z);
        Vec3 q = transform_point3(m, p);
        mn = vmin(mn, q);
        mx = vmax(mx, q);
    }
    return Bounds{mn, mx};
}

struct TriangleHit {
    float t_hit;
    Vec3 position;
    Vec3 normal;
    float bary[3];
};

struct TraversalCounters {
    uint64_t box_tests = 0, tri_tests = 0;
};

// math::intersect_triangle (math/src/ray.rs:44-182): watertight test with f64 fallback and conservative t > delta_t.
inline bool intersect_triangle(const Ray& ray, float t_max, const Vec3 ps[3], TriangleHit* out) {
    if (length_squared(cross(ps[1] - ps[0], ps[2] - ps[0])) == 0.0f) return false;
    Vec3 p0o = ps[0] - ray.o, p1o = ps[1] - ray.o, p2o = ps[2] - ray.o;
    Vec3 d = ray.d;
    int kz = max_position(vabs(d));
    int kx = (kz + 1) % 3;
    int ky = (kx + 1) % 3;
    d = Vec3(d[kx], d[ky], d[kz]);
    Vec3 p0(p0o[kx], p0o[ky], p0o[kz]), p1(p1o[kx], p1o[ky], p1o[kz]), p2(p2o[kx], p2o[ky], p2o[kz]);
    float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
    p0.x += sx * p0.z; p0.y += sy * p0.z;
    p1.x += sx * p1.z; p1.y += sy * p1.z;
    p2.x += sx * p2.z; p2.y += sy * p2.z;
    float e0 = p2.x * p1.y - p2.y * p1.x;
    float e1 = p0.x * p2.y - p0.y * p2.x;
    float e2 = p1.x * p0.y - p1.y * p0.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        e0 = (float)((double)p2.x * (double)p1.y - (double)p2.y * (double)p1.x);
        e1 = (float)((double)p0.x * (double)p2.y - (double)p0.y * (double)p2.x);
        e2 = (float)((double)p1.x * (double)p0.y - (double)p1.y * (double)p0.x);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0.z *= sz; p1.z *= sz; p2.z *= sz;
    float t_scaled = e0 * p0.z + e1 * p1.z + e2 * p2.z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return false;
    float inv_det = 1.0f / det;
    float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    float t_hit = t_scaled * inv_det;
    // gamma(n) = n*eps/(1-n*eps), eps = 2^-24 (const fn in f32, ray.rs:137-140)
    const float EPS = 5.9604644775390625e-8f;
    auto gamma = [&](int n) { return ((float)n * EPS) / (1.0f - (float)n * EPS); };
    float max_zt = max_element(vabs(Vec3(p0.z, p1.z, p2.z)));
    float delta_z = gamma(3) * max_zt;
    float max_xt = max_element(vabs(Vec3(p0.x, p1.x, p2.x)));
    float max_yt = max_element(vabs(Vec3(p0.y, p1.y, p2.y)));
    float delta_x = gamma(5) * max_xt;
    float delta_y = gamma(5) * max_yt;
    float delta_e = 2.0f * (gamma(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    float max_e = max_element(vabs(Vec3(e0, e1, e2)));
    float delta_t = 3.0f * (gamma(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * std::fabs(inv_det);
    if (t_hit < delta_t) return false;
    if (out) {
        out->t_hit = t_hit;
        out->bary[0] = b0; out->bary[1] = b1; out->bary[2] = b2;
        out->position = ps[0] * b0 + ps[1] * b1 + ps[2] * b2;
        out->normal = make_normal(normalize(cross(ps[1] - ps[0], ps[2] - ps[0])));
    }
    return true;
}

}  // namespace orc
