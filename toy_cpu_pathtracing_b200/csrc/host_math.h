// Host-side math for scene preparation (product code; plain C++, no CUDA).
// The reference prepares scenes with glam 0.30.3 f32 arithmetic; topology parity of the BVH and of the per-instance matrices
// depends on reproducing its operation order, so nothing here may be FMA-contracted: this translation
// unit is compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace tcpt {

struct V2 { float x, y; };
struct V3 {
    float x, y, z;
    float get(int i) const { return i == 0 ? x : i == 1 ? y : z; }
};
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 scale(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline float dot3(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline V3 cross3(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
static inline float len3(V3 a) { return std::sqrt(dot3(a, a)); }
static inline V3 unit(V3 a) { return scale(a, 1.0f / len3(a)); }
static inline float sel_min(float a, float b) { return a < b ? a : b; }
static inline float sel_max(float a, float b) { return a > b ? a : b; }
static inline V3 min3(V3 a, V3 b) { return {sel_min(a.x, b.x), sel_min(a.y, b.y), sel_min(a.z, b.z)}; }
static inline V3 max3(V3 a, V3 b) { return {sel_max(a.x, b.x), sel_max(a.y, b.y), sel_max(a.z, b.z)}; }
static inline bool any_nan(V3 a) { return std::isnan(a.x) || std::isnan(a.y) || std::isnan(a.z); }

struct Box {
    V3 lo, hi;
    void grow(const Box& o) { lo = min3(lo, o.lo); hi = max3(hi, o.hi); }
    float centre(int axis) const { return (lo.get(axis) + hi.get(axis)) * 0.5f; }      // math/src/bounds.rs:59-62
    float half_area2() const { V3 d = sub(hi, lo); return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); }  // bounds.rs:66-69
};

// column-major 4x4, m[col*4+row] (glam Mat4 memory order)
struct M4 {
    float m[16];
    float at(int col, int row) const { return m[col * 4 + row]; }
    float& at(int col, int row) { return m[col * 4 + row]; }
};
static inline M4 m4_identity() { M4 r; std::memset(&r, 0, sizeof r); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
static inline M4 m4_translation(V3 t) { M4 r = m4_identity(); r.m[12] = t.x; r.m[13] = t.y; r.m[14] = t.z; return r; }
// a * b with glam's column accumulation order ((c0*x + c1*y) + c2*z) + c3*w
static inline M4 m4_mul(const M4& a, const M4& b) {
    M4 r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r.at(j, i) = ((a.at(0, i) * b.at(j, 0) + a.at(1, i) * b.at(j, 1)) + a.at(2, i) * b.at(j, 2)) + a.at(3, i) * b.at(j, 3);
    return r;
}
static inline V3 m4_point(const M4& a, V3 p) {
    return {((a.at(0, 0) * p.x + a.at(1, 0) * p.y) + a.at(2, 0) * p.z) + a.at(3, 0),
            ((a.at(0, 1) * p.x + a.at(1, 1) * p.y) + a.at(2, 1) * p.z) + a.at(3, 1),
            ((a.at(0, 2) * p.x + a.at(1, 2) * p.y) + a.at(2, 2) * p.z) + a.at(3, 2)};
}
// general inverse by the cofactor scheme glam uses (GLM-derived), accumulated in the same order
static inline M4 m4_inverse(const M4& s) {
    const float a00 = s.at(0, 0), a01 = s.at(0, 1), a02 = s.at(0, 2), a03 = s.at(0, 3);
    const float a10 = s.at(1, 0), a11 = s.at(1, 1), a12 = s.at(1, 2), a13 = s.at(1, 3);
    const float a20 = s.at(2, 0), a21 = s.at(2, 1), a22 = s.at(2, 2), a23 = s.at(2, 3);
    const float a30 = s.at(3, 0), a31 = s.at(3, 1), a32 = s.at(3, 2), a33 = s.at(3, 3);
    const float c00 = a22 * a33 - a32 * a23, c02 = a12 * a33 - a32 * a13, c03 = a12 * a23 - a22 * a13;
    const float c04 = a21 * a33 - a31 * a23, c06 = a11 * a33 - a31 * a13, c07 = a11 * a23 - a21 * a13;
    const float c08 = a21 * a32 - a31 * a22, c10 = a11 * a32 - a31 * a12, c11 = a11 * a22 - a21 * a12;
    const float c12 = a20 * a33 - a30 * a23, c14 = a10 * a33 - a30 * a13, c15 = a10 * a23 - a20 * a13;
    const float c16 = a20 * a32 - a30 * a22, c18 = a10 * a32 - a30 * a12, c19 = a10 * a22 - a20 * a12;
    const float c20 = a20 * a31 - a30 * a21, c22 = a10 * a31 - a30 * a11, c23 = a10 * a21 - a20 * a11;
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {a10, a00, a00, a00}, v1[4] = {a11, a01, a01, a01}, v2[4] = {a12, a02, a02, a02}, v3_[4] = {a13, a03, a03, a03};
    M4 inv;
    for (int i = 0; i < 4; ++i) {
        const float sgn_a = (i & 1) ? -1.0f : 1.0f, sgn_b = -sgn_a;
        inv.at(0, i) = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3_[i] * f2[i]) * sgn_a;
        inv.at(1, i) = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3_[i] * f4[i]) * sgn_b;
        inv.at(2, i) = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3_[i] * f5[i]) * sgn_a;
        inv.at(3, i) = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sgn_b;
    }
    const float det = ((s.at(0, 0) * inv.at(0, 0) + s.at(0, 1) * inv.at(1, 0)) + s.at(0, 2) * inv.at(2, 0)) + s.at(0, 3) * inv.at(3, 0);
    const float rcp = 1.0f / det;
    for (float& x : inv.m) x = x * rcp;
    return inv;
}
static inline bool m4_is_identity(const M4& a) { M4 i = m4_identity(); return std::memcmp(a.m, i.m, sizeof a.m) == 0; }

// AABB of the eight transformed corners of a local box (math/src/transform.rs:61-74)
static inline Box transform_box(const M4& m, const Box& b) {
    const float inf = INFINITY;
    Box r{{inf, inf, inf}, {-inf, -inf, -inf}};
    for (int k = 0; k < 8; ++k) {
        V3 c = {(k & 1) ? b.hi.x : b.lo.x, (k & 2) ? b.hi.y : b.lo.y, (k & 4) ? b.hi.z : b.lo.z};
        V3 q = m4_point(m, c);
        r.lo = min3(r.lo, q);
        r.hi = max3(r.hi, q);
    }
    return r;
}

}  // namespace tcpt
