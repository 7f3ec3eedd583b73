// Two-level BVH traversal on the device (software; B200 has no RT cores).
//
// Reproduces the RESULT of the reference's exhaustive recursive traversal (/root/reference/scene/src/bvh.rs:344-520) with an
// ordered, t-shrinking, stack-based walk:
//   * slab test = math/src/bounds.rs:27-55, same operations in the same order (sub, mul, compare-selects; no FMA);
//   * triangle test = math/src/ray.rs:44-158 (watertight shear, f64 fallback when an edge function is 0, conservative t > delta_t);
//   * instance transform = primitive/impls/triangle_mesh.rs:97 (ray parameter t preserved, direction not re-normalised);
//   * the reference never shrinks t_max and keeps candidates by  Node: ties -> second child,  Leaf: ties -> earlier item
//     (bvh.rs:384-388, 413-420).  That is the total order  (t, later leaf first, earlier item first)  applied per level
//     (TLAS, then BLAS), so any visiting order that sees every candidate the winner competes with gives the same hit.
//     Boxes are culled against  t_best * (1 + 2^-10) + 2^-10  instead of t_best: the slab interval of a box and the
//     watertight t of a triangle inside it are rounded independently, so the margin keeps every box that could still hold an
//     equal-or-smaller t (the triangle test itself always runs with the caller's t_max, like the reference).
#pragma once
#include "dcommon.cuh"

namespace tcpt {

struct DHit {
    float t, b0, b1, b2;
    int prim;      // primitive index, -1 = miss
    uint32_t tri;  // triangle index within the geometry
};

struct RayXform {  // per-space ray constants for the watertight test (math/src/ray.rs:63-78)
    float3 o, d, inv_d;
    int kx, ky, kz;
    float sx, sy, sz;
};

__device__ __forceinline__ void ray_setup(RayXform& r, float3 o, float3 d) {
    r.o = o; r.d = d;
    r.inv_d = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // bvh.rs:433
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { m = ay; kz = 1; }
    if (az > m) { kz = 2; }
    r.kz = kz; r.kx = (kz + 1) % 3; r.ky = (r.kx + 1) % 3;
    const float dx = comp(d, r.kx), dy = comp(d, r.ky), dz = comp(d, r.kz);
    r.sx = -dx / dz; r.sy = -dy / dz; r.sz = 1.0f / dz;
}

// Bounds::intersect; NaNs (0 * inf) are ignored by the ordered compare-selects exactly as in the reference
__device__ __forceinline__ bool slab_test(const float4 lo, const float4 hi, const RayXform& r, float t_max, float* t_entry) {
    float t0 = 0.0f, t1 = t_max;
    {
        float tn = (lo.x - r.o.x) * r.inv_d.x, tf = (hi.x - r.o.x) * r.inv_d.x;
        if (tn > tf) { const float s = tn; tn = tf; tf = s; }
        t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
    }
    {
        float tn = (lo.y - r.o.y) * r.inv_d.y, tf = (hi.y - r.o.y) * r.inv_d.y;
        if (tn > tf) { const float s = tn; tn = tf; tf = s; }
        t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
    }
    {
        float tn = (lo.z - r.o.z) * r.inv_d.z, tf = (hi.z - r.o.z) * r.inv_d.z;
        if (tn > tf) { const float s = tn; tn = tf; tf = s; }
        t0 = tn > t0 ? tn : t0; t1 = tf < t1 ? tf : t1;
    }
    *t_entry = t0;
    return !(t0 > t1);
}

// math::intersect_triangle up to the accept decision; returns true and t/barycentrics on a hit
__device__ __forceinline__ bool tri_test(const float4 v0, const float4 v1, const float4 v2, const RayXform& r, float t_max, float* t_out, float* b0, float* b1, float* b2) {
    if (__float_as_uint(v1.w) != 0u) return false;  // degenerate (|e1 x e2|^2 == 0), decided on the host
    const float3 a = f3(v0.x - r.o.x, v0.y - r.o.y, v0.z - r.o.z), b = f3(v1.x - r.o.x, v1.y - r.o.y, v1.z - r.o.z), c = f3(v2.x - r.o.x, v2.y - r.o.y, v2.z - r.o.z);
    float p0x = comp(a, r.kx), p0y = comp(a, r.ky), p0z = comp(a, r.kz);
    float p1x = comp(b, r.kx), p1y = comp(b, r.ky), p1z = comp(b, r.kz);
    float p2x = comp(c, r.kx), p2y = comp(c, r.ky), p2z = comp(c, r.kz);
    p0x += r.sx * p0z; p0y += r.sy * p0z;
    p1x += r.sx * p1z; p1y += r.sy * p1z;
    p2x += r.sx * p2z; p2y += r.sy * p2z;
    float e0 = p2x * p1y - p2y * p1x;
    float e1 = p0x * p2y - p0y * p2x;
    float e2 = p1x * p0y - p1y * p0x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        e0 = (float)__dsub_rn(__dmul_rn((double)p2x, (double)p1y), __dmul_rn((double)p2y, (double)p1x));
        e1 = (float)__dsub_rn(__dmul_rn((double)p0x, (double)p2y), __dmul_rn((double)p0y, (double)p2x));
        e2 = (float)__dsub_rn(__dmul_rn((double)p1x, (double)p0y), __dmul_rn((double)p1y, (double)p0x));
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    const float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0z *= r.sz; p1z *= r.sz; p2z *= r.sz;
    const float t_scaled = e0 * p0z + e1 * p1z + e2 * p2z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return false;
    const float inv_det = 1.0f / det;
    const float t_hit = t_scaled * inv_det;
    // conservative t > delta_t (ray.rs:137-158); gamma(n) = n*eps/(1-n*eps) evaluated in f32 like the reference's const fn
    const float EPS = 5.9604644775390625e-8f;
    const float g2 = (2.0f * EPS) / (1.0f - 2.0f * EPS), g3 = (3.0f * EPS) / (1.0f - 3.0f * EPS), g5 = (5.0f * EPS) / (1.0f - 5.0f * EPS);
    const float max_zt = rmax(fabsf(p0z), rmax(fabsf(p1z), fabsf(p2z)));
    const float delta_z = g3 * max_zt;
    const float max_xt = rmax(fabsf(p0x), rmax(fabsf(p1x), fabsf(p2x)));
    const float max_yt = rmax(fabsf(p0y), rmax(fabsf(p1y), fabsf(p2y)));
    const float delta_x = g5 * max_xt, delta_y = g5 * max_yt;
    const float delta_e = 2.0f * (g2 * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    const float max_e = rmax(fabsf(e0), rmax(fabsf(e1), fabsf(e2)));
    const float delta_t = 3.0f * (g3 * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
    if (t_hit < delta_t) return false;
    *t_out = t_hit; *b0 = e0 * inv_det; *b1 = e1 * inv_det; *b2 = e2 * inv_det;
    return true;
}

__device__ __forceinline__ float cull_limit(float t_best) { return t_best * 1.0009765625f + 0.0009765625f; }

#define TCPT_STACK_SENTINEL 0xffffffffu

// Closest hit.  COUNT adds box/triangle test counters (algorithmic work B, T of SURVEY.md section 8d).
template <bool COUNT>
__device__ inline DHit trace_closest(const DScene& sc, float3 o, float3 d, float t_max, uint32_t* n_box, uint32_t* n_tri) {
    DHit best; best.prim = -1; best.t = t_max; best.b0 = best.b1 = best.b2 = 0.0f; best.tri = 0;
    // tie-break state of the current best: TLAS leaf, position of the primitive inside that leaf, BLAS leaf
    uint32_t best_tleaf = 0, best_titem = 0, best_bleaf = 0;
    float limit = t_max;  // box-culling bound (t_max until the first hit)

    RayXform rw; ray_setup(rw, o, d);
    RayXform rl = rw;
    uint32_t stack[TCPT_TRAVERSAL_STACK];
    int sp = 0;
    // TLAS root
    {
        const float4 lo = __ldg(&sc.nodes[0]), hi = __ldg(&sc.nodes[1]);
        float te; if (COUNT) (*n_box)++;
        if (!slab_test(lo, hi, rw, limit, &te)) return best;
    }
    uint32_t node = 0;           // current node (absolute index into sc.nodes/2)
    bool in_blas = false;
    uint32_t blas_base = 0, slot_base = 0;
    int cur_prim = -1; uint32_t cur_tleaf = 0, cur_titem = 0;
    // pending TLAS leaf iteration state
    uint32_t tl_first = 0, tl_count = 0, tl_next = 0;

    for (;;) {
        const float4 nlo = __ldg(&sc.nodes[2 * (size_t)node]), nhi = __ldg(&sc.nodes[2 * (size_t)node + 1]);
        const uint32_t a = __float_as_uint(nlo.w), cnt = __float_as_uint(nhi.w);
        bool pop = false;
        if (cnt == 0) {
            // inner: test both children, descend into the nearer, push the farther
            const uint32_t base = in_blas ? blas_base : 0u;
            const uint32_t c0 = node + 1, c1 = base + a;
            const float4 lo0 = __ldg(&sc.nodes[2 * (size_t)c0]), hi0 = __ldg(&sc.nodes[2 * (size_t)c0 + 1]);
            const float4 lo1 = __ldg(&sc.nodes[2 * (size_t)c1]), hi1 = __ldg(&sc.nodes[2 * (size_t)c1 + 1]);
            float t0e, t1e;
            const RayXform& r = in_blas ? rl : rw;
            const bool h0 = slab_test(lo0, hi0, r, limit, &t0e);
            const bool h1 = slab_test(lo1, hi1, r, limit, &t1e);
            if (COUNT) (*n_box) += 2;
            if (h0 && h1) {
                if (t1e < t0e) { stack[sp++] = c0; node = c1; } else { stack[sp++] = c1; node = c0; }
            } else if (h0) node = c0;
            else if (h1) node = c1;
            else pop = true;
        } else if (!in_blas) {
            // TLAS leaf: iterate its primitives one at a time (each may open a BLAS)
            tl_first = a; tl_count = cnt; tl_next = 0; cur_tleaf = node;
            pop = true;  // falls into the TLAS-leaf continuation below
            stack[sp++] = TCPT_STACK_SENTINEL;  // marker: resume TLAS leaf iteration
        } else {
            // BLAS leaf: test triangles in order
            for (uint32_t i = 0; i < cnt; ++i) {
                const size_t s = 3 * (size_t)(slot_base + a + i);
                const float4 v0 = __ldg(&sc.tri_verts[s]), v1 = __ldg(&sc.tri_verts[s + 1]), v2 = __ldg(&sc.tri_verts[s + 2]);
                float t, b0, b1, b2;
                if (COUNT) (*n_tri)++;
                if (tri_test(v0, v1, v2, rl, t_max, &t, &b0, &b1, &b2)) {
                    // total order: smaller t; then (TLAS) later leaf, earlier primitive slot; then (BLAS) later leaf, earlier item
                    bool take;
                    if (best.prim < 0) take = true;
                    else if (t != best.t) take = t < best.t;
                    else if (cur_tleaf != best_tleaf) take = cur_tleaf > best_tleaf;
                    else if (cur_titem != best_titem) take = cur_titem < best_titem;
                    else if (node != best_bleaf) take = node > best_bleaf;
                    else take = false;  // same leaf: earlier item already recorded
                    if (take) {
                        best.t = t; best.b0 = b0; best.b1 = b1; best.b2 = b2; best.prim = cur_prim; best.tri = __float_as_uint(v0.w);
                        best_tleaf = cur_tleaf; best_titem = cur_titem; best_bleaf = node;
                        limit = fminf(t_max, cull_limit(t));
                    }
                }
            }
            pop = true;
        }
        while (pop) {
            if (sp == 0) return best;
            const uint32_t top = stack[--sp];
            if (top == TCPT_STACK_SENTINEL) {
                // continue the TLAS leaf: open the next primitive's BLAS, or finish the leaf
                in_blas = false;
                if (tl_next < tl_count) {
                    cur_titem = tl_next;
                    cur_prim = __ldg(&sc.tlas_items[tl_first + tl_next]);
                    tl_next++;
                    const tcpt_flat_primitive& P = sc.primitives[cur_prim];
                    const tcpt_flat_geometry& G = sc.geometries[P.geometry];
                    if (P.identity) rl = rw;
                    else ray_setup(rl, xf_point(P.r2l, o), xf_vector(P.r2l, d));
                    blas_base = G.node_base; slot_base = G.slot_base;
                    stack[sp++] = TCPT_STACK_SENTINEL;
                    // BLAS root box with the caller's semantics
                    const float4 lo = __ldg(&sc.nodes[2 * (size_t)blas_base]), hi = __ldg(&sc.nodes[2 * (size_t)blas_base + 1]);
                    float te; if (COUNT) (*n_box)++;
                    if (slab_test(lo, hi, rl, limit, &te)) { in_blas = true; node = blas_base; pop = false; }
                } else {
                    // the sentinel pushed for this leaf is consumed; restore the enclosing TLAS leaf state is not needed:
                    // TLAS leaves never nest
                }
            } else {
                // the culling bound may have shrunk since this node was pushed: re-test lazily by just visiting it
                node = top; pop = false;
            }
        }
    }
}

// Any hit (Scene::intersect_p, scene.rs:93-103): order independent, returns at the first accepted triangle
template <bool COUNT>
__device__ inline bool trace_any(const DScene& sc, float3 o, float3 d, float t_max, uint32_t* n_box, uint32_t* n_tri) {
    RayXform rw; ray_setup(rw, o, d);
    RayXform rl = rw;
    uint32_t stack[TCPT_TRAVERSAL_STACK];
    int sp = 0;
    {
        const float4 lo = __ldg(&sc.nodes[0]), hi = __ldg(&sc.nodes[1]);
        float te; if (COUNT) (*n_box)++;
        if (!slab_test(lo, hi, rw, t_max, &te)) return false;
    }
    uint32_t node = 0;
    bool in_blas = false;
    uint32_t blas_base = 0, slot_base = 0;
    uint32_t tl_first = 0, tl_count = 0, tl_next = 0;
    for (;;) {
        const float4 nlo = __ldg(&sc.nodes[2 * (size_t)node]), nhi = __ldg(&sc.nodes[2 * (size_t)node + 1]);
        const uint32_t a = __float_as_uint(nlo.w), cnt = __float_as_uint(nhi.w);
        bool pop = false;
        if (cnt == 0) {
            const uint32_t base = in_blas ? blas_base : 0u;
            const uint32_t c0 = node + 1, c1 = base + a;
            const float4 lo0 = __ldg(&sc.nodes[2 * (size_t)c0]), hi0 = __ldg(&sc.nodes[2 * (size_t)c0 + 1]);
            const float4 lo1 = __ldg(&sc.nodes[2 * (size_t)c1]), hi1 = __ldg(&sc.nodes[2 * (size_t)c1 + 1]);
            float t0e, t1e;
            const RayXform& r = in_blas ? rl : rw;
            const bool h0 = slab_test(lo0, hi0, r, t_max, &t0e);
            const bool h1 = slab_test(lo1, hi1, r, t_max, &t1e);
            if (COUNT) (*n_box) += 2;
            if (h0 && h1) { stack[sp++] = c1; node = c0; }
            else if (h0) node = c0;
            else if (h1) node = c1;
            else pop = true;
        } else if (!in_blas) {
            tl_first = a; tl_count = cnt; tl_next = 0;
            pop = true;
            stack[sp++] = TCPT_STACK_SENTINEL;
        } else {
            for (uint32_t i = 0; i < cnt; ++i) {
                const size_t s = 3 * (size_t)(slot_base + a + i);
                const float4 v0 = __ldg(&sc.tri_verts[s]), v1 = __ldg(&sc.tri_verts[s + 1]), v2 = __ldg(&sc.tri_verts[s + 2]);
                float t, b0, b1, b2;
                if (COUNT) (*n_tri)++;
                if (tri_test(v0, v1, v2, rl, t_max, &t, &b0, &b1, &b2)) return true;
            }
            pop = true;
        }
        while (pop) {
            if (sp == 0) return false;
            const uint32_t top = stack[--sp];
            if (top == TCPT_STACK_SENTINEL) {
                in_blas = false;
                if (tl_next < tl_count) {
                    const int prim = __ldg(&sc.tlas_items[tl_first + tl_next]);
                    tl_next++;
                    const tcpt_flat_primitive& P = sc.primitives[prim];
                    const tcpt_flat_geometry& G = sc.geometries[P.geometry];
                    if (P.identity) rl = rw;
                    else ray_setup(rl, xf_point(P.r2l, o), xf_vector(P.r2l, d));
                    blas_base = G.node_base; slot_base = G.slot_base;
                    stack[sp++] = TCPT_STACK_SENTINEL;
                    const float4 lo = __ldg(&sc.nodes[2 * (size_t)blas_base]), hi = __ldg(&sc.nodes[2 * (size_t)blas_base + 1]);
                    float te; if (COUNT) (*n_box)++;
                    if (slab_test(lo, hi, rl, t_max, &te)) { in_blas = true; node = blas_base; pop = false; }
                }
            } else { node = top; pop = false; }
        }
    }
}

}  // namespace tcpt
