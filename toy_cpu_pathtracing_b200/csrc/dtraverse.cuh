// Two-level BVH traversal on the device (software; B200 has no RT cores).
//
// Reproduces the RESULT of the reference's exhaustive recursive traversal (/root/reference/scene/src/bvh.rs:344-520) with an
// ordered, t-shrinking, stack-based walk over 64-byte child-pair records (include/tcpt_flat.h):
//   * slab test = math/src/bounds.rs:27-55, same operations in the same order (sub, mul, compare-selects; no FMA);
//   * triangle test = math/src/ray.rs:44-158 (watertight shear, f64 fallback when an edge function is 0, conservative t > delta_t);
//   * instance transform = primitive/impls/triangle_mesh.rs:97 (ray parameter t preserved, direction not re-normalised);
//   * the reference never shrinks t_max and keeps candidates by  Node: ties -> second child,  Leaf: ties -> earlier item
//     (bvh.rs:384-388, 413-420).  That is the total order  (t, later leaf first, earlier item first)  applied per level
//     (TLAS, then BLAS), so any visiting order that sees every candidate the winner competes with gives the same hit.
//     Every box on the path to a candidate is tested with the reference's own slab arithmetic, so the candidate set is a
//     subset of the reference's; boxes are additionally culled against  t_best * (1 + 2^-10) + 2^-10  instead of t_best:
//     the slab interval of a box and the watertight t of a triangle inside it are rounded independently, so the margin keeps
//     every box that could still hold an equal-or-smaller t (the triangle test itself always runs with the caller's t_max).
#pragma once
#include "dcommon.cuh"

// Three micro-variants of the walk, measured one by one on the 4K frame and left off (trace 46.1 ms per step without them):
//   TCPT_OPT_TOS    top stack entry in a register, so a pop does not wait for local memory           46.7 ms
//   TCPT_OPT_RL     one "current ray" record restored at BLAS exit instead of a per-visit select     46.6 ms
//   TCPT_OPT_BALLOT idle lanes derived from the next iteration's two masks (one ballot less)         46.5 ms
// Each removes instructions from the loop but moves the 72-register allocation to a worse place.
#ifndef TCPT_OPT_TOS
#define TCPT_OPT_TOS 0
#endif
#ifndef TCPT_OPT_RL
#define TCPT_OPT_RL 0
#endif
#ifndef TCPT_OPT_BALLOT
#define TCPT_OPT_BALLOT 0
#endif
#ifndef TCPT_SPECULATE
#define TCPT_SPECULATE 0          // 1: walk one leaf ahead of the triangle tests (Aila & Laine postponed leaf); measured 6 % slower (3.52 vs 3.32 ms/spp)
#endif

namespace tcpt {

struct DHit {
    float t, b0, b1, b2;
    int prim;      // primitive index, -1 = miss
    uint32_t tri;  // triangle index within the geometry
};

struct RayXform {  // per-space ray constants for the slab and watertight tests (math/src/ray.rs:63-78)
    float3 o, inv_d;
    int kz;
    float sx, sy, sz;
};

__device__ __forceinline__ void ray_setup(RayXform& r, float3 o, float3 d) {
    r.o = o;
    r.inv_d = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // bvh.rs:433
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { m = ay; kz = 1; }
    if (az > m) { kz = 2; }
    r.kz = kz;
    const float dx = kz == 0 ? d.y : (kz == 1 ? d.z : d.x), dy = kz == 0 ? d.z : (kz == 1 ? d.x : d.y), dz = kz == 0 ? d.x : (kz == 1 ? d.y : d.z);
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
    const float ax = v0.x - r.o.x, ay = v0.y - r.o.y, az = v0.z - r.o.z;
    const float bx = v1.x - r.o.x, by = v1.y - r.o.y, bz = v1.z - r.o.z;
    const float cx = v2.x - r.o.x, cy = v2.y - r.o.y, cz = v2.z - r.o.z;
    float p0x, p0y, p0z, p1x, p1y, p1z, p2x, p2y, p2z;
    if (r.kz == 0) { p0x = ay; p0y = az; p0z = ax; p1x = by; p1y = bz; p1z = bx; p2x = cy; p2y = cz; p2z = cx; }
    else if (r.kz == 1) { p0x = az; p0y = ax; p0z = ay; p1x = bz; p1y = bx; p1z = by; p2x = cz; p2y = cx; p2z = cy; }
    else { p0x = ax; p0y = ay; p0z = az; p1x = bx; p1y = by; p1z = bz; p2x = cx; p2y = cy; p2z = cz; }
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

#define TCPT_TLAS_ITEM_BIT 0x80000000u
#define TCPT_ABSENT 0xffffffffu

// One ray in flight.  The walk is an explicit state machine so that a warp can keep its lanes busy:
//   * a lane whose ray is finished picks up the next ray of the queue instead of idling until the slowest ray of the warp is
//     done (persistent threads with dynamic fetch; first profile, one ray per thread: 4.2 of 32 lanes active on bounce rays);
//   * reaching a leaf only RECORDS its triangle slots; the warp switches to a triangle phase (one triangle per lane per
//     iteration) once enough lanes hold pending triangles (second profile: the inline leaf loops ran with 1.9 lanes active and
//     were 60 % of the kernel's warp instructions).
struct Traversal {
    float3 o, d;                 // the ray in Render space
    float t_max, limit;          // caller's t_max; box-culling bound (t_max until the first hit)
    RayXform rw, rl;             // Render-space and current BLAS-space ray constants
    DHit best;
    uint32_t best_tleaf, best_tslot, best_bleaf, best_bslot;  // tie-break keys of `best`: TLAS (leaf first slot, slot), BLAS (same)
    int sp, blas_sp;             // stack height; stack height at BLAS entry (-1 = traversing the TLAS)
    uint32_t tos;                // the top stack entry lives in a register (entries 0 .. sp-2 in local memory): a pop hands it out at once
                                 // and the load of the entry below overlaps the node visit that follows (the pop's local-memory load
                                 // was 6 % of the kernel's stall samples)
    uint32_t node_base, slot_base, node;
    int cur_prim; uint32_t cur_tleaf, cur_tslot;
    uint32_t pend_slot, pend_cnt;  // triangle slots (relative to slot_base) recorded but not yet tested
#if TCPT_SPECULATE
    uint32_t pend2_slot, pend2_cnt;  // a second recorded leaf range: the walk may run one leaf ahead of the triangle tests
#endif
    bool need_pop;                 // the next node comes from the stack (deferred so pending triangles keep their BLAS state)

    __device__ __forceinline__ void init(float3 o_, float3 d_, float t_max_) {
        o = o_; d = d_; t_max = t_max_; limit = t_max_;
        best.prim = -1; best.t = t_max_; best.b0 = best.b1 = best.b2 = 0.0f; best.tri = 0;
        best_tleaf = best_tslot = best_bleaf = best_bslot = 0;
        ray_setup(rw, o, d); rl = rw;
        sp = 0; tos = 0; blas_sp = -1; node_base = 0; slot_base = 0; node = 0; cur_prim = -1; cur_tleaf = cur_tslot = 0;
        pend_slot = 0; pend_cnt = 0; need_pop = false;
#if TCPT_SPECULATE
        pend2_slot = 0; pend2_cnt = 0;
#endif
    }
#if TCPT_SPECULATE
    // With triangles pending the walk may still advance inside the same BLAS (the pending slots keep their meaning) as long as
    // the second range is free; leaving the BLAS, or finding a third leaf, has to wait for the triangle phase.
    __device__ __forceinline__ bool can_walk() const { return pend_cnt == 0u || (pend2_cnt == 0u && (!need_pop || sp > blas_sp)); }
#else
    __device__ __forceinline__ bool can_walk() const { return pend_cnt == 0u; }
#endif

#if TCPT_OPT_TOS
    __device__ __forceinline__ void push(uint32_t* stack, uint32_t v) { if (sp > 0) stack[sp - 1] = tos; tos = v; ++sp; }
    __device__ __forceinline__ uint32_t pop(uint32_t* stack) { const uint32_t v = tos; --sp; if (sp > 0) tos = stack[sp - 1]; return v; }
#else
    __device__ __forceinline__ void push(uint32_t* stack, uint32_t v) { stack[sp++] = v; }
    __device__ __forceinline__ uint32_t pop(uint32_t* stack) { return stack[--sp]; }
#endif

    // Tests ONE pending triangle.  ANY = Scene::intersect_p (scene.rs:93-103): returns true (ray finished) at the first accepted triangle.
    template <bool ANY, bool COUNT>
    __device__ __forceinline__ bool tri_step(const DScene& sc, uint32_t* n_tri) {
        const uint32_t bslot = pend_slot;
        pend_slot += 1; pend_cnt -= 1;
#if TCPT_SPECULATE
        if (pend_cnt == 0u && pend2_cnt != 0u) { pend_slot = pend2_slot; pend_cnt = pend2_cnt; pend2_cnt = 0u; }
#endif
        const size_t s = 3 * (size_t)(slot_base + bslot);
        const float4 v0 = __ldg(&sc.tri_verts[s]), v1 = __ldg(&sc.tri_verts[s + 1]), v2 = __ldg(&sc.tri_verts[s + 2]);
        float t, b0, b1, b2;
        if (COUNT) (*n_tri)++;
        if (tri_test(v0, v1, v2, rl, t_max, &t, &b0, &b1, &b2)) {
            if (ANY) { best.prim = 0; best.t = t; return true; }
            // total order: smaller t; then (TLAS) later leaf, earlier slot; then (BLAS) later leaf, earlier slot
            const uint32_t bleaf = __float_as_uint(v2.w);
            bool take;
            if (best.prim < 0) take = true;
            else if (t != best.t) take = t < best.t;
            else if (cur_tleaf != best_tleaf) take = cur_tleaf > best_tleaf;
            else if (cur_tslot != best_tslot) take = cur_tslot < best_tslot;
            else if (bleaf != best_bleaf) take = bleaf > best_bleaf;
            else take = bslot < best_bslot;
            if (take) {
                best.t = t; best.b0 = b0; best.b1 = b1; best.b2 = b2; best.prim = cur_prim; best.tri = __float_as_uint(v0.w);
                best_tleaf = cur_tleaf; best_tslot = cur_tslot; best_bleaf = bleaf; best_bslot = bslot;
                limit = fminf(t_max, cull_limit(t));
            }
        }
        return false;
    }

    // Visits one child-pair record (both slab tests; leaf children are recorded / TLAS items queued; descend, or defer a pop).
    // Returns true when the ray is finished (stack empty and nothing pending).
    template <bool COUNT>
    __device__ __forceinline__ bool node_step(const DScene& sc, uint32_t* stack, uint32_t* n_box) {
        if (need_pop) {
            need_pop = false;
            if (sp == blas_sp) { blas_sp = -1; node_base = 0; if (TCPT_OPT_RL) rl = rw; }  // this BLAS is exhausted: back in the TLAS (and to the Render-space ray)
            if (sp == 0) return true;
            const uint32_t top = pop(stack);
            if (top & TCPT_TLAS_ITEM_BIT) {
                const uint32_t tslot = top & ~TCPT_TLAS_ITEM_BIT;
                const int2 item = __ldg(&sc.tlas_items[tslot]);
                cur_prim = item.x; cur_tleaf = (uint32_t)item.y; cur_tslot = tslot;
                const tcpt_flat_primitive& P = sc.primitives[cur_prim];
                const tcpt_flat_geometry& G = sc.geometries[P.geometry];
                // local_to_render.inverse() * ray (primitive/impls/triangle_mesh.rs:97).  An instance without rotation or scale maps the
                // direction onto the same bits, and everything ray_setup derives (1/d, the shear constants) depends on the direction
                // alone: keep the Render-space values instead of six IEEE divisions (ray_setup was 12 % of the kernel's instructions)
                const float3 ol = xf_point(P.r2l, o), dl = xf_vector(P.r2l, d);
                if (__float_as_uint(dl.x) == __float_as_uint(d.x) && __float_as_uint(dl.y) == __float_as_uint(d.y) && __float_as_uint(dl.z) == __float_as_uint(d.z)) { rl = rw; rl.o = ol; }
                else ray_setup(rl, ol, dl);
                node_base = G.node_base; slot_base = G.slot_base;
                blas_sp = sp;
                if (G.single) { pend_slot = 0; pend_cnt = 1; need_pop = true; return false; }  // SingleTriangle: straight to the triangle test, no boxes
                node = node_base;  // entry record: tests the BLAS root box
            } else {
                node = top;
            }
        }
        const float4* rec = sc.nodes + 4 * (size_t)node;
        const float4 q0 = __ldg(rec), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const bool in_blas = blas_sp >= 0;
        const RayXform& r = TCPT_OPT_RL ? rl : (in_blas ? rl : rw);   // TCPT_OPT_RL: the CURRENT ray: Render space in the TLAS (restored at BLAS exit), instance space inside a BLAS -- no per-visit selects
        const uint32_t ref0 = __float_as_uint(q0.w), cnt0 = __float_as_uint(q1.w), ref1 = __float_as_uint(q2.w), cnt1 = __float_as_uint(q3.w);
        float te0, te1;
        const bool h0 = slab_test(q0, q1, r, limit, &te0);
        const bool h1 = (ref1 != TCPT_ABSENT) && slab_test(q2, q3, r, limit, &te1);
        if (COUNT) (*n_box) += (ref1 != TCPT_ABSENT) ? 2u : 1u;
        const bool l0 = h0 && cnt0 != 0, l1 = h1 && cnt1 != 0;
        if (in_blas) {
            // sibling leaves occupy adjacent slots (leaf order = DFS order), so both fit one pending range
            uint32_t ns = 0, nc = 0;
            if (l0) { ns = ref0; nc = cnt0 + (l1 ? cnt1 : 0u); }
            else if (l1) { ns = ref1; nc = cnt1; }
#if TCPT_SPECULATE
            if (nc != 0u) { if (pend_cnt == 0u) { pend_slot = ns; pend_cnt = nc; } else { pend2_slot = ns; pend2_cnt = nc; } }
#else
            if (nc != 0u) { pend_slot = ns; pend_cnt = nc; }
#endif
        } else {
            // TLAS leaf: queue its primitives (each opens a BLAS when popped)
            if (l0) for (uint32_t i = 0; i < cnt0; ++i) push(stack, TCPT_TLAS_ITEM_BIT | (ref0 + cnt0 - 1u - i));
            if (l1) for (uint32_t i = 0; i < cnt1; ++i) push(stack, TCPT_TLAS_ITEM_BIT | (ref1 + cnt1 - 1u - i));
        }
        const bool i0 = h0 && cnt0 == 0, i1 = h1 && cnt1 == 0;
        if (i0 && i1) {
            const uint32_t c0 = node_base + ref0, c1 = node_base + ref1;
            if (te1 < te0) { push(stack, c0); node = c1; } else { push(stack, c1); node = c0; }
        } else if (i0) node = node_base + ref0;
        else if (i1) node = node_base + ref1;
        else need_pop = true;
        return false;
    }
};

#ifndef TCPT_REFILL_IDLE_LANES
#define TCPT_REFILL_IDLE_LANES 12  // a warp fetches new rays once this many of its lanes are idle (swept 4..16 with the pooled fetch: 12 is best)
#endif
#ifndef TCPT_TRI_PHASE_LANES
#define TCPT_TRI_PHASE_LANES 8    // a warp runs a triangle iteration once this many lanes hold pending triangles
#endif

// Traces the rays of a queue with persistent warps.  `work` is a global counter zeroed before the launch.  A finished ray's
// result stays in the lane until the warp's next refill point, where `commit(i, hit)` is called for all finished lanes
// together (converged): result stores, bucket filing and shadow accumulation then cost one memory round trip per refill
// instead of one per finishing lane in a divergent branch (third profile: the filing atomic at 4 lanes was the top stall).
// Every warp of the grid must call this with all 32 lanes.
template <bool ANY, bool COUNT, class Commit>
__device__ __forceinline__ void trace_queue(const DScene& sc, const float4* __restrict__ q_o, const float4* __restrict__ q_d, uint32_t n,
                                            uint32_t* work, uint32_t* n_box, uint32_t* n_tri, Commit&& commit) {
    const uint32_t FULL = 0xffffffffu, NONE = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t stack[TCPT_TRAVERSAL_STACK];
    Traversal T;
    T.pend_cnt = 0;
    uint32_t ray = NONE, fin = NONE;
    // Ray indices are reserved from the global counter a CHUNK at a time and handed out from a warp-local pool, so most refills
    // cost no global atomic (fourth profile: the warp waiting on the counter's round trip at every refill was the top stall).
    // The chunk shrinks with the queue so that short queues (deep bounces) still spread over the whole grid.
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t chunk = n / (n_warps * 8u);
    chunk = chunk < 32u ? 32u : (chunk > 512u ? 512u : chunk);
    uint32_t pool = 0, pool_end = 0;  // warp-uniform: indices [pool, pool_end) belong to this warp
    bool drained = false;             // the global counter has passed n
    for (;;) {
        const uint32_t idle = __ballot_sync(FULL, ray == NONE);
        const uint32_t n_idle = (uint32_t)__popc(idle);
        const uint32_t left = pool_end - pool;
        const bool fetch = n_idle > left && !drained;
        uint32_t fetched = 0;
        if (fetch && lane == 0) fetched = atomicAdd(work, chunk);  // in flight while the finished rays are committed
        if (fin != NONE) { commit(fin, T.best); fin = NONE; }
        if (n_idle != 0u) {
            uint32_t base_b = 0;
            if (fetch) {
                base_b = __shfl_sync(FULL, fetched, 0);
                drained = base_b + chunk >= n;
            }
            if (ray == NONE) {
                // the first `left` idle lanes take what remains of the old chunk, the others start the new one
                const uint32_t r = (uint32_t)__popc(idle & ((1u << lane) - 1u));
                uint32_t mine = NONE;
                if (r < left) mine = pool + r;
                else if (fetch) mine = base_b + (r - left);
                if (mine < n) {
                    const float4 o = q_o[mine], d = q_d[mine];
                    T.init(f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), o.w);
                    ray = mine;
                }
            }
            if (fetch) { pool = base_b + (n_idle - left); pool_end = base_b + chunk; if (pool_end > n) pool_end = n; if (pool > pool_end) pool = pool_end; }
            else { pool += n_idle; if (pool > pool_end) pool = pool_end; }
        }
        const bool exhausted = drained && pool >= pool_end;
        if (__ballot_sync(FULL, ray != NONE) == 0u) break;  // nothing in flight and nothing left to fetch
        const uint32_t stop_at = exhausted ? 32u : (uint32_t)TCPT_REFILL_IDLE_LANES;
#if TCPT_OPT_BALLOT
        // A ray in flight either holds pending triangles or may walk (never neither), so the lanes without a ray are the complement
        // of the two masks: the masks of the next iteration double as this iteration's idle count (one ballot less per iteration).
        uint32_t n_idle_now;
        bool has_tri = ray != NONE && T.pend_cnt != 0u;
        bool walks = ray != NONE && T.can_walk();
        uint32_t tri_mask = __ballot_sync(FULL, has_tri);
        uint32_t node_mask = __ballot_sync(FULL, walks);
        do {
            bool finished = false;
            if (tri_mask != 0u && ((uint32_t)__popc(tri_mask) >= (uint32_t)TCPT_TRI_PHASE_LANES || node_mask == 0u)) {
                if (has_tri) finished = T.template tri_step<ANY, COUNT>(sc, n_tri);
            } else {
                if (walks) finished = T.template node_step<COUNT>(sc, stack, n_box);
            }
            if (finished) { fin = ray; ray = NONE; T.pend_cnt = 0; }
            has_tri = ray != NONE && T.pend_cnt != 0u;
            walks = ray != NONE && T.can_walk();
            tri_mask = __ballot_sync(FULL, has_tri);
            node_mask = __ballot_sync(FULL, walks);
            n_idle_now = 32u - (uint32_t)__popc(tri_mask | node_mask);
        } while (n_idle_now < stop_at);
#else
        uint32_t n_idle_now;
        do {
            const bool has_tri = ray != NONE && T.pend_cnt != 0u;
            const bool walks = ray != NONE && T.can_walk();
            const uint32_t tri_mask = __ballot_sync(FULL, has_tri);
            const uint32_t node_mask = __ballot_sync(FULL, walks);
            bool finished = false;
            if (tri_mask != 0u && ((uint32_t)__popc(tri_mask) >= (uint32_t)TCPT_TRI_PHASE_LANES || node_mask == 0u)) {
                if (has_tri) finished = T.template tri_step<ANY, COUNT>(sc, n_tri);
            } else {
                if (walks) finished = T.template node_step<COUNT>(sc, stack, n_box);
            }
            if (finished) { fin = ray; ray = NONE; T.pend_cnt = 0; }
            n_idle_now = (uint32_t)__popc(__ballot_sync(FULL, ray == NONE));
        } while (n_idle_now < stop_at);
#endif
    }
}

}  // namespace tcpt
