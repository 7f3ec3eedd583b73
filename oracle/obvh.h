// ORACLE (test infrastructure).  Literal restatement of the reference's generic binary BVH:
// full-sweep SAH build, pre-order flatten, exhaustive recursive closest-hit and early-out any-hit traversal.
// Follows /root/reference/scene/src/bvh.rs (build :92-230, flatten :234-295, intersect :344-444, intersect_p :447-520).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <memory>
#include <vector>

#include "omath.h"

namespace orc {

enum NodeKind : uint32_t { NODE_INNER = 0, NODE_LEAF = 1, NODE_ITEM = 2 };
struct FlatNode {
    uint32_t kind;
    Bounds bounds;   // inner / leaf
    uint32_t value;  // inner: second_offset ; leaf: item_count ; item: item id
};

struct BvhBuildTree {
    bool leaf = true;
    Bounds bounds;
    std::vector<uint32_t> items;
    std::unique_ptr<BvhBuildTree> first, second;
};

class Bvh {
   public:
    std::vector<FlatNode> nodes;

    // item_bounds[i] = BvhItem::bounds of item i; items enumerated 0..n in item_list() order.
    // `literal` = O(N^2) per axis exactly as bvh.rs:114-137; otherwise prefix/suffix sweeps (bit-identical: min/max are exact).
    void build(const std::vector<Bounds>& item_bounds, bool literal) {
        nodes.clear();
        if (item_bounds.empty()) return;
        ib_ = &item_bounds;
        literal_ = literal;
        auto root = std::make_unique<BvhBuildTree>();
        root->items.resize(item_bounds.size());
        for (size_t i = 0; i < item_bounds.size(); ++i) root->items[i] = (uint32_t)i;
        root->bounds = list_bounds(root->items, 0, root->items.size());
        build_rec(*root);
        size_t index = 0;
        flatten(*root, index);
    }
    bool empty() const { return nodes.empty(); }
    Bounds bounds() const { return nodes[0].bounds; }

    struct Hit { bool hit = false; float t = 0; };

    // closest hit; item_fn(item, &t) returns true on hit and sets t; keep(item) is called when a candidate becomes current best
    // within the recursion exactly like the reference: Node ties -> second, Leaf ties -> earlier item.
    template <class ItemFn>
    bool intersect(const Ray& ray, float t_max, ItemFn&& item_fn, float* t_out, uint64_t* payload_out, TraversalCounters* ctr) const {
        if (nodes.empty()) return false;
        Vec3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        Cand c = shrink ? traverse_shrink(0, ray, t_max, inv_dir, item_fn, ctr) : traverse(0, ray, t_max, inv_dir, item_fn, ctr);
        if (!c.hit) return false;
        *t_out = c.t;
        *payload_out = c.payload;
        return true;
    }

    template <class ItemFn>
    bool intersect_p(const Ray& ray, float t_max, ItemFn&& item_fn, TraversalCounters* ctr) const {
        if (nodes.empty()) return false;
        Vec3 inv_dir(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        return traverse_p(0, ray, t_max, inv_dir, item_fn, ctr);
    }

    // "optimised CPU" cost model for bench.py's cpu_baseline only: near child first, the far child and its triangles see
    // t_max shrunk to the best hit so far.  The combine rules are the reference's (node ties -> second child, leaf ties -> earlier
    // item) and equal-t candidates still pass the shrunk bound, so the hit is the reference's except when a candidate's scaled-t
    // comparison rounds differently against the tighter t_max.  Parity tests never use this mode.
    bool shrink = false;

   private:
    const std::vector<Bounds>* ib_ = nullptr;
    bool literal_ = true;
    struct Cand { bool hit = false; float t = 0; uint64_t payload = 0; };

    Bounds list_bounds(const std::vector<uint32_t>& items, size_t b, size_t e) const {
        Bounds r = (*ib_)[items[b]];
        for (size_t i = b + 1; i < e; ++i) r = r.merge((*ib_)[items[i]]);
        return r;
    }

    struct Split { std::vector<uint32_t> first, second; float cost; bool valid; };

    Split split(const std::vector<uint32_t>& in, int axis, float parent_area) const {
        std::vector<uint32_t> items = in;
        const std::vector<Bounds>& ib = *ib_;
        std::stable_sort(items.begin(), items.end(), [&](uint32_t a, uint32_t b) { return ib[a].center()[axis] < ib[b].center()[axis]; });
        const size_t n = items.size();
        float min_cost = INFINITY;
        size_t best = 0;
        std::vector<Bounds> suffix;
        Bounds prefix{};
        if (!literal_) {
            suffix.resize(n);
            suffix[n - 1] = ib[items[n - 1]];
            for (size_t i = n - 1; i-- > 0;) suffix[i] = ib[items[i]].merge(suffix[i + 1]);
            prefix = ib[items[0]];
        }
        for (size_t i = 1; i < n; ++i) {
            Bounds b0, b1;
            if (literal_) {
                b0 = list_bounds(items, 0, i);
                b1 = list_bounds(items, i, n);
            } else {
                if (i > 1) prefix = prefix.merge(ib[items[i - 1]]);
                b0 = prefix;
                b1 = suffix[i];
            }
            // COST_NODE + COST_LEAF*area0/parent*len0 + COST_LEAF*area1/parent*len1   (bvh.rs:125-129)
            float cost = (1.0f + ((1.0f * b0.area()) / parent_area) * (float)i) + ((1.0f * b1.area()) / parent_area) * (float)(n - i);
            if (cost < min_cost) { min_cost = cost; best = i; }
        }
        Split s;
        s.valid = best != 0;  // the reference unwraps (panics) if no split index was ever selected
        s.cost = min_cost;
        if (s.valid) {
            s.first.assign(items.begin(), items.begin() + best);
            s.second.assign(items.begin() + best, items.end());
        }
        return s;
    }

    void build_rec(BvhBuildTree& node) const {
        if (node.items.size() <= 1) return;
        float min_cost = 1.0f * (float)node.items.size();
        bool have = false;
        Split best;
        for (int axis = 0; axis < 3; ++axis) {
            Split s = split(node.items, axis, node.bounds.area());
            if (!s.valid) continue;  // (reference would panic; unreachable for non-degenerate input)
            if (s.cost < min_cost) { min_cost = s.cost; best = std::move(s); have = true; }
        }
        if (!have) return;
        node.leaf = false;
        node.first = std::make_unique<BvhBuildTree>();
        node.second = std::make_unique<BvhBuildTree>();
        node.first->items = std::move(best.first);
        node.second->items = std::move(best.second);
        node.first->bounds = list_bounds(node.first->items, 0, node.first->items.size());
        node.second->bounds = list_bounds(node.second->items, 0, node.second->items.size());
        node.items.clear();
        build_rec(*node.first);
        build_rec(*node.second);
    }

    void flatten(const BvhBuildTree& t, size_t& index) {
        if (!t.leaf) {
            size_t node_index = index;
            nodes.push_back(FlatNode{NODE_INNER, t.bounds, 0});
            index += 1;
            flatten(*t.first, index);
            nodes[node_index].value = (uint32_t)index - (uint32_t)node_index;
            flatten(*t.second, index);
        } else {
            index += t.items.size() + 1;
            nodes.push_back(FlatNode{NODE_LEAF, t.bounds, (uint32_t)t.items.size()});
            for (uint32_t it : t.items) nodes.push_back(FlatNode{NODE_ITEM, Bounds{}, it});
        }
    }

    template <class ItemFn>
    Cand traverse(size_t index, const Ray& ray, float t_max, Vec3 inv_dir, ItemFn& item_fn, TraversalCounters* ctr) const {
        const FlatNode& n = nodes[index];
        if (ctr) ctr->box_tests++;
        if (!n.bounds.intersect(ray, t_max, inv_dir)) return Cand{};
        if (n.kind == NODE_INNER) {
            Cand first = traverse(index + 1, ray, t_max, inv_dir, item_fn, ctr);
            Cand second = traverse(index + n.value, ray, t_max, inv_dir, item_fn, ctr);
            if (first.hit && second.hit) return first.t < second.t ? first : second;
            if (first.hit) return first;
            return second;
        }
        Cand best;
        for (uint32_t i = 1; i <= n.value; ++i) {
            uint32_t item = nodes[index + i].value;
            float t;
            uint64_t payload;
            if (ctr) ctr->tri_tests++;
            if (item_fn(item, t_max, &t, &payload)) {
                if (best.hit) {
                    if (t < best.t) best = Cand{true, t, payload};
                } else {
                    best = Cand{true, t, payload};
                }
            }
        }
        return best;
    }

    static bool slab_entry(const Bounds& b, const Ray& ray, float t_max, Vec3 inv_dir, float* t_entry) {
        float t0 = 0.0f, t1 = t_max;
        for (int i = 0; i < 3; ++i) {
            float t_near = (b.mn[i] - ray.o[i]) * inv_dir[i];
            float t_far = (b.mx[i] - ray.o[i]) * inv_dir[i];
            if (t_near > t_far) { float t = t_near; t_near = t_far; t_far = t; }
            t0 = t_near > t0 ? t_near : t0;
            t1 = t_far < t1 ? t_far : t1;
            if (t0 > t1) return false;
        }
        *t_entry = t0;
        return true;
    }
    // the node's own box has been tested by the caller
    template <class ItemFn>
    Cand traverse_shrink_in(size_t index, const Ray& ray, float t_max, Vec3 inv_dir, ItemFn& item_fn, TraversalCounters* ctr) const {
        const FlatNode& n = nodes[index];
        if (n.kind == NODE_INNER) {
            const size_t i1 = index + 1, i2 = index + n.value;
            float e1 = 0, e2 = 0;
            if (ctr) ctr->box_tests += 2;
            const bool h1 = slab_entry(nodes[i1].bounds, ray, t_max, inv_dir, &e1), h2 = slab_entry(nodes[i2].bounds, ray, t_max, inv_dir, &e2);
            Cand first, second;
            if (h1 && h2) {
                if (e2 < e1) {
                    second = traverse_shrink_in(i2, ray, t_max, inv_dir, item_fn, ctr);
                    const float lim = second.hit ? second.t : t_max;
                    if (!second.hit || e1 <= lim) first = traverse_shrink_in(i1, ray, lim, inv_dir, item_fn, ctr);
                } else {
                    first = traverse_shrink_in(i1, ray, t_max, inv_dir, item_fn, ctr);
                    const float lim = first.hit ? first.t : t_max;
                    if (!first.hit || e2 <= lim) second = traverse_shrink_in(i2, ray, lim, inv_dir, item_fn, ctr);
                }
            } else if (h1) first = traverse_shrink_in(i1, ray, t_max, inv_dir, item_fn, ctr);
            else if (h2) second = traverse_shrink_in(i2, ray, t_max, inv_dir, item_fn, ctr);
            if (first.hit && second.hit) return first.t < second.t ? first : second;
            if (first.hit) return first;
            return second;
        }
        Cand best;
        float lim = t_max;
        for (uint32_t i = 1; i <= n.value; ++i) {
            uint32_t item = nodes[index + i].value;
            float t;
            uint64_t payload;
            if (ctr) ctr->tri_tests++;
            if (item_fn(item, lim, &t, &payload)) {
                if (best.hit) {
                    if (t < best.t) { best = Cand{true, t, payload}; lim = t; }
                } else {
                    best = Cand{true, t, payload}; lim = t;
                }
            }
        }
        return best;
    }
    template <class ItemFn>
    Cand traverse_shrink(size_t index, const Ray& ray, float t_max, Vec3 inv_dir, ItemFn& item_fn, TraversalCounters* ctr) const {
        float e;
        if (ctr) ctr->box_tests++;
        if (!slab_entry(nodes[index].bounds, ray, t_max, inv_dir, &e)) return Cand{};
        return traverse_shrink_in(index, ray, t_max, inv_dir, item_fn, ctr);
    }

    template <class ItemFn>
    bool traverse_p(size_t index, const Ray& ray, float t_max, Vec3 inv_dir, ItemFn& item_fn, TraversalCounters* ctr) const {
        const FlatNode& n = nodes[index];
        if (ctr) ctr->box_tests++;
        if (!n.bounds.intersect(ray, t_max, inv_dir)) return false;
        if (n.kind == NODE_INNER) {
            if (traverse_p(index + 1, ray, t_max, inv_dir, item_fn, ctr)) return true;
            return traverse_p(index + n.value, ray, t_max, inv_dir, item_fn, ctr);
        }
        for (uint32_t i = 1; i <= n.value; ++i) {
            if (ctr) ctr->tri_tests++;
            if (item_fn(nodes[index + i].value, t_max)) return true;
        }
        return false;
    }
};

}  // namespace orc
