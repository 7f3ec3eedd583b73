// Host BVH builder (product code): reproduces the TOPOLOGY of the reference's full-sweep SAH builder
// (/root/reference/scene/src/bvh.rs:92-230) bit for bit, in O(N log^2 N) instead of the reference's O(N^2) per split level.
//
// Why it is exact: the reference stably sorts a copy of the node's CURRENT item order by box
// centre on the axis, then for every split index folds the boxes of both halves and compares
//     cost = (1 + ((1*area0)/parent_area)*n0) + ((1*area1)/parent_area)*n1        bvh.rs:125-129
// with strict `<` (first minimum along the sweep wins, earlier axis wins ties, "no split" wins ties).  Folding boxes is a
// sequence of exact min/max operations, so a prefix sweep and a suffix sweep give bit-identical boxes; the cost expression
// is evaluated verbatim.  std::stable_sort has the tie behaviour of Rust's sort_by.
#pragma once
#include <algorithm>
#include <cstdint>
#include <future>
#include <memory>
#include <vector>

#include "host_math.h"

namespace tcpt {

struct BuildNode {
    Box box;
    uint32_t second = 0;      // inner: index of the second child in the node-only pre-order array
    uint32_t first_item = 0;  // leaf: first slot
    uint32_t count = 0;       // leaf: item count (> 0); inner: 0
};

struct BuiltBvh {
    std::vector<BuildNode> nodes;  // pre-order, inner/leaf only
    std::vector<uint32_t> items;   // leaf items in leaf order (slot -> item id)
    uint32_t depth = 0;
};

class SahBuilder {
   public:
    explicit SahBuilder(const std::vector<Box>& item_boxes) : boxes_(item_boxes) {}

    BuiltBvh build() {
        BuiltBvh out;
        if (boxes_.empty()) return out;
        std::vector<uint32_t> items(boxes_.size());
        for (size_t i = 0; i < items.size(); ++i) items[i] = (uint32_t)i;
        Tree root = build_subtree(std::move(items), 0);
        emit(root, out, 1);
        return out;
    }

   private:
    struct Tree {
        Box box;
        std::vector<uint32_t> items;  // leaf
        std::unique_ptr<Tree> a, b;   // inner
    };
    const std::vector<Box>& boxes_;

    Box fold(const uint32_t* ids, size_t n) const {
        if (n == 0) return Box{{0, 0, 0}, {0, 0, 0}};
        Box r = boxes_[ids[0]];
        for (size_t i = 1; i < n; ++i) r.grow(boxes_[ids[i]]);
        return r;
    }

    struct Candidate {
        float cost = INFINITY;
        size_t split = 0;
        std::vector<uint32_t> order;
    };

    Candidate sweep(const std::vector<uint32_t>& in, int axis, float parent_area) const {
        Candidate c;
        c.order = in;
        std::stable_sort(c.order.begin(), c.order.end(), [&](uint32_t l, uint32_t r) { return boxes_[l].centre(axis) < boxes_[r].centre(axis); });
        const size_t n = c.order.size();
        std::vector<Box> suffix(n);
        suffix[n - 1] = boxes_[c.order[n - 1]];
        for (size_t i = n - 1; i-- > 0;) { suffix[i] = boxes_[c.order[i]]; suffix[i].grow(suffix[i + 1]); }
        Box prefix = boxes_[c.order[0]];
        for (size_t i = 1; i < n; ++i) {
            if (i > 1) prefix.grow(boxes_[c.order[i - 1]]);
            const float cost = (1.0f + ((1.0f * prefix.half_area2()) / parent_area) * (float)i) + ((1.0f * suffix[i].half_area2()) / parent_area) * (float)(n - i);
            if (cost < c.cost) { c.cost = cost; c.split = i; }
        }
        return c;
    }

    Tree build_subtree(std::vector<uint32_t> items, int level) const {
        Tree t;
        t.box = fold(items.data(), items.size());
        if (items.size() <= 1) { t.items = std::move(items); return t; }
        float best_cost = 1.0f * (float)items.size();  // cost of not splitting (bvh.rs:187)
        Candidate best;
        bool have = false;
        const float parent_area = t.box.half_area2();
        if (items.size() > 20000) {
            std::future<Candidate> f[3];
            for (int axis = 0; axis < 3; ++axis) f[axis] = std::async(std::launch::async, [&, axis] { return sweep(items, axis, parent_area); });
            for (int axis = 0; axis < 3; ++axis) {
                Candidate c = f[axis].get();
                if (c.split != 0 && c.cost < best_cost) { best_cost = c.cost; best = std::move(c); have = true; }
            }
        } else {
            for (int axis = 0; axis < 3; ++axis) {
                Candidate c = sweep(items, axis, parent_area);
                if (c.split != 0 && c.cost < best_cost) { best_cost = c.cost; best = std::move(c); have = true; }
            }
        }
        if (!have) { t.items = std::move(items); return t; }
        std::vector<uint32_t> left(best.order.begin(), best.order.begin() + best.split);
        std::vector<uint32_t> right(best.order.begin() + best.split, best.order.end());
        t.a.reset(new Tree());
        t.b.reset(new Tree());
        if (level < 3 && left.size() + right.size() > 4096) {
            auto fut = std::async(std::launch::async, [&] { return build_subtree(std::move(left), level + 1); });
            *t.b = build_subtree(std::move(right), level + 1);
            *t.a = fut.get();
        } else {
            *t.a = build_subtree(std::move(left), level + 1);
            *t.b = build_subtree(std::move(right), level + 1);
        }
        return t;
    }

    // pre-order emission (bvh.rs:234-295 without the inline item records)
    void emit(const Tree& t, BuiltBvh& out, uint32_t depth) const {
        if (depth > out.depth) out.depth = depth;
        const uint32_t self = (uint32_t)out.nodes.size();
        out.nodes.push_back(BuildNode{t.box, 0, 0, 0});
        if (t.a) {
            emit(*t.a, out, depth + 1);
            out.nodes[self].second = (uint32_t)out.nodes.size();
            emit(*t.b, out, depth + 1);
        } else {
            out.nodes[self].first_item = (uint32_t)out.items.size();
            out.nodes[self].count = (uint32_t)t.items.size();
            for (uint32_t it : t.items) out.items.push_back(it);
        }
    }
};

// Fast binned-SAH builder for the synthetic triangle soups: the reference's builder is O(N^2) and
// cannot define a topology at 1M+ triangles, so soups are outside topology-parity scope.
class BinnedBuilder {
   public:
    explicit BinnedBuilder(const std::vector<Box>& item_boxes, uint32_t leaf_size = 4) : boxes_(item_boxes), leaf_(leaf_size) {}
    BuiltBvh build() {
        BuiltBvh out;
        if (boxes_.empty()) return out;
        ids_.resize(boxes_.size());
        for (size_t i = 0; i < ids_.size(); ++i) ids_[i] = (uint32_t)i;
        out.nodes.reserve(boxes_.size());
        recurse(0, ids_.size(), out, 1);
        out.items = ids_;
        return out;
    }

   private:
    const std::vector<Box>& boxes_;
    uint32_t leaf_;
    std::vector<uint32_t> ids_;
    void recurse(size_t b, size_t e, BuiltBvh& out, uint32_t depth) {
        if (depth > out.depth) out.depth = depth;
        const uint32_t self = (uint32_t)out.nodes.size();
        Box box = boxes_[ids_[b]], cb{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
        for (size_t i = b; i < e; ++i) {
            box.grow(boxes_[ids_[i]]);
            V3 c = {boxes_[ids_[i]].centre(0), boxes_[ids_[i]].centre(1), boxes_[ids_[i]].centre(2)};
            cb.lo = min3(cb.lo, c); cb.hi = max3(cb.hi, c);
        }
        out.nodes.push_back(BuildNode{box, 0, 0, 0});
        const size_t n = e - b;
        auto make_leaf = [&] { out.nodes[self].first_item = (uint32_t)b; out.nodes[self].count = (uint32_t)n; };
        if (n <= leaf_) { make_leaf(); return; }
        V3 ext = sub(cb.hi, cb.lo);
        int axis = ext.x > ext.y ? (ext.x > ext.z ? 0 : 2) : (ext.y > ext.z ? 1 : 2);
        if (!(ext.get(axis) > 0.0f)) {
            if (n <= 64) { make_leaf(); return; }
            size_t mid = b + n / 2;  // identical centres: split by count
            recurse(b, mid, out, depth + 1);
            out.nodes[self].second = (uint32_t)out.nodes.size();
            recurse(mid, e, out, depth + 1);
            return;
        }
        constexpr int NB = 16;
        Box bb[NB];
        uint32_t cnt[NB] = {0};
        for (auto& x : bb) x = Box{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
        const float k = NB * (1.0f - 1e-6f) / ext.get(axis), lo = cb.lo.get(axis);
        auto bin_of = [&](uint32_t id) { int v = (int)((boxes_[id].centre(axis) - lo) * k); return v < 0 ? 0 : (v >= NB ? NB - 1 : v); };
        for (size_t i = b; i < e; ++i) { int bi = bin_of(ids_[i]); cnt[bi]++; bb[bi].grow(boxes_[ids_[i]]); }
        float best = INFINITY; int best_split = -1;
        Box r[NB]; uint32_t rc[NB];
        Box acc{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; uint32_t c = 0;
        for (int i = NB - 1; i > 0; --i) { acc.grow(bb[i]); c += cnt[i]; r[i] = acc; rc[i] = c; }
        acc = Box{{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}}; c = 0;
        for (int i = 0; i < NB - 1; ++i) {
            acc.grow(bb[i]); c += cnt[i];
            if (c == 0 || rc[i + 1] == 0) continue;
            float cost = acc.half_area2() * c + r[i + 1].half_area2() * rc[i + 1];
            if (cost < best) { best = cost; best_split = i; }
        }
        size_t mid;
        if (best_split < 0) mid = b + n / 2;
        else mid = std::partition(ids_.begin() + b, ids_.begin() + e, [&](uint32_t id) { return bin_of(id) <= best_split; }) - ids_.begin();
        if (mid == b || mid == e) mid = b + n / 2;
        recurse(b, mid, out, depth + 1);
        out.nodes[self].second = (uint32_t)out.nodes.size();
        recurse(mid, e, out, depth + 1);
    }
};

}  // namespace tcpt
