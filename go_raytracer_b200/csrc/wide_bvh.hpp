// wide_bvh.hpp — the device-internal 4-wide BVH, built at upload from the ABI's binary GrtNode array.
//
// The ABI (include/grt.h) carries BuildBVH's binary tree in the reference's child order (bvh.go:35-61).  A binary
// tree costs one dependent memory round trip per level, and on the 10^5..10^6-node trees of configs C4/C5 the extend
// kernel was bound by exactly that latency (profiles/README.md, round 1).  Here every binary node absorbs
// grandchildren until it has four children (the child with the largest box is opened first), children keep their
// left-to-right order, and each child — inner node OR leaf — carries its own box:
//
//   * the closest hit is unchanged: boxes only cull, and every box is conservative (rounded outwards);
//   * a subtree that contains a constantMedium is flagged "in order": its children are visited in the reference's
//     order, because a medium test draws a random number (medium.go:47) and the draw index must not depend on the
//     visiting order; a culled medium never reaches its draw (both boundary hits lie inside its box), so culling is
//     unobservable there too;
//   * everywhere else children are visited nearest first; which of two EXACTLY equidistant primitives wins then
//     differs from bvh.go:73-79 — the documented tie, already true of round 1's axis hints;
//   * a leaf run of up to 8 consecutive primitives is named in the child ref itself (DREF_RUN), so reaching the
//     primitives costs no list-entry fetch;
//   * a binary subtree WITHOUT a medium is only a set of surfaces as far as the closest hit goes, so its topology is
//     not kept at all: the leaves are regrouped top-down by the surface area heuristic (binned, 16 bins per axis; an
//     exact sweep for sets of <= 16 that keeps groups of four together), straight into 4-wide nodes.  BuildBVH splits
//     at the object median of the longest axis, which puts the 1000-unit ground sphere of the book-1 cover next to
//     0.2-unit spheres in every upper node: regrouping cuts the box tests per segment from 73 to 24 there (+21 %
//     paths/s) and from 62 to 54 on the 1M-triangle mesh (+7 %); GRT_WIDE_SAH=0 keeps the widened BuildBVH tree;
//   * the same holds BETWEEN two media of a list: a medium test sees the closest hit among the items before it
//     (hittable.go:129-136, medium.go:38-47) and nothing of the items after it, so every maximal run of medium-free list
//     items is one set of surfaces, regrouped into one nearest-first subtree (nested lists and BVHs dissolved into it),
//     while the media and the runs keep their list order.  Book 2's world list becomes [6 surfaces + 400 boxes], medium,
//     medium, [2 spheres + 1000 spheres]: +10 % paths/s (GRT_WIDE_SAH_LISTS=0: lists keep their items one by one).
//
// Node layout (128 bytes = one L1 line, 8 x float4): lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] ref[4] meta[4];
// meta[0] bit 0 = children may be visited nearest first, meta[1] = number of children.  Empty slots hold the empty
// box (+inf, -inf) and a NONE ref.  Nodes are numbered breadth first from the roots, so the first k nodes are the
// top of the tree: that prefix is what the kernels stage in shared memory.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include "../../include/grt.h"

namespace grt {
namespace wide {

// device-internal child ref of a primitive run: bit 31 | type << 28 | (count - 1) << 25 | first index (25 bits)
#define GRT_DREF_RUN_BIT 0x80000000u
#define GRT_DREF_RUN_MAX_COUNT 8u
#define GRT_DREF_RUN_INDEX_BITS 25u
#define GRT_DREF_RUN_INDEX_MASK ((1u << GRT_DREF_RUN_INDEX_BITS) - 1u)
#define GRT_WNODE_FLOATS 32

struct Box {
    double lo[3], hi[3];
    static Box none() { Box b; for (int a = 0; a < 3; a++) { b.lo[a] = std::numeric_limits<double>::infinity(); b.hi[a] = -std::numeric_limits<double>::infinity(); } return b; }
    static Box all() { Box b; for (int a = 0; a < 3; a++) { b.lo[a] = -std::numeric_limits<double>::infinity(); b.hi[a] = std::numeric_limits<double>::infinity(); } return b; }
    // (a NaN coordinate is ignored, like fmin / fmax would)
    void add(const double p[3]) { for (int a = 0; a < 3; a++) { lo[a] = p[a] < lo[a] ? p[a] : lo[a]; hi[a] = p[a] > hi[a] ? p[a] : hi[a]; } }
    void add(const Box& o) { for (int a = 0; a < 3; a++) { lo[a] = o.lo[a] < lo[a] ? o.lo[a] : lo[a]; hi[a] = o.hi[a] > hi[a] ? o.hi[a] : hi[a]; } }
    bool empty() const { return !(lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]); }
    bool finite() const { for (int a = 0; a < 3; a++) if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return false; return true; }
    double area() const {
        if (empty()) return 0.0;
        if (!finite()) return std::numeric_limits<double>::infinity();
        double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return x * y + y * z + z * x;
    }
};

struct Result {
    std::vector<float> wnodes;         // GRT_WNODE_FLOATS per node
    std::vector<uint32_t> node_map;    // binary node index -> wide node index (0xFFFFFFFF: absorbed into its parent)
    std::vector<uint32_t> list_map;    // list entry index -> wide node index of the list that starts there (0xFFFFFFFF: none)
    std::vector<uint32_t> media_boundary;   // device ref of every medium's boundary
    uint32_t root = 0;                 // device ref of the world
    uint32_t n_wide = 0;
    int need_main = 0, need_boundary = 0;   // traversal stack entries a closest-hit query can hold at once
    std::string error;
};

class Builder {
   public:
    // nodes / entries / media: already remapped to device refs (LIST refs index `entries`, pairs {first ref | LAST, count})
    Builder(const GrtScene* s, const std::vector<GrtNode>& nodes, const std::vector<uint32_t>& entries, const std::vector<GrtMedium>& media)
        : S(s), N(nodes), E(entries), M(media) {}

    bool run(uint32_t root, Result& R) {
        const uint32_t NONE_IDX = 0xFFFFFFFFu;
        Rp = &R;
        R.node_map.assign(N.size(), NONE_IDX);
        R.list_map.assign(E.size() / 2, NONE_IDX);
        hasMedium.assign(N.size(), -1);
        // height of every binary subtree (the flattener emits parents before children: one reverse pass)
        height.assign(N.size(), 0);
        for (size_t b = N.size(); b-- > 0;) {
            int hgt = 0;
            const uint32_t c[2] = {N[b].left & ~GRT_NODE_HINT_BIT, N[b].right & ~GRT_NODE_HINT_BIT};
            for (int k = 0; k < 2; k++) if (type(c[k]) == GRT_REF_NODE && idx(c[k]) < N.size()) hgt = std::max(hgt, (idx(c[k]) > b ? height[idx(c[k])] : 0) + 1);
            height[b] = hgt;
        }
        // HittableLists become wide nodes too when the scene has a BVH at all (list-only scenes — the Cornell boxes —
        // keep their ordered runs and never compile the node code): every item gets its own box, a list without a
        // medium is visited nearest first, one with a medium in list order (hittable.go:129-136).
        wideLists = !N.empty() && !(getenv("GRT_WIDE_LISTS") && atoi(getenv("GRT_WIDE_LISTS")) == 0);   // (env: A/B knob)
        // work items: a binary node to collapse, or a consecutive range of a list's items to group
        struct Work { bool list; uint32_t w, b; std::vector<uint32_t> items; size_t lo = 0, hi = 0; Box box = Box::none(); bool haveBox = false; };   // w: wide index to fill; b: binary node; [lo, hi): range of `leaves` (SAH rebuild)
        std::deque<Work> queue;
        auto alloc = [&]() -> uint32_t {
            uint32_t w = R.n_wide++;
            R.wnodes.resize((size_t)R.n_wide * GRT_WNODE_FLOATS, 0.0f);
            return w;
        };
        std::function<uint32_t(uint32_t)> deviceRef = [&](uint32_t r) -> uint32_t {   // a remapped ABI ref as the device sees it
            if (type(r) == GRT_REF_NODE) {
                uint32_t b = idx(r);
                if (R.node_map[b] == NONE_IDX) { R.node_map[b] = alloc(); queue.push_back(Work{false, R.node_map[b], b, {}}); }
                return GRT_MAKE_REF(GRT_REF_NODE, R.node_map[b]);
            }
            if (type(r) == GRT_REF_LIST && wideLists) {
                uint32_t e0 = idx(r);
                if (R.list_map[e0] == NONE_IDX) {
                    std::vector<uint32_t> items;
                    listItems(e0, items);
                    if (items.empty()) return GRT_MAKE_REF(GRT_REF_NONE, 0);
                    if (items.size() == 1 && (items[0] & GRT_DREF_RUN_BIT)) return items[0];   // a list that is one short run
                    R.list_map[e0] = alloc();
                    queue.push_back(Work{true, R.list_map[e0], 0, std::move(items)});
                }
                return GRT_MAKE_REF(GRT_REF_NODE, R.list_map[e0]);
            }
            if (type(r) == GRT_REF_LIST) {   // a list kept as a list: the nodes it names still have to be built
                for (uint32_t k = idx(r);; k++) {
                    const uint32_t e = E[2 * k], first = e & ~GRT_LIST_LAST;
                    if (type(first) == GRT_REF_NODE || type(first) == GRT_REF_LIST) deviceRef(first);
                    if (e & GRT_LIST_LAST) break;
                }
                return r;
            }
            // run refs are decoded by the node code only: list-only scenes keep plain refs
            return wideLists ? asRun(r) : r;
        };
        R.root = deviceRef(root);
        R.media_boundary.clear();
        // a medium whose boundary is a single primitive keeps the plain ref: the device solves it directly (dev_trace.cuh)
        for (const GrtMedium& m : M) R.media_boundary.push_back(isPrim(type(m.boundary)) ? m.boundary : deviceRef(m.boundary));
        std::vector<uint32_t> kids;
        while (!queue.empty()) {
            Work wk = std::move(queue.front());
            queue.pop_front();
            bool ordered;
            const uint32_t w = wk.w;
            kids.clear();
            std::vector<Box> kbox;
            std::vector<uint32_t> kref;
            if (wk.list && sah && sahLists) {
                // A medium test only depends on the closest hit among the items BEFORE it (hittable.go:129-136 hands every
                // item rayT.Max = closest so far; medium.go:38-47 clamps to it and then draws), so the items between two
                // media are again just a set of surfaces: every maximal run of medium-free items is regrouped by SAH into
                // one nearest-first subtree (nested lists and BVHs dissolved into it); the media and the runs keep their
                // list order.  A list without any medium becomes that subtree itself.
                std::vector<uint32_t>& items = wk.items;
                std::vector<uint32_t> out;
                size_t i = 0;
                while (i < items.size()) {
                    size_t j = i;
                    while (j < items.size() && !subtreeHasMedium(items[j]) && !isSet(items[j])) j++;
                    const bool whole = i == 0 && j == items.size();
                    if (j - i >= 2 || (j - i == 1 && !(items[i] & GRT_DREF_RUN_BIT) && type(items[i]) == GRT_REF_LIST)) {
                        const size_t lo = leaves.size();
                        for (size_t k = i; k < j; k++) appendItemLeaves(items[k], 0);
                        const size_t hi = leaves.size();
                        if (hi - lo >= 2 && whole) {
                            wk.list = false; wk.lo = lo; wk.hi = hi; wk.box = rangeBox(lo, hi); wk.haveBox = true;
                            break;
                        } else if (hi - lo >= 2) {
                            const uint32_t cw = alloc();
                            Work sub{false, cw, 0, {}};
                            sub.lo = lo; sub.hi = hi; sub.box = rangeBox(lo, hi); sub.haveBox = true;
                            setBox[cw] = sub.box;
                            queue.push_back(std::move(sub));
                            out.push_back(GRT_DREF_RUN_BIT | ((uint32_t)GRT_REF_NODE << GRT_REF_SHIFT) | cw);
                        } else {
                            leaves.resize(lo);
                            for (size_t k = i; k < j; k++) out.push_back(items[k]);
                        }
                    } else for (size_t k = i; k < j; k++) out.push_back(items[k]);
                    if (j < items.size()) { out.push_back(items[j]); j++; }
                    i = j;
                }
                if (wk.list) items.swap(out);
            }
            if (!wk.list && wk.hi == 0 && sah && !subtreeHasMedium(GRT_MAKE_REF(GRT_REF_NODE, wk.b)) && height[wk.b] >= sahMinHeight)
                gatherLeaves(wk.b, wk.lo, wk.hi);   // a medium-free binary subtree: only its SET of leaves matters, regroup it
            if (wk.hi > wk.lo) {
                // Surface-area-heuristic regrouping (binned, 16 bins per axis on the leaf centroids): the set is split until
                // it has four parts, always the part with the largest box next; a part of one leaf becomes a leaf child.
                // The closest hit over a set of surfaces does not depend on how the set is boxed (header comment).
                ordered = true;
                struct Part { size_t lo, hi; Box box; };
                std::vector<Part> parts{{wk.lo, wk.hi, wk.haveBox ? wk.box : rangeBox(wk.lo, wk.hi)}};
                while (parts.size() < 4) {
                    int best = -1;
                    double bestA = -1.0;
                    for (size_t i = 0; i < parts.size(); i++) {
                        const size_t cnt = parts[i].hi - parts[i].lo;
                        // a part of 2..4 leaves is a full node of its own unless ALL of its leaves fit into this node
                        if (cnt < 2 || (cnt <= 4 && sahPack && parts.size() - 1 + cnt > 4)) continue;
                        const double a = parts[i].box.area();
                        if (a > bestA) { bestA = a; best = (int)i; }
                    }
                    if (best < 0) break;
                    const size_t lo = parts[best].lo, hi = parts[best].hi;
                    if (hi - lo <= 4 && sahPack) {
                        parts.erase(parts.begin() + best);
                        for (size_t i = lo; i < hi; i++) parts.insert(parts.begin() + best + (i - lo), Part{i, i + 1, leaves[i].box});
                        continue;
                    }
                    Box lb, rb;
                    const size_t mid = sahSplit(lo, hi, lb, rb);
                    parts[best] = Part{lo, mid, lb};
                    parts.insert(parts.begin() + best + 1, Part{mid, hi, rb});
                }
                for (auto& pr : parts) {
                    if (pr.box.empty()) continue;
                    if (pr.hi - pr.lo == 1) {
                        kbox.push_back(leaves[pr.lo].box);
                        kref.push_back(deviceRef(leaves[pr.lo].ref));
                    } else {
                        uint32_t cw = alloc();
                        Work sub{false, cw, 0, {}};
                        sub.lo = pr.lo; sub.hi = pr.hi; sub.box = pr.box; sub.haveBox = true;
                        queue.push_back(std::move(sub));
                        kbox.push_back(pr.box);
                        kref.push_back(GRT_MAKE_REF(GRT_REF_NODE, cw));
                    }
                }
            } else if (!wk.list) {
                const uint32_t b = wk.b;
                pushChildren(kids, b);
                // open the child with the largest box (the surface-area heuristic's choice) among the children whose
                // subtree is within two levels of the tallest, until there are four; children keep their left-to-right
                // order.  Opening by area alone can leave the deepest branch one binary level per wide level (a 2^20-leaf
                // tree: 20 wide levels x 3 waiting siblings overflows the traversal stack); by height alone it ignores
                // the boxes (-7 % on the 1M-triangle mesh).
                while (kids.size() < 4) {
                    int best = -1;
                    double bestKey = -1.0;
                    int maxH = -1;
                    for (size_t i = 0; i < kids.size(); i++) if (type(kids[i]) == GRT_REF_NODE) maxH = std::max(maxH, height[idx(kids[i])]);
                    // near the leaves (subtrees of height <= bottomH) the tallest child is opened instead, which packs four
                    // leaves per node where area order would strand pairs
                    const bool byHeight = mode == 1 || (mode == 2 && maxH <= bottomH);
                    for (int pass = 0; pass < 2 && best < 0; pass++)   // second pass: no tall candidate fits, take any that does
                        for (size_t i = 0; i < kids.size(); i++) {
                            if (type(kids[i]) != GRT_REF_NODE || (pass == 0 && height[idx(kids[i])] < maxH - 2)) continue;
                            std::vector<uint32_t> sub;
                            pushChildren(sub, idx(kids[i]));
                            if (kids.size() - 1 + sub.size() > 4) continue;
                            const double a = nodeBox(idx(kids[i])).area();
                            const double key = byHeight ? (double)height[idx(kids[i])] + a / (1.0 + a) * 0.5 : a;
                            if (key > bestKey) { bestKey = key; best = (int)i; }
                        }
                    if (best < 0) break;
                    std::vector<uint32_t> sub;
                    pushChildren(sub, idx(kids[best]));
                    kids.erase(kids.begin() + best);
                    kids.insert(kids.begin() + best, sub.begin(), sub.end());
                }
                ordered = !subtreeHasMedium(GRT_MAKE_REF(GRT_REF_NODE, b));
                for (uint32_t c : kids) {
                    if (type(c) == GRT_REF_NONE) continue;
                    Box bx;
                    if (type(c) == GRT_REF_NODE) {
                        // the ABI's node boxes are the fp64 geometry rounded outwards; the fp32 primitives the device tests
                        // (v0, e0, e1 rounded separately) can sit an ulp OF THE LARGEST COORDINATE outside of them
                        bx = nodeBox(idx(c));
                        grow(bx, 5e-7);
                    } else bx = refBox(c);
                    if (bx.empty()) continue;    // nothing below this child can be hit
                    kbox.push_back(bx);
                    kref.push_back(deviceRef(c));
                }
            } else {
                // a range of list items: up to four go straight into the node, more are split into four consecutive groups
                std::vector<uint32_t>& items = wk.items;
                ordered = true;
                for (uint32_t it : items) if (subtreeHasMedium(it)) ordered = false;
                const size_t n = items.size(), groups = n <= 4 ? n : 4;
                size_t at = 0;
                for (size_t gi = 0; gi < groups; gi++) {
                    const size_t cnt = (n - at + (groups - gi) - 1) / (groups - gi);
                    std::vector<uint32_t> sub(items.begin() + at, items.begin() + at + cnt);
                    at += cnt;
                    Box bx = Box::none();
                    for (uint32_t it : sub) bx.add(itemBox(it));
                    if (bx.empty()) continue;
                    if (cnt == 1) kref.push_back(isSet(sub[0]) ? GRT_MAKE_REF(GRT_REF_NODE, sub[0] & GRT_DREF_RUN_INDEX_MASK) : ((sub[0] & GRT_DREF_RUN_BIT) ? sub[0] : deviceRef(sub[0])));
                    else {
                        uint32_t cw = alloc();
                        queue.push_back(Work{true, cw, 0, std::move(sub)});
                        kref.push_back(GRT_MAKE_REF(GRT_REF_NODE, cw));
                    }
                    kbox.push_back(bx);
                }
            }
            // write the node
            float* d = R.wnodes.data() + (size_t)w * GRT_WNODE_FLOATS;
            const float INF = std::numeric_limits<float>::infinity();
            uint32_t refs[4], meta[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++) {
                for (int a = 0; a < 3; a++) { d[4 * a + k] = INF; d[12 + 4 * a + k] = -INF; }
                refs[k] = GRT_MAKE_REF(GRT_REF_NONE, 0);
            }
            uint32_t nk = 0;
            for (size_t k = 0; k < kref.size() && nk < 4; k++) {
                if (type(kref[k]) == GRT_REF_NONE && !(kref[k] & GRT_DREF_RUN_BIT)) continue;
                for (int a = 0; a < 3; a++) { d[4 * a + nk] = down(kbox[k].lo[a]); d[12 + 4 * a + nk] = up(kbox[k].hi[a]); }
                refs[nk++] = kref[k];
            }
            meta[0] = ordered ? 1u : 0u;
            meta[1] = nk;
            memcpy(d + 24, refs, 16);
            memcpy(d + 28, meta, 16);
        }
        // stack needs
        needW.assign(R.n_wide, -1);
        R.need_main = needRef(R.root) + 2;
        R.need_boundary = 0;
        for (uint32_t b : R.media_boundary) R.need_boundary = std::max(R.need_boundary, needRef(b) + 2);
        return R.error.empty();
    }

    // a remapped ABI ref held outside the node array (list entries of list-only scenes) as the device sees it
    uint32_t wideRef(uint32_t r) const {
        if (type(r) == GRT_REF_NODE) return GRT_MAKE_REF(GRT_REF_NODE, Rp->node_map[idx(r)]);
        return r;
    }

   private:
    const GrtScene* S;
    const std::vector<GrtNode>& N;
    const std::vector<uint32_t>& E;
    const std::vector<GrtMedium>& M;
    std::vector<int> hasMedium, needW, height;
    Result* Rp = nullptr;
    bool wideLists = false;
    // collapse order (A/B knob, results are identical): 0 = largest box first, 1 = tallest subtree first, 2 = largest box,
    // tallest near the leaves
    int mode = getenv("GRT_WIDE_MODE") ? atoi(getenv("GRT_WIDE_MODE")) : 2;
    int bottomH = getenv("GRT_WIDE_BOTTOM") ? atoi(getenv("GRT_WIDE_BOTTOM")) : 2;

    // ---- SAH regrouping of medium-free subtrees -----------------------------------------------------------------
   public:
    // 0: keep BuildBVH's topology (longest-axis object median, bvh.go:35-61) and only widen it; 1: regroup by SAH
    int sah = getenv("GRT_WIDE_SAH") ? atoi(getenv("GRT_WIDE_SAH")) : 1;
   private:
    int sahMinHeight = 2;
    int sahPack = getenv("GRT_WIDE_PACK") ? atoi(getenv("GRT_WIDE_PACK")) : 1;
    int sahSweepMax = getenv("GRT_WIDE_SWEEP") ? atoi(getenv("GRT_WIDE_SWEEP")) : 16;
    struct SahLeaf { uint32_t ref; Box box; double c[3]; };
    std::vector<SahLeaf> leaves;

    // a list item that stands for an SAH-regrouped run of medium-free items: RUN bit | NODE type | wide node index
    int sahLists = getenv("GRT_WIDE_SAH_LISTS") ? atoi(getenv("GRT_WIDE_SAH_LISTS")) : 1;
    std::map<uint32_t, Box> setBox;
    static bool isSet(uint32_t it) { return (it & GRT_DREF_RUN_BIT) && ((it >> GRT_REF_SHIFT) & 7u) == GRT_REF_NODE; }
    void pushLeaf(uint32_t ref) {
        SahLeaf L;
        L.ref = ref;
        L.box = refBox(ref);
        if (L.box.empty()) return;
        for (int a = 0; a < 3; a++) L.c[a] = 0.5 * (L.box.lo[a] + L.box.hi[a]);
        leaves.push_back(L);
    }
    // the surfaces below one medium-free list item, appended to `leaves`: runs expanded, nested lists and BVHs dissolved
    void appendItemLeaves(uint32_t it, int depth) {
        if (depth > 16) { pushLeaf(it); return; }
        if (it & GRT_DREF_RUN_BIT) {
            const uint32_t t = (it >> GRT_REF_SHIFT) & 7u, cnt = ((it >> GRT_DREF_RUN_INDEX_BITS) & 7u) + 1u, first = it & GRT_DREF_RUN_INDEX_MASK;
            for (uint32_t j = 0; j < cnt; j++) pushLeaf(GRT_MAKE_REF(t, first + j));
            return;
        }
        if (type(it) == GRT_REF_NODE && idx(it) < N.size()) { size_t lo, hi; gatherLeaves(idx(it), lo, hi, true); return; }
        if (type(it) == GRT_REF_LIST) {
            std::vector<uint32_t> sub;
            listItems(idx(it), sub);
            for (uint32_t x : sub) appendItemLeaves(x, depth + 1);
            return;
        }
        if (type(it) != GRT_REF_NONE) pushLeaf(it);
    }

    // the leaves of binary subtree b (remapped ABI refs that are not inner nodes of it), appended to `leaves`
    void gatherLeaves(uint32_t b0, size_t& lo, size_t& hi, bool append = false) {
        lo = leaves.size();
        std::vector<uint32_t> st{b0};
        while (!st.empty()) {
            const uint32_t b = st.back();
            st.pop_back();
            const uint32_t l = N[b].left & ~GRT_NODE_HINT_BIT, r = N[b].right & ~GRT_NODE_HINT_BIT;
            for (int k = 0; k < 2; k++) {
                const uint32_t c = k ? r : l;
                if (k && r == l) continue;   // a span-1 node names its object twice (bvh.go:44-46)
                if (type(c) == GRT_REF_NONE) continue;
                if (type(c) == GRT_REF_NODE && idx(c) < N.size()) { st.push_back(idx(c)); continue; }
                SahLeaf L;
                L.ref = c;
                L.box = refBox(c);
                if (L.box.empty()) continue;
                for (int a = 0; a < 3; a++) L.c[a] = 0.5 * (L.box.lo[a] + L.box.hi[a]);
                leaves.push_back(L);
            }
        }
        hi = leaves.size();
        if (!append && hi - lo < 2) { leaves.resize(lo); hi = lo = 0; }   // nothing to regroup: the plain collapse handles it
    }
    Box rangeBox(size_t lo, size_t hi) const {
        Box b = Box::none();
        for (size_t i = lo; i < hi; i++) b.add(leaves[i].box);
        return b;
    }
    // best SAH split of leaves[lo, hi) (reordered in place; lb / rb: the boxes of the two sides); falls back to the object
    // median on the longest axis
    size_t sahSplit(size_t lo, size_t hi, Box& lb, Box& rb) {
        constexpr int NB = 16;
        const size_t n = hi - lo;
        if ((int)n <= sahSweepMax) {
            // a small set: exact sweep over the sorted centroids, and only splits that leave whole groups of four on one
            // side (a 4-wide node with two leaves costs a full node visit for half the work)
            double bestCost = std::numeric_limits<double>::infinity();
            int bestAxis = -1;
            size_t bestK = 0;
            std::vector<SahLeaf> tmp(leaves.begin() + lo, leaves.begin() + hi), bestOrder;
            std::vector<double> ra(n + 1);
            for (int a = 0; a < 3; a++) {
                std::stable_sort(tmp.begin(), tmp.end(), [&](const SahLeaf& x, const SahLeaf& y) { return x.c[a] < y.c[a]; });
                Box acc = Box::none();
                for (size_t k = n; k-- > 1;) { acc.add(tmp[k].box); ra[k] = acc.area(); }
                acc = Box::none();
                for (size_t k = 1; k < n; k++) {
                    acc.add(tmp[k - 1].box);
                    if (n > 4 && (k % 4) != 0 && ((n - k) % 4) != 0) continue;
                    const double cost = acc.area() * (double)k + ra[k] * (double)(n - k);
                    if (std::isfinite(cost) && cost < bestCost) { bestCost = cost; bestAxis = a; bestK = k; bestOrder = tmp; }
                }
            }
            if (bestAxis >= 0) {
                std::copy(bestOrder.begin(), bestOrder.end(), leaves.begin() + lo);
                lb = rangeBox(lo, lo + bestK); rb = rangeBox(lo + bestK, hi);
                return lo + bestK;
            }
        }
        // binned: 16 bins per axis on the centroids, all three axes filled in one pass
        Box cb = Box::none();
        for (size_t i = lo; i < hi; i++) cb.add(leaves[i].c);
        double bestCost = std::numeric_limits<double>::infinity(), scale[3] = {0, 0, 0};
        int bestAxis = -1, bestBin = 0;
        if (cb.finite()) {
            Box bb[3][NB];
            size_t cnt[3][NB];
            for (int a = 0; a < 3; a++) {
                const double ext = cb.hi[a] - cb.lo[a];
                scale[a] = ext > 0 ? NB / ext : 0.0;
                for (int k = 0; k < NB; k++) { bb[a][k] = Box::none(); cnt[a][k] = 0; }
            }
            auto bin = [&](const SahLeaf& L, int a) { int k = (int)((L.c[a] - cb.lo[a]) * scale[a]); return k < 0 ? 0 : (k >= NB ? NB - 1 : k); };
            for (size_t i = lo; i < hi; i++)
                for (int a = 0; a < 3; a++) { const int k = bin(leaves[i], a); bb[a][k].add(leaves[i].box); cnt[a][k]++; }
            for (int a = 0; a < 3; a++) {
                if (!(scale[a] > 0)) continue;
                double ra[NB];
                size_t rc[NB];
                Box acc = Box::none();
                size_t c = 0;
                for (int k = NB - 1; k > 0; k--) { acc.add(bb[a][k]); c += cnt[a][k]; ra[k] = acc.area(); rc[k] = c; }
                acc = Box::none(); c = 0;
                for (int k = 1; k < NB; k++) {
                    acc.add(bb[a][k - 1]); c += cnt[a][k - 1];
                    if (c == 0 || rc[k] == 0) continue;
                    const double cost = acc.area() * (double)c + ra[k] * (double)rc[k];
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = k; }
                }
            }
            if (bestAxis >= 0) {
                const int a = bestAxis;
                auto mid = std::partition(leaves.begin() + lo, leaves.begin() + hi, [&](const SahLeaf& L) { return bin(L, a) < bestBin; });
                const size_t m = (size_t)(mid - leaves.begin());
                if (m > lo && m < hi) {
                    lb = Box::none(); rb = Box::none();
                    for (int k = 0; k < NB; k++) (k < bestBin ? lb : rb).add(bb[a][k]);
                    return m;
                }
            }
        }
        // coincident centroids (or non-finite boxes): object median on the longest axis of the boxes
        Box bx = rangeBox(lo, hi);
        int a = 0;
        double ext = -1;
        for (int k = 0; k < 3; k++) { const double e = bx.hi[k] - bx.lo[k]; if (std::isfinite(e) && e > ext) { ext = e; a = k; } }
        std::nth_element(leaves.begin() + lo, leaves.begin() + lo + n / 2, leaves.begin() + hi, [&](const SahLeaf& x, const SahLeaf& y) { return x.c[a] < y.c[a]; });
        lb = rangeBox(lo, lo + n / 2); rb = rangeBox(lo + n / 2, hi);
        return lo + n / 2;
    }

    // the items of the list starting at entry e0, as device refs of runs (long runs split) or remapped ABI refs
    void listItems(uint32_t e0, std::vector<uint32_t>& out) const {
        for (uint32_t k = e0;; k++) {
            const uint32_t e = E[2 * k], cnt = E[2 * k + 1], first = e & ~GRT_LIST_LAST;
            if (isPrim(type(first))) {
                for (uint32_t at = 0; at < cnt; at += GRT_DREF_RUN_MAX_COUNT) {
                    const uint32_t c = std::min(cnt - at, GRT_DREF_RUN_MAX_COUNT);
                    if (idx(first) + at + c - 1 > GRT_DREF_RUN_INDEX_MASK) { for (uint32_t j = 0; j < c; j++) out.push_back(first + at + j); continue; }
                    out.push_back(GRT_DREF_RUN_BIT | (type(first) << GRT_REF_SHIFT) | ((c - 1) << GRT_DREF_RUN_INDEX_BITS) | (idx(first) + at));
                }
            } else if (type(first) != GRT_REF_NONE) out.push_back(first);
            if (e & GRT_LIST_LAST) break;
        }
    }
    Box itemBox(uint32_t it) const {
        if (isSet(it)) return setBox.at(it & GRT_DREF_RUN_INDEX_MASK);
        if (it & GRT_DREF_RUN_BIT) {
            Box b = Box::none();
            const uint32_t t = (it >> GRT_REF_SHIFT) & 7u, cnt = ((it >> GRT_DREF_RUN_INDEX_BITS) & 7u) + 1u, first = it & GRT_DREF_RUN_INDEX_MASK;
            for (uint32_t j = 0; j < cnt; j++) b.add(primBox(t, first + j));
            return b;
        }
        if (type(it) == GRT_REF_NODE) { Box b = nodeBox(idx(it)); grow(b, 5e-7); return b; }
        return refBox(it);
    }

    static uint32_t type(uint32_t r) { return GRT_REF_TYPE(r); }
    static uint32_t idx(uint32_t r) { return r & GRT_REF_MASK; }
    static bool isPrim(uint32_t t) { return t == GRT_REF_SPHERE || t == GRT_REF_QUAD || t == GRT_REF_TRI || t == GRT_REF_BOX; }

    // outward rounding with a margin of a few ulps (the device slab test is fp32)
    static float down(double x) {
        if (!std::isfinite(x)) return (float)x;
        float f = (float)(x - 4e-7 * fabs(x) - 1e-30);
        if ((double)f > x) f = nextafterf(f, -std::numeric_limits<float>::infinity());
        return nextafterf(f, -std::numeric_limits<float>::infinity());
    }
    static float up(double x) {
        if (!std::isfinite(x)) return (float)x;
        float f = (float)(x + 4e-7 * fabs(x) + 1e-30);
        if ((double)f < x) f = nextafterf(f, std::numeric_limits<float>::infinity());
        return nextafterf(f, std::numeric_limits<float>::infinity());
    }

    // children of binary node b in the reference's order; a span-1 node names one object twice (bvh.go:44-46):
    // testing a surface twice cannot change the closest hit, a medium draws per test and is kept twice
    void pushChildren(std::vector<uint32_t>& out, uint32_t b) {
        const uint32_t l = N[b].left & ~GRT_NODE_HINT_BIT, r = N[b].right & ~GRT_NODE_HINT_BIT;
        out.push_back(l);
        if (r != l || subtreeHasMedium(l)) out.push_back(r);
    }

    Box nodeBox(uint32_t b) const {
        Box x;
        for (int a = 0; a < 3; a++) { x.lo[a] = N[b].bmin[a]; x.hi[a] = N[b].bmax[a]; }
        return x;
    }

    bool subtreeHasMedium(uint32_t r) {
        if (r & GRT_DREF_RUN_BIT) return false;
        switch (type(r)) {
            case GRT_REF_MEDIUM: return true;
            case GRT_REF_NODE: {
                // iterative post-order so a million-node tree does not recurse a million deep
                const uint32_t b0 = idx(r);
                if (hasMedium[b0] >= 0) return hasMedium[b0] != 0;
                std::vector<uint32_t> st{b0};
                while (!st.empty()) {
                    const uint32_t b = st.back();
                    if (hasMedium[b] >= 0) { st.pop_back(); continue; }
                    const uint32_t c[2] = {N[b].left & ~GRT_NODE_HINT_BIT, N[b].right & ~GRT_NODE_HINT_BIT};
                    bool ready = true, any = false;
                    for (int k = 0; k < 2; k++) {
                        if (type(c[k]) == GRT_REF_NODE) {
                            if (hasMedium[idx(c[k])] < 0) { st.push_back(idx(c[k])); ready = false; }
                            else any = any || hasMedium[idx(c[k])] != 0;
                        } else any = any || subtreeHasMedium(c[k]);
                    }
                    if (ready) { hasMedium[b] = any ? 1 : 0; st.pop_back(); }
                }
                return hasMedium[b0] != 0;
            }
            case GRT_REF_LIST: {
                for (uint32_t k = idx(r);; k++) {
                    const uint32_t e = E[2 * k];
                    if (subtreeHasMedium(e & ~GRT_LIST_LAST)) return true;
                    if (e & GRT_LIST_LAST) break;
                }
                return false;
            }
        }
        return false;
    }

    // ---- boxes of things that are not inner nodes -------------------------------------------------------------
    Box primBox(uint32_t t, uint32_t i) const {
        Box b = Box::none();
        if (t == GRT_REF_SPHERE) {   // objects.go:23-37
            const GrtSphere& s = S->spheres[i];
            for (int k = 0; k < 2; k++) {
                double lo[3], hi[3];
                for (int a = 0; a < 3; a++) { double c = s.c0[a] + (double)k * (double)s.dc[a]; lo[a] = c - fabs(s.r); hi[a] = c + fabs(s.r); }
                b.add(lo); b.add(hi);
            }
            grow(b, 1e-6);
        } else if (t == GRT_REF_TRI) {   // objects.go:317-354
            const GrtTri& tr = S->tris[i];
            double v0[3], v1[3], v2[3];
            for (int a = 0; a < 3; a++) { v0[a] = tr.v0[a]; v1[a] = (double)tr.v0[a] + tr.e0[a]; v2[a] = (double)tr.v0[a] + tr.e1[a]; }
            b.add(v0); b.add(v1); b.add(v2);
            if (S->tri_v64) { const double* v = S->tri_v64 + 9 * (size_t)i; b.add(v); b.add(v + 3); b.add(v + 6); }
            grow(b, 1e-6);
        } else if (t == GRT_REF_QUAD) {   // objects.go:129-147; u, v recovered from alpha = A.(p-Q), beta = B.(p-Q)
            const GrtQuad& q = S->quads[i];
            const double n[3] = {q.n[0], q.n[1], q.n[2]}, A[3] = {q.A[0], q.A[1], q.A[2]}, B[3] = {q.B[0], q.B[1], q.B[2]};
            double bn[3], na[3];
            cross(B, n, bn); cross(n, A, na);
            const double su = dot(A, bn), sv = dot(B, na);
            if (!(fabs(su) > 1e-300) || !(fabs(sv) > 1e-300) || !std::isfinite(su) || !std::isfinite(sv)) return Box::all();
            double p[4][3];
            for (int a = 0; a < 3; a++) {
                const double u = bn[a] / su, v = na[a] / sv;
                p[0][a] = q.Q[a]; p[1][a] = q.Q[a] + u; p[2][a] = q.Q[a] + v; p[3][a] = q.Q[a] + u + v;
            }
            for (int k = 0; k < 4; k++) b.add(p[k]);
            grow(b, 1e-4);
        } else if (t == GRT_REF_BOX) {   // objects.go:208-240 under the baked rotateY + translate
            const GrtBox& x = S->boxes[i];
            for (int k = 0; k < 8; k++) {
                const double ox = (k & 1) ? x.mx[0] : x.mn[0], oy = (k & 2) ? x.mx[1] : x.mn[1], oz = (k & 4) ? x.mx[2] : x.mn[2];
                const double p[3] = {(double)x.rc * ox + (double)x.rs * oz + x.T[0], oy + x.T[1], -(double)x.rs * ox + (double)x.rc * oz + x.T[2]};
                b.add(p);
            }
            grow(b, 1e-5);
        }
        return b;
    }
    static void grow(Box& b, double rel) {
        if (b.empty()) return;
        for (int a = 0; a < 3; a++) {
            const double m = rel * (fabs(b.lo[a]) + fabs(b.hi[a]) + (b.hi[a] - b.lo[a])) + 1e-7;
            b.lo[a] -= m; b.hi[a] += m;
        }
    }
    static void cross(const double a[3], const double b[3], double o[3]) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
    static double dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

    Box refBox(uint32_t r, int depth = 0) const {
        if (depth > 16) return Box::all();
        const uint32_t t = type(r), i = idx(r);
        if (isPrim(t)) return primBox(t, i);
        if (t == GRT_REF_NODE) return nodeBox(i);
        if (t == GRT_REF_MEDIUM) return refBox(M[i].boundary, depth + 1);
        if (t == GRT_REF_LIST) {
            Box b = Box::none();
            for (uint32_t k = i;; k++) {
                const uint32_t e = E[2 * k], cnt = E[2 * k + 1], first = e & ~GRT_LIST_LAST;
                if (isPrim(type(first))) for (uint32_t j = 0; j < cnt; j++) b.add(primBox(type(first), idx(first) + j));
                else b.add(refBox(first, depth + 1));
                if (e & GRT_LIST_LAST) break;
            }
            return b;
        }
        return Box::none();
    }

    // a child that is one short run of primitives is named directly (no list-entry fetch on the device)
    uint32_t asRun(uint32_t r) const {
        uint32_t first = r, cnt = 1;
        if (type(r) == GRT_REF_LIST) {
            const uint32_t e = E[2 * idx(r)];
            if (!(e & GRT_LIST_LAST)) return r;
            first = e & ~GRT_LIST_LAST; cnt = E[2 * idx(r) + 1];
        }
        if (!isPrim(type(first)) || cnt < 1 || cnt > GRT_DREF_RUN_MAX_COUNT || idx(first) + cnt - 1 > GRT_DREF_RUN_INDEX_MASK) return r;
        return GRT_DREF_RUN_BIT | (type(first) << GRT_REF_SHIFT) | ((cnt - 1) << GRT_DREF_RUN_INDEX_BITS) | idx(first);
    }

    // ---- traversal stack need (entries held at once, worst case over visiting orders) ------------------------------
    int needRef(uint32_t r, int depth = 0) {
        if (depth > 64) { Rp->error = "scene nesting too deep"; return 1 << 20; }
        if (r & GRT_DREF_RUN_BIT) return 0;
        switch (type(r)) {
            case GRT_REF_NODE: return needWide(idx(r), depth);
            case GRT_REF_LIST: {
                int need = 0;
                for (uint32_t k = idx(r);; k++) {
                    const uint32_t e = E[2 * k];
                    need = std::max(need, 1 + needRef(wideRef(e & ~GRT_LIST_LAST), depth + 1));   // the rest of the list waits on the stack
                    if (e & GRT_LIST_LAST) break;
                }
                return need;
            }
        }
        return 0;   // primitives; a medium's boundary query runs on its own stack (need_boundary)
    }
    int needWide(uint32_t w0, int depth) {
        if (needW[w0] >= 0) return needW[w0];
        // iterative post-order over wide nodes
        std::vector<uint32_t> st{w0};
        while (!st.empty()) {
            const uint32_t w = st.back();
            if (needW[w] >= 0) { st.pop_back(); continue; }
            const float* d = Rp->wnodes.data() + (size_t)w * GRT_WNODE_FLOATS;
            uint32_t refs[4], meta[4];
            memcpy(refs, d + 24, 16); memcpy(meta, d + 28, 16);
            bool ready = true;
            int mx = 0;
            for (uint32_t k = 0; k < meta[1]; k++) {
                const uint32_t c = refs[k];
                if (!(c & GRT_DREF_RUN_BIT) && type(c) == GRT_REF_NODE) {
                    if (needW[idx(c)] < 0) { st.push_back(idx(c)); ready = false; }
                    else mx = std::max(mx, needW[idx(c)]);
                } else mx = std::max(mx, needRef(c, depth + 1));
            }
            if (ready) { needW[w] = (int)meta[1] - 1 + mx; if (needW[w] < 0) needW[w] = 0; st.pop_back(); }
        }
        return needW[w0];
    }
};

}  // namespace wide
}  // namespace grt
