// flatten.hpp — turns the host-built scene (scene_ir.hpp, the reference's
// constructor vocabulary) into the flat arrays of include/grt.h.
//
// What happens here, all in fp64 on the host, once per scene:
//   * BuildBVH: the reference's median split (bvh.go:35-61) — bbox union,
//     LongestAxis (aabb.go:73-87), sort by boxCompare (bvh.go:25-32), split at
//     span/2, span 1 duplicates the object, span 2 keeps list order — run in
//     OBJECT space with the reference's bbox rules (aabb.go:25-59,118-129,
//     objects.go:23-37,143-147,317-354, transformation.go:21-24,48-77) so the
//     tree TOPOLOGY and child order are the reference's.
//   * translate / rotateY instances (transformation.go) are baked: primitives
//     are moved to world space, node boxes are re-fitted bottom-up in world
//     space and rounded OUTWARD to fp32.
//   * quads get the reference's derived fields (objects.go:129-141) plus the
//     precomputed A = v×w, B = w×u used by the device interior test.
// This is the Go-side `hittable.Flatten` of INTEGRATION.md written in C++.
#pragma once
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <functional>
#include "scene_ir.hpp"
#include "../../include/grt.h"
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <map>

namespace grt {
namespace flat {

using ir::V3;
static const double kInf = std::numeric_limits<double>::infinity();

inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline double length(V3 a) { return std::sqrt(dot(a, a)); }

// ---- reference bbox arithmetic (aabb.go, interval.go) ----------------------
struct RefBox {
    double lo[3], hi[3];
};
inline void padToMinimum(RefBox& b) {  // aabb.go:118-129, interval.go:47-50
    const double delta = 0.0001;
    for (int a = 0; a < 3; a++)
        if (b.hi[a] - b.lo[a] < delta) { double p = delta / 2; b.lo[a] -= p; b.hi[a] += p; }
}
inline RefBox emptyBox() { RefBox b; for (int a = 0; a < 3; a++) { b.lo[a] = kInf; b.hi[a] = -kInf; } padToMinimum(b); return b; }  // aabb.go:20
inline RefBox fromPoints(V3 p, V3 q) {  // aabb.go:31-52
    RefBox b;
    double P[3] = {p.x, p.y, p.z}, Q[3] = {q.x, q.y, q.z};
    for (int a = 0; a < 3; a++) { if (P[a] < Q[a]) { b.lo[a] = P[a]; b.hi[a] = Q[a]; } else { b.lo[a] = Q[a]; b.hi[a] = P[a]; } }
    padToMinimum(b);
    return b;
}
inline RefBox fromBoxes(const RefBox& p, const RefBox& q) {  // aabb.go:54-59
    RefBox b;
    for (int a = 0; a < 3; a++) { b.lo[a] = std::fmin(p.lo[a], q.lo[a]); b.hi[a] = std::fmax(p.hi[a], q.hi[a]); }
    padToMinimum(b);
    return b;
}
inline int longestAxis(const RefBox& b) {  // aabb.go:73-87
    double sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
    if (sx > sy) return sx > sz ? 0 : 2;
    return sy > sz ? 1 : 2;
}

// A rigid Y-rotation followed by a translation: world = R(obj) + T with
// R(v) = (c v.x + s v.z, v.y, -s v.x + c v.z)   (transformation.go:87-93).
struct Xform {
    double c = 1, s = 0;
    V3 T;
    V3 rot(V3 v) const { return V3(c * v.x + s * v.z, v.y, -s * v.x + c * v.z); }
    V3 point(V3 p) const { return rot(p) + T; }
    bool identity() const { return c == 1 && s == 0 && T.x == 0 && T.y == 0 && T.z == 0; }
};

struct FlatScene {
    std::vector<GrtNode> nodes;
    std::vector<GrtSphere> spheres;
    std::vector<GrtQuad> quads;
    std::vector<GrtBox> boxes;
    std::vector<GrtTri> tris;
    std::vector<GrtTriShade> tri_shade;
    std::vector<double> tri_v64;
    std::vector<uint32_t> items;
    std::vector<GrtMedium> media;
    std::vector<GrtMaterial> materials;
    std::vector<GrtTexture> textures;
    std::vector<GrtImage> images;
    std::vector<uint8_t> texels;
    std::vector<GrtPerlin> perlins;
    std::vector<GrtLight> lights;
    uint32_t root = 0, lights_mode = GRT_LIGHTS_LIST, max_depth_hint = 0;
    bool any_tri_shade = false;

    GrtScene view() const {
        GrtScene s;
        memset(&s, 0, sizeof(s));
        s.abi_version = GRT_ABI_VERSION;
        s.root = root;
        s.nodes = nodes.data(); s.n_nodes = (uint32_t)nodes.size();
        s.spheres = spheres.data(); s.n_spheres = (uint32_t)spheres.size();
        s.quads = quads.data(); s.n_quads = (uint32_t)quads.size();
        s.boxes = boxes.data(); s.n_boxes = (uint32_t)boxes.size();
        s.tris = tris.data(); s.n_tris = (uint32_t)tris.size();
        s.tri_shade = any_tri_shade ? tri_shade.data() : nullptr;
        s.tri_v64 = tri_v64.empty() ? nullptr : tri_v64.data();
        s.items = items.data(); s.n_items = (uint32_t)items.size();
        s.media = media.data(); s.n_media = (uint32_t)media.size();
        s.materials = materials.data(); s.n_materials = (uint32_t)materials.size();
        s.textures = textures.data(); s.n_textures = (uint32_t)textures.size();
        s.images = images.data(); s.n_images = (uint32_t)images.size();
        s.texels = texels.data(); s.n_texel_bytes = texels.size();
        s.perlins = perlins.data(); s.n_perlins = (uint32_t)perlins.size();
        s.lights = lights.data(); s.n_lights = (uint32_t)lights.size();
        s.lights_mode = lights_mode;
        s.max_depth_hint = max_depth_hint;
        return s;
    }
};

// Flatten-time tuning (results are identical for every setting; only the amount
// of box culling changes).  A BVH subtree with few leaves costs more to
// traverse on SIMT hardware (divergent box/primitive alternation) than to test
// linearly, so such subtrees are emitted as a HittableList-style ordered run of
// their leaves in the reference's depth-first, left-first visiting order.
struct FlattenOptions {
    int collapse_whole = 32;   // a BuildBVH result with at most this many leaves becomes one list
    int collapse_leaf = 1;     // inside larger trees, subtrees with at most this many leaves become lists (1: every primitive is its own leaf
                               // and gets its own box in the 4-wide device BVH: measured best of 1/2/4/8, profiles/README.md round 2)
    bool box_prims = true;     // NewBox results are found with one slab test (GrtBox) instead of six quad tests
    bool order_hints = true;   // nodes carry the split axis so a ray may visit the nearer child first
    // BuildBVH's object order computed elsewhere (the GPU, grt_bvh_order) for lists of at least gpu_order_min
    // objects: boxes = n x {lo.xyz, hi.xyz}, order[p] = list index of the object at position p.  Returns false
    // to make the flattener sort on the host.
    std::function<bool(const double* boxes, uint32_t n, uint32_t* order)> gpu_order;
    size_t gpu_order_min = 32768;
};

class Flattener {
   public:
    explicit Flattener(const ir::Scene& s, FlattenOptions o = FlattenOptions()) : S(s), opt(o) {}
    std::string error;

    bool run(FlatScene& out) {
        F = &out;
        try {
            if (S.world < 0 || S.lights < 0) throw std::runtime_error("scene has no world or no lights (Camera.Render needs both)");
            refBoxCache.assign(S.hittables.size(), RefBox());
            refBoxDone.assign(S.hittables.size(), 0);
            emitMaterials();
            const double t_emit = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
            Emitted e = emit(S.world, Xform(), false);
            if (getenv("GRT_FLATTEN_TRACE")) fprintf(stderr, "[flatten] emit (incl. BuildBVH) %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t_emit);
            F->root = e.ref;
            F->max_depth_hint = (uint32_t)std::max(e.need + 1, mediumNeed + 1);
            emitLights();
        } catch (const std::exception& ex) {
            error = ex.what();
            return false;
        }
        return true;
    }

   private:
    const ir::Scene& S;
    FlattenOptions opt;
    FlatScene* F = nullptr;
    std::vector<RefBox> refBoxCache;
    std::vector<char> refBoxDone;
    int mediumNeed = 0;

    // ---- BuildBVH topology in object space (bvh.go:35-61) ---------------
    struct BuildNode { int left, right; bool leftIsNode, rightIsNode; int leaves; int axis; };  // child: build-node index or hittable id
    struct Topology { std::vector<BuildNode> nodes; int root = -1; };

    bool isBoxPrim(int listPayload) const {
        if (!opt.box_prims || S.box_of_list[listPayload] < 0) return false;
        const ir::BoxP& b = S.boxes[S.box_of_list[listPayload]];
        return b.mx.x > b.mn.x && b.mx.y > b.mn.y && b.mx.z > b.mn.z;   // a flat "box" keeps its six quads
    }
    // number of primitive tests a linear run over this hittable would make (-1: not collapsible)
    int leafCount(int hid) {
        const ir::Hittable& h = S.hittables[hid];
        switch (h.type) {
            case ir::H_SPHERE: case ir::H_QUAD: case ir::H_TRI: return 1;
            case ir::H_MEDIUM: return 1;   // stays a single ref inside a run
            case ir::H_TRANSLATE: case ir::H_ROTATEY: return leafCount(h.child);
            case ir::H_LIST: { int n = 0; for (int c : S.lists[h.a]) { int k = leafCount(c); n += k; if (n > (1 << 20)) return 1 << 20; } return n; }
            case ir::H_BVH: if (isBoxPrim(h.a)) return 1;
                return topology(h.a).nodes[topology(h.a).root].leaves;
        }
        return 1;
    }
    std::map<int, Topology> topoCache;  // by list payload index

    // Reference bbox of a hittable in ITS OWN space (what obj.BBox() returns in Go).
    const RefBox& refBox(int hid) {
        if (refBoxDone[hid]) return refBoxCache[hid];
        const ir::Hittable& h = S.hittables[hid];
        RefBox b;
        switch (h.type) {
            case ir::H_SPHERE: {  // objects.go:23-37
                const ir::SphereP& p = S.spheres[h.a];
                V3 rv(p.r, p.r, p.r);
                if (p.dc.x == 0 && p.dc.y == 0 && p.dc.z == 0) b = fromPoints(p.c0 - rv, p.c0 + rv);
                else {
                    V3 c0 = p.c0 + p.dc * 0.0, c1 = p.c0 + p.dc * 1.0;
                    b = fromBoxes(fromPoints(c0 - rv, c0 + rv), fromPoints(c1 - rv, c1 + rv));
                }
                break;
            }
            case ir::H_QUAD: {  // objects.go:143-147
                const ir::QuadP& q = S.quads[h.a];
                b = fromBoxes(fromPoints(q.Q, q.Q + q.u + q.v), fromPoints(q.Q + q.u, q.Q + q.v));
                break;
            }
            case ir::H_TRI: {  // objects.go:317-354
                const ir::TriP& t = S.tris[h.a];
                double mn[3] = {kInf, kInf, kInf}, mx[3] = {-kInf, -kInf, -kInf};
                for (int k = 0; k < 3; k++) {
                    double v[3] = {t.v[k].x, t.v[k].y, t.v[k].z};
                    for (int a = 0; a < 3; a++) { mn[a] = std::fmin(v[a], mn[a]); mx[a] = std::fmax(v[a], mx[a]); }
                }
                const double epsilon = 1e-8;
                for (int a = 0; a < 3; a++) if (mx[a] - mn[a] < epsilon) { mx[a] += epsilon; mn[a] -= epsilon; }
                for (int a = 0; a < 3; a++) { b.lo[a] = mn[a]; b.hi[a] = mx[a]; }
                padToMinimum(b);
                break;
            }
            case ir::H_LIST: {  // hittable.go:105-116
                b = emptyBox();
                for (int c : S.lists[h.a]) b = fromBoxes(b, refBox(c));
                break;
            }
            case ir::H_BVH: {  // bvh.go:36-39 (root bbox = running union over the list)
                b = emptyBox();
                for (int c : S.lists[h.a]) b = fromBoxes(b, refBox(c));
                break;
            }
            case ir::H_TRANSLATE: {  // transformation.go:21-24, aabb.go:131
                const RefBox& cb = refBox(h.child);
                V3 o = S.xforms[h.a].offset;
                double off[3] = {o.x, o.y, o.z};
                for (int a = 0; a < 3; a++) { b.lo[a] = cb.lo[a] + off[a]; b.hi[a] = cb.hi[a] + off[a]; }
                padToMinimum(b);
                break;
            }
            case ir::H_ROTATEY: {  // transformation.go:48-77
                const RefBox& cb = refBox(h.child);
                double radians = S.xforms[h.a].degrees * 3.14159265358979323846 / 180.0;
                double sn = std::sin(radians), cs = std::cos(radians);
                double mn[3] = {kInf, kInf, kInf}, mx[3] = {-kInf, -kInf, -kInf};
                for (int i = 0; i < 2; i++)
                    for (int j = 0; j < 2; j++)
                        for (int k = 0; k < 2; k++) {
                            double x = (double)i * cb.hi[0] + (double)(1 - i) * cb.lo[0];
                            double y = (double)j * cb.hi[1] + (double)(1 - j) * cb.lo[1];
                            double z = (double)k * cb.hi[2] + (double)(1 - k) * cb.lo[2];
                            double t[3] = {cs * x + sn * z, y, -sn * x + cs * z};
                            for (int c = 0; c < 3; c++) { mn[c] = std::fmin(mn[c], t[c]); mx[c] = std::fmax(mx[c], t[c]); }
                        }
                b = fromPoints(V3(mn[0], mn[1], mn[2]), V3(mx[0], mx[1], mx[2]));
                break;
            }
            case ir::H_MEDIUM: b = refBox(h.child); break;  // medium.go:60-62
            default: throw std::runtime_error("unknown hittable type");
        }
        refBoxCache[hid] = b;
        refBoxDone[hid] = 1;
        return refBoxCache[hid];
    }

    // `presorted`: objs already stands in BuildBVH's final order (grt_bvh_order), so nothing is sorted here and a span's
    // box — needed only for its longest axis — is the union of its halves' boxes instead of a pass over all its objects.
    int buildRange(Topology& T, std::vector<int>& objs, size_t start, size_t end, bool presorted = false, RefBox* box_out = nullptr) {
        size_t span = end - start;
        BuildNode n;
        n.axis = -1;   // only a sorted split (span >= 3) orders its children along `axis`
        RefBox bb = emptyBox();
        if (!presorted || span <= 2) for (size_t i = start; i < end; i++) bb = fromBoxes(bb, refBox(objs[i]));
        if (span == 1) { n.left = n.right = objs[start]; n.leftIsNode = n.rightIsNode = false; n.leaves = leafCount(objs[start]); }
        else if (span == 2) { n.left = objs[start]; n.right = objs[start + 1]; n.leftIsNode = n.rightIsNode = false; n.leaves = leafCount(objs[start]) + leafCount(objs[start + 1]); }
        else {
            size_t mid = start + span / 2;
            if (!presorted) {
                int axis = longestAxis(bb);
                // boxCompare (bvh.go:25-32).  Go's sort.Slice is not stable; equal keys are
                // documented as unordered, we keep list order for them.
                std::stable_sort(objs.begin() + start, objs.begin() + end, [&](int a, int b) {
                    const RefBox &A = refBoxCache[a], &B = refBoxCache[b];
                    if (A.lo[axis] != B.lo[axis]) return A.lo[axis] < B.lo[axis];
                    return A.hi[axis] < B.hi[axis];
                });
                n.left = buildRange(T, objs, start, mid);
                n.right = buildRange(T, objs, mid, end);
                n.axis = axis;
            } else {
                RefBox lb, rb;
                n.left = buildRange(T, objs, start, mid, true, &lb);
                n.right = buildRange(T, objs, mid, end, true, &rb);
                bb = fromBoxes(lb, rb);
                n.axis = longestAxis(bb);
            }
            n.leftIsNode = n.rightIsNode = true;
            n.leaves = T.nodes[n.left].leaves + T.nodes[n.right].leaves;
        }
        if (box_out) *box_out = bb;
        T.nodes.push_back(n);
        return (int)T.nodes.size() - 1;
    }
    const Topology& topology(int listPayload) {
        auto it = topoCache.find(listPayload);
        if (it != topoCache.end()) return it->second;
        Topology T;
        std::vector<int> objs = S.lists[listPayload];
        if (objs.empty()) throw std::runtime_error("BuildBVH of an empty list (the reference would index out of range)");
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const bool trace = getenv("GRT_FLATTEN_TRACE") != nullptr && objs.size() > 1000;
        double t0 = now();
        for (int o : objs) (void)refBox(o);
        double t1 = now();
        bool presorted = false;
        if (opt.gpu_order && objs.size() >= opt.gpu_order_min) {
            // the final order of all recursive sorts, computed level by level on the device (grt_bvh.cu)
            std::vector<double> boxes(objs.size() * 6);
            for (size_t i = 0; i < objs.size(); i++) {
                const RefBox& b = refBoxCache[objs[i]];
                for (int a = 0; a < 3; a++) { boxes[6 * i + a] = b.lo[a]; boxes[6 * i + 3 + a] = b.hi[a]; }
            }
            std::vector<uint32_t> order(objs.size());
            if (opt.gpu_order(boxes.data(), (uint32_t)objs.size(), order.data())) {
                std::vector<int> sorted(objs.size());
                for (size_t p = 0; p < objs.size(); p++) sorted[p] = objs[order[p]];
                objs.swap(sorted);
                presorted = true;
            }
        }
        double t2 = now();
        T.root = buildRange(T, objs, 0, objs.size(), presorted);
        if (trace) fprintf(stderr, "[flatten] BuildBVH of %zu objects: boxes %.3f s, %s order %.3f s, tree %.3f s\n", objs.size(), t1 - t0,
                           presorted ? "GPU" : "no separate", t2 - t1, now() - t2);
        return topoCache.emplace(listPayload, std::move(T)).first->second;
    }

    // ---- world-space emission ---------------------------------------------
    struct WBox {
        double lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
        void add(V3 p) { double v[3] = {p.x, p.y, p.z}; for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], v[a]); hi[a] = std::fmax(hi[a], v[a]); } }
        void add(const WBox& b) { for (int a = 0; a < 3; a++) { lo[a] = std::fmin(lo[a], b.lo[a]); hi[a] = std::fmax(hi[a], b.hi[a]); } }
    };
    struct Emitted { uint32_t ref; WBox box; int need; bool spliceable = false; uint32_t list_first = 0, list_count = 0; bool has_medium = false; };

    static float roundDown(double x) {
        if (x == -kInf) return -std::numeric_limits<float>::infinity();
        float f = (float)x;
        if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
        return std::nextafterf(f, -std::numeric_limits<float>::infinity());
    }
    static float roundUp(double x) {
        if (x == kInf) return std::numeric_limits<float>::infinity();
        float f = (float)x;
        if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
        return std::nextafterf(f, std::numeric_limits<float>::infinity());
    }

    // Leaves of a subtree in the order BVHNode.Hit visits them (left first).  A span-1 node
    // names the same object twice (bvh.go:44-46); testing a surface primitive twice cannot
    // change the closest hit, so it is listed once — a constantMedium draws a fresh random
    // number per test (medium.go:47) and is therefore kept twice.
    void collectLeaves(const Topology& T, int ni, std::vector<int>& out) {
        const BuildNode& bn = T.nodes[ni];
        if (bn.leftIsNode) collectLeaves(T, bn.left, out); else out.push_back(bn.left);
        if (bn.rightIsNode) collectLeaves(T, bn.right, out);
        else if (bn.leftIsNode || bn.right != bn.left || S.hittables[bn.right].type == ir::H_MEDIUM) out.push_back(bn.right);
    }
    // Emits an ordered run of hittables as one list; nested lists are spliced in place
    // (HittableList.Hit of a nested list is the same sequential scan, hittable.go:122-138).
    Emitted emitRun(const std::vector<int>& kids, const Xform& X, bool inBoundary) {
        Emitted e;
        e.need = 0;
        std::vector<Emitted> ch;
        for (int c : kids) ch.push_back(emit(c, X, inBoundary));
        std::vector<uint32_t> refs;
        int need = 0;
        for (size_t i = 0; i < ch.size(); i++) {
            e.box.add(ch[i].box);
            e.has_medium = e.has_medium || ch[i].has_medium;
            uint32_t r = ch[i].ref;
            if (GRT_REF_TYPE(r) == GRT_REF_NONE) continue;
            if (GRT_REF_TYPE(r) == GRT_REF_LIST && ch[i].spliceable) {
                // splice: the child's items were appended most recently and contain only leaf refs
                for (uint32_t k = ch[i].list_first; k < ch[i].list_first + ch[i].list_count; k++) refs.push_back(F->items[k] & ~GRT_LIST_LAST);
                continue;
            }
            refs.push_back(r);
            need = std::max(need, 1 + ch[i].need);
        }
        if (refs.empty()) { e.ref = GRT_MAKE_REF(GRT_REF_NONE, 0); return e; }
        uint32_t first = (uint32_t)F->items.size();
        bool leafOnly = true;
        for (size_t i = 0; i < refs.size(); i++) {
            uint32_t t = GRT_REF_TYPE(refs[i]);
            if (t != GRT_REF_SPHERE && t != GRT_REF_QUAD && t != GRT_REF_TRI && t != GRT_REF_BOX) leafOnly = false;
            F->items.push_back(refs[i] | (i + 1 == refs.size() ? GRT_LIST_LAST : 0u));
        }
        e.ref = GRT_MAKE_REF(GRT_REF_LIST, first);
        e.need = std::max(need, 2);
        e.spliceable = leafOnly;
        e.list_first = first;
        e.list_count = (uint32_t)refs.size();
        return e;
    }

    Emitted emitTopo(const Topology& T, int ni, const Xform& X, bool inBoundary) {
        const BuildNode& bn = T.nodes[ni];
        int limit = (T.nodes[T.root].leaves <= opt.collapse_whole) ? opt.collapse_whole : opt.collapse_leaf;
        if (bn.leaves <= limit && bn.leaves > 0) {
            std::vector<int> leaves;
            collectLeaves(T, ni, leaves);
            return emitRun(leaves, X, inBoundary);
        }
        uint32_t idx = (uint32_t)F->nodes.size();
        if (idx > GRT_REF_MASK) throw std::runtime_error("too many BVH nodes");
        F->nodes.emplace_back();  // depth-first, left-first order
        Emitted l = bn.leftIsNode ? emitTopo(T, bn.left, X, inBoundary) : emit(bn.left, X, inBoundary);
        Emitted r;
        if (!bn.rightIsNode && !bn.leftIsNode && bn.right == bn.left) r = l;  // span 1: same object twice (bvh.go:44-46)
        else r = bn.rightIsNode ? emitTopo(T, bn.right, X, inBoundary) : emit(bn.right, X, inBoundary);
        WBox wb = l.box;
        wb.add(r.box);
        // minimum thickness like aabb.go:118-129 so no slab is degenerate, then outward rounding
        for (int a = 0; a < 3; a++) if (wb.hi[a] - wb.lo[a] < 0.0001) { wb.lo[a] -= 0.00005; wb.hi[a] += 0.00005; }
        GrtNode& n = F->nodes[idx];
        for (int a = 0; a < 3; a++) { n.bmin[a] = roundDown(wb.lo[a]); n.bmax[a] = roundUp(wb.hi[a]); }
        n.left = l.ref;
        n.right = r.ref;
        Emitted e;
        e.has_medium = l.has_medium || r.has_medium;
        // traversal hint (include/grt.h): only for an unrotated split with no medium below it
        if (opt.order_hints && bn.axis >= 0 && !e.has_medium && X.c == 1 && X.s == 0) {
            uint32_t h = (uint32_t)bn.axis + 1u;
            if (h & 1u) n.left |= GRT_NODE_HINT_BIT;
            if (h & 2u) n.right |= GRT_NODE_HINT_BIT;
        }
        e.ref = GRT_MAKE_REF(GRT_REF_NODE, idx);
        e.box = wb;
        e.need = std::max(2, std::max(1 + l.need, r.need));
        return e;
    }

    Emitted emit(int hid, const Xform& X, bool inBoundary) {
        const ir::Hittable& h = S.hittables[hid];
        Emitted e;
        e.need = 0;
        switch (h.type) {
            case ir::H_SPHERE: {
                const ir::SphereP& p = S.spheres[h.a];
                GrtSphere g;
                memset(&g, 0, sizeof(g));
                V3 c0 = X.point(p.c0), dc = X.rot(p.dc);
                g.c0[0] = c0.x; g.c0[1] = c0.y; g.c0[2] = c0.z; g.r = p.r;
                g.dc[0] = (float)dc.x; g.dc[1] = (float)dc.y; g.dc[2] = (float)dc.z;
                g.mat = (uint32_t)h.mat; g.id = (uint32_t)hid;
                g.uvrot[0] = (float)X.c; g.uvrot[1] = (float)X.s;
                F->spheres.push_back(g);
                e.ref = GRT_MAKE_REF(GRT_REF_SPHERE, F->spheres.size() - 1);
                V3 rv(p.r, p.r, p.r);
                e.box.add(c0 - rv); e.box.add(c0 + rv);
                if (!(dc.x == 0 && dc.y == 0 && dc.z == 0)) { e.box.add(c0 + dc - rv); e.box.add(c0 + dc + rv); }
                break;
            }
            case ir::H_QUAD: {
                const ir::QuadP& p = S.quads[h.a];
                V3 Q = X.point(p.Q), u = X.rot(p.u), v = X.rot(p.v);
                // NewQuad (objects.go:129-141)
                V3 n = cross(u, v);
                double nn = dot(n, n);
                if (!(nn > 0)) {
                    // u x v = 0: NewQuad's normal, D and w are NaN (objects.go:132-136), every comparison in
                    // quad.Hit is false and the quad can never be hit — emit nothing, keep its extent
                    e.ref = GRT_MAKE_REF(GRT_REF_NONE, 0);
                    e.box.add(Q); e.box.add(Q + u); e.box.add(Q + v); e.box.add(Q + u + v);
                    break;
                }
                V3 normal = n * (1 / length(n));
                double D = dot(normal, Q);
                V3 w = n * (1 / nn);
                V3 A = cross(v, w), B = cross(w, u);
                GrtQuad g;
                memset(&g, 0, sizeof(g));
                g.n[0] = (float)normal.x; g.n[1] = (float)normal.y; g.n[2] = (float)normal.z; g.D = (float)D;
                g.Q[0] = (float)Q.x; g.Q[1] = (float)Q.y; g.Q[2] = (float)Q.z;
                g.A[0] = (float)A.x; g.A[1] = (float)A.y; g.A[2] = (float)A.z;
                g.B[0] = (float)B.x; g.B[1] = (float)B.y; g.B[2] = (float)B.z;
                g.n64[0] = normal.x; g.n64[1] = normal.y; g.n64[2] = normal.z; g.D64 = D;
                int nz = (normal.x != 0) + (normal.y != 0) + (normal.z != 0);
                // axis-aligned AND fp32-exact plane: then D - n.o and n.d are single-rounding in fp32
                bool exactQ = (double)g.Q[0] == Q.x && (double)g.Q[1] == Q.y && (double)g.Q[2] == Q.z;
                if (nz == 1 && exactQ) {
                    g.flags |= GRT_QUAD_AXIS_ALIGNED;
                    for (int a = 0; a < 3; a++) g.n[a] = g.n[a] > 0 ? 1.0f : (g.n[a] < 0 ? -1.0f : 0.0f);
                    g.D = (float)(g.n[0] * Q.x + g.n[1] * Q.y + g.n[2] * Q.z);
                }
                g.mat = (uint32_t)h.mat; g.id = (uint32_t)hid;
                F->quads.push_back(g);
                e.ref = GRT_MAKE_REF(GRT_REF_QUAD, F->quads.size() - 1);
                e.box.add(Q); e.box.add(Q + u); e.box.add(Q + v); e.box.add(Q + u + v);
                break;
            }
            case ir::H_TRI: {
                const ir::TriP& p = S.tris[h.a];
                V3 v0 = X.point(p.v[0]), v1 = X.point(p.v[1]), v2 = X.point(p.v[2]);
                V3 e0 = v1 - v0, e1 = v2 - v0;
                GrtTri g;
                memset(&g, 0, sizeof(g));
                g.v0[0] = (float)v0.x; g.v0[1] = (float)v0.y; g.v0[2] = (float)v0.z;
                g.e0[0] = (float)e0.x; g.e0[1] = (float)e0.y; g.e0[2] = (float)e0.z;
                g.e1[0] = (float)e1.x; g.e1[1] = (float)e1.y; g.e1[2] = (float)e1.z;
                g.mat = (uint32_t)h.mat; g.id = (uint32_t)hid;
                g.flags = (p.hasNormals ? GRT_TRI_HAS_NORMALS : 0) | (p.hasUV ? GRT_TRI_HAS_UV : 0);
                GrtTriShade sh;
                memset(&sh, 0, sizeof(sh));
                if (p.hasNormals) {
                    V3 n0 = X.rot(p.n[0]), n1 = X.rot(p.n[1]), n2 = X.rot(p.n[2]);
                    sh.n0[0] = (float)n0.x; sh.n0[1] = (float)n0.y; sh.n0[2] = (float)n0.z;
                    sh.n1[0] = (float)n1.x; sh.n1[1] = (float)n1.y; sh.n1[2] = (float)n1.z;
                    sh.n2[0] = (float)n2.x; sh.n2[1] = (float)n2.y; sh.n2[2] = (float)n2.z;
                }
                if (p.hasUV) for (int k = 0; k < 3; k++) { sh.uv[2 * k] = (float)p.uv[k][0]; sh.uv[2 * k + 1] = (float)p.uv[k][1]; }
                if (g.flags) F->any_tri_shade = true;
                F->tris.push_back(g);
                F->tri_shade.push_back(sh);
                { const V3* vv[3] = {&v0, &v1, &v2}; for (int k = 0; k < 3; k++) { F->tri_v64.push_back(vv[k]->x); F->tri_v64.push_back(vv[k]->y); F->tri_v64.push_back(vv[k]->z); } }
                e.ref = GRT_MAKE_REF(GRT_REF_TRI, F->tris.size() - 1);
                e.box.add(v0); e.box.add(v1); e.box.add(v2);
                break;
            }
            case ir::H_LIST: {
                e = emitRun(S.lists[h.a], X, inBoundary);
                break;
            }
            case ir::H_BVH: {
                if (isBoxPrim(h.a)) {
                    // NewBox: emit its six quads contiguously in the reference's order, plus the slab record
                    const ir::BoxP& bp = S.boxes[S.box_of_list[h.a]];
                    uint32_t first = (uint32_t)F->quads.size();
                    for (int i = 0; i < 6; i++) { Emitted q = emit(bp.quads[i], X, inBoundary); e.box.add(q.box); }
                    GrtBox b;
                    memset(&b, 0, sizeof(b));
                    b.mn[0] = (float)bp.mn.x; b.mn[1] = (float)bp.mn.y; b.mn[2] = (float)bp.mn.z;
                    b.mx[0] = (float)bp.mx.x; b.mx[1] = (float)bp.mx.y; b.mx[2] = (float)bp.mx.z;
                    b.first_quad = first;
                    b.T[0] = (float)X.T.x; b.T[1] = (float)X.T.y; b.T[2] = (float)X.T.z;
                    b.rc = (float)X.c; b.rs = (float)X.s;
                    F->boxes.push_back(b);
                    e.ref = GRT_MAKE_REF(GRT_REF_BOX, F->boxes.size() - 1);
                    e.need = 0;
                    break;
                }
                const Topology& T = topology(h.a);
                e = emitTopo(T, T.root, X, inBoundary);
                break;
            }
            case ir::H_TRANSLATE: {
                Xform Y = X;
                Y.T = X.rot(S.xforms[h.a].offset) + X.T;
                e = emit(h.child, Y, inBoundary);
                break;
            }
            case ir::H_ROTATEY: {
                double radians = S.xforms[h.a].degrees * 3.14159265358979323846 / 180.0;  // util.DegressToRadians
                double sn = std::sin(radians), cs = std::cos(radians);
                Xform Y = X;
                Y.c = X.c * cs - X.s * sn;
                Y.s = X.s * cs + X.c * sn;
                e = emit(h.child, Y, inBoundary);
                break;
            }
            case ir::H_MEDIUM: {
                if (inBoundary) throw std::runtime_error("constantMedium nested inside a medium boundary is not supported");
                Emitted b = emit(h.child, X, true);
                GrtMedium m;
                m.boundary = b.ref;
                m.neg_inv_density = (float)(-1 / S.media[h.a].density);
                m.mat = (uint32_t)S.media[h.a].phase;
                m.id = (uint32_t)hid;
                F->media.push_back(m);
                e.ref = GRT_MAKE_REF(GRT_REF_MEDIUM, F->media.size() - 1);
                e.box = b.box;
                e.has_medium = true;
                mediumNeed = std::max(mediumNeed, b.need + 1);
                break;
            }
            default: throw std::runtime_error("unknown hittable type");
        }
        return e;
    }

    void emitMaterials() {
        for (const ir::Image& im : S.images) {
            GrtImage g;
            g.width = (uint32_t)im.width; g.height = (uint32_t)im.height; g.offset = F->texels.size();
            F->texels.insert(F->texels.end(), im.rgb.begin(), im.rgb.end());
            F->images.push_back(g);
        }
        while (F->texels.size() % 16) F->texels.push_back(0);
        for (const ir::Perlin& p : S.perlins) {
            GrtPerlin g;
            memset(&g, 0, sizeof(g));
            for (int i = 0; i < 256; i++) {
                for (int k = 0; k < 3; k++) g.grad[i][k] = (float)p.vec[i][k];
                for (int a = 0; a < 3; a++) g.perm[a][i] = (uint8_t)p.perm[a][i];
            }
            F->perlins.push_back(g);
        }
        for (const ir::Texture& t : S.textures) {
            GrtTexture g;
            memset(&g, 0, sizeof(g));
            g.type = (uint32_t)t.type;
            g.color[0] = (float)t.color.x; g.color[1] = (float)t.color.y; g.color[2] = (float)t.color.z;
            switch (t.type) {
                case ir::TEX_SOLID: break;
                case ir::TEX_CHECKER: g.scale = (float)(1 / t.scale); g.even = (uint32_t)t.even; g.odd = (uint32_t)t.odd; break;  // texture.go:37
                case ir::TEX_IMAGE: g.aux = (uint32_t)t.image; break;
                case ir::TEX_NOISE: g.scale = (float)t.scale; g.aux = (uint32_t)t.perlin | ((uint32_t)t.variant << 16); break;
            }
            F->textures.push_back(g);
        }
        for (const ir::Material& m : S.materials) {
            GrtMaterial g;
            memset(&g, 0, sizeof(g));
            g.type = (uint32_t)m.type;
            g.tex = m.tex >= 0 ? (uint32_t)m.tex : 0;
            g.albedo[0] = (float)m.albedo.x; g.albedo[1] = (float)m.albedo.y; g.albedo[2] = (float)m.albedo.z;
            g.fuzz = (float)m.fuzz; g.ior = (float)m.ior;
            F->materials.push_back(g);
        }
    }

    void addLight(int hid) {
        const ir::Hittable& h = S.hittables[hid];
        GrtLight L;
        memset(&L, 0, sizeof(L));
        L.prim = GRT_MAKE_REF(GRT_REF_NONE, 0);
        switch (h.type) {
            case ir::H_SPHERE: {
                const ir::SphereP& p = S.spheres[h.a];
                L.type = GRT_LIGHT_SPHERE;
                L.p[0] = p.c0.x; L.p[1] = p.c0.y; L.p[2] = p.c0.z; L.p[3] = p.r;  // Center.At(0), objects.go:57,64
                // the sphere light moves with ray time 0 in PdfValue (ray.New -> time 0), so dc is irrelevant there
                break;
            }
            case ir::H_QUAD: {
                const ir::QuadP& q = S.quads[h.a];
                V3 n = cross(q.u, q.v);
                double area = length(n);
                V3 normal = n * (1 / area);
                double D = dot(normal, q.Q);
                V3 w = n * (1 / dot(n, n));
                L.type = GRT_LIGHT_QUAD;
                double vals[17] = {q.Q.x, q.Q.y, q.Q.z, q.u.x, q.u.y, q.u.z, q.v.x, q.v.y, q.v.z, normal.x, normal.y, normal.z, w.x, w.y, w.z, D, area};
                for (int i = 0; i < 17; i++) L.p[i] = vals[i];
                break;
            }
            case ir::H_TRI: {
                const ir::TriP& t = S.tris[h.a];
                L.type = GRT_LIGHT_TRI;
                for (int k = 0; k < 3; k++) { L.p[3 * k] = t.v[k].x; L.p[3 * k + 1] = t.v[k].y; L.p[3 * k + 2] = t.v[k].z; }
                L.p[9] = length(cross(t.v[1] - t.v[0], t.v[2] - t.v[0])) / 2.0;  // objects.go:263
                if (t.hasNormals) {
                    L.flags = GRT_TRI_HAS_NORMALS;
                    for (int k = 0; k < 3; k++) { L.p[10 + 3 * k] = t.n[k].x; L.p[11 + 3 * k] = t.n[k].y; L.p[12 + 3 * k] = t.n[k].z; }
                }
                break;
            }
            default:
                // BVHNode / translate / rotateY / constantMedium embed defaultPdfImpl whose PdfValue is
                // log.Fatal("hit an invalid PDF function") (hittable.go:69-72)
                throw std::runtime_error("lights must be spheres, quads or triangles (reference: hit an invalid PDF function)");
        }
        for (int i = 0; i < 24; i++) L.f[i] = (float)L.p[i];
        F->lights.push_back(L);
    }
    void emitLights() {
        const ir::Hittable& h = S.hittables[S.lights];
        if (h.type == ir::H_LIST) {
            F->lights_mode = GRT_LIGHTS_LIST;
            for (int c : S.lists[h.a]) {
                if (S.hittables[c].type == ir::H_LIST) throw std::runtime_error("nested light lists are not supported");
                addLight(c);
            }
        } else {
            F->lights_mode = GRT_LIGHTS_BARE;
            addLight(S.lights);
        }
    }
};

// Camera.initialize (camera.go:179-253) -> GrtCamera, in fp64.
inline bool deriveCamera(const ir::CameraConfig& in, GrtCamera& out, std::string& err) {
    ir::CameraConfig c = in;
    if (c.AspectRatio == 0) c.AspectRatio = 1.0;
    if (c.Width == 0) c.Width = 100;
    if (c.SamplesPerPixel == 0) c.SamplesPerPixel = 100;
    if (c.MaxDepth == 0) c.MaxDepth = 10;
    if (c.VerticalFOV == 0) c.VerticalFOV = 90;
    if (c.FocusDistance == 0) c.FocusDistance = 10;
    if (c.MaxContribution == 0) c.MaxContribution = 1.5;
    if (c.Width < 0 || c.SamplesPerPixel < 0 || c.MaxDepth < 0) { err = "negative camera field"; return false; }
    int imageHeight = std::max(1, (int)((double)c.Width / c.AspectRatio));
    int sppSqrt = (int)std::sqrt((double)c.SamplesPerPixel);
    if (sppSqrt < 1) { err = "SamplesPerPixel < 1"; return false; }
    const double PI = 3.14159265358979323846;
    double theta = c.VerticalFOV * PI / 180.0;
    double h = std::tan(theta / 2);
    double viewportHeight = 2.0 * h * c.FocusDistance;
    double viewportWidth = viewportHeight * ((double)c.Width / (double)imageHeight);
    V3 d = c.lookFrom - c.lookAt;
    V3 w = d * (1 / length(d));
    V3 cu = cross(c.vup, w);
    V3 u = cu * (1 / length(cu));
    V3 v = cross(w, u);
    V3 viewportU = u * viewportWidth;
    V3 viewportV = ir::neg(v) * viewportHeight;
    V3 du = viewportU * (1.0 / (double)c.Width);
    V3 dv = viewportV * (1.0 / (double)imageHeight);
    V3 topLeft = c.lookFrom - w * c.FocusDistance - viewportU * 0.5 - viewportV * 0.5;
    V3 p00 = topLeft + (du + dv) * 0.5;
    double defocusRadius = c.FocusDistance * std::tan((c.DefocusAngle / 2.0) * PI / 180.0);
    V3 defU = u * defocusRadius, defV = v * defocusRadius;
    memset(&out, 0, sizeof(out));
    out.width = c.Width; out.height = imageHeight; out.spp_sqrt = sppSqrt; out.max_depth = c.MaxDepth;
    double* dst[6] = {out.center, out.pixel00, out.delta_u, out.delta_v, out.defocus_u, out.defocus_v};
    V3 src[6] = {c.lookFrom, p00, du, dv, defU, defV};
    for (int i = 0; i < 6; i++) { dst[i][0] = src[i].x; dst[i][1] = src[i].y; dst[i][2] = src[i].z; }
    out.defocus_angle = c.DefocusAngle;
    out.background[0] = c.Background.x; out.background[1] = c.Background.y; out.background[2] = c.Background.z;
    out.max_contribution = c.MaxContribution;
    return true;
}

}  // namespace flat
}  // namespace grt
