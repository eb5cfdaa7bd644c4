// dev_trace.cuh — device scene view, primitive intersection and the stack-based
// BVH traversal (closest hit) in the reference's child order.
//
// Reference: internal/hittable/bvh.go:69-82 (BVHNode.Hit), hittable.go:122-138
// (HittableList.Hit), aabb/aabb.go:90-113 (slab test), objects.go:83-115
// (sphere.Hit), :167-206 (quad.Hit, isInterior), :408-461 (Triangle.Hit),
// medium.go:27-58 (constantMedium.Hit).
#pragma once
#include "grt_internal.h"
#include "dev_math.cuh"
#include "../../include/grt.h"

namespace grtd {

// feature bits: kernels are compiled for feature supersets so that e.g. the
// Cornell box never carries sphere / triangle / texture code.
enum : uint32_t {
    F_SPHERE = 1u << 0, F_QUAD = 1u << 1, F_TRI = 1u << 2, F_LIST = 1u << 3, F_MEDIUM = 1u << 4,
    F_SPECULAR = 1u << 5, F_TEXTURE = 1u << 6, F_ISOTROPIC = 1u << 7, F_ROTQUAD = 1u << 8,
    F_SPHERE_LIGHT = 1u << 9, F_QUAD_LIGHT = 1u << 10, F_TRI_LIGHT = 1u << 11, F_TRISHADE = 1u << 12,
    F_DEFOCUS = 1u << 13, F_NODE = 1u << 14,
    F_DUPIDS = 1u << 15,   // some object id names more than one flat primitive: exclude by id, not by flat ref
    F_BOX = 1u << 17,      // NewBox results tested as one slab test (GrtBox)
    F_ALL = ((1u << 16) - 1) | F_BOX,
    // not a scene feature: the C-ABI ray query accepts rays that start ON a plane without naming it, so its
    // t >= tmin decisions fall back to the fp64 plane when the origin is within fp32 resolution of that plane.
    // The integrator always names the primitive it starts on (self exclusion), which leaves such events at
    // ~1e-7 per segment, and compiles the fallback out of its quad loop.
    F_TMIN_F64 = 1u << 16,
    // not a scene feature either: the scene arrays are read from global memory (nothing staged in shared memory), so the
    // traversal fetches them through the read-only path (ld.global.nc, __ldg) instead of generic loads
    F_GMEM = 1u << 18,
    // the first SceneView::tl_n nodes (the top of the tree: nodes are numbered breadth first) are also held in shared
    // memory by this kernel and fetched from there
    F_TREELET = 1u << 19
};
// Which loads take the read-only path: the node fetches always do (F_GMEM); the primitive records only where it paid —
// ptxas schedules ld.global.nc loads earlier and wider, which cost the all-feature extend kernel 1.1 KB of spills.
#ifndef GRT_NC_PRIMS
#define GRT_NC_PRIMS 0
#endif
#define GRT_PRIM_NC(FEAT) (GRT_NC_PRIMS != 0 && ((FEAT) & F_GMEM) != 0)
#ifndef GRT_EAGER_LEAF
#define GRT_EAGER_LEAF 0   /* measured: 444 vs 469 (mesh) and 486 vs 562 (book 2) Mpaths/s with it on — see node_step */
#endif
#ifndef GRT_NC_NODES
#define GRT_NC_NODES 1
#endif
// load through the read-only data path when the pointer is known to be global memory
template <bool NC, class T>
__device__ __forceinline__ T ldro(const T* p) {
    if constexpr (NC) return __ldg(p);
    else return *p;
}
// feature-set variants the kernels are instantiated for (a scene runs on the smallest one that covers it)
#define V_CORNELL (F_QUAD | F_BOX | F_LIST | F_ROTQUAD | F_QUAD_LIGHT)
#define V_SMOKE (V_CORNELL | F_MEDIUM | F_ISOTROPIC)
#define V_SPHERES (F_NODE | F_SPHERE | F_LIST | F_SPECULAR | F_TEXTURE | F_SPHERE_LIGHT | F_DEFOCUS)
#define V_MESH (F_NODE | F_SPHERE | F_TRI | F_LIST | F_SPECULAR | F_SPHERE_LIGHT | F_TRI_LIGHT | F_TRISHADE | F_DEFOCUS)
#define V_FULL F_ALL
#define V_FULL_UNIQ (F_ALL & ~F_DUPIDS)   /* everything, quads excluded by flat ref (no duplicated quad ids) */
#define GRT_NEEDS_F64(FEAT) (((FEAT) & F_SPHERE) != 0)
// The fp64 copies of the ray (14 registers) stay live only in the sphere-only variant, where a segment makes ~40 sphere
// tests; variants that also traverse triangle BVHs convert on the fly (6 cvt + 3 DFMA per sphere test) — in the mesh
// extend kernel those registers were spilled around every node visit (ncu: local loads + stores = 45 % of its L1 sectors)
#define GRT_RAY_KEEPS_F64(FEAT) ((((FEAT) & F_SPHERE) != 0) && (((FEAT) & F_TRI) == 0))

// Device-internal quad records, repacked from GrtQuad at upload.
// Hot (48 B, read by every test): plane, and the interior test folded into two affine forms
//   alpha = A.xyz . p + A.w   (A.w = -A.Q),   beta = B.xyz . p + B.w
// Cold (80 B, read once per hit): normal, the precomputed orthonormal basis of onb.go:13-25 for
// the FRONT-face normal (the back face is (u, -v, -n)), ids, and the fp64 plane for refinement.
struct DQuadHot { float4 plane, A, B; };
struct DQuadCold {
    float n[3]; uint32_t flags;
    float ou[3]; uint32_t mat;
    float ov[3]; uint32_t id;
    double n64[3]; double D64;
};

// Device-internal fp32 record of a quad light (fast path of dev_shade.cuh), built at upload from GrtLight.p:
// the re-intersection of objects.go:152-160 needs the plane and the two affine interior forms, sampling
// (objects.go:161-165) needs Q, u, v.
struct DLight { float4 plane, A, B, Qa, U, V; };   // Qa.w = area

// The small, hot arrays live in one blob (byte offsets below) so a block can
// stage the whole thing in shared memory when it fits; large arrays stay in HBM.
struct DevScene {
    const unsigned char* blob;   // HBM copy of the blob
    uint32_t blob_bytes;
    uint32_t stage_bytes;        // prefix of the blob every block copies to shared memory: all of it, the hot arrays only, or 0
    uint32_t off_nodes, off_spheres, off_quads, off_quads_cold, off_boxes, off_items, off_media, off_materials, off_textures, off_lights, off_dlights, off_images;
    uint32_t n_nodes, n_spheres, n_quads, n_boxes, n_items, n_media, n_materials, n_textures, n_lights, n_images;
    const GrtTri* tris;          // HBM
    const GrtTriShade* tri_shade;
    const double* tri_v64;
    const uint8_t* texels;
    const GrtPerlin* perlins;
    uint32_t n_tris, n_perlins;
    uint32_t root, lights_mode, features, stack_need;
};

// View over either the shared-memory or the HBM copy of the blob.
// The blob is laid out hot arrays first (nodes, spheres, quad test records, boxes, list entries, media), then the
// per-hit arrays (quad cold records, materials, textures, lights, images).  `base` serves the hot arrays, `cold`
// the rest; both point into shared memory when the whole blob is staged.
struct SceneView {
    const unsigned char* base;
    const unsigned char* cold;
    const DevScene* ds;
    uint32_t tl = 0, tl_n = 0;   // F_TREELET: shared-window address of the staged top-of-tree nodes, and how many
    __device__ __forceinline__ const float4* nodes() const { return (const float4*)(base + ds->off_nodes); }
    __device__ __forceinline__ const GrtSphere* spheres() const { return (const GrtSphere*)(base + ds->off_spheres); }
    __device__ __forceinline__ const DQuadHot* quads() const { return (const DQuadHot*)(base + ds->off_quads); }
    __device__ __forceinline__ const DQuadCold* quads_cold() const { return (const DQuadCold*)(cold + ds->off_quads_cold); }
    __device__ __forceinline__ const GrtBox* boxes() const { return (const GrtBox*)(base + ds->off_boxes); }
    // run-length list entries built at upload: x = first ref (| GRT_LIST_LAST), y = number of consecutive primitives
    __device__ __forceinline__ const uint2* entries() const { return (const uint2*)(base + ds->off_items); }
    __device__ __forceinline__ const GrtMedium* media() const { return (const GrtMedium*)(base + ds->off_media); }
    __device__ __forceinline__ const GrtMaterial* materials() const { return (const GrtMaterial*)(cold + ds->off_materials); }
    __device__ __forceinline__ const GrtTexture* textures() const { return (const GrtTexture*)(cold + ds->off_textures); }
    __device__ __forceinline__ const GrtLight* lights() const { return (const GrtLight*)(cold + ds->off_lights); }
    __device__ __forceinline__ const DLight* dlights() const { return (const DLight*)(cold + ds->off_dlights); }
    __device__ __forceinline__ const GrtImage* images() const { return (const GrtImage*)(cold + ds->off_images); }
};

// Cooperative copy of the staged prefix of the blob into shared memory (16-byte vectors).
__device__ __forceinline__ void stage_blob(unsigned char* smem, const DevScene& ds) {
    const uint4* src = (const uint4*)ds.blob;
    uint4* dst = (uint4*)smem;
    uint32_t n = ds.stage_bytes >> 4;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}

// STAGED: 0 = read everything from HBM (through L1/L2), 1 = hot arrays in shared memory, 2 = the whole blob in
// shared memory (then `cold` IS `base`, and costs no register).
template <int STAGED>
__device__ __forceinline__ SceneView make_view(const DevScene& ds, unsigned char* smem) {
    SceneView sv;
    sv.ds = &ds;
    if (STAGED) stage_blob(smem, ds);
    sv.base = STAGED ? smem : ds.blob;
    sv.cold = STAGED == 2 ? smem : ds.blob;
    return sv;
}

struct RayD {
    f3 o, d, invd;
    float time;
    // float4 index, inside a 4-wide node, of the NEAR x / y / z planes of its four children (0/3, 1/4, 2/5 by the sign
    // of 1/d; the far planes are 3 - nx, 5 - ny, 7 - nz): the slab test needs no per-axis min/max or swap (aabb.go:102-104)
    uint32_t nx, ny, nz;
    // fp64 copies for the cancellation-prone sums (sphere quadratic, rotated-quad planes)
    d3 o64, d64;
    double a64;  // d.d
};
template <uint32_t FEAT>
__device__ __forceinline__ void ray_setup(RayD& r, f3 o, f3 d, float time) {
    r.o = o; r.d = d; r.time = time;
    if (FEAT & F_NODE) {
        r.invd = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // aabb.go:96 (only box tests use it)
        r.nx = r.invd.x < 0.0f ? 3u : 0u; r.ny = r.invd.y < 0.0f ? 4u : 1u; r.nz = r.invd.z < 0.0f ? 5u : 2u;
    }
    if (GRT_RAY_KEEPS_F64(FEAT)) {
        r.o64 = tod3(o); r.d64 = tod3(d);
        r.a64 = dot(r.d64, r.d64);
    }
}

struct HitInfo {
    float t;
    uint32_t ref;   // flat ref of the primitive / medium that was hit
    float u, v;     // quad alpha/beta, triangle barycentrics (sphere: unused)
};

struct TraceCounters {
    uint32_t box, sphere, quad, tri, medium;
};

// ---- slab test, aabb.go:90-113 -------------------------------------------
// Same expression as the reference: t = (bound - origin) * (1/d), swap when
// 1/d < 0, shrink, reject when max <= min.  fminf/fmaxf drop NaN where Go's
// min/max propagate it; the difference only makes Go accept boxes that are
// geometrically missed, never changes a closest hit.
__device__ __forceinline__ bool box_hit(float lx, float ly, float lz, float hx, float hy, float hz, const RayD& r, float tmin, float tmax) {
    float t0x = (lx - r.o.x) * r.invd.x, t1x = (hx - r.o.x) * r.invd.x;
    float t0y = (ly - r.o.y) * r.invd.y, t1y = (hy - r.o.y) * r.invd.y;
    float t0z = (lz - r.o.z) * r.invd.z, t1z = (hz - r.o.z) * r.invd.z;
    float ax = r.invd.x < 0 ? t1x : t0x, bx = r.invd.x < 0 ? t0x : t1x;
    float ay = r.invd.y < 0 ? t1y : t0y, by = r.invd.y < 0 ? t0y : t1y;
    float az = r.invd.z < 0 ? t1z : t0z, bz = r.invd.z < 0 ? t0z : t1z;
    float lo = fmaxf(fmaxf(ax, ay), fmaxf(az, tmin));
    float hi = fminf(fminf(bx, by), fminf(bz, tmax));
    return !(hi <= lo);
}

// Device BVH node: 4-wide, 128 bytes (8 x float4), built at upload from the ABI's binary GrtNode array (wide_bvh.hpp):
//   f4[0..2] = lo.x, lo.y, lo.z of the four children   f4[3..5] = hi.x, hi.y, hi.z
//   f4[6]    = the four child refs (inner node, DREF_RUN primitive run, list, medium; NONE = empty slot, empty box)
//   f4[7].x  = bit 0: children may be visited nearest first (no constantMedium below), .y = number of children
// Every child carries its own box — leaf runs too, which the reference does not box-test (bvh.go:73-79); a
// conservative box cannot change a closest hit, it only saves fetching primitives the ray passes by.
#define GRT_WNODE_F4 8
// child ref of a run of consecutive primitives: bit 31 | type << 28 | (count - 1) << 25 | first index (wide_bvh.hpp)
#define GRT_DREF_RUN_BIT 0x80000000u
#define GRT_DREF_RUN_INDEX_MASK 0x01FFFFFFu

// ---- sphere.Hit, objects.go:83-115 ----------------------------------------
// The cancellation-prone sums (c = |oc|^2 - r^2, disc = h^2 - a c) are fp64;
// the roots use the cancellation-free pair {c/q, q/a}, q = h + sign(h) sqrt(disc),
// which equals {(h-s)/a, (h+s)/a} of the reference up to rounding.
// `self`: the ray origin lies on this sphere; exact arithmetic has c == 0.
template <bool KEEPS_F64, bool NC = false>
__device__ __forceinline__ bool sphere_roots(const GrtSphere& sg, const RayD& r, bool self, float& nearr, float& farr) {
    const d3 o64 = KEEPS_F64 ? r.o64 : tod3(r.o), d64 = KEEPS_F64 ? r.d64 : tod3(r.d);
    const double a64 = KEEPS_F64 ? r.a64 : dot(d64, d64);
    // c0.xy | c0.z, r | dc.xyz, mat: three 16-byte loads of the 64-byte record
    struct { double c0[3], r; float dc[3]; } s;
    if constexpr (NC) {
        const double2 c01 = __ldg((const double2*)&sg.c0[0]), c2r = __ldg((const double2*)&sg.c0[2]);
        const float4 dcm = __ldg((const float4*)&sg.dc[0]);
        s.c0[0] = c01.x; s.c0[1] = c01.y; s.c0[2] = c2r.x; s.r = c2r.y; s.dc[0] = dcm.x; s.dc[1] = dcm.y; s.dc[2] = dcm.z;
    } else {
        s.c0[0] = sg.c0[0]; s.c0[1] = sg.c0[1]; s.c0[2] = sg.c0[2]; s.r = sg.r; s.dc[0] = sg.dc[0]; s.dc[1] = sg.dc[1]; s.dc[2] = sg.dc[2];
    }
    double cx = s.c0[0] + (double)r.time * (double)s.dc[0];
    double cy = s.c0[1] + (double)r.time * (double)s.dc[1];
    double cz = s.c0[2] + (double)r.time * (double)s.dc[2];
    double ocx = cx - o64.x, ocy = cy - o64.y, ocz = cz - o64.z;
    double h = d64.x * ocx + d64.y * ocy + d64.z * ocz;
    double c = ocx * ocx + ocy * ocy + ocz * ocz - s.r * s.r;
    if (self) c = 0.0;
    double disc = h * h - a64 * c;
    if (disc < 0) return false;
    float hf = (float)h, af = (float)a64, cf = (float)c;
    float sq = sqrtf((float)disc);
    if (hf >= 0) { float q = hf + sq; farr = q / af; nearr = cf / q; }
    else { float q = hf - sq; nearr = q / af; farr = cf / q; }
    return true;
}
// Surrounds: open interval, nearer root first (objects.go:101-105)
__device__ __forceinline__ bool sphere_pick(float nearr, float farr, float tmin, float tmax, float& t_out) {
    float root = nearr;
    if (!(tmin < root && root < tmax)) {
        root = farr;
        if (!(tmin < root && root < tmax)) return false;
    }
    t_out = root;
    return true;
}
template <bool KEEPS_F64, bool NC = false>
__device__ __forceinline__ bool sphere_hit(const GrtSphere& s, const RayD& r, float tmin, float tmax, bool self, float& t_out) {
    float nearr, farr;
    if (!sphere_roots<KEEPS_F64, NC>(s, r, self, nearr, farr)) return false;
    return sphere_pick(nearr, farr, tmin, tmax, t_out);
}

// One MUFU.RCP (1 ulp) and a multiply; __fdividef adds ~5 range-scaling instructions we do not need
// (a denominator below 1e-8 is rejected anyway).
__device__ __forceinline__ float fast_div(float x, float y) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return x * r;
}

// ---- quad.Hit + isInterior, objects.go:167-206 -----------------------------
// Branch-free fp32 candidate test on the 48-byte hot record.  For a quad that
// is not axis-aligned the fp32 plane distance D - n.o loses relative accuracy
// when the origin is close to the plane; the WINNING hit is therefore refined
// in fp64 by quad_refine_t (one refinement per segment instead of fp64
// arithmetic in every test).
// `uncertain` is set when the plane distance at tmin, D - n.(o + tmin d), is within fp32 rounding of zero (the
// origin lies almost on this plane): the accept/reject decision t >= tmin then needs the fp64 plane.
template <bool NC = false>
__device__ __forceinline__ bool quad_hit(const DQuadHot* q, const RayD& r, float tmin, float tmax, float& t_out, float& a_out, float& b_out, bool& uncertain) {
    const float4 P = ldro<NC>(&q->plane), A = ldro<NC>(&q->A), B = ldro<NC>(&q->B);
    const float denom = P.x * r.d.x + P.y * r.d.y + P.z * r.d.z;
    const float num = P.w - (P.x * r.o.x + P.y * r.o.y + P.z * r.o.z);
    const float t = fast_div(num, denom);          // 2 ulp, far inside the 1e-5 budget
    const float px = fmaf(t, r.d.x, r.o.x), py = fmaf(t, r.d.y, r.o.y), pz = fmaf(t, r.d.z, r.o.z);
    const float alpha = fmaf(A.x, px, fmaf(A.y, py, fmaf(A.z, pz, A.w)));
    const float beta = fmaf(B.x, px, fmaf(B.y, py, fmaf(B.z, pz, B.w)));
    // 0 <= x <= 1  <=>  x - x*x >= 0 (one FMA, exact in sign; closed like objects.go:199).  Moves four compares
    // from the ALU pipe (the busiest pipe of this loop, profiles/r1_mega_v4) to the FMA pipe.
    const float ia = fmaf(-alpha, alpha, alpha), ib = fmaf(-beta, beta, beta);
    const bool ok = (fabsf(denom) >= 1e-8f) & (tmin <= t) & (t <= tmax)            // objects.go:171,177 (closed interval)
                    & (fminf(ia, ib) >= 0.0f) & (alpha == alpha) & (beta == beta);
    t_out = t; a_out = alpha; b_out = beta;
    uncertain = fabsf(fmaf(-tmin, denom, num)) < 2.5e-4f;
    return ok;
}
// The same test with the fp64 plane (rare path: origin within fp32 resolution of the plane).
__device__ __forceinline__ bool quad_hit_f64(const DQuadHot* q, const DQuadCold* c, const RayD& r, float tmin, float tmax, float& t_out, float& a_out, float& b_out) {
    const double denom = c->n64[0] * (double)r.d.x + c->n64[1] * (double)r.d.y + c->n64[2] * (double)r.d.z;
    if (fabs(denom) < 1e-8) return false;
    const double num = c->D64 - (c->n64[0] * (double)r.o.x + c->n64[1] * (double)r.o.y + c->n64[2] * (double)r.o.z);
    const double t64 = num / denom;
    if (!((double)tmin <= t64 && t64 <= (double)tmax)) return false;
    const float t = (float)t64;
    const float4 A = q->A, B = q->B;
    const float px = fmaf(t, r.d.x, r.o.x), py = fmaf(t, r.d.y, r.o.y), pz = fmaf(t, r.d.z, r.o.z);
    const float alpha = fmaf(A.x, px, fmaf(A.y, py, fmaf(A.z, pz, A.w)));
    const float beta = fmaf(B.x, px, fmaf(B.y, py, fmaf(B.z, pz, B.w)));
    if (!((0.0f <= alpha) & (alpha <= 1.0f) & (0.0f <= beta) & (beta <= 1.0f))) return false;
    t_out = t; a_out = alpha; b_out = beta;
    return true;
}
__device__ __forceinline__ float quad_refine_t(const DQuadCold* q, const RayD& r, float t32) {
    if (q->flags & GRT_QUAD_AXIS_ALIGNED) return t32;   // D - n.o and n.d are single-rounding there
    const double denom = q->n64[0] * (double)r.d.x + q->n64[1] * (double)r.d.y + q->n64[2] * (double)r.d.z;
    const double num = q->D64 - (q->n64[0] * (double)r.o.x + q->n64[1] * (double)r.o.y + q->n64[2] * (double)r.o.z);
    return __fdividef((float)num, (float)denom);
}

// ---- NewBox as one primitive (objects.go:208-240) ------------------------------
// The six quads of a box are found with one slab test in the box's own frame (the ray is taken
// to object space exactly as translate.Hit / rotateY.Hit do, transformation.go:25-34,94-107).
// A line meets a convex box in at most two points, t_enter <= t_exit, each on one face; quad.Hit
// over the six faces returns the smaller one that lies in [tmin, tmax] (closed, objects.go:177).
// Returns the face index in NewBox's order (front, right, back, left, top, bottom) or -1.
// `excl_face`: face of THIS box the ray starts on (-1 if none).  `near_tmin` reports a candidate
// within fp32 resolution of tmin (see F_TMIN_F64).
struct BoxSlab { float t_enter, t_exit, d_enter, d_exit; int f_enter, f_exit; };
template <bool NC = false>
__device__ __forceinline__ bool box_prim_slab(const GrtBox* bx, const RayD& r, BoxSlab& o) {
    const float4 b0 = ldro<NC>((const float4*)&bx->mn[0]), b1 = ldro<NC>((const float4*)&bx->mx[0]), b2 = ldro<NC>((const float4*)&bx->T[0]);
    const float rs = ldro<NC>(&bx->rs), rc = b2.w;
    const float px = r.o.x - b2.x, py = r.o.y - b2.y, pz = r.o.z - b2.z;
    const float ox = rc * px - rs * pz, oz = rs * px + rc * pz;          // rayTranslationHelper, transformation.go:79-85
    const float dx = rc * r.d.x - rs * r.d.z, dz = rs * r.d.x + rc * r.d.z, dy = r.d.y;
    float ix, iy, iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(dx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(dy));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(dz));
    const float ax = (b0.x - ox) * ix, bx_ = (b1.x - ox) * ix;
    const float ay = (b0.y - py) * iy, by = (b1.y - py) * iy;
    const float az = (b0.z - oz) * iz, bz = (b1.z - oz) * iz;
    const float lox = fminf(ax, bx_), hix = fmaxf(ax, bx_);
    const float loy = fminf(ay, by), hiy = fmaxf(ay, by);
    const float loz = fminf(az, bz), hiz = fmaxf(az, bz);
    o.t_enter = fmaxf(fmaxf(lox, loy), loz); o.t_exit = fminf(fminf(hix, hiy), hiz);
    if (!(o.t_enter <= o.t_exit)) return false;
    // faces: 0 front z=max, 1 right x=max, 2 back z=min, 3 left x=min, 4 top y=max, 5 bottom y=min
    // entering through axis a: the min face when d_a > 0, else the max face; leaving: the other way round
    // (selects, no branches: the three axes are equally likely and would split the warp three ways)
    const int fx_in = dx > 0 ? 3 : 1, fy_in = dy > 0 ? 5 : 4, fz_in = dz > 0 ? 2 : 0;
    const bool ex = (lox >= loy) & (lox >= loz), ey = loy >= loz;
    o.f_enter = ex ? fx_in : (ey ? fy_in : fz_in);
    o.d_enter = ex ? dx : (ey ? dy : dz);
    const bool xx = (hix <= hiy) & (hix <= hiz), xy = hiy <= hiz;
    o.f_exit = xx ? (4 - fx_in) : (xy ? (9 - fy_in) : (2 - fz_in));   // the opposite face of the same axis
    o.d_exit = xx ? dx : (xy ? dy : dz);
    return true;
}
// the smaller of (t_enter, t_exit) that lies in [tmin, tmax] (closed, objects.go:177) and is not on the excluded face
__device__ __forceinline__ int box_prim_pick(const BoxSlab& o, float tmin, float tmax, int excl_face, float& t_out, bool& near_tmin) {
    near_tmin = (fabsf((o.t_enter - tmin) * o.d_enter) < 2.5e-4f) | (fabsf((o.t_exit - tmin) * o.d_exit) < 2.5e-4f);
    if (tmin <= o.t_enter && o.t_enter <= tmax && o.f_enter != excl_face) { t_out = o.t_enter; return o.f_enter; }
    if (tmin <= o.t_exit && o.t_exit <= tmax && o.f_exit != excl_face) { t_out = o.t_exit; return o.f_exit; }
    return -1;
}
template <bool NC = false>
__device__ __forceinline__ int box_prim_hit(const GrtBox* bx, const RayD& r, float tmin, float tmax, int excl_face, float& t_out, bool& near_tmin) {
    BoxSlab o;
    near_tmin = false;
    if (!box_prim_slab<NC>(bx, r, o)) return -1;
    return box_prim_pick(o, tmin, tmax, excl_face, t_out, near_tmin);
}

// ---- Triangle.Hit (Möller–Trumbore), objects.go:408-461 --------------------
// The same test on the fp64 vertices, in the reference's expression order (objects.go:409-433): used by the ray-query
// kernel when an fp32 barycentric lies within fp32 error of an edge (see tri_hit).
static __device__ __noinline__ bool tri_hit_f64(const double* v, f3 o32, f3 d32, float tmin, float tmax, float& t_out, float& u_out, float& v_out) {
    const d3 v0 = ldd3(v), e0 = ldd3(v + 3) - v0, e1 = ldd3(v + 6) - v0;
    const d3 o = tod3(o32), d = tod3(d32);
    const d3 pvec = cross(d, e1);
    const double det = dot(e0, pvec);
    if (fabs(det) < 1e-8) return false;
    const double invDet = 1.0 / det;
    const d3 tvec = o - v0;
    const double u = dot(tvec, pvec) * invDet;
    if (u < 0 || u > 1) return false;
    const d3 qvec = cross(tvec, e0);
    const double vv = dot(d, qvec) * invDet;
    if (vv < 0 || (u + vv) > 1) return false;
    const double tl = dot(e1, qvec) * invDet;
    if (tl < (double)tmin || tl > (double)tmax) return false;
    t_out = (float)tl; u_out = (float)u; v_out = (float)vv;
    return true;
}

// EDGE64 (the C-ABI ray query, F_TMIN_F64 builds): an fp32 barycentric within 2e-3 of an edge — one fp32 ulp of a
// vertex 20 units from the origin is 3e-5 of an edge of the 1M-triangle mesh, more for a grazing ray — is decided on the
// fp64 vertices like the reference does, so that the hit id on a shared edge is the reference's.  The integrator keeps
// the fp32 decision: either neighbour of a shared edge is the same surface point.
template <bool EDGE64 = false>
__device__ __forceinline__ bool tri_hit(const GrtTri* tp, const double* v64, const RayD& r, float tmin, float tmax, uint32_t self_id, float& t_out, float& u_out, float& v_out) {
    const float4 a = __ldg((const float4*)tp);        // v0, mat
    const float4 b = __ldg((const float4*)tp + 1);    // e0, id
    const float4 c = __ldg((const float4*)tp + 2);    // e1, flags
    if (__float_as_uint(b.w) == self_id) return false;
    f3 e0 = mk3(b.x, b.y, b.z), e1 = mk3(c.x, c.y, c.z);
    f3 pvec = cross(r.d, e1);
    float det = dot(e0, pvec);
    if (fabsf(det) < 1e-8f) return false;
    float invDet = 1.0f / det;
    f3 tvec = r.o - mk3(a.x, a.y, a.z);
    float u = dot(tvec, pvec) * invDet;
    const float E = EDGE64 ? 2e-3f : 0.0f;
    if (u < -E || u > 1 + E) return false;
    f3 qvec = cross(tvec, e0);
    float v = dot(r.d, qvec) * invDet;
    if (v < -E || (u + v) > 1 + E) return false;
    if (EDGE64) {
        if (v64 && ((u < E) | (v < E) | (u + v > 1 - E))) return tri_hit_f64(v64, r.o, r.d, tmin, tmax, t_out, u_out, v_out);
        if (u < 0 || u > 1 || v < 0 || (u + v) > 1) return false;
    }
    float tl = dot(e1, qvec) * invDet;
    if (tl < tmin || tl > tmax) return false;
    t_out = tl; u_out = u; v_out = v;
    return true;
}

// t of the winning triangle recomputed in fp64 from the fp64 vertices (same expression order as objects.go:409-433).
__device__ __forceinline__ float tri_refine_t(const double* v, const RayD& r, float t32) {
    const d3 v0 = ldd3(v), e0 = ldd3(v + 3) - v0, e1 = ldd3(v + 6) - v0;
    const d3 o = tod3(r.o), d = tod3(r.d);
    const d3 pvec = cross(d, e1);
    const double det = dot(e0, pvec);
    if (fabs(det) < 1e-300) return t32;
    const d3 qvec = cross(o - v0, e0);
    return (float)(dot(e1, qvec) / det);
}

#ifndef GRT_STACK_MAIN
#define GRT_STACK_MAIN 64
#endif
#define GRT_STACK_BOUNDARY 32

// Context for the medium's random draw (medium.go:47): stream (pixel, sample,
// bounce, MEDIUM), draw index = running count of medium tests on this segment.
struct MediumRngCtx {
    uint32_t pixel, sample, bounce, k0, k1, count;
};

// Closest-hit traversal.  Visits children left first, right second
// (bvh.go:73-79); the current closest t plays the role of rayT.Max.  A list
// (HittableList, or a collapsed BVH subtree) is scanned in order in a tight
// loop (hittable.go:129-136); a non-primitive item parks the rest of the list
// on the stack as a continuation entry, so a long list never overflows it.
//
// The traversal is resumable: its whole state (stack, closest candidate so far) lives in a TravState, and
// trav_run<.., VOTE=true> returns, state intact, once most lanes of the warp are done.  The megakernel uses
// that on scenes with a real BVH: the lanes whose rays finished go on to shading and their next ray instead
// of idling until the warp's longest traversal ends (results do not depend on where a traversal is interrupted).
// STRIDE == 1: the stack is a thread-local array.  STRIDE > 1: the stack is this thread's column of a shared-memory
// array [STACK][STRIDE] (`ext` points at row 0): a local-memory stack of 113 664 resident threads does not fit the
// L1/L2, so every pop of a deep traversal was an L2 or DRAM round trip on the critical path.
template <bool BOUNDARY, int STRIDE = 1, int NS = GRT_TRAV_SMEM>
struct TravState {
    static constexpr int STACK = BOUNDARY ? GRT_STACK_BOUNDARY : GRT_STACK_MAIN;
    static constexpr int NSMEM = STRIDE == 1 ? 0 : NS;   // entries held in shared memory; deeper ones overflow
    uint32_t local[STACK - NSMEM > 0 ? STACK - NSMEM : 1];
    uint32_t ext;        // shared-window address of this thread's column (set with set_ext)
    // Explicit ld.shared / st.shared: a reference chosen at run time between the local array and the shared column made
    // every push and pop a GENERIC load / store (the kernels held no LDS / STS at all, cuobjdump round 2), which costs an
    // address-space lookup in the LSU and shows up as long-scoreboard stalls on the stack lines.
    __device__ __forceinline__ void set_ext(const uint32_t* column) { ext = (uint32_t)__cvta_generic_to_shared(column); }
    __device__ __forceinline__ uint32_t get(int i) const {
        if (STRIDE == 1 || i >= NSMEM) return local[i - NSMEM];
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ext + (uint32_t)i * (uint32_t)(STRIDE * 4)) : "memory");
        return v;
    }
    __device__ __forceinline__ void put(int i, uint32_t v) {
        if (STRIDE == 1 || i >= NSMEM) { local[i - NSMEM] = v; return; }
        asm volatile("st.shared.u32 [%0], %1;" : : "r"(ext + (uint32_t)i * (uint32_t)(STRIDE * 4)), "r"(v) : "memory");
    }
    int sp;
    float tmax;          // the closest t so far; plays the role of rayT.Max (bvh.go:78)
    uint32_t ref;        // the closest primitive so far, GRT_REF_NONE: none
    float u, v;
};
template <typename TS>
__device__ __forceinline__ void trav_begin(TS& ts, uint32_t root, float tmax) {
    ts.put(0, root); ts.sp = 1; ts.tmax = tmax; ts.ref = GRT_MAKE_REF(GRT_REF_NONE, 0); ts.u = 0; ts.v = 0;
}

template <uint32_t FEAT, bool BOUNDARY, bool STATS>
__device__ bool closest_hit(const SceneView& sv, uint32_t root, const RayD& r, float tmin, float tmax,
                            uint32_t self_id, uint32_t self_ref, MediumRngCtx* mrng, HitInfo& hit, TraceCounters* tc);

// Self exclusion: the primitive the ray starts on is named by its object id (`self_id`, the C ABI's
// notion) and by its flat ref (`self_ref`).  When every object id maps to one flat primitive the cheap
// ref comparison is used; with F_DUPIDS the id of every candidate is compared.
// Returns true when the traversal is complete.
template <uint32_t FEAT, bool BOUNDARY, bool STATS, bool VOTE, typename TS>
__device__ __forceinline__ bool trav_run(const SceneView& sv, TS& ts, const RayD& r, float tmin,
                                         uint32_t self_id, uint32_t self_ref, MediumRngCtx* mrng, TraceCounters* tc,
                                         unsigned mask, int exit16, int vote16 = 8) {
    int sp = ts.sp;
    float tmax = ts.tmax;
    HitInfo hit;
    hit.ref = ts.ref; hit.u = ts.u; hit.v = ts.v;
    const float4* nodes = sv.nodes();
    const uint32_t excl = BOUNDARY ? GRT_NO_ID : self_id;
    const uint32_t excl_ref = BOUNDARY ? 0xFFFFFFFFu : self_ref;

    // n consecutive primitives starting at `ref`; returns true when `ref` names a primitive type
    auto test_prims = [&](uint32_t ref, uint32_t n) -> bool {
        const uint32_t type = GRT_REF_TYPE(ref), idx = ref & GRT_REF_MASK;
        if ((FEAT & F_QUAD) && type == GRT_REF_QUAD) {
            const DQuadHot* q = sv.quads() + idx;
            if (STATS) tc->quad += n;
            // one quad: branch-free candidate test + predicated update of (tmax, ref)
            auto one = [&](uint32_t k) {
                float t, a, b;
                bool unc;
                bool ok = quad_hit<GRT_PRIM_NC(FEAT)>(q + k, r, tmin, tmax, t, a, b, unc);
                if ((FEAT & F_TMIN_F64) && (FEAT & F_ROTQUAD) && unc) ok = quad_hit_f64(q + k, sv.quads_cold() + idx + k, r, tmin, tmax, t, a, b);
                // a planar primitive cannot be re-hit by a ray leaving it (the fp64 reference finds t ~ 1e-13 < tmin)
                if (FEAT & F_DUPIDS) ok = ok && (sv.quads_cold()[idx + k].id != excl);
                else ok = ok & ((ref + k) != excl_ref);
                if (ok) { tmax = t; hit.ref = ref + k; hit.u = a; hit.v = b; }
            };
            // short runs (every run of a small scene) are straight-line code: no loop counters, and the
            // scheduler can overlap the loads of one quad with the arithmetic of the previous one
            switch (n) {
                case 1: one(0); break;
                case 2: one(0); one(1); break;
                case 3: one(0); one(1); one(2); break;
                case 4: one(0); one(1); one(2); one(3); break;
                case 5: one(0); one(1); one(2); one(3); one(4); break;
                case 6: one(0); one(1); one(2); one(3); one(4); one(5); break;
                default:
#pragma unroll 2
                    for (uint32_t k = 0; k < n; k++) one(k);
            }
            return true;
        }
        if ((FEAT & F_BOX) && type == GRT_REF_BOX) {
            const GrtBox* bx = sv.boxes() + idx;
            if (STATS) tc->box += n;
#pragma unroll 1   // (ptxas otherwise unrolls this by four: +120 instructions in the Cornell megakernel for runs of two boxes, -6 %)
            for (uint32_t k = 0; k < n; k++, bx++) {
                const uint32_t first_quad = ldro<GRT_PRIM_NC(FEAT)>(&bx->first_quad);
                const uint32_t qref = GRT_MAKE_REF(GRT_REF_QUAD, first_quad);
                int excl_face = -1;
                if (FEAT & F_DUPIDS) { for (int f = 0; f < 6; f++) if (sv.quads_cold()[bx->first_quad + f].id == excl) excl_face = f; }
                else if (excl_ref - qref < 6u) excl_face = (int)(excl_ref - qref);
                float t; bool near;
                int face = box_prim_hit<GRT_PRIM_NC(FEAT)>(bx, r, tmin, tmax, excl_face, t, near);
                if ((FEAT & F_TMIN_F64) && near) {
                    // a candidate within fp32 resolution of tmin: decide with the six quads (fp64 plane fallback inside)
                    const DQuadHot* q = sv.quads() + bx->first_quad;
                    for (int f = 0; f < 6; f++) {
                        float tq, a, b; bool unc;
                        bool ok = quad_hit(q + f, r, tmin, tmax, tq, a, b, unc);
                        if (unc) ok = quad_hit_f64(q + f, sv.quads_cold() + bx->first_quad + f, r, tmin, tmax, tq, a, b);
                        if (ok && f != excl_face) { tmax = tq; hit.ref = qref + f; hit.u = a; hit.v = b; }
                    }
                } else if (face >= 0) { tmax = t; hit.ref = qref + (uint32_t)face; hit.u = 0; hit.v = 0; }
            }
            return true;
        }
        if ((FEAT & F_SPHERE) && type == GRT_REF_SPHERE) {
            const GrtSphere* s = sv.spheres() + idx;
            if (STATS) tc->sphere += n;
            for (uint32_t k = 0; k < n; k++, s++) {
                float t;
                if (sphere_hit<GRT_RAY_KEEPS_F64(FEAT), GRT_PRIM_NC(FEAT)>(*s, r, tmin, tmax, ldro<GRT_PRIM_NC(FEAT)>(&s->id) == excl, t)) { tmax = t; hit.ref = ref + k; hit.u = 0; hit.v = 0; }
            }
            return true;
        }
        if ((FEAT & F_TRI) && type == GRT_REF_TRI) {
            if (STATS) tc->tri += n;
            for (uint32_t k = 0; k < n; k++) {
                float t, u, v;
                constexpr bool EDGE64 = (FEAT & F_TMIN_F64) != 0;
                const double* v64 = (EDGE64 && sv.ds->tri_v64) ? sv.ds->tri_v64 + 9 * (size_t)(idx + k) : nullptr;
                if (tri_hit<EDGE64>(sv.ds->tris + idx + k, v64, r, tmin, tmax, excl, t, u, v)) { tmax = t; hit.ref = ref + k; hit.u = u; hit.v = v; }
            }
            return true;
        }
        return type == GRT_REF_QUAD || type == GRT_REF_SPHERE || type == GRT_REF_TRI || type == GRT_REF_BOX || type == GRT_REF_NONE;
    };

    // One item that is not an inner node: a run of primitives named by its parent node, a list (scanned in order), a
    // medium, or a single primitive.  How the primitive runs reach test_prims depends on the variant, and both shapes are
    // measured (profiles/README.md, round 2):
    //  * variants with a BVH or media: ONE call site, inside a loop.  With several, nvcc outlines the lambda in the
    //    all-feature variants and everything it captures by reference (the ray, the scene view, the closest hit) lives
    //    in local memory for the whole kernel (book 2: 560 -> 636 Mpaths/s once it stayed inline; smoke 4620 -> 5290).
    //  * the surface-only list variant (the Cornell box): the list scan with its own call site plus one for a bare
    //    primitive; funnelling both through one loop costs the unrolled 6-quad run its straight-line schedule
    //    (9400 -> 7770).
    auto leaf = [&](uint32_t ref) {
        uint32_t type, idx;
        if constexpr ((FEAT & (F_NODE | F_MEDIUM)) != 0) {
            uint32_t pref = ref, cnt = 1u, lidx = 0u;
            bool in_list = false, last = true;
            if ((FEAT & F_NODE) && (ref & GRT_DREF_RUN_BIT)) {   // (type, count, first index) packed by wide_bvh.hpp
                pref = (ref & (7u << GRT_REF_SHIFT)) | (ref & GRT_DREF_RUN_INDEX_MASK);
                cnt = ((ref >> 25) & 7u) + 1u;
            } else if ((FEAT & F_LIST) && GRT_REF_TYPE(ref) == GRT_REF_LIST) {
                in_list = true;
                lidx = ref & GRT_REF_MASK;
            }
            for (;;) {
                if ((FEAT & F_LIST) && in_list) {
                    const uint2 e = sv.entries()[lidx];
                    last = (e.x & GRT_LIST_LAST) != 0;
                    pref = e.x & ~GRT_LIST_LAST;
                    cnt = e.y;
                }
                if (!test_prims(pref, cnt)) break;        // not a primitive type: handled below
                if (!in_list || last) return;
                lidx++;
            }
            if (in_list && !last) ts.put(sp++, GRT_MAKE_REF(GRT_REF_LIST, lidx + 1));   // the rest of the list, after this item
            ref = pref;
            type = GRT_REF_TYPE(ref);
            idx = ref & GRT_REF_MASK;
            if (type == GRT_REF_LIST || type == GRT_REF_NODE) { ts.put(sp++, ref); return; }   // nested list / a BVH inside a list: next pop
        } else {
            type = GRT_REF_TYPE(ref);
            idx = ref & GRT_REF_MASK;
            if ((FEAT & F_LIST) && type == GRT_REF_LIST) {
                const uint2* entries = sv.entries();
                for (;;) {
                    const uint2 e = entries[idx];
                    const bool last = (e.x & GRT_LIST_LAST) != 0;
                    ref = e.x & ~GRT_LIST_LAST;
                    if (test_prims(ref, e.y)) {
                        if (last) { ref = GRT_MAKE_REF(GRT_REF_NONE, 0); break; }
                        idx++;
                        continue;
                    }
                    if (!last) ts.put(sp++, GRT_MAKE_REF(GRT_REF_LIST, idx + 1));   // the rest of the list, after this item
                    break;
                }
                type = GRT_REF_TYPE(ref);
                idx = ref & GRT_REF_MASK;
                if (type == GRT_REF_LIST || type == GRT_REF_NODE) { ts.put(sp++, ref); return; }   // nested list: next pop
                if (type == GRT_REF_NONE) return;
            }
            if (type != GRT_REF_MEDIUM) { test_prims(ref, 1); return; }   // a bare primitive, or NONE
        }
        if (!BOUNDARY && (FEAT & F_MEDIUM) && type == GRT_REF_MEDIUM) {
            // constantMedium.Hit, medium.go:27-58
            const GrtMedium m = sv.media()[idx];
            if (STATS) tc->medium++;
            HitInfo h1, h2;
            const float INF = __int_as_float(0x7f800000);
            // the two boundary queries of medium.go:32-37: (-inf, inf), then (t1 + 0.0001, inf).  When the boundary is a
            // single sphere or NewBox (every medium of main.go) both answers come from ONE solve — the same floats the two
            // general queries would return, without two nested traversals
            const uint32_t btype = GRT_REF_TYPE(m.boundary), bidx = m.boundary & GRT_REF_MASK;
            if ((FEAT & F_SPHERE) && btype == GRT_REF_SPHERE) {
                if (STATS) tc->sphere += 2;
                float nearr, farr;
                if (!sphere_roots<GRT_RAY_KEEPS_F64(FEAT), GRT_PRIM_NC(FEAT)>(sv.spheres()[bidx], r, false, nearr, farr)) return;
                if (!sphere_pick(nearr, farr, -INF, INF, h1.t)) return;
                if (!sphere_pick(nearr, farr, h1.t + 0.0001f, INF, h2.t)) return;
            } else if ((FEAT & F_BOX) && !(FEAT & F_TMIN_F64) && btype == GRT_REF_BOX) {
                if (STATS) tc->box += 2;
                BoxSlab bs;
                bool nt;
                if (!box_prim_slab<GRT_PRIM_NC(FEAT)>(sv.boxes() + bidx, r, bs)) return;
                if (box_prim_pick(bs, -INF, INF, -1, h1.t, nt) < 0) return;
                if (box_prim_pick(bs, h1.t + 0.0001f, INF, -1, h2.t, nt) < 0) return;
            } else {
                // A general boundary (a BVH or list): two nested queries through closest_hit, which is a real function call
                // here.  It gets COPIES of the view and the ray: handing it references to `sv` and `r` forces both into
                // local memory for the whole kernel — on the book-2 cover, whose boundaries never take this path, the
                // extend kernel read every ray component with LDL (153 M local vs 69 M global load sectors per launch).
                const SceneView svb = sv;
                const RayD rb = r;
                if (!closest_hit<FEAT, true, STATS>(svb, m.boundary, rb, -INF, INF, GRT_NO_ID, 0xFFFFFFFFu, nullptr, h1, tc)) return;
                if (!closest_hit<FEAT, true, STATS>(svb, m.boundary, rb, h1.t + 0.0001f, INF, GRT_NO_ID, 0xFFFFFFFFu, nullptr, h2, tc)) return;
            }
            float t1 = fmaxf(h1.t, tmin), t2 = fminf(h2.t, tmax);
            if (t1 >= t2) return;
            t1 = fmaxf(0.0f, t1);
            float rayLength = sqrtf(len2(r.d));
            float inside = (t2 - t1) * rayLength;
            // medium.go:47 draws only once both boundary hits exist and overlap the interval
            float u = rng_draw(mrng->pixel, mrng->sample, mrng->bounce, GRT_STREAM_MEDIUM, mrng->count++, mrng->k0, mrng->k1);
            float hitDistance = m.neg_inv_density * logf(u);
            if (hitDistance > inside) return;
            float t = t1 + hitDistance / rayLength;
            tmax = t; hit.ref = ref; hit.u = 0; hit.v = 0;
            return;
        }
    };
    // one inner node: four slab tests, then the nearest hit child (or, under a medium, the first in the reference's
    // order) becomes current and the other hit children wait on the stack, nearest on top
    auto node_step = [&](uint32_t& ref) -> bool {   // false: the stack ran dry
        const float4* np = nodes + GRT_WNODE_F4 * (ref & GRT_REF_MASK);
        constexpr bool NC = GRT_NC_NODES != 0 && (FEAT & F_GMEM) != 0;   // 128-bit read-only (ld.global.nc) fetches unless the nodes sit in shared memory
        float4 pnx, pny, pnz, pfx, pfy, pfz;
        uint4 ch;
        uint2 meta;
        const uint32_t ni = ref & GRT_REF_MASK;
        if ((FEAT & F_TREELET) && ni < sv.tl_n) {
            // the top of the tree, staged in shared memory by this block (128-bit ld.shared)
            const uint32_t a = sv.tl + ni * (GRT_WNODE_F4 * 16u);
            auto lds4 = [](uint32_t addr) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v; };
            pnx = lds4(a + 16u * r.nx); pny = lds4(a + 16u * r.ny); pnz = lds4(a + 16u * r.nz);
            pfx = lds4(a + 16u * (3u - r.nx)); pfy = lds4(a + 16u * (5u - r.ny)); pfz = lds4(a + 16u * (7u - r.nz));
            const float4 c4 = lds4(a + 96u), m4 = lds4(a + 112u);
            ch = make_uint4(__float_as_uint(c4.x), __float_as_uint(c4.y), __float_as_uint(c4.z), __float_as_uint(c4.w));
            meta = make_uint2(__float_as_uint(m4.x), __float_as_uint(m4.y));
        } else {
            pnx = ldro<NC>(np + r.nx); pny = ldro<NC>(np + r.ny); pnz = ldro<NC>(np + r.nz);
            pfx = ldro<NC>(np + (3u - r.nx)); pfy = ldro<NC>(np + (5u - r.ny)); pfz = ldro<NC>(np + (7u - r.nz));
            ch = ldro<NC>((const uint4*)(np + 6));
            meta = ldro<NC>((const uint2*)(np + 7));
        }
        if (STATS) tc->box += meta.y;
        const float INF = __int_as_float(0x7f800000);
        const bool ordered = (meta.x & 1u) != 0u;
        // aabb.go:94-110 per child: t = (plane - origin) * (1/d), shrink [tmin, tmax], reject when max < min.
        // The test must be CONSERVATIVE in fp32 (a box may only be rejected if fp64 rejects it too): a leaf box around a
        // planar primitive is ~1e-4 thin, and seen from 1000 units away its two planes round to the same t.  Every t
        // carries a relative error <= 1.5 * 2^-23 (1/d and the product each round once; plane - origin is exact to half
        // an ulp of the result and monotonic), so the far side is widened by 4 * 2^-23 and equality counts as a hit.
#define GRT_SLAB(K, C)                                                                                                          \
        const float tn##K = fmaxf(fmaxf((pnx.C - r.o.x) * r.invd.x, (pny.C - r.o.y) * r.invd.y), fmaxf((pnz.C - r.o.z) * r.invd.z, tmin)); \
        const float tf##K = fminf(fminf(fminf((pfx.C - r.o.x) * r.invd.x, (pfy.C - r.o.y) * r.invd.y), (pfz.C - r.o.z) * r.invd.z) * 1.00000048f, tmax); \
        float k##K = (tn##K <= tf##K) ? (ordered ? tn##K : (float)K) : INF;
        GRT_SLAB(0, x) GRT_SLAB(1, y) GRT_SLAB(2, z) GRT_SLAB(3, w)
#undef GRT_SLAB
        uint32_t c0 = ch.x, c1 = ch.y, c2 = ch.z, c3 = ch.w;
        // sorting network on (key, ref): misses (key = inf) sink to the end
#define GRT_CSWAP(A, B) { const bool sw_ = k##B < k##A; const float lo_ = fminf(k##A, k##B), hi_ = fmaxf(k##A, k##B); \
                          const uint32_t ca_ = sw_ ? c##B : c##A, cb_ = sw_ ? c##A : c##B; k##A = lo_; k##B = hi_; c##A = ca_; c##B = cb_; }
        GRT_CSWAP(0, 1) GRT_CSWAP(2, 3) GRT_CSWAP(0, 2) GRT_CSWAP(1, 3) GRT_CSWAP(1, 2)
#undef GRT_CSWAP
        if (GRT_EAGER_LEAF && (FEAT & F_TRI) && ordered) {
            // Eager leaves: a hit child that is a run of triangles is tested right here, nearest first, instead of
            // waiting on the stack for the warp's next leaf phase (one vote round trip and two shared-memory accesses
            // per leaf saved, and tmax shrinks before the sibling nodes are entered, which are then culled for free).
            // Only where order is free (no medium below); other leaf types keep their own phase.
            // OFF by default: it measured 5 % (mesh) to 13 % (book 2) SLOWER — the triangle test inside the node step runs at
            // a few lanes while the rest of the warp waits, and the four inlined copies cost registers in every variant.
            auto eager = [&](uint32_t c, float k) __attribute__((always_inline)) -> bool {
                if (!(k < INF) || !(c & GRT_DREF_RUN_BIT) || ((c >> GRT_REF_SHIFT) & 7u) != GRT_REF_TRI) return false;
                if (k <= tmax) {
                    const uint32_t first = c & GRT_DREF_RUN_INDEX_MASK, cnt = ((c >> 25) & 7u) + 1u;
                    if (STATS) tc->tri += cnt;
                    for (uint32_t j = 0; j < cnt; j++) {
                        float t, u, v;
                        constexpr bool EDGE64 = (FEAT & F_TMIN_F64) != 0;
                        const double* v64 = (EDGE64 && sv.ds->tri_v64) ? sv.ds->tri_v64 + 9 * (size_t)(first + j) : nullptr;
                        if (tri_hit<EDGE64>(sv.ds->tris + first + j, v64, r, tmin, tmax, excl, t, u, v)) {
                            tmax = t; hit.ref = GRT_MAKE_REF(GRT_REF_TRI, first + j); hit.u = u; hit.v = v;
                        }
                    }
                }
                return true;
            };
            if (eager(c0, k0)) k0 = INF;
            if (eager(c1, k1)) k1 = INF;
            if (eager(c2, k2)) k2 = INF;
            if (eager(c3, k3)) k3 = INF;
            // what is left are inner nodes (and other items) in near-to-far order, minus those behind the new tmax
            k0 = k0 <= tmax ? k0 : INF; k1 = k1 <= tmax ? k1 : INF; k2 = k2 <= tmax ? k2 : INF; k3 = k3 <= tmax ? k3 : INF;
            const bool h0 = k0 < INF, h1 = k1 < INF, h2 = k2 < INF, h3 = k3 < INF;
            if (h3 && (h0 | h1 | h2)) ts.put(sp++, c3);
            if (h2 && (h0 | h1)) ts.put(sp++, c2);
            if (h1 && h0) ts.put(sp++, c1);
            if (h0 | h1 | h2 | h3) { ref = h0 ? c0 : (h1 ? c1 : (h2 ? c2 : c3)); return true; }
        } else if (k0 < INF) {
            if (k3 < INF) ts.put(sp++, c3);
            if (k2 < INF) ts.put(sp++, c2);
            if (k1 < INF) ts.put(sp++, c1);
            ref = c0;
            return true;
        }
        if (sp == 0) { ref = GRT_MAKE_REF(GRT_REF_NONE, 0); return false; }
        ref = ts.get(--sp);
        return true;
    };

    if (!VOTE) {
        while (sp > 0) {
            uint32_t ref = ts.get(--sp);
            if (FEAT & F_NODE) {
                // walk down through inner nodes first ("while-while" traversal)
                while (GRT_REF_TYPE(ref) == GRT_REF_NODE) if (!node_step(ref)) break;
            }
            leaf(ref);
        }
    } else {
        // Warp-synchronous traversal.  The lanes in `mask` (they all enter together) vote on every step: while at
        // least half of the lanes that still have work stand on an inner node, those lanes do one box test each in
        // lock step (vote16/16 instead of half: where the leaves are expensive — fp64 sphere tests, box slabs, media — it
        // pays to let more lanes gather on leaves first, 4/16; on triangle meshes half is best; measured, profiles/README.md);
        // otherwise the lanes standing on a leaf item (list, medium, primitive) process it.  Left to the
        // compiler's reconvergence the lanes of a warp drift apart and run the loop almost serially (measured:
        // 5 of 32 lanes active per instruction on the 1M-triangle mesh).  The slice ends for the whole warp once
        // fewer than exit16/16 of the entering lanes still have work: those keep their state and resume on the
        // next call, next to lanes that have shaded and started a new segment in the meantime.
        const uint32_t NONE = GRT_MAKE_REF(GRT_REF_NONE, 0);
        const int n_enter = __popc(__ballot_sync(mask, sp > 0));
        uint32_t ref = NONE;
        if (sp > 0) ref = ts.get(--sp);
        for (;;) {
            const bool have = GRT_REF_TYPE(ref) != GRT_REF_NONE;
            const bool on_node = (FEAT & F_NODE) && GRT_REF_TYPE(ref) == GRT_REF_NODE;
            const int n_have = __popc(__ballot_sync(mask, have));
            const int n_node = __popc(__ballot_sync(mask, on_node));
            if (n_have * 16 < n_enter * exit16 || n_have == 0) break;
            if (16 * n_node >= vote16 * n_have) {
                if (on_node) node_step(ref);
            } else if (have && !on_node) {
                leaf(ref);
                ref = NONE;
            }
            if (GRT_REF_TYPE(ref) == GRT_REF_NONE && sp > 0) ref = ts.get(--sp);
        }
        if (GRT_REF_TYPE(ref) != GRT_REF_NONE) ts.put(sp++, ref);
    }
    ts.sp = sp; ts.tmax = tmax; ts.ref = hit.ref; ts.u = hit.u; ts.v = hit.v;
    return sp == 0;
}

// Closes a finished traversal: the winning primitive's t is refined in fp64 where fp32 is not enough.
template <uint32_t FEAT, typename TS>
__device__ __forceinline__ bool trav_end(const SceneView& sv, const TS& ts, const RayD& r, HitInfo& hit) {
    hit.ref = ts.ref; hit.u = ts.u; hit.v = ts.v;
    const bool any = hit.ref != GRT_MAKE_REF(GRT_REF_NONE, 0);
    hit.t = ts.tmax;
    if ((FEAT & F_ROTQUAD) && any && GRT_REF_TYPE(hit.ref) == GRT_REF_QUAD) hit.t = quad_refine_t(sv.quads_cold() + (hit.ref & GRT_REF_MASK), r, hit.t);
    if ((FEAT & F_TRI) && any && GRT_REF_TYPE(hit.ref) == GRT_REF_TRI && sv.ds->tri_v64) hit.t = tri_refine_t(sv.ds->tri_v64 + 9 * (size_t)(hit.ref & GRT_REF_MASK), r, hit.t);
    return any;
}

template <uint32_t FEAT, bool BOUNDARY, bool STATS>
__device__ bool closest_hit(const SceneView& sv, uint32_t root, const RayD& r, float tmin, float tmax,
                            uint32_t self_id, uint32_t self_ref, MediumRngCtx* mrng, HitInfo& hit, TraceCounters* tc) {
    TravState<BOUNDARY> ts;
    trav_begin(ts, root, tmax);
    trav_run<FEAT, BOUNDARY, STATS, false>(sv, ts, r, tmin, self_id, self_ref, mrng, tc, 0u, 0);
    return trav_end<FEAT>(sv, ts, r, hit);
}

// ---- full hit record (HitRecord, hittable.go:14-34) ------------------------
struct Surface {
    f3 p, n;            // hit point, face-forwarded normal
    float u, v;
    uint32_t mat, id;
    bool front;
    bool is_surface;    // false for a medium scatter point
    bool planar;
    bool has_onb;       // ou/ov hold the precomputed basis of the FRONT normal (quads)
    f3 ou, ov;
};

__device__ __forceinline__ void set_face_normal(Surface& s, f3 d, f3 outward) {  // hittable.go:27-34
    s.front = dot(d, outward) < 0;
    s.n = s.front ? outward : -outward;
}

// calculateSphereUV on the OBJECT-space outward normal (objects.go:44-50,113);
// uvrot undoes the baked rotateY.
__device__ __forceinline__ void sphere_uv(const GrtSphere& sp, Surface& s) {
    f3 outward = s.front ? s.n : -s.n;
    float c = sp.uvrot[0], sn = sp.uvrot[1];
    f3 on = mk3(c * outward.x - sn * outward.z, outward.y, sn * outward.x + c * outward.z);
    float theta = acosf(fminf(fmaxf(-on.y, -1.0f), 1.0f));
    float phi = atan2f(-on.z, on.x) + GRT_PI_F;
    s.u = phi / (2 * GRT_PI_F);
    s.v = theta / GRT_PI_F;
}

template <uint32_t FEAT>
__device__ __forceinline__ void finish_hit(const SceneView& sv, const RayD& r, const HitInfo& h, bool want_sphere_uv, Surface& s) {
    uint32_t type = GRT_REF_TYPE(h.ref), idx = h.ref & GRT_REF_MASK;
    s.p = mk3(fmaf(h.t, r.d.x, r.o.x), fmaf(h.t, r.d.y, r.o.y), fmaf(h.t, r.d.z, r.o.z));  // Ray.At, ray.go:35
    s.u = h.u; s.v = h.v;
    s.is_surface = true; s.planar = true; s.has_onb = false;
    if ((FEAT & F_QUAD) && type == GRT_REF_QUAD) {
        const float4* c = (const float4*)(sv.quads_cold() + idx);
        const float4 c0 = c[0], c1 = c[1], c2 = c[2];
        s.mat = __float_as_uint(c1.w); s.id = __float_as_uint(c2.w);
        set_face_normal(s, r.d, mk3(c0.x, c0.y, c0.z));
        s.has_onb = true; s.ou = mk3(c1.x, c1.y, c1.z); s.ov = mk3(c2.x, c2.y, c2.z);
        if ((FEAT & F_BOX) && (FEAT & (F_TEXTURE | F_TMIN_F64))) {   // a box hit carries no (alpha, beta): recompute them (objects.go:187-188)
            const float4 A = sv.quads()[idx].A, B = sv.quads()[idx].B;
            s.u = fmaf(A.x, s.p.x, fmaf(A.y, s.p.y, fmaf(A.z, s.p.z, A.w)));
            s.v = fmaf(B.x, s.p.x, fmaf(B.y, s.p.y, fmaf(B.z, s.p.z, B.w)));
        }
        return;
    }
    if ((FEAT & F_SPHERE) && type == GRT_REF_SPHERE) {
        const GrtSphere& sp = sv.spheres()[idx];
        s.mat = sp.mat; s.id = sp.id; s.planar = false;
        // outward = (p - centre)/r with the hit point evaluated in fp64 (objects.go:109-111)
        double cx = sp.c0[0] + (double)r.time * (double)sp.dc[0];
        double cy = sp.c0[1] + (double)r.time * (double)sp.dc[1];
        double cz = sp.c0[2] + (double)r.time * (double)sp.dc[2];
        double t = (double)h.t;
        double inv = 1.0 / sp.r;
        const d3 o64 = GRT_RAY_KEEPS_F64(FEAT) ? r.o64 : tod3(r.o), d64 = GRT_RAY_KEEPS_F64(FEAT) ? r.d64 : tod3(r.d);
        f3 outward = mk3((float)((o64.x + t * d64.x - cx) * inv), (float)((o64.y + t * d64.y - cy) * inv), (float)((o64.z + t * d64.z - cz) * inv));
        set_face_normal(s, r.d, outward);
        if (want_sphere_uv) sphere_uv(sp, s);
        return;
    }
    if ((FEAT & F_TRI) && type == GRT_REF_TRI) {
        const GrtTri* tp = sv.ds->tris + idx;
        const float4 a = __ldg((const float4*)tp), b = __ldg((const float4*)tp + 1), c = __ldg((const float4*)tp + 2);
        s.mat = __float_as_uint(a.w); s.id = __float_as_uint(b.w);
        uint32_t flags = __float_as_uint(c.w);
        f3 nrm;
        if ((FEAT & F_TRISHADE) && (flags & (GRT_TRI_HAS_NORMALS | GRT_TRI_HAS_UV))) {
            const float4* sh = (const float4*)(sv.ds->tri_shade + idx);
            float4 s0 = __ldg(sh), s1 = __ldg(sh + 1), s2 = __ldg(sh + 2), s3 = __ldg(sh + 3);
            float w = 1.0f - h.u - h.v;
            if (flags & GRT_TRI_HAS_UV) {  // objects.go:437-441
                // uv = s2.y.. : n0(3) n1(3) n2(3) uv(6) -> floats 9..14
                float uv0 = s2.y, uv1 = s2.z, uv2 = s2.w, uv3 = s3.x, uv4 = s3.y, uv5 = s3.z;
                s.u = w * uv0 + h.u * uv2 + h.v * uv4;
                s.v = w * uv1 + h.u * uv3 + h.v * uv5;
            }
            if (flags & GRT_TRI_HAS_NORMALS) {  // interpolateNormal, objects.go:389-405
                f3 n0 = mk3(s0.x, s0.y, s0.z), n1 = mk3(s0.w, s1.x, s1.y), n2 = mk3(s1.z, s1.w, s2.x);
                nrm = unit(mk3(w * n0.x + h.u * n1.x + h.v * n2.x, w * n0.y + h.u * n1.y + h.v * n2.y, w * n0.z + h.u * n1.z + h.v * n2.z));
            } else {
                nrm = unit(cross(mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z)));
            }
        } else {
            nrm = unit(cross(mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z)));  // objects.go:268-270
        }
        set_face_normal(s, r.d, nrm);
        return;
    }
    if ((FEAT & F_MEDIUM) && type == GRT_REF_MEDIUM) {  // medium.go:52-56
        const GrtMedium m = sv.media()[idx];
        s.mat = m.mat; s.id = m.id;
        s.n = mk3(1, 0, 0); s.front = true; s.is_surface = false; s.planar = false;
        s.u = 0; s.v = 0;   // the reference leaves u,v stale here
        return;
    }
    s.mat = 0; s.id = GRT_NO_ID; s.n = mk3(0, 1, 0); s.front = true;
}

}  // namespace grtd
