// grt_wavefront.cu — the wavefront variant of the render loop (GRT_VARIANT_WAVEFRONT).
//
// Same arithmetic and the same Philox streams as the megakernel (dev_trace.cuh / dev_shade.cuh),
// different execution shape: a pool of path slots lives in HBM and every bounce is a sequence of
// coherent kernels
//
//   wf_generate   free slots take the next (pixel, stratum) and a camera ray       camera.go:256-290
//   wf_extend     closest hit for every live slot; the slot index is appended to    bvh.go:69, hittable.go:122
//                 the queue of its material class with a warp ballot/popc compaction
//   wf_shade<Q>   one launch per queue: TERMINAL (miss / light: unwind + accumulate), camera.go:301-314
//                 DIFFUSE (Lambertian / Isotropic: PDF mixture), SPECULAR            camera.go:315-330
//
// so that shading never runs with a partially filled warp.  The price is moving ~100 bytes of path
// state through HBM per segment and one atomicAdd per finished path (paths of one pixel are spread
// over the pool), which also makes the fp32 sums order-dependent.  profiles/README.md compares the two
// variants with ncu counters.
#include <cuda_runtime.h>
#include <string.h>
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "dev_shade.cuh"
#include "grt_internal.h"

using namespace grtd;

#define WF_FREE 0xFFFFFFFFu
enum { Q_TERMINAL = 0, Q_DIFFUSE = 1, Q_SPECULAR = 2, Q_COUNT = 3 };
// counters[]: 0..2 queue sizes, 3 live slots after extend, 4 finished flag scratch
enum { C_LIVE = 3, C_NEXT = 5, C_WORDS = 8 };   // C_NEXT: next pool slot handed out by wf_extend_dyn

struct WfParams {
    DevScene scene;
    DevCamera cam;
    uint32_t k0, k1;
    uint32_t sample_first, sample_stride, n_my;
    int x0, y0, ww, wh;
    uint32_t n_pixels;
    uint32_t P;                       // pool size (slots)
    unsigned long long total;         // n_pixels * n_my
    float4* S0;                       // o.xyz, time
    float4* S1;                       // d.xyz, self_ref
    float4* S2;                       // T.xyz, zinfo
    uint4* S3;                        // pixel index, sample, bounce | sp << 8 (WF_FREE = empty), self_id
    float4* H;                        // t, ref, u, v
    float4* rstack;                   // [WEIGHT_STACK][P], depth-major
    uint32_t* queues;                 // [Q_COUNT][P]
    uint32_t* counters;               // C_WORDS
    unsigned long long* next_sample;
    float* rgb_sum;
    int sys_atomics;                  // rgb_sum may live on a peer GPU (grt_render_multi): system-scope atomics
    int exit16;                       // wf_extend_dyn: a traversal slice ends when fewer than exit16/16 lanes have work
    int vote16;                       // wf_extend_dyn: a node step runs while at least vote16/16 of the lanes with work stand on an inner node
    uint32_t treelet_nodes;           // wf_extend_dyn<.., TREELET>: wide nodes staged in shared memory per block
};

template <int STAGED>
__device__ __forceinline__ SceneView wf_view(const WfParams& P, unsigned char* smem) { return make_view<STAGED>(P.scene, smem); }

// ---- generate ----------------------------------------------------------------------------------
template <uint32_t FEAT>
__global__ void __launch_bounds__(256) wf_generate(const __grid_constant__ WfParams P) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    bool want = slot < P.P && P.S3[slot].z == WF_FREE;
    // warp-aggregated claim of sample indices
    unsigned mask = __ballot_sync(FULL, want);
    if (!mask) return;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long base = 0;
    if (lane == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(P.next_sample, (unsigned long long)__popc(mask));
    base = __shfl_sync(FULL, base, __ffs(mask) - 1);
    if (!want) return;
    unsigned long long idx = base + __popc(mask & ((1u << lane) - 1u));
    if (idx >= P.total) return;
    // stratum-major order: consecutive slots get consecutive pixels of one stratum (coherent camera rays)
    uint32_t k = (uint32_t)(idx / P.n_pixels), q = (uint32_t)(idx - (unsigned long long)k * P.n_pixels);
    uint32_t row = q / (uint32_t)P.ww;
    int px = P.x0 + (int)(q - row * (uint32_t)P.ww), py = P.y0 + (int)row;
    uint32_t sample = P.sample_first + k * P.sample_stride;
    uint32_t pixel_index = (uint32_t)(py * P.cam.width + px);
    f3 o, d; float time;
    camera_ray<FEAT>(P.cam, px, py, sample, pixel_index, P.k0, P.k1, o, d, time);
    P.S0[slot] = make_float4(o.x, o.y, o.z, time);
    P.S1[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(0xFFFFFFFFu));
    P.S2[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));
    P.S3[slot] = make_uint4(pixel_index, sample, 0u, GRT_NO_ID);
}

// ---- extend + enqueue ----------------------------------------------------------------------------
#ifndef WF_EXT_MIN_BLOCKS
#define WF_EXT_MIN_BLOCKS 4   /* 64 registers, 32 warps/SM: best of 3/4/5 on the book scenes (C4 342 / 385 / 340 Mpaths/s) */
#endif
// BVH scenes keep the first WF_EXT_SMEM traversal-stack entries of every thread in shared memory (one column per
// thread, conflict-free): as a local array every push / pop was an LDL / STL through the L1 (7 % of this kernel's
// instructions on the book-2 cover, ncu round 1) next to the node and primitive fetches it competes with.
#ifndef WF_EXT_SMEM
#define WF_EXT_SMEM 16
#endif
template <uint32_t FEAT, int STAGED>
__global__ void __launch_bounds__(256, WF_EXT_MIN_BLOCKS) wf_extend(const __grid_constant__ WfParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    SceneView sv = wf_view<STAGED>(P, smem);
    constexpr bool BVH = (FEAT & F_NODE) != 0;
    constexpr uint32_t TF = FEAT | (STAGED == 0 ? F_GMEM : 0u);
    __shared__ uint32_t trav_smem[BVH ? WF_EXT_SMEM : 1][BVH ? 256 : 1];
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    int q = -1;
    if (slot < P.P) {
        const uint4 s3 = P.S3[slot];
        if (s3.z != WF_FREE) {
            const float4 s0 = P.S0[slot], s1 = P.S1[slot];
            RayD ray;
            ray_setup<FEAT>(ray, mk3(s0.x, s0.y, s0.z), mk3(s1.x, s1.y, s1.z), s0.w);
            MediumRngCtx mr;
            mr.pixel = s3.x; mr.sample = s3.y; mr.bounce = s3.z & 255u; mr.k0 = P.k0; mr.k1 = P.k1; mr.count = 0;
            HitInfo h;
            const float INF = __int_as_float(0x7f800000);
            TravState<false, BVH ? 256 : 1, WF_EXT_SMEM> ts;
            ts.set_ext(&trav_smem[0][BVH ? threadIdx.x : 0]);
            trav_begin(ts, P.scene.root, INF);
            trav_run<TF, false, false, false>(sv, ts, ray, 0.001f, s3.w, __float_as_uint(s1.w), &mr, nullptr, 0u, 0);
            const bool hit = trav_end<FEAT>(sv, ts, ray, h);
            if (!hit) { h.t = INF; h.ref = GRT_MAKE_REF(GRT_REF_NONE, 0); h.u = h.v = 0; q = Q_TERMINAL; }
            else {
                uint32_t type = GRT_REF_TYPE(h.ref), idx = h.ref & GRT_REF_MASK, mat;
                if ((FEAT & F_QUAD) && type == GRT_REF_QUAD) mat = sv.quads_cold()[idx].mat;
                else if ((FEAT & F_SPHERE) && type == GRT_REF_SPHERE) mat = sv.spheres()[idx].mat;
                else if ((FEAT & F_TRI) && type == GRT_REF_TRI) mat = P.scene.tris[idx].mat;
                else mat = sv.media()[idx].mat;
                uint32_t mt = sv.materials()[mat].type;
                q = (mt == GRT_MAT_DIFFUSE_LIGHT) ? Q_TERMINAL : ((mt == GRT_MAT_METAL || mt == GRT_MAT_DIELECTRIC) ? Q_SPECULAR : Q_DIFFUSE);
            }
            P.H[slot] = make_float4(h.t, __uint_as_float(h.ref), h.u, h.v);
        }
    }
    // per-material queues, compacted with ballot/popc: one atomicAdd per warp per queue
#pragma unroll
    for (int k = 0; k < Q_COUNT; k++) {
        unsigned m = __ballot_sync(FULL, q == k);
        if (!m) continue;
        uint32_t base = 0;
        int leader = __ffs(m) - 1;
        if ((int)lane == leader) base = atomicAdd(P.counters + k, (uint32_t)__popc(m));
        base = __shfl_sync(FULL, base, leader);
        if (q == k) P.queues[(size_t)k * P.P + base + __popc(m & ((1u << lane) - 1u))] = slot;
    }
    unsigned live = __ballot_sync(FULL, q >= 0);
    if (live && lane == 0) atomicAdd(P.counters + C_LIVE, (uint32_t)__popc(live));
}


// ---- extend for BVH scenes: persistent warps, dynamic ray fetch ----------------------------------
// Traversal lengths on a large BVH differ by orders of magnitude (a ray that misses the mesh is done after one
// node, a grazing one visits hundreds), so one-thread-per-slot leaves most lanes of a warp idle behind its
// longest ray.  Here the warps are persistent: all 32 lanes run warp-synchronous traversal slices
// (trav_run<VOTE>, shared-memory stacks), and whenever a slice ends the lanes whose ray is finished write
// their hit, enqueue the slot and fetch the next live slot from a global counter.  Slots are independent, so
// unlike the megakernel nothing ties a lane to a pixel.
#define WF_DYN_THREADS 128
#ifndef WF_DYN_TREELET
#define WF_DYN_TREELET 0          /* wide nodes (128 B each) staged in shared memory per block for trees of at least ... (0: off — with the
                                     SAH-regrouped tree 0 / 32 / 64 nodes measure 544 / 531 / 526 Mpaths/s on the mesh; GRT_WF_TREELET=n turns it on) */
#endif
#ifndef WF_DYN_TREELET_MIN_NODES
#define WF_DYN_TREELET_MIN_NODES 16384u   /* ... this many wide nodes (smaller trees are L1-resident: staging only shrinks the L1) */
#endif
#ifndef WF_DYN_SMEM
#define WF_DYN_SMEM 24   /* traversal-stack entries per thread in shared memory (deeper ones in local memory): 16 / 24 / 32 measure the same */
#endif
#ifndef WF_DYN_MIN_BLOCKS
#define WF_DYN_MIN_BLOCKS 7   /* 72 registers, 28 warps/SM: measured best of 6/7/8 with the 4-wide BVH (430 / 462 / 442 Mpaths/s on the 1M-triangle mesh) */
#endif
template <uint32_t FEAT, int STAGED, bool TREELET>
__global__ void __launch_bounds__(WF_DYN_THREADS, WF_DYN_MIN_BLOCKS) wf_extend_dyn(const __grid_constant__ WfParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    SceneView sv = wf_view<STAGED>(P, smem);
    constexpr uint32_t TF = FEAT | (STAGED == 0 ? F_GMEM : 0u) | (TREELET ? F_TREELET : 0u);
    __shared__ uint32_t trav_smem[WF_DYN_SMEM][WF_DYN_THREADS];
    // Top-of-tree nodes in shared memory (TREELET builds, large trees only): the wide nodes are numbered breadth first, so
    // the first P.treelet_nodes of them are the levels every ray walks through; the block copies them once into its
    // dynamic shared memory and node_step fetches them with ld.shared.  Measured (profiles/README.md, round 2): +1-2 % on
    // the 1M-triangle mesh with 32 nodes, -3 ... -10 % on the small book scenes at any size (their whole tree already
    // lives in the L1, and every KB of shared memory is a KB less of it), so small trees run the build without it.
    if (TREELET) {
        float4* treelet = (float4*)smem;
        const uint32_t n = P.treelet_nodes * GRT_WNODE_F4;
        const float4* src = sv.nodes();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) treelet[i] = __ldg(src + i);
        __syncthreads();
        sv.tl = (uint32_t)__cvta_generic_to_shared(treelet);
        sv.tl_n = P.treelet_nodes;
    }
    TravState<false, WF_DYN_THREADS, WF_DYN_SMEM> ts;
    ts.set_ext(&trav_smem[0][threadIdx.x]);
    ts.sp = 0;
    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const float INF = __int_as_float(0x7f800000);
    bool have = false, exhausted = false;
    uint32_t slot = 0, self_ref = 0xFFFFFFFFu, med_count = 0;
    uint4 s3 = make_uint4(0, 0, 0, GRT_NO_ID);
    RayD ray;
    ray_setup<FEAT>(ray, mk3(0, 0, 0), mk3(0, 0, 1), 0.0f);
    // Every pass of the loop: (1) lanes whose ray finished in the last slice hand their slot to the queue of its
    // material class, (2) idle lanes take the next slots of the pool, (3) all lanes run one traversal slice.  The up to
    // four counters involved (three queues + the pool cursor) are bumped by lane 0 back to back, so the warp waits for
    // ONE atomic round trip per pass instead of four dependent ones (8.5 % of this kernel's stall samples, ncu round 2).
    int q = -1;                 // queue of the ray this lane just finished (-1: none)
    uint32_t done_slot = 0;
    for (;;) {
        const unsigned m0 = __ballot_sync(FULL, q == 0), m1 = __ballot_sync(FULL, q == 1), m2 = __ballot_sync(FULL, q == 2);
        const unsigned need = exhausted ? 0u : __ballot_sync(FULL, !have);
        uint32_t b0 = 0, b1 = 0, b2 = 0, bn = 0;
        if (lane == 0) {
            if (m0) b0 = atomicAdd(P.counters + 0, (uint32_t)__popc(m0));
            if (m1) b1 = atomicAdd(P.counters + 1, (uint32_t)__popc(m1));
            if (m2) b2 = atomicAdd(P.counters + 2, (uint32_t)__popc(m2));
            if (need) bn = atomicAdd(P.counters + C_NEXT, (uint32_t)__popc(need));
            if (m0 | m1 | m2) atomicAdd(P.counters + C_LIVE, (uint32_t)__popc(m0 | m1 | m2));
        }
        b0 = __shfl_sync(FULL, b0, 0); b1 = __shfl_sync(FULL, b1, 0); b2 = __shfl_sync(FULL, b2, 0); bn = __shfl_sync(FULL, bn, 0);
        if (q >= 0) {
            const unsigned m = q == 0 ? m0 : (q == 1 ? m1 : m2);
            const uint32_t base = q == 0 ? b0 : (q == 1 ? b1 : b2);
            P.queues[(size_t)q * P.P + base + __popc(m & lt_mask)] = done_slot;
            q = -1;
        }
        if (need) {
            if (bn >= P.P) exhausted = true;
            if (!have) {
                const uint32_t sl = bn + (uint32_t)__popc(need & lt_mask);
                if (sl < P.P) {
                    const uint4 v = P.S3[sl];
                    if (v.z != WF_FREE) {
                        slot = sl; s3 = v;
                        const float4 s0 = P.S0[sl], s1 = P.S1[sl];
                        ray_setup<FEAT>(ray, mk3(s0.x, s0.y, s0.z), mk3(s1.x, s1.y, s1.z), s0.w);
                        self_ref = __float_as_uint(s1.w);
                        trav_begin(ts, P.scene.root, INF);
                        med_count = 0;
                        have = true;
                    }
                }
            }
        }
        if (!__any_sync(FULL, have)) { if (exhausted) break; else continue; }
        MediumRngCtx mr;
        mr.pixel = s3.x; mr.sample = s3.y; mr.bounce = s3.z & 255u; mr.k0 = P.k0; mr.k1 = P.k1; mr.count = med_count;
        const bool fin = trav_run<TF, false, false, true>(sv, ts, ray, 0.001f, s3.w, self_ref, &mr, nullptr, FULL, P.exit16, P.vote16);
        med_count = mr.count;
        if (have && fin) {
            HitInfo h;
            const bool hit = trav_end<FEAT>(sv, ts, ray, h);
            if (!hit) { h.t = INF; h.ref = GRT_MAKE_REF(GRT_REF_NONE, 0); h.u = h.v = 0; q = Q_TERMINAL; }
            else {
                uint32_t type = GRT_REF_TYPE(h.ref), idx = h.ref & GRT_REF_MASK, mat;
                if ((FEAT & F_QUAD) && type == GRT_REF_QUAD) mat = sv.quads_cold()[idx].mat;
                else if ((FEAT & F_SPHERE) && type == GRT_REF_SPHERE) mat = sv.spheres()[idx].mat;
                else if ((FEAT & F_TRI) && type == GRT_REF_TRI) mat = P.scene.tris[idx].mat;
                else mat = sv.media()[idx].mat;
                uint32_t mt = sv.materials()[mat].type;
                q = (mt == GRT_MAT_DIFFUSE_LIGHT) ? Q_TERMINAL : ((mt == GRT_MAT_METAL || mt == GRT_MAT_DIELECTRIC) ? Q_SPECULAR : Q_DIFFUSE);
            }
            P.H[slot] = make_float4(h.t, __uint_as_float(h.ref), h.u, h.v);
            done_slot = slot;
            have = false;
        }
    }
    // (the loop only ends on a pass that found no lane with work, and such a pass has already flushed every pending q)
}

// ---- shade ------------------------------------------------------------------------------------------
__device__ __forceinline__ void wf_finish_path(const WfParams& P, uint32_t slot, uint32_t pixel_index, f3 L) {
    if (L.x != 0.0f || L.y != 0.0f || L.z != 0.0f) {   // NaN != 0 is true: a NaN sample is accumulated (color.go:28-36)
        float* dst = P.rgb_sum + (size_t)pixel_index * 3;
        if (P.sys_atomics) { atomicAdd_system(dst, L.x); atomicAdd_system(dst + 1, L.y); atomicAdd_system(dst + 2, L.z); }
        else { atomicAdd(dst, L.x); atomicAdd(dst + 1, L.y); atomicAdd(dst + 2, L.z); }
    }
    P.S3[slot].z = WF_FREE;
}

template <uint32_t FEAT, int STAGED, int Q>
__global__ void __launch_bounds__(256) wf_shade(const __grid_constant__ WfParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    SceneView sv = wf_view<STAGED>(P, smem);
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P.counters[Q]) return;
    const uint32_t slot = P.queues[(size_t)Q * P.P + j];
    const uint4 s3 = P.S3[slot];
    const float4 s0 = P.S0[slot], s1 = P.S1[slot], s2 = P.S2[slot], hh = P.H[slot];
    const uint32_t bounce = s3.z & 255u;
    int sp = (int)(s3.z >> 8);
    f3 T = mk3(s2.x, s2.y, s2.z);
    uint32_t zinfo = __float_as_uint(s2.w);
    const float4* rs = P.rstack + slot;
    RayD ray;
    ray_setup<FEAT>(ray, mk3(s0.x, s0.y, s0.z), mk3(s1.x, s1.y, s1.z), s0.w);
    HitInfo h;
    h.t = hh.x; h.ref = __float_as_uint(hh.y); h.u = hh.z; h.v = hh.w;

    if (Q == Q_TERMINAL) {
        f3 E;
        if (GRT_REF_TYPE(h.ref) == GRT_REF_NONE) E = P.cam.background;          // camera.go:301
        else {
            Surface s;
            finish_hit<FEAT>(sv, ray, h, false, s);
            const GrtMaterial mat = sv.materials()[s.mat];
            if ((FEAT & F_SPHERE) && (FEAT & F_TEXTURE) && GRT_REF_TYPE(h.ref) == GRT_REF_SPHERE && material_needs_uv<FEAT>(sv, mat))
                sphere_uv(sv.spheres()[h.ref & GRT_REF_MASK], s);
            E = s.front ? texture_value<FEAT>(sv, mat.tex, s.u, s.v, s.p) : mk3(0, 0, 0);   // materials.go:150-155
        }
        f3 L = mk3(0, 0, 0);
        if (E.x != 0.0f || E.y != 0.0f || E.z != 0.0f) L = unwind_clamp(T, zinfo, E, rs, sp, P.cam.max_contribution, P.P);
        wf_finish_path(P, slot, s3.x, L);
        return;
    }
    Surface s;
    finish_hit<FEAT>(sv, ray, h, false, s);
    const GrtMaterial mat = sv.materials()[s.mat];
    if ((FEAT & F_SPHERE) && (FEAT & F_TEXTURE) && GRT_REF_TYPE(h.ref) == GRT_REF_SPHERE && material_needs_uv<FEAT>(sv, mat))
        sphere_uv(sv.spheres()[h.ref & GRT_REF_MASK], s);
    ShadeResult R = shade_vertex<FEAT>(sv, ray, s, mat, s3.x, s3.y, bounce, P.k0, P.k1, nullptr);
    if (R.kind == SHADE_NAN) { const float qn = __int_as_float(0x7fc00000); wf_finish_path(P, slot, s3.x, mk3(qn, qn, qn)); return; }
    if (R.kind == SHADE_TERMINATE) {   // not reached for Q_DIFFUSE / Q_SPECULAR materials
        wf_finish_path(P, slot, s3.x, mk3(0, 0, 0));
        return;
    }
    if (R.kind == SHADE_SPECULAR) apply_factor(T, zinfo, R.value, sp);
    else {
        if (R.value.x == 0.0f && R.value.y == 0.0f && R.value.z == 0.0f) { wf_finish_path(P, slot, s3.x, mk3(0, 0, 0)); return; }
        P.rstack[(size_t)sp * P.P + slot] = recip_factor(T);
        sp++;
        apply_factor(T, zinfo, R.value, sp);
    }
    if ((int)bounce + 1 > P.cam.max_depth) { wf_finish_path(P, slot, s3.x, mk3(0, 0, 0)); return; }   // camera.go:294-296
    P.S0[slot] = make_float4(s.p.x, s.p.y, s.p.z, s0.w);
    P.S1[slot] = make_float4(R.dir.x, R.dir.y, R.dir.z, __uint_as_float(s.is_surface ? h.ref : 0xFFFFFFFFu));
    P.S2[slot] = make_float4(T.x, T.y, T.z, __uint_as_float(zinfo));
    P.S3[slot] = make_uint4(s3.x, s3.y, (bounce + 1u) | ((uint32_t)sp << 8), s.is_surface ? s.id : GRT_NO_ID);
}

__global__ void wf_init_slots(uint4* S3, uint32_t P) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) S3[i] = make_uint4(0, 0, WF_FREE, GRT_NO_ID);
}

// ---- host driver ---------------------------------------------------------------------------------------
static thread_local GrtTiming g_timing;
extern "C" int grt_last_timing(GrtTiming* out) {
    if (!out) { grt_set_error("grt_last_timing: NULL argument"); return GRT_E_INVALID; }
    *out = g_timing;
    return GRT_OK;
}
void grt_internal_set_timing(const GrtTiming& t) { g_timing = t; }

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { grt_set_error(std::string("wavefront: ") + #call + ": " + cudaGetErrorString(e_)); rc = GRT_E_CUDA; goto done; } \
    } while (0)

template <uint32_t FEAT>
static int wf_run(GrtSceneHandle h, WfParams& P, cudaStream_t st, uint32_t* h_counters, bool timing) {
    int rc = GRT_OK;
    // Persistent extend with dynamic ray fetch (wf_extend_dyn) vs one thread per slot (wf_extend), measured with the 4-wide
    // BVH, one primitive per leaf and lists as wide nodes (profiles/README.md, round 2): triangle meshes 462 vs 232
    // Mpaths/s, the sphere BVH of book 1 796 vs 696, the book-2 cover 538 vs 510 (round 1's binary tree with 4-primitive
    // leaves had book 2 the other way round: 252 vs 311).
    // Since the medium-free subtrees are regrouped by SAH (wide_bvh.hpp) book 1 visits ~7 nodes of a 140-node tree per ray:
    // that small a scene sits in shared memory whole and the plain kernel wins again (1662 vs 1547), book 2 (521 nodes, two
    // media) stays with the persistent one (743 vs 657).
    constexpr bool can_dyn = (FEAT & F_NODE) != 0;
    bool dyn = can_dyn && !(P.scene.n_nodes < 256u && P.scene.n_media == 0u);
    if (const char* e = getenv("GRT_WF_DYN")) dyn = can_dyn && (atoi(e) == 2 || (dyn && atoi(e) != 0));   // 0: never, 2: whenever compiled in
    const bool staged = grt_internal_staged(h) == 2 && !dyn;   // whole-blob staging or none
    {
        uint32_t tl = WF_DYN_TREELET;
        if (const char* e = getenv("GRT_WF_TREELET")) tl = (uint32_t)atoi(e);   // (A/B knob: nodes per block, 0 = off)
        else if (P.scene.n_nodes < WF_DYN_TREELET_MIN_NODES) tl = 0;
        P.treelet_nodes = dyn ? (tl < P.scene.n_nodes ? tl : P.scene.n_nodes) : 0u;
        if (P.treelet_nodes > 256u) P.treelet_nodes = 256u;
    }
    const unsigned dyn_blocks = (unsigned)grt_internal_sm_count(h) * WF_DYN_MIN_BLOCKS;
    const size_t smem = staged ? P.scene.stage_bytes : 0;
    const unsigned blocks = (P.P + 255) / 256;
    const bool has_spec = (P.scene.features & F_SPECULAR) != 0;
    uint64_t launches = 0;
    if (smem) {
        cudaFuncSetAttribute(wf_extend<FEAT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(wf_shade<FEAT, 2, Q_TERMINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(wf_shade<FEAT, 2, Q_DIFFUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(wf_shade<FEAT, 2, Q_SPECULAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    // one bounce: reset the queue counters, refill free slots, extend, shade the three queues
    std::vector<cudaEvent_t> tev;   // GRT_OPT_TIMING: four events per bounce (before generate / extend / shade, after shade)
    auto mark = [&](cudaStream_t s_) {
        if (!timing) return;
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, s_); tev.push_back(e); }
    };
    auto one_iteration = [&](cudaStream_t s_) -> cudaError_t {
        cudaError_t e = cudaMemsetAsync(P.counters, 0, C_WORDS * 4, s_);
        if (e != cudaSuccess) return e;
        mark(s_);
        wf_generate<FEAT><<<blocks, 256, 0, s_>>>(P);
        mark(s_);
        if (dyn) {
            if constexpr (can_dyn) {
                // the scene is read from global memory (L1/L2): the shared memory holds the traversal stacks
                if (P.treelet_nodes) wf_extend_dyn<FEAT, 0, true><<<dyn_blocks, WF_DYN_THREADS, P.treelet_nodes * GRT_WNODE_F4 * 16, s_>>>(P);
                else wf_extend_dyn<FEAT, 0, false><<<dyn_blocks, WF_DYN_THREADS, 0, s_>>>(P);
                mark(s_);
                wf_shade<FEAT, 0, Q_TERMINAL><<<blocks, 256, 0, s_>>>(P);
                wf_shade<FEAT, 0, Q_DIFFUSE><<<blocks, 256, 0, s_>>>(P);
                if (has_spec) wf_shade<FEAT, 0, Q_SPECULAR><<<blocks, 256, 0, s_>>>(P);
            }
        } else if (staged) {
            wf_extend<FEAT, 2><<<blocks, 256, smem, s_>>>(P);
            mark(s_);
            wf_shade<FEAT, 2, Q_TERMINAL><<<blocks, 256, smem, s_>>>(P);
            wf_shade<FEAT, 2, Q_DIFFUSE><<<blocks, 256, smem, s_>>>(P);
            if (has_spec) wf_shade<FEAT, 2, Q_SPECULAR><<<blocks, 256, smem, s_>>>(P);
        } else {
            wf_extend<FEAT, 0><<<blocks, 256, 0, s_>>>(P);
            mark(s_);
            wf_shade<FEAT, 0, Q_TERMINAL><<<blocks, 256, 0, s_>>>(P);
            wf_shade<FEAT, 0, Q_DIFFUSE><<<blocks, 256, 0, s_>>>(P);
            if (has_spec) wf_shade<FEAT, 0, Q_SPECULAR><<<blocks, 256, 0, s_>>>(P);
        }
        mark(s_);
        return cudaSuccess;
    };
    const uint64_t per_iter = has_spec ? 5 : 4;
    // The bounce loop is launch-bound on small frames (five short kernels and a memset per bounce), so eight bounces
    // plus the read-back of the live-slot counter are captured once into a CUDA graph and replayed: one launch per
    // eight bounces.  Capture needs a real stream (the caller's may be the legacy default stream), so the loop runs on
    // a private stream ordered after, and waited for by, the caller's stream.  GRT_WF_GRAPH=0: plain launches.
    // (Measured: +0.5 % — the launches were already hidden behind the running kernels; kept because it removes 47 of
    // every 48 host calls from the render thread.)
    static const bool graph_env = [] { const char* e = getenv("GRT_WF_GRAPH"); return !(e && atoi(e) == 0); }();
    const bool use_graph = graph_env && !timing;   // events inside a captured graph cannot be read back: plain launches when timing
    cudaStream_t ws = nullptr;
    cudaEvent_t e_in = nullptr, e_out = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (use_graph) {
        CU(cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&e_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&e_out, cudaEventDisableTiming));
        CU(cudaEventRecord(e_in, st));
        CU(cudaStreamWaitEvent(ws, e_in, 0));
    }
    {
        cudaStream_t s0 = use_graph ? ws : st;
        wf_init_slots<<<blocks, 256, 0, s0>>>(P.S3, P.P);
        launches++;
        CU(cudaMemsetAsync(P.next_sample, 0, 8, s0));
    }
    if (use_graph) {
        const int BOUNCES = 8;
        CU(cudaStreamBeginCapture(ws, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < BOUNCES; k++) CU(one_iteration(ws));
        CU(cudaMemcpyAsync(h_counters, P.counters, C_WORDS * 4, cudaMemcpyDeviceToHost, ws));
        CU(cudaStreamEndCapture(ws, &graph));
        CU(cudaGraphInstantiate(&exec, graph, 0));
        for (uint64_t rounds = 0;; rounds++) {
            CU(cudaGraphLaunch(exec, ws));
            CU(cudaStreamSynchronize(ws));
            launches += per_iter * BOUNCES;
            if (h_counters[C_LIVE] == 0) break;
            if (rounds > (1ull << 36)) { grt_set_error("wavefront: did not drain"); rc = GRT_E_CUDA; goto done; }
        }
        CU(cudaEventRecord(e_out, ws));
        CU(cudaStreamWaitEvent(st, e_out, 0));
    } else {
        for (uint64_t iter = 0;; iter++) {
            CU(one_iteration(st));
            launches += per_iter;
            if ((iter & 7u) == 7u || iter < 2) {   // the host only needs to know when the pool has drained
                CU(cudaMemcpyAsync(h_counters, P.counters, C_WORDS * 4, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (h_counters[C_LIVE] == 0) break;
            }
            if (iter > (1ull << 40)) { grt_set_error("wavefront: did not drain"); rc = GRT_E_CUDA; goto done; }
        }
    }
    CU(cudaGetLastError());
    if (timing) {
        CU(cudaStreamSynchronize(st));
        GrtTiming T;
        memset(&T, 0, sizeof(T));
        for (size_t k = 0; k + 3 < tev.size(); k += 4) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tev[k], tev[k + 1]);
            cudaEventElapsedTime(&b, tev[k + 1], tev[k + 2]);
            cudaEventElapsedTime(&c, tev[k + 2], tev[k + 3]);
            T.generate_ms += a; T.extend_ms += b; T.shade_ms += c;
            T.extend_launches++;
        }
        if (tev.size() >= 4) { float t = 0; cudaEventElapsedTime(&t, tev.front(), tev.back()); T.total_ms = t; }
        T.launches = launches;
        T.extend_kernel = dyn ? 2u : 1u;
        grt_internal_set_timing(T);
    }
done:
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (e_in) cudaEventDestroy(e_in);
    if (e_out) cudaEventDestroy(e_out);
    if (ws) { cudaStreamSynchronize(ws); cudaStreamDestroy(ws); }
    grt_count_launch(launches);
    if (getenv("GRT_WF_TRACE")) fprintf(stderr, "[wavefront] %llu launches, dyn=%d, pool=%u slots\n", (unsigned long long)launches, (int)dyn, P.P);
    return rc;
}

int grt_render_wavefront(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt, float* d_rgb_sum, cudaStream_t st, GrtStats* d_stats) {
    (void)d_stats;
    int rc = GRT_OK;
    WfParams P;
    memset(&P, 0, sizeof(P));
    if ((rc = grt_make_dev_camera(cam, &P.cam))) return rc;
    P.scene = *grt_internal_dev_scene(h);
    P.k0 = (uint32_t)opt->seed; P.k1 = (uint32_t)(opt->seed >> 32);
    const uint32_t S2 = (uint32_t)cam->spp_sqrt * (uint32_t)cam->spp_sqrt;
    const uint32_t stride = opt->sample_stride ? opt->sample_stride : 1u;
    if (opt->sample_first >= S2) return GRT_OK;
    P.sample_first = opt->sample_first; P.sample_stride = stride;
    P.n_my = (S2 - opt->sample_first + stride - 1) / stride;
    int x0 = opt->x0, y0 = opt->y0, x1 = opt->x1, y1 = opt->y1;
    if (x0 == 0 && y0 == 0 && x1 == 0 && y1 == 0) { x1 = cam->width; y1 = cam->height; }
    if (x0 < 0 || y0 < 0 || x1 > cam->width || y1 > cam->height || x1 <= x0 || y1 <= y0) { grt_set_error("GrtOptions pixel window out of range"); return GRT_E_INVALID; }
    P.x0 = x0; P.y0 = y0; P.ww = x1 - x0; P.wh = y1 - y0;
    P.n_pixels = (uint32_t)P.ww * (uint32_t)P.wh;
    P.total = (unsigned long long)P.n_pixels * P.n_my;
    // Pool size: 2 Mi slots (~2.4 GB of state) for big frames — 1 / 2 / 4 / 8 Mi measure 426 / 468 / 471 / 433 Mpaths/s on
    // the 1M-triangle mesh — and a sixteenth of the frame's paths for small ones, so that the drain of the last pool (a
    // few long paths in an almost empty pool) stays a small part of the frame: book 1 at its shipped 400x225x100
    // (9 M paths) renders at 561 Mpaths/s with 512 Ki slots, 497 with 2 Mi, 401 with 4 Mi.
    unsigned long long pool_slots = P.total / 16;
    if (pool_slots < (1ull << 18)) pool_slots = 1ull << 18;
    if (pool_slots > (1ull << 21)) pool_slots = 1ull << 21;
    if (const char* e = getenv("GRT_WF_SLOTS")) { unsigned long long v = strtoull(e, nullptr, 10); if (v >= 256) pool_slots = v; }
    unsigned long long want = P.total < pool_slots ? P.total : pool_slots;
    P.P = (uint32_t)((want + 255) / 256 * 256);
    P.rgb_sum = d_rgb_sum;
    P.sys_atomics = (opt->flags & GRT_OPT_ATOMIC_SUM) ? 1 : 0;
    // Slice exit: the warp leaves a traversal slice (to flush finished rays and fetch new ones) once fewer than exit16/16
    // of the lanes that entered still have work.  A refill costs a pass of atomics, ray fetches and stack resets, so it
    // only pays where traversal lengths differ a lot: the 1M-triangle mesh 9/16 (2 / 4 / 6 / 8 / 9 / 10 / 12: 445 / 484 /
    // 509 / 518 / 522 / 519 / 497 Mpaths/s), book 2 with its media 3/16 (0 / 2 / 3 / 4 / 8 / 12: 718 / 737 / 743 / 745 /
    // 720 / 658); book 1's 140-node tree would run every ray of the warp to its end (0 / 2 / 4 / 8 / 12: 1544 / 1470 / 1409 /
    // 1303 / 1079) but takes the plain kernel anyway (wf_run).
    // Meshes of every size want the refill (2 k / 32 k triangles: 659 / 454 Mpaths/s with 0, 749 / 531 with 6, 708 / 541 with 9).
    P.exit16 = (P.scene.n_tris > P.scene.n_spheres + P.scene.n_boxes) ? 9 : 3;
    if (const char* e = getenv("GRT_WF_EXIT16")) { int v = atoi(e); if (v >= 0 && v <= 16) P.exit16 = v; }
    // node-step quorum: triangle meshes 10/16, scenes whose leaves are spheres / boxes / media 4/16
    // (mesh, 6 / 8 / 10 / 12 / 14 sixteenths: 514 / 520 / 530 / 522 / 499; book 2 on its final tree, 1 / 2 / 4: 803 / 819 / 833;
    // on the earlier trees 2 / 4 / 8: 660 / 653 / 634)
    P.vote16 = (P.scene.n_tris > P.scene.n_spheres + P.scene.n_boxes) ? 10 : 4;
    if (const char* e = getenv("GRT_WF_VOTE16")) { int v = atoi(e); if (v >= 1 && v <= 16) P.vote16 = v; }
    void* pool = nullptr;
    uint32_t* h_counters = nullptr;
    const bool timing = (opt->flags & GRT_OPT_TIMING) != 0;
    const bool trace = getenv("GRT_WF_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_a = now(), t_b = 0, t_c = 0;
    {
        size_t n = P.P;
        size_t bytes = n * 16 * 5 + n * 16 * (size_t)(cam->max_depth + 1) + n * 4 * Q_COUNT + C_WORDS * 4 + 64;
        void* pinned = nullptr;
        pool = grt_internal_wf_pool(h, bytes, &pinned);
        if (!pool) { rc = GRT_E_CUDA; goto done; }   // (grt_internal_wf_pool has set the error text)
        h_counters = (uint32_t*)pinned;
        unsigned char* p = (unsigned char*)pool;
        P.S0 = (float4*)p; p += n * 16; P.S1 = (float4*)p; p += n * 16; P.S2 = (float4*)p; p += n * 16;
        P.S3 = (uint4*)p; p += n * 16; P.H = (float4*)p; p += n * 16;
        P.rstack = (float4*)p; p += n * 16 * (size_t)(cam->max_depth + 1);
        P.queues = (uint32_t*)p; p += n * 4 * Q_COUNT;
        P.counters = (uint32_t*)p; p += C_WORDS * 4;
        p = (unsigned char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
        P.next_sample = (unsigned long long*)p;
    }
    t_b = now();
    {
        uint32_t f = (P.scene.features & ~F_DUPIDS) | (cam->defocus_angle > 0 ? F_DEFOCUS : 0u);
        bool dup = (P.scene.features & F_DUPIDS) != 0;
        if (!dup && (f & ~V_CORNELL) == 0) rc = wf_run<V_CORNELL>(h, P, st, h_counters, timing);
        else if (!dup && (f & ~V_SMOKE) == 0) rc = wf_run<V_SMOKE>(h, P, st, h_counters, timing);
        else if (!dup && (f & ~V_SPHERES) == 0) rc = wf_run<V_SPHERES>(h, P, st, h_counters, timing);
        else if (!dup && (f & ~V_MESH) == 0) rc = wf_run<V_MESH>(h, P, st, h_counters, timing);
        else if (!dup) rc = wf_run<V_FULL_UNIQ>(h, P, st, h_counters, timing);
        else rc = wf_run<F_ALL>(h, P, st, h_counters, timing);
    }
done:
    if (pool) { cudaStreamSynchronize(st); t_c = now(); }   // the pool stays with the scene handle
    if (trace) fprintf(stderr, "[wavefront] alloc %.1f ms, render %.1f ms, free %.1f ms\n", t_b - t_a, t_c - t_b, now() - t_c);
    return rc;
}
