// grt_kernels.cu — sm_100a kernels and the C ABI of include/grt.h.
//
// Kernels:
//   trace_batch_kernel   closest hit for a ray batch (parity checks 1 and 2)
//   render_mega_kernel   persistent megakernel: camera ray -> loop { BVH
//                        traversal, intersection, BSDF scatter, light/cosine
//                        PDF mixture } -> recursive-clamp unwind -> per-pixel sum
//   tonemap_kernel       color.go:14-46
// The wavefront variant lives in grt_wavefront.cu.
//
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <algorithm>
#include <math.h>
#include <limits>

#include "dev_shade.cuh"
#include "grt_internal.h"
#include "wide_bvh.hpp"

using namespace grtd;

// ===========================================================================
// error plumbing
// ===========================================================================
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches(0);

void grt_set_error(const std::string& s) { g_last_error = s; }

cudaError_t grt_dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    static std::mutex m;
    static bool tuned[64] = {false};
    {
        std::lock_guard<std::mutex> lk(m);
        if (dev < 64 && !tuned[dev]) {   // keep freed blocks in the pool instead of returning them to the driver
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            tuned[dev] = true;
        }
    }
    if ((e = cudaMallocAsync(p, bytes ? bytes : 1, 0)) != cudaSuccess) return e;
    return cudaStreamSynchronize(0);
}
void grt_dev_free(void* p) {
    if (!p) return;
    cudaDeviceSynchronize();
    cudaFreeAsync(p, 0);
}
void grt_count_launch(uint64_t n) { g_launches += n; }

#define CUDA_TRY(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            grt_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
            return GRT_E_CUDA;                                                                  \
        }                                                                                       \
    } while (0)

// ===========================================================================
// trace_batch: one thread per ray
// ===========================================================================
template <uint32_t FEAT, int STAGED>
__global__ void __launch_bounds__(128) trace_batch_kernel(const __grid_constant__ DevScene ds, const GrtRay* __restrict__ rays, uint64_t n, GrtHit* __restrict__ hits) {
    extern __shared__ __align__(16) unsigned char smem[];
    SceneView sv = make_view<STAGED>(ds, smem);
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4* rp = (const float4*)(rays + i);
    float4 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
    RayD r;
    ray_setup<FEAT>(r, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), c.x);
    uint32_t self_id = __float_as_uint(c.y);
    MediumRngCtx mr;
    mr.pixel = (uint32_t)i; mr.sample = (uint32_t)(i >> 32); mr.bounce = 0; mr.k0 = 0; mr.k1 = 0; mr.count = 0;
    HitInfo h;
    GrtHit out;
    if (closest_hit<FEAT | F_DUPIDS | F_TMIN_F64 | (STAGED == 0 ? F_GMEM : 0u), false, false>(sv, ds.root, r, a.w, b.w, self_id, 0xFFFFFFFFu, &mr, h, nullptr)) {
        Surface s;
        finish_hit<FEAT | F_TMIN_F64>(sv, r, h, true, s);
        out.t = h.t; out.id = s.id; out.ref = h.ref; out.front_face = s.front ? 1u : 0u;
        out.p[0] = s.p.x; out.p[1] = s.p.y; out.p[2] = s.p.z; out.u = s.u;
        out.n[0] = s.n.x; out.n[1] = s.n.y; out.n[2] = s.n.z; out.v = s.v;
    } else {
        out.t = __int_as_float(0x7f800000); out.id = GRT_NO_ID; out.ref = GRT_MAKE_REF(GRT_REF_NONE, 0); out.front_face = 0;
        out.p[0] = out.p[1] = out.p[2] = 0; out.u = 0; out.n[0] = out.n[1] = out.n[2] = 0; out.v = 0;
    }
    hits[i] = out;
}

// ===========================================================================
// megakernel
// ===========================================================================
// One warp renders one pixel at a time.  Its 32 lanes pull strata from a
// warp-uniform counter (ballot + popc ranks), so a lane whose path ends takes
// the next sample immediately and the warp stays full until the pixel's
// samples run out.  Lanes that find the current pixel exhausted start on the
// warp's NEXT pixel (two pixels in flight), so there is no drain bubble even
// with few samples per pixel (multi-GPU strata sharding).  Every pixel has one
// writer and every sample is keyed by its global index, so a frame is
// reproducible up to fp32 summation order (which lane sums which strata depends
// on the pixel the warp rendered before, and pixels are claimed from an atomic
// counter): two renders of the headline frame agree to 2e-6 relative.
struct RenderParams {
    DevScene scene;
    DevCamera cam;
    uint32_t k0, k1;
    uint32_t sample_first, sample_stride, n_my;   // strata s = first + k*stride, k < n_my
    int x0, y0, ww, wh;                            // pixel window
    uint32_t n_pixels;                             // ww * wh
    uint32_t claim;                                // pixels claimed per atomicAdd
    int atomic_sum;                                // rgb_sum is shared with other renders / peer GPUs: accumulate with system-scope atomics
    int trav_exit16;                               // a traversal slice ends when fewer than trav_exit16/16 of its lanes have work left
    float* rgb_sum;                                // device, W*H*3, += per pixel
    unsigned int* counter;
    GrtStats* stats;
};

// The resumable, warp-synchronous traversal pays off where traversal lengths within a warp differ by orders of
// magnitude (a large mesh next to empty space: +8 % on the 1M-triangle config); on the sphere/box BVHs of the book
// scenes the plain per-lane loop is faster (profiles/README.md), so only the mesh variant uses it.
__host__ __device__ constexpr bool mega_resumable(uint32_t feat) { return feat == (V_MESH) && (GRT_RESUMABLE_REQ) == 0u; }
__host__ __device__ constexpr int mega_min_blocks(uint32_t feat) { return mega_resumable(feat) ? GRT_MEGA_MIN_BLOCKS_BVH : GRT_MEGA_MIN_BLOCKS; }

template <uint32_t FEAT, int STAGED, bool STATS>
__global__ void __launch_bounds__(GRT_MEGA_THREADS, mega_min_blocks(FEAT)) render_mega_kernel(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    SceneView sv = make_view<STAGED>(P.scene, smem);

    const unsigned FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const DevCamera& cam = P.cam;
    const int max_depth = cam.max_depth;

    // warp-uniform scheduling state
    uint32_t blk_next = 0, blk_end = 0;      // unclaimed remainder of the warp's current pixel block
    uint32_t pix_cur = 0xffffffffu;          // pixel being retired next (ordinal in the window)
    uint32_t pix_nxt = 0xffffffffu;          // the pixel after it (claimed lazily)
    uint32_t k_alloc = 0;                    // next stratum ordinal of the pixel being handed out
    bool alloc_on_nxt = false;               // samples are being handed out from pix_nxt
    bool exhausted = false;                  // no more pixels to claim

    // pixel coordinates packed as (y << 16) | x.  One integer division per claimed BLOCK of pixels (next to the
    // atomic, so it is not speculated into the loop), then incremental.
    uint32_t blk_x = 0, blk_y = 0, claimed_xy = 0;
    auto claim_pixel = [&]() -> uint32_t {
        if (blk_next == blk_end && !exhausted) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(P.counter, P.claim);
            base = __shfl_sync(FULL, base, 0);
            if (base >= P.n_pixels) { exhausted = true; }
            else {
                blk_next = base; blk_end = min(base + P.claim, P.n_pixels);
                blk_y = base / (uint32_t)P.ww; blk_x = base - blk_y * (uint32_t)P.ww;
            }
        }
        if (blk_next < blk_end) {
            claimed_xy = (((uint32_t)P.y0 + blk_y) << 16) | ((uint32_t)P.x0 + blk_x);
            if (++blk_x == (uint32_t)P.ww) { blk_x = 0; blk_y++; }
            return blk_next++;
        }
        return 0xffffffffu;
    };
    pix_cur = claim_pixel();
    if (pix_cur == 0xffffffffu) return;
    uint32_t xy_cur = claimed_xy, xy_nxt = 0;

    // lane state
    bool active = false;
    uint32_t my_parity = 0;                  // 0: my path belongs to pix_cur, 1: to pix_nxt
    f3 acc0 = mk3(0, 0, 0), acc1 = mk3(0, 0, 0);
    RayD ray;
    uint32_t self_id = GRT_NO_ID, self_ref = 0xFFFFFFFFu, my_sample = 0, my_pixel_index = 0;
    int bounce = 0;
    // Recursive firefly clamp (camera.go:327-341) without recursion.  With T = running product of all
    // weights/attenuations, E = terminal radiance and P0 = T (x) E, the radiance returned at clamped
    // vertex j is P_j * S_j with P_j = P0 / T_before_j (componentwise) and
    // S_j = min(S_{j+1}, M / sum(P_j)); hence L0 = P0 * min(1, M / max_j dot(P0, 1/T_before_j)).
    // The stack keeps R_j = 1/T_before_j per clamped vertex; the unwind is one dot product and a max
    // per entry (no dependent divide/compare chain).  An exactly-zero factor component (e.g. gold's
    // blue albedo, main.go:382) cannot be divided out again, so zero factors are kept OUT of T and
    // tracked in zinfo: byte c = number of stack entries whose suffix contains a zero in component c,
    // bit 24+c = a zero was seen at all (then the final component is 0).
    f3 T = mk3(1, 1, 1);
    uint32_t zinfo = 0;
    int sp = 0;
    // the first GRT_RS_SMEM entries live in shared memory (one column per thread, conflict-free); deeper entries —
    // rare — in local memory.  Keeping the hot entries out of local memory matters: 113 664 resident threads x a
    // 1 KB frame does not fit the L2, and the write-backs showed up as 11 GB/s of DRAM writes (profiles/README.md).
    __shared__ float4 rs_smem[GRT_RS_SMEM][GRT_MEGA_THREADS];
    float4 rs_deep[WEIGHT_STACK - GRT_RS_SMEM];
    struct ClampStack {
        float4* sm; float4* deep;
        __device__ __forceinline__ float4 operator[](size_t i) const { return i < GRT_RS_SMEM ? sm[i * GRT_MEGA_THREADS] : deep[i - GRT_RS_SMEM]; }
        __device__ __forceinline__ void put(int i, float4 v) { if (i < GRT_RS_SMEM) sm[i * GRT_MEGA_THREADS] = v; else deep[i - GRT_RS_SMEM] = v; }
    } rstack;
    rstack.sm = &rs_smem[0][threadIdx.x]; rstack.deep = rs_deep;
    // resumable traversal state (BVH scenes only)
    constexpr bool RESUMABLE = mega_resumable(FEAT);
    __shared__ uint32_t trav_smem[RESUMABLE ? GRT_TRAV_SMEM : 1][RESUMABLE ? GRT_MEGA_THREADS : 1];
    TravState<false, RESUMABLE ? GRT_MEGA_THREADS : 1> ts;
    ts.set_ext(&trav_smem[0][RESUMABLE ? threadIdx.x : 0]);
    ts.sp = 0;
    bool tracing = false;
    uint32_t med_count = 0;
    // stats
    uint32_t st_lane_iters = 0, st_paths = 0, st_segments = 0, st_diffuse = 0, st_specular = 0, st_lightpdf = 0, st_nan = 0, st_iters = 0;
    TraceCounters tc;
    tc.box = tc.sphere = tc.quad = tc.tri = tc.medium = 0;

    for (;;) {
        // ---- hand out strata to idle lanes ---------------------------------
        unsigned need = __ballot_sync(FULL, !active);
        while (need) {
            uint32_t pix_alloc = alloc_on_nxt ? pix_nxt : pix_cur;
            uint32_t avail = (pix_alloc == 0xffffffffu) ? 0u : (P.n_my - k_alloc);
            uint32_t rank = __popc(need & lt_mask);
            bool want = ((need >> lane) & 1u) != 0;
            if (want && rank < avail) {
                uint32_t k = k_alloc + rank;
                my_sample = P.sample_first + k * P.sample_stride;
                my_parity = alloc_on_nxt ? 1u : 0u;
                const uint32_t xy = alloc_on_nxt ? xy_nxt : xy_cur;   // warp-uniform, computed once per pixel
                int px = (int)(xy & 0xffffu), py = (int)(xy >> 16);
                my_pixel_index = (uint32_t)(py * cam.width + px);
                f3 o, d; float time;
                camera_ray<FEAT>(cam, px, py, my_sample, my_pixel_index, P.k0, P.k1, o, d, time);
                ray_setup<FEAT>(ray, o, d, time);
                self_id = GRT_NO_ID; self_ref = 0xFFFFFFFFu; bounce = 0; T = mk3(1, 1, 1); zinfo = 0; sp = 0;
                active = true;
                if (STATS) st_paths++;
            }
            uint32_t taken = min((uint32_t)__popc(need), avail);
            k_alloc += taken;
            need = __ballot_sync(FULL, !active);
            if (!need) break;
            // this pixel is used up: move allocation to the next pixel, at most one ahead
            if (alloc_on_nxt) break;
            if (pix_nxt == 0xffffffffu) { pix_nxt = claim_pixel(); xy_nxt = claimed_xy; }
            if (pix_nxt == 0xffffffffu) break;
            alloc_on_nxt = true; k_alloc = 0;
        }
        // ---- retire pix_cur once no lane works on it and its strata are all handed out
        bool cur_done_alloc = alloc_on_nxt || (k_alloc >= P.n_my);
        if (cur_done_alloc && __ballot_sync(FULL, active && my_parity == 0u) == 0u) {
            f3 s = acc0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s.x += __shfl_xor_sync(FULL, s.x, off);
                s.y += __shfl_xor_sync(FULL, s.y, off);
                s.z += __shfl_xor_sync(FULL, s.z, off);
            }
            if (lane == 0) {
                int px = (int)(xy_cur & 0xffffu), py = (int)(xy_cur >> 16);
                float* dst = P.rgb_sum + ((size_t)py * cam.width + px) * 3;
                if (P.atomic_sum) {   // the buffer is shared (another GPU's shard over NVLink peer memory, grt_multi.cu)
                    atomicAdd_system(dst, s.x); atomicAdd_system(dst + 1, s.y); atomicAdd_system(dst + 2, s.z);
                } else {
                    dst[0] += s.x; dst[1] += s.y; dst[2] += s.z;   // this warp owns the pixel for the whole launch
                }
            }
            acc0 = acc1; acc1 = mk3(0, 0, 0);
            if (active) my_parity = 0u;   // survivors were on pix_nxt
            if (alloc_on_nxt) { pix_cur = pix_nxt; xy_cur = xy_nxt; pix_nxt = 0xffffffffu; alloc_on_nxt = false; }
            else { pix_cur = claim_pixel(); xy_cur = claimed_xy; k_alloc = 0; }
            if (pix_cur == 0xffffffffu) break;
            continue;
        }
        if (STATS && lane == 0) st_iters++;
        if constexpr (!RESUMABLE) { if (!active) continue; }
        if (STATS && active) st_lane_iters++;

        // ---- one path segment: trace ------------------------------------------
        MediumRngCtx mr;
        mr.pixel = my_pixel_index; mr.sample = my_sample; mr.bounce = (uint32_t)bounce; mr.k0 = P.k0; mr.k1 = P.k1; mr.count = 0;
        HitInfo h;
        const float INF = __int_as_float(0x7f800000);
        bool hit;
        if constexpr (RESUMABLE) {
            // a BVH scene: all 32 lanes run one warp-synchronous traversal slice (idle lanes carry an empty stack), then
            // the lanes that are done shade and start their next segment while the others resume (dev_trace.cuh)
            if (active && !tracing) { trav_begin(ts, P.scene.root, INF); tracing = true; med_count = 0; if (STATS) st_segments++; }
            mr.count = med_count;
            const bool fin = trav_run<FEAT | (STAGED == 0 ? F_GMEM : 0u), false, STATS, true>(sv, ts, ray, 0.001f, self_id, self_ref, &mr, &tc, FULL, P.trav_exit16);
            med_count = mr.count;
            if (!active || !fin) continue;
            tracing = false;
            hit = trav_end<FEAT>(sv, ts, ray, h);
        } else {
            if (STATS) st_segments++;
            hit = closest_hit<FEAT | (STAGED == 0 ? F_GMEM : 0u), false, STATS>(sv, P.scene.root, ray, 0.001f, INF, self_id, self_ref, &mr, h, &tc);   // camera.go:300
        }

        f3 Lterm = mk3(0, 0, 0);
        bool terminate = false, isnan_path = false;
        if (!hit) { Lterm = cam.background; terminate = true; }   // camera.go:301
        else {
            Surface s;
            finish_hit<FEAT>(sv, ray, h, false, s);
            const GrtMaterial mat = sv.materials()[s.mat];
            if ((FEAT & F_SPHERE) && (FEAT & F_TEXTURE) && GRT_REF_TYPE(h.ref) == GRT_REF_SPHERE && material_needs_uv<FEAT>(sv, mat))
                sphere_uv(sv.spheres()[h.ref & GRT_REF_MASK], s);   // only image textures read a sphere's (u,v)
            uint32_t nlp = 0;
            ShadeResult R = shade_vertex<FEAT>(sv, ray, s, mat, my_pixel_index, my_sample, (uint32_t)bounce, P.k0, P.k1, STATS ? &nlp : nullptr);
            if (STATS) st_lightpdf += nlp;
            if (R.kind == SHADE_TERMINATE) { Lterm = R.value; terminate = true; }
            else if (R.kind == SHADE_NAN) { isnan_path = true; terminate = true; }
            else {
                if (R.kind == SHADE_SPECULAR) {   // camera.go:315-317: unclamped, folds into the product
                    if (STATS) st_specular++;
                    apply_factor(T, zinfo, R.value, sp);
                } else {                          // camera.go:319-330: a clamped vertex
                    if (STATS) st_diffuse++;
                    if (R.value.x == 0.0f && R.value.y == 0.0f && R.value.z == 0.0f) { terminate = true; }  // weight 0: the sample is 0 whatever follows
                    else {
                        // 1/T_before_j; a zero component stays zero in P0 as well, so its reciprocal is irrelevant
                        rstack.put(sp++, recip_factor(T));
                        apply_factor(T, zinfo, R.value, sp);
                    }
                }
                if (!terminate) {
                    bounce++;
                    if (bounce > max_depth) { terminate = true; }   // depth < 0 -> (0,0,0), camera.go:294-296
                    else {
                        ray_setup<FEAT>(ray, s.p, R.dir, ray.time);
                        self_id = s.is_surface ? s.id : GRT_NO_ID;
                        self_ref = s.is_surface ? h.ref : 0xFFFFFFFFu;
                    }
                }
            }
        }
        if (terminate) {
            f3 L = Lterm;
            if (isnan_path) { L = mk3(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000), __int_as_float(0x7fc00000)); if (STATS) st_nan++; }
            else if (L.x != 0.0f || L.y != 0.0f || L.z != 0.0f) {
                // unwind the recursion of camera.go:327-330 from the terminal radiance
                // (paths through media collect many clamped vertices: there the single loop measured 4-9 % faster than the split one,
                // on the surface-only Cornell box the split one 3 % faster)
                if constexpr ((FEAT & F_MEDIUM) != 0) L = unwind_clamp(T, zinfo, L, rstack, sp, cam.max_contribution);
                else L = unwind_clamp_split<GRT_RS_SMEM>(T, zinfo, L, [&](int i) { return rstack.sm[i * GRT_MEGA_THREADS]; },
                                                         [&](int i) { return rstack.deep[i - GRT_RS_SMEM]; }, sp, cam.max_contribution);
            }
            { const f3 Z = mk3(0, 0, 0); acc0 = acc0 + (my_parity == 0u ? L : Z); acc1 = acc1 + (my_parity == 0u ? Z : L); }
            active = false;
        }
    }

    if (STATS) {
        unsigned long long v[13] = {st_paths, st_segments, tc.box, tc.sphere, tc.quad, tc.tri, tc.medium, st_diffuse, st_specular, st_lightpdf, st_nan, st_iters, st_lane_iters};
        unsigned long long* dst = (unsigned long long*)P.stats;
#pragma unroll
        for (int i = 0; i < 13; i++) {
            unsigned long long x = v[i];
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
            if (lane == 0 && x) atomicAdd(dst + i, x);
        }
    }
}

// ===========================================================================
// tonemap, color.go:14-46
// ===========================================================================
__global__ void tonemap_kernel(const float* __restrict__ sum, uint8_t* __restrict__ out, uint64_t n, float scale) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c = sum[i] * scale;          // pixelColor.Scale(pixelSamplesScale), camera.go:103
    if (isnan(c)) c = 0.0f;            // color.go:28-36
    c = c <= 0.0f ? 0.0f : sqrtf(c);   // linearToGamma :14-19
    c = c < 0.0f ? 0.0f : (c > 0.99999f ? 0.99999f : c);   // intensity.Clamp :11,40-42
    out[i] = (uint8_t)(int)(c * 256.0f);
}

// ===========================================================================
// host side: upload, dispatch
// ===========================================================================
struct GrtSceneDev {
    int device = 0;
    DevScene ds;
    void* d_blob = nullptr;
    void* d_tris = nullptr;
    void* d_tri_shade = nullptr;
    void* d_tri_v64 = nullptr;
    void* d_texels = nullptr;
    void* d_perlins = nullptr;
    unsigned int* d_counter = nullptr;
    void* wf_pool = nullptr;         // wavefront path pool, kept between renders (allocating 2 GB per call stalls for up to 0.4 s)
    size_t wf_pool_bytes = 0;
    void* wf_pinned = nullptr;
    int staged = 0;   // 0 none, 1 hot arrays, 2 whole blob
    int sm_count = 0;
};

static uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

static uint32_t scan_features(const GrtScene* s) {
    uint32_t f = 0;
    if (s->n_nodes) f |= F_NODE;
    if (s->n_spheres) f |= F_SPHERE;
    if (s->n_quads) f |= F_QUAD;
    if (s->n_boxes) f |= F_BOX;
    if (s->n_tris) f |= F_TRI;
    if (s->n_items) f |= F_LIST;
    if (s->n_media) f |= F_MEDIUM;
    for (uint32_t i = 0; i < s->n_materials; i++) {
        uint32_t t = s->materials[i].type;
        if (t == GRT_MAT_METAL || t == GRT_MAT_DIELECTRIC) f |= F_SPECULAR;
        if (t == GRT_MAT_ISOTROPIC) f |= F_ISOTROPIC;
    }
    for (uint32_t i = 0; i < s->n_textures; i++) if (s->textures[i].type != GRT_TEX_SOLID) f |= F_TEXTURE;
    for (uint32_t i = 0; i < s->n_quads; i++) if (!(s->quads[i].flags & GRT_QUAD_AXIS_ALIGNED)) f |= F_ROTQUAD;
    for (uint32_t i = 0; i < s->n_lights; i++) {
        if (s->lights[i].type == GRT_LIGHT_SPHERE) f |= F_SPHERE_LIGHT;
        if (s->lights[i].type == GRT_LIGHT_QUAD) f |= F_QUAD_LIGHT;
        if (s->lights[i].type == GRT_LIGHT_TRI) f |= F_TRI_LIGHT;
    }
    if (s->tri_shade) f |= F_TRISHADE;
    {   // does any QUAD object id name more than one flat quad (an object listed or instanced twice)?  Spheres and
        // triangles carry their id in the hot record and are always excluded by id; quads are excluded by flat ref
        // unless this flag says a ref no longer identifies the object.
        std::vector<uint32_t> ids;
        ids.reserve(s->n_quads);
        for (uint32_t i = 0; i < s->n_quads; i++) ids.push_back(s->quads[i].id);
        std::sort(ids.begin(), ids.end());
        for (size_t i = 1; i < ids.size(); i++) if (ids[i] == ids[i - 1]) { f |= F_DUPIDS; break; }
    }
    return f;
}

static int validate_scene(const GrtScene* s) {
    if (!s) { grt_set_error("scene is NULL"); return GRT_E_INVALID; }
    if (s->abi_version != GRT_ABI_VERSION) { grt_set_error("GrtScene.abi_version mismatch"); return GRT_E_INVALID; }
    auto check_ref = [&](uint32_t ref) -> bool {
        uint32_t t = GRT_REF_TYPE(ref), i = ref & GRT_REF_MASK;
        switch (t) {
            case GRT_REF_NODE: return i < s->n_nodes;
            case GRT_REF_SPHERE: return i < s->n_spheres;
            case GRT_REF_QUAD: return i < s->n_quads;
            case GRT_REF_BOX: return i < s->n_boxes;
            case GRT_REF_TRI: return i < s->n_tris;
            case GRT_REF_LIST: return i < s->n_items;
            case GRT_REF_MEDIUM: return i < s->n_media;
            case GRT_REF_NONE: return true;
        }
        return false;
    };
    if (!check_ref(s->root)) { grt_set_error("scene root ref out of range"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_nodes; i++)
        if (!check_ref(s->nodes[i].left & ~GRT_NODE_HINT_BIT) || !check_ref(s->nodes[i].right & ~GRT_NODE_HINT_BIT)) { grt_set_error("BVH node child ref out of range"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_items; i++)
        if (!check_ref(s->items[i] & ~GRT_LIST_LAST)) { grt_set_error("list item ref out of range"); return GRT_E_INVALID; }
    if (s->n_items && !(s->items[s->n_items - 1] & GRT_LIST_LAST)) { grt_set_error("items[] does not end with GRT_LIST_LAST"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_media; i++) {
        if (!check_ref(s->media[i].boundary) || s->media[i].mat >= s->n_materials) { grt_set_error("medium ref out of range"); return GRT_E_INVALID; }
    }
    for (uint32_t i = 0; i < s->n_spheres; i++) if (s->spheres[i].mat >= s->n_materials) { grt_set_error("sphere material out of range"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_boxes; i++) if ((uint64_t)s->boxes[i].first_quad + 6 > s->n_quads) { grt_set_error("box quad range out of range"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_quads; i++) if (s->quads[i].mat >= s->n_materials) { grt_set_error("quad material out of range"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_tris; i++) if (s->tris[i].mat >= s->n_materials) { grt_set_error("triangle material out of range"); return GRT_E_INVALID; }
    for (uint32_t i = 0; i < s->n_materials; i++) {
        const GrtMaterial& m = s->materials[i];
        if (m.type > GRT_MAT_ISOTROPIC) { grt_set_error("unknown material type"); return GRT_E_INVALID; }
        if ((m.type == GRT_MAT_LAMBERTIAN || m.type == GRT_MAT_DIFFUSE_LIGHT || m.type == GRT_MAT_ISOTROPIC) && m.tex >= s->n_textures) { grt_set_error("material texture out of range"); return GRT_E_INVALID; }
    }
    for (uint32_t i = 0; i < s->n_textures; i++) {
        const GrtTexture& t = s->textures[i];
        if (t.type == GRT_TEX_CHECKER && (t.even >= s->n_textures || t.odd >= s->n_textures)) { grt_set_error("checker texture ids out of range"); return GRT_E_INVALID; }
        if (t.type == GRT_TEX_IMAGE && t.aux >= s->n_images) { grt_set_error("image index out of range"); return GRT_E_INVALID; }
        if (t.type == GRT_TEX_NOISE && (t.aux & 0xFFFFu) >= s->n_perlins) { grt_set_error("perlin index out of range"); return GRT_E_INVALID; }
    }
    if (s->max_depth_hint >= GRT_STACK_BOUNDARY) {
        grt_set_error("scene needs a deeper traversal stack than the kernels provide");
        return GRT_E_UNSUPPORTED;
    }
    return GRT_OK;
}

extern "C" int grt_abi_version(void) { return GRT_ABI_VERSION; }
extern "C" const char* grt_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t grt_launch_count(void) { return g_launches.load(); }
extern "C" int grt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int need_device(int device) {
    int n = grt_device_count();
    if (n <= 0) { grt_set_error("no CUDA device available (libgrt_cuda has no CPU fallback)"); return GRT_E_NO_DEVICE; }
    if (device < 0 || device >= n) { grt_set_error("CUDA device ordinal out of range"); return GRT_E_NO_DEVICE; }
    return GRT_OK;
}

// ---- device-internal repacking (host side of the upload; also reachable without a device, for the CPU tests) ------
struct Repacked {
    std::vector<uint32_t> entries;     // run-length list entries, pairs {first ref | LAST, count}
    std::vector<GrtMedium> media;      // boundary refs remapped
    grt::wide::Result wide;            // the 4-wide BVH
    uint32_t root = 0, n_entries = 0;
};
static int repack_scene(const GrtScene* s, Repacked& R) {
    // run-length list entries: consecutive items of one primitive type with
    // consecutive indices collapse into {first ref, count}; LIST refs are remapped to entry indices
    std::vector<uint32_t> item2entry(s->n_items + 1, 0);
    std::vector<uint32_t>& entries = R.entries;
    for (uint32_t i = 0; i < s->n_items;) {
        uint32_t ref = s->items[i] & ~GRT_LIST_LAST;
        uint32_t t = GRT_REF_TYPE(ref);
        bool prim = t == GRT_REF_SPHERE || t == GRT_REF_QUAD || t == GRT_REF_TRI || t == GRT_REF_BOX;
        uint32_t n = 1;
        item2entry[i] = (uint32_t)(entries.size() / 2);
        bool last = (s->items[i] & GRT_LIST_LAST) != 0;
        while (prim && !last && i + n < s->n_items) {
            uint32_t nx = s->items[i + n] & ~GRT_LIST_LAST;
            if (nx != ref + n) break;
            item2entry[i + n] = (uint32_t)(entries.size() / 2);   // only list STARTS are ever referenced
            last = (s->items[i + n] & GRT_LIST_LAST) != 0;
            n++;
        }
        entries.push_back(ref | (last ? GRT_LIST_LAST : 0u));
        entries.push_back(n);
        i += n;
    }
    auto remap = [&](uint32_t ref) -> uint32_t {
        uint32_t flag = ref & GRT_LIST_LAST, r = ref & ~GRT_LIST_LAST;
        if (GRT_REF_TYPE(r) == GRT_REF_LIST) r = GRT_MAKE_REF(GRT_REF_LIST, item2entry[r & GRT_REF_MASK]);
        return r | flag;
    };
    for (size_t k = 0; k < entries.size(); k += 2) entries[k] = remap(entries[k]);
    std::vector<GrtNode> nodes(s->nodes, s->nodes + s->n_nodes);
    for (auto& n : nodes) { n.left = remap(n.left); n.right = remap(n.right); }
    R.media.assign(s->media, s->media + s->n_media);
    for (auto& m : R.media) m.boundary = remap(m.boundary);
    R.n_entries = (uint32_t)(entries.size() / 2);
    // the 4-wide BVH (wide_bvh.hpp): NODE refs held by lists, media and the root now index the wide array
    grt::wide::Builder wb(s, nodes, entries, R.media);
    if (!wb.run(remap(s->root), R.wide)) { grt_set_error("wide BVH build: " + R.wide.error); return GRT_E_UNSUPPORTED; }
    if (wb.sah && (R.wide.need_main > GRT_STACK_MAIN || R.wide.need_boundary > GRT_STACK_BOUNDARY)) {
        // a regrouped tree may be arbitrarily unbalanced; BuildBVH's median splits are not
        grt::wide::Builder plain(s, nodes, entries, R.media);
        plain.sah = 0;
        R.wide = grt::wide::Result();
        if (!plain.run(remap(s->root), R.wide)) { grt_set_error("wide BVH build: " + R.wide.error); return GRT_E_UNSUPPORTED; }
    }
    if (R.wide.need_main > GRT_STACK_MAIN || R.wide.need_boundary > GRT_STACK_BOUNDARY) {
        grt_set_error("scene needs a deeper traversal stack than the kernels provide");
        return GRT_E_UNSUPPORTED;
    }
    // NODE refs held by list entries (list-only scenes have none; in BVH scenes the lists themselves became wide nodes
    // and the entries are no longer referenced), by media and by the root now name wide nodes
    auto to_wide = [&](uint32_t ref) -> uint32_t {
        uint32_t flag = ref & GRT_LIST_LAST, r = ref & ~GRT_LIST_LAST;
        if (GRT_REF_TYPE(r) == GRT_REF_NODE) r = GRT_MAKE_REF(GRT_REF_NODE, R.wide.node_map[r & GRT_REF_MASK]);
        return r | flag;
    };
    for (size_t k = 0; k < entries.size(); k += 2) entries[k] = to_wide(entries[k]);
    for (size_t i = 0; i < R.media.size(); i++) R.media[i].boundary = R.wide.media_boundary[i];
    R.root = R.wide.root;
    return GRT_OK;
}

// Test hook (no device needed): the wide BVH and run-length entries the upload would build for this scene.
// wnodes: capacity cap_nodes x 32 floats; entries: capacity cap_entries x 2 words.  Returns GRT_E_INVALID when too small.
extern "C" int grt_debug_repack(const GrtScene* s, float* wnodes, uint32_t cap_nodes, uint32_t* n_wide, uint32_t* entries, uint32_t cap_entries,
                                uint32_t* n_entries, uint32_t* root, uint32_t* media_boundaries, int* need_main, int* need_boundary) {
    int rc = validate_scene(s);
    if (rc) return rc;
    Repacked R;
    if ((rc = repack_scene(s, R))) return rc;
    if (n_wide) *n_wide = R.wide.n_wide;
    if (n_entries) *n_entries = R.n_entries;
    if (root) *root = R.root;
    if (need_main) *need_main = R.wide.need_main;
    if (need_boundary) *need_boundary = R.wide.need_boundary;
    if (R.wide.n_wide > cap_nodes || R.n_entries > cap_entries) { grt_set_error("grt_debug_repack: output buffers too small"); return GRT_E_INVALID; }
    if (wnodes) memcpy(wnodes, R.wide.wnodes.data(), (size_t)R.wide.n_wide * GRT_WNODE_FLOATS * sizeof(float));
    if (entries) memcpy(entries, R.entries.data(), (size_t)R.n_entries * 8);
    if (media_boundaries) for (size_t i = 0; i < R.media.size(); i++) media_boundaries[i] = R.media[i].boundary;
    return GRT_OK;
}

extern "C" int grt_scene_upload(const GrtScene* s, int device, GrtSceneHandle* out) {
    if (!out) { grt_set_error("out handle is NULL"); return GRT_E_INVALID; }
    *out = nullptr;
    int rc = validate_scene(s);
    if (rc) return rc;
    rc = need_device(device);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    GrtSceneDev* h = new GrtSceneDev();
    struct Guard { GrtSceneDev* h; ~Guard() { if (h) grt_scene_free(h); } } guard{h};   // every early return frees the handle
    h->device = device;
    DevScene& ds = h->ds;
    memset(&ds, 0, sizeof(ds));
    // ---- device-internal repacking (list entries, 4-wide BVH) --------------------
    Repacked RP;
    if ((rc = repack_scene(s, RP))) return rc;
    std::vector<uint32_t>& entries = RP.entries;
    std::vector<GrtMedium>& media = RP.media;
    grt::wide::Result& wide = RP.wide;
    const uint32_t n_entries = RP.n_entries;
    ds.root = RP.root;
    // ---- pack the blob ------------------------------------------------------
    uint32_t off = 0;
    auto place = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    ds.off_nodes = place(wide.n_wide * (uint32_t)(GRT_WNODE_F4 * 16));   // offset 0: 128-byte aligned like the allocation
    ds.off_spheres = place(s->n_spheres * (uint32_t)sizeof(GrtSphere));
    ds.off_quads = place(s->n_quads * (uint32_t)sizeof(DQuadHot));
    ds.off_boxes = place(s->n_boxes * (uint32_t)sizeof(GrtBox));
    ds.off_items = place(n_entries * 8u);
    ds.off_media = place(s->n_media * (uint32_t)sizeof(GrtMedium));
    const uint32_t hot_end = off;   // everything above is read inside the traversal loop
    ds.off_quads_cold = place(s->n_quads * (uint32_t)sizeof(DQuadCold));
    ds.off_materials = place(s->n_materials * (uint32_t)sizeof(GrtMaterial));
    ds.off_textures = place(s->n_textures * (uint32_t)sizeof(GrtTexture));
    ds.off_lights = place(s->n_lights * (uint32_t)sizeof(GrtLight));
    ds.off_dlights = place(s->n_lights * (uint32_t)sizeof(DLight));
    ds.off_images = place(s->n_images * (uint32_t)sizeof(GrtImage));
    if (off == 0) off = 16;
    std::vector<unsigned char> blob(off, 0);
    auto put = [&](uint32_t o, const void* p, size_t bytes) { if (bytes) memcpy(blob.data() + o, p, bytes); };
    put(ds.off_nodes, wide.wnodes.data(), (size_t)wide.n_wide * GRT_WNODE_F4 * 16);
    put(ds.off_spheres, s->spheres, s->n_spheres * sizeof(GrtSphere));
    {
        std::vector<DQuadHot> hot(s->n_quads);
        std::vector<DQuadCold> cold(s->n_quads);
        for (uint32_t i = 0; i < s->n_quads; i++) {
            const GrtQuad& q = s->quads[i];
            hot[i].plane = make_float4(q.n[0], q.n[1], q.n[2], q.D);
            double aq = (double)q.A[0] * q.Q[0] + (double)q.A[1] * q.Q[1] + (double)q.A[2] * q.Q[2];
            double bq = (double)q.B[0] * q.Q[0] + (double)q.B[1] * q.Q[1] + (double)q.B[2] * q.Q[2];
            hot[i].A = make_float4(q.A[0], q.A[1], q.A[2], (float)-aq);
            hot[i].B = make_float4(q.B[0], q.B[1], q.B[2], (float)-bq);
            DQuadCold& c = cold[i];
            memset(&c, 0, sizeof(c));
            // NewONB(normal), onb.go:13-25, in fp64
            double n[3] = {q.n64[0], q.n64[1], q.n64[2]};
            double a[3] = {fabs(n[0]) > 0.9 ? 0.0 : 1.0, fabs(n[0]) > 0.9 ? 1.0 : 0.0, 0.0};
            double v[3] = {n[1] * a[2] - n[2] * a[1], n[2] * a[0] - n[0] * a[2], n[0] * a[1] - n[1] * a[0]};
            double vl = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            for (int k = 0; k < 3; k++) v[k] /= vl;
            double u[3] = {n[1] * v[2] - n[2] * v[1], n[2] * v[0] - n[0] * v[2], n[0] * v[1] - n[1] * v[0]};
            double ul = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
            for (int k = 0; k < 3; k++) { u[k] /= ul; c.n[k] = q.n[k]; c.ou[k] = (float)u[k]; c.ov[k] = (float)v[k]; c.n64[k] = q.n64[k]; }
            c.flags = q.flags; c.mat = q.mat; c.id = q.id; c.D64 = q.D64;
        }
        put(ds.off_quads, hot.data(), hot.size() * sizeof(DQuadHot));
        put(ds.off_quads_cold, cold.data(), cold.size() * sizeof(DQuadCold));
    }
    put(ds.off_boxes, s->boxes, s->n_boxes * sizeof(GrtBox));
    put(ds.off_items, entries.data(), n_entries * 8u);
    put(ds.off_media, media.data(), s->n_media * sizeof(GrtMedium));
    put(ds.off_materials, s->materials, s->n_materials * sizeof(GrtMaterial));
    put(ds.off_textures, s->textures, s->n_textures * sizeof(GrtTexture));
    put(ds.off_lights, s->lights, s->n_lights * sizeof(GrtLight));
    {
        std::vector<DLight> dl(s->n_lights);
        for (uint32_t i = 0; i < s->n_lights; i++) {
            memset(&dl[i], 0, sizeof(DLight));
            if (s->lights[i].type != GRT_LIGHT_QUAD) continue;
            const double* p = s->lights[i].p;   // Q[3], u[3], v[3], n[3], w[3], D, area
            const double *Q = p, *u = p + 3, *v = p + 6, *n = p + 9, *w = p + 12;
            double A[3] = {v[1] * w[2] - v[2] * w[1], v[2] * w[0] - v[0] * w[2], v[0] * w[1] - v[1] * w[0]};   // v x w: alpha = A.(p - Q)
            double B[3] = {w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0]};   // w x u: beta  = B.(p - Q)
            dl[i].plane = make_float4((float)n[0], (float)n[1], (float)n[2], (float)p[15]);
            dl[i].A = make_float4((float)A[0], (float)A[1], (float)A[2], (float)-(A[0] * Q[0] + A[1] * Q[1] + A[2] * Q[2]));
            dl[i].B = make_float4((float)B[0], (float)B[1], (float)B[2], (float)-(B[0] * Q[0] + B[1] * Q[1] + B[2] * Q[2]));
            dl[i].Qa = make_float4((float)Q[0], (float)Q[1], (float)Q[2], (float)p[16]);
            dl[i].U = make_float4((float)u[0], (float)u[1], (float)u[2], 0.0f);
            dl[i].V = make_float4((float)v[0], (float)v[1], (float)v[2], 0.0f);
        }
        put(ds.off_dlights, dl.data(), dl.size() * sizeof(DLight));
    }
    put(ds.off_images, s->images, s->n_images * sizeof(GrtImage));
    ds.blob_bytes = off;
    ds.stage_bytes = off <= GRT_STAGE_MAX_BYTES ? off : (hot_end <= GRT_STAGE_MAX_BYTES ? hot_end : 0u);
    ds.n_nodes = wide.n_wide; ds.n_spheres = s->n_spheres; ds.n_quads = s->n_quads; ds.n_boxes = s->n_boxes; ds.n_items = n_entries; ds.n_media = s->n_media;
    ds.n_materials = s->n_materials; ds.n_textures = s->n_textures; ds.n_lights = s->n_lights; ds.n_images = s->n_images;
    ds.n_tris = s->n_tris; ds.n_perlins = s->n_perlins;
    ds.lights_mode = s->lights_mode; ds.stack_need = (uint32_t)wide.need_main;
    ds.features = scan_features(s);
    auto upload = [&](void** dptr, const void* src, size_t bytes) -> int {
        *dptr = nullptr;
        if (!bytes) return GRT_OK;
        CUDA_TRY(grt_dev_alloc(dptr, bytes));
        CUDA_TRY(cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
        return GRT_OK;
    };
    if ((rc = upload(&h->d_blob, blob.data(), blob.size()))) return rc;
    if ((rc = upload(&h->d_tris, s->tris, (size_t)s->n_tris * sizeof(GrtTri)))) return rc;
    if (s->tri_shade && (rc = upload(&h->d_tri_shade, s->tri_shade, (size_t)s->n_tris * sizeof(GrtTriShade)))) return rc;
    if (s->tri_v64 && (rc = upload(&h->d_tri_v64, s->tri_v64, (size_t)s->n_tris * 9 * sizeof(double)))) return rc;
    if ((rc = upload(&h->d_texels, s->texels, (size_t)s->n_texel_bytes))) return rc;
    if ((rc = upload(&h->d_perlins, s->perlins, (size_t)s->n_perlins * sizeof(GrtPerlin)))) return rc;
    ds.blob = (const unsigned char*)h->d_blob;
    ds.tris = (const GrtTri*)h->d_tris;
    ds.tri_shade = (const GrtTriShade*)h->d_tri_shade;
    ds.tri_v64 = (const double*)h->d_tri_v64;
    ds.texels = (const uint8_t*)h->d_texels;
    ds.perlins = (const GrtPerlin*)h->d_perlins;
    CUDA_TRY(grt_dev_alloc((void**)&h->d_counter, 256));
    h->staged = ds.stage_bytes == 0 ? 0 : (ds.stage_bytes == ds.blob_bytes ? 2 : 1);
    CUDA_TRY(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));   // (cudaGetDeviceProperties takes milliseconds)
    guard.h = nullptr;
    *out = h;
    return GRT_OK;
}

extern "C" int grt_scene_free(GrtSceneHandle h) {
    if (!h) return GRT_OK;
    cudaSetDevice(h->device);
    grt_dev_free(h->d_blob); grt_dev_free(h->d_tris); grt_dev_free(h->d_tri_shade); grt_dev_free(h->d_tri_v64); grt_dev_free(h->d_texels); grt_dev_free(h->d_perlins); grt_dev_free(h->d_counter);
    grt_dev_free(h->wf_pool);
    if (h->wf_pinned) cudaFreeHost(h->wf_pinned);
    delete h;
    return GRT_OK;
}

const DevScene* grt_internal_dev_scene(GrtSceneHandle h) { return &h->ds; }
int grt_internal_sm_count(GrtSceneHandle h) { return h->sm_count; }
int grt_internal_staged(GrtSceneHandle h) { return h->staged; }
unsigned int* grt_internal_counter(GrtSceneHandle h) { return h->d_counter; }
void* grt_internal_wf_pool(GrtSceneHandle h, size_t bytes, void** pinned64) {
    if (bytes > h->wf_pool_bytes) {
        if (h->wf_pool) { grt_dev_free(h->wf_pool); h->wf_pool = nullptr; h->wf_pool_bytes = 0; }
        cudaError_t e = grt_dev_alloc(&h->wf_pool, bytes);
        if (e != cudaSuccess) { grt_set_error(std::string("wavefront: cannot allocate the path pool (") + std::to_string(bytes >> 20) + " MiB): " + cudaGetErrorString(e)); h->wf_pool = nullptr; return nullptr; }
        h->wf_pool_bytes = bytes;
    }
    if (!h->wf_pinned) {
        cudaError_t e = cudaMallocHost(&h->wf_pinned, 64);
        if (e != cudaSuccess) { grt_set_error(std::string("wavefront: cannot allocate the pinned counter block: ") + cudaGetErrorString(e)); h->wf_pinned = nullptr; return nullptr; }
    }
    *pinned64 = h->wf_pinned;
    return h->wf_pool;
}

// ---- feature-variant dispatch ------------------------------------------------
// Each variant is a feature SUPERSET compiled as its own kernel.
#define V_NODUP(v) ((v) & ~F_DUPIDS)

template <uint32_t FEAT>
static int launch_trace(GrtSceneDev* h, const GrtRay* d_rays, uint64_t n, GrtHit* d_hits, cudaStream_t st) {
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (h->staged == 2) {
        CUDA_TRY(cudaFuncSetAttribute(trace_batch_kernel<FEAT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->ds.stage_bytes));
        trace_batch_kernel<FEAT, 2><<<blocks, 128, h->ds.stage_bytes, st>>>(h->ds, d_rays, n, d_hits);
    } else {   // (hot-prefix staging is not worth a third kernel for the one-shot ray query)
        trace_batch_kernel<FEAT, 0><<<blocks, 128, 0, st>>>(h->ds, d_rays, n, d_hits);
    }
    grt_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GRT_OK;
}

extern "C" int grt_trace_batch_device(GrtSceneHandle h, const GrtRay* d_rays, uint64_t n, GrtHit* d_hits, void* stream) {
    if (!h || (!d_rays && n) || (!d_hits && n)) { grt_set_error("grt_trace_batch_device: NULL argument"); return GRT_E_INVALID; }
    if (n == 0) return GRT_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t f = h->ds.features & ~F_DUPIDS;   // the trace kernels always exclude by object id
    if ((f & ~V_CORNELL) == 0) return launch_trace<V_CORNELL>(h, d_rays, n, d_hits, st);
    if ((f & ~V_SMOKE) == 0) return launch_trace<V_SMOKE>(h, d_rays, n, d_hits, st);
    if ((f & ~V_SPHERES) == 0) return launch_trace<V_SPHERES>(h, d_rays, n, d_hits, st);
    if ((f & ~V_MESH) == 0) return launch_trace<V_MESH>(h, d_rays, n, d_hits, st);
    return launch_trace<V_FULL>(h, d_rays, n, d_hits, st);
}

extern "C" int grt_trace_batch(GrtSceneHandle h, const GrtRay* rays, uint64_t n, GrtHit* hits) {
    if (!h || (!rays && n) || (!hits && n)) { grt_set_error("grt_trace_batch: NULL argument"); return GRT_E_INVALID; }
    if (n == 0) return GRT_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    GrtRay* d_rays = nullptr; GrtHit* d_hits = nullptr;
    CUDA_TRY(grt_dev_alloc((void**)&d_rays, n * sizeof(GrtRay)));
    cudaError_t e = grt_dev_alloc((void**)&d_hits, n * sizeof(GrtHit));
    if (e != cudaSuccess) { grt_dev_free(d_rays); grt_set_error(cudaGetErrorString(e)); return GRT_E_CUDA; }
    int rc = GRT_OK;
    do {
        if ((e = cudaMemcpy(d_rays, rays, n * sizeof(GrtRay), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        rc = grt_trace_batch_device(h, d_rays, n, d_hits, nullptr);
        if (rc) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        if ((e = cudaMemcpy(hits, d_hits, n * sizeof(GrtHit), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    grt_dev_free(d_rays); grt_dev_free(d_hits);
    if (e != cudaSuccess) { grt_set_error(cudaGetErrorString(e)); return GRT_E_CUDA; }
    return rc;
}

int grt_make_dev_camera(const GrtCamera* c, DevCamera* out) {
    if (!c) { grt_set_error("camera is NULL"); return GRT_E_INVALID; }
    if (c->width <= 0 || c->height <= 0 || c->spp_sqrt <= 0 || c->max_depth < 0) { grt_set_error("camera: width/height/spp_sqrt must be positive"); return GRT_E_INVALID; }
    if (c->width > 65535 || c->height > 65535) { grt_set_error("camera: image larger than 65535 pixels on a side"); return GRT_E_UNSUPPORTED; }
    if (c->max_depth + 1 > WEIGHT_STACK) { grt_set_error("camera: MaxDepth exceeds the kernel's weight stack (63)"); return GRT_E_UNSUPPORTED; }
    if ((uint64_t)c->spp_sqrt * (uint64_t)c->spp_sqrt > 0xffffffffull) { grt_set_error("camera: more than 2^32 strata per pixel"); return GRT_E_UNSUPPORTED; }
    DevCamera d;
    auto f = [](const double* v) { return f3{(float)v[0], (float)v[1], (float)v[2]}; };
    d.center = f(c->center);
    double rel[3] = {c->pixel00[0] - c->center[0], c->pixel00[1] - c->center[1], c->pixel00[2] - c->center[2]};
    d.p00_rel = f(rel);
    d.du = f(c->delta_u); d.dv = f(c->delta_v); d.defu = f(c->defocus_u); d.defv = f(c->defocus_v);
    d.background = f(c->background);
    d.max_contribution = (float)c->max_contribution;
    d.recip_spp_sqrt = (float)(1.0 / (double)c->spp_sqrt);
    d.defocus_angle = (float)c->defocus_angle;
    d.width = c->width; d.height = c->height; d.spp_sqrt = c->spp_sqrt; d.max_depth = c->max_depth;
    *out = d;
    return GRT_OK;
}

template <uint32_t FEAT>
static int launch_mega(GrtSceneDev* h, RenderParams& P, bool stats, cudaStream_t st) {
    int blocks = h->sm_count * mega_min_blocks(FEAT);
    size_t smem = h->staged ? h->ds.stage_bytes : 0;
#define GRT_LAUNCH(STAGED_, STATS_)                                                                                           \
    do {                                                                                                                      \
        if (smem) CUDA_TRY(cudaFuncSetAttribute(render_mega_kernel<FEAT, STAGED_, STATS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        render_mega_kernel<FEAT, STAGED_, STATS_><<<blocks, GRT_MEGA_THREADS, smem, st>>>(P);                                 \
    } while (0)
    // quad-only variants are for tiny scenes: whole-blob staging or none (keeps the number of kernels down)
    constexpr bool tiny = (FEAT & (F_NODE | F_SPHERE | F_TRI)) == 0;
    int mode = h->staged;
    if (tiny && mode == 1) { mode = 0; smem = 0; }
    if (mega_resumable(FEAT) && smem > GRT_STAGE_MAX_BYTES_BVH) { mode = 0; smem = 0; }   // the traversal stack needs the shared memory
    if (mode == 2) { if (stats) GRT_LAUNCH(2, true); else GRT_LAUNCH(2, false); }
    else if (mode == 1) { if constexpr (!tiny) { if (stats) GRT_LAUNCH(1, true); else GRT_LAUNCH(1, false); } }
    else { if (stats) GRT_LAUNCH(0, true); else GRT_LAUNCH(0, false); }
#undef GRT_LAUNCH
    grt_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GRT_OK;
}

int grt_render_wavefront(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt, float* d_rgb_sum, cudaStream_t st, GrtStats* d_stats);

int grt_internal_resolve_variant(GrtSceneHandle h, const GrtOptions* opt, bool has_stats_buffer) {
    if (opt->variant != GRT_VARIANT_AUTO) return opt->variant;
    return ((h->ds.features & F_NODE) && !((opt->flags & GRT_OPT_STATS) && has_stats_buffer)) ? GRT_VARIANT_WAVEFRONT : GRT_VARIANT_MEGAKERNEL;
}

// dst[i] += src[i], src in ANOTHER device's memory (NVLink peer loads, 128-bit, coalesced): devices[0] pulls the private
// sums of a peer's wavefront shard once that shard is complete (grt_render_multi).  No atomics: nobody else touches dst
// at that point.  (The first version pushed from the peer with one system-scope atomic per value: 25 M remote atomics
// for a 3840x2160 frame took longer than the shard's render at 16 spp, 593 vs 951 Mpaths/s for ncclReduce.)
__global__ void peer_pull_kernel(float* __restrict__ dst, const float* __restrict__ src, uint64_t n) {
    const uint64_t n4 = n / 4, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = ((const float4*)src)[i];
        float4 d = ((float4*)dst)[i];
        d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w;
        ((float4*)dst)[i] = d;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[n4 * 4 + threadIdx.x] += src[n4 * 4 + threadIdx.x];
}
int grt_internal_peer_pull(float* d_dst, const float* d_src, uint64_t n, cudaStream_t st) {
    if (!n) return GRT_OK;
    uint64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    peer_pull_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_dst, d_src, n);
    grt_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GRT_OK;
}

extern "C" int grt_render_device(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt, float* d_rgb_sum, void* stream, GrtStats* d_stats) {
    if (!h || !cam || !opt || !d_rgb_sum) { grt_set_error("grt_render_device: NULL argument"); return GRT_E_INVALID; }
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    // AUTO: BVH scenes (book covers, meshes) go to the wavefront kernels, whose extend step copes with long, uneven
    // traversals and many material classes; tiny list scenes (Cornell) to the megakernel.  Event counters exist only
    // in the megakernel.
    const int variant = grt_internal_resolve_variant(h, opt, d_stats != nullptr);
    if (variant == GRT_VARIANT_WAVEFRONT) return grt_render_wavefront(h, cam, opt, d_rgb_sum, st, d_stats);
    if (variant != GRT_VARIANT_MEGAKERNEL) { grt_set_error("unknown GrtOptions.variant"); return GRT_E_INVALID; }
    RenderParams P;
    memset(&P, 0, sizeof(P));
    int rc = grt_make_dev_camera(cam, &P.cam);
    if (rc) return rc;
    P.scene = h->ds;
    P.k0 = (uint32_t)opt->seed; P.k1 = (uint32_t)(opt->seed >> 32);
    uint32_t S2 = (uint32_t)cam->spp_sqrt * (uint32_t)cam->spp_sqrt;
    uint32_t stride = opt->sample_stride ? opt->sample_stride : 1u;
    if (opt->sample_first >= S2) return GRT_OK;   // nothing to do for this shard
    P.sample_first = opt->sample_first; P.sample_stride = stride;
    P.n_my = (S2 - opt->sample_first + stride - 1) / stride;
    int x0 = opt->x0, y0 = opt->y0, x1 = opt->x1, y1 = opt->y1;
    if (x0 == 0 && y0 == 0 && x1 == 0 && y1 == 0) { x1 = cam->width; y1 = cam->height; }
    if (x0 < 0 || y0 < 0 || x1 > cam->width || y1 > cam->height || x1 <= x0 || y1 <= y0) { grt_set_error("GrtOptions pixel window out of range"); return GRT_E_INVALID; }
    P.x0 = x0; P.y0 = y0; P.ww = x1 - x0; P.wh = y1 - y0;
    P.n_pixels = (uint32_t)P.ww * (uint32_t)P.wh;
    // pixels per claim: enough work per atomic, small enough for a short tail
    uint64_t paths_per_pixel = P.n_my;
    // (one claim is the unit of the end-of-frame tail: 8192 paths per claim left a 2-3 ms tail, 3 % of a 62 ms shard
    // frame at 8 GPUs; an atomic per 1024 paths is still free)
    static const uint32_t claim_paths = [] { const char* e = getenv("GRT_CLAIM_PATHS"); int v = e ? atoi(e) : 0; return (uint32_t)(v > 0 ? v : 1024); }();
    uint32_t claim = (uint32_t)(claim_paths / (paths_per_pixel ? paths_per_pixel : 1));
    if (claim < 1) claim = 1;
    if (claim > 64) claim = 64;
    // small images: every resident warp must get several claims, or most of the machine idles
    uint32_t per_warp = P.n_pixels / ((uint32_t)h->sm_count * GRT_MEGA_MIN_BLOCKS * (GRT_MEGA_THREADS / 32) * 4u);
    if (per_warp < 1) per_warp = 1;
    if (claim > per_warp) claim = per_warp;
    P.claim = claim;
    {   // slice exit threshold in sixteenths; the result does not depend on it (tests/test_gpu_parity.py)
        static const int exit16 = [] { const char* e = getenv("GRT_TRAV_EXIT16"); int b = e ? atoi(e) : -1; return b >= 0 ? b : GRT_TRAV_EXIT16; }();
        P.trav_exit16 = exit16;
    }
    P.rgb_sum = d_rgb_sum;
    P.atomic_sum = (opt->flags & GRT_OPT_ATOMIC_SUM) ? 1 : 0;
    P.counter = h->d_counter;
    P.stats = d_stats;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, 4, st));
    bool stats = (opt->flags & GRT_OPT_STATS) && d_stats;
    uint32_t f = h->ds.features | (cam->defocus_angle > 0 ? F_DEFOCUS : 0u);
    const bool timing = (opt->flags & GRT_OPT_TIMING) != 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (timing) { CUDA_TRY(cudaEventCreate(&t0)); CUDA_TRY(cudaEventCreate(&t1)); CUDA_TRY(cudaEventRecord(t0, st)); }
    if ((f & ~V_CORNELL) == 0) rc = launch_mega<V_CORNELL>(h, P, stats, st);
    else if ((f & ~V_SMOKE) == 0) rc = launch_mega<V_SMOKE>(h, P, stats, st);
    else if ((f & ~V_SPHERES) == 0) rc = launch_mega<V_SPHERES>(h, P, stats, st);
    else if ((f & ~V_MESH) == 0) rc = launch_mega<V_MESH>(h, P, stats, st);
    else if ((f & ~V_FULL_UNIQ) == 0) rc = launch_mega<V_FULL_UNIQ>(h, P, stats, st);
    else rc = launch_mega<V_FULL>(h, P, stats, st);
    if (timing) {
        if (!rc) {
            GrtTiming T;
            memset(&T, 0, sizeof(T));
            float ms = 0;
            if (cudaEventRecord(t1, st) == cudaSuccess && cudaEventSynchronize(t1) == cudaSuccess) cudaEventElapsedTime(&ms, t0, t1);
            T.total_ms = T.extend_ms = ms; T.extend_launches = T.launches = 1;
            grt_internal_set_timing(T);
        }
        cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
    return rc;
}

extern "C" int grt_tonemap_device(const float* d_rgb_sum, uint8_t* d_rgb8, uint64_t n_values, float scale, void* stream) {
    if (!d_rgb_sum || !d_rgb8) { grt_set_error("grt_tonemap_device: NULL argument"); return GRT_E_INVALID; }
    if (!n_values) return GRT_OK;
    tonemap_kernel<<<(unsigned)((n_values + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_rgb_sum, d_rgb8, n_values, scale);
    grt_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GRT_OK;
}

extern "C" int grt_render(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt, float* rgb_sum, uint8_t* rgb8, GrtStats* stats) {
    if (!h || !cam || !opt || !rgb_sum) { grt_set_error("grt_render: NULL argument"); return GRT_E_INVALID; }
    if (cam->width <= 0 || cam->height <= 0) { grt_set_error("camera: width/height must be positive"); return GRT_E_INVALID; }
    CUDA_TRY(cudaSetDevice(h->device));
    size_t nval = (size_t)cam->width * cam->height * 3;
    float* d_sum = nullptr; uint8_t* d_rgb8 = nullptr; GrtStats* d_stats = nullptr;
    int rc = GRT_OK;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = grt_dev_alloc((void**)&d_sum, nval * sizeof(float))) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_sum, rgb_sum, nval * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if (stats) {
            if ((e = grt_dev_alloc((void**)&d_stats, sizeof(GrtStats))) != cudaSuccess) break;
            if ((e = cudaMemset(d_stats, 0, sizeof(GrtStats))) != cudaSuccess) break;
        }
        GrtOptions o = *opt;
        if (stats) o.flags |= GRT_OPT_STATS;
        rc = grt_render_device(h, cam, &o, d_sum, nullptr, d_stats);
        if (rc) break;
        if (rgb8) {
            if ((e = grt_dev_alloc((void**)&d_rgb8, nval)) != cudaSuccess) break;
            float scale = 1.0f / (float)((double)cam->spp_sqrt * (double)cam->spp_sqrt);
            rc = grt_tonemap_device(d_sum, d_rgb8, nval, scale, nullptr);
            if (rc) break;
        }
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        if ((e = cudaMemcpy(rgb_sum, d_sum, nval * sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (rgb8 && (e = cudaMemcpy(rgb8, d_rgb8, nval, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (stats && (e = cudaMemcpy(stats, d_stats, sizeof(GrtStats), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    grt_dev_free(d_sum); grt_dev_free(d_rgb8); grt_dev_free(d_stats);
    if (e != cudaSuccess) { grt_set_error(std::string("grt_render: ") + cudaGetErrorString(e)); return GRT_E_CUDA; }
    return rc;
}
