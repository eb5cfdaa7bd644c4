/*
 * grt.h — C ABI of libgrt_cuda: the B200 (sm_100a) backend for go_raytracer's
 * per-pixel Monte Carlo render loop.
 *
 * This is the drop-in boundary.  Everything above it (scene construction, BVH
 * build, flattening, camera set-up, PPM text) is host code in the reference's
 * own vocabulary; everything below it is hand-written CUDA.  The entry points
 * are what a cgo package `internal/cuda` would bind:
 *
 *   grt_render*      replaces  (*Camera).Render's renderer goroutines
 *                    reference: internal/camera/camera.go:90-153 (renderRow,
 *                    threadedRenderer, syncRenderer), :256-290 (getRay),
 *                    :293-341 (rayColor, clampContribution)
 *   grt_trace_batch  replaces  world.Hit(r, [tmin,tmax], &rec)
 *                    reference: internal/hittable/bvh.go:69, hittable.go:122,
 *                    objects.go:83,167,408, medium.go:27, aabb/aabb.go:90
 *   grt_tonemap_*    replaces  Vec3.PrintColor's numeric part
 *                    reference: internal/vec/color.go:14-46
 *
 * All structs are plain-old-data, little-endian, naturally aligned.  Host
 * buffers are caller-allocated and caller-owned.  Every function returns 0 on
 * success or a negative GRT_E_* code; grt_last_error() gives the text.  Nothing
 * here aborts the process (the reference's log.Fatal sites become error codes).
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with GRT_E_NO_DEVICE.
 */
#ifndef GRT_H
#define GRT_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRT_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------- */
#define GRT_OK            0
#define GRT_E_INVALID    -1   /* bad argument / malformed scene            */
#define GRT_E_NO_DEVICE  -2   /* no CUDA device, or device index invalid   */
#define GRT_E_CUDA       -3   /* CUDA runtime error (text in last_error)   */
#define GRT_E_UNSUPPORTED -4  /* scene uses something the backend rejects  */
#define GRT_E_NCCL       -5   /* NCCL error on the in-process multi-GPU path */

/* ---- child references (BVH children, list items, light refs) ------------
 * ref = (type << 28) | index.  The reference's Hittable interface value
 * (hittable.go:60) becomes this tagged index. */
#define GRT_REF_SHIFT   28u
#define GRT_REF_MASK    0x0FFFFFFFu
#define GRT_REF_TYPE(ref) (((ref) >> GRT_REF_SHIFT) & 7u)
#define GRT_REF_NODE    0u   /* BVHNode              bvh.go:12       */
#define GRT_REF_SPHERE  1u   /* sphere               objects.go:14   */
#define GRT_REF_QUAD    2u   /* quad                 objects.go:117  */
#define GRT_REF_TRI     3u   /* Triangle             objects.go:242  */
#define GRT_REF_LIST    4u   /* HittableList: index of first item in items[] */
#define GRT_REF_MEDIUM  5u   /* constantMedium       medium.go:13    */
#define GRT_REF_BOX     6u   /* NewBox: six quads tested as one slab test  objects.go:208-240 */
#define GRT_REF_NONE    7u   /* e.g. an empty HittableList: never hits */
#define GRT_MAKE_REF(type, idx) (((uint32_t)(type) << GRT_REF_SHIFT) | ((uint32_t)(idx) & GRT_REF_MASK))
#define GRT_LIST_LAST   0x80000000u  /* bit 31 of an items[] word: last item of its list */

/* ---- BVH node: 32 bytes, 32-byte aligned, reference child order ----------
 * Nodes are stored in depth-first, left-first order (the order BVHNode.Hit
 * visits them, bvh.go:69-82).  Boxes are the reference's fp64 boxes rounded
 * OUTWARD to fp32, so the fp32 slab test never rejects a box the fp64 test
 * accepts.  Children may be nodes or primitives; primitives are not box-tested
 * (bvh.go:73,79 call child.Hit directly).
 * Bit 31 of `left` and of `right` together form a 2-bit traversal hint h = left>>31 | (right>>31)<<1:
 * h = 0: visit left then right, always (the reference's order; required when the subtree holds a
 * constantMedium, whose random draw makes the visiting order observable); h = 1..3: the children were
 * split along axis h-1 with `left` on the low side, so a ray with a negative direction component on that
 * axis may visit `right` first (closest hits are order-independent up to exact ties). */
#define GRT_NODE_HINT_BIT 0x80000000u
typedef struct GrtNode {
    float    bmin[3];
    float    bmax[3];
    uint32_t left;
    uint32_t right;
} GrtNode;

/* ---- primitives ----------------------------------------------------------
 * Instances (translate / rotateY, transformation.go) are baked into world
 * space by the flattener in fp64.  `id` is the caller's object id for the
 * source primitive and is what grt_trace_batch reports. */
typedef struct GrtSphere {           /* 64 B */
    double   c0[3];                  /* centre at time 0  (Center.At(0))   */
    double   r;                      /* radius                              */
    float    dc[3];                  /* centre motion per unit time         */
    uint32_t mat;
    uint32_t id;
    uint32_t flags;                  /* reserved                            */
    float    uvrot[2];               /* cos,sin of the baked rotateY angle: sphere UV is
                                        computed from the OBJECT-space normal (objects.go:113) */
} GrtSphere;

#define GRT_QUAD_AXIS_ALIGNED 1u     /* normal has a single non-zero component */
typedef struct GrtQuad {             /* 96 B */
    float    n[3];  float D;         /* unit normal, D = n.Q      objects.go:134-135 */
    float    Q[3];  uint32_t flags;
    float    A[3];  uint32_t mat;    /* alpha = A.(p-Q), A = v x w  (== w.(p x v), objects.go:187) */
    float    B[3];  uint32_t id;     /* beta  = B.(p-Q), B = w x u  (== w.(u x p), objects.go:188) */
    double   n64[3]; double D64;     /* fp64 plane, used when not axis-aligned */
} GrtQuad;

/* NewBox (objects.go:208-240) builds six quads and a BVH over them.  The six quads are still emitted
 * (quads[first_quad .. first_quad+5] in the reference's order front, right, back, left, top, bottom) and a hit
 * reports the quad that was hit; GrtBox only lets the traversal find that quad with one slab test in the
 * box's own frame instead of six plane tests.  world = R(obj) + T, R(v) = (rc v.x + rs v.z, v.y, -rs v.x + rc v.z). */
typedef struct GrtBox {              /* 64 B */
    float    mn[3]; uint32_t first_quad;
    float    mx[3]; uint32_t flags;
    float    T[3];  float rc;
    float    rs;    float pad[3];
} GrtBox;

#define GRT_TRI_HAS_NORMALS 1u
#define GRT_TRI_HAS_UV      2u
typedef struct GrtTri {              /* 48 B, intersection data */
    float    v0[3]; uint32_t mat;
    float    e0[3]; uint32_t id;     /* v1-v0 */
    float    e1[3]; uint32_t flags;  /* v2-v0 */
} GrtTri;
typedef struct GrtTriShade {         /* 64 B, read only on a confirmed hit  */
    float    n0[3], n1[3], n2[3];    /* vertex normals                      */
    float    uv[6];                  /* texCoords[3][2]                     */
    float    pad;
} GrtTriShade;

typedef struct GrtMedium {           /* constantMedium, medium.go:13-18     */
    uint32_t boundary;               /* ref of the boundary sub-tree        */
    float    neg_inv_density;
    uint32_t mat;                    /* isotropic phase function            */
    uint32_t id;
} GrtMedium;

/* ---- materials / textures (tagged unions) -------------------------------- */
#define GRT_MAT_LAMBERTIAN    0u     /* materials.go:30  */
#define GRT_MAT_METAL         1u     /* materials.go:61  */
#define GRT_MAT_DIELECTRIC    2u     /* materials.go:85  */
#define GRT_MAT_DIFFUSE_LIGHT 3u     /* materials.go:132 */
#define GRT_MAT_ISOTROPIC     4u     /* materials.go:157 */
typedef struct GrtMaterial {         /* 32 B */
    uint32_t type;
    uint32_t tex;                    /* lambertian / light / isotropic      */
    float    albedo[3];              /* metal                               */
    float    fuzz;                   /* metal                               */
    float    ior;                    /* dielectric                          */
    uint32_t pad;
} GrtMaterial;

#define GRT_TEX_SOLID   0u           /* texture.go:14  */
#define GRT_TEX_CHECKER 1u           /* texture.go:29  */
#define GRT_TEX_IMAGE   2u           /* texture.go:62  */
#define GRT_TEX_NOISE   3u           /* texture.go:98  */
#define GRT_NOISE_PERLIN    1u       /* texture.go:93-96 */
#define GRT_NOISE_MARBLE    2u
#define GRT_NOISE_TURBULENT 3u
typedef struct GrtTexture {          /* 32 B */
    uint32_t type;
    float    color[3];               /* solid                               */
    float    scale;                  /* checker: inv_scale; noise: scale    */
    uint32_t even;                   /* checker: texture ids                */
    uint32_t odd;
    uint32_t aux;                    /* image: image index; noise: perlin index | variant<<16 */
} GrtTexture;

typedef struct GrtImage {
    uint32_t width, height;
    uint64_t offset;                 /* byte offset of RGB8 texel 0 in texels[] */
} GrtImage;

/* One Perlin generator (perlin.go:12-17): 256 unit gradients + 3 permutations. */
typedef struct GrtPerlin {
    float    grad[256][4];           /* xyz, w unused                       */
    uint8_t  perm[3][256];           /* permX, permY, permZ                 */
} GrtPerlin;

/* ---- lights (the `lights` argument of Camera.Render) ---------------------
 * PdfValue / Random of the reference re-intersect and sample the light in
 * fp64; the backend keeps that arithmetic in fp64. */
#define GRT_LIGHT_SPHERE 1u
#define GRT_LIGHT_QUAD   2u
#define GRT_LIGHT_TRI    3u
typedef struct GrtLight {            /* 304 B */
    uint32_t type;
    uint32_t prim;                   /* ref of the primitive in the world (or NONE) */
    uint32_t flags;                  /* tri: GRT_TRI_HAS_NORMALS            */
    uint32_t pad;
    /* sphere: c[3], r
     * quad:   Q[3], u[3], v[3], n[3], w[3], D, area
     * tri:    v0[3], v1[3], v2[3], area, n0[3], n1[3], n2[3]               */
    double   p[24];
    float    f[24];                  /* p[] rounded to fp32: fast path, falls back to p[] near an edge */
} GrtLight;

#define GRT_LIGHTS_LIST 0u           /* lights is a HittableList (hittable.go:89-103) */
#define GRT_LIGHTS_BARE 1u           /* lights is a single primitive (main.go:274)    */

/* ---- the flattened scene -------------------------------------------------- */
typedef struct GrtScene {
    uint32_t abi_version;
    uint32_t root;                   /* ref of `world`                      */

    const GrtNode*     nodes;      uint32_t n_nodes;
    const GrtSphere*   spheres;    uint32_t n_spheres;
    const GrtQuad*     quads;      uint32_t n_quads;
    const GrtBox*      boxes;      uint32_t n_boxes;
    const GrtTri*      tris;       uint32_t n_tris;
    const GrtTriShade* tri_shade;  /* n_tris entries, or NULL if no tri has normals/uv */
    const double*      tri_v64;    /* n_tris x 9 doubles (v0,v1,v2), or NULL: fp64 vertices used to refine the
                                      WINNING triangle's t (fp32 Moller-Trumbore loses relative accuracy when the
                                      origin is almost coplanar, e.g. a ray leaving a neighbouring mesh triangle) */
    const uint32_t*    items;      uint32_t n_items;   /* HittableList items: child refs in list order, GRT_LIST_LAST on each list's last */
    const GrtMedium*   media;      uint32_t n_media;
    const GrtMaterial* materials;  uint32_t n_materials;
    const GrtTexture*  textures;   uint32_t n_textures;
    const GrtImage*    images;     uint32_t n_images;
    const uint8_t*     texels;     uint64_t n_texel_bytes;
    const GrtPerlin*   perlins;    uint32_t n_perlins;
    const GrtLight*    lights;     uint32_t n_lights;
    uint32_t           lights_mode;
    uint32_t           max_depth_hint;  /* deepest traversal stack the flattener saw */
} GrtScene;

/* ---- camera: the DERIVED state of Camera.initialize (camera.go:179-253) --
 * computed on the host in fp64 exactly as the reference does, so the device
 * never re-derives it in another precision. */
typedef struct GrtCamera {
    int32_t  width, height;          /* Width, imageHeight                  */
    int32_t  spp_sqrt;               /* sppSqrt = int(sqrt(SamplesPerPixel))*/
    int32_t  max_depth;              /* MaxDepth                            */
    double   center[3];
    double   pixel00[3];             /* pixel00Loc                          */
    double   delta_u[3], delta_v[3]; /* pixelDeltaU/V                       */
    double   defocus_u[3], defocus_v[3];
    double   defocus_angle;          /* DefocusAngle (<=0: pinhole)         */
    double   background[3];
    double   max_contribution;       /* MaxContribution (firefly clamp)     */
} GrtCamera;

#define GRT_VARIANT_MEGAKERNEL 0
#define GRT_VARIANT_WAVEFRONT  1
#define GRT_VARIANT_AUTO       2   /* wavefront for scenes with a BVH (measured faster there), megakernel otherwise */
typedef struct GrtOptions {
    uint64_t seed;                   /* Philox key                          */
    int32_t  variant;                /* GRT_VARIANT_*                       */
    int32_t  device;                 /* CUDA device ordinal                 */
    /* strata subset rendered by this call: samples s = first + k*stride,
     * s < spp_sqrt^2 (multi-GPU sharding splits the strata set). */
    uint32_t sample_first;
    uint32_t sample_stride;          /* 0 is treated as 1                   */
    /* optional pixel window (0,0,0,0 = whole image)                         */
    int32_t  x0, y0, x1, y1;
    uint32_t flags;                  /* GRT_OPT_*                           */
    uint32_t pad;
} GrtOptions;
#define GRT_OPT_STATS 1u             /* fill GrtStats (slower: counts events) */
#define GRT_OPT_TIMING 4u            /* time each kernel class of the render with CUDA events on the launching stream
                                        (the wavefront variant then issues plain launches instead of its CUDA graph);
                                        read the result with grt_last_timing() after synchronising */
#define GRT_OPT_ATOMIC_SUM 2u        /* accumulate into rgb_sum with system-scope atomics: several renders, also from
                                        peer GPUs over NVLink, may then share one buffer (grt_render_multi does) */

typedef struct GrtStats {            /* event counters for the roofline table */
    uint64_t paths, segments;        /* rayColor calls at depth=MaxDepth / all */
    uint64_t box_tests, sphere_tests, quad_tests, tri_tests, medium_tests;
    uint64_t shade_diffuse, shade_specular, light_pdf_evals;
    uint64_t nan_samples;
    uint64_t warp_iterations;        /* megakernel: loop iterations summed over warps          */
    uint64_t lane_iterations;        /* ... and the number of lanes that traced a segment in them */
} GrtStats;

/* ---- ray batch (parity checks 1 and 2) ----------------------------------- */
typedef struct GrtRay {
    float    o[3]; float tmin;
    float    d[3]; float tmax;
    float    time;
    uint32_t self_id;                /* object id of the primitive the origin lies on, or GRT_NO_ID */
    uint32_t pad[2];
} GrtRay;
#define GRT_NO_ID 0xFFFFFFFFu
typedef struct GrtHit {
    float    t;                      /* +inf on miss                         */
    uint32_t id;                     /* caller object id, GRT_NO_ID on miss  */
    uint32_t ref;                    /* flat primitive / medium ref          */
    uint32_t front_face;
    float    p[3]; float u;
    float    n[3]; float v;
} GrtHit;

typedef struct GrtSceneDev* GrtSceneHandle;

/* ---- entry points -------------------------------------------------------- */
int         grt_abi_version(void);
int         grt_device_count(void);
const char* grt_last_error(void);

/* Copies the flat scene to HBM on `device` (small scenes are additionally
 * staged into shared memory by the kernels).  The GrtScene may be freed after
 * the call returns.  A handle carries per-render scratch (the pixel counter,
 * the wavefront path pool): run one render at a time per handle; different
 * handles, also on one device, are independent.  Device memory comes from a
 * process-wide pool and returns to it on grt_scene_free. */
int grt_scene_upload(const GrtScene* scene, int device, GrtSceneHandle* out);
int grt_scene_free(GrtSceneHandle h);

/* Closest-hit query for n rays; rays/hits are HOST buffers. */
int grt_trace_batch(GrtSceneHandle h, const GrtRay* rays, uint64_t n, GrtHit* hits);
/* Same with DEVICE buffers on `stream` (a cudaStream_t, may be NULL). */
int grt_trace_batch_device(GrtSceneHandle h, const GrtRay* d_rays, uint64_t n, GrtHit* d_hits, void* stream);

/* Renders the strata subset selected by `opt` and ADDS the per-pixel radiance
 * sums into rgb_sum (HOST, width*height*3 floats, row-major, caller zeroes
 * it).  If rgb8 is non-NULL also writes the quantised image
 * (color.go:23-46) using scale 1/spp_sqrt^2 — only meaningful when the call
 * covers all strata.  stats may be NULL. */
int grt_render(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt,
               float* rgb_sum, uint8_t* rgb8, GrtStats* stats);

/* Device-resident variant: accumulates into d_rgb_sum (DEVICE, same layout)
 * asynchronously on `stream`.  This is what a one-process-per-GPU driver
 * calls before its ncclReduce. */
int grt_render_device(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt,
                      float* d_rgb_sum, void* stream, GrtStats* d_stats);

/* rgb8[i] = int(clamp(sqrt(max(x,0)), 0, 0.99999) * 256), NaN -> 0, x = sum*scale
 * (color.go:14-46).  DEVICE buffers. */
int grt_tonemap_device(const float* d_rgb_sum, uint8_t* d_rgb8, uint64_t n_values,
                       float scale, void* stream);

/* Device time per kernel class of the calling thread's last render that had GRT_OPT_TIMING set (CUDA events on the
 * launching stream; measurement support for bench.py's roofline).  Megakernel renders report everything
 * as `extend_ms` = `total_ms` (one kernel does the whole path). */
typedef struct GrtTiming {
    double   total_ms;               /* first launch .. last launch of the render                         */
    double   generate_ms;            /* camera rays            camera.go:256-290                          */
    double   extend_ms;              /* BVH traversal + primitive intersection (the dominant kernel)      */
    double   shade_ms;               /* BSDF scatter, PDF mixture, clamp unwind, accumulation             */
    uint64_t extend_launches;
    uint64_t launches;
    uint64_t extend_kernel;          /* which kernel `extend_ms` timed: 0 render_mega_kernel, 1 wf_extend (one thread per
                                        slot), 2 wf_extend_dyn (persistent warps) — the library picks per scene         */
} GrtTiming;
int grt_last_timing(GrtTiming* out);

/* In-process multi-GPU render: replicates the scene on devices[0..n), splits
 * the strata set s = g mod n, and combines the fp32 sums with one ncclReduce
 * to devices[0] (NCCL is dlopen'ed; GRT_E_NCCL if unavailable).  Outputs as
 * grt_render. */
int grt_render_multi(const GrtScene* scene, const GrtCamera* cam, const GrtOptions* opt,
                     const int* devices, int n_devices,
                     float* rgb_sum, uint8_t* rgb8, double* kernel_ms);

/* BuildBVH's object order on the GPU: replaces the per-node sort.Slice of bvhHelper (bvh.go:35-61, boxCompare
 * bvh.go:25-32, LongestAxis aabb.go:73-87) for large lists.  boxes = n x {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z}, the
 * objects' BBox() in list order; order_out[p] = list index of the object that stands at position p after all
 * recursive sorts (ties keep list order).  The tree's shape follows from n alone (median splits). */
int grt_bvh_order(const double* boxes, uint32_t n, int device, uint32_t* order_out);

/* Kernel launch count since library load (bench.py's gpu_launches claim). */
uint64_t grt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GRT_H */
