/*
 * grt_host.h — host-side mirror of the reference's scene/camera vocabulary,
 * as a C API over the C++ implementation (scene_ir.hpp, scenes.hpp,
 * flatten.hpp).  The Go toolchain is absent in this environment, so the
 * reference-side host code (main.go scene functions, package hittable
 * constructors, Camera.initialize / Render) is mirrored here in C++ with the
 * same names and argument meaning; INTEGRATION.md shows the Go originals next
 * to the cgo binding a maintainer would add.
 *
 * All functions return an id / 0 on success or a negative number on failure;
 * grt_host_last_error() gives the text (the reference's log.Fatal sites).
 */
#ifndef GRT_HOST_H
#define GRT_HOST_H
#include "grt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct GrtHostScene GrtHostScene;

/* Public fields of camera.Camera (camera.go:24-36) and PositionCamera (:65). */
typedef struct GrtCameraConfig {
    double  AspectRatio;
    int32_t Width, SamplesPerPixel, MaxDepth, MaxThreads;
    double  VerticalFOV, DefocusAngle, FocusDistance;
    double  Background[3];
    double  MaxContribution;
    double  lookFrom[3], lookAt[3], vup[3];
} GrtCameraConfig;

typedef struct GrtSceneOptions {         /* overrides for the built-in scenes */
    int32_t  width, spp;                 /* 0 = value shipped in main.go      */
    double   aspect;                     /* 0 = shipped                       */
    uint64_t seed;                       /* 0 = per-scene default             */
    int32_t  mesh_segments;              /* scene 8: UV-sphere segments       */
    int32_t  image_w, image_h;           /* decoded earthmap for scenes 2, 5  */
    const uint8_t* image_rgb;
} GrtSceneOptions;

const char* grt_host_last_error(void);
GrtHostScene* grt_host_scene_new(void);
void grt_host_scene_free(GrtHostScene* s);

/* textures — texture.go */
int grt_host_solid_color(GrtHostScene* s, double r, double g, double b);                 /* NewSolidColor            */
int grt_host_checkerboard(GrtHostScene* s, double scale, int even_tex, int odd_tex);     /* NewCheckerboard          */
int grt_host_image(GrtHostScene* s, int w, int h, const uint8_t* rgb);                   /* decoded RTImage          */
/* LoadImage for a JPEG held in memory, with the arithmetic of Go's image/jpeg and color.YCbCr (imageLoader.go:28-84):
 * call with rgb == NULL to get the size, then with a width*height*3 buffer.  Baseline JPEG only. */
int grt_host_decode_jpeg(const uint8_t* data, size_t n, int* width, int* height, uint8_t* rgb, size_t cap);
int grt_host_image_texture(GrtHostScene* s, int image);                                  /* NewImageTexture          */
int grt_host_noise_texture(GrtHostScene* s, double scale, int variant, uint64_t seed);   /* NewNoiseTextureWithType  */
/* materials — materials.go */
int grt_host_lambertian(GrtHostScene* s, int tex);                                       /* NewTexturedLambertian    */
int grt_host_metal(GrtHostScene* s, double r, double g, double b, double fuzz);          /* NewMetal                 */
int grt_host_dielectric(GrtHostScene* s, double ior);                                    /* NewDielectric            */
int grt_host_diffuse_light(GrtHostScene* s, int tex);                                    /* NewDiffuseLightTextured  */
int grt_host_isotropic(GrtHostScene* s, int tex);                                        /* NewIsotropicTexture      */
/* hittables — objects.go, hittable.go, bvh.go, transformation.go, medium.go */
int grt_host_sphere(GrtHostScene* s, const double c[3], double r, int mat);              /* NewSphere                */
int grt_host_motion_sphere(GrtHostScene* s, const double c1[3], const double c2[3], double r, int mat);  /* NewMotionSphere */
int grt_host_quad(GrtHostScene* s, const double Q[3], const double u[3], const double v[3], int mat);    /* NewQuad         */
int grt_host_box(GrtHostScene* s, const double a[3], const double b[3], int mat);        /* NewBox                   */
int grt_host_triangle(GrtHostScene* s, const double v[9], const double* n9, const double* uv6, int mat); /* NewTriangle*    */
int grt_host_list(GrtHostScene* s);                                                      /* NewHittableList          */
int grt_host_list_add(GrtHostScene* s, int list, int obj);                               /* (*HittableList).Add      */
int grt_host_bvh(GrtHostScene* s, int list);                                             /* BuildBVH                 */
int grt_host_translate(GrtHostScene* s, int obj, const double off[3]);                   /* Translate                */
int grt_host_rotate_y(GrtHostScene* s, int obj, double degrees);                         /* RotateY                  */
int grt_host_constant_medium(GrtHostScene* s, int boundary, double density, int tex);    /* ConstantMediumTexture    */
int grt_host_set_world(GrtHostScene* s, int obj);                                        /* Render's `world`         */
int grt_host_set_lights(GrtHostScene* s, int obj);                                       /* Render's `lights`        */

/* internal/objLoader: LoadObjWithOptions on in-memory text (objLoader.go:72; mtl_text may be NULL/empty).
 * model_out receives the BVH over the triangles, lights_out the list of emissive triangles (objLoader.go:489-512).
 * default_mat < 0: Lambertian(0.8).  Image maps (map_Kd/map_Ka) are not resolved here: a material that needs one
 * makes the load fail, as a missing file does in the reference. */
typedef struct GrtObjOptions {
    double  ScaleFactor;
    int32_t FlipYZ, IgnoreNormals, Center, FlipFaces, IgnoreMtl, FindWindows;
    double  Position[3];
    int32_t DefaultMaterial;
} GrtObjOptions;
int grt_host_load_obj(GrtHostScene* s, const char* obj_text, const char* mtl_text, const GrtObjOptions* opt,
                      int* model_out, int* lights_out, int* n_triangles_out);

/* The same from a file: `mtllib` is resolved next to the OBJ as objLoader.go:117-125 does; a missing MTL file falls
 * back to the default material (objLoader.go:136-139). */
int grt_host_load_obj_file(GrtHostScene* s, const char* path, const GrtObjOptions* opt,
                           int* model_out, int* lights_out, int* n_triangles_out);

/* main.go's scene functions, -S 1..8 (main.go:449-476); fills *cam like the scene function does. */
int grt_host_builtin_scene(GrtHostScene* s, int scene_id, const GrtSceneOptions* opt, GrtCameraConfig* cam);

/* hittable.Flatten: BuildBVH + bake + flatten.  *out points into storage owned by s
 * (valid until the next flatten or grt_host_scene_free). */
int grt_host_flatten(GrtHostScene* s, GrtScene* out);
/* Same with explicit tuning: a BuildBVH result with <= collapse_whole leaves, and inside larger trees any
 * subtree with <= collapse_leaf leaves, is emitted as an ordered run of its leaves (visiting order of
 * bvh.go:69-82) instead of nodes.  (0, 0) reproduces the reference's tree node for node.  Closest-hit
 * results do not depend on these values. */
int grt_host_flatten_opts(GrtHostScene* s, int collapse_whole, int collapse_leaf, GrtScene* out);
/* Camera.initialize (camera.go:179-253). */
int grt_host_camera_derive(const GrtCameraConfig* cfg, GrtCamera* out);
/* P3 text exactly as camera.go:160 + color.go:45 write it; returns bytes written or -1. */
long grt_host_write_ppm(const uint8_t* rgb8, int width, int height, char* out, long cap);
/* The same pixels in binary containers (the reference's README lists other output formats as future work,
 * README.md:63): P6 ("P6\n%d %d\n255\n" + raw RGB) and PNG (8-bit RGB, stored deflate blocks; cap >= 100 + 1.001 *
 * (3*width+1)*height).  Return bytes written or -1. */
long grt_host_write_p6(const uint8_t* rgb8, int width, int height, unsigned char* out, long cap);
long grt_host_write_png(const uint8_t* rgb8, int width, int height, unsigned char* out, long cap);

/* Camera.Render(world, lights) behind the CUDA backend: flatten, upload, render
 * on n_gpus devices (strata split + ncclReduce when n_gpus > 1), tonemap, P3
 * text into ppm_out (may be NULL).  rgb_sum_out (may be NULL) receives the
 * per-pixel radiance sums.  Returns 0 or a GRT_E_* code. */
int grt_host_camera_render(GrtHostScene* s, const GrtCameraConfig* cfg, uint64_t seed, int variant, int n_gpus,
                           float* rgb_sum_out, char* ppm_out, long ppm_cap, long* ppm_len, double* kernel_ms);

/* Same, also handing back the quantised pixels (rgb8_out: width*height*3 bytes, may be NULL). */
int grt_host_camera_render_rgb8(GrtHostScene* s, const GrtCameraConfig* cfg, uint64_t seed, int variant, int n_gpus,
                                float* rgb_sum_out, uint8_t* rgb8_out, char* ppm_out, long ppm_cap, long* ppm_len, double* kernel_ms);

/* Opaque pointer to the scene description, consumed by the test oracle only. */
const void* grt_host_scene_description(GrtHostScene* s);

#ifdef __cplusplus
}
#endif
#endif
