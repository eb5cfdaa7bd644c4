// host.cpp — C API over the host-side mirror (include/grt_host.h).
#include "jpeg_go.hpp"
#include "../../include/grt_host.h"
#include "scene_ir.hpp"
#include "scenes.hpp"
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <stdio.h>
#include "flatten.hpp"
#include "obj_loader.hpp"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <memory>

using grt::ir::V3;

static thread_local std::string g_host_error;
static int fail(const std::string& e) { g_host_error = e; return -1; }

struct GrtHostScene {
    grt::ir::Scene ir;
    std::unique_ptr<grt::flat::FlatScene> flat;
};

static V3 v3(const double* p) { return V3(p[0], p[1], p[2]); }

#define GUARD(expr)                                   \
    try { return (expr); }                            \
    catch (const std::exception& ex) { return fail(ex.what()); }

extern "C" {

const char* grt_host_last_error(void) { return g_host_error.c_str(); }
GrtHostScene* grt_host_scene_new(void) { return new GrtHostScene(); }
void grt_host_scene_free(GrtHostScene* s) { delete s; }

int grt_host_solid_color(GrtHostScene* s, double r, double g, double b) { GUARD(s->ir.NewSolidColor(V3(r, g, b))); }
int grt_host_checkerboard(GrtHostScene* s, double scale, int even_tex, int odd_tex) {
    if (even_tex < 0 || odd_tex < 0 || even_tex >= (int)s->ir.textures.size() || odd_tex >= (int)s->ir.textures.size()) return fail("bad texture id");
    GUARD(s->ir.NewCheckerboard(scale, even_tex, odd_tex));
}
int grt_host_image(GrtHostScene* s, int w, int h, const uint8_t* rgb) {
    if (w <= 0 || h <= 0 || !rgb) return fail("bad image");
    GUARD(s->ir.AddImage(w, h, rgb));
}
// imageLoader.LoadImage (imageLoader.go:28-47) for JPEG files: Go's image/jpeg + color.YCbCr arithmetic (jpeg_go.hpp), so
// the texels are the ones the reference renders with.  Two calls: rgb == NULL reports the size, then with a buffer.
int grt_host_decode_jpeg(const uint8_t* data, size_t n, int* width, int* height, uint8_t* rgb, size_t cap) {
    if (!data || !width || !height) return fail("grt_host_decode_jpeg: NULL argument");
    grt::jpeg::Image img;
    std::string err = grt::jpeg::decode(data, n, img);
    if (!err.empty()) return fail("jpeg: " + err);
    *width = img.width; *height = img.height;
    if (rgb) {
        if (cap < img.rgb.size()) return fail("grt_host_decode_jpeg: buffer too small");
        memcpy(rgb, img.rgb.data(), img.rgb.size());
    }
    return 0;
}
int grt_host_image_texture(GrtHostScene* s, int image) {
    if (image < 0 || image >= (int)s->ir.images.size()) return fail("bad image id");
    GUARD(s->ir.NewImageTexture(image));
}
int grt_host_noise_texture(GrtHostScene* s, double scale, int variant, uint64_t seed) {
    if (variant < 1 || variant > 3) return fail("bad noise variant");
    GUARD(s->ir.NewNoiseTextureWithType(scale, variant, seed));
}
static bool okTex(GrtHostScene* s, int t) { return t >= 0 && t < (int)s->ir.textures.size(); }
int grt_host_lambertian(GrtHostScene* s, int tex) { if (!okTex(s, tex)) return fail("bad texture id"); GUARD(s->ir.NewTexturedLambertian(tex)); }
int grt_host_metal(GrtHostScene* s, double r, double g, double b, double fuzz) { GUARD(s->ir.NewMetal(V3(r, g, b), fuzz)); }
int grt_host_dielectric(GrtHostScene* s, double ior) { GUARD(s->ir.NewDielectric(ior)); }
int grt_host_diffuse_light(GrtHostScene* s, int tex) { if (!okTex(s, tex)) return fail("bad texture id"); GUARD(s->ir.NewDiffuseLightTextured(tex)); }
int grt_host_isotropic(GrtHostScene* s, int tex) { if (!okTex(s, tex)) return fail("bad texture id"); GUARD(s->ir.NewIsotropicTexture(tex)); }

int grt_host_sphere(GrtHostScene* s, const double c[3], double r, int mat) { GUARD(s->ir.NewSphere(v3(c), r, mat)); }
int grt_host_motion_sphere(GrtHostScene* s, const double c1[3], const double c2[3], double r, int mat) { GUARD(s->ir.NewMotionSphere(v3(c1), v3(c2), r, mat)); }
int grt_host_quad(GrtHostScene* s, const double Q[3], const double u[3], const double v[3], int mat) { GUARD(s->ir.NewQuad(v3(Q), v3(u), v3(v), mat)); }
int grt_host_box(GrtHostScene* s, const double a[3], const double b[3], int mat) { GUARD(s->ir.NewBox(v3(a), v3(b), mat)); }
int grt_host_triangle(GrtHostScene* s, const double v[9], const double* n9, const double* uv6, int mat) {
    V3 vv[3] = {v3(v), v3(v + 3), v3(v + 6)};
    V3 nn[3];
    double uv[3][2];
    if (n9) for (int i = 0; i < 3; i++) nn[i] = v3(n9 + 3 * i);
    if (uv6) for (int i = 0; i < 3; i++) { uv[i][0] = uv6[2 * i]; uv[i][1] = uv6[2 * i + 1]; }
    GUARD(s->ir.NewTriangleFull(vv, n9 ? nn : nullptr, uv6 ? uv : nullptr, mat));
}
int grt_host_list(GrtHostScene* s) { GUARD(s->ir.NewHittableList()); }
int grt_host_list_add(GrtHostScene* s, int list, int obj) {
    try { s->ir.Add(list, obj); return 0; } catch (const std::exception& ex) { return fail(ex.what()); }
}
int grt_host_bvh(GrtHostScene* s, int list) { GUARD(s->ir.BuildBVH(list)); }
int grt_host_translate(GrtHostScene* s, int obj, const double off[3]) { GUARD(s->ir.Translate(obj, v3(off))); }
int grt_host_rotate_y(GrtHostScene* s, int obj, double degrees) { GUARD(s->ir.RotateY(obj, degrees)); }
int grt_host_constant_medium(GrtHostScene* s, int boundary, double density, int tex) {
    if (!okTex(s, tex)) return fail("bad texture id");
    GUARD(s->ir.ConstantMediumTexture(boundary, density, tex));
}
int grt_host_set_world(GrtHostScene* s, int obj) { try { s->ir.checkH(obj); s->ir.world = obj; return 0; } catch (const std::exception& ex) { return fail(ex.what()); } }
int grt_host_set_lights(GrtHostScene* s, int obj) { try { s->ir.checkH(obj); s->ir.lights = obj; return 0; } catch (const std::exception& ex) { return fail(ex.what()); } }

int grt_host_load_obj(GrtHostScene* s, const char* obj_text, const char* mtl_text, const GrtObjOptions* o,
                      int* model_out, int* lights_out, int* n_triangles_out) {
    if (!s || !obj_text) return fail("NULL argument");
    grt::obj::LoadObjOptions opt;
    if (o) {
        opt.ScaleFactor = o->ScaleFactor; opt.FlipYZ = o->FlipYZ != 0; opt.IgnoreNormals = o->IgnoreNormals != 0; opt.Center = o->Center != 0;
        opt.FlipFaces = o->FlipFaces != 0; opt.IgnoreMtl = o->IgnoreMtl != 0; opt.FindWindows = o->FindWindows != 0;
        opt.Position = v3(o->Position); opt.DefaultMaterial = o->DefaultMaterial;
    }
    grt::obj::LoadResult lr;
    std::string err;
    try {
        if (!grt::obj::loadObj(s->ir, obj_text, mtl_text ? mtl_text : "", opt, lr, err)) return fail(err);
    } catch (const std::exception& ex) { return fail(ex.what()); }
    if (model_out) *model_out = lr.model;
    if (lights_out) *lights_out = lr.lights;
    if (n_triangles_out) *n_triangles_out = lr.nTriangles;
    return 0;
}

// LoadObjWithOptions on a file (objLoader.go:72-140): the OBJ is read from `path`; unless IgnoreMtl is set, the first
// `mtllib` line names the material library, looked up next to the OBJ (filepath.Join(filepath.Dir(filename), name),
// objLoader.go:117-125).  A missing MTL file is a warning in the reference (:136-139): the default material is used.
int grt_host_load_obj_file(GrtHostScene* s, const char* path, const GrtObjOptions* o, int* model_out, int* lights_out, int* n_triangles_out) {
    if (!s || !path) return fail("NULL argument");
    auto slurp = [](const std::string& p, std::string& out) -> bool {
        FILE* f = fopen(p.c_str(), "rb");
        if (!f) return false;
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
        fclose(f);
        return true;
    };
    std::string obj, mtl;
    if (!slurp(path, obj)) return fail(std::string("cannot open OBJ file: ") + path);   // objLoader.go:84-87
    if (!(o && o->IgnoreMtl)) {
        size_t pos = 0;
        while (pos < obj.size()) {
            size_t eol = obj.find('\n', pos);
            if (eol == std::string::npos) eol = obj.size();
            size_t a = pos, b = eol;
            while (a < b && isspace((unsigned char)obj[a])) a++;
            while (b > a && isspace((unsigned char)obj[b - 1])) b--;
            if (b - a > 7 && obj.compare(a, 6, "mtllib") == 0 && isspace((unsigned char)obj[a + 6])) {
                size_t c = a + 6;
                while (c < b && isspace((unsigned char)obj[c])) c++;
                std::string name;   // strings.Join(strings.Fields(line)[1:], " ")
                bool gap = false;
                for (size_t i = c; i < b; i++) {
                    if (isspace((unsigned char)obj[i])) { gap = true; continue; }
                    if (gap && !name.empty()) name += ' ';
                    gap = false;
                    name += obj[i];
                }
                std::string dir(path);
                size_t slash = dir.find_last_of('/');
                dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
                slurp(dir + "/" + name, mtl);   // absent: continue with the default material
                break;
            }
            pos = eol + 1;
        }
    }
    return grt_host_load_obj(s, obj.c_str(), mtl.empty() ? nullptr : mtl.c_str(), o, model_out, lights_out, n_triangles_out);
}

static void toC(const grt::ir::CameraConfig& c, GrtCameraConfig* o) {
    o->AspectRatio = c.AspectRatio; o->Width = c.Width; o->SamplesPerPixel = c.SamplesPerPixel; o->MaxDepth = c.MaxDepth; o->MaxThreads = c.MaxThreads;
    o->VerticalFOV = c.VerticalFOV; o->DefocusAngle = c.DefocusAngle; o->FocusDistance = c.FocusDistance;
    o->Background[0] = c.Background.x; o->Background[1] = c.Background.y; o->Background[2] = c.Background.z;
    o->MaxContribution = c.MaxContribution;
    const V3* src[3] = {&c.lookFrom, &c.lookAt, &c.vup};
    double* dst[3] = {o->lookFrom, o->lookAt, o->vup};
    for (int i = 0; i < 3; i++) { dst[i][0] = src[i]->x; dst[i][1] = src[i]->y; dst[i][2] = src[i]->z; }
}
static grt::ir::CameraConfig fromC(const GrtCameraConfig* o) {
    grt::ir::CameraConfig c;
    c.AspectRatio = o->AspectRatio; c.Width = o->Width; c.SamplesPerPixel = o->SamplesPerPixel; c.MaxDepth = o->MaxDepth; c.MaxThreads = o->MaxThreads;
    c.VerticalFOV = o->VerticalFOV; c.DefocusAngle = o->DefocusAngle; c.FocusDistance = o->FocusDistance;
    c.Background = v3(o->Background); c.MaxContribution = o->MaxContribution;
    c.lookFrom = v3(o->lookFrom); c.lookAt = v3(o->lookAt); c.vup = v3(o->vup);
    return c;
}

int grt_host_builtin_scene(GrtHostScene* s, int scene_id, const GrtSceneOptions* opt, GrtCameraConfig* cam) {
    if (!s || !cam) return fail("NULL argument");
    grt::scenes::SceneOptions o;
    if (opt) {
        o.width = opt->width; o.spp = opt->spp; o.aspect = opt->aspect; o.seed = opt->seed; o.mesh_segments = opt->mesh_segments;
        o.image_rgb = opt->image_rgb; o.image_w = opt->image_w; o.image_h = opt->image_h;
    }
    grt::ir::CameraConfig c;
    try {
        if (!grt::scenes::buildScene(scene_id, s->ir, c, o)) return fail("unknown scene id (main.go -S accepts 1..8; the default scene is empty)");
    } catch (const std::exception& ex) { return fail(ex.what()); }
    toC(c, cam);
    return 0;
}

int grt_host_flatten(GrtHostScene* s, GrtScene* out) {
    grt::flat::FlattenOptions o;
    return grt_host_flatten_opts(s, o.collapse_whole, o.collapse_leaf, out);
}
int grt_host_flatten_opts(GrtHostScene* s, int collapse_whole, int collapse_leaf, GrtScene* out) {
    if (!s || !out) return fail("NULL argument");
    s->flat.reset(new grt::flat::FlatScene());
    grt::flat::FlattenOptions fo;
    fo.collapse_whole = collapse_whole < 0 ? 0 : collapse_whole; fo.collapse_leaf = collapse_leaf;
    if (collapse_whole == 0 && collapse_leaf == 0) { fo.box_prims = false; fo.order_hints = false; }   // (0, 0): the reference's tree, node for node
    {   // BuildBVH's sorts run on the GPU for large lists when a device is present (GRT_BVH_BUILD=cpu|gpu overrides:
        // "cpu" never, "gpu" for every list of three or more objects)
        const char* e = getenv("GRT_BVH_BUILD");
        const bool never = e && strcmp(e, "cpu") == 0;
        if (!never && grt_device_count() > 0) {
            fo.gpu_order = [](const double* boxes, uint32_t n, uint32_t* order) { return grt_bvh_order(boxes, n, 0, order) == 0; };
            if (e && strcmp(e, "gpu") == 0) fo.gpu_order_min = 3;
        }
    }
    grt::flat::Flattener f(s->ir, fo);
    if (!f.run(*s->flat)) { s->flat.reset(); return fail(f.error); }
    *out = s->flat->view();
    return 0;
}

int grt_host_camera_derive(const GrtCameraConfig* cfg, GrtCamera* out) {
    if (!cfg || !out) return fail("NULL argument");
    std::string err;
    if (!grt::flat::deriveCamera(fromC(cfg), *out, err)) return fail(err);
    return 0;
}

long grt_host_write_ppm(const uint8_t* rgb8, int width, int height, char* out, long cap) {
    if (!rgb8 || !out || cap < 32) return -1;
    long n = snprintf(out, (size_t)cap, "P3\n%d %d\n255\n", width, height);   // camera.go:160
    // "%d %d %d\n" per pixel, color.go:45 — hand-rolled, this is 10^6..10^7 lines
    static const char digits[] = "0123456789";
    for (long i = 0; i < (long)width * height; i++) {
        if (n + 13 > cap) return -1;
        for (int k = 0; k < 3; k++) {
            unsigned v = rgb8[3 * i + k];
            if (v >= 100) { out[n++] = digits[v / 100]; out[n++] = digits[(v / 10) % 10]; out[n++] = digits[v % 10]; }
            else if (v >= 10) { out[n++] = digits[v / 10]; out[n++] = digits[v % 10]; }
            else out[n++] = digits[v];
            out[n++] = k == 2 ? '\n' : ' ';
        }
    }
    if (n < cap) out[n] = 0;
    return n;
}

long grt_host_write_p6(const uint8_t* rgb8, int width, int height, unsigned char* out, long cap) {
    if (!rgb8 || !out || cap < 32) return -1;
    long n = snprintf((char*)out, (size_t)cap, "P6\n%d %d\n255\n", width, height);
    const long body = (long)width * height * 3;
    if (n + body > cap) return -1;
    memcpy(out + n, rgb8, (size_t)body);
    return n + body;
}

// PNG (RFC 2083) with the scanlines in stored (uncompressed) deflate blocks: no compressor in the image, and the point
// is an exact, viewer-readable container for the same bytes the P3 text carries.
long grt_host_write_png(const uint8_t* rgb8, int width, int height, unsigned char* out, long cap) {
    if (!rgb8 || !out || width <= 0 || height <= 0) return -1;
    static uint32_t crc_table[256];
    static bool have_table = false;
    if (!have_table) {
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; crc_table[i] = c; }
        have_table = true;
    }
    const size_t row = (size_t)width * 3 + 1, raw = row * (size_t)height;
    const size_t n_blocks = (raw + 65534) / 65535;
    const size_t zlen = 2 + raw + 5 * n_blocks + 4;
    const size_t total = 8 + (12 + 13) + (12 + zlen) + 12;
    if ((size_t)cap < total) return -1;
    unsigned char* p = out;
    auto be32 = [](unsigned char* q, uint32_t v) { q[0] = (unsigned char)(v >> 24); q[1] = (unsigned char)(v >> 16); q[2] = (unsigned char)(v >> 8); q[3] = (unsigned char)v; };
    auto chunk_end = [&](unsigned char* type_at, size_t len) {   // CRC over type + data
        uint32_t c = 0xFFFFFFFFu;
        for (size_t i = 0; i < len + 4; i++) c = crc_table[(c ^ type_at[i]) & 0xFFu] ^ (c >> 8);
        be32(type_at + 4 + len, c ^ 0xFFFFFFFFu);
    };
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    memcpy(p, sig, 8); p += 8;
    be32(p, 13); memcpy(p + 4, "IHDR", 4); be32(p + 8, (uint32_t)width); be32(p + 12, (uint32_t)height);
    p[16] = 8; p[17] = 2; p[18] = 0; p[19] = 0; p[20] = 0;   // 8 bits, RGB, deflate, no filter method, no interlace
    chunk_end(p + 4, 13); p += 12 + 13;
    be32(p, (uint32_t)zlen); memcpy(p + 4, "IDAT", 4);
    unsigned char* z = p + 8;
    *z++ = 0x78; *z++ = 0x01;
    uint32_t a1 = 1, a2 = 0;   // Adler-32 of the raw scanline stream
    size_t done = 0, y = 0, x = 0;   // position in the raw stream = filter byte + row bytes
    for (size_t b = 0; b < n_blocks; b++) {
        const size_t len = raw - done < 65535 ? raw - done : 65535;
        *z++ = (b + 1 == n_blocks) ? 1 : 0;
        *z++ = (unsigned char)(len & 0xFF); *z++ = (unsigned char)(len >> 8);
        *z++ = (unsigned char)(~len & 0xFF); *z++ = (unsigned char)((~len >> 8) & 0xFF);
        for (size_t i = 0; i < len; i++) {
            unsigned char v = (x == 0) ? 0 : rgb8[y * (size_t)width * 3 + (x - 1)];   // filter type 0 (None) starts every row
            *z++ = v;
            a1 += v; if (a1 >= 65521u) a1 -= 65521u;
            a2 += a1; if (a2 >= 65521u) a2 -= 65521u;
            if (++x == row) { x = 0; y++; }
        }
        done += len;
    }
    be32(z, (a2 << 16) | a1);
    chunk_end(p + 4, zlen); p += 12 + zlen;
    be32(p, 0); memcpy(p + 4, "IEND", 4); chunk_end(p + 4, 0); p += 12;
    return (long)(p - out);
}

int grt_host_camera_render(GrtHostScene* s, const GrtCameraConfig* cfg, uint64_t seed, int variant, int n_gpus,
                           float* rgb_sum_out, char* ppm_out, long ppm_cap, long* ppm_len, double* kernel_ms) {
    return grt_host_camera_render_rgb8(s, cfg, seed, variant, n_gpus, rgb_sum_out, nullptr, ppm_out, ppm_cap, ppm_len, kernel_ms);
}

int grt_host_camera_render_rgb8(GrtHostScene* s, const GrtCameraConfig* cfg, uint64_t seed, int variant, int n_gpus,
                                float* rgb_sum_out, uint8_t* rgb8_out, char* ppm_out, long ppm_cap, long* ppm_len, double* kernel_ms) {
    if (!s || !cfg) { fail("NULL argument"); return GRT_E_INVALID; }
    GrtScene scene;
    if (grt_host_flatten(s, &scene)) return GRT_E_INVALID;
    GrtCamera cam;
    if (grt_host_camera_derive(cfg, &cam)) return GRT_E_INVALID;
    GrtOptions opt;
    memset(&opt, 0, sizeof(opt));
    opt.seed = seed; opt.variant = variant; opt.sample_stride = 1;
    size_t nval = (size_t)cam.width * cam.height * 3;
    std::vector<float> sum(nval, 0.0f);
    std::vector<uint8_t> rgb8(nval, 0);
    int rc;
    if (n_gpus <= 1) {
        GrtSceneHandle h = nullptr;
        rc = grt_scene_upload(&scene, 0, &h);
        if (rc) { fail(grt_last_error()); return rc; }
        rc = grt_render(h, &cam, &opt, sum.data(), rgb8.data(), nullptr);
        grt_scene_free(h);
        if (kernel_ms) *kernel_ms = 0;
    } else {
        std::vector<int> devs(n_gpus);
        for (int i = 0; i < n_gpus; i++) devs[i] = i;
        rc = grt_render_multi(&scene, &cam, &opt, devs.data(), n_gpus, sum.data(), rgb8.data(), kernel_ms);
    }
    if (rc) { fail(grt_last_error()); return rc; }
    if (rgb_sum_out) memcpy(rgb_sum_out, sum.data(), nval * sizeof(float));
    if (rgb8_out) memcpy(rgb8_out, rgb8.data(), nval);
    if (ppm_out) {
        long n = grt_host_write_ppm(rgb8.data(), cam.width, cam.height, ppm_out, ppm_cap);
        if (n < 0) { fail("ppm buffer too small"); return GRT_E_INVALID; }
        if (ppm_len) *ppm_len = n;
    }
    return GRT_OK;
}

const void* grt_host_scene_description(GrtHostScene* s) { return s ? &s->ir : nullptr; }

}  // extern "C"
