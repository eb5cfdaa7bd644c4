// scene_ir.hpp — host-side scene description in the reference's own vocabulary.
//
// The Go program builds its scene by calling constructors of package hittable
// (NewSphere, NewQuad, NewBox, BuildBVH, Translate, RotateY, ConstantMedium,
// NewLambertian, ...; reference: internal/hittable/*.go, main.go:19-409).  This
// header records exactly those calls, with their fp64 arguments, as plain data.
// It contains NO intersection, traversal, shading or BVH-build algorithm: the
// flattener (flatten.cpp, product) and the oracle (oracle/, test
// infrastructure) each consume this description independently.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>
#include <stdexcept>

namespace grt {
namespace ir {

struct V3 {
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, double t) { return V3(a.x * t, a.y * t, a.z * t); }
inline V3 neg(V3 a) { return V3(-a.x, -a.y, -a.z); }

// Deterministic host RNG standing in for Go's unseeded global math/rand at
// SCENE-BUILD time (main.go:40-41,64,107,155; perlin.go:25,87).  splitmix64.
struct HostRng {
    uint64_t s;
    explicit HostRng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double Float64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    int Intn(int n) { return (int)(next() % (uint64_t)n); }
    double RangeRange(double lo, double hi) { return lo + (hi - lo) * Float64(); }   // util/utilities.go:12
};

enum TexType { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_IMAGE = 2, TEX_NOISE = 3 };
enum NoiseVariant { NOISE_PERLIN = 1, NOISE_MARBLE = 2, NOISE_TURBULENT = 3 };  // texture.go:93-96
enum MatType { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
enum HType { H_SPHERE = 0, H_QUAD, H_TRI, H_LIST, H_BVH, H_TRANSLATE, H_ROTATEY, H_MEDIUM };

struct Texture {
    int type = TEX_SOLID;
    V3 color;          // solid
    double scale = 1;  // checker: scale (consumer takes 1/scale, texture.go:37); noise: scale
    int even = -1, odd = -1;
    int image = -1;
    int perlin = -1;
    int variant = NOISE_PERLIN;
};
struct Material {
    int type = MAT_LAMBERTIAN;
    int tex = -1;
    V3 albedo;
    double fuzz = 0;
    double ior = 1;
};
struct Image {
    int width = 0, height = 0;
    std::vector<uint8_t> rgb;  // RTImage.bdata as RGB8, row-major (imageLoader.go:78-84)
};
struct Perlin {  // perlin.go:12-17
    double vec[256][3];
    int perm[3][256];
};
struct SphereP { V3 c0, dc; double r; };                 // Center ray (origin, direction), objects.go:14-36
struct QuadP { V3 Q, u, v; };                            // objects.go:117-141
struct TriP {                                            // objects.go:242-313
    V3 v[3];
    V3 n[3];
    double uv[3][2];
    bool hasNormals = false, hasUV = false;
};
struct XformP { V3 offset; double degrees = 0; };        // transformation.go:13-19,36-42
struct MediumP { double density; int phase; };           // medium.go:13-25
struct BoxP { V3 mn, mx; int quads[6]; };                // NewBox, objects.go:208-240 (remembered so the six quads can share one slab test)

struct Hittable {
    uint8_t type;
    int32_t mat;    // primitives
    int32_t a;      // payload index (sphere/quad/tri/xform/medium/list) or list id for BVH
    int32_t child;  // wrapped object for translate/rotateY/medium; list id for BVH
};

struct Scene {
    std::vector<Texture> textures;
    std::vector<Material> materials;
    std::vector<Image> images;
    std::vector<Perlin> perlins;
    std::vector<Hittable> hittables;
    std::vector<SphereP> spheres;
    std::vector<QuadP> quads;
    std::vector<TriP> tris;
    std::vector<XformP> xforms;
    std::vector<MediumP> media;
    std::vector<std::vector<int>> lists;
    std::vector<BoxP> boxes;
    std::vector<int> box_of_list;   // list payload index -> boxes index, or -1
    int world = -1;
    int lights = -1;

    // ---- textures (texture.go) ------------------------------------------
    int NewSolidColor(V3 albedo) { Texture t; t.type = TEX_SOLID; t.color = albedo; textures.push_back(t); return (int)textures.size() - 1; }
    int NewCheckerboard(double scale, int even, int odd) {
        Texture t; t.type = TEX_CHECKER; t.scale = scale; t.even = even; t.odd = odd; textures.push_back(t); return (int)textures.size() - 1;
    }
    int NewCheckerboardColors(double scale, V3 even, V3 odd) {
        int e = NewSolidColor(even), o = NewSolidColor(odd);
        return NewCheckerboard(scale, e, o);
    }
    int AddImage(int w, int h, const uint8_t* rgb) {
        Image im; im.width = w; im.height = h; im.rgb.assign(rgb, rgb + (size_t)w * h * 3); images.push_back(std::move(im));
        return (int)images.size() - 1;
    }
    int NewImageTexture(int image) { Texture t; t.type = TEX_IMAGE; t.image = image; textures.push_back(t); return (int)textures.size() - 1; }
    // NewPerlin (perlin.go:20-31) with the draw order of the reference:
    // 256 x RangeRandom(-1,1).UnitVector(), then permute(permX), permY, permZ.
    int NewPerlin(uint64_t seed) {
        HostRng rng(seed);
        Perlin p;
        for (int i = 0; i < 256; i++) {
            double x = rng.RangeRange(-1, 1), y = rng.RangeRange(-1, 1), z = rng.RangeRange(-1, 1);
            double s = 1.0 / std::sqrt(x * x + y * y + z * z);
            p.vec[i][0] = x * s; p.vec[i][1] = y * s; p.vec[i][2] = z * s;
        }
        for (int a = 0; a < 3; a++) {
            for (int i = 0; i < 256; i++) p.perm[a][i] = i;
        }
        for (int a = 0; a < 3; a++) {
            for (int i = 255; i > 0; i--) {  // perlin.go:85-90 (rand.Intn(i), not i+1)
                int target = rng.Intn(i);
                int tmp = p.perm[a][i]; p.perm[a][i] = p.perm[a][target]; p.perm[a][target] = tmp;
            }
        }
        perlins.push_back(p);
        return (int)perlins.size() - 1;
    }
    int NewNoiseTextureWithType(double scale, int variant, uint64_t seed) {
        Texture t; t.type = TEX_NOISE; t.scale = scale; t.variant = variant; t.perlin = NewPerlin(seed);
        textures.push_back(t); return (int)textures.size() - 1;
    }

    // ---- materials (materials.go) ---------------------------------------
    int addMat(const Material& m) { materials.push_back(m); return (int)materials.size() - 1; }
    int NewTexturedLambertian(int tex) { Material m; m.type = MAT_LAMBERTIAN; m.tex = tex; return addMat(m); }
    int NewLambertian(V3 albedo) { return NewTexturedLambertian(NewSolidColor(albedo)); }
    int NewMetal(V3 albedo, double fuzz) { Material m; m.type = MAT_METAL; m.albedo = albedo; m.fuzz = fuzz; return addMat(m); }
    int NewDielectric(double ior) { Material m; m.type = MAT_DIELECTRIC; m.ior = ior; return addMat(m); }
    int NewDiffuseLightTextured(int tex) { Material m; m.type = MAT_DIFFUSE_LIGHT; m.tex = tex; return addMat(m); }
    int NewDiffuseLight(V3 color) { return NewDiffuseLightTextured(NewSolidColor(color)); }
    int NewIsotropicTexture(int tex) { Material m; m.type = MAT_ISOTROPIC; m.tex = tex; return addMat(m); }
    int NewIsotropic(V3 albedo) { return NewIsotropicTexture(NewSolidColor(albedo)); }

    // ---- hittables (objects.go, hittable.go, bvh.go, transformation.go, medium.go)
    int addH(uint8_t type, int mat, int a, int child) {
        Hittable h; h.type = type; h.mat = mat; h.a = a; h.child = child; hittables.push_back(h);
        return (int)hittables.size() - 1;
    }
    void checkMat(int m) const { if (m < 0 || m >= (int)materials.size()) throw std::invalid_argument("bad material id"); }
    void checkH(int h) const { if (h < 0 || h >= (int)hittables.size()) throw std::invalid_argument("bad hittable id"); }
    int NewSphere(V3 center, double radius, int mat) {
        checkMat(mat);
        SphereP s; s.c0 = center; s.dc = V3(0, 0, 0); s.r = radius; spheres.push_back(s);
        return addH(H_SPHERE, mat, (int)spheres.size() - 1, -1);
    }
    int NewMotionSphere(V3 c1, V3 c2, double radius, int mat) {
        checkMat(mat);
        SphereP s; s.c0 = c1; s.dc = c2 - c1; s.r = radius; spheres.push_back(s);
        return addH(H_SPHERE, mat, (int)spheres.size() - 1, -1);
    }
    int NewQuad(V3 Q, V3 u, V3 v, int mat) {
        checkMat(mat);
        QuadP q; q.Q = Q; q.u = u; q.v = v; quads.push_back(q);
        return addH(H_QUAD, mat, (int)quads.size() - 1, -1);
    }
    int NewTriangleFull(const V3 v[3], const V3* n, const double (*uv)[2], int mat) {
        checkMat(mat);
        TriP t;
        for (int i = 0; i < 3; i++) t.v[i] = v[i];
        t.hasNormals = n != nullptr; t.hasUV = uv != nullptr;
        for (int i = 0; i < 3; i++) {
            t.n[i] = n ? n[i] : V3();
            t.uv[i][0] = uv ? uv[i][0] : 0; t.uv[i][1] = uv ? uv[i][1] : 0;
        }
        tris.push_back(t);
        return addH(H_TRI, mat, (int)tris.size() - 1, -1);
    }
    int NewTriangle(const V3 v[3], int mat) { return NewTriangleFull(v, nullptr, nullptr, mat); }
    int NewTriangleWithNormals(const V3 v[3], const V3 n[3], int mat) { return NewTriangleFull(v, n, nullptr, mat); }
    int NewTexturedTriangle(const V3 v[3], const double uv[3][2], int mat) { return NewTriangleFull(v, nullptr, uv, mat); }
    int NewTexturedTriangleWithNormals(const V3 v[3], const V3 n[3], const double uv[3][2], int mat) { return NewTriangleFull(v, n, uv, mat); }

    int NewHittableList() { lists.emplace_back(); box_of_list.push_back(-1); return addH(H_LIST, -1, (int)lists.size() - 1, -1); }
    void Add(int list, int obj) {
        checkH(list); checkH(obj);
        if (hittables[list].type != H_LIST) throw std::invalid_argument("Add: not a list");
        if (box_of_list[hittables[list].a] >= 0) box_of_list[hittables[list].a] = -1;   // a box's side list was modified: no longer a plain box
        lists[hittables[list].a].push_back(obj);
    }
    // BuildBVH(list): recorded lazily; consumers run bvhHelper (bvh.go:35-61)
    // over the list's CURRENT contents at consume time.  The reference sorts
    // list.objects in place; consumers sort a copy (only matters if the same
    // list is also used as `lights`, where the pick is uniform anyway).
    int BuildBVH(int list) {
        checkH(list);
        if (hittables[list].type != H_LIST) throw std::invalid_argument("BuildBVH: not a list");
        return addH(H_BVH, -1, hittables[list].a, list);
    }
    // NewBox (objects.go:208-240): six quads in the reference's order, BVH over them.
    int NewBox(V3 a, V3 b, int mat) {
        int sides = NewHittableList();
        V3 mn(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z));
        V3 mx(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z));
        V3 dx(mx.x - mn.x, 0, 0), dy(0, mx.y - mn.y, 0), dz(0, 0, mx.z - mn.z);
        Add(sides, NewQuad(V3(mn.x, mn.y, mx.z), dx, dy, mat));        // front
        Add(sides, NewQuad(V3(mx.x, mn.y, mx.z), neg(dz), dy, mat));   // right
        Add(sides, NewQuad(V3(mx.x, mn.y, mn.z), neg(dx), dy, mat));   // back
        Add(sides, NewQuad(V3(mn.x, mn.y, mn.z), dz, dy, mat));        // left
        Add(sides, NewQuad(V3(mn.x, mx.y, mx.z), dx, neg(dz), mat));   // top
        Add(sides, NewQuad(V3(mn.x, mn.y, mn.z), dx, dz, mat));        // bottom
        BoxP bp; bp.mn = mn; bp.mx = mx;
        for (int i = 0; i < 6; i++) bp.quads[i] = lists[hittables[sides].a][i];
        boxes.push_back(bp);
        box_of_list[hittables[sides].a] = (int)boxes.size() - 1;
        return BuildBVH(sides);
    }
    int Translate(int obj, V3 offset) {
        checkH(obj);
        XformP x; x.offset = offset; xforms.push_back(x);
        return addH(H_TRANSLATE, -1, (int)xforms.size() - 1, obj);
    }
    int RotateY(int obj, double degrees) {
        checkH(obj);
        XformP x; x.degrees = degrees; xforms.push_back(x);
        return addH(H_ROTATEY, -1, (int)xforms.size() - 1, obj);
    }
    int ConstantMediumTexture(int boundary, double density, int tex) {
        checkH(boundary);
        MediumP m; m.density = density; m.phase = NewIsotropicTexture(tex); media.push_back(m);
        return addH(H_MEDIUM, m.phase, (int)media.size() - 1, boundary);
    }
    int ConstantMedium(int boundary, double density, V3 albedo) {
        return ConstantMediumTexture(boundary, density, NewSolidColor(albedo));
    }
};

// Public fields of camera.Camera (camera.go:24-36) + PositionCamera (:65-81).
struct CameraConfig {
    double AspectRatio = 0;
    int Width = 0;
    int SamplesPerPixel = 0;
    int MaxDepth = 0;
    int MaxThreads = 1;
    double VerticalFOV = 0;
    double DefocusAngle = 0;
    double FocusDistance = 0;
    V3 Background;
    double MaxContribution = 0;
    V3 lookFrom = V3(0, 0, 0), lookAt = V3(0, 0, -1), vup = V3(0, 1, 0);
    void PositionCamera(V3 from, V3 at, V3 up) { lookFrom = from; lookAt = at; vup = up; }
};

}  // namespace ir
}  // namespace grt
