// scenes.hpp — the reference's eight scene functions (main.go:19-409) restated
// as scene-description builders, plus the synthetic mesh of config C5.
//
// Only PARAMETERS live here (positions, colours, camera fields).  Go's scenes
// draw their random geometry from the unseeded global math/rand, so there is
// no canonical geometry to match; we use fixed seeds and the
// reference's draw ORDER.
#pragma once
#include <charconv>
#include "scene_ir.hpp"
#include "obj_loader.hpp"
#include <cstdio>

namespace grt {
namespace scenes {

using ir::V3;
using ir::Scene;
using ir::CameraConfig;

struct SceneOptions {
    int width = 0;            // 0 = shipped value
    int spp = 0;              // 0 = shipped value
    double aspect = 0;        // 0 = shipped value
    uint64_t seed = 0;        // 0 = per-scene default
    int mesh_segments = 0;    // C5: UV-sphere segments (0 = 708 -> ~1.0 M triangles)
    const uint8_t* image_rgb = nullptr;  // earthmap texels for scenes 2 and 5
    int image_w = 0, image_h = 0;
};

static const uint64_t SEED_BOOK1 = 0x5EED0001ull;
static const uint64_t SEED_BOOK2 = 0x5EED0004ull;

inline void applyOverrides(CameraConfig& c, const SceneOptions& o) {
    if (o.width > 0) c.Width = o.width;
    if (o.spp > 0) c.SamplesPerPixel = o.spp;
    if (o.aspect > 0) c.AspectRatio = o.aspect;
}

// Procedural stand-in used when no decoded earthmap is supplied (the JPEG
// decode itself is out of scope).
inline int proceduralImage(Scene& s, int w, int h) {
    std::vector<uint8_t> px((size_t)w * h * 3);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            double u = (double)x / w, v = (double)y / h;
            double land = std::sin(u * 37.0) * std::sin(v * 23.0) + 0.5 * std::sin(u * 91.0 + v * 57.0);
            uint8_t* p = &px[((size_t)y * w + x) * 3];
            if (land > 0.3) { p[0] = 60; p[1] = (uint8_t)(120 + 60 * v); p[2] = 40; }
            else { p[0] = 20; p[1] = (uint8_t)(60 + 40 * u); p[2] = (uint8_t)(150 + 80 * v); }
        }
    return s.AddImage(w, h, px.data());
}
inline int earthImage(Scene& s, const SceneOptions& o) {
    if (o.image_rgb && o.image_w > 0 && o.image_h > 0) return s.AddImage(o.image_w, o.image_h, o.image_rgb);
    return proceduralImage(s, 1024, 512);
}

// main.go:19-91
inline void book1Scene(Scene& s, CameraConfig& c, const SceneOptions& o) {
    c.AspectRatio = 16.0 / 9.0;
    c.Width = 400;
    c.SamplesPerPixel = 100;
    c.MaxDepth = 50;
    c.VerticalFOV = 20;
    c.PositionCamera(V3(13, 2, 3), V3(0, 0, 0), V3(0, 1, 0));
    c.DefocusAngle = 0.6;
    c.FocusDistance = 10.0;
    c.Background = V3(0.70, 0.80, 1.00);

    ir::HostRng rng(o.seed ? o.seed : SEED_BOOK1);
    int world = s.NewHittableList();
    int lights = s.NewHittableList();
    int glass = s.NewDielectric(1.5);
    int checker = s.NewCheckerboardColors(0.32, V3(.2, .3, .1), V3(.9, .9, .9));
    s.Add(world, s.NewSphere(V3(0, -1000, 0), 1000, s.NewTexturedLambertian(checker)));
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            double mat = rng.Float64();
            double cx = (double)a + 0.9 * rng.Float64();
            double cz = (double)b + 0.9 * rng.Float64();
            V3 center(cx, 0.2, cz);
            V3 d = center + ir::neg(V3(4, 0.2, 0));
            if (std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z) > 0.9) {
                if (mat < 0.6) {
                    V3 r1(rng.Float64(), rng.Float64(), rng.Float64());
                    V3 r2(rng.Float64(), rng.Float64(), rng.Float64());
                    V3 albedo(r1.x * r2.x, r1.y * r2.y, r1.z * r2.z);
                    int m = s.NewLambertian(albedo);
                    s.Add(world, s.NewMotionSphere(center, center + V3(0, rng.RangeRange(0, 0.5), 0), 0.2, m));
                } else if (mat < 0.8) {
                    // perlin orbs: the reference builds a material and never adds a sphere
                    // (main.go:52-60 has no world.Add); we only consume the Intn draw.
                    (void)rng.Intn(10);
                } else if (mat < 0.95) {
                    V3 albedo(rng.RangeRange(0.5, 1.0), rng.RangeRange(0.5, 1.0), rng.RangeRange(0.5, 1.0));
                    double fuzz = rng.Float64();
                    s.Add(world, s.NewSphere(center, 0.2, s.NewMetal(albedo, fuzz)));
                } else {
                    s.Add(world, s.NewSphere(center, 0.2, glass));
                }
            }
        }
    }
    s.Add(world, s.NewSphere(V3(0, 1, 0), 1.0, glass));
    s.Add(world, s.NewSphere(V3(-4, 1, 0), 1.0, s.NewLambertian(V3(0.4, 0.2, 0.1))));
    s.Add(world, s.NewSphere(V3(4, 1, 0), 1.0, s.NewMetal(V3(.7, .6, .5), 0)));
    int sun = s.NewSphere(V3(0, 100, 0), 50, s.NewDiffuseLight(V3(5, 5, 5)));
    s.Add(world, sun);
    s.Add(lights, sun);
    s.world = s.BuildBVH(world);
    s.lights = lights;
    applyOverrides(c, o);
}

// main.go:94-174
inline void book2Scene(Scene& s, CameraConfig& c, const SceneOptions& o) {
    uint64_t seed = o.seed ? o.seed : SEED_BOOK2;
    ir::HostRng rng(seed);
    int boxes1 = s.NewHittableList();
    int ground = s.NewLambertian(V3(.48, .83, .53));
    const int boxesPerSide = 20;
    for (int i = 0; i < boxesPerSide; i++) {
        for (int j = 0; j < boxesPerSide; j++) {
            double w = 100.0;
            double x0 = -1000.0 + (double)i * w;
            double z0 = -1000.0 + (double)j * w;
            double y0 = 0.0;
            double x1 = x0 + w;
            double y1 = rng.RangeRange(1, 101);
            double z1 = z0 + w;
            s.Add(boxes1, s.NewBox(V3(x0, y0, z0), V3(x1, y1, z1), ground));
        }
    }
    int world = s.NewHittableList();
    s.Add(world, s.BuildBVH(boxes1));
    int lights = s.NewHittableList();
    int light = s.NewQuad(V3(123, 554, 147), V3(300, 0, 0), V3(0, 0, 265), s.NewDiffuseLight(V3(7, 7, 7)));
    s.Add(world, light);
    s.Add(lights, light);
    V3 c1(400, 400, 200);
    V3 c2 = c1 + V3(30, 0, 0);
    s.Add(world, s.NewMotionSphere(c1, c2, 50, s.NewLambertian(V3(.7, .3, .1))));
    s.Add(world, s.NewSphere(V3(260, 150, 45), 50, s.NewDielectric(1.5)));
    s.Add(world, s.NewSphere(V3(0, 150, 145), 50, s.NewMetal(V3(0.8, 0.8, 0.9), 1.0)));
    int boundary = s.NewSphere(V3(360, 150, 145), 70, s.NewDielectric(1.5));
    s.Add(world, boundary);
    s.Add(world, s.ConstantMedium(boundary, .2, V3(0.2, 0.4, 0.9)));
    int b2 = s.NewSphere(V3(0, 0, 0), 5000, s.NewDielectric(1.5));
    s.Add(world, s.ConstantMedium(b2, .0001, V3(1, 1, 1)));
    int eMat = s.NewTexturedLambertian(s.NewImageTexture(earthImage(s, o)));
    s.Add(world, s.NewSphere(V3(400, 200, 400), 100, eMat));
    int p = s.NewTexturedLambertian(s.NewNoiseTextureWithType(.2, ir::NOISE_MARBLE, seed + 1));
    s.Add(world, s.NewSphere(V3(220, 280, 300), 80, p));
    int boxes2 = s.NewHittableList();
    int white = s.NewLambertian(V3(.73, .73, .73));
    for (int k = 0; k < 1000; k++) {
        V3 ctr(rng.RangeRange(0, 165), rng.RangeRange(0, 165), rng.RangeRange(0, 165));
        s.Add(boxes2, s.NewSphere(ctr, 10, white));
    }
    s.Add(world, s.Translate(s.RotateY(s.BuildBVH(boxes2), 15), V3(-100, 270, 395)));
    c.AspectRatio = 1.0;
    c.Width = 800;
    c.SamplesPerPixel = 100;
    c.MaxDepth = 40;
    c.Background = V3(0, 0, 0);
    c.VerticalFOV = 40;
    c.PositionCamera(V3(478, 278, -600), V3(278, 278, 0), V3(0, 1, 0));
    c.DefocusAngle = 0;
    s.world = world;
    s.lights = lights;
    applyOverrides(c, o);
}

inline void cornellWalls(Scene& s, int world, int red, int white, int green) {
    s.Add(world, s.NewQuad(V3(555, 0, 0), V3(0, 555, 0), V3(0, 0, 555), green));
    s.Add(world, s.NewQuad(V3(0, 0, 0), V3(0, 555, 0), V3(0, 0, 555), red));
    s.Add(world, s.NewQuad(V3(0, 0, 0), V3(555, 0, 0), V3(0, 0, 555), white));
    s.Add(world, s.NewQuad(V3(555, 555, 555), V3(-555, 0, 0), V3(0, 0, -555), white));
    s.Add(world, s.NewQuad(V3(0, 0, 555), V3(555, 0, 0), V3(0, 555, 0), white));
}
inline void cornellCamera(CameraConfig& c, int width, int spp) {
    c.AspectRatio = 1.0;
    c.Width = width;
    c.SamplesPerPixel = spp;
    c.MaxDepth = 50;
    c.Background = V3(0, 0, 0);
    c.VerticalFOV = 40;
    c.PositionCamera(V3(278, 278, -800), V3(278, 278, 0), V3(0, 1, 0));
    c.DefocusAngle = 0;
}

// main.go:177-218
inline void book3Scene(Scene& s, CameraConfig& c, const SceneOptions& o) {
    int world = s.NewHittableList();
    int red = s.NewLambertian(V3(.65, .05, .05));
    int white = s.NewLambertian(V3(.73, .73, .73));
    int green = s.NewLambertian(V3(.12, .45, .15));
    int light = s.NewDiffuseLight(V3(15, 15, 15));
    cornellWalls(s, world, red, white, green);
    int lights = s.NewHittableList();
    s.Add(lights, s.NewQuad(V3(343, 550, 332), V3(-130, 0, 0), V3(0, 0, -105), light));
    s.Add(world, lights);
    int b1 = s.NewBox(V3(0, 0, 0), V3(165, 330, 165), white);
    b1 = s.RotateY(b1, 15);
    b1 = s.Translate(b1, V3(265, 0, 295));
    s.Add(world, b1);
    int sp = s.NewSphere(V3(190, 90, 190), 90, s.NewDielectric(1.5));
    s.Add(lights, sp);
    s.Add(world, sp);
    cornellCamera(c, 600, 10);
    s.world = s.BuildBVH(world);
    s.lights = lights;
    applyOverrides(c, o);
}

// main.go:220-247
inline void quadsScene(Scene& s, CameraConfig& c, const SceneOptions& o) {
    int world = s.NewHittableList();
    int lights = s.NewHittableList();
    int leftEarth = s.NewTexturedLambertian(s.NewImageTexture(earthImage(s, o)));
    int backLight = s.NewDiffuseLight(V3(3, 3, 3));
    int rightPerlin = s.NewTexturedLambertian(s.NewNoiseTextureWithType(5, ir::NOISE_MARBLE, (o.seed ? o.seed : 0x5EED0005ull)));
    int upperMetal = s.NewMetal(V3(0.8, 0.6, 0.2), 0);
    int lowerTeal = s.NewLambertian(V3(0.2, 0.8, 0.8));
    s.Add(world, s.NewQuad(V3(-3, -2, 5), V3(0, 0, -4), V3(0, 4, 0), leftEarth));
    int light = s.NewQuad(V3(-2, -2, 0), V3(4, 0, 0), V3(0, 4, 0), backLight);
    s.Add(world, light);
    s.Add(world, s.NewQuad(V3(3, -2, 1), V3(0, 0, 4), V3(0, 4, 0), rightPerlin));
    s.Add(world, s.NewQuad(V3(-2, 3, 1), V3(4, 0, 0), V3(0, 0, 4), upperMetal));
    s.Add(world, s.NewQuad(V3(-2, -3, 5), V3(4, 0, 0), V3(0, 0, -4), lowerTeal));
    s.Add(lights, light);
    c.AspectRatio = 1.0;
    c.Width = 400;
    c.SamplesPerPixel = 100;
    c.MaxDepth = 50;
    c.Background = V3(0.70, 0.80, 1.00);
    c.VerticalFOV = 80;
    c.PositionCamera(V3(0, 0, 9), V3(0, 0, 0), V3(0, 1, 0));
    c.DefocusAngle = 0;
    s.world = s.BuildBVH(world);
    s.lights = lights;
    applyOverrides(c, o);
}

// main.go:249-275 — world is a plain HittableList, lights is a BARE quad.
inline void simpleLight(Scene& s, CameraConfig& c, const SceneOptions& o) {
    int world = s.NewHittableList();
    int p = s.NewNoiseTextureWithType(4, ir::NOISE_MARBLE, (o.seed ? o.seed : 0x5EED0006ull));
    int l = s.NewDiffuseLight(V3(4, 4, 4));
    int s1 = s.NewSphere(V3(0, -1000, 0), 1000, s.NewTexturedLambertian(p));
    int s2 = s.NewSphere(V3(0, 2, 0), 2, s.NewTexturedLambertian(p));
    int q = s.NewQuad(V3(3, 1, -2), V3(2, 0, 0), V3(0, 2, 0), l);
    int sl = s.NewSphere(V3(0, 7, 0), 2, l);
    s.Add(world, s1);
    s.Add(world, sl);
    s.Add(world, q);
    s.Add(world, s2);
    c.AspectRatio = 16.0 / 9.0;
    c.Width = 400;
    c.SamplesPerPixel = 100;
    c.MaxDepth = 50;
    c.Background = V3(0, 0, 0);
    c.VerticalFOV = 20;
    c.PositionCamera(V3(26, 3, 6), V3(0, 2, 0), V3(0, 1, 0));
    c.DefocusAngle = 0;
    s.world = world;
    s.lights = q;
    applyOverrides(c, o);
}

// main.go:278-320
inline void cornellBox(Scene& s, CameraConfig& c, const SceneOptions& o) {
    int world = s.NewHittableList();
    int red = s.NewLambertian(V3(.65, .05, .05));
    int white = s.NewLambertian(V3(.73, .73, .73));
    int green = s.NewLambertian(V3(.12, .45, .15));
    int light = s.NewDiffuseLight(V3(15, 15, 15));
    cornellWalls(s, world, red, white, green);
    int lights = s.NewHittableList();
    s.Add(lights, s.NewQuad(V3(343, 550, 332), V3(-130, 0, 0), V3(0, 0, -105), light));
    s.Add(world, lights);
    int b1 = s.NewBox(V3(0, 0, 0), V3(165, 330, 165), white);
    b1 = s.RotateY(b1, 15);
    b1 = s.Translate(b1, V3(265, 0, 295));
    int b2 = s.NewBox(V3(0, 0, 0), V3(165, 165, 165), white);
    b2 = s.RotateY(b2, -18);
    b2 = s.Translate(b2, V3(130, 0, 65));
    s.Add(world, b1);
    s.Add(world, b2);
    cornellCamera(c, 600, 100);
    s.world = s.BuildBVH(world);
    s.lights = lights;
    applyOverrides(c, o);
}

// main.go:323-367
inline void cornellSmoke(Scene& s, CameraConfig& c, const SceneOptions& o) {
    int world = s.NewHittableList();
    int lights = s.NewHittableList();
    int red = s.NewLambertian(V3(.65, .05, .05));
    int white = s.NewLambertian(V3(.73, .73, .73));
    int green = s.NewLambertian(V3(.12, .45, .15));
    int light = s.NewDiffuseLight(V3(15, 15, 15));
    s.Add(world, s.NewQuad(V3(555, 0, 0), V3(0, 555, 0), V3(0, 0, 555), green));
    s.Add(world, s.NewQuad(V3(0, 0, 0), V3(0, 555, 0), V3(0, 0, 555), red));
    int lightQuad = s.NewQuad(V3(343, 550, 332), V3(-130, 0, 0), V3(0, 0, -105), light);
    s.Add(world, lightQuad);
    s.Add(lights, lightQuad);
    s.Add(world, s.NewQuad(V3(0, 0, 0), V3(555, 0, 0), V3(0, 0, 555), white));
    s.Add(world, s.NewQuad(V3(555, 555, 555), V3(-555, 0, 0), V3(0, 0, -555), white));
    s.Add(world, s.NewQuad(V3(0, 0, 555), V3(555, 0, 0), V3(0, 555, 0), white));
    int b1 = s.NewBox(V3(0, 0, 0), V3(165, 330, 165), white);
    b1 = s.RotateY(b1, 15);
    b1 = s.Translate(b1, V3(265, 0, 295));
    int b2 = s.NewBox(V3(0, 0, 0), V3(165, 165, 165), white);
    b2 = s.RotateY(b2, -18);
    b2 = s.Translate(b2, V3(130, 0, 65));
    s.Add(world, s.ConstantMedium(b1, .01, V3(0, 0, 0)));
    s.Add(world, s.ConstantMedium(b2, .01, V3(1, 1, 1)));
    cornellCamera(c, 600, 10);
    s.world = s.BuildBVH(world);
    s.lights = lights;
    applyOverrides(c, o);
}

// Synthetic mesh for config C5 as OBJ TEXT: a UV sphere of nseg x nseg quad faces with radius
// 1 + 0.08 sin(7θ) sin(5φ) + 0.02 sin(31θ+17φ) and smooth vertex normals (`v`, `vn`, `f a//a b//b d//d c//c`),
// deterministic, no RNG.  It is then loaded through the objLoader mirror (obj_loader.hpp) exactly like
// modelExample loads dragon.obj: ScaleFactor 5, Center, Position (main.go:377-383), fan triangulation
// (objLoader.go:396-467), NewTriangleWithNormals.  %.17g keeps every double exact through the text.
inline std::string displacedSphereObj(int nseg) {
    const double PI = 3.14159265358979323846;
    auto radius = [&](double th, double ph) { return 1.0 + 0.08 * std::sin(7 * th) * std::sin(5 * ph) + 0.02 * std::sin(31 * th + 17 * ph); };
    auto pos = [&](double th, double ph) {
        double r = radius(th, ph);
        return V3(r * std::sin(th) * std::cos(ph), r * std::cos(th), r * std::sin(th) * std::sin(ph));
    };
    std::string out;
    out.reserve((size_t)(nseg + 1) * (nseg + 1) * 140 + (size_t)nseg * nseg * 60);
    out += "# displaced UV sphere, generated (SURVEY.md 8d, config C5)\n";
    char buf[256];
    for (int i = 0; i <= nseg; i++) {
        for (int j = 0; j <= nseg; j++) {
            double th = PI * (double)i / nseg, ph = 2 * PI * (double)j / nseg;
            V3 p = pos(th, ph);
            double e = 1e-4;   // smooth normal from central differences of the displaced surface, oriented outward
            V3 a = pos(th + e, ph) - pos(th - e, ph), b = pos(th, ph + e) - pos(th, ph - e);
            V3 n(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
            double len = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
            if (len < 1e-12 || i == 0 || i == nseg) n = p;
            else if (n.x * p.x + n.y * p.y + n.z * p.z < 0) n = ir::neg(n);
            // shortest round-trip decimals (std::to_chars): the same doubles as "%.17g" would carry, several times faster
            char* q = buf;
            char* const qe = buf + sizeof(buf) - 1;
            auto ch = [&](char c) { if (q < qe) *q++ = c; };
            auto num = [&](double x) { q = std::to_chars(q, qe, x).ptr; };
            ch('v'); ch(' '); num(p.x); ch(' '); num(p.y); ch(' '); num(p.z); ch('\n');
            ch('v'); ch('n'); ch(' '); num(n.x); ch(' '); num(n.y); ch(' '); num(n.z); ch('\n');
            out.append(buf, (size_t)(q - buf));
        }
    }
    for (int i = 0; i < nseg; i++) {
        for (int j = 0; j < nseg; j++) {
            int a = i * (nseg + 1) + j + 1, b = a + 1, c = a + (nseg + 1), d = c + 1;   // OBJ indices are 1-based
            char* q = buf;
            char* const qe = buf + sizeof(buf) - 1;
            auto ch = [&](char c) { if (q < qe) *q++ = c; };
            auto idx = [&](int k) { q = std::to_chars(q, qe, k).ptr; ch('/'); ch('/'); q = std::to_chars(q, qe, k).ptr; };
            ch('f'); ch(' '); idx(a); ch(' '); idx(b); ch(' '); idx(d); ch(' '); idx(c); ch('\n');
            out.append(buf, (size_t)(q - buf));
        }
    }
    return out;
}

// main.go:371-409 with the dragon replaced by the synthetic mesh.
inline void modelExample(Scene& s, CameraConfig& c, const SceneOptions& o) {
    int world = s.NewHittableList();
    s.Add(world, s.NewSphere(V3(0, -1000, 0), 1000, s.NewLambertian(V3(.4, .4, .4))));
    int gold = s.NewMetal(V3(255.0 / 255.0, 215.0 / 255.0, 0), 0.5);
    int nseg = o.mesh_segments > 0 ? o.mesh_segments : 708;
    obj::LoadObjOptions opt;              // objLoader.DefaultLoadOptions(), then main.go:378-382
    opt.ScaleFactor = 5;
    opt.Center = true;
    opt.Position = V3(0, 1.8, 0);
    opt.DefaultMaterial = gold;
    obj::LoadResult lr;
    std::string err;
    if (!obj::loadObj(s, displacedSphereObj(nseg), std::string(), opt, lr, err)) throw std::runtime_error(err);
    int lights = lr.lights;               // no emissive triangles in this mesh (objLoader.go:492-510)
    s.Add(world, s.RotateY(lr.model, 180));
    int light = s.NewSphere(V3(7, 13, 7), 5, s.NewDiffuseLight(V3(4, 4, 4)));
    s.Add(world, light);
    s.Add(lights, light);
    c.AspectRatio = 16.0 / 9.0;
    c.Width = 600;
    c.SamplesPerPixel = 250;
    c.MaxDepth = 50;
    c.Background = V3(0, 0, 0);
    c.VerticalFOV = 40;
    c.MaxContribution = 2.0;
    c.PositionCamera(V3(10, 5, 10), V3(0, 0, 0), V3(0, 1, 0));
    c.DefocusAngle = .1;
    s.world = world;
    s.lights = lights;
    applyOverrides(c, o);
}

// -S values of main.go:449-476.
inline bool buildScene(int id, Scene& s, CameraConfig& c, const SceneOptions& o) {
    switch (id) {
        case 1: book1Scene(s, c, o); return true;
        case 2: book2Scene(s, c, o); return true;
        case 3: book3Scene(s, c, o); return true;
        case 4: simpleLight(s, c, o); return true;
        case 5: quadsScene(s, c, o); return true;
        case 6: cornellBox(s, c, o); return true;
        case 7: cornellSmoke(s, c, o); return true;
        case 8: modelExample(s, c, o); return true;
        default: return false;  // defaultScene is empty (main.go:412)
    }
}

}  // namespace scenes
}  // namespace grt
