// obj_loader.hpp — host-side mirror of internal/objLoader (objLoader.go, mtlLoader.go): OBJ v/vt/vn/f/usemtl
// parsing with scale / flipYZ / centre / position, fan triangulation, MTL -> material heuristics and the
// emissive-triangle light list.  SURVEY.md §8f rank 3 ("next" row): it runs once on the host and produces
// ordinary Triangles, which ARE on the hot path.  Works on in-memory text so that the harness (and config C5's
// synthetic mesh) needs no files.
//
// Deviations from the reference, all on error paths: log.Fatalf sites return an error string instead of
// exiting; map_Kd / map_Ka image maps need a decoded image supplied by the caller (image decode is out of
// scope) — without one the load fails like Go's "Could not open <file>".
#pragma once
#include <charconv>
#include <cstring>
#include "scene_ir.hpp"
#include <map>
#include <sstream>
#include <functional>
#include <cstdlib>
#include <array>

namespace grt {
namespace obj {

using ir::V3;

struct LoadObjOptions {   // objLoader.go:17-45
    double ScaleFactor = 1.0;
    bool FlipYZ = false;
    bool IgnoreNormals = false;
    bool Center = true;
    bool FlipFaces = false;
    V3 Position = V3(0, 0, 0);
    int DefaultMaterial = -1;   // material id, -1 = Lambertian(0.8) (objLoader.go:88-90)
    bool IgnoreMtl = false;
    bool FindWindows = false;
    // resolves a texture file name to an ir image id (>= 0) or -1 if unavailable
    std::function<int(const std::string&)> imageResolver;
};

struct MtlMaterial {      // mtlLoader.go:18-35 with the defaults of :86-97
    std::string Name;
    V3 Ambient = V3(0.2, 0.2, 0.2), Diffuse = V3(0.8, 0.8, 0.8), Specular = V3(0, 0, 0), Emission = V3(0, 0, 0), Tf = V3(0, 0, 0);
    double SpecExp = 0.0, Dissolve = 1.0, Refraction = 1.0;
    int Illum = 2;
    std::string MapKd, MapKa;
    int Material = -1;
};

inline std::vector<std::string> fields(const std::string& line) {   // strings.Fields
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    auto sp = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    while (i < n) {
        while (i < n && sp(line[i])) i++;
        size_t j = i;
        while (j < n && !sp(line[j])) j++;
        if (j > i) out.emplace_back(line, i, j - i);
        i = j;
    }
    return out;
}
// The same split into a vector that is reused from line to line (a 100 MB OBJ has millions of lines: no per-line
// allocation of the line, the field vector or the fields; tokens of up to 15 characters live in the strings' own buffers).
inline void fieldsInto(const char* a, const char* b, std::vector<std::string>& out) {
    auto sp = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    size_t k = 0;
    while (a < b) {
        while (a < b && sp(*a)) a++;
        const char* e = a;
        while (e < b && !sp(*e)) e++;
        if (e > a) { if (k < out.size()) out[k].assign(a, (size_t)(e - a)); else out.emplace_back(a, (size_t)(e - a)); k++; }
        a = e;
    }
    out.resize(k);
}
// bufio.Scanner over the text: successive lines [a, b) without their '\n' (the last line may lack one)
struct LineScan {
    const char* cur;
    const char* end;
    explicit LineScan(const std::string& t) : cur(t.data()), end(t.data() + t.size()) {}
    bool next(const char*& a, const char*& b) {
        if (cur >= end) return false;
        a = cur;
        const char* nl = (const char*)memchr(cur, '\n', (size_t)(end - cur));
        b = nl ? nl : end;
        cur = nl ? nl + 1 : end;
        return true;
    }
};
// strings.TrimSpace on a range
inline void trimRange(const char*& a, const char*& b) {
    auto sp = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    while (a < b && sp(*a)) a++;
    while (b > a && sp(b[-1])) b--;
}
inline std::string trim(const std::string& s) {   // strings.TrimSpace
    size_t a = s.find_first_not_of(" \t\r\n\v\f"), b = s.find_last_not_of(" \t\r\n\v\f");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline bool parseFloat(const std::string& s, double& v) {   // strconv.ParseFloat(s, 64): the whole token must parse
    if (s.empty()) return false;
    // plain decimal tokens (all of a mesh file) through from_chars: correctly rounded like strtod, several times faster;
    // whatever it does not take whole (a leading '+', hex floats, "Inf") goes to strtod
    const char* b = s.data();
    const char* e = b + s.size();
    auto r = std::from_chars(b, e, v);
    if (r.ec == std::errc() && r.ptr == e) return true;
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return end && *end == 0;
}
inline bool parseInt(const std::string& s, int& v) {        // strconv.Atoi
    if (s.empty()) return false;
    char* end = nullptr;
    long x = std::strtol(s.c_str(), &end, 10);
    if (!end || *end != 0) return false;
    v = (int)x;
    return true;
}
inline std::string joinFrom(const std::vector<std::string>& p, size_t i) {
    std::string s;
    for (size_t k = i; k < p.size(); k++) { if (k > i) s += " "; s += p[k]; }
    return s;
}

// ConvertToRaytracerMaterial, mtlLoader.go:233-326
inline int convertMaterial(ir::Scene& sc, const MtlMaterial& m, const LoadObjOptions& opt, std::string& err) {
    auto imageTex = [&](const std::string& name) -> int {
        int img = opt.imageResolver ? opt.imageResolver(name) : -1;
        if (img < 0) { err = "Could not open " + name; return -1; }   // imageLoader.go:30-33
        return sc.NewImageTexture(img);
    };
    if ((m.Dissolve < 0.95 && m.Refraction > 1.0) || m.Illum == 4 || m.Illum == 6 || m.Illum == 7) {
        double ri = m.Refraction;
        if (ri <= 1.01) ri = 1.5;
        return sc.NewDielectric(ri);
    }
    if (m.Dissolve < 0.95) return sc.NewIsotropic(m.Diffuse);
    double emissive = m.Emission.x + m.Emission.y + m.Emission.z;
    if (emissive > 0.1) {
        if (!m.MapKd.empty()) { int t = imageTex(m.MapKd); return t < 0 ? -1 : sc.NewDiffuseLightTextured(t); }
        if (!m.MapKa.empty()) { int t = imageTex(m.MapKa); return t < 0 ? -1 : sc.NewDiffuseLightTextured(t); }
        return sc.NewDiffuseLight(m.Emission);
    }
    double spec = m.Specular.x + m.Specular.y + m.Specular.z;
    double diff = m.Diffuse.x + m.Diffuse.y + m.Diffuse.z;
    if (spec > 0.1 && spec > diff * 0.5) {
        double roughness;
        if (m.SpecExp <= 0.0) roughness = 1.0;
        else if (m.SpecExp >= 1000.0) roughness = 0.0;
        else {
            roughness = std::pow(1.0 - m.SpecExp / 1000.0, 2.0);
            roughness = std::fmax(0.0, std::fmin(1.0, roughness));
        }
        V3 color = m.Specular;
        if (spec < 0.2) {
            double blend = 1.0 - (spec / 0.2);
            color = V3((1.0 - blend) * m.Specular.x + blend * m.Diffuse.x, (1.0 - blend) * m.Specular.y + blend * m.Diffuse.y,
                       (1.0 - blend) * m.Specular.z + blend * m.Diffuse.z);
        }
        return sc.NewMetal(color, roughness);
    }
    auto diffuse = [&]() -> int {
        if (!m.MapKd.empty()) { int t = imageTex(m.MapKd); return t < 0 ? -1 : sc.NewTexturedLambertian(t); }
        if (!m.MapKa.empty()) { int t = imageTex(m.MapKa); return t < 0 ? -1 : sc.NewTexturedLambertian(t); }
        return sc.NewLambertian(m.Diffuse);
    };
    switch (m.Illum) {
        case 0: case 1: case 2: return diffuse();
        case 3: case 4: case 5: return sc.NewMetal(m.Specular, 0.3);
        default: return diffuse();
    }
}

// LoadMTL, mtlLoader.go:53-230 (on text)
inline bool loadMtl(ir::Scene& sc, const std::string& text, const LoadObjOptions& opt, std::map<std::string, MtlMaterial>& lib, std::string& err) {
    std::istringstream in(text);
    std::string raw;
    MtlMaterial* cur = nullptr;
    auto rgb = [&](const std::vector<std::string>& p, V3& out) {   // ParseFloat errors are ignored there (value 0)
        double r = 0, g = 0, b = 0;
        parseFloat(p[1], r); parseFloat(p[2], g); parseFloat(p[3], b);
        out = V3(r, g, b);
    };
    while (std::getline(in, raw)) {
        std::string line = trim(raw);
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> p = fields(line);
        if (p.empty()) continue;
        const std::string& k = p[0];
        if (k == "newmtl") {
            if (p.size() < 2) continue;
            MtlMaterial m; m.Name = p[1];
            lib[p[1]] = m;
            cur = &lib[p[1]];
            continue;
        }
        if (!cur) continue;
        if ((k == "Ka" || k == "Kd" || k == "Ks" || k == "Ke" || k == "Tf") && p.size() >= 4) {
            V3 v; rgb(p, v);
            if (k == "Ka") cur->Ambient = v; else if (k == "Kd") cur->Diffuse = v; else if (k == "Ks") cur->Specular = v;
            else if (k == "Ke") cur->Emission = v;
            else { cur->Tf = v; cur->Dissolve = (v.x + v.y + v.z) / 3.0; }
        } else if ((k == "Ns" || k == "d" || k == "Ni") && p.size() >= 2) {
            double v = 0; parseFloat(p[1], v);
            if (k == "Ns") cur->SpecExp = v; else if (k == "d") cur->Dissolve = v; else cur->Refraction = v;
        } else if (k == "illum" && p.size() >= 2) { int v = 0; parseInt(p[1], v); cur->Illum = v; }
        else if (k == "map_Kd" && p.size() >= 2) cur->MapKd = joinFrom(p, 1);
        else if (k == "map_Ka" && p.size() >= 2) cur->MapKa = joinFrom(p, 1);
    }
    for (auto& kv : lib) {
        kv.second.Material = convertMaterial(sc, kv.second, opt, err);
        if (kv.second.Material < 0) return false;
    }
    return true;
}

inline int fixIndex(int i, int length) {   // objLoader.go:47-61
    if (i < 0) i = length + i; else i = i - 1;
    if (i < 0 || i >= length) i = (int)std::fmax(0, std::fmin((double)i, (double)(length - 1)));
    return i;
}

struct LoadResult { int model = -1; int lights = -1; int nVertices = 0, nNormals = 0, nTriangles = 0, nLights = 0; };

// LoadObjWithOptions, objLoader.go:72-538 (on text; mtlText may be empty = "no mtllib found / could not load").
inline bool loadObj(ir::Scene& sc, const std::string& objText, const std::string& mtlText, LoadObjOptions opt, LoadResult& res, std::string& err) {
    if (opt.DefaultMaterial < 0) opt.DefaultMaterial = sc.NewLambertian(V3(0.8, 0.8, 0.8));
    std::map<std::string, MtlMaterial> lib;
    bool haveLib = false;
    if (!opt.IgnoreMtl && !mtlText.empty()) {
        // the reference only loads the library when the OBJ names one (mtllib), :105-133
        LineScan scan(objText);
        const char *la, *lb;
        std::vector<std::string> p;
        bool named = false;
        while (scan.next(la, lb)) {
            fieldsInto(la, lb, p);
            if (p.size() >= 2 && p[0] == "mtllib") { named = true; break; }
        }
        if (named) { if (!loadMtl(sc, mtlText, opt, lib, err)) return false; haveLib = true; }
    }
    // first pass: vertices, texture coordinates, bounds (:143-208)
    std::vector<V3> rawVertices, normals;
    std::vector<std::array<double, 2>> texCoords;
    double mn[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
    double mx[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
    {
        LineScan in(objText);
        const char *la, *lb;
        std::vector<std::string> p;
        while (in.next(la, lb)) {
            trimRange(la, lb);
            if (la == lb || *la == '#') continue;
            fieldsInto(la, lb, p);
            if (p.empty()) continue;
            if (p[0] == "vt") {
                if (p.size() < 3) continue;
                double u, v;
                if (!parseFloat(p[1], u) || !parseFloat(p[2], v)) continue;
                texCoords.push_back({u, v});
            }
            if (p[0] == "v") {
                if (p.size() < 4) continue;
                double x, y, z;
                if (!parseFloat(p[1], x) || !parseFloat(p[2], y) || !parseFloat(p[3], z)) continue;
                x *= opt.ScaleFactor; y *= opt.ScaleFactor; z *= opt.ScaleFactor;
                if (opt.FlipYZ) std::swap(y, z);
                rawVertices.push_back(V3(x, y, z));
                mn[0] = std::fmin(mn[0], x); mn[1] = std::fmin(mn[1], y); mn[2] = std::fmin(mn[2], z);
                mx[0] = std::fmax(mx[0], x); mx[1] = std::fmax(mx[1], y); mx[2] = std::fmax(mx[2], z);
            }
        }
    }
    V3 center((mn[0] + mx[0]) / 2, (mn[1] + mx[1]) / 2, (mn[2] + mx[2]) / 2);   // :211-215
    std::vector<V3> vertices;
    vertices.reserve(rawVertices.size());
    for (const V3& v : rawVertices) {                                            // :237-250 (AddInplace twice)
        V3 t = v;
        if (opt.Center) { t = t + ir::neg(center); t = t + opt.Position; }
        vertices.push_back(t);
    }
    // second pass: normals and faces (:284-470)
    int currentMaterial = opt.DefaultMaterial;
    std::vector<int> triangles;
    LineScan in(objText);
    const char *la, *lb;
    std::vector<std::string> p, idx;
    std::vector<V3> fv, fn;
    std::vector<std::array<double, 2>> ft;
    while (in.next(la, lb)) {
        trimRange(la, lb);
        if (la == lb || *la == '#') continue;
        fieldsInto(la, lb, p);
        if (p.empty()) continue;
        if (p[0] == "vn") {
            if (p.size() < 4) continue;
            double nx, ny, nz;
            if (!parseFloat(p[1], nx) || !parseFloat(p[2], ny) || !parseFloat(p[3], nz)) continue;
            if (opt.FlipYZ) std::swap(ny, nz);
            double length = std::sqrt(nx * nx + ny * ny + nz * nz);
            V3 n(nx, ny, nz);
            if (length > 0) n = n * (1.0 / length);
            normals.push_back(n);
        } else if (p[0] == "usemtl") {
            if (opt.IgnoreMtl || !haveLib || p.size() < 2) continue;
            auto it = lib.find(p[1]);
            currentMaterial = it != lib.end() ? it->second.Material : opt.DefaultMaterial;
        } else if (p[0] == "f") {
            if (p.size() < 4) continue;
            fv.clear(); fn.clear(); ft.clear();
            for (size_t i = 1; i < p.size(); i++) {
                // strings.Split(parts[i], "/") into the reused vector
                {
                    size_t a = 0, k = 0;
                    for (;;) {
                        size_t b = p[i].find('/', a);
                        size_t len = (b == std::string::npos ? p[i].size() : b) - a;
                        if (k < idx.size()) idx[k].assign(p[i], a, len); else idx.emplace_back(p[i], a, len);
                        k++;
                        if (b == std::string::npos) break;
                        a = b + 1;
                    }
                    idx.resize(k);
                }
                if (!idx.empty() && !idx[0].empty()) {
                    int k;
                    if (!parseInt(idx[0], k)) continue;
                    int vi = fixIndex(k, (int)vertices.size());
                    if (vi >= 0 && vi < (int)vertices.size()) fv.push_back(vertices[vi]); else continue;
                }
                if (idx.size() > 1 && !idx[1].empty() && !texCoords.empty()) {
                    int k;
                    if (parseInt(idx[1], k)) { int ti = fixIndex(k, (int)texCoords.size()); if (ti >= 0 && ti < (int)texCoords.size()) ft.push_back(texCoords[ti]); }
                }
                if (idx.size() > 2 && !idx[2].empty() && !normals.empty() && !opt.IgnoreNormals) {
                    int k;
                    if (parseInt(idx[2], k)) { int ni = fixIndex(k, (int)normals.size()); if (ni >= 0 && ni < (int)normals.size()) fn.push_back(normals[ni]); }
                }
            }
            if (fv.size() < 3) continue;
            for (size_t i = 2; i < fv.size(); i++) {   // fan triangulation, :396-467
                size_t a = 0, b = i - 1, c = i;
                if (opt.FlipFaces) std::swap(b, c);
                bool hasTex = ft.size() >= fv.size() && ft.size() > i;
                bool hasN = fn.size() >= fv.size() && fn.size() > i && !opt.IgnoreNormals;
                V3 v[3] = {fv[a], fv[b], fv[c]};
                V3 n[3];
                double uv[3][2];
                if (hasN) { n[0] = fn[a]; n[1] = fn[b]; n[2] = fn[c]; }
                if (hasTex) { const size_t o[3] = {a, b, c}; for (int k = 0; k < 3; k++) { uv[k][0] = ft[o[k]][0]; uv[k][1] = ft[o[k]][1]; } }
                triangles.push_back(sc.NewTriangleFull(v, hasN ? n : nullptr, hasTex ? uv : nullptr, currentMaterial));
            }
        }
    }
    if (triangles.empty()) { err = "No triangles found in OBJ file"; return false; }   // :485-487
    // model list, light list, BVH (:489-512)
    int model = sc.NewHittableList();
    int lights = sc.NewHittableList();
    int nl = 0;
    for (int t : triangles) {
        sc.Add(model, t);
        int mt = sc.materials[sc.hittables[t].mat].type;
        if (mt == ir::MAT_DIELECTRIC) { if (opt.FindWindows) { sc.Add(lights, t); nl++; } }
        else if (mt == ir::MAT_DIFFUSE_LIGHT) { sc.Add(lights, t); nl++; }
    }
    res.model = sc.BuildBVH(model);
    res.lights = lights;
    res.nVertices = (int)vertices.size(); res.nNormals = (int)normals.size(); res.nTriangles = (int)triangles.size(); res.nLights = nl;
    return true;
}

}  // namespace obj
}  // namespace grt
