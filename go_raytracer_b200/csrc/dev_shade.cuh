// dev_shade.cuh — textures, light PDFs / sampling, BSDF scatter and the
// per-vertex weight of the reference's integrator.
//
// Reference: internal/hittable/texture.go, perlin.go, materials.go, pdf.go,
// onb.go, objects.go:52-80,152-165,356-385 (light PdfValue/Random),
// hittable.go:89-103 (HittableList.PdfValue/Random), vec/vec.go:136-186,
// camera/camera.go:256-290 (getRay), :319-330 (mixture weighting).
#pragma once
#include "dev_trace.cuh"

namespace grtd {

// ---- Perlin noise, perlin.go:34-111 ----------------------------------------
__device__ __forceinline__ float perlin_noise(const GrtPerlin* P, f3 p) {
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    int i = (int)fx, j = (int)fy, k = (int)fz;
    float uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    float acc = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; di++)
#pragma unroll
        for (int dj = 0; dj < 2; dj++)
#pragma unroll
            for (int dk = 0; dk < 2; dk++) {
                uint32_t h = P->perm[0][(i + di) & 255] ^ P->perm[1][(j + dj) & 255] ^ P->perm[2][(k + dk) & 255];
                float4 g = __ldg((const float4*)&P->grad[h][0]);
                float wx = u - (float)di, wy = v - (float)dj, wz = w - (float)dk;
                acc += ((di ? uu : 1 - uu) * (dj ? vv : 1 - vv) * (dk ? ww : 1 - ww)) * (g.x * wx + g.y * wy + g.z * wz);
            }
    return acc;
}
__device__ __forceinline__ float perlin_turbulence(const GrtPerlin* P, f3 p, int depth) {  // perlin.go:57-69
    float accum = 0.0f, weight = 1.0f;
    for (int i = 0; i < depth; i++) {
        accum += weight * perlin_noise(P, p);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(accum);
}

// ---- Texture.Value, texture.go:25,50,70,112 ----------------------------------
template <uint32_t FEAT>
__device__ __forceinline__ f3 texture_value(const SceneView& sv, uint32_t tex, float u, float v, f3 p) {
    const GrtTexture* T = sv.textures() + tex;
    if (FEAT & F_TEXTURE) {
        // checkerboards may nest (even/odd are textures): walk down
        for (int guard = 0; guard < 8 && T->type == GRT_TEX_CHECKER; guard++) {  // texture.go:50-59
            int x = (int)floorf(T->scale * p.x), y = (int)floorf(T->scale * p.y), z = (int)floorf(T->scale * p.z);
            T = sv.textures() + (((x + y + z) % 2 == 0) ? T->even : T->odd);
        }
        if (T->type == GRT_TEX_IMAGE) {  // texture.go:70-91, imageLoader.go:52-62
            const GrtImage im = sv.images()[T->aux];
            if (im.height == 0) return mk3(0, 1, 1);
            float uu = fabsf(fmodf(u, 1.0f));
            float vv = 1.0f - fabsf(fmodf(v, 1.0f));
            int i = (int)(uu * (float)(im.width - 1));
            int j = (int)(vv * (float)(im.height - 1));
            i = min(max(i, 0), (int)im.width); j = min(max(j, 0), (int)im.height);
            size_t idx = (size_t)j * im.width + i;
            if (idx >= (size_t)im.width * im.height) return mk3(1, 0, 1);  // magenta past the end
            const uint8_t* px = sv.ds->texels + im.offset + idx * 3;
            const float s = 1.0f / 255.0f;
            return mk3((float)px[0] * s, (float)px[1] * s, (float)px[2] * s);
        }
        if (T->type == GRT_TEX_NOISE) {  // texture.go:112-125
            const GrtPerlin* P = sv.ds->perlins + (T->aux & 0xFFFFu);
            uint32_t variant = T->aux >> 16;
            if (variant == GRT_NOISE_MARBLE) {
                float s = 0.5f * (1.0f + sinf(T->scale * p.z + 10.0f * perlin_turbulence(P, p, 7)));
                return mk3(s, s, s);
            }
            if (variant == GRT_NOISE_TURBULENT) {
                float s = perlin_turbulence(P, p, 7);
                return mk3(s, s, s);
            }
            float s = 0.5f * (1.0f + perlin_noise(P, p * T->scale));
            return mk3(s, s, s);
        }
    }
    return mk3(T->color[0], T->color[1], T->color[2]);
}
// does evaluating this material's texture need the sphere's (u,v)?
template <uint32_t FEAT>
__device__ __forceinline__ bool material_needs_uv(const SceneView& sv, const GrtMaterial& m) {
    if (!(FEAT & F_TEXTURE)) return false;
    if (m.type == GRT_MAT_METAL || m.type == GRT_MAT_DIELECTRIC) return false;
    uint32_t t = sv.textures()[m.tex].type;
    return t == GRT_TEX_IMAGE || t == GRT_TEX_CHECKER;  // a checker may contain an image
}

// ---- orthonormal basis, onb.go:13-43 (fp32 and fp64) --------------------------
struct Onb {
    f3 u, v, w;
};
__device__ __forceinline__ Onb make_onb(f3 n) {
    Onb b;
    b.w = unit(n);
    f3 a = fabsf(n.x) > 0.9f ? mk3(0, 1, 0) : mk3(1, 0, 0);
    b.v = unit(cross(n, a));
    b.u = unit(cross(n, b.v));
    return b;
}
__device__ __forceinline__ f3 onb_transform(const Onb& b, f3 v) { return b.u * v.x + b.v * v.y + b.w * v.z; }

__device__ __forceinline__ d3 unit64(d3 a) { return a * (1.0 / sqrt(dot(a, a))); }

// ---- light Random / PdfValue in fp64 ----------------------------------------
// The reference samples a light point and then RE-INTERSECTS the light to get
// its pdf (objects.go:52-62,152-160,356-367).  A sample on the light's edge
// must not be lost to rounding (a zero pdf with a below-horizon direction is
// 0/0 = NaN and blacks the pixel, color.go:28-36), so this arithmetic stays in
// fp64 like the reference; it is ~1 light evaluation per diffuse vertex.
template <uint32_t FEAT>
__device__ __forceinline__ d3 light_random(const GrtLight* L, d3 origin, float r1, float r2) {
    if ((FEAT & F_QUAD_LIGHT) && L->type == GRT_LIGHT_QUAD) {  // objects.go:161-165
        d3 Q = ldd3(L->p), u = ldd3(L->p + 3), v = ldd3(L->p + 6);
        d3 p = Q + u * (double)r1 + v * (double)r2;
        return p - origin;
    }
    if ((FEAT & F_SPHERE_LIGHT) && L->type == GRT_LIGHT_SPHERE) {  // objects.go:63-80
        d3 dir = ldd3(L->p) - origin;
        double distSquared = dot(dir, dir);
        double R = L->p[3];
        // NewONB(direction)
        d3 w = unit64(dir);
        d3 a = fabs(dir.x) > 0.9 ? mkd3(0, 1, 0) : mkd3(1, 0, 0);
        d3 vv = unit64(cross(dir, a));
        d3 uu = unit64(cross(dir, vv));
        double z = 1 + (double)r2 * (sqrt(1 - R * R / distSquared) - 1);
        double phi = 2 * GRT_PI_D * (double)r1;
        double t = sqrt(1 - z * z);
        double sn, cs;
        sincos(phi, &sn, &cs);
        double x = cs * t, y = sn * t;
        return uu * x + vv * y + w * z;
    }
    if ((FEAT & F_TRI_LIGHT) && L->type == GRT_LIGHT_TRI) {  // objects.go:369-385
        double a1 = (double)r1;
        double a2 = (double)r2 * (1 - a1);
        double a = 1 - a1 - a2;
        d3 p = ldd3(L->p) * a + ldd3(L->p + 3) * a1 + ldd3(L->p + 6) * a2;
        return p - origin;
    }
    return mkd3(1, 0, 0);  // defaultPdfImpl.Random, hittable.go:73-75
}

template <uint32_t FEAT>
__device__ __forceinline__ double light_pdf(const GrtLight* L, d3 o, d3 d) {
    if ((FEAT & F_QUAD_LIGHT) && L->type == GRT_LIGHT_QUAD) {  // objects.go:152-160 + quad.Hit
        d3 n = ldd3(L->p + 9);
        double denom = dot(n, d);
        if (fabs(denom) < 1e-8) return 0;
        double t = (L->p[15] - dot(n, o)) / denom;
        if (!(0.001 <= t)) return 0;
        d3 Q = ldd3(L->p), u = ldd3(L->p + 3), v = ldd3(L->p + 6), w = ldd3(L->p + 12);
        d3 planar = (o + d * t) - Q;
        double alpha = dot(w, cross(planar, v));
        double beta = dot(w, cross(u, planar));
        if (!(0 <= alpha && alpha <= 1) || !(0 <= beta && beta <= 1)) return 0;
        double dd = dot(d, d);
        double distSquared = t * t * dd;
        double cosine = fabs(denom / sqrt(dd));
        return distSquared / (cosine * L->p[16]);
    }
    if ((FEAT & F_SPHERE_LIGHT) && L->type == GRT_LIGHT_SPHERE) {  // objects.go:52-62 + sphere.Hit at time 0
        d3 oc = ldd3(L->p) - o;
        double R = L->p[3];
        double a = dot(d, d), h = dot(d, oc), c = dot(oc, oc) - R * R;
        double disc = h * h - a * c;
        if (disc < 0) return 0;
        double sq = sqrt(disc);
        double root = (h - sq) / a;
        if (!(0.0001 < root)) {
            root = (h + sq) / a;
            if (!(0.0001 < root)) return 0;
        }
        double distSquared = dot(oc, oc);
        double cosThetaMax = sqrt(1 - R * R / distSquared);
        double solidAngle = 2 * GRT_PI_D * (1 - cosThetaMax);
        return 1 / solidAngle;
    }
    if ((FEAT & F_TRI_LIGHT) && L->type == GRT_LIGHT_TRI) {  // objects.go:356-367 + Triangle.Hit
        d3 v0 = ldd3(L->p), e0 = ldd3(L->p + 3) - v0, e1 = ldd3(L->p + 6) - v0;
        d3 pvec = cross(d, e1);
        double det = dot(e0, pvec);
        if (fabs(det) < 1e-8) return 0;
        double invDet = 1.0 / det;
        d3 tvec = o - v0;
        double u = dot(tvec, pvec) * invDet;
        if (u < 0 || u > 1) return 0;
        d3 qvec = cross(tvec, e0);
        double v = dot(d, qvec) * invDet;
        if (v < 0 || (u + v) > 1) return 0;
        double tl = dot(e1, qvec) * invDet;
        if (tl < 0.001) return 0;
        d3 nrm;
        if (L->flags & GRT_TRI_HAS_NORMALS) {
            double w = 1.0 - u - v;
            nrm = unit64(ldd3(L->p + 10) * w + ldd3(L->p + 13) * u + ldd3(L->p + 16) * v);
        } else {
            nrm = unit64(cross(e0, e1));
        }
        double dd = dot(d, d);
        double distSquared = tl * tl * dd;
        double cosine = fabs(dot(d, nrm) / sqrt(dd));
        return distSquared / (cosine * L->p[9]);
    }
    return 0;
}

// fp32 fast path of the quad light pdf (objects.go:152-160 + quad.Hit) on the packed DLight record.
// Returns 1 with *pdf set when the fp32 result is certain (the hit is at least 1e-4 inside / outside
// every decision boundary: alpha, beta in [0,1], t >= 0.001, not grazing), 0 when a decision is within
// fp32 error of a boundary and the fp64 evaluation must be used.
__device__ __forceinline__ int quad_light_pdf_fast(const DLight* L, f3 o, f3 d, float* pdf) {
    const float4 P = L->plane, A = L->A, B = L->B;
    const float denom = P.x * d.x + P.y * d.y + P.z * d.z;
    const float t = fast_div(P.w - (P.x * o.x + P.y * o.y + P.z * o.z), denom);
    const float px = fmaf(t, d.x, o.x), py = fmaf(t, d.y, o.y), pz = fmaf(t, d.z, o.z);
    const float alpha = fmaf(A.x, px, fmaf(A.y, py, fmaf(A.z, pz, A.w)));
    const float beta = fmaf(B.x, px, fmaf(B.y, py, fmaf(B.z, pz, B.w)));
    const float dd = d.x * d.x + d.y * d.y + d.z * d.z;
    const float rlen = rsqrtf(dd);
    const float EPS = 1e-4f;
    const float cosine = fabsf(denom) * rlen;                       // |d.n| / |d|   (objects.go:158)
    const float lo = fminf(fminf(alpha, beta), fminf(1.0f - alpha, 1.0f - beta));   // signed distance to the nearest edge
    const bool inside = (lo > EPS) & (t > 0.002f);
    const bool outside = (lo < -EPS) | (t < 0.0005f);
    if (!(cosine >= 1e-3f) || !(inside || outside)) return 0;      // grazing, near an edge / tmin, or NaN
    *pdf = inside ? fast_div(t * t * dd, cosine * L->Qa.w) : 0.0f;  // distSquared / (cosine * area)
    return 1;
}

// ---- vec.go sampling routines -------------------------------------------------
__device__ __forceinline__ f3 random_unit_vector(Rng& rng) {  // vec.go:159-167
    for (;;) {
        float x = -1.0f + 2.0f * rng.next(), y = -1.0f + 2.0f * rng.next(), z = -1.0f + 2.0f * rng.next();
        float l2 = x * x + y * y + z * z;
        if (1e-30f < l2 && l2 <= 1.0f) { float s = 1.0f / sqrtf(l2); return mk3(x * s, y * s, z * s); }
    }
}
__device__ __forceinline__ f3 random_cosine_direction(float r1, float r2) {  // vec.go:177-186
    // phi = 2 pi r1 in (0, 2 pi); evaluate at phi - pi in (-pi, pi), where sin.approx / cos.approx are accurate to
    // 2^-21 absolute (a 5e-7 rad perturbation of a random direction), and flip the signs
    float sn, cs;
    __sincosf(2.0f * GRT_PI_F * (r1 - 0.5f), &sn, &cs);
    sn = -sn; cs = -cs;
    float sr = sqrtf(r2);
    return mk3(cs * sr, sn * sr, sqrtf(1 - r2));
}
__device__ __forceinline__ f3 reflect(f3 v, f3 n) { return v - n * (dot(n, v) * 2); }  // vec.go:136
__device__ __forceinline__ f3 refract(f3 v, f3 n, float eta) {                           // vec.go:141-146
    float cosTheta = fminf(dot(-v, n), 1.0f);
    f3 rPerp = (v + n * cosTheta) * eta;
    f3 rPar = n * (-sqrtf(fabsf(1.0f - len2(rPerp))));
    return rPerp + rPar;
}

// ---- camera ray, camera.go:256-290 ----------------------------------------------
struct DevCamera {
    f3 center, p00_rel, du, dv, defu, defv;   // p00_rel = pixel00Loc - center (fp64 difference, rounded once)
    f3 background;
    float max_contribution, recip_spp_sqrt, defocus_angle;
    int width, height, spp_sqrt, max_depth;
};
template <uint32_t FEAT>
__device__ __forceinline__ void camera_ray(const DevCamera& cam, int px, int py, uint32_t sample, uint32_t pixel_index, uint32_t k0, uint32_t k1,
                                           f3& o, f3& d, float& time) {
    Rng rng;
    rng.init(pixel_index, sample, 0, GRT_STREAM_CAMERA, k0, k1);
    // renderRow calls getRay(j, row, s_j, s_i) (camera.go:97-99): the x stratum is the inner loop index
    uint32_t s_y;
    if (cam.spp_sqrt <= 1024) s_y = (uint32_t)(((float)sample + 0.5f) * cam.recip_spp_sqrt);   // exact: sample < 2^20, margin 0.5/S
    else s_y = sample / (uint32_t)cam.spp_sqrt;
    const uint32_t s_x = sample - s_y * (uint32_t)cam.spp_sqrt;
    float ox = (((float)s_x + rng.next()) * cam.recip_spp_sqrt) - 0.5f;   // sampleSquareStratified :277-282
    float oy = (((float)s_y + rng.next()) * cam.recip_spp_sqrt) - 0.5f;
    f3 rel = cam.p00_rel + cam.du * ((float)px + ox) + cam.dv * ((float)py + oy);   // pixelSample - center
    o = cam.center;
    if ((FEAT & F_DEFOCUS) && cam.defocus_angle > 0) {   // defocusDiskSample :285-290, RandomUnitDisk vec.go:149-156
        float dx, dy;
        for (;;) {
            dx = -1.0f + 2.0f * rng.next(); dy = -1.0f + 2.0f * rng.next();
            if (dx * dx + dy * dy < 1.0f) break;
        }
        f3 off = cam.defu * dx + cam.defv * dy;
        o = cam.center + off;
        rel = rel - off;
    }
    d = rel;
    time = rng.next();   // camera.go:268
}

// ---- one shading vertex ------------------------------------------------------------
// Outcome of Material.Scatter + the PDF mixture of camera.go:312-328.
enum ShadeKind { SHADE_TERMINATE = 0, SHADE_SPECULAR = 1, SHADE_DIFFUSE = 2, SHADE_NAN = 3 };
struct ShadeResult {
    int kind;
    f3 value;     // TERMINATE: radiance returned by this vertex (emission / 0); SPECULAR: attenuation;
                  // DIFFUSE: weight = Attenuation * scatterPdf / pdfValue
    f3 dir;       // next direction
};

template <uint32_t FEAT>
__device__ __forceinline__ ShadeResult shade_vertex(const SceneView& sv, const RayD& ray, const Surface& s, const GrtMaterial& m,
                                                    uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t k0, uint32_t k1, uint32_t* n_lightpdf) {
    ShadeResult R;
    R.dir = mk3(0, 0, 0);
    if (m.type == GRT_MAT_DIFFUSE_LIGHT) {  // materials.go:142-155, camera.go:305-314
        R.kind = SHADE_TERMINATE;
        R.value = s.front ? texture_value<FEAT>(sv, m.tex, s.u, s.v, s.p) : mk3(0, 0, 0);
        return R;
    }
    Rng rng;
    rng.init(pixel, sample, bounce, GRT_STREAM_SHADE, k0, k1);
    if ((FEAT & F_SPECULAR) && m.type == GRT_MAT_METAL) {  // materials.go:70-79
        f3 refl = reflect(ray.d, s.n);
        refl = unit(refl) + random_unit_vector(rng) * m.fuzz;
        R.kind = SHADE_SPECULAR; R.value = mk3(m.albedo[0], m.albedo[1], m.albedo[2]); R.dir = refl;
        return R;
    }
    if ((FEAT & F_SPECULAR) && m.type == GRT_MAT_DIELECTRIC) {  // materials.go:94-130
        float ri = s.front ? 1.0f / m.ior : m.ior;
        f3 ud = unit(ray.d);
        float cosTheta = fminf(dot(-ud, s.n), 1.0f);
        float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
        bool cannotRefract = ri * sinTheta > 1.0f;
        bool refl = cannotRefract;
        if (!refl) {
            float r0 = (1.0f - m.ior) / (1.0f + m.ior);
            r0 *= r0;
            float x = 1 - cosTheta, x2 = x * x;
            float reflectance = r0 + (1 - r0) * (x2 * x2 * x);   // math.Pow(1-cos, 5)
            refl = reflectance > rng.next();
        }
        R.kind = SHADE_SPECULAR; R.value = mk3(1, 1, 1);
        R.dir = refl ? reflect(ud, s.n) : refract(ud, s.n, ri);
        return R;
    }
    // Lambertian (CosinePdf) or Isotropic (SpherePdf): mixture with the light pdf
    const bool iso = (FEAT & F_ISOTROPIC) && m.type == GRT_MAT_ISOTROPIC;
    f3 att = texture_value<FEAT>(sv, m.tex, s.u, s.v, s.p);
    const uint32_t nl = sv.ds->n_lights;
    const GrtLight* lights = sv.lights();
    f3 wn = unit(s.n);   // onb.W
    f3 dir;
    bool from_light = false;
    uint32_t li = 0;
    float r1 = 0, r2 = 0;
    if (rng.next() < 0.5f) {  // pdf.go:69-74: p[0] = light
        if (sv.ds->lights_mode == GRT_LIGHTS_LIST && nl == 0) {  // hittable.go:98-101: vec.Random()
            float x = rng.next(), y = rng.next(), z = rng.next();
            dir = mk3(x, y, z);
        } else {
            if (sv.ds->lights_mode == GRT_LIGHTS_LIST) { li = (uint32_t)(rng.next() * (float)nl); if (li >= nl) li = nl - 1; }  // rand.Intn
            r1 = rng.next(); r2 = rng.next();
            from_light = true;
            const GrtLight* L = lights + li;
            if ((FEAT & F_QUAD_LIGHT) && L->type == GRT_LIGHT_QUAD) {   // objects.go:161-165
                const DLight* D = sv.dlights() + li;
                const float4 Q = D->Qa, U = D->U, V = D->V;
                dir = mk3(fmaf(V.x, r2, fmaf(U.x, r1, Q.x)) - s.p.x, fmaf(V.y, r2, fmaf(U.y, r1, Q.y)) - s.p.y, fmaf(V.z, r2, fmaf(U.z, r1, Q.z)) - s.p.z);
            }
            else dir = tof3(light_random<FEAT>(L, tod3(s.p), r1, r2));
        }
    } else {
        if (iso) dir = random_unit_vector(rng);                                   // pdf.go:21-23
        else {  // pdf.go:38-40
            float c1 = rng.next(), c2 = rng.next();
            Onb b;
            if (s.has_onb) { b.u = s.ou; b.v = s.front ? s.ov : -s.ov; b.w = s.n; }   // NewONB(-n) = (u, -v, -n)
            else b = make_onb(s.n);
            dir = onb_transform(b, random_cosine_direction(c1, c2));
        }
    }
    // mixPdf.Value: 0.5 * lights.PdfValue + 0.5 * material pdf  (pdf.go:65-67, hittable.go:89-96)
    float lp = 0.0f;
    if (nl > 0) {
        const float weight = 1.0f / (float)nl;
        for (uint32_t i = 0; i < nl; i++) {
            const GrtLight* L = lights + i;
            float pv;
            if (!((FEAT & F_QUAD_LIGHT) && L->type == GRT_LIGHT_QUAD && quad_light_pdf_fast(sv.dlights() + i, s.p, dir, &pv))) {
                // within fp32 error of a decision boundary (or not a quad): the reference's fp64 arithmetic.
                // A direction sampled ON this light is regenerated in fp64 so that it cannot fall off its edge.
#ifndef GRT_NO_FALLBACK_FENCE
                asm volatile("" ::: "memory");   // keep the loads of this 0.1 % path from being scheduled above the branch
#endif
                d3 p64 = tod3(s.p);
                d3 dir64 = (from_light && i == li) ? light_random<FEAT>(L, p64, r1, r2) : tod3(dir);
                pv = (float)light_pdf<FEAT>(L, p64, dir64);
#ifdef GRT_DEBUG_COUNT_F64
                if (n_lightpdf) *n_lightpdf += 1000000u;   // (debug builds: counts fp64 fallbacks in the high digits of light_pdf_evals)
#endif
            }
            lp += weight * pv;
        }
        if (n_lightpdf) *n_lightpdf += nl;
    }
    float mp, sp;
    if (iso) { mp = 1.0f / (4 * GRT_PI_F); sp = mp; }
    else {
        float cosTheta = dot(unit(dir), wn);
        mp = fmaxf(0.0f, cosTheta * (1.0f / GRT_PI_F));               // pdf.go:33-36
        sp = cosTheta < 0 ? 0.0f : cosTheta * (1.0f / GRT_PI_F);      // materials.go:51-57
    }
    float pdfValue = 0.5f * lp + 0.5f * mp;
    if (pdfValue == 0.0f && sp == 0.0f) { R.kind = SHADE_NAN; R.value = mk3(0, 0, 0); return R; }  // 0 * x * (1/0) = NaN (camera.go:328)
    R.kind = SHADE_DIFFUSE;
    R.value = att * __fdividef(sp, pdfValue);
    R.dir = dir;
    return R;
}

__device__ __forceinline__ f3 clamp_contribution(f3 c, float maxValue) {  // camera.go:334-341
    float intensity = c.x + c.y + c.z;
    if (intensity > maxValue) return c * (maxValue / intensity);
    return c;
}

// ---- recursive firefly clamp (camera.go:327-341) without recursion -------------------------
// See DESIGN.md §3.2.  T = running product of the non-zero weight components, zinfo tracks
// exactly-zero components, rstack[j] = 1/T_before_j for each clamped vertex.
#define WEIGHT_STACK 64

// T *= w, except that exactly-zero components are recorded in zinfo instead: byte c = number of stack
// entries whose suffix contains a zero in component c, bit 24+c = a zero was seen at all.
__device__ __forceinline__ void apply_factor(f3& T, uint32_t& zinfo, f3 w, int sp_after) {
    if (w.x == 0.0f | w.y == 0.0f | w.z == 0.0f) {
        if (w.x == 0.0f) { zinfo = (zinfo & ~0x000000ffu) | (uint32_t)sp_after | (1u << 24); w.x = 1.0f; }
        if (w.y == 0.0f) { zinfo = (zinfo & ~0x0000ff00u) | ((uint32_t)sp_after << 8) | (1u << 25); w.y = 1.0f; }
        if (w.z == 0.0f) { zinfo = (zinfo & ~0x00ff0000u) | ((uint32_t)sp_after << 16) | (1u << 26); w.z = 1.0f; }
    }
    T = T * w;
}
__device__ __forceinline__ float4 recip_factor(f3 T) { return make_float4(fast_div(1.0f, T.x), fast_div(1.0f, T.y), fast_div(1.0f, T.z), 0.0f); }

// L0 = P0 * min(1, M / max_j sum(P_j)),  P0 = T (x) E,  sum(P_j) = P0 . rstack[j] over the components whose
// suffix from j holds no zero factor.  `stride` lets the wavefront variant keep the stack depth-major in HBM.
template <class StackPtr>
__device__ __forceinline__ f3 unwind_clamp(f3 T, uint32_t zinfo, f3 E, StackPtr rstack, int sp, float max_contribution, size_t stride = 1) {
    f3 L = T * E;
    float worst = 0.0f;   // fmaxf drops the NaN of 0 * inf (an underflowed T component), which contributes nothing anyway
    if (zinfo == 0u) {
        for (int i = 0; i < sp; i++) {
            const float4 rj = rstack[(size_t)i * stride];
            worst = fmaxf(worst, fmaf(L.x, rj.x, fmaf(L.y, rj.y, L.z * rj.z)));
        }
    } else {
        const int zx = (int)(zinfo & 255u), zy = (int)((zinfo >> 8) & 255u), zz = (int)((zinfo >> 16) & 255u);
        for (int i = 0; i < sp; i++) {
            const float4 rj = rstack[(size_t)i * stride];
            float sj = (i >= zx ? L.x * rj.x : 0.0f) + (i >= zy ? L.y * rj.y : 0.0f) + (i >= zz ? L.z * rj.z : 0.0f);
            worst = fmaxf(worst, sj);
        }
        if (zinfo & (1u << 24)) L.x = 0.0f;
        if (zinfo & (1u << 25)) L.y = 0.0f;
        if (zinfo & (1u << 26)) L.z = 0.0f;
    }
    if (worst > max_contribution) L = L * fast_div(max_contribution, worst);
    return L;
}

// The same unwind for a stack whose first NS entries live in shared memory (`near(i)`) and the rest in local memory
// (`deep(i)`), as the megakernel keeps it.  One predicated, fully unrolled pass over the NS near entries and a (rare)
// loop over the deep ones: choosing the memory per element inside one loop compiled to a branch per entry (7 % of
// the Cornell kernel's instructions, ncu round 2).
template <int NS, class Near, class Deep>
__device__ __forceinline__ f3 unwind_clamp_split(f3 T, uint32_t zinfo, f3 E, Near near, Deep deep, int sp, float max_contribution) {
    f3 L = T * E;
    float worst = 0.0f;
    if (zinfo == 0u) {
#pragma unroll
        for (int i = 0; i < NS; i++) {
            if (i < sp) { const float4 rj = near(i); worst = fmaxf(worst, fmaf(L.x, rj.x, fmaf(L.y, rj.y, L.z * rj.z))); }
        }
        for (int i = NS; i < sp; i++) { const float4 rj = deep(i); worst = fmaxf(worst, fmaf(L.x, rj.x, fmaf(L.y, rj.y, L.z * rj.z))); }
    } else {
        const int zx = (int)(zinfo & 255u), zy = (int)((zinfo >> 8) & 255u), zz = (int)((zinfo >> 16) & 255u);
        for (int i = 0; i < sp; i++) {
            const float4 rj = i < NS ? near(i) : deep(i);
            float sj = (i >= zx ? L.x * rj.x : 0.0f) + (i >= zy ? L.y * rj.y : 0.0f) + (i >= zz ? L.z * rj.z : 0.0f);
            worst = fmaxf(worst, sj);
        }
        if (zinfo & (1u << 24)) L.x = 0.0f;
        if (zinfo & (1u << 25)) L.y = 0.0f;
        if (zinfo & (1u << 26)) L.z = 0.0f;
    }
    if (worst > max_contribution) L = L * fast_div(max_contribution, worst);
    return L;
}

}  // namespace grtd
