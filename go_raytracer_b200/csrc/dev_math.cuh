// dev_math.cuh — fp32 vector helpers and the counter-based RNG (Philox4x32-10)
// that replaces the reference's global math/rand stream (util/utilities.go:12,
// every rand.Float64 call site listed in SURVEY.md §8c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace grtd {

struct f3 {
    float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float t) { return mk3(a.x * t, a.y * t, a.z * t); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float len2(f3 a) { return dot(a, a); }
__device__ __forceinline__ f3 unit(f3 a) { return a * rsqrtf(len2(a)); }
__device__ __forceinline__ f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ d3 mkd3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 tod3(f3 a) { return mkd3((double)a.x, (double)a.y, (double)a.z); }
__device__ __forceinline__ f3 tof3(d3 a) { return mk3((float)a.x, (float)a.y, (float)a.z); }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mkd3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mkd3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double t) { return mkd3(a.x * t, a.y * t, a.z * t); }
__device__ __forceinline__ double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross(d3 a, d3 b) { return mkd3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ d3 ldd3(const double* p) { return mkd3(p[0], p[1], p[2]); }

#define GRT_PI_F 3.14159265358979323846f
#define GRT_PI_D 3.14159265358979323846

// ---- Philox4x32-10 (Salmon et al., SC'11), same constants as oracle.cpp ----
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// uniform in (0,1): (2*(w>>9)+1) / 2^24 — an odd multiple of 2^-24, exactly
// representable in fp32 and strictly inside the interval.
// Built from bits: 0x3f800000 | (w>>9) is 1 + k/2^23 in [1,2); subtracting 1 - 2^-24 gives k/2^23 + 2^-24 exactly
// (the result is representable, so the subtraction is exact).  Two instructions, no int->float conversion.
__host__ __device__ __forceinline__ float u01(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(0x3f800000u | (w >> 9)) - 0.99999994f;   // 0.99999994f == 1 - 2^-24
#else
    return (float)(2u * (w >> 9) + 1u) * (1.0f / 16777216.0f);
#endif
}

#define GRT_STREAM_SHADE 0u
#define GRT_STREAM_MEDIUM 1u
#define GRT_STREAM_CAMERA 2u

// Sequential view of one stream: draw i = word (i&3) of Philox(ctr=(pixel,sample,dim,i>>2)).
struct Rng {
    uint32_t pixel, sample, dim, idx, k0, k1;
    uint32_t b0, b1, b2, b3;
    __device__ __forceinline__ void init(uint32_t pixel_, uint32_t sample_, uint32_t bounce, uint32_t stream, uint32_t k0_, uint32_t k1_) {
        pixel = pixel_; sample = sample_; dim = (bounce << 2) | stream; idx = 0; k0 = k0_; k1 = k1_;
    }
    __device__ __forceinline__ float next() {
        uint32_t lane = idx & 3u;
        if (lane == 0) {
            b0 = pixel; b1 = sample; b2 = dim; b3 = idx >> 2;
            philox4x32_10(b0, b1, b2, b3, k0, k1);
        }
        uint32_t w = lane == 0 ? b0 : (lane == 1 ? b1 : (lane == 2 ? b2 : b3));
        idx++;
        return u01(w);
    }
};
// One isolated draw (used for the rare medium draws so no block is cached).
__device__ __forceinline__ float rng_draw(uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t stream, uint32_t i, uint32_t k0, uint32_t k1) {
    uint32_t c0 = pixel, c1 = sample, c2 = (bounce << 2) | stream, c3 = i >> 2;
    philox4x32_10(c0, c1, c2, c3, k0, k1);
    uint32_t lane = i & 3u;
    uint32_t w = lane == 0 ? c0 : (lane == 1 ? c1 : (lane == 2 ? c2 : c3));
    return u01(w);
}

}  // namespace grtd
