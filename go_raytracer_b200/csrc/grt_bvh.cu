// grt_bvh.cu — BuildBVH's object order on the GPU.
//
// bvhHelper (bvh.go:35-61) sorts every span of three or more objects along the longest axis of the span's
// bounding box (boxCompare: box minimum, then box maximum, bvh.go:25-32), splits it at the median and recurses:
// O(n log^2 n) comparisons through an interface call, 20 levels deep for a million triangles.  Because the split
// is always the median, the SHAPE of the tree depends only on n; the only data-dependent result is the final
// order of the objects.  This file computes that order level by level on the device:
//
//   per level:  span of every position (binary search in the level's span table)
//               -> bounding box of every span (warp-aggregated atomic min/max on order-preserving integers)
//               -> longest axis of every span (aabb.go:73-87, with the padding of aabb.go:118-129)
//               -> one stable sort of all positions by (span, box min on the span's axis, box max)
//
// The sort is cub::DeviceMergeSort::StableSortPairs with the reference's comparison on fp64 keys, so the result is
// the reference's order with ties kept in list order (Go's sort.Slice is unstable; ties are documented as
// unordered).  The host flattener (flatten.hpp) then walks the known shape without sorting.
// One documented difference: the reference pads the running union after every object (aabb.go:54-59); here the
// union is padded once.  The two can only differ for spans thinner than 2e-4 on their LONGEST axis.
#include <cuda_runtime.h>
#include <cub/device/device_merge_sort.cuh>
#include <stdint.h>
#include <string>
#include <vector>
#include "grt_internal.h"
#include "../../include/grt.h"

namespace {

struct SortKey {
    uint32_t group;   // first position of the span being sorted, or the position itself outside such spans
    double kmin, kmax;
};
struct SortLess {   // boxCompare, bvh.go:25-32, below the span grouping
    __device__ __forceinline__ bool operator()(const SortKey& a, const SortKey& b) const {
        if (a.group != b.group) return a.group < b.group;
        if (a.kmin != b.kmin) return a.kmin < b.kmin;
        return a.kmax < b.kmax;
    }
};

// order-preserving map double -> uint64 (for atomicMin / atomicMax); -0.0 is folded onto +0.0 first
__device__ __forceinline__ unsigned long long ord64(double x) {
    x += 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unord64(unsigned long long u) {
    unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

// spans: this level's sorted list of [start, end) with end - start >= 3, ascending and disjoint
__global__ void bvh_span_of_pos(const uint2* __restrict__ spans, uint32_t n_spans, uint32_t n, uint32_t* __restrict__ span_of) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t lo = 0, hi = n_spans;   // last span with start <= p
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (spans[m].x <= p) lo = m + 1; else hi = m; }
    uint32_t s = 0xFFFFFFFFu;
    if (lo > 0 && p < spans[lo - 1].y) s = lo - 1;
    span_of[p] = s;
}

__global__ void bvh_init_boxes(unsigned long long* __restrict__ sbox, uint32_t n_spans) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spans * 6u) return;
    sbox[i] = (i % 6u) < 3u ? 0xFFFFFFFFFFFFFFFFull : 0ull;   // lo: +max, hi: 0 (below every ordered value)
}

// bounding box of every span: union of the boxes of its objects (aabb.go:54-59 without the per-step padding)
__global__ void bvh_span_boxes(const double* __restrict__ boxes, const uint32_t* __restrict__ order, const uint32_t* __restrict__ span_of,
                               uint32_t n, unsigned long long* __restrict__ sbox) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    uint32_t s = p < n ? span_of[p] : 0xFFFFFFFFu;
    unsigned long long v[6];
    if (s != 0xFFFFFFFFu) {
        const double* b = boxes + 6 * (size_t)order[p];
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = ord64(b[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = k < 3 ? 0xFFFFFFFFFFFFFFFFull : 0ull;
    }
    // the 32 positions of a warp usually belong to one span: reduce in the warp, one lane does the atomics
    const uint32_t s0 = __shfl_sync(FULL, s, 0);
    if (__all_sync(FULL, s == s0)) {
        if (s0 == 0xFFFFFFFFu) return;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            unsigned long long x = v[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                unsigned long long y = __shfl_xor_sync(FULL, x, off);
                x = k < 3 ? (y < x ? y : x) : (y > x ? y : x);
            }
            v[k] = x;
        }
        if ((threadIdx.x & 31u) == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) { atomicMin(sbox + 6 * (size_t)s0 + k, v[k]); atomicMax(sbox + 6 * (size_t)s0 + 3 + k, v[3 + k]); }
        }
    } else if (s != 0xFFFFFFFFu) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(sbox + 6 * (size_t)s + k, v[k]); atomicMax(sbox + 6 * (size_t)s + 3 + k, v[3 + k]); }
    }
}

// LongestAxis (aabb.go:73-87) of the padded box (aabb.go:118-129, interval.go:47-50)
__global__ void bvh_span_axis(const unsigned long long* __restrict__ sbox, uint32_t n_spans, uint32_t* __restrict__ axis) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spans) return;
    double size[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double lo = unord64(sbox[6 * (size_t)i + a]), hi = unord64(sbox[6 * (size_t)i + 3 + a]);
        if (hi - lo < 0.0001) { const double pad = 0.0001 / 2; lo = lo - pad; hi = hi + pad; }
        size[a] = hi - lo;
    }
    uint32_t ax;
    if (size[0] > size[1]) ax = size[0] > size[2] ? 0u : 2u;
    else ax = size[1] > size[2] ? 1u : 2u;
    axis[i] = ax;
}

__global__ void bvh_make_keys(const double* __restrict__ boxes, const uint32_t* __restrict__ order, const uint32_t* __restrict__ span_of,
                              const uint2* __restrict__ spans, const uint32_t* __restrict__ axis, uint32_t n, SortKey* __restrict__ keys) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t s = span_of[p];
    SortKey k;
    if (s == 0xFFFFFFFFu) { k.group = p; k.kmin = 0.0; k.kmax = 0.0; }
    else {
        const uint32_t a = axis[s];
        const double* b = boxes + 6 * (size_t)order[p];
        k.group = spans[s].x; k.kmin = b[a]; k.kmax = b[3 + a];
    }
    keys[p] = k;
}

__global__ void bvh_iota(uint32_t* order, uint32_t n) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) order[p] = p;
}

}  // namespace

#define BVH_TRY(call)                                                                                                  \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) { grt_set_error(std::string("grt_bvh_order: ") + #call + ": " + cudaGetErrorString(e_)); rc = GRT_E_CUDA; goto done; } \
    } while (0)

extern "C" int grt_bvh_order(const double* boxes, uint32_t n, int device, uint32_t* order_out) {
    if (!boxes || !order_out) { grt_set_error("grt_bvh_order: NULL argument"); return GRT_E_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { grt_set_error("no CUDA device (the backend has no CPU fallback)"); return GRT_E_NO_DEVICE; }
    if (device < 0 || device >= ndev) { grt_set_error("grt_bvh_order: device out of range"); return GRT_E_INVALID; }
    if (n == 0) return GRT_OK;
    // the shape of the tree (bvh.go:41-58): spans of >= 3 objects are sorted and split at the median
    std::vector<std::vector<uint2>> levels;
    {
        std::vector<uint2> cur;
        if (n >= 3) cur.push_back(make_uint2(0u, n));
        while (!cur.empty()) {
            std::vector<uint2> next;
            next.reserve(cur.size() * 2);
            for (const uint2& s : cur) {
                const uint32_t mid = s.x + (s.y - s.x) / 2;
                if (mid - s.x >= 3) next.push_back(make_uint2(s.x, mid));
                if (s.y - mid >= 3) next.push_back(make_uint2(mid, s.y));
            }
            levels.push_back(std::move(cur));
            cur = std::move(next);
        }
    }
    int rc = GRT_OK;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    double* d_boxes = nullptr;
    uint32_t *d_order = nullptr, *d_span_of = nullptr, *d_axis = nullptr;
    uint2* d_spans = nullptr;
    unsigned long long* d_sbox = nullptr;
    SortKey* d_keys = nullptr;
    void* d_temp = nullptr;
    size_t temp_bytes = 0, max_spans = 1;
    uint64_t launches = 0;
    for (auto& l : levels) max_spans = l.size() > max_spans ? l.size() : max_spans;
    const unsigned T = 256, B = (n + T - 1) / T;
    BVH_TRY(cudaSetDevice(device));
    BVH_TRY(grt_dev_alloc((void**)&d_boxes, (size_t)n * 48));
    BVH_TRY(grt_dev_alloc((void**)&d_order, (size_t)n * 4));
    BVH_TRY(grt_dev_alloc((void**)&d_span_of, (size_t)n * 4));
    BVH_TRY(grt_dev_alloc((void**)&d_keys, (size_t)n * sizeof(SortKey)));
    BVH_TRY(grt_dev_alloc((void**)&d_spans, max_spans * sizeof(uint2)));
    BVH_TRY(grt_dev_alloc((void**)&d_axis, max_spans * 4));
    BVH_TRY(grt_dev_alloc((void**)&d_sbox, max_spans * 48));
    BVH_TRY(cub::DeviceMergeSort::StableSortPairs(nullptr, temp_bytes, d_keys, d_order, (int64_t)n, SortLess()));
    BVH_TRY(grt_dev_alloc((void**)&d_temp, temp_bytes ? temp_bytes : 16));
    BVH_TRY(cudaMemcpy(d_boxes, boxes, (size_t)n * 48, cudaMemcpyHostToDevice));
    bvh_iota<<<B, T>>>(d_order, n);
    launches++;
    for (auto& l : levels) {
        const uint32_t ns = (uint32_t)l.size();
        BVH_TRY(cudaMemcpyAsync(d_spans, l.data(), (size_t)ns * sizeof(uint2), cudaMemcpyHostToDevice, 0));
        bvh_span_of_pos<<<B, T>>>(d_spans, ns, n, d_span_of);
        bvh_init_boxes<<<(ns * 6 + T - 1) / T, T>>>(d_sbox, ns);
        bvh_span_boxes<<<B, T>>>(d_boxes, d_order, d_span_of, n, d_sbox);
        bvh_span_axis<<<(ns + T - 1) / T, T>>>(d_sbox, ns, d_axis);
        bvh_make_keys<<<B, T>>>(d_boxes, d_order, d_span_of, d_spans, d_axis, n, d_keys);
        BVH_TRY(cub::DeviceMergeSort::StableSortPairs(d_temp, temp_bytes, d_keys, d_order, (int64_t)n, SortLess()));
        launches += 6;
    }
    BVH_TRY(cudaGetLastError());
    BVH_TRY(cudaMemcpy(order_out, d_order, (size_t)n * 4, cudaMemcpyDeviceToHost));
done:
    grt_dev_free(d_boxes); grt_dev_free(d_order); grt_dev_free(d_span_of); grt_dev_free(d_keys); grt_dev_free(d_spans); grt_dev_free(d_axis); grt_dev_free(d_sbox); grt_dev_free(d_temp);
    cudaSetDevice(prev_dev);
    grt_count_launch(launches);
    return rc;
}
