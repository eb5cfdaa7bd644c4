// grt_multi.cu — in-process multi-GPU render (the `-gpus N` path of the CLI and
// of a cgo caller, which is ONE process).  Each device holds a scene replica
// and renders the strata s ≡ g (mod N).
//
// Fused path (NVLink / NVSwitch with native peer atomics): ONE accumulation buffer lives on devices[0], every other
// device maps it as peer memory, and the render kernels add their per-pixel sums straight into it with system-scope
// atomics as each pixel is retired — the exchange is spread over the whole render (12.6 MB of 4-byte reductions over
// tens of milliseconds) instead of following it, and no collective runs at all.  That holds for the megakernel, which
// retires one sum per PIXEL.  The wavefront kernels retire one sum per PATH (paths of a pixel are spread over the pool),
// and a remote atomic per path is 10^8 NVLink transactions per frame: there every device accumulates at home (plain
// device-scope atomics), and devices[0] PULLS each peer's finished buffer over NVLink with coalesced 128-bit peer loads
// and adds it to its own (peer_pull_kernel, ~0.2 ms for a 100 MB frame; the pool memory of the peers is opened to
// devices[0] with cudaMemPoolSetAccess).
// Fallback (no peer access): private buffers and ONE ncclReduce(sum, root = devices[0]) before tonemap.
// The devices render concurrently (one host thread each); kernel_ms is the slowest device's e0..e1 time.
// (bench.py instead runs one process per GPU and reduces through torch.distributed's NCCL communicator.)
//
// NCCL is bound at run time with dlopen so that libgrt_cuda has no link-time
// dependency on a particular libnccl (PyTorch bundles its own).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include "grt_internal.h"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0 };
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        Reduce = (decltype(Reduce))dlsym(lib, "ncclReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Reduce;
    }
};
Nccl g_nccl;
}  // namespace

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) { grt_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); rc = GRT_E_CUDA; goto done; } \
    } while (0)
#define NC(call)                                                                                 \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if (r_ != 0) { grt_set_error(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); rc = GRT_E_NCCL; goto done; } \
    } while (0)

// One plain-cudaMalloc buffer per device, kept for the life of the process (cudaFree costs 30-500 ms once a process
// holds a few GB); a second concurrent caller gets a fresh allocation of its own.
namespace {
struct PeerBuf { float* p = nullptr; size_t bytes = 0; bool busy = false; };
std::mutex g_peer_m;
PeerBuf g_peer_buf[64];
float* peer_buf_acquire(int dev, size_t bytes, int* own) {
    *own = 0;
    if (cudaSetDevice(dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_peer_m);
    if (dev >= 0 && dev < 64 && !g_peer_buf[dev].busy) {
        PeerBuf& b = g_peer_buf[dev];
        if (b.bytes < bytes) {
            if (b.p) cudaFree(b.p);
            b.p = nullptr; b.bytes = 0;
            if (cudaMalloc((void**)&b.p, bytes) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return nullptr; }
            b.bytes = bytes;
        }
        b.busy = true; *own = 1;
        return b.p;
    }
    float* p = nullptr;
    if (cudaMalloc((void**)&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *own = 2;
    return p;
}
void peer_buf_release(int dev, float* p, int own) {
    if (own == 2) { cudaSetDevice(dev); cudaFree(p); return; }
    std::lock_guard<std::mutex> lk(g_peer_m);
    if (dev >= 0 && dev < 64 && g_peer_buf[dev].p == p) g_peer_buf[dev].busy = false;
}
}   // namespace

extern "C" int grt_render_multi(const GrtScene* scene, const GrtCamera* cam, const GrtOptions* opt, const int* devices, int n,
                                float* rgb_sum, uint8_t* rgb8, double* kernel_ms) {
    if (!scene || !cam || !opt || !devices || n < 1 || !rgb_sum) { grt_set_error("grt_render_multi: bad argument"); return GRT_E_INVALID; }
    int have = grt_device_count();
    if (have <= 0) { grt_set_error("no CUDA device available (libgrt_cuda has no CPU fallback)"); return GRT_E_NO_DEVICE; }
    for (int i = 0; i < n; i++) if (devices[i] < 0 || devices[i] >= have) { grt_set_error("device ordinal out of range"); return GRT_E_NO_DEVICE; }
    // can every other device reach devices[0]'s memory with native atomics?
    bool fused = n > 1;
    if (const char* e = getenv("GRT_MULTI_P2P")) fused = fused && atoi(e) != 0;
    for (int g = 1; g < n && fused; g++) {
        int can = 0, at = 0;
        if (devices[g] == devices[0]) { fused = false; break; }
        if (cudaDeviceCanAccessPeer(&can, devices[g], devices[0]) != cudaSuccess || !can) fused = false;
        else if (cudaDeviceGetP2PAttribute(&at, cudaDevP2PAttrNativeAtomicSupported, devices[g], devices[0]) != cudaSuccess || !at) fused = false;
        else if (cudaDeviceCanAccessPeer(&can, devices[0], devices[g]) != cudaSuccess || !can) fused = false;   // (the pull direction)
    }
    if (n > 1 && !fused && !g_nccl.load()) { grt_set_error("libnccl.so.2 could not be loaded"); return GRT_E_NCCL; }

    int rc = GRT_OK;
    size_t nval = (size_t)cam->width * cam->height * 3;
    std::vector<GrtSceneHandle> hs(n, nullptr);
    std::vector<float*> d_sum(n, nullptr);
    std::vector<cudaStream_t> st(n, nullptr);
    std::vector<cudaEvent_t> e0(n, nullptr), e1(n, nullptr);
    std::vector<ncclComm_t> comms(n, nullptr);
    uint8_t* d_rgb8 = nullptr;
    float* d_shared = nullptr;      // fused path: the one accumulation buffer, plain cudaMalloc (peer-mappable) on devices[0]
    cudaEvent_t e_init = nullptr;
    float* d_total = nullptr;       // where the complete sums end up (on devices[0])
    bool wf = false;                // fused path, wavefront shards: private sums, pulled by devices[0] at the end
    std::vector<int> own_sum(n, 0); // d_sum[g] came from peer_buf_acquire: 1 = the kept buffer of that device, 2 = a fresh cudaMalloc

    for (int g = 0; g < n; g++) {
        rc = grt_scene_upload(scene, devices[g], &hs[g]);
        if (rc) goto done;
        CU(cudaSetDevice(devices[g]));
        CU(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
        CU(cudaEventCreate(&e0[g]));
        CU(cudaEventCreate(&e1[g]));
        if (fused) {
            if (g == 0) wf = grt_internal_resolve_variant(hs[0], opt, false) == GRT_VARIANT_WAVEFRONT;
            if (g > 0 && wf) {
                // private sums, pulled into the shared buffer at the end: plain cudaMalloc memory (peer-mappable; opening
                // the stream-ordered pool to devices[0] with cudaMemPoolSetAccess made the NEXT big pool allocation of
                // this device fail with "out of memory" on these boxes), kept per device across calls
                d_sum[g] = peer_buf_acquire(devices[g], nval * sizeof(float), &own_sum[g]);
                if (!d_sum[g]) { grt_set_error("grt_render_multi: cannot allocate the private accumulation buffer"); rc = GRT_E_CUDA; goto done; }
                CU(cudaMemsetAsync(d_sum[g], 0, nval * sizeof(float), st[g]));
                CU(cudaSetDevice(devices[0]));
                cudaError_t pe = cudaDeviceEnablePeerAccess(devices[g], 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { grt_set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)); rc = GRT_E_CUDA; goto done; }
                cudaGetLastError();
                CU(cudaSetDevice(devices[g]));
            }
            if (g == 0) {
                CU(cudaMalloc((void**)&d_shared, nval * sizeof(float)));
                CU(cudaMemcpyAsync(d_shared, rgb_sum, nval * sizeof(float), cudaMemcpyHostToDevice, st[0]));
                CU(cudaEventCreateWithFlags(&e_init, cudaEventDisableTiming));
                CU(cudaEventRecord(e_init, st[0]));
            } else {
                cudaError_t pe = cudaDeviceEnablePeerAccess(devices[0], 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { grt_set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)); rc = GRT_E_CUDA; goto done; }
                cudaGetLastError();
                CU(cudaStreamWaitEvent(st[g], e_init, 0));   // the shared buffer is initialised before anyone adds to it
            }
            continue;
        }
        CU(grt_dev_alloc((void**)&d_sum[g], nval * sizeof(float)));
        if (g == 0) CU(cudaMemcpyAsync(d_sum[g], rgb_sum, nval * sizeof(float), cudaMemcpyHostToDevice, st[g]));
        else CU(cudaMemsetAsync(d_sum[g], 0, nval * sizeof(float), st[g]));
    }
    if (n > 1 && !fused) {
        // communicators are kept for the life of the process (one set per device list): creating them costs ~0.3 s and the
        // first collective on a fresh communicator another ~0.4 s of connection setup
        static std::mutex cm;
        static std::vector<int> kept_devs;
        static std::vector<ncclComm_t> kept;
        std::lock_guard<std::mutex> lk(cm);
        if (kept_devs != std::vector<int>(devices, devices + n)) {
            for (ncclComm_t c : kept) if (c) g_nccl.CommDestroy(c);
            kept.assign(n, nullptr);
            kept_devs.clear();
            NC(g_nccl.CommInitAll(kept.data(), n, devices));
            kept_devs.assign(devices, devices + n);
        }
        comms = kept;
    }

    // every device renders its strata shard, concurrently: one host thread per device, because the wavefront variant's
    // bounce loop blocks its caller (it reads the live-path counter back between graph launches) — issued from one
    // thread the devices would render one after another
    {
        std::vector<int> rcs(n, GRT_OK);
        std::vector<std::string> errs(n);
        auto work = [&](int g) {
            GrtOptions o = *opt;
            uint32_t base_stride = opt->sample_stride ? opt->sample_stride : 1u;
            o.sample_first = opt->sample_first + (uint32_t)g * base_stride;
            o.sample_stride = base_stride * (uint32_t)n;
            o.device = devices[g];
            const bool direct = fused && !d_sum[g];   // this shard adds straight into the shared buffer as it goes
            if (direct && !wf) o.flags |= GRT_OPT_ATOMIC_SUM;   // (wavefront: devices[0] is alone on the shared buffer until the pull)
            cudaError_t e = cudaSetDevice(devices[g]);
            if (e == cudaSuccess) e = cudaEventRecord(e0[g], st[g]);
            if (e != cudaSuccess) { rcs[g] = GRT_E_CUDA; errs[g] = cudaGetErrorString(e); return; }
            rcs[g] = grt_render_device(hs[g], cam, &o, direct ? d_shared : d_sum[g], st[g], nullptr);
            if (rcs[g]) errs[g] = grt_last_error();   // the error text is thread-local: carry it to the caller
        };
        if (n == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int g = 0; g < n; g++) th.emplace_back(work, g);
            for (auto& t : th) t.join();
        }
        for (int g = 0; g < n; g++) if (rcs[g]) { grt_set_error(errs[g]); rc = rcs[g]; goto done; }
    }
    if (n > 1 && !fused) {
        NC(g_nccl.GroupStart());
        for (int g = 0; g < n; g++) {
            ncclResult_t r_ = g_nccl.Reduce(d_sum[g], d_sum[g], nval, ncclFloat32, ncclSum, 0, comms[g], st[g]);
            if (r_ != 0) { g_nccl.GroupEnd(); grt_set_error("ncclReduce failed"); rc = GRT_E_NCCL; goto done; }
        }
        NC(g_nccl.GroupEnd());
    }
    for (int g = 1; g < n; g++) {
        CU(cudaSetDevice(devices[g]));
        CU(cudaEventRecord(e1[g], st[g]));
    }
    CU(cudaSetDevice(devices[0]));
    if (fused) for (int g = 1; g < n; g++) {
        CU(cudaStreamWaitEvent(st[0], e1[g], 0));   // every shard has landed (or is complete at home) before tonemap / read-back
        if (wf && (rc = grt_internal_peer_pull(d_shared, d_sum[g], nval, st[0]))) goto done;
    }
    CU(cudaEventRecord(e1[0], st[0]));
    d_total = fused ? d_shared : d_sum[0];
    if (rgb8) {
        CU(grt_dev_alloc((void**)&d_rgb8, nval));
        float scale = 1.0f / (float)((double)cam->spp_sqrt * (double)cam->spp_sqrt);
        rc = grt_tonemap_device(d_total, d_rgb8, nval, scale, st[0]);
        if (rc) goto done;
    }
    {
        double worst = 0;
        for (int g = 0; g < n; g++) {
            CU(cudaSetDevice(devices[g]));
            CU(cudaStreamSynchronize(st[g]));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0[g], e1[g]));
            if (ms > worst) worst = ms;
        }
        if (kernel_ms) *kernel_ms = worst;
    }
    CU(cudaSetDevice(devices[0]));
    CU(cudaMemcpy(rgb_sum, d_total, nval * sizeof(float), cudaMemcpyDeviceToHost));
    if (rgb8) CU(cudaMemcpy(rgb8, d_rgb8, nval, cudaMemcpyDeviceToHost));

done:
    for (int g = 0; g < n; g++) {
        if (hs[g]) cudaSetDevice(devices[g]);
        if (d_sum[g]) { if (own_sum[g]) peer_buf_release(devices[g], d_sum[g], own_sum[g]); else grt_dev_free(d_sum[g]); }
        if (e0[g]) cudaEventDestroy(e0[g]);
        if (e1[g]) cudaEventDestroy(e1[g]);
        if (st[g]) cudaStreamDestroy(st[g]);
        if (hs[g]) grt_scene_free(hs[g]);
    }
    if (d_rgb8) { cudaSetDevice(devices[0]); grt_dev_free(d_rgb8); }
    if (e_init) cudaEventDestroy(e_init);
    if (d_shared) { cudaSetDevice(devices[0]); cudaFree(d_shared); }
    return rc;
}
