// grt_internal.h — shared between the .cu translation units of libgrt_cuda.
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "../../include/grt.h"

#ifndef GRT_MEGA_THREADS
#define GRT_MEGA_THREADS 128
#endif
#ifndef GRT_MEGA_MIN_BLOCKS
#define GRT_MEGA_MIN_BLOCKS 6   /* 80 registers, 24 warps/SM: measured best of 4/5/6 (profiles/README.md) */
#endif
// the scene blob is staged into shared memory when every resident block can hold a copy
#define GRT_STAGE_MAX_BYTES (28u * 1024u)
// clamp-stack entries per thread kept in shared memory (16 B each)
#define GRT_RS_SMEM 4
/* BVH scenes (variants with F_NODE): resident blocks per SM, shared-memory traversal stack entries per thread
   (deeper entries overflow to local memory), and the staging limit that still lets all blocks fit */
#ifndef GRT_MEGA_MIN_BLOCKS_BVH
#define GRT_MEGA_MIN_BLOCKS_BVH 6
#endif
#define GRT_TRAV_SMEM 32
#define GRT_STAGE_MAX_BYTES_BVH (12u * 1024u)
/* a traversal slice of the megakernel on BVH scenes ends when fewer than GRT_TRAV_EXIT16/16 of the lanes that
   entered it still have work (env GRT_TRAV_EXIT16 overrides) */
#define GRT_TRAV_EXIT16 4
/* non-zero: never use the resumable, warp-synchronous traversal in the megakernel (A/B builds) */
#ifndef GRT_RESUMABLE_REQ
#define GRT_RESUMABLE_REQ 0u
#endif

namespace grtd { struct DevScene; struct DevCamera; }

void grt_set_error(const std::string& s);
void grt_count_launch(uint64_t n);
const grtd::DevScene* grt_internal_dev_scene(GrtSceneHandle h);
int grt_internal_sm_count(GrtSceneHandle h);
int grt_internal_staged(GrtSceneHandle h);   // 0 none, 1 hot arrays, 2 whole blob
unsigned int* grt_internal_counter(GrtSceneHandle h);
/* Device memory from the device's stream-ordered pool, kept by the process instead of being handed back to the driver
   on free: on the B200 boxes a cudaFree costs 30-500 ms once a process holds a few GB, which made scene upload + free
   the largest item of an end-to-end step after the kernel itself.  grt_dev_free waits for the device first (what
   cudaFree does implicitly). */
cudaError_t grt_dev_alloc(void** p, size_t bytes);
void grt_dev_free(void* p);
/* the wavefront variant's path pool (grown on demand, freed with the scene) and 64 pinned bytes for its counters */
void* grt_internal_wf_pool(GrtSceneHandle h, size_t bytes, void** pinned64);
void grt_internal_set_timing(const GrtTiming& t);
int grt_internal_resolve_variant(GrtSceneHandle h, const GrtOptions* opt, bool has_stats_buffer);   /* what GRT_VARIANT_AUTO picks */
int grt_internal_peer_pull(float* d_dst, const float* d_src, uint64_t n, cudaStream_t st);
int grt_make_dev_camera(const GrtCamera* c, grtd::DevCamera* out);
