// grt_wavefront.cu — wavefront variant (placeholder until implemented).
#include "dev_shade.cuh"
#include "grt_internal.h"
int grt_render_wavefront(GrtSceneHandle h, const GrtCamera* cam, const GrtOptions* opt, float* d_rgb_sum, cudaStream_t st, GrtStats* d_stats) {
    grt_set_error("wavefront variant not built yet");
    return GRT_E_UNSUPPORTED;
}
