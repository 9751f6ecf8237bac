// Internal interface between the C-ABI wrappers and the render launcher (tr_render.cu).
#pragma once
#include <cuda_runtime.h>

#include "tr_scene.h"

namespace tr {
// Render the pixels of `pixel_subset` (device array of global pixel ids; NULL = the pixels owned by
// (tile, rank, world)) into film/steady, which are indexed by GLOBAL pixel id.
int render_device(TrScene *scene, const TrRenderParams *p, float *film, float *steady, TrCounters *counters,
                  cudaStream_t stream, const uint32_t *pixel_subset, uint32_t n_subset);
int validate_params(const SceneImpl &sc, const TrRenderParams *p);
}  // namespace tr
