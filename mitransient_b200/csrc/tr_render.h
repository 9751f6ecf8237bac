// Internal interface between the C-ABI wrappers and the render launcher (tr_render.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "tr_scene.h"

namespace tr {
// Render the pixels of `pixel_subset` (device array of global pixel ids; NULL = the pixels owned by
// (tile, rank, world)) into film/steady, which are indexed by GLOBAL pixel id.
// compact: film/steady are indexed by the position in the pixel list instead (sharded host output).
int render_device(TrScene *scene, const TrRenderParams *p, float *film, float *steady, TrCounters *counters,
                  cudaStream_t stream, const uint32_t *pixel_subset, uint32_t n_subset, bool compact = false);
// device list of the pixels owned by (tile_size, rank, world) of `p`, in render order
int owned_pixels(SceneImpl &sc, const TrRenderParams *p, cudaStream_t stream, const uint32_t **d_list, uint32_t *n);
void list_owned_pixels(uint32_t W, uint32_t H, uint32_t tile, uint32_t rank, uint32_t world, std::vector<uint32_t> &px);
int validate_params(const SceneImpl &sc, const TrRenderParams *p);
}  // namespace tr
