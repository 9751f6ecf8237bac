// tr_render_to_host: the reference-facing call with HOST output buffers (what mi.render returns on
// a CPU variant, mitransient/integrators/common.py:212-213).  The film is rendered in bands of
// rows; the device->host copy of band i runs on a second stream while band i+1 renders, so the
// PCIe transfer of the (large) film overlaps the path tracing.
#include <algorithm>
#include <vector>

#include "tr_common.h"
#include "tr_render.h"

extern "C" int tr_render_to_host(const TrScene *scene_c, const TrRenderParams *params, float *film_host, float *steady_host,
                                 TrCounters *counters_host, void *stream_v) {
    using namespace tr;
    if (!scene_c) return fail(TR_ERR_INVALID_ARGUMENT, "scene is NULL");
    if (!film_host) return fail(TR_ERR_INVALID_ARGUMENT, "film_host is NULL");
    TrScene *scene = const_cast<TrScene *>(scene_c);
    SceneImpl &sc = scene->impl;
    int rc = validate_params(sc, params);
    if (rc != TR_OK) return rc;
    if (params->world != 1) return fail(TR_ERR_UNSUPPORTED, "tr_render_to_host renders the whole film (world must be 1)");
    if (params->accumulate) return fail(TR_ERR_UNSUPPORTED, "tr_render_to_host does not accumulate");
    cudaStream_t stream = (cudaStream_t) stream_v;

    const uint32_t W = sc.film.width, H = sc.film.height, T = sc.film.temporal_bins;
    const int C = params->film_channels;
    const size_t row_floats = (size_t) W * T * C;
    // band height: ~192 MiB of film per band, at least 8 pixels per SM-resident CTA chunk
    uint32_t rows = (uint32_t) std::max<size_t>(1, ((size_t) 192 << 20) / (row_floats * 4));
    rows = std::min(rows, H);
    const uint32_t n_bands = (H + rows - 1) / rows;

    float *d_film[2] = {nullptr, nullptr}, *d_steady = nullptr;
    uint32_t *d_pix = nullptr; TrCounters *d_cnt = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    int status = TR_OK;
    auto cleanup = [&]() {
        for (int i = 0; i < 2; ++i) {
            if (d_film[i]) cudaFree(d_film[i]);
            if (rendered[i]) cudaEventDestroy(rendered[i]);
            if (copied[i]) cudaEventDestroy(copied[i]);
        }
        if (d_steady) cudaFree(d_steady);
        if (d_pix) cudaFree(d_pix);
        if (d_cnt) cudaFree(d_cnt);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    };
#define TR_TRY(expr)                                                                          \
    do { cudaError_t _e = (expr);                                                             \
         if (_e != cudaSuccess) { status = fail_cuda(_e, #expr, __FILE__, __LINE__); cleanup(); return status; } } while (0)

    const size_t band_bytes = (size_t) rows * row_floats * 4;
    for (int i = 0; i < (n_bands > 1 ? 2 : 1); ++i) TR_TRY(cudaMalloc(&d_film[i], band_bytes));
    TR_TRY(cudaMalloc(&d_steady, (size_t) W * H * 3 * 4));
    TR_TRY(cudaMalloc(&d_cnt, sizeof(TrCounters)));
    TR_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(TrCounters), stream));
    {   // pixel ids in row-major order: band b covers ids [b*rows*W, ...)
        std::vector<uint32_t> ids((size_t) W * H);
        for (size_t i = 0; i < ids.size(); ++i) ids[i] = (uint32_t) i;
        TR_TRY(cudaMalloc(&d_pix, ids.size() * 4));
        TR_TRY(cudaMemcpyAsync(d_pix, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, stream));
        TR_TRY(cudaStreamSynchronize(stream));
    }
    TR_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        TR_TRY(cudaEventCreateWithFlags(&rendered[i], cudaEventDisableTiming));
        TR_TRY(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    }

    for (uint32_t b = 0; b < n_bands; ++b) {
        const int buf = (int) (b & 1u);
        const uint32_t y0 = b * rows, nrows = std::min(rows, H - y0);
        const size_t first_pix = (size_t) y0 * W, npix = (size_t) nrows * W;
        if (b >= 2) TR_TRY(cudaStreamWaitEvent(stream, copied[buf], 0));   // band b-2 has left the buffer
        // the kernel indexes the film by global pixel id: bias the band buffer accordingly
        float *film_biased = d_film[buf] - first_pix * (size_t) T * C;
        rc = render_device(scene, params, film_biased, steady_host ? d_steady : nullptr, counters_host ? d_cnt : nullptr,
                           stream, d_pix + first_pix, (uint32_t) npix);
        if (rc != TR_OK) { cleanup(); return rc; }
        TR_TRY(cudaEventRecord(rendered[buf], stream));
        TR_TRY(cudaStreamWaitEvent(copy_stream, rendered[buf], 0));
        TR_TRY(cudaMemcpyAsync(film_host + first_pix * (size_t) T * C, d_film[buf], npix * (size_t) T * C * 4,
                               cudaMemcpyDeviceToHost, copy_stream));
        TR_TRY(cudaEventRecord(copied[buf], copy_stream));
    }
    if (steady_host) TR_TRY(cudaMemcpyAsync(steady_host, d_steady, (size_t) W * H * 3 * 4, cudaMemcpyDeviceToHost, stream));
    if (counters_host) TR_TRY(cudaMemcpyAsync(counters_host, d_cnt, sizeof(TrCounters), cudaMemcpyDeviceToHost, stream));
    TR_TRY(cudaStreamSynchronize(stream));
    TR_TRY(cudaStreamSynchronize(copy_stream));
#undef TR_TRY
    cleanup();
    return TR_OK;
}
