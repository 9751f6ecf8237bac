// tr_render_to_host: the reference-facing call with HOST output buffers (what mi.render returns on
// a CPU variant, mitransient/integrators/common.py:212-213).  The film is rendered in bands of
// rows; the device->host copy of band i runs on a second stream while band i+1 renders, so the
// PCIe transfer of the (large) film overlaps the path tracing.  Band buffers, streams and events
// live in a per-device workspace that is created on first use and reused by later calls.
#include <algorithm>
#include <mutex>
#include <vector>

#include "tr_common.h"
#include "tr_render.h"

namespace tr {
namespace {

constexpr int kBandBuffers = 3;

struct HostWorkspace {
    int device = -1;
    float *d_film[kBandBuffers] = {nullptr, nullptr, nullptr};
    size_t band_bytes = 0;
    float *d_steady = nullptr; size_t steady_bytes = 0;
    uint32_t *d_pix = nullptr; size_t n_pix = 0;
    TrCounters *d_cnt = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t rendered[kBandBuffers] = {nullptr, nullptr, nullptr}, copied[kBandBuffers] = {nullptr, nullptr, nullptr};
};

std::mutex g_ws_mutex;
std::vector<HostWorkspace *> g_ws;   // one per device, never freed (process lifetime)

int get_workspace(size_t band_bytes, size_t steady_bytes, size_t n_pix, cudaStream_t stream, HostWorkspace **out) {
    int dev = 0;
    TR_CUDA(cudaGetDevice(&dev));
    HostWorkspace *ws = nullptr;
    for (auto *w : g_ws) if (w->device == dev) ws = w;
    if (!ws) {
        ws = new HostWorkspace(); ws->device = dev;
        TR_CUDA(cudaStreamCreateWithFlags(&ws->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < kBandBuffers; ++i) {
            TR_CUDA(cudaEventCreateWithFlags(&ws->rendered[i], cudaEventDisableTiming));
            TR_CUDA(cudaEventCreateWithFlags(&ws->copied[i], cudaEventDisableTiming));
        }
        TR_CUDA(cudaMalloc(&ws->d_cnt, sizeof(TrCounters)));
        g_ws.push_back(ws);
    }
    if (ws->band_bytes < band_bytes) {
        for (int i = 0; i < kBandBuffers; ++i) {
            if (ws->d_film[i]) { cudaFree(ws->d_film[i]); ws->d_film[i] = nullptr; }
            TR_CUDA(cudaMalloc(&ws->d_film[i], band_bytes));
        }
        ws->band_bytes = band_bytes;
    }
    if (ws->steady_bytes < steady_bytes) {
        if (ws->d_steady) { cudaFree(ws->d_steady); ws->d_steady = nullptr; }
        TR_CUDA(cudaMalloc(&ws->d_steady, steady_bytes));
        ws->steady_bytes = steady_bytes;
    }
    if (ws->n_pix < n_pix) {   // identity pixel-id table: band b renders ids [b*rows*W, ...)
        if (ws->d_pix) { cudaFree(ws->d_pix); ws->d_pix = nullptr; }
        std::vector<uint32_t> ids(n_pix);
        for (size_t i = 0; i < n_pix; ++i) ids[i] = (uint32_t) i;
        TR_CUDA(cudaMalloc(&ws->d_pix, n_pix * 4));
        TR_CUDA(cudaMemcpyAsync(ws->d_pix, ids.data(), n_pix * 4, cudaMemcpyHostToDevice, stream));
        TR_CUDA(cudaStreamSynchronize(stream));
        ws->n_pix = n_pix;
    }
    *out = ws;
    return TR_OK;
}

}  // namespace
}  // namespace tr

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

    if (sc.film.laser_scan_width) return fail(TR_ERR_UNSUPPORTED, "tr_render_to_host does not support exhaustive_scan films");
    const uint32_t W = sc.film.width, H = sc.film.height;
    // floats per pixel: T * C (time histogram) or 2F + 1 (phasor film, film_channels = 2F + 1)
    const uint32_t T = sc.film.type == TR_FILM_PHASOR_HDR ? 1u : sc.film.temporal_bins;
    const int C = params->film_channels;
    const size_t row_floats = (size_t) W * T * C;
    // band height: ~96 MiB of film per band (the first copy starts early, the last one ends soon after the last kernel)
    uint32_t rows = (uint32_t) std::max<size_t>(1, ((size_t) 96 << 20) / (row_floats * 4));
    rows = std::min(rows, H);
    const uint32_t n_bands = (H + rows - 1) / rows;
    const size_t band_bytes = (size_t) rows * row_floats * 4;

    std::lock_guard<std::mutex> guard(g_ws_mutex);
    HostWorkspace *ws = nullptr;
    rc = get_workspace(band_bytes, (size_t) W * H * 3 * 4, (size_t) W * H, stream, &ws);
    if (rc != TR_OK) return rc;
    if (counters_host) TR_CUDA(cudaMemsetAsync(ws->d_cnt, 0, sizeof(TrCounters), stream));

    for (uint32_t b = 0; b < n_bands; ++b) {
        const int buf = (int) (b % kBandBuffers);
        const uint32_t y0 = b * rows, nrows = std::min(rows, H - y0);
        const size_t first_pix = (size_t) y0 * W, npix = (size_t) nrows * W;
        if (b >= (uint32_t) kBandBuffers) TR_CUDA(cudaStreamWaitEvent(stream, ws->copied[buf], 0));   // buffer is free again
        // the kernel indexes the film by global pixel id: bias the band buffer accordingly
        float *film_biased = ws->d_film[buf] - first_pix * (size_t) T * C;
        rc = render_device(scene, params, film_biased, steady_host ? ws->d_steady : nullptr, counters_host ? ws->d_cnt : nullptr,
                           stream, ws->d_pix + first_pix, (uint32_t) npix);
        if (rc != TR_OK) { cudaStreamSynchronize(stream); cudaStreamSynchronize(ws->copy_stream); return rc; }
        TR_CUDA(cudaEventRecord(ws->rendered[buf], stream));
        TR_CUDA(cudaStreamWaitEvent(ws->copy_stream, ws->rendered[buf], 0));
        TR_CUDA(cudaMemcpyAsync(film_host + first_pix * (size_t) T * C, ws->d_film[buf], npix * (size_t) T * C * 4,
                                cudaMemcpyDeviceToHost, ws->copy_stream));
        TR_CUDA(cudaEventRecord(ws->copied[buf], ws->copy_stream));
    }
    if (steady_host) TR_CUDA(cudaMemcpyAsync(steady_host, ws->d_steady, (size_t) W * H * 3 * 4, cudaMemcpyDeviceToHost, stream));
    if (counters_host) TR_CUDA(cudaMemcpyAsync(counters_host, ws->d_cnt, sizeof(TrCounters), cudaMemcpyDeviceToHost, stream));
    TR_CUDA(cudaStreamSynchronize(stream));
    TR_CUDA(cudaStreamSynchronize(ws->copy_stream));
    return TR_OK;
}
