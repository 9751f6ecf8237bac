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
    if (params->accumulate) return fail(TR_ERR_UNSUPPORTED, "tr_render_to_host does not accumulate");
    cudaStream_t stream = (cudaStream_t) stream_v;

    if (sc.film.laser_scan_width) return fail(TR_ERR_UNSUPPORTED, "tr_render_to_host does not support exhaustive_scan films");
    const uint32_t W = sc.film.width, H = sc.film.height;
    // floats per pixel: T * C (time histogram) or 2F + 1 (phasor film, film_channels = 2F + 1)
    const uint32_t T = sc.film.type == TR_FILM_PHASOR_HDR ? 1u : sc.film.temporal_bins;
    const int C = params->film_channels;
    const size_t pixel_floats = (size_t) T * C;
    const bool sharded = params->world > 1;
    // world == 1: bands of whole rows, the host film is indexed by pixel id.  world > 1: bands of consecutive OWNED
    // pixels (tr_owned_pixels order), the host film is the compact float[n_owned][T][C] array of this rank.
    const uint32_t *d_list = nullptr; uint32_t n_list = W * H;
    if (sharded) {
        rc = owned_pixels(sc, params, stream, &d_list, &n_list);
        if (rc != TR_OK) return rc;
        if (n_list == 0) return TR_OK;
    }
    // band size: ~96 MiB of film per band (the first copy starts early, the last one ends soon after the last kernel)
    uint32_t band_pix = (uint32_t) std::max<size_t>(1, ((size_t) 96 << 20) / (pixel_floats * 4));
    if (!sharded) band_pix = std::max(1u, band_pix / W) * W;          // whole rows
    band_pix = std::min(band_pix, n_list);
    const uint32_t n_bands = (n_list + band_pix - 1) / band_pix;
    const size_t band_bytes = (size_t) band_pix * pixel_floats * 4;

    std::lock_guard<std::mutex> guard(g_ws_mutex);
    HostWorkspace *ws = nullptr;
    rc = get_workspace(band_bytes, (size_t) W * H * 3 * 4, (size_t) W * H, stream, &ws);
    if (rc != TR_OK) return rc;
    if (!sharded) d_list = ws->d_pix;
    if (counters_host) TR_CUDA(cudaMemsetAsync(ws->d_cnt, 0, sizeof(TrCounters), stream));

    for (uint32_t b = 0; b < n_bands; ++b) {
        const int buf = (int) (b % kBandBuffers);
        const size_t first = (size_t) b * band_pix, npix = std::min<size_t>(band_pix, n_list - first);
        if (b >= (uint32_t) kBandBuffers) TR_CUDA(cudaStreamWaitEvent(stream, ws->copied[buf], 0));   // buffer is free again
        // whole-film mode: the kernel indexes the film by global pixel id (= position in the identity list): bias the
        // band buffer.  Sharded mode: the kernel writes pixel j of the band at position j (compact).
        float *film_dst = sharded ? ws->d_film[buf] : ws->d_film[buf] - first * pixel_floats;
        float *steady_dst = !steady_host ? nullptr : (sharded ? ws->d_steady + first * 3 : ws->d_steady);
        rc = render_device(scene, params, film_dst, steady_dst, counters_host ? ws->d_cnt : nullptr, stream, d_list + first,
                           (uint32_t) npix, sharded);
        if (rc != TR_OK) { cudaStreamSynchronize(stream); cudaStreamSynchronize(ws->copy_stream); return rc; }
        TR_CUDA(cudaEventRecord(ws->rendered[buf], stream));
        TR_CUDA(cudaStreamWaitEvent(ws->copy_stream, ws->rendered[buf], 0));
        TR_CUDA(cudaMemcpyAsync(film_host + first * pixel_floats, ws->d_film[buf], npix * pixel_floats * 4,
                                cudaMemcpyDeviceToHost, ws->copy_stream));
        TR_CUDA(cudaEventRecord(ws->copied[buf], ws->copy_stream));
    }
    if (steady_host) TR_CUDA(cudaMemcpyAsync(steady_host, ws->d_steady, (size_t) n_list * 3 * 4, cudaMemcpyDeviceToHost, stream));
    if (counters_host) TR_CUDA(cudaMemcpyAsync(counters_host, ws->d_cnt, sizeof(TrCounters), cudaMemcpyDeviceToHost, stream));
    TR_CUDA(cudaStreamSynchronize(stream));
    TR_CUDA(cudaStreamSynchronize(ws->copy_stream));
    return TR_OK;
}

// Host-side helper (no GPU): the pixels (y * W + x) owned by `rank`, in the order tr_render_to_host lays them out.
extern "C" int tr_owned_pixels(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t rank, uint32_t world,
                               uint32_t *out, uint64_t *n_out) {
    using namespace tr;
    if (width == 0 || height == 0 || world == 0 || rank >= world || !n_out)
        return fail(TR_ERR_INVALID_ARGUMENT, "tr_owned_pixels: bad argument");
    std::vector<uint32_t> px;
    list_owned_pixels(width, height, tile_size ? tile_size : 32u, rank, world, px);
    if (out) {
        if (*n_out < px.size()) return fail(TR_ERR_INVALID_ARGUMENT, "tr_owned_pixels: output array too small");
        std::copy(px.begin(), px.end(), out);
    }
    *n_out = px.size();
    return TR_OK;
}
