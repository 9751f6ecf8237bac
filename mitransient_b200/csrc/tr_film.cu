// Film-side kernels that are callable on their own through the C ABI:
//   tr_film_scatter_add  = TransientHDRFilm.add_transient_data + TransientImageBlock.put_/accum for a
//                          stream of deposits (transient_hdr_film.py:263-276, transient_image_block.py:130-149)
//   tr_film_develop      = TransientHDRFilm.develop_transient_ (transient_hdr_film.py:220-248)
#include <algorithm>
#include <cstdint>

#include "tr_common.h"

namespace tr {

// One deposit per thread.  Deposits that land in the same (pixel, bin) within a warp are summed in
// registers first (match.any + shuffle reduction) so that one vector RED per distinct bin reaches L2.
__global__ void __launch_bounds__(256) film_scatter_add_kernel(float *__restrict__ film, const TrDeposit *__restrict__ dep,
                                                               unsigned long long n, uint32_t W, uint32_t H, uint32_t T,
                                                               float start_opl, float bin_width, int C, float crop_x, float crop_y) {
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    const float Tf = (float) T;
    // loop bound is warp-uniform so that the warp-level primitives below are convergent
    const unsigned long long n_round = (n + 31ull) & ~31ull;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool ok = i < n;
        float r = 0.f, g = 0.f, b = 0.f;
        unsigned long long key = ~0ull - lane;  // distinct sentinel per lane: never merges
        if (ok) {
            TrDeposit d = dep[i];
            r = d.r; g = d.g; b = d.b;
            float pos = (d.opl - start_opl) / bin_width;              // transient_hdr_film.py:263
            ok = (pos >= 0.0f) && (pos < Tf) && !(r == 0.0f && g == 0.0f && b == 0.0f);   // :265
            if (ok) {
                float fx = floorf(d.x) - crop_x, fy = floorf(d.y) - crop_y;   // floor(pos) - offset  transient_image_block.py:131
                ok = fx >= 0.0f && fy >= 0.0f && fx < (float) W && fy < (float) H;       // :146
                if (ok) {
                    uint32_t x = (uint32_t) fx, y = (uint32_t) fy, z = (uint32_t) floorf(pos);
                    if (z >= T) ok = false;
                    else key = ((unsigned long long) y * W + x) * T + z;   // :142-144 (before * channel_count)
                }
            }
        }
        uint32_t grp = __match_any_sync(0xffffffffu, key);
        bool leader = ok && ((uint32_t) (__ffs(grp) - 1) == lane);
        uint32_t multi = __ballot_sync(0xffffffffu, leader && (grp & (grp - 1)));
        while (multi) {  // rare: several lanes hit one bin -> sum them into the group leader
            int Lr = __ffs(multi) - 1; multi &= multi - 1;
            uint32_t gm = __shfl_sync(0xffffffffu, grp, Lr);
            bool in = (gm >> lane) & 1u;
            float sr = in ? r : 0.f, sg = in ? g : 0.f, sb = in ? b : 0.f;
            for (int o = 16; o > 0; o >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o); sg += __shfl_xor_sync(0xffffffffu, sg, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
            }
            if ((int) lane == Lr) { r = sr; g = sg; b = sb; }
        }
        if (leader) {
            float *p = film + key * (unsigned long long) C;
            if (C == 4) {
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r), "f"(g), "f"(b), "f"(0.0f) : "memory");
            } else {
                atomicAdd(p, r); atomicAdd(p + 1, g); atomicAdd(p + 2, b);
            }
        }
    }
}

// raw [n][4] -> developed [n][3], four bins per thread: 4 x 128-bit loads, 3 x 128-bit stores.
__global__ void __launch_bounds__(256) film_develop4_kernel(const float4 *__restrict__ raw, float4 *__restrict__ out,
                                                            unsigned long long n_quads) {
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += stride) {
        float4 a = __ldcs(raw + 4 * i), b = __ldcs(raw + 4 * i + 1), c = __ldcs(raw + 4 * i + 2), d = __ldcs(raw + 4 * i + 3);
        float wa = a.w == 0.0f ? 1.0f : a.w, wb = b.w == 0.0f ? 1.0f : b.w;
        float wc = c.w == 0.0f ? 1.0f : c.w, wd = d.w == 0.0f ? 1.0f : d.w;
        float4 o0 = make_float4(a.x / wa, a.y / wa, a.z / wa, b.x / wb);
        float4 o1 = make_float4(b.y / wb, b.z / wb, c.x / wc, c.y / wc);
        float4 o2 = make_float4(c.z / wc, d.x / wd, d.y / wd, d.z / wd);
        __stcs(out + 3 * i, o0); __stcs(out + 3 * i + 1, o1); __stcs(out + 3 * i + 2, o2);
    }
}

__global__ void __launch_bounds__(256) film_develop_generic_kernel(const float *__restrict__ raw, float *__restrict__ out,
                                                                   unsigned long long n_bins, int C) {
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    const int oc = C - 1;
    for (unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; i < n_bins; i += stride) {
        float w = raw[i * C + C - 1];
        float div = (w == 0.0f) ? 1.0f : w;
        for (int c = 0; c < oc; ++c) out[i * oc + c] = raw[i * C + c] / div;
    }
}

}  // namespace tr

extern "C" {

int tr_film_scatter_add(float *film, const TrFilmDesc *fd, int32_t channels, const TrDeposit *deposits, uint64_t n,
                        void *stream) {
    if (!film || !fd || (!deposits && n)) return tr::fail(TR_ERR_INVALID_ARGUMENT, "NULL argument");
    if (channels != 3 && channels != 4) return tr::fail(TR_ERR_INVALID_ARGUMENT, "channels must be 3 or 4");
    if (fd->width == 0 || fd->height == 0 || fd->temporal_bins == 0 || !(fd->bin_width_opl > 0.0f))
        return tr::fail(TR_ERR_INVALID_ARGUMENT, "bad film description");
    int rc = tr_device_check();
    if (rc != TR_OK) return rc;
    if (n == 0) return TR_OK;
    int dev = 0, n_sm = 0;
    TR_CUDA(cudaGetDevice(&dev));
    TR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    unsigned long long blocks = (n + 255) / 256;
    int grid = (int) std::min<unsigned long long>(blocks, (unsigned long long) n_sm * 8);
    tr::film_scatter_add_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>(film, deposits, n, fd->width, fd->height,
                                                                        fd->temporal_bins, fd->start_opl, fd->bin_width_opl,
                                                                        channels, (float) fd->crop_offset_x, (float) fd->crop_offset_y);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

int tr_film_develop(const float *raw, float *out, uint64_t n_bins, int32_t in_channels, void *stream) {
    if (!raw || !out) return tr::fail(TR_ERR_INVALID_ARGUMENT, "NULL argument");
    if (in_channels < 2) return tr::fail(TR_ERR_INVALID_ARGUMENT, "in_channels must be >= 2");
    int rc = tr_device_check();
    if (rc != TR_OK) return rc;
    if (n_bins == 0) return TR_OK;
    int dev = 0, n_sm = 0;
    TR_CUDA(cudaGetDevice(&dev));
    TR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const bool aligned = (((uintptr_t) raw | (uintptr_t) out) & 15u) == 0;
    if (in_channels == 4 && (n_bins % 4 == 0) && aligned) {
        unsigned long long quads = n_bins / 4;
        int grid = (int) std::min<unsigned long long>((quads + 255) / 256, (unsigned long long) n_sm * 16);
        tr::film_develop4_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>((const float4 *) raw, (float4 *) out, quads);
    } else {
        int grid = (int) std::min<unsigned long long>((n_bins + 255) / 256, (unsigned long long) n_sm * 16);
        tr::film_develop_generic_kernel<<<grid, 256, 0, (cudaStream_t) stream>>>(raw, out, n_bins, in_channels);
    }
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

}  // extern "C"
