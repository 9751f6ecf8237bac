/*
 * mitransient_b200.h -- C ABI of the B200-native transient path-tracing hot path.
 *
 * This is the drop-in boundary for the ONE path this repository accelerates:
 * mitransient's per-bounce transient path-tracing wavefront plus the scatter-add of
 * per-path radiance into the (y, x, time) film histogram.  Every entry point cites the
 * reference interface (file:line under the mitransient source tree) it replaces.
 *
 * Conventions
 *  - plain C, no torch / C++ types; pointers + sizes only.
 *  - every function returns 0 on success or a negative TrStatus; the message of the last
 *    failure on the calling thread is available from tr_last_error().
 *  - "device" pointers are CUDA device pointers on the CURRENT device of the calling
 *    thread; "host" pointers are ordinary host memory.  Output buffers are always
 *    allocated by the caller (PyTorch on the host side); the library only owns the
 *    opaque TrScene handle.
 *  - all work is enqueued on the caller-supplied cudaStream_t (passed as void*).  No hidden
 *    device synchronisation, except where a function is documented as synchronous.
 *  - film layout is the reference's:  film[y][x][t][c]  (row-major, channel innermost;
 *    mitransient/render/transient_image_block.py:63-68,142-144).
 */
#ifndef MITRANSIENT_B200_H
#define MITRANSIENT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TR_ABI_VERSION 14

/* ------------------------------------------------------------------ status codes */
typedef enum TrStatus {
    TR_OK = 0,
    TR_ERR_INVALID_ARGUMENT = -1, /* bad pointer / size / enum value                         */
    TR_ERR_UNSUPPORTED = -2,      /* combination the reference rejects, or not built here    */
    TR_ERR_CUDA = -3,             /* a CUDA runtime call failed (text in tr_last_error)      */
    TR_ERR_NO_DEVICE = -4,        /* no sm_100 device visible: the product has NO CPU path   */
    TR_ERR_OUT_OF_MEMORY = -5
} TrStatus;

/* ------------------------------------------------------------------ scene description
 * A flat restatement of what mi.load_dict()/XML hands to the integrator: geometry as a list
 * of planar primitives (triangles and parallelograms) grouped into shapes, a BSDF table, an
 * emitter table, one sensor and its film (mitransient/utils.py:78-220 is the benchmark
 * scene expressed this way).
 *
 * A primitive is four vertices (p0, p1, p2, p3); e_k = p_k - p0.  TR_PRIM_TRIANGLE covers
 * {p0 + u e1 + v e2 : u, v >= 0, u + v <= 1}; TR_PRIM_QUAD covers the parallelogram {u, v in [0, 1]}
 * whose fourth corner is p1 + p2 - p0 (Mitsuba's analytic `rectangle`, or two coplanar mesh triangles
 * merged by the loader); for both the geometric normal is normalize(cross(e1, e2)) and p3 is ignored.
 * TR_PRIM_BOX is the SOLID parallelepiped {p0 + u e1 + v e2 + w e3 : u, v, w in [0, 1]} with outward
 * face normals (Mitsuba's `cube`: one slab test instead of twelve triangles); det(e1, e2, e3) != 0.   */
enum { TR_PRIM_TRIANGLE = 0, TR_PRIM_QUAD = 1, TR_PRIM_BOX = 2 };

enum { TR_SHAPE_MESH = 0, TR_SHAPE_RECTANGLE = 1 };
enum { TR_SHAPE_FLAG_SENSOR = 1 };       /* the shape carrying the nlos_capture_meter      */

typedef struct TrShapeDesc {
    int32_t type;        /* TR_SHAPE_*: rectangle = Mitsuba `rectangle` (exactly one quad primitive,
                            to_world-based uniform sampling); mesh = obj/ply/cube               */
    int32_t bsdf;        /* index into TrSceneDesc.bsdfs                                    */
    int32_t emitter;     /* index into TrSceneDesc.emitters (area light) or -1              */
    int32_t flags;       /* TR_SHAPE_FLAG_*                                                 */
    int32_t interior_medium; /* index into media or -1 (transient_prbvolpath)              */
    int32_t exterior_medium; /* index into media or -1                                     */
    uint32_t prim_begin; /* first primitive of this shape in prim_verts                     */
    uint32_t prim_count;
    float to_world[12];  /* row-major 3x4 affine; meaningful for rectangles                 */
} TrShapeDesc;

/* BSDFs (the per-bounce bsdf.sample / bsdf.eval_pdf of transientpath.py:207-227).  diffuse, null: every integrator.
 * conductor (Mitsuba `conductor`: perfect mirror weighted by the complex-IOR Fresnel term) and dielectric (Mitsuba
 * `dielectric`: delta reflection / refraction, radiance scaled by eta_ti^2, optical path length scaled by the
 * accumulated relative IOR, transientpath.py:154,233): transient_path only.                                       */
enum { TR_BSDF_DIFFUSE = 0, TR_BSDF_NULL = 1, TR_BSDF_CONDUCTOR = 2, TR_BSDF_DIELECTRIC = 3 };
enum { TR_BSDF_FLAG_TWOSIDED = 1 };

typedef struct TrBsdfDesc {
    int32_t type;        /* TR_BSDF_*                                                       */
    int32_t flags;       /* TR_BSDF_FLAG_*                                                  */
    float reflectance[3];             /* diffuse: reflectance; conductor / dielectric: specular_reflectance */
    float eta[3];                     /* conductor: real part of the IOR; dielectric: eta[0] = int_ior / ext_ior */
    float k[3];                       /* conductor: extinction coefficient                  */
    float specular_transmittance[3];  /* dielectric                                         */
    float _pad[2];
} TrBsdfDesc;

/* TR_EMITTER_ANGULARAREA: mitransient's own `angulararea` emitter (mitransient/emitters/angulararea.py): an area light
 * whose radiance falls off with the angle to the surface normal (1 inside beam_width, linear down to 0 at
 * cutoff_angle); its sample_direction carries the reference's extra 1/dist^2 factor (:121-123).  transient_path only. */
enum { TR_EMITTER_AREA = 0, TR_EMITTER_PROJECTOR = 1, TR_EMITTER_ANGULARAREA = 2 };

typedef struct TrEmitterDesc {
    int32_t type;        /* TR_EMITTER_*                                                    */
    int32_t shape;       /* area: index of the emitting shape; projector: -1                */
    float radiance[3];   /* area: radiance; projector: irradiance                           */
    float fov_x;         /* projector: horizontal field of view, degrees                    */
    float to_world[12];  /* projector: row-major 3x4 affine (origin + orientation)          */
    float world_to_local[12];    /* projector: inverse of to_world (row-major 3x4 affine)       */
    float camera_to_sample[16];  /* projector: Mitsuba's perspective_projection(1x1, fov_x, 1e-4,
                                    1e4): local space -> [0,1]^2 texture coordinates            */
    float cutoff_angle;          /* angulararea: radians (angulararea.py:57-58)                  */
    float beam_width;            /* angulararea: radians, <= cutoff_angle (:59-60)               */
    float cos_cutoff_angle;      /* angulararea: the reference computes these three in double     */
    float cos_beam_width;        /*   precision at construction (:61-67) and uses them as fp32    */
    float inv_transition_width;  /*   constants: 1 / (cutoff - beam), +inf when they are equal    */
    float _pad;
} TrEmitterDesc;

typedef struct TrMediumDesc { /* homogeneous medium + Henyey-Greenstein phase function      */
    float sigma_t[3];
    float albedo[3];
    float g;
    float scale;
} TrMediumDesc;

enum { TR_SENSOR_PERSPECTIVE = 0, TR_SENSOR_NLOS_CAPTURE_METER = 1 };

typedef struct TrSensorDesc {
    int32_t type;           /* TR_SENSOR_*                                                  */
    int32_t shape;          /* nlos_capture_meter: index of the relay-wall shape            */
    float to_world[12];     /* perspective: camera-to-world, row-major 3x4                  */
    float sample_to_camera[16]; /* perspective: inverse of Mitsuba's perspective_projection */
    float near_clip, far_clip;
    float origin[3];        /* nlos_capture_meter: sensor_origin
                               (mitransient/sensors/nloscapturemeter.py:100-102)            */
    int32_t medium;         /* index of the medium the sensor sits in, or -1                */
    int32_t is_confocal;    /* nlos_capture_meter with original_film_width/height: a 1x1 film whose every ray goes to
                               laser_target (nloscapturemeter.py:113-125, 142-144)          */
    float laser_target[3];
} TrSensorDesc;

/* Film kinds.  TR_FILM_TRANSIENT_HDR: time histogram, film[y][x][t][c]
 * (mitransient/films/transient_hdr_film.py, render/transient_image_block.py:63-68).  With
 * laser_scan_width/height > 0 it is the `exhaustive_scan` layout film[y][x][ly][lx][t][c] addressed with
 * the reference's flat index ((((y*W + x)*Lw + laser_x)*Lh + laser_y)*T + t)*C
 * (transient_image_block.py:59-66, 134-140).
 * TR_FILM_PHASOR_HDR: frequency-domain film (mitransient/films/phasor_hdr_film.py,
 * render/phasor_image_block.py:43-67): film[y][x][2*F + 1] = (Re, Im) per stored frequency + weight;
 * a deposit (spec, opl) adds spec.x * (cos, sin)(fmod(-2 pi f (opl - start_opl), 2 pi)) for every
 * frequency f; there is no time window (every finite opl counts).                                    */
enum { TR_FILM_TRANSIENT_HDR = 0, TR_FILM_PHASOR_HDR = 1 };

typedef struct TrFilmDesc {  /* mitransient/films/transient_hdr_film.py:114-121             */
    uint32_t width, height;        /* the rendered window = film.crop_size() (the whole film unless a crop window is
                                      set): the block the reference allocates (transient_hdr_film.py:192-197)          */
    uint32_t temporal_bins;
    float start_opl;
    float bin_width_opl;
    uint32_t laser_scan_width, laser_scan_height; /* exhaustive_scan film (transient_hdr_film.py:119-121); 0, 0 = off */
    uint32_t type;                 /* TR_FILM_*                                             */
    uint32_t n_frequencies;        /* phasor film: F                                        */
    uint32_t crop_offset_x, crop_offset_y; /* film.crop_offset(): where the window sits in the full film.  Enters only the
                                      film-sample arithmetic of ADIntegrator.sample_rays [upstream]: pos = pixel + offset,
                                      sample = fma(pos + jitter, 1/crop_size, -offset/crop_size); sample_to_camera is built
                                      for the window.  The block subtracts the offset again
                                      (transient_image_block.py:131), so film[y][x] is window-relative.                 */
    uint32_t _pad;
    const float *frequencies;      /* phasor film: host, [F] (phasor_hdr_film.py:125-137)   */
} TrFilmDesc;

typedef struct TrSceneDesc {
    uint32_t abi_version;          /* must be TR_ABI_VERSION                                */
    uint32_t n_prims;
    const float *prim_verts;       /* host, [n_prims][4 vertices][xyz], world space         */
    const uint8_t *prim_types;     /* host, [n_prims] TR_PRIM_*                             */
    uint32_t n_shapes;
    const TrShapeDesc *shapes;     /* host                                                  */
    uint32_t n_bsdfs;
    const TrBsdfDesc *bsdfs;       /* host                                                  */
    uint32_t n_emitters;
    const TrEmitterDesc *emitters; /* host                                                  */
    uint32_t n_media;
    const TrMediumDesc *media;     /* host (may be NULL when n_media == 0)                  */
    TrSensorDesc sensor;
    TrFilmDesc film;
} TrSceneDesc;

/* ------------------------------------------------------------------ render parameters
 * Integrator properties keep the reference's names and defaults
 * (mitransient/integrators/common.py:22-30, transientpath.py:43-86,
 *  transientnlospath.py:201-249).                                                         */

enum { TR_INTEGRATOR_TRANSIENT_PATH = 0,
       TR_INTEGRATOR_TRANSIENT_NLOS_PATH = 1,
       TR_INTEGRATOR_TRANSIENT_PRBVOLPATH = 2 };

enum { TR_FLAG_CAMERA_UNWARP = 1,          /* common.py:25                                  */
       TR_FLAG_DISCARD_DIRECT_LIGHT = 2,   /* common.py:27                                  */
       TR_FLAG_HIDE_EMITTERS = 4,          /* transientpath.py:161-163                      */
       /* transient_nlos_path only (transientnlospath.py:201-249): */
       TR_FLAG_NLOS_LASER_SAMPLING = 8,
       TR_FLAG_NLOS_HG_SAMPLING = 16,
       TR_FLAG_NLOS_HG_SAMPLING_DO_RROULETTE = 32,
       TR_FLAG_NLOS_HG_SAMPLING_INCLUDES_RELAY_WALL = 64,
       TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES = 128,
       TR_FLAG_DISCARD_DIRECT_PATHS = 256 };

enum { TR_HIST_AUTO = 0, TR_HIST_SHARED = 1, TR_HIST_GLOBAL = 2 };
enum { TR_CAPTURE_SINGLE = 0, TR_CAPTURE_CONFOCAL = 1, TR_CAPTURE_EXHAUSTIVE = 2 };

typedef struct TrRenderParams {
    int32_t integrator;     /* TR_INTEGRATOR_*                                              */
    uint32_t max_depth;     /* 0xffffffff == unbounded (reference: max_depth = -1)          */
    uint32_t rr_depth;
    uint32_t flags;         /* TR_FLAG_*                                                    */
    uint32_t spp;           /* samples per pixel of THIS pass                               */
    uint32_t total_spp;     /* sum over passes; deposits are scaled by 1/total_spp
                               (mitransient/integrators/common.py:150-155,173)              */
    uint32_t seed;          /* sampler seed of this pass (common.py:52,72)                  */
    /* pixel-tile sharding: tiles of tile_size x tile_size pixels are dealt
       round-robin with a one-tile skew per tile row; this call renders the tiles t (row-major
       index, tile row ty) with (t + ty) % world == rank.  world == 1
       renders everything.  RNG streams are indexed by the GLOBAL sample index, so the
       film does not depend on `world`.                                                    */
    uint32_t tile_size;
    uint32_t rank, world;
    int32_t film_channels;  /* layout of `film` passed to tr_render: 4 = raw RGBW
                               (develop(raw=True)), 3 = developed RGB; phasor film: 2F + 1  */
    int32_t accumulate;     /* 0: owned pixels are overwritten; 1: added to (multi-pass)    */
    /* transient_nlos_path */
    int32_t capture_type;   /* TR_CAPTURE_*                                                 */
    int32_t filter_depth;   /* -1 = off                                                     */
    /* Exhaustive captures (transientnlospath.py:340-379): with force_equal_illumination_scanning the illuminated
       points are the sensor's scanned points; otherwise they come from a laser_scan_width x laser_scan_height grid
       of rays shot through a copy of the projector with the field of view illumination_scan_fov --
       laser_scan_sample_to_camera is that copy's sample -> camera matrix (inverse of Mitsuba's
       perspective_projection(1x1, fov, 1e-4, 1e4)), row-major 4x4.                                        */
    int32_t force_equal_grids;
    float laser_scan_sample_to_camera[16];
    /* Where the kernel accumulates a pixel's (time x RGBW) histogram -- a device-side choice with no counterpart in
       the reference; both give the same film (the same deposits, summed in a different order).  AUTO picks by
       measured cost ("Two histogram paths"); SHARED / GLOBAL state a preference, honoured whenever that
       path can serve the call (tr_last_render_info tells which one ran).                                  */
    uint32_t histogram_path; /* TR_HIST_*                                                    */
    uint32_t _reserved[5];  /* (the `single` laser target and the per-pixel sensor targets,
                               transientnlospath.py:295-339, are traced on the device)      */
} TrRenderParams;

/* Exact work counters; identical integers are produced by the CPU oracle, which makes
 * them a bit-exact parity check on the integer/index side of the path.                   */
typedef struct TrCounters {
    uint64_t paths;             /* camera paths started                                     */
    uint64_t vertices;          /* N_v: loop iterations executed (live path vertices)       */
    uint64_t deposits;          /* non-zero deposits offered to the film                    */
    uint64_t deposits_in_range; /* ... that passed the [start, start + T*bw) mask           */
    uint64_t bin_checksum;      /* sum over in-range deposits of (pixel_index * T + bin)    */
    uint64_t shadow_rays;       /* visibility tests issued                                  */
    uint64_t reserved[2];
} TrCounters;

typedef struct TrScene TrScene; /* opaque: device copy of the scene + BVH                   */

/* One recorded deposit = the arguments of TransientHDRFilm.add_transient_data
 * (mitransient/films/transient_hdr_film.py:250-276) after the 1/spp pre-scale.           */
typedef struct TrDeposit {
    float x, y;     /* splatting position (integer valued for the box filter)               */
    float opl;      /* optical path length ("distance")                                     */
    float r, g, b;  /* spec * sample_scale * ray_weight                                     */
} TrDeposit;

/* ------------------------------------------------------------------ entry points */

/* Library / device probe.  Returns TR_OK iff an sm_100 class device is current.           */
int tr_device_check(void);
int tr_abi_version(void);
const char *tr_last_error(void);

/* Upload a scene, build its BVH.  Replaces mi.load_dict()'s hand-off of shapes / bsdfs /
 * emitters / sensor / film to the integrator (scene services called from
 * mitransient/integrators/transientpath.py:134-231).                                      */
int tr_scene_create(const TrSceneDesc *desc, TrScene **out_scene);
void tr_scene_destroy(TrScene *scene);
/* Upload the packed scene again (host -> device, asynchronous on `stream`, from page-locked host memory): the
 * per-render input transfer of a caller that keeps its scene on the host between renders (Dr.Jit re-uploads the
 * scene parameters it traced; mi.render on a CPU variant owns no device state).  Used by bench.py's e2e leg.     */
int tr_scene_refresh(TrScene *scene, void *stream);
/* bytes of the packed scene blob, number of BVH nodes (introspection for tests)           */
int tr_scene_info(const TrScene *scene, uint64_t *blob_bytes, uint32_t *n_nodes, uint32_t *n_prims);

/* THE hot path.  Replaces TransientADIntegrator.render's per-pass body
 * (mitransient/integrators/common.py:157-206): sample_rays -> TransientPath.sample
 * (transientpath.py:88-326) / TransientNLOSPath.sample (transientnlospath.py:673-927)
 * -> add_transient_data (transient_hdr_film.py:250-276) -> TransientImageBlock.put_/accum
 * (transient_image_block.py:103-151,79-81), plus the steady-image splat.
 *   film     device, float[H][W][T][film_channels]; exhaustive_scan: float[H][W][Lh][Lw][T][film_channels]
 *            (must be zero-filled by the caller, accumulate = 1); phasor film: float[H][W][2F+1]
 *   steady   device, float[H][W][3]   (mean of L over total_spp), may be NULL
 *   counters device, TrCounters accumulated with atomics, may be NULL
 * Asynchronous on `stream`.                                                               */
int tr_render(const TrScene *scene, const TrRenderParams *params,
              float *film, float *steady, TrCounters *counters, void *stream);

/* Same, but film / steady are HOST buffers (pinned or pageable): the film is rendered in
 * row bands on the device and copied out band by band so the device->host copy of band i
 * overlaps the rendering of band i+1.  Synchronous (returns when the host buffers are
 * complete).  This is what mi.render() returns on a CPU variant.                          */
int tr_render_to_host(const TrScene *scene, const TrRenderParams *params,
                      float *film_host, float *steady_host, TrCounters *counters_host,
                      void *stream);
/* With params->world > 1 tr_render_to_host renders only the pixels this rank owns and film_host /
 * steady_host are the COMPACT arrays float[n_owned][T][film_channels] / float[n_owned][3], pixels in the
 * order tr_owned_pixels reports (owned tiles row-major, pixels row-major inside a tile): every GPU of a
 * node then moves only its share of the film over its own PCIe link.  tr_owned_pixels needs no GPU:
 * out == NULL only counts; otherwise *n_out holds the capacity of `out` on entry.                     */
int tr_owned_pixels(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t rank, uint32_t world,
                    uint32_t *out, uint64_t *n_out);

/* K4 on its own: TransientImageBlock.put_ for a stream of deposits
 * (transient_hdr_film.py:263-276 + transient_image_block.py:130-149).  `deposits` is a
 * device array; film is device float[H][W][T][channels], accumulated into (not cleared). */
int tr_film_scatter_add(float *film, const TrFilmDesc *film_desc, int32_t channels,
                        const TrDeposit *deposits, uint64_t n_deposits, void *stream);

/* TransientHDRFilm.develop_transient_ (transient_hdr_film.py:220-248): drop the weight
 * channel, dividing by it where it is non-zero.  in: float[n_bins][in_channels] with the
 * weight last; out: float[n_bins][in_channels-1].                                         */
int tr_film_develop(const float *film_raw, float *out, uint64_t n_bins, int32_t in_channels,
                    void *stream);

/* ------------------------------------------------------------------ multi-GPU film assembly
 * Pixel tiles have exactly one writer, so the ranks of one node can flush their finished pixel
 * histograms straight into ONE film that lives on the destination rank, mapped into every process through
 * CUDA IPC (NVLink peer memory).  The reference has no counterpart (single process); this replaces the
 * "NCCL reduce of the film" with stores fused into the render kernel.
 *   tr_peer_alloc  on the destination rank: zero-filled device buffer + a 64-byte handle to send to the peers
 *   tr_peer_open   on the other ranks: map it (enables peer access); pass the pointer as `film` to tr_render
 *   tr_peer_close / tr_peer_free   undo the above                                                          */
#define TR_PEER_HANDLE_BYTES 64
int tr_peer_alloc(uint64_t bytes, void **dev_ptr, unsigned char *handle /* [TR_PEER_HANDLE_BYTES] */);
int tr_peer_open(const unsigned char *handle /* [TR_PEER_HANDLE_BYTES] */, void **dev_ptr);
int tr_peer_close(void *dev_ptr);
int tr_peer_free(void *dev_ptr);

/* ------------------------------------------------------------------ self test of the arithmetic contract
 * The ray / primitive tests divide through rcp_c(x): the correctly rounded 1/x, which the kernels compute as
 * MUFU.RCP + one Newton step without nvcc's range guard (DESIGN.md "Arithmetic contract").  This runs that sequence
 * over EVERY fp32 x with 2^-126 <= |x| < 2^126 on the device and counts the values that differ from the IEEE
 * division 1.0f / x (expected: 0).  No reference counterpart (the reference delegates intersection to Embree/OptiX). */
/* What the last tr_render / tr_render_sharded / tr_render_to_host launch of THIS host thread used (introspection for
 * tests and bench.py): the histogram path that ran (TR_HIST_SHARED or TR_HIST_GLOBAL), its shared-memory histogram
 * slots per CTA, threads per CTA, dynamic shared memory per CTA in bytes.  All zero before the first launch.        */
typedef struct TrRenderInfo { uint32_t histogram_path, slots, threads, smem_bytes; } TrRenderInfo;
void tr_last_render_info(TrRenderInfo *out);

int tr_selftest_rcp(uint64_t *n_checked, uint64_t *n_mismatch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MITRANSIENT_B200_H */
