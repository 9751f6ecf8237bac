// Packed device scene ("blob") shared by tr_scene.cu (host packer / BVH builder) and the kernels.
//
// The whole scene is ONE contiguous, 16-byte aligned byte blob so that a CTA can stage it into
// shared memory with a single TMA bulk copy (cp.async.bulk, SASS UBLKCP) when it is small enough
// (Cornell box: ~7 KB), and read it through the read-only path from global memory otherwise.
#pragma once
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mitransient_b200.h"

namespace tr {

struct BlobHeader {            // 112 bytes
    uint32_t total_bytes;      // multiple of 16
    uint32_t n_prims, n_nodes, n_shapes, n_bsdfs, n_emitters, n_media;
    uint32_t off_prims, off_nodes, off_shapes, off_bsdfs, off_emitters, off_media, off_sampling;
    uint32_t n_area_emitters;  // emitters usable by Scene::sample_emitter_direction
    uint32_t off_geom;         // DevGeom[n_prims], ORIGINAL primitive order
    uint32_t off_occluders;    // DevPrim[n_occluders]: the non-boundary primitives (brute-force shadow loop)
    uint32_t n_occluders;
    uint32_t off_box_extra;    // DevBoxExtra[n_boxes]: sampling data of box primitives
    uint32_t n_boxes;
    uint32_t n_planar;         // brute-force order: DevPrim[0, n_planar) planar, [n_planar, n_prims) boxes
    uint32_t n_occ_planar;     // same split inside the occluder list
    // brute-force order inside the planar range: axis-aligned rectangles with normal axis x in [0, aa_end[0]), y in
    // [aa_end[0], aa_end[1]), z in [aa_end[1], aa_end[2]), the other planar primitives in [aa_end[2], n_planar)
    uint32_t aa_end[3];
    uint32_t occ_aa_end[3];    // same split inside the occluder list
};
static_assert(sizeof(BlobHeader) == 112, "BlobHeader must be 112 bytes");

// Intersection record of one planar primitive (triangle or parallelogram), 80 bytes = 5 x float4.
//   N = cross(e1,e2), d0 = dot(N,p0), a1 = cross(e2,N)/|N|^2, c1 = dot(p0,a1), a2 = cross(N,e1)/|N|^2, c2 = dot(p0,a2)
// Order in the blob: when the scene is small enough for the lock-step brute-force loop (n_nodes == 0), planar
// primitives in original order followed by boxes in original order (the oracle tests them in that order); BVH
// leaf order otherwise.
struct DevPrim {
    float N[3];  float d0;     // box: a1, c1   (duals of the parallelepiped: u = a1.x - c1, ...)
    float a1[3]; float c1;     // box: a2, c2
    float a2[3]; float c2;     // box: a3, c3
    float lim;                 // u + v <= |lim| : 1 for triangles, 3 for parallelograms; NEGATIVE for boundary
                               // primitives (never occlude a shadow ray, see tr_scene.cu classify_boundary);
                               // 5 marks a TR_PRIM_BOX; +-7 an axis-aligned rectangle, whose first six words are
                               // (c, lo_a, hi_a, lo_b | hi_b, k): plane x_k = c (k stored as an integer), extent along
                               // the in-plane axes a = (k + 1) % 3, b = (k + 2) % 3
    uint32_t shape;
    uint32_t prim;             // original primitive index (tie-break key)
    float area;
    float n[3];  uint32_t _r0; // unit geometric normal (shading); box: 1/|a1|, 1/|a2|, 1/|a3|
};
static_assert(sizeof(DevPrim) == 80, "DevPrim must be 80 bytes");

struct DevGeom {               // 48 bytes, used by position sampling (emitters, hidden geometry)
    float p0[3]; uint32_t quad;   // TR_PRIM_* (0 triangle, 1 quad, 2 box)
    float e1[3]; float area;
    float e2[3]; uint32_t extra;  // box: byte offset in the blob of its DevBoxExtra
};
static_assert(sizeof(DevGeom) == 48, "DevGeom must be 48 bytes");

struct DevBoxExtra {           // 80 bytes: what sampling a point on a box needs beyond DevGeom
    float e3[3]; float fa0;       // third edge; face areas |e2 x e3|, |e3 x e1|, |e1 x e2|
    float a1[3]; float fa1;
    float a2[3]; float fa2;
    float a3[3]; float il0;       // duals and their inverse lengths (outward normals = +- a_k / |a_k|)
    float il1, il2, _r0, _r1;
};
static_assert(sizeof(DevBoxExtra) == 80, "DevBoxExtra must be 80 bytes");

struct DevNode {               // 64 bytes = 4 x float4: one INNER node of the BVH with the (padded) boxes of both children,
    float lo0[3], hi0[3];      // breadth-first order.  Words: lo0.xyz hi0.x | hi0.yz lo1.xy | lo1.z hi1.xyz | child0 child1 - -
    float lo1[3], hi1[3];
    uint32_t child0, child1;   // inner node index, or 0x80000000 | first_prim << 4 | count for a leaf
    uint32_t _r0, _r1;
};
static_assert(sizeof(DevNode) == 64, "DevNode must be 64 bytes");

struct DevShape {              // 128 bytes
    int32_t type, bsdf, emitter, flags;
    uint32_t prim_begin, prim_count, samp_off /* byte offset in blob of the pmf | cdf arrays */, _r0;
    float inv_area, area_sum, area_norm, _r1;
    float to_world[12];
    float n[3]; int32_t interior_medium;
    float s[3]; int32_t exterior_medium;
};
static_assert(sizeof(DevShape) == 128, "DevShape must be 128 bytes");

struct DevBsdf { int32_t type, flags; float r[3]; float eta[3]; float k[3]; float tr[3]; float _pad[2]; };   // 64 bytes = 4 x float4
static_assert(sizeof(DevBsdf) == 64, "DevBsdf must be 64 bytes");

struct DevEmitter {            // 96 bytes
    int32_t type, shape; float radiance[3]; float fov_x;   // 24
    float to_world[12];                                    // 72
    float cutoff_angle, beam_width, cos_cutoff, cos_beam;  // angulararea (angulararea.py:57-67), offsets 72..84
    float inv_transition, _pad;
};
static_assert(sizeof(DevEmitter) == 96, "DevEmitter must be 96 bytes");

struct DevMedium { float sigma_t[3]; float g; float albedo[3]; float scale; };  // 32 bytes

// Host-side packed scene + device copies; this is what the opaque TrScene points to.
struct SceneImpl {
    std::vector<unsigned char> blob;     // host copy
    void *d_blob = nullptr;              // device copy
    void *h_pinned = nullptr;            // page-locked host copy (tr_scene_refresh uploads from it)
    TrSensorDesc sensor;
    TrFilmDesc film;
    uint32_t n_nodes = 0, n_prims = 0;
    bool has_delta_bsdfs = false;        // conductor / dielectric / angulararea present: the general-material kernel instantiation is used
    bool has_angular_emitters = false;
    bool has_null_bsdfs = false;         // transient_path walks through null boundaries as delta lobes (general-material kernels)
    int device = -1;
    // cached owned-pixel list (tile sharding), rebuilt when (tile, rank, world) changes
    void *d_pixels = nullptr; uint32_t n_pixels = 0; uint32_t px_tile = 0, px_rank = 0, px_world = 0;
    // scratch for counters / work queue
    void *d_scratch = nullptr;
    // host copies used by the NLOS prepare step (transientnlospath.py:251-383)
    std::vector<TrShapeDesc> shapes; std::vector<TrEmitterDesc> emitters; std::vector<float> shape_area;
    std::vector<float> frequencies; void *d_freqs = nullptr;   // phasor film: host / device copy of TrFilmDesc.frequencies
    void *d_nlos = nullptr;              // float4 targets[H*W + 1] | float hg_table[2 * n_shapes]
    uint32_t nlos_key = 0xffffffffu; float hg_sum = 0.f, hg_norm = 0.f;
    int nlos_grid_valid = 1;             // does the film's own laser grid (Exhaustive capture) hit the scene anywhere? (:370-377)
    int nlos_laser_valid = -1;           // does the laser axis hit the scene? (-1: not computed yet; transientnlospath.py:335)
    std::mutex mu;                       // tr_render mutates the caches above: one render call per scene at a time
    size_t nlos_capacity = 0; uint32_t nlos_grid = 0; float nlos_s2c[16] = {0};   // laser grid of an Exhaustive capture with its own fov
};

bool pack_scene(const TrSceneDesc *desc, SceneImpl &out, std::string &err);

}  // namespace tr

struct TrScene { tr::SceneImpl impl; };
