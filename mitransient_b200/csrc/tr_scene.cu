// Host side of the scene: validation, per-triangle precomputation, BVH build, blob packing,
// upload.  Replaces what Mitsuba's Scene / Shape constructors + Embree/OptiX acceleration build do
// for the integrator (scene services used at mitransient/integrators/transientpath.py:134-231).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tr_common.h"
#include "tr_scene.h"

namespace tr {
namespace {

struct H3 { float x, y, z; };
inline H3 h3(float x, float y, float z) { return H3{x, y, z}; }
inline H3 operator-(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline H3 operator*(H3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float hdot(H3 a, H3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline H3 hcross(H3 a, H3 b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
inline H3 hxf_vector(const float *m, H3 v) {
    return h3(fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)), fmaf(m[4], v.x, fmaf(m[5], v.y, m[6] * v.z)),
              fmaf(m[8], v.x, fmaf(m[9], v.y, m[10] * v.z)));
}

struct BuildTri { float lo[3], hi[3], c[3]; uint32_t idx; };
struct BuildNode { float lo[3], hi[3]; int left = -1, right = -1; uint32_t first = 0, count = 0; };

constexpr uint32_t kLeafMax = 4;                // primitives per BVH leaf (the leaf code holds up to 15)
constexpr uint32_t kBruteForceMaxPrims = 40;   // up to this many primitives: no BVH, lock-step loop in the kernel.  Measured (round 2, 256 spp):
                                               // 27 primitives 27.5 ms lock-step; 87 primitives 69.3 ms lock-step vs 39.6 ms BVH; 18: 9.8 vs 18.3 ms
constexpr int kSahBins = 16;
constexpr int kSahMaxDepth = 40;               // deeper than this the build switches to median splits: bounds the traversal stack

inline float half_area(const float *lo, const float *hi) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// Binned surface-area-heuristic build (16 bins per axis, all three axes), leaves of up to kLeafMax primitives.
int build_rec(std::vector<BuildNode> &nodes, std::vector<BuildTri> &tris, uint32_t begin, uint32_t end, int depth, int &max_depth) {
    BuildNode nd;
    float clo[3], chi[3];
    for (int a = 0; a < 3; ++a) { nd.lo[a] = clo[a] = INFINITY; nd.hi[a] = chi[a] = -INFINITY; }
    for (uint32_t i = begin; i < end; ++i)
        for (int a = 0; a < 3; ++a) {
            nd.lo[a] = std::min(nd.lo[a], tris[i].lo[a]); nd.hi[a] = std::max(nd.hi[a], tris[i].hi[a]);
            clo[a] = std::min(clo[a], tris[i].c[a]); chi[a] = std::max(chi[a], tris[i].c[a]);
        }
    int me = (int) nodes.size();
    nodes.push_back(nd);
    max_depth = std::max(max_depth, depth);
    const uint32_t n = end - begin;
    if (n <= 2) { nodes[me].first = begin; nodes[me].count = n; return me; }
    uint32_t mid = 0;
    bool split = false;
    if (depth < kSahMaxDepth) {
        float best_cost = INFINITY; int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.0f)) continue;
            uint32_t cnt[kSahBins] = {0};
            float blo[kSahBins][3], bhi[kSahBins][3];
            for (int b = 0; b < kSahBins; ++b) for (int k = 0; k < 3; ++k) { blo[b][k] = INFINITY; bhi[b][k] = -INFINITY; }
            const float scale = (float) kSahBins / ext;
            for (uint32_t i = begin; i < end; ++i) {
                int b = std::min(kSahBins - 1, std::max(0, (int) ((tris[i].c[a] - clo[a]) * scale)));
                cnt[b]++;
                for (int k = 0; k < 3; ++k) { blo[b][k] = std::min(blo[b][k], tris[i].lo[k]); bhi[b][k] = std::max(bhi[b][k], tris[i].hi[k]); }
            }
            float ra[kSahBins]; uint32_t rn[kSahBins];          // right-side sweeps
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}; uint32_t c = 0;
            for (int b = kSahBins - 1; b > 0; --b) {
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                c += cnt[b]; rn[b] = c; ra[b] = c ? half_area(lo, hi) : 0.0f;
            }
            for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
            c = 0;
            for (int b = 0; b < kSahBins - 1; ++b) {            // split between bin b and b + 1
                for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], blo[b][k]); hi[k] = std::max(hi[k], bhi[b][k]); }
                c += cnt[b];
                if (c == 0 || rn[b + 1] == 0) continue;
                float cost = half_area(lo, hi) * (float) c + ra[b + 1] * (float) rn[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
            }
        }
        // leaf when splitting does not pay (cost model: one box test ~ one primitive test) and the leaf is small enough
        const float leaf_cost = half_area(nd.lo, nd.hi) * (float) n;
        if (n <= kLeafMax && !(best_cost + half_area(nd.lo, nd.hi) < leaf_cost)) { nodes[me].first = begin; nodes[me].count = n; return me; }
        if (best_axis >= 0) {
            const float scale = (float) kSahBins / (chi[best_axis] - clo[best_axis]);
            const float c0 = clo[best_axis];
            auto it = std::stable_partition(tris.begin() + begin, tris.begin() + end, [=](const BuildTri &t) {
                int b = std::min(kSahBins - 1, std::max(0, (int) ((t.c[best_axis] - c0) * scale)));
                return b <= best_bin;
            });
            mid = (uint32_t) (it - tris.begin());
            split = mid > begin && mid < end;
        }
    }
    if (!split) {   // identical centroids, or too deep: median split along the widest axis (by index when degenerate)
        int axis = 0;
        for (int a = 1; a < 3; ++a) if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
        if (n <= kLeafMax && !(chi[axis] - clo[axis] > 0.0f)) { nodes[me].first = begin; nodes[me].count = n; return me; }
        mid = begin + n / 2;
        std::nth_element(tris.begin() + begin, tris.begin() + mid, tris.begin() + end,
                         [axis](const BuildTri &a, const BuildTri &b) {
                             return a.c[axis] < b.c[axis] || (a.c[axis] == b.c[axis] && a.idx < b.idx);
                         });
    }
    int l = build_rec(nodes, tris, begin, mid, depth + 1, max_depth);
    int r = build_rec(nodes, tris, mid, end, depth + 1, max_depth);
    nodes[me].left = l; nodes[me].right = r;
    return me;
}

inline uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

}  // namespace

bool pack_scene(const TrSceneDesc *desc, SceneImpl &out, std::string &err) {
    if (!desc) { err = "scene descriptor is NULL"; return false; }
    if (desc->abi_version != TR_ABI_VERSION) { err = "TrSceneDesc.abi_version does not match this library"; return false; }
    if (desc->n_prims == 0 || !desc->prim_verts || !desc->prim_types) { err = "scene has no primitives"; return false; }
    if (desc->n_shapes == 0 || !desc->shapes) { err = "scene has no shapes"; return false; }
    if (desc->n_bsdfs == 0 || !desc->bsdfs) { err = "scene has no bsdfs"; return false; }
    if (desc->n_emitters && !desc->emitters) { err = "emitters pointer is NULL"; return false; }
    if (desc->n_prims >= (1u << 26)) { err = "too many primitives"; return false; }
    const TrFilmDesc &f = desc->film;
    if (f.width == 0 || f.height == 0 || f.temporal_bins == 0) { err = "film has a zero dimension"; return false; }
    if (!(f.bin_width_opl > 0.0f)) { err = "bin_width_opl must be > 0"; return false; }
    if ((uint64_t) f.width * f.height >= (1ull << 31)) { err = "film too large"; return false; }
    if (f.type != TR_FILM_TRANSIENT_HDR && f.type != TR_FILM_PHASOR_HDR) { err = "unknown film type"; return false; }
    if (f.type == TR_FILM_PHASOR_HDR && (f.n_frequencies == 0 || !f.frequencies || f.n_frequencies > 4096)) {
        err = "phasor film needs 1..4096 frequencies"; return false;
    }
    if ((f.laser_scan_width == 0) != (f.laser_scan_height == 0)) { err = "laser_scan_width and laser_scan_height must both be set"; return false; }
    if (f.laser_scan_width && f.type != TR_FILM_TRANSIENT_HDR) { err = "exhaustive_scan is a transient_hdr_film layout"; return false; }
    if ((uint64_t) f.width * f.height * std::max(1u, f.laser_scan_width) * std::max(1u, f.laser_scan_height) * f.temporal_bins >= (1ull << 40)) {
        err = "film too large"; return false;
    }

    const uint32_t np = desc->n_prims, ns = desc->n_shapes;
    std::vector<uint32_t> prim_shape(np, 0xffffffffu);
    for (uint32_t s = 0; s < ns; ++s) {
        const TrShapeDesc &sd = desc->shapes[s];
        if ((uint64_t) sd.prim_begin + sd.prim_count > np) { err = "shape primitive range out of bounds"; return false; }
        if (sd.prim_count == 0) { err = "shape without primitives"; return false; }
        if (sd.bsdf < 0 || (uint32_t) sd.bsdf >= desc->n_bsdfs) { err = "shape bsdf index out of range"; return false; }
        if (sd.emitter >= (int32_t) desc->n_emitters) { err = "shape emitter index out of range"; return false; }
        if (sd.type == TR_SHAPE_RECTANGLE && (sd.prim_count != 1 || desc->prim_types[sd.prim_begin] != TR_PRIM_QUAD)) {
            err = "a rectangle shape must consist of exactly one quad primitive"; return false;
        }
        for (uint32_t k = 0; k < sd.prim_count; ++k) prim_shape[sd.prim_begin + k] = s;
    }
    for (uint32_t i = 0; i < np; ++i)
        if (prim_shape[i] == 0xffffffffu) { err = "primitive not covered by any shape"; return false; }

    // ---- per-primitive data (same arithmetic as the oracle's build_scene)
    std::vector<DevPrim> prims(np);
    std::vector<DevGeom> geom(np);
    std::vector<BuildTri> bprims(np);
    float slo[3] = {INFINITY, INFINITY, INFINITY}, shi[3] = {-INFINITY, -INFINITY, -INFINITY};
    std::vector<DevBoxExtra> box_extra;
    std::vector<uint32_t> box_index(np, 0xffffffffu);
    for (uint32_t i = 0; i < np; ++i) {
        const float *v = desc->prim_verts + 12 * (size_t) i;
        for (int k = 0; k < 12; ++k) if (!std::isfinite(v[k])) { err = "non-finite vertex"; return false; }
        if (desc->prim_types[i] > TR_PRIM_BOX) { err = "unknown primitive type"; return false; }
        const bool quad = desc->prim_types[i] == TR_PRIM_QUAD, box = desc->prim_types[i] == TR_PRIM_BOX;
        H3 p0 = h3(v[0], v[1], v[2]);
        H3 e1 = h3(v[3], v[4], v[5]) - p0, e2 = h3(v[6], v[7], v[8]) - p0, e3 = h3(v[9], v[10], v[11]) - p0;
        H3 N = hcross(e1, e2);
        float nn = hdot(N, N);
        if (!(nn > 0.0f)) { err = "degenerate primitive (zero area)"; return false; }
        float inv_nn = 1.0f / nn;
        float len = sqrtf(nn);
        H3 n = N * (1.0f / len);
        DevPrim &t = prims[i];
        memset(&t, 0, sizeof(t));
        DevGeom &g = geom[i];
        memset(&g, 0, sizeof(g));
        if (box) {   // same arithmetic as the oracle's build_scene (TR_PRIM_BOX)
            H3 c23 = hcross(e2, e3), c31 = hcross(e3, e1);
            float det = hdot(e1, c23);
            if (!(std::fabs(det) > 0.0f)) { err = "degenerate box primitive"; return false; }
            float inv_det = 1.0f / det;
            H3 a1 = c23 * inv_det, a2 = c31 * inv_det, a3 = N * inv_det;
            t.N[0] = a1.x; t.N[1] = a1.y; t.N[2] = a1.z; t.d0 = hdot(p0, a1);
            t.a1[0] = a2.x; t.a1[1] = a2.y; t.a1[2] = a2.z; t.c1 = hdot(p0, a2);
            t.a2[0] = a3.x; t.a2[1] = a3.y; t.a2[2] = a3.z; t.c2 = hdot(p0, a3);
            float il0 = 1.0f / sqrtf(hdot(a1, a1)), il1 = 1.0f / sqrtf(hdot(a2, a2)), il2 = 1.0f / sqrtf(hdot(a3, a3));
            t.n[0] = il0; t.n[1] = il1; t.n[2] = il2;
            float fa0 = sqrtf(hdot(c23, c23)), fa1 = sqrtf(hdot(c31, c31)), fa2 = len;
            t.lim = 5.0f;
            t.area = 2.0f * ((fa2 + fa0) + fa1);
            DevBoxExtra be; memset(&be, 0, sizeof(be));
            be.e3[0] = e3.x; be.e3[1] = e3.y; be.e3[2] = e3.z; be.fa0 = fa0;
            be.a1[0] = a1.x; be.a1[1] = a1.y; be.a1[2] = a1.z; be.fa1 = fa1;
            be.a2[0] = a2.x; be.a2[1] = a2.y; be.a2[2] = a2.z; be.fa2 = fa2;
            be.a3[0] = a3.x; be.a3[1] = a3.y; be.a3[2] = a3.z; be.il0 = il0; be.il1 = il1; be.il2 = il2;
            box_index[i] = (uint32_t) box_extra.size();
            box_extra.push_back(be);
        } else {
            H3 a1 = hcross(e2, N) * inv_nn, a2 = hcross(N, e1) * inv_nn;
            t.N[0] = N.x; t.N[1] = N.y; t.N[2] = N.z; t.d0 = hdot(N, p0);
            t.a1[0] = a1.x; t.a1[1] = a1.y; t.a1[2] = a1.z; t.c1 = hdot(p0, a1);
            t.a2[0] = a2.x; t.a2[1] = a2.y; t.a2[2] = a2.z; t.c2 = hdot(p0, a2);
            t.lim = quad ? 3.0f : 1.0f;
            t.area = quad ? len : 0.5f * len;
            t.n[0] = n.x; t.n[1] = n.y; t.n[2] = n.z;
        }
        t.shape = prim_shape[i]; t.prim = i;
        g.p0[0] = p0.x; g.p0[1] = p0.y; g.p0[2] = p0.z; g.quad = desc->prim_types[i];
        g.e1[0] = e1.x; g.e1[1] = e1.y; g.e1[2] = e1.z; g.area = t.area;
        g.e2[0] = e2.x; g.e2[1] = e2.y; g.e2[2] = e2.z;
        BuildTri &b = bprims[i];
        b.idx = i;
        for (int a = 0; a < 3; ++a) {
            float x0 = v[a], x1 = v[3 + a], x2 = v[6 + a], x3 = v[9 + a];
            float lo, hi;
            if (box) {   // all eight corners p0 + {0,1} e1 + {0,1} e2 + {0,1} e3
                float d1 = x1 - x0, d2 = x2 - x0, d3 = x3 - x0;
                lo = x0 + std::min(d1, 0.0f) + std::min(d2, 0.0f) + std::min(d3, 0.0f);
                hi = x0 + std::max(d1, 0.0f) + std::max(d2, 0.0f) + std::max(d3, 0.0f);
            } else {
                float x4 = quad ? (x1 + x2 - x0) : x0;
                lo = std::min(std::min(x0, x4), std::min(x1, x2)); hi = std::max(std::max(x0, x4), std::max(x1, x2));
            }
            b.lo[a] = lo; b.hi[a] = hi;
            b.c[a] = (b.lo[a] + b.hi[a]) * 0.5f;
            slo[a] = std::min(slo[a], b.lo[a]); shi[a] = std::max(shi[a], b.hi[a]);
        }
    }

    // ---- boundary primitives: every scene vertex lies on the front side of the supporting plane (same
    // arithmetic as the oracle's classify_boundary_prims).  They can never occlude a shadow ray.
    uint32_t n_occ = 0;
    {
        std::vector<H3> verts;
        verts.reserve((size_t) np * 4);
        for (uint32_t i = 0; i < np; ++i) {
            const DevGeom &g = geom[i];
            H3 p0 = h3(g.p0[0], g.p0[1], g.p0[2]), e1 = h3(g.e1[0], g.e1[1], g.e1[2]), e2 = h3(g.e2[0], g.e2[1], g.e2[2]);
            H3 p1 = h3(p0.x + e1.x, p0.y + e1.y, p0.z + e1.z), p2 = h3(p0.x + e2.x, p0.y + e2.y, p0.z + e2.z);
            H3 p12 = h3(p1.x + e2.x, p1.y + e2.y, p1.z + e2.z);
            verts.push_back(p0); verts.push_back(p1); verts.push_back(p2);
            if (g.quad != TR_PRIM_TRIANGLE) verts.push_back(p12);
            if (g.quad == TR_PRIM_BOX) {
                const DevBoxExtra &be = box_extra[box_index[i]];
                H3 e3 = h3(be.e3[0], be.e3[1], be.e3[2]);
                verts.push_back(h3(p0.x + e3.x, p0.y + e3.y, p0.z + e3.z)); verts.push_back(h3(p1.x + e3.x, p1.y + e3.y, p1.z + e3.z));
                verts.push_back(h3(p2.x + e3.x, p2.y + e3.y, p2.z + e3.z)); verts.push_back(h3(p12.x + e3.x, p12.y + e3.y, p12.z + e3.z));
            }
        }
        float lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY};
        const size_t n_scene_verts = verts.size();
        // end points of shadow rays that are not scene vertices (oracle: classify_boundary_prims): projector and sensor origins
        for (uint32_t e = 0; e < desc->n_emitters; ++e)
            if (desc->emitters[e].type == TR_EMITTER_PROJECTOR)
                verts.push_back(h3(desc->emitters[e].to_world[3], desc->emitters[e].to_world[7], desc->emitters[e].to_world[11]));
        if (desc->sensor.type == TR_SENSOR_NLOS_CAPTURE_METER) verts.push_back(h3(desc->sensor.origin[0], desc->sensor.origin[1], desc->sensor.origin[2]));
        else verts.push_back(h3(desc->sensor.to_world[3], desc->sensor.to_world[7], desc->sensor.to_world[11]));
        for (size_t vi = 0; vi < n_scene_verts; ++vi) {
            const H3 &x = verts[vi];
            lo3[0] = fminf(lo3[0], x.x); hi3[0] = fmaxf(hi3[0], x.x);
            lo3[1] = fminf(lo3[1], x.y); hi3[1] = fmaxf(hi3[1], x.y);
            lo3[2] = fminf(lo3[2], x.z); hi3[2] = fmaxf(hi3[2], x.z);
        }
        const float ext = fmaxf(fmaxf(hi3[0] - lo3[0], hi3[1] - lo3[1]), hi3[2] - lo3[2]);
        for (uint32_t i = 0; i < np; ++i) {
            DevPrim &t = prims[i];
            if (t.lim == 5.0f) { n_occ++; continue; }     // a solid box is never a boundary
            H3 N = h3(t.N[0], t.N[1], t.N[2]);
            float tol = 1e-4f * sqrtf(hdot(N, N)) * ext;
            bool boundary = true;
            for (const H3 &x : verts) {
                float sd = fmaf(N.z, x.z, fmaf(N.y, x.y, fmaf(N.x, x.x, -t.d0)));
                if (!(sd >= -tol)) { boundary = false; break; }
            }
            if (boundary) t.lim = -t.lim; else n_occ++;
        }
    }

    // ---- contract checks: the kernels' reciprocal (MUFU.RCP + one Newton step) is the correctly rounded 1/x for
    // 2^-126 <= |x| < 2^126; denominators are N.d / a_k.d with |d| = 1, so plane / dual vectors must stay far below that
    for (uint32_t i = 0; i < np; ++i) {
        const DevPrim &t = prims[i];
        const float *w = t.N;
        for (int k = 0; k < 12; ++k) if (!(std::fabs(w[k]) < 1.2676506e30f)) { err = "primitive too large or too small for fp32 (|N|, |a_k| must be < 2^100)"; return false; }
    }
    // ---- axis-aligned rectangles: their own primitive class (oracle: aa_rect_test).  Detected on the vertices: both
    // edges run along a coordinate axis exactly.
    std::vector<int> aa_axis(np, -1);
    for (uint32_t i = 0; i < np; ++i) {
        if (desc->prim_types[i] != TR_PRIM_QUAD) continue;
        const float *v0 = desc->prim_verts + 12 * (size_t) i, *v1 = v0 + 3, *v2 = v0 + 6;
        int ax1 = -1, ax2 = -1, n1 = 0, n2 = 0;
        for (int c = 0; c < 3; ++c) { if (v1[c] != v0[c]) { ax1 = c; n1++; } if (v2[c] != v0[c]) { ax2 = c; n2++; } }
        if (!(n1 == 1 && n2 == 1 && ax1 != ax2)) continue;
        const int k = 3 - ax1 - ax2;
        DevPrim &t = prims[i];
        float lo[2], hi[2];
        for (int j = 0; j < 2; ++j) {
            const int a = (k + 1 + j) % 3;
            lo[j] = fminf(fminf(v0[a], v1[a]), v2[a]); hi[j] = fmaxf(fmaxf(v0[a], v1[a]), v2[a]);
        }
        aa_axis[i] = k;
        t.N[0] = v0[k]; t.N[1] = lo[0]; t.N[2] = hi[0]; t.d0 = lo[1];
        t.a1[0] = hi[1]; uint32_t ku = (uint32_t) k; memcpy(&t.a1[1], &ku, 4); t.a1[2] = 0.0f; t.c1 = 0.0f;
        t.a2[0] = t.a2[1] = t.a2[2] = 0.0f; t.c2 = 0.0f;
        t.lim = t.lim < 0.0f ? -7.0f : 7.0f;
    }

    // ---- acceleration: none for small scenes (the kernel runs a lock-step loop over all primitives, original
    // order: the contract's test order); otherwise a binned-SAH BVH, breadth-first child-pair nodes.  Boxes are padded so that the fp32 slab test is
    // conservative w.r.t. the primitive test: culling never changes the result.
    std::vector<DevNode> nodes;
    std::vector<DevPrim> prims_out;
    uint32_t nn_nodes = 0;
    uint32_t n_planar = 0, aa_end[3] = {0, 0, 0}, occ_aa_end[3] = {0, 0, 0};
    if (np <= kBruteForceMaxPrims) {
        // test order of the contract: axis-aligned rectangles by normal axis, other planar primitives, boxes
        for (int k = 0; k < 3; ++k) {
            for (uint32_t i = 0; i < np; ++i) if (aa_axis[i] == k) { prims_out.push_back(prims[i]); n_planar++; }
            aa_end[k] = n_planar;
        }
        for (uint32_t i = 0; i < np; ++i) if (prims[i].lim != 5.0f && aa_axis[i] < 0) { prims_out.push_back(prims[i]); n_planar++; }
        for (uint32_t i = 0; i < np; ++i) if (prims[i].lim == 5.0f) prims_out.push_back(prims[i]);
    } else {
        std::vector<BuildNode> bnodes;
        bnodes.reserve(2 * np);
        int max_depth = 0;
        build_rec(bnodes, bprims, 0, np, 0, max_depth);
        if (max_depth >= 60) { err = "internal: BVH deeper than the traversal stack"; return false; }
        float extent = std::max(shi[0] - slo[0], std::max(shi[1] - slo[1], shi[2] - slo[2]));
        float absmax = 0.0f;
        for (int a = 0; a < 3; ++a) absmax = std::max(absmax, std::max(std::fabs(slo[a]), std::fabs(shi[a])));
        const float pad = 1e-4f * std::max(extent, absmax) + 1e-6f;
        prims_out.resize(np);
        for (uint32_t i = 0; i < np; ++i) prims_out[i] = prims[bprims[i].idx];
        // device nodes: one 64-byte record per INNER node holding the padded boxes of both children, breadth-first
        // (the top of the tree is contiguous at the start of the array).  Child code: inner node index, or
        // 0x80000000 | first << 4 | count for a leaf.
        std::vector<int> order; order.reserve(bnodes.size());          // inner build nodes in BFS order
        std::vector<uint32_t> dev_index(bnodes.size(), 0xffffffffu);
        if (bnodes[0].left < 0) { err = "internal: BVH root is a leaf"; return false; }
        order.push_back(0); dev_index[0] = 0;
        for (size_t h = 0; h < order.size(); ++h) {
            const BuildNode &b = bnodes[order[h]];
            for (int ch : {b.left, b.right})
                if (bnodes[ch].left >= 0) { dev_index[ch] = (uint32_t) order.size(); order.push_back(ch); }
        }
        nn_nodes = (uint32_t) order.size();
        nodes.resize(nn_nodes);
        auto code = [&](int ch) -> uint32_t {
            const BuildNode &c = bnodes[ch];
            return c.left >= 0 ? dev_index[ch] : (0x80000000u | (c.first << 4) | c.count);
        };
        for (uint32_t i = 0; i < nn_nodes; ++i) {
            const BuildNode &b = bnodes[order[i]];
            const BuildNode &c0 = bnodes[b.left], &c1 = bnodes[b.right];
            DevNode &d = nodes[i];
            for (int a = 0; a < 3; ++a) { d.lo0[a] = c0.lo[a] - pad; d.hi0[a] = c0.hi[a] + pad; d.lo1[a] = c1.lo[a] - pad; d.hi1[a] = c1.hi[a] + pad; }
            d.child0 = code(b.left); d.child1 = code(b.right); d._r0 = d._r1 = 0;
        }
    }

    // ---- shapes (+ area pmf / cdf tables for meshes)
    std::vector<DevShape> shapes(ns);
    std::vector<float> shape_rect_area(ns, 0.0f);
    std::vector<unsigned char> sampling;  // per mesh shape: pmf[n] | cdf[n]
    for (uint32_t s = 0; s < ns; ++s) {
        const TrShapeDesc &sd = desc->shapes[s];
        DevShape &d = shapes[s];
        memset(&d, 0, sizeof(d));
        d.type = sd.type; d.bsdf = sd.bsdf; d.emitter = sd.emitter; d.flags = sd.flags;
        d.prim_begin = sd.prim_begin; d.prim_count = sd.prim_count;
        d.interior_medium = sd.interior_medium; d.exterior_medium = sd.exterior_medium;
        memcpy(d.to_world, sd.to_world, sizeof(d.to_world));
        double acc = 0.0;
        std::vector<float> pmf(sd.prim_count), cdf(sd.prim_count);
        for (uint32_t k = 0; k < sd.prim_count; ++k) {
            float a = prims[sd.prim_begin + k].area;
            acc += (double) a; pmf[k] = a; cdf[k] = (float) acc;
        }
        d.area_sum = (float) acc; d.area_norm = (float) (1.0 / acc);
        if (sd.type == TR_SHAPE_RECTANGLE) {
            H3 dpdu = hxf_vector(sd.to_world, h3(2.0f, 0.0f, 0.0f));
            H3 dpdv = hxf_vector(sd.to_world, h3(0.0f, 2.0f, 0.0f));
            H3 c = hcross(dpdu, dpdv);
            shape_rect_area[s] = sqrtf(hdot(c, c));
            d.inv_area = 1.0f / shape_rect_area[s];
            const DevPrim &t0 = prims[sd.prim_begin];
            H3 n = h3(t0.n[0], t0.n[1], t0.n[2]);
            float dd = hdot(n, dpdu);
            H3 sv = h3(fmaf(-n.x, dd, dpdu.x), fmaf(-n.y, dd, dpdu.y), fmaf(-n.z, dd, dpdu.z));
            float inv = 1.0f / sqrtf(hdot(sv, sv));
            d.n[0] = n.x; d.n[1] = n.y; d.n[2] = n.z;
            d.s[0] = sv.x * inv; d.s[1] = sv.y * inv; d.s[2] = sv.z * inv;
            d.samp_off = 0;
        } else {
            d.inv_area = d.area_norm;
            d.samp_off = (uint32_t) sampling.size();  // relative for now, fixed up below
            size_t base = sampling.size();
            sampling.resize(base + (size_t) sd.prim_count * 8);
            memcpy(&sampling[base], pmf.data(), sd.prim_count * 4);
            memcpy(&sampling[base + sd.prim_count * 4], cdf.data(), sd.prim_count * 4);
        }
    }

    std::vector<DevBsdf> bsdfs(desc->n_bsdfs);
    for (uint32_t i = 0; i < desc->n_bsdfs; ++i) {
        memset(&bsdfs[i], 0, sizeof(DevBsdf));
        bsdfs[i].type = desc->bsdfs[i].type; bsdfs[i].flags = desc->bsdfs[i].flags;
        if (bsdfs[i].type < TR_BSDF_DIFFUSE || bsdfs[i].type > TR_BSDF_DIELECTRIC) { err = "unknown bsdf type"; return false; }
        if (bsdfs[i].type == TR_BSDF_DIELECTRIC && !(desc->bsdfs[i].eta[0] > 0.0f)) { err = "dielectric needs eta = int_ior / ext_ior > 0"; return false; }
        if (bsdfs[i].type == TR_BSDF_DIELECTRIC && (bsdfs[i].flags & TR_BSDF_FLAG_TWOSIDED)) { err = "twosided cannot wrap a transmissive bsdf"; return false; }
        for (int c = 0; c < 3; ++c) {
            bsdfs[i].r[c] = desc->bsdfs[i].reflectance[c]; bsdfs[i].eta[c] = desc->bsdfs[i].eta[c];
            bsdfs[i].k[c] = desc->bsdfs[i].k[c]; bsdfs[i].tr[c] = desc->bsdfs[i].specular_transmittance[c];
        }
        if (bsdfs[i].type >= TR_BSDF_CONDUCTOR) out.has_delta_bsdfs = true;
        if (bsdfs[i].type == TR_BSDF_NULL) out.has_null_bsdfs = true;
    }
    std::vector<DevEmitter> emitters(std::max(1u, desc->n_emitters));
    memset(emitters.data(), 0, emitters.size() * sizeof(DevEmitter));
    uint32_t n_area = 0;
    for (uint32_t i = 0; i < desc->n_emitters; ++i) {
        const TrEmitterDesc &e = desc->emitters[i];
        emitters[i].type = e.type; emitters[i].shape = e.shape; emitters[i].fov_x = e.fov_x;
        for (int c = 0; c < 3; ++c) emitters[i].radiance[c] = e.radiance[c];
        memcpy(emitters[i].to_world, e.to_world, sizeof(e.to_world));
        if (e.type == TR_EMITTER_AREA || e.type == TR_EMITTER_ANGULARAREA) {
            if (e.type == TR_EMITTER_ANGULARAREA) {
                if (!(e.cutoff_angle >= e.beam_width) || !(e.beam_width >= 0.0f)) { err = "angulararea: need cutoff_angle >= beam_width >= 0"; return false; }
                emitters[i].cutoff_angle = e.cutoff_angle; emitters[i].beam_width = e.beam_width;
                emitters[i].cos_cutoff = e.cos_cutoff_angle; emitters[i].cos_beam = e.cos_beam_width; emitters[i].inv_transition = e.inv_transition_width;
                out.has_delta_bsdfs = true; out.has_angular_emitters = true;
            }
            if (e.shape < 0 || (uint32_t) e.shape >= ns) { err = "area emitter without a valid shape"; return false; }
            if (desc->shapes[e.shape].emitter != (int32_t) i) { err = "area emitter / shape cross reference mismatch"; return false; }
            n_area++;
        } else if (e.type != TR_EMITTER_PROJECTOR) { err = "unknown emitter type"; return false; }
    }
    std::vector<DevMedium> media(std::max(1u, desc->n_media));
    memset(media.data(), 0, media.size() * sizeof(DevMedium));
    for (uint32_t i = 0; i < desc->n_media; ++i) {
        for (int c = 0; c < 3; ++c) { media[i].sigma_t[c] = desc->media[i].sigma_t[c]; media[i].albedo[c] = desc->media[i].albedo[c]; }
        media[i].g = desc->media[i].g; media[i].scale = desc->media[i].scale;
    }

    // ---- blob
    BlobHeader h; memset(&h, 0, sizeof(h));
    h.n_prims = np; h.n_nodes = nn_nodes; h.n_shapes = ns; h.n_bsdfs = desc->n_bsdfs; h.n_emitters = desc->n_emitters;
    h.n_media = desc->n_media; h.n_area_emitters = n_area;
    uint32_t off = sizeof(BlobHeader);
    h.off_prims = off; off += np * (uint32_t) sizeof(DevPrim);
    h.off_nodes = off; off += nn_nodes * (uint32_t) sizeof(DevNode);
    h.off_shapes = off; off += ns * (uint32_t) sizeof(DevShape);
    h.off_bsdfs = off; off += (uint32_t) (bsdfs.size() * sizeof(DevBsdf));
    h.off_emitters = off; off += (uint32_t) (emitters.size() * sizeof(DevEmitter));
    h.off_media = off; off += (uint32_t) (media.size() * sizeof(DevMedium));
    h.off_geom = off; off += np * (uint32_t) sizeof(DevGeom);
    h.off_sampling = off; off += align16((uint32_t) sampling.size());
    h.off_box_extra = off; h.n_boxes = (uint32_t) box_extra.size(); off += (uint32_t) (box_extra.size() * sizeof(DevBoxExtra));
    for (uint32_t i = 0; i < np; ++i) if (box_index[i] != 0xffffffffu) geom[i].extra = h.off_box_extra + box_index[i] * (uint32_t) sizeof(DevBoxExtra);
    h.off_occluders = off; h.n_occluders = n_occ; h.n_planar = n_planar;
    if (nn_nodes == 0) off += n_occ * (uint32_t) sizeof(DevPrim);
    h.total_bytes = align16(off);
    for (uint32_t s = 0; s < ns; ++s)
        if (shapes[s].type != TR_SHAPE_RECTANGLE) shapes[s].samp_off += h.off_sampling;
    out.blob.assign(h.total_bytes, 0);
    memcpy(&out.blob[0], &h, sizeof(h));
    memcpy(&out.blob[h.off_prims], prims_out.data(), np * sizeof(DevPrim));
    if (nn_nodes) memcpy(&out.blob[h.off_nodes], nodes.data(), nn_nodes * sizeof(DevNode));
    memcpy(&out.blob[h.off_shapes], shapes.data(), ns * sizeof(DevShape));
    memcpy(&out.blob[h.off_bsdfs], bsdfs.data(), bsdfs.size() * sizeof(DevBsdf));
    memcpy(&out.blob[h.off_emitters], emitters.data(), emitters.size() * sizeof(DevEmitter));
    memcpy(&out.blob[h.off_media], media.data(), media.size() * sizeof(DevMedium));
    memcpy(&out.blob[h.off_geom], geom.data(), np * sizeof(DevGeom));
    if (!sampling.empty()) memcpy(&out.blob[h.off_sampling], sampling.data(), sampling.size());
    if (!box_extra.empty()) memcpy(&out.blob[h.off_box_extra], box_extra.data(), box_extra.size() * sizeof(DevBoxExtra));
    if (nn_nodes == 0) {   // brute-force mode: compact copy of the possible occluders, planar first then boxes
        uint32_t k = 0, kp = 0;
        for (uint32_t i = 0; i < np; ++i) {     // prims_out is in test order already: the axis-aligned groups come first
            if (i == aa_end[0]) occ_aa_end[0] = kp;
            if (i == aa_end[1]) occ_aa_end[1] = kp;
            if (i == aa_end[2]) occ_aa_end[2] = kp;
            if (prims_out[i].lim > 0.0f && prims_out[i].lim != 5.0f) { memcpy(&out.blob[h.off_occluders + (k++) * sizeof(DevPrim)], &prims_out[i], sizeof(DevPrim)); kp++; }
        }
        if (np == aa_end[0]) occ_aa_end[0] = kp;
        if (np == aa_end[1]) occ_aa_end[1] = kp;
        if (np == aa_end[2]) occ_aa_end[2] = kp;
        for (uint32_t i = 0; i < np; ++i)
            if (prims_out[i].lim == 5.0f) memcpy(&out.blob[h.off_occluders + (k++) * sizeof(DevPrim)], &prims_out[i], sizeof(DevPrim));
        BlobHeader *hh = reinterpret_cast<BlobHeader *>(out.blob.data());
        hh->n_occ_planar = kp;
        for (int a = 0; a < 3; ++a) { hh->aa_end[a] = aa_end[a]; hh->occ_aa_end[a] = occ_aa_end[a]; }
    }
    out.sensor = desc->sensor; out.film = desc->film; out.n_nodes = nn_nodes; out.n_prims = np;
    out.frequencies.clear();
    if (f.type == TR_FILM_PHASOR_HDR) out.frequencies.assign(f.frequencies, f.frequencies + f.n_frequencies);
    out.film.frequencies = nullptr;      // the caller's array is not ours to keep
    out.shapes.assign(desc->shapes, desc->shapes + ns);
    out.emitters.assign(desc->emitters, desc->emitters + desc->n_emitters);
    out.shape_area.resize(ns);
    for (uint32_t s = 0; s < ns; ++s)   // Shape::surface_area(): rectangle |dp_du x dp_dv|, mesh = sum of primitive areas
        out.shape_area[s] = (shapes[s].type == TR_SHAPE_RECTANGLE) ? shape_rect_area[s] : shapes[s].area_sum;
    if (desc->sensor.type == TR_SENSOR_NLOS_CAPTURE_METER) {
        if (desc->sensor.shape < 0 || (uint32_t) desc->sensor.shape >= ns || desc->shapes[desc->sensor.shape].type != TR_SHAPE_RECTANGLE) {
            err = "nlos_capture_meter must be attached to a rectangle shape"; return false;
        }
    } else if (desc->sensor.type != TR_SENSOR_PERSPECTIVE) { err = "unknown sensor type"; return false; }
    return true;
}

}  // namespace tr

extern "C" {

int tr_scene_create(const TrSceneDesc *desc, TrScene **out_scene) {
    if (!out_scene) return tr::fail(TR_ERR_INVALID_ARGUMENT, "out_scene is NULL");
    *out_scene = nullptr;
    int rc = tr_device_check();
    if (rc != TR_OK) return rc;
    TrScene *s = new (std::nothrow) TrScene();
    if (!s) return tr::fail(TR_ERR_OUT_OF_MEMORY, "host allocation failed");
    std::string err;
    if (!tr::pack_scene(desc, s->impl, err)) { delete s; return tr::fail(TR_ERR_INVALID_ARGUMENT, err); }
    TR_CUDA_OR(cudaGetDevice(&s->impl.device), { delete s; });
    TR_CUDA_OR(cudaMalloc(&s->impl.d_blob, s->impl.blob.size()), { delete s; });
    TR_CUDA_OR(cudaMemcpy(s->impl.d_blob, s->impl.blob.data(), s->impl.blob.size(), cudaMemcpyHostToDevice),
               { cudaFree(s->impl.d_blob); delete s; });
    TR_CUDA_OR(cudaMalloc(&s->impl.d_scratch, 4096), { cudaFree(s->impl.d_blob); delete s; });
    TR_CUDA_OR(cudaHostAlloc(&s->impl.h_pinned, s->impl.blob.size(), cudaHostAllocDefault), { cudaFree(s->impl.d_scratch); cudaFree(s->impl.d_blob); delete s; });
    memcpy(s->impl.h_pinned, s->impl.blob.data(), s->impl.blob.size());
    if (!s->impl.frequencies.empty()) {
        const size_t fb = s->impl.frequencies.size() * sizeof(float);
        TR_CUDA_OR(cudaMalloc(&s->impl.d_freqs, fb), { cudaFree(s->impl.d_scratch); cudaFree(s->impl.d_blob); delete s; });
        TR_CUDA_OR(cudaMemcpy(s->impl.d_freqs, s->impl.frequencies.data(), fb, cudaMemcpyHostToDevice),
                   { cudaFree(s->impl.d_freqs); cudaFree(s->impl.d_scratch); cudaFree(s->impl.d_blob); delete s; });
    }
    *out_scene = s;
    return TR_OK;
}

void tr_scene_destroy(TrScene *scene) {
    if (!scene) return;
    if (scene->impl.d_blob) cudaFree(scene->impl.d_blob);
    if (scene->impl.d_pixels) cudaFree(scene->impl.d_pixels);
    if (scene->impl.d_scratch) cudaFree(scene->impl.d_scratch);
    if (scene->impl.d_nlos) cudaFree(scene->impl.d_nlos);
    if (scene->impl.d_freqs) cudaFree(scene->impl.d_freqs);
    if (scene->impl.h_pinned) cudaFreeHost(scene->impl.h_pinned);
    delete scene;
}

int tr_scene_refresh(TrScene *scene, void *stream) {
    if (!scene) return tr::fail(TR_ERR_INVALID_ARGUMENT, "scene is NULL");
    TR_CUDA(cudaMemcpyAsync(scene->impl.d_blob, scene->impl.h_pinned, scene->impl.blob.size(), cudaMemcpyHostToDevice,
                            (cudaStream_t) stream));
    return TR_OK;
}

int tr_scene_info(const TrScene *scene, uint64_t *blob_bytes, uint32_t *n_nodes, uint32_t *n_prims) {
    if (!scene) return tr::fail(TR_ERR_INVALID_ARGUMENT, "scene is NULL");
    if (blob_bytes) *blob_bytes = scene->impl.blob.size();
    if (n_nodes) *n_nodes = scene->impl.n_nodes;
    if (n_prims) *n_prims = scene->impl.n_prims;
    return TR_OK;
}

}  // extern "C"
