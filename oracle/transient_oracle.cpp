// ===========================================================================================
//  oracle/transient_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
//  A scalar, line-by-line restatement of mitransient's transient path-tracing hot path.  It
//  is the checker the CUDA kernels are compared against; it is NOT part of the product:
//  only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
//  may load it.  Nothing under mitransient_b200/ links, imports or calls it.
//
//  PARITY STATUS: pinned against the reference's own code (executed), down to the Mitsuba primitives.
//   - PINNED, film half: add_transient_data, TransientImageBlock.put_/accum (plain and exhaustive_scan index),
//     develop_transient_, the phasor film, the benchmark scene dictionary, parameter parsing, pass split.
//   - PINNED, integrator loops: TransientADIntegrator.render, TransientPath.sample, TransientNLOSPath.prepare /
//     sample / emitter_*_sample / hidden_geometry_sample, NLOSCaptureMeter, TransientPRBVolpathIntegrator.sample /
//     sample_emitter / first_surface and AngularAreaLight run UNMODIFIED over a NumPy stand-in for drjit/mitsuba
//     (tests/golden/refshim.py) whose scene objects call this file's osv_* exports; every deposit of
//     trace_transient_path / trace_nlos_path / trace_volpath must equal one of theirs bit for bit
//     (tests/golden/make_reference_golden.py, tests/test_reference_golden.py).
//   - UNPINNED: the Mitsuba primitives themselves (ray casting, BSDF / emitter / medium / sensor sampling, the
//     sampler).  Their arithmetic lives in two third-party, un-vendored dependencies -- mitsuba
//     (pinned >=3.6.0,<3.9.0, mitransient/version.py:5-7, setup.py:21) and drjit (1.1.0 in the notebooks) --
//     which are not installable here (no wheel, no network), and the reference's only test is stale and
//     shape-only (tests/integration/test_nlos.py:81-118).  They restate the published Mitsuba 3 algorithms and
//     are validated by the analytic known-answer tests in tests/test_oracle_kat.py.
//
//  What is restated (file:line relative to the mitransient tree):
//    render driver        mitransient/integrators/common.py:122-213   -> oracle_render
//    sample_rays          mitsuba ADIntegrator.sample_rays [upstream] -> sample_camera_ray
//    TransientPath.sample mitransient/integrators/transientpath.py:109-326 -> trace_path
//    add_transient_f      mitransient/integrators/common.py:411-422   -> Film::add_transient
//    add_transient_data   mitransient/films/transient_hdr_film.py:250-276
//    put_/accum           mitransient/render/transient_image_block.py:130-149, 79-81
//    develop              mitransient/films/transient_hdr_film.py:220-248 -> oracle_develop
//  Upstream (Mitsuba 3 / Dr.Jit) pieces restated from their published source: PCG32 + TEA
//  seeding of the `independent` sampler, perspective sensor, `diffuse` BSDF, `area` emitter,
//  Scene::sample_emitter_direction / pdf_emitter_direction, rectangle / mesh position
//  sampling, SurfaceInteraction::spawn_ray(_to), mis_weight (power heuristic).
//
//  ARITHMETIC CONTRACT (shared with the CUDA kernels by *specification*, not by code): all
//  maths is IEEE fp32, no implicit contraction (-ffp-contract=off); fused multiply-adds are
//  written out as fmaf() exactly where the contract says so:
//      dot(a,b)   = fmaf(a.z,b.z, fmaf(a.y,b.y, a.x*b.x))
//      cross(a,b) = ( fmaf(a.y,b.z, -(a.z*b.y)), fmaf(a.z,b.x, -(a.x*b.z)), fmaf(a.x,b.y, -(a.y*b.x)) )
//      v / s      = v * (1.0f / s)            normalize(v) = v * (1.0f / sqrtf(dot(v,v)))
//  Ray / primitive intersection (the reference delegates it to Embree/OptiX, whose bits cannot be
//  reproduced): plane form with per-primitive precomputed duals, see Prim / prim_test below; the one
//  division of a ray / primitive test is t = num * rcp_c(den) with rcp_c(x) = the correctly rounded 1/x for
//  |x| >= 2^-100 and +-inf below (a ray parallel to the plane / slab); axis-aligned rectangles are their own
//  primitive class (slab arithmetic with the ray's three reciprocals rcp_c(d.x|y|z), aa_rect_test).
//  Everywhere else sqrt and division are correctly rounded; sin/cos use the Cephes polynomial below.  With
//  the same PCG32 streams the GPU then walks bit-identical paths; only the order of the fp32
//  additions into a film bin differs.
// ===========================================================================================
#include "../include/mitransient_b200.h"

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------------------ vectors
struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
inline V3 fma3(V3 a, float s, V3 c) { return {fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z)}; }
inline V3 normalize(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline float max3(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
inline float mulsign(float a, float b) {  // a with the sign of b multiplied in (drjit mulsign)
    uint32_t ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
    ua ^= (ub & 0x80000000u); float r; memcpy(&r, &ua, 4); return r;
}

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kRayEpsilon = 8.940696716308594e-05f;     // Mitsuba math::RayEpsilon<float> = 2^-24 * 1500
constexpr float kShadowEpsilon = 8.940696716308594e-04f;  // RayEpsilon * 10
constexpr float kInf = std::numeric_limits<float>::infinity();

// Cephes single-precision sincos (the polynomial Dr.Jit uses), valid for |x| < 8192.
inline void sincos_cephes(float x, float *s_out, float *c_out) {
    float xa = fabsf(x);
    int j = (int) (xa * 1.2732395447351628f);  // 4/pi
    j = (j + 1) & ~1;
    float y = (float) j;
    // Cody-Waite range reduction
    float z = fmaf(y, -0.78515625f, xa);
    z = fmaf(y, -2.4187564849853515625e-4f, z);
    z = fmaf(y, -3.77489497744594108e-8f, z);
    float z2 = z * z;
    float sp = fmaf(z2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, z2, -1.6666654611e-1f);
    sp = fmaf(sp * z2, z, z);
    float cp = fmaf(z2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, z2, 4.166664568298827e-2f);
    cp = fmaf(cp * z2, z2, fmaf(z2, -0.5f, 1.0f));
    bool swap = (j & 2) != 0;
    float s = swap ? cp : sp, c = swap ? sp : cp;
    if (j & 4) s = -s;
    if ((j + 2) & 4) c = -c;
    if (x < 0.0f) s = -s;
    *s_out = s; *c_out = c;
}

// Cephes single-precision exp / log (explicit fmaf; shared by specification with the CUDA kernels).
inline float exp_cephes(float x) {
    x = fminf(fmaxf(x, -87.0f), 88.0f);
    float n = floorf(fmaf(1.44269504088896341f, x, 0.5f));
    x = fmaf(n, -0.693359375f, x);
    x = fmaf(n, 2.12194440e-4f, x);
    float z = x * x;
    float p = fmaf(1.9875691500e-4f, x, 1.3981999507e-3f);
    p = fmaf(p, x, 8.3334519073e-3f);
    p = fmaf(p, x, 4.1665795894e-2f);
    p = fmaf(p, x, 1.6666665459e-1f);
    p = fmaf(p, x, 5.0000001201e-1f);
    p = fmaf(p, z, x) + 1.0f;
    uint32_t bits = (uint32_t) ((int32_t) n + 127) << 23;      // 2^n, n in [-126, 127]
    float scale; memcpy(&scale, &bits, 4);
    return p * scale;
}
inline float log_cephes(float x) {          // x > 0, normal
    uint32_t bits; memcpy(&bits, &x, 4);
    int32_t e = (int32_t) ((bits >> 23) & 0xffu) - 126;        // frexp: x = m * 2^e, m in [0.5, 1)
    bits = (bits & 0x007fffffu) | 0x3f000000u;
    float m; memcpy(&m, &bits, 4);
    if (m < 0.707106781186547524f) { e -= 1; m = (m + m) - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = fmaf(7.0376836292e-2f, m, -1.1514610310e-1f);
    y = fmaf(y, m, 1.1676998740e-1f);
    y = fmaf(y, m, -1.2420140846e-1f);
    y = fmaf(y, m, 1.4249322787e-1f);
    y = fmaf(y, m, -1.6668057665e-1f);
    y = fmaf(y, m, 2.0000714765e-1f);
    y = fmaf(y, m, -2.4999993993e-1f);
    y = fmaf(y, m, 3.3333331174e-1f);
    y = (y * m) * z;
    float fe = (float) e;
    y = fmaf(-2.12194440e-4f, fe, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(0.693359375f, fe, r);
}

// ------------------------------------------------------------------------------ sampler
// Mitsuba `independent` sampler: one PCG32 per lane, seeded through the Tiny Encryption
// Algorithm from (seed, lane index) [upstream: Sampler::seed, sample_tea_32, drjit PCG32].
struct Pcg32 {
    uint64_t state, inc;
    void seed(uint64_t initstate, uint64_t initseq) {
        state = 0; inc = (initseq << 1) | 1u; next_u32(); state += initstate; next_u32();
    }
    uint32_t next_u32() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    float next_1d() {
        uint32_t u = (next_u32() >> 9) | 0x3f800000u; float f; memcpy(&f, &u, 4); return f - 1.0f;
    }
};
inline void sample_tea_32(uint32_t v0, uint32_t v1, uint32_t *o0, uint32_t *o1) {
    uint32_t sum = 0;
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    *o0 = v0; *o1 = v1;
}

// ------------------------------------------------------------------------------ scene
// A planar primitive (triangle or parallelogram) in intersection form:
//   N = cross(e1,e2), d0 = dot(N,p0), a1 = cross(e2,N)/|N|^2, c1 = dot(p0,a1), a2 = cross(N,e1)/|N|^2, c2 = dot(p0,a2)
// so that for a point p of the plane  u = dot(p,a1) - c1,  v = dot(p,a2) - c2  (p = p0 + u e1 + v e2).
struct Prim {
    V3 p0, e1, e2, N, a1, a2, n, s, t; float d0, c1, c2, area; uint32_t shape; bool quad; bool boundary;
    // TR_PRIM_BOX: solid parallelepiped p0 + u e1 + v e2 + w e3; (a1,c1),(a2,c2),(a3,c3) are the duals giving (u,v,w)
    bool box; V3 e3, a3; float c3, inv_len[3], face_area[3];
    // axis-aligned rectangle (a quad whose two edges run along coordinate axes): normal axis (-1: not one), plane
    // coordinate, and the extent along the two in-plane axes (aa_axis + 1) % 3, (aa_axis + 2) % 3
    int aa_axis; float aa_c, aa_lo[2], aa_hi[2];
};
struct Shape {
    TrShapeDesc d;
    float inv_area;            // 1 / surface area
    V3 n;                      // rectangle: frame normal
    std::vector<float> pmf;    // mesh: primitive areas (unnormalised pmf)
    std::vector<float> cdf;    // mesh: running area sums (float, accumulated in double)
    float area_sum, area_norm; // DiscreteDistribution m_sum / m_normalization
    float area;                // Shape::surface_area()
};
struct Scene {
    std::vector<Prim> prims;
    std::vector<Shape> shapes;
    std::vector<TrBsdfDesc> bsdfs;
    std::vector<TrEmitterDesc> emitters;
    std::vector<TrMediumDesc> media;
    TrSensorDesc sensor;
    TrFilmDesc film;
};

// Duff et al. branchless orthonormal basis (Mitsuba coordinate_system).
inline void coordinate_system(V3 n, V3 *s, V3 *t) {
    float sign = copysignf(1.0f, n.z);
    float a = -(1.0f / (sign + n.z));
    float b = n.x * n.y * a;
    *s = v3(mulsign(n.x * n.x * a, n.z) + 1.0f, mulsign(b, n.z), mulsign(-n.x, n.z));
    *t = v3(b, fmaf(n.y, n.y * a, sign), -n.y);
}

inline V3 xf_point(const float *m, float x, float y, float z) {  // row-major 3x4 affine
    return v3(fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3]))),
              fmaf(m[4], x, fmaf(m[5], y, fmaf(m[6], z, m[7]))),
              fmaf(m[8], x, fmaf(m[9], y, fmaf(m[10], z, m[11]))));
}
inline V3 xf_vector(const float *m, V3 v) {
    return v3(fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)),
              fmaf(m[4], v.x, fmaf(m[5], v.y, m[6] * v.z)),
              fmaf(m[8], v.x, fmaf(m[9], v.y, m[10] * v.z)));
}

// A primitive is a BOUNDARY primitive when every vertex of the scene lies on the front side of (or within a
// tolerance of) its supporting plane.  A segment between two scene points can then only touch that plane at
// its end points, so a boundary primitive never occludes a shadow ray: scene.ray_test skips it (this removes
// the fp self-shadowing flicker at room corners and a third of the occlusion tests of a closed room).
inline float plane_side(const Prim &pr, V3 x) { return fmaf(pr.N.z, x.z, fmaf(pr.N.y, x.y, fmaf(pr.N.x, x.x, -pr.d0))); }
void classify_boundary_prims(Scene &sc) {
    float lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
    std::vector<V3> verts;
    for (const Prim &q : sc.prims) {
        V3 c[8] = {q.p0, q.p0 + q.e1, q.p0 + q.e2, (q.p0 + q.e1) + q.e2,
                   q.p0 + q.e3, (q.p0 + q.e1) + q.e3, (q.p0 + q.e2) + q.e3, ((q.p0 + q.e1) + q.e2) + q.e3};
        for (int k = 0; k < (q.box ? 8 : (q.quad ? 4 : 3)); ++k) {
            verts.push_back(c[k]);
            lo[0] = fminf(lo[0], c[k].x); hi[0] = fmaxf(hi[0], c[k].x);
            lo[1] = fminf(lo[1], c[k].y); hi[1] = fmaxf(hi[1], c[k].y);
            lo[2] = fminf(lo[2], c[k].z); hi[2] = fmaxf(hi[2], c[k].z);
        }
    }
    // end points of shadow rays that are not scene vertices: the origin of every emitter that is not attached to a
    // shape (the projector of transient_nlos_path: emitter_nee_sample tests visibility towards it) and of the sensor
    for (const TrEmitterDesc &e : sc.emitters)
        if (e.type == TR_EMITTER_PROJECTOR) verts.push_back(v3(e.to_world[3], e.to_world[7], e.to_world[11]));
    if (sc.sensor.type == TR_SENSOR_NLOS_CAPTURE_METER) verts.push_back(v3(sc.sensor.origin[0], sc.sensor.origin[1], sc.sensor.origin[2]));
    else verts.push_back(v3(sc.sensor.to_world[3], sc.sensor.to_world[7], sc.sensor.to_world[11]));
    float extent = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    for (Prim &pr : sc.prims) {
        float tol = 1e-4f * sqrtf(dot(pr.N, pr.N)) * extent;
        bool b = !pr.box;      // a solid box is never a boundary
        if (b) for (const V3 &x : verts) if (!(plane_side(pr, x) >= -tol)) { b = false; break; }
        pr.boundary = b;
    }
}

bool build_scene(const TrSceneDesc *desc, Scene &sc, std::string &err) {
    if (!desc || desc->abi_version != TR_ABI_VERSION) { err = "bad scene descriptor / ABI version"; return false; }
    sc.sensor = desc->sensor; sc.film = desc->film;
    sc.bsdfs.assign(desc->bsdfs, desc->bsdfs + desc->n_bsdfs);
    sc.emitters.assign(desc->emitters, desc->emitters + desc->n_emitters);
    if (desc->n_media) sc.media.assign(desc->media, desc->media + desc->n_media);
    sc.prims.resize(desc->n_prims);
    sc.shapes.resize(desc->n_shapes);
    for (uint32_t si = 0; si < desc->n_shapes; ++si) {
        Shape &sh = sc.shapes[si];
        sh.d = desc->shapes[si];
        if (sh.d.prim_begin + sh.d.prim_count > desc->n_prims) { err = "shape primitive range out of bounds"; return false; }
        double area_acc = 0.0;
        sh.cdf.clear(); sh.pmf.clear();
        for (uint32_t k = 0; k < sh.d.prim_count; ++k) {
            uint32_t pi = sh.d.prim_begin + k;
            Prim &t = sc.prims[pi];
            const float *v = desc->prim_verts + 12 * (size_t) pi;
            t.quad = desc->prim_types[pi] == TR_PRIM_QUAD;
            t.box = desc->prim_types[pi] == TR_PRIM_BOX;
            t.p0 = v3(v[0], v[1], v[2]);
            t.e1 = v3(v[3], v[4], v[5]) - t.p0;
            t.e2 = v3(v[6], v[7], v[8]) - t.p0;
            t.e3 = v3(v[9], v[10], v[11]) - t.p0;
            t.N = cross(t.e1, t.e2);
            float nn = dot(t.N, t.N);
            float inv_nn = 1.0f / nn;
            float len = sqrtf(nn);
            t.n = t.N * (1.0f / len);
            t.d0 = dot(t.N, t.p0);
            if (t.box) {
                // duals of the parallelepiped: u = a1.(x - p0), v = a2.(x - p0), w = a3.(x - p0)
                V3 c23 = cross(t.e2, t.e3), c31 = cross(t.e3, t.e1), c12 = t.N;
                float inv_det = 1.0f / dot(t.e1, c23);
                t.a1 = c23 * inv_det; t.c1 = dot(t.p0, t.a1);
                t.a2 = c31 * inv_det; t.c2 = dot(t.p0, t.a2);
                t.a3 = c12 * inv_det; t.c3 = dot(t.p0, t.a3);
                t.inv_len[0] = 1.0f / sqrtf(dot(t.a1, t.a1)); t.inv_len[1] = 1.0f / sqrtf(dot(t.a2, t.a2));
                t.inv_len[2] = 1.0f / sqrtf(dot(t.a3, t.a3));
                t.face_area[0] = sqrtf(dot(c23, c23)); t.face_area[1] = sqrtf(dot(c31, c31)); t.face_area[2] = len;
            } else {
                t.a1 = cross(t.e2, t.N) * inv_nn; t.c1 = dot(t.p0, t.a1);
                t.a2 = cross(t.N, t.e1) * inv_nn; t.c2 = dot(t.p0, t.a2);
                t.a3 = v3(0, 0, 0); t.c3 = 0.0f;
            }
            t.aa_axis = -1; t.aa_c = 0.0f; t.aa_lo[0] = t.aa_lo[1] = t.aa_hi[0] = t.aa_hi[1] = 0.0f;
            if (t.quad && !t.box) {
                // both edges along a coordinate axis (exactly): an axis-aligned rectangle
                const float *v0 = v, *v1 = v + 3, *v2 = v + 6;
                int ax1 = -1, ax2 = -1, n1 = 0, n2 = 0;
                for (int c = 0; c < 3; ++c) { if (v1[c] != v0[c]) { ax1 = c; n1++; } if (v2[c] != v0[c]) { ax2 = c; n2++; } }
                if (n1 == 1 && n2 == 1 && ax1 != ax2) {
                    const int k = 3 - ax1 - ax2;
                    t.aa_axis = k; t.aa_c = v0[k];
                    for (int j = 0; j < 2; ++j) {
                        const int a = (k + 1 + j) % 3;
                        t.aa_lo[j] = fminf(fminf(v0[a], v1[a]), v2[a]); t.aa_hi[j] = fmaxf(fmaxf(v0[a], v1[a]), v2[a]);
                    }
                }
            }
            t.shape = si;
            t.area = t.box ? 2.0f * ((t.face_area[2] + t.face_area[0]) + t.face_area[1]) : (t.quad ? len : 0.5f * len);
            area_acc += (double) t.area;
            sh.pmf.push_back(t.area);
            sh.cdf.push_back((float) area_acc);
            if (sh.d.type == TR_SHAPE_RECTANGLE) {
                // Rectangle: tangent from dp_du = to_world * (2,0,0), Gram-Schmidt against n
                // (Mitsuba SurfaceInteraction::initialize_sh_frame).
                V3 dpdu = xf_vector(sh.d.to_world, v3(2.0f, 0.0f, 0.0f));
                float d = dot(t.n, dpdu);
                t.s = normalize(v3(fmaf(-t.n.x, d, dpdu.x), fmaf(-t.n.y, d, dpdu.y), fmaf(-t.n.z, d, dpdu.z)));
            } else {
                V3 tt; coordinate_system(t.n, &t.s, &tt);
            }
            t.t = cross(t.n, t.s);
        }
        sh.area_sum = (float) area_acc;
        sh.area_norm = (float) (1.0 / area_acc);
        if (sh.d.type == TR_SHAPE_RECTANGLE) {
            V3 dpdu = xf_vector(sh.d.to_world, v3(2.0f, 0.0f, 0.0f));
            V3 dpdv = xf_vector(sh.d.to_world, v3(0.0f, 2.0f, 0.0f));
            V3 c = cross(dpdu, dpdv);
            sh.area = sqrtf(dot(c, c));
            sh.inv_area = 1.0f / sh.area;
            sh.n = sh.d.prim_count ? sc.prims[sh.d.prim_begin].n : v3(0, 0, 1);
        } else {
            sh.area = sh.area_sum;
            sh.inv_area = sh.area_norm;
            sh.n = v3(0, 0, 0);
        }
    }
    classify_boundary_prims(sc);
    return true;
}

// ------------------------------------------------------------------------------ rays
struct Ray { V3 o, d; float maxt; };
struct Hit {
    bool valid; float t; V3 p, n, s, tt; uint32_t shape, prim;
};

// Reciprocal of the arithmetic contract: the correctly rounded 1/x for |x| >= 2^-100; below that (a ray parallel to
// the plane / slab, up to rounding) +-inf with the sign of x.  (The kernel: MUFU.RCP + one Newton step, which is the
// correctly rounded reciprocal on that range -- checked exhaustively on the GPU, tests/test_gpu_parity.py.)
constexpr float kRcpMin = 7.888609052210118e-31f;   // 2^-100
inline float rcp_c(float x) { return fabsf(x) >= kRcpMin ? 1.0f / x : copysignf(kInf, x); }

// Ray / primitive test of the arithmetic contract: intersect the supporting plane, then read the
// parallelogram coordinates (u, v) of the hit point off the precomputed dual vectors.
inline bool prim_test(const Prim &pr, const Ray &r, float *t_out, V3 *p_out) {
    float denom = dot(pr.N, r.d);
    float tnum = fmaf(-pr.N.z, r.o.z, fmaf(-pr.N.y, r.o.y, fmaf(-pr.N.x, r.o.x, pr.d0)));
    float t = tnum * rcp_c(denom);
    V3 p = fma3(r.d, t, r.o);
    float u = fmaf(p.z, pr.a1.z, fmaf(p.y, pr.a1.y, fmaf(p.x, pr.a1.x, -pr.c1)));
    float v = fmaf(p.z, pr.a2.z, fmaf(p.y, pr.a2.y, fmaf(p.x, pr.a2.x, -pr.c2)));
    bool inside = pr.quad ? (u <= 1.0f && v <= 1.0f) : (u + v <= 1.0f);
    bool ok = (u >= 0.0f) && (v >= 0.0f) && inside && (t >= 0.0f) && (t <= r.maxt);
    *t_out = t; *p_out = p;
    return ok;
}

// Axis-aligned rectangle (normal along axis k, plane x_k = c): slab arithmetic with the ray's reciprocal direction
// inv_d[k] = rcp_c(d[k]), then the hit point's two in-plane coordinates against the rectangle's extent.
inline bool aa_rect_test(const Prim &pr, const Ray &r, const float inv_d[3], float *t_out) {
    const float o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    const int k = pr.aa_axis, a = (k + 1) % 3, b = (k + 2) % 3;
    float t = (pr.aa_c - o[k]) * inv_d[k];
    float pa = fmaf(t, d[a], o[a]), pb = fmaf(t, d[b], o[b]);
    *t_out = t;
    return (pa >= pr.aa_lo[0]) && (pa <= pr.aa_hi[0]) && (pb >= pr.aa_lo[1]) && (pb <= pr.aa_hi[1]) && (t >= 0.0f) && (t <= r.maxt);
}

// Ray / box test of the arithmetic contract: three slabs in the parallelepiped's (u, v, w) coordinates.  A ray
// that starts inside leaves through the far face.  *face = axis * 2 + (outward normal along +a_axis ? 1 : 0).
inline bool box_test(const Prim &pr, const Ray &r, float *t_out, int *face) {
    const V3 a[3] = {pr.a1, pr.a2, pr.a3}; const float c[3] = {pr.c1, pr.c2, pr.c3};
    float tmin[3], tmax[3], ud[3];
    for (int k = 0; k < 3; ++k) {
        float uo = fmaf(a[k].z, r.o.z, fmaf(a[k].y, r.o.y, fmaf(a[k].x, r.o.x, -c[k])));
        ud[k] = dot(a[k], r.d);
        float inv = rcp_c(ud[k]);
        float t0 = (-uo) * inv, t1 = (1.0f - uo) * inv;
        tmin[k] = fminf(t0, t1); tmax[k] = fmaxf(t0, t1);
    }
    float t_near = fmaxf(fmaxf(tmin[0], tmin[1]), tmin[2]), t_far = fminf(fminf(tmax[0], tmax[1]), tmax[2]);
    bool entering = t_near >= 0.0f;
    int k;
    if (entering) k = (tmin[0] >= tmin[1] && tmin[0] >= tmin[2]) ? 0 : (tmin[1] >= tmin[2] ? 1 : 2);
    else k = (tmax[0] <= tmax[1] && tmax[0] <= tmax[2]) ? 0 : (tmax[1] <= tmax[2] ? 1 : 2);
    bool plus = entering ? !(ud[k] > 0.0f) : (ud[k] > 0.0f);
    *face = k * 2 + (plus ? 1 : 0);
    *t_out = entering ? t_near : t_far;
    return (t_near <= t_far) && (t_far >= 0.0f) && (*t_out <= r.maxt);
}
inline V3 box_face_normal(const Prim &pr, int face) {
    const V3 a = (face >> 1) == 0 ? pr.a1 : ((face >> 1) == 1 ? pr.a2 : pr.a3);
    float il = pr.inv_len[face >> 1];
    return (face & 1) ? a * il : (-a) * il;
}

// Test order of the contract (the first primitive tested wins a tie in t): axis-aligned rectangles by normal axis
// (x, then y, then z), then the other planar primitives, then boxes; ascending primitive index inside each class.
inline int prim_class(const Prim &pr) { return pr.box ? 4 : (pr.aa_axis >= 0 ? pr.aa_axis : 3); }

Hit ray_intersect(const Scene &sc, const Ray &r) {  // scene.ray_intersect: closest hit, brute force
    Hit h; h.valid = false; h.t = kInf; h.prim = 0xffffffffu; h.shape = 0;
    h.p = v3(0, 0, 0); h.n = v3(0, 0, 0); h.s = h.tt = v3(0, 0, 0);
    int best_face = 0;
    const float inv_d[3] = {rcp_c(r.d.x), rcp_c(r.d.y), rcp_c(r.d.z)};
    for (int pass = 0; pass < 5; ++pass)
        for (uint32_t i = 0; i < sc.prims.size(); ++i) {
            const Prim &pr = sc.prims[i];
            if (prim_class(pr) != pass) continue;
            float t; V3 p; int face = 0;
            bool ok = pr.box ? box_test(pr, r, &t, &face) : (pr.aa_axis >= 0 ? aa_rect_test(pr, r, inv_d, &t) : prim_test(pr, r, &t, &p));
            if (ok && t < h.t) { h.valid = true; h.t = t; h.prim = i; best_face = face; }
        }
    if (h.valid) {
        const Prim &pr = sc.prims[h.prim];
        h.p = fma3(r.d, h.t, r.o);
        h.shape = pr.shape;
        if (pr.box) { h.n = box_face_normal(pr, best_face); V3 tt; coordinate_system(h.n, &h.s, &tt); h.tt = cross(h.n, h.s); }
        else { h.n = pr.n; h.s = pr.s; h.tt = pr.t; }
    }
    return h;
}

bool ray_test(const Scene &sc, const Ray &r) {  // scene.ray_test: any hit
    const float inv_d[3] = {rcp_c(r.d.x), rcp_c(r.d.y), rcp_c(r.d.z)};
    for (uint32_t i = 0; i < sc.prims.size(); ++i) {
        const Prim &pr = sc.prims[i];
        if (pr.boundary) continue;   // see classify_boundary_prims
        float t; V3 p; int face;
        if (pr.box ? box_test(pr, r, &t, &face) : (pr.aa_axis >= 0 ? aa_rect_test(pr, r, inv_d, &t) : prim_test(pr, r, &t, &p))) return true;
    }
    return false;
}

// SurfaceInteraction::offset_p / spawn_ray / spawn_ray_to [upstream interaction.h]
inline V3 offset_p(V3 p, V3 n, V3 dir) {
    float mag = (1.0f + fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * kRayEpsilon;
    mag = mulsign(mag, dot(n, dir));
    return fma3(n, mag, p);
}
inline Ray spawn_ray(V3 p, V3 n, V3 d) { return Ray{offset_p(p, n, d), d, std::numeric_limits<float>::max()}; }
inline Ray spawn_ray_to(V3 p, V3 n, V3 target) {
    V3 o = offset_p(p, n, target - p);
    V3 d = target - o;
    float dist = sqrtf(dot(d, d));
    d = d * (1.0f / dist);
    return Ray{o, d, dist * (1.0f - kShadowEpsilon)};
}

// ------------------------------------------------------------------------------ warps
inline void square_to_uniform_disk_concentric(float sx, float sy, float *ox, float *oy) {
    float x = fmaf(2.0f, sx, -1.0f), y = fmaf(2.0f, sy, -1.0f);
    bool is_zero = (x == 0.0f) && (y == 0.0f);
    bool quadrant_1_or_3 = fabsf(x) < fabsf(y);
    float r = quadrant_1_or_3 ? y : x, rp = quadrant_1_or_3 ? x : y;
    float phi = 0.25f * kPi * rp * (1.0f / r);
    if (quadrant_1_or_3) phi = 0.5f * kPi - phi;
    if (is_zero) phi = 0.0f;
    float s, c; sincos_cephes(phi, &s, &c);
    *ox = r * c; *oy = r * s;
}
inline V3 square_to_cosine_hemisphere(float sx, float sy) {
    float px, py; square_to_uniform_disk_concentric(sx, sy, &px, &py);
    float z2 = 1.0f - fmaf(py, py, px * px);
    float z = sqrtf(fmaxf(z2, 0.0f));  // safe_sqrt
    return v3(px, py, z);
}

inline float mis_weight(float pdf_a, float pdf_b) {  // mitsuba.ad.integrators.common.mis_weight
    float a2 = pdf_a * pdf_a, b2 = pdf_b * pdf_b;
    float w = a2 / (a2 + b2);
    return std::isfinite(w) ? w : 0.0f;
}

// ------------------------------------------------------------------------------ emitters
struct DirectionSample { V3 p, n, d; float dist, pdf; bool delta; };

// Shape::sample_position for the two shape kinds [upstream rectangle.cpp / mesh.cpp]
inline void shape_sample_position(const Scene &sc, const Shape &sh, float u0, float u1, V3 *p, V3 *n, float *pdf) {
    if (sh.d.type == TR_SHAPE_RECTANGLE) {
        *p = xf_point(sh.d.to_world, fmaf(u0, 2.0f, -1.0f), fmaf(u1, 2.0f, -1.0f), 0.0f);
        *n = sh.n; *pdf = sh.inv_area;
    } else {
        // Mesh::sample_position: face by area CDF reusing sample.y, then a uniform point on the
        // primitive (triangle: Mitsuba's square_to_uniform_triangle; quad: p0 + u e1 + v e2)
        float value = u1 * sh.area_sum;
        uint32_t lo = 0, hi = sh.d.prim_count ? sh.d.prim_count - 1 : 0;  // first idx with cdf[idx] >= value
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (sh.cdf[mid] < value) lo = mid + 1; else hi = mid; }
        uint32_t idx = lo;
        float cdf_prev = idx ? sh.cdf[idx - 1] : 0.0f;
        float pmf = sh.pmf[idx] * sh.area_norm;
        float u1r = (u1 - cdf_prev * sh.area_norm) / pmf;
        const Prim &tr = sc.prims[sh.d.prim_begin + idx];
        if (tr.box) {
            // a face by area (reusing u0), then a uniform point on it; faces in the order u=0, u=1, v=0, v=1, w=0, w=1
            const float fa[6] = {tr.face_area[0], tr.face_area[0], tr.face_area[1], tr.face_area[1], tr.face_area[2], tr.face_area[2]};
            float x = u0 * tr.area;
            float cdf[5]; cdf[0] = fa[0];
            for (int k = 1; k < 5; ++k) cdf[k] = cdf[k - 1] + fa[k];
            int f = x < cdf[0] ? 0 : (x < cdf[1] ? 1 : (x < cdf[2] ? 2 : (x < cdf[3] ? 3 : (x < cdf[4] ? 4 : 5))));
            float prev = f == 0 ? 0.0f : cdf[f - 1];
            float s0 = fminf(fmaxf((x - prev) / fa[f], 0.0f), 1.0f);
            const int axis = f >> 1; const bool far_side = (f & 1) != 0;
            const V3 ea = axis == 0 ? tr.e2 : (axis == 1 ? tr.e3 : tr.e1), eb = axis == 0 ? tr.e3 : (axis == 1 ? tr.e1 : tr.e2);
            const V3 en = axis == 0 ? tr.e1 : (axis == 1 ? tr.e2 : tr.e3);
            V3 base = far_side ? tr.p0 + en : tr.p0;
            *p = fma3(ea, s0, fma3(eb, u1r, base));
            *n = box_face_normal(tr, f);
            *pdf = sh.area_norm;
            return;
        }
        float b0, b1;
        if (tr.quad) { b0 = u0; b1 = u1r; }
        else { float t = sqrtf(fmaxf(1.0f - u0, 0.0f)); b0 = 1.0f - t; b1 = t * u1r; }
        *p = fma3(tr.e1, b0, fma3(tr.e2, b1, tr.p0));
        *n = tr.n; *pdf = sh.area_norm;
    }
}

// acos for the angulararea falloff (Cephes asinf/acosf polynomial; part of the arithmetic contract like sincos_cephes)
inline float asin_cephes(float x) {
    float a = fabsf(x), z, w; bool big = a > 0.5f;
    if (big) { z = 0.5f * (1.0f - a); w = sqrtf(z); } else { w = a; z = a * a; }
    float p = fmaf(4.2163199048e-2f, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f); p = fmaf(p, z, 7.4953002686e-2f); p = fmaf(p, z, 1.6666752422e-1f);
    float r = fmaf(p * z, w, w);
    if (big) r = 1.57079632679489661923f - (r + r);
    return x < 0.0f ? -r : r;
}
inline float acos_cephes(float x) {
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * asin_cephes(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * asin_cephes(sqrtf(0.5f * (1.0f - x)));
    return 1.57079632679489661923f - asin_cephes(x);
}
// AngularAreaLight._fallof_curve (mitransient/emitters/angulararea.py:73-81) on a direction in the emitter's local frame
struct AngularParams { float cutoff, cos_cutoff, cos_beam, inv_transition; };
inline AngularParams angular_params(const TrEmitterDesc &em) {                  // :57-67 (computed by the host in double, like the reference)
    AngularParams a; a.cutoff = em.cutoff_angle; a.inv_transition = em.inv_transition_width;
    a.cos_cutoff = em.cos_cutoff_angle; a.cos_beam = em.cos_beam_width;
    return a;
}
inline float angular_falloff(const TrEmitterDesc &em, V3 d) {
    AngularParams a = angular_params(em);
    V3 local_dir = normalize(d);                                               // :74
    float cos_theta = local_dir.z;
    float beam_res = cos_theta >= a.cos_beam ? 1.0f : (a.cutoff - acos_cephes(cos_theta)) * a.inv_transition;   // :76-79
    return cos_theta > a.cos_cutoff ? beam_res : 0.0f;                         // :81
}
// Frame3f(n).to_local(v) with Mitsuba's coordinate_system(n)
inline V3 frame_to_local(V3 n, V3 v) { V3 s, t; coordinate_system(n, &s, &t); return v3(dot(v, s), dot(v, t), dot(v, n)); }

// Emitter(area)::sample_direction -> Shape::sample_direction (area pdf -> solid angle).
// Returns the (unoccluded) weight radiance / pdf, or 0.
inline V3 area_sample_direction(const Scene &sc, const TrEmitterDesc &em, V3 ref_p, float u0, float u1, DirectionSample *ds) {
    const Shape &sh = sc.shapes[em.shape];
    shape_sample_position(sc, sh, u0, u1, &ds->p, &ds->n, &ds->pdf);
    V3 d = ds->p - ref_p;
    float dist2 = dot(d, d);
    ds->dist = sqrtf(dist2);
    ds->d = d * (1.0f / ds->dist);
    float dp = fabsf(dot(ds->d, ds->n));
    float x = dist2 / dp;
    ds->pdf *= std::isfinite(x) ? x : 0.0f;
    ds->delta = false;
    bool active = (dot(ds->d, ds->n) < 0.0f) && (ds->pdf != 0.0f);
    if (!active) return v3(0, 0, 0);
    if (em.type == TR_EMITTER_ANGULARAREA) {                                   // AngularAreaLight.sample_direction  angulararea.py:106-127
        V3 local_d = frame_to_local(ds->n, -ds->d);                            // :113-114
        float inv_dist = 1.0f / ds->dist;                                      // :115
        float falloff = angular_falloff(em, local_d);                          // :118
        if (!(falloff > 0.0f)) return v3(0, 0, 0);                             // :119
        float f = falloff * (inv_dist * inv_dist);                             // :122-123  radiance * (falloff * inv_dist^2) / ds.pdf
        V3 spec = v3(em.radiance[0], em.radiance[1], em.radiance[2]) * f;
        return v3(spec.x / ds->pdf, spec.y / ds->pdf, spec.z / ds->pdf);
    }
    float inv = 1.0f / ds->pdf;
    return v3(em.radiance[0], em.radiance[1], em.radiance[2]) * inv;
}
// Emitter::eval(si) on the emitter's surface: `area` = radiance on the front side; `angulararea` (:93-101) adds the falloff
inline V3 emitter_eval(const TrEmitterDesc &em, V3 wi) {
    if (!(wi.z > 0.0f)) return v3(0, 0, 0);
    V3 r = v3(em.radiance[0], em.radiance[1], em.radiance[2]);
    if (em.type == TR_EMITTER_ANGULARAREA) {
        float falloff = angular_falloff(em, wi);
        return falloff > 0.0f ? r * falloff : v3(0, 0, 0);
    }
    return r;
}

// Scene::sample_emitter_direction without the visibility test [upstream scene.cpp]: uniform emitter choice with
// sample reuse, Emitter::sample_direction, pdf scaled by the pick probability, weight by its inverse.
inline V3 scene_sample_emitter_direction(const Scene &sc, V3 ref_p, float u0, float u1, DirectionSample *ds, uint32_t *em_index) {
    const uint32_t n_em = (uint32_t) sc.emitters.size();
    uint32_t ei = 0; float emw = 1.0f, pdf_pick = 1.0f;
    if (n_em > 1) {
        float fi = u0 * (float) n_em;
        ei = (uint32_t) fi; if (ei > n_em - 1) ei = n_em - 1;
        u0 = fi - (float) ei; emw = (float) n_em; pdf_pick = 1.0f / (float) n_em;
    }
    const TrEmitterDesc &em = sc.emitters[ei];
    V3 em_weight = area_sample_direction(sc, em, ref_p, u0, u1, ds);
    ds->pdf *= pdf_pick; em_weight = em_weight * emw;
    *em_index = ei;
    return em_weight;
}

// Scene::pdf_emitter_direction for an area emitter hit by BSDF sampling.
inline float area_pdf_direction(const Scene &sc, const TrEmitterDesc &em, V3 d, V3 n_light, float dist) {
    float dp = dot(d, n_light);
    if (!(dp < 0.0f)) return 0.0f;
    if (em.type == TR_EMITTER_ANGULARAREA && !(angular_falloff(em, frame_to_local(n_light, -d)) > 0.0f)) return 0.0f;   // :129-141
    const Shape &sh = sc.shapes[em.shape];
    float adp = fabsf(dp);
    float pdf = sh.inv_area * (adp != 0.0f ? (dist * dist) / adp : 0.0f);
    return pdf;
}

// ------------------------------------------------------------------------------ film
struct Film {
    const TrFilmDesc &fd; float *data; int C;
    TrCounters *cnt; std::vector<TrDeposit> *log;
    // add_transient_f + add_transient_data + put_ (box filter path); laser_x / laser_y address the
    // exhaustive_scan layout (transient_image_block.py:134-140) and are ignored by a plain film (:141-144)
    inline void add_transient(float px, float py, V3 spec, float distance, float sample_scale, float ray_weight,
                              uint32_t laser_x = 0, uint32_t laser_y = 0) {
        V3 value = (spec * sample_scale) * ray_weight;
        if (value.x == 0.0f && value.y == 0.0f && value.z == 0.0f) return;  // x + 0 == x: skipping is exact
        cnt->deposits++;
        if (log)    // positions are logged as the reference's add_transient_data receives them: film coordinates (window + offset)
            log->push_back(TrDeposit{px + (float) fd.crop_offset_x, py + (float) fd.crop_offset_y, distance, value.x, value.y, value.z});
        if (fd.type == TR_FILM_PHASOR_HDR) { add_phasor(px, py, value, distance); return; }
        float pos_distance = (distance - fd.start_opl) / fd.bin_width_opl;
        bool mask = (pos_distance >= 0.0f) && (pos_distance < (float) fd.temporal_bins);
        if (!mask) return;
        if (!(px >= 0.0f && py >= 0.0f)) return;                                             // :146  0 <= p
        uint32_t x = (uint32_t) floorf(px), y = (uint32_t) floorf(py), z = (uint32_t) floorf(pos_distance);
        if (!(x < fd.width && y < fd.height && z < fd.temporal_bins)) return;
        uint64_t pix = (uint64_t) y * fd.width + x;
        if (fd.laser_scan_width && fd.laser_scan_height)                                     // :134-138
            pix = (pix * fd.laser_scan_width + laser_x) * fd.laser_scan_height + laser_y;
        uint64_t index = (pix * fd.temporal_bins + z) * (uint64_t) C;
        data[index + 0] += value.x; data[index + 1] += value.y; data[index + 2] += value.z;
        // channel 3 (W) receives alpha = 0.0 in the reference (transient_image_block.py:99,148): no-op
        cnt->deposits_in_range++;
        cnt->bin_checksum += pix * fd.temporal_bins + z;
    }
    // PhasorHDRFilm.add_transient_data (phasor_hdr_film.py:241-264) + PhasorImageBlock.put
    // (phasor_image_block.py:43-67) + ImageBlock.put (box): monochromatic, two channels per frequency
    inline void add_phasor(float px, float py, V3 value, float distance) {
        float opl = distance - fd.start_opl;                                                 // :255
        if (!std::isfinite(opl)) return;                                                     // phasor_image_block.py:50
        if (!(px >= 0.0f && py >= 0.0f)) return;
        uint32_t x = (uint32_t) floorf(px), y = (uint32_t) floorf(py);
        if (!(x < fd.width && y < fd.height)) return;
        uint64_t pix = (uint64_t) y * fd.width + x;
        float *o = data + pix * (2ull * fd.n_frequencies + 1ull);
        const float two_pi = 6.283185307179586f, neg_two_pi = -6.283185307179586f;
        for (uint32_t f = 0; f < fd.n_frequencies; ++f) {
            float xx = (neg_two_pi * fd.frequencies[f]) * opl;                               // :56  -2 * pi * freq * opl
            float phase = xx - two_pi * floorf(xx / two_pi);                                 // :52-53 fmod
            float sn, cs; sincos_cephes(phase, &sn, &cs);
            o[2 * f] += value.x * cs; o[2 * f + 1] += value.x * sn;                          // :57-58
        }
        cnt->deposits_in_range++;
        cnt->bin_checksum += pix;
    }
};

// ------------------------------------------------------------------------------ sensor
inline Ray sample_camera_ray(const Scene &sc, uint32_t px, uint32_t py, float jx, float jy) {
    // ADIntegrator.sample_rays + perspective.sample_ray [upstream]
    const TrSensorDesc &se = sc.sensor;
    // pos = pixel + crop_offset; pos_adjusted = fma(pos + jitter, rcp(crop_size), -crop_offset * rcp(crop_size))
    float scx = 1.0f / (float) sc.film.width, scy = 1.0f / (float) sc.film.height;
    float fx = (float) (px + sc.film.crop_offset_x) + jx, fy = (float) (py + sc.film.crop_offset_y) + jy;
    float sx = fmaf(fx, scx, -(float) sc.film.crop_offset_x * scx), sy = fmaf(fy, scy, -(float) sc.film.crop_offset_y * scy);
    const float *m = se.sample_to_camera;
    float cx = fmaf(m[0], sx, fmaf(m[1], sy, m[3]));
    float cy = fmaf(m[4], sx, fmaf(m[5], sy, m[7]));
    float cz = fmaf(m[8], sx, fmaf(m[9], sy, m[11]));
    float cw = fmaf(m[12], sx, fmaf(m[13], sy, m[15]));
    float iw = 1.0f / cw;
    V3 d = normalize(v3(cx * iw, cy * iw, cz * iw));
    V3 dw = xf_vector(se.to_world, d);
    float inv_z = 1.0f / d.z;
    float near_t = se.near_clip * inv_z, far_t = se.far_clip * inv_z;
    V3 o = v3(se.to_world[3], se.to_world[7], se.to_world[11]);
    o = fma3(dw, near_t, o);
    return Ray{o, dw, far_t - near_t};
}

// ------------------------------------------------------------------------------ transient_path
// ---- Mitsuba 3 fresnel.h [upstream], restated ------------------------------------------------------------------
// fresnel_conductor(cos_theta_i, eta + i k): unpolarized reflectance of a conductor ("Optics", K.D. Moeller)
inline float safe_sqrt(float x) { return sqrtf(fmaxf(x, 0.0f)); }
inline float fresnel_conductor(float cos_theta_i, float eta_r, float eta_i) {
    float cos_theta_i_2 = cos_theta_i * cos_theta_i, sin_theta_i_2 = 1.0f - cos_theta_i_2, sin_theta_i_4 = sin_theta_i_2 * sin_theta_i_2;
    float temp_1 = eta_r * eta_r - eta_i * eta_i - sin_theta_i_2;
    float a_2_pb_2 = safe_sqrt(temp_1 * temp_1 + 4.0f * eta_i * eta_i * eta_r * eta_r);
    float a = safe_sqrt(0.5f * (a_2_pb_2 + temp_1));
    float term_1 = a_2_pb_2 + cos_theta_i_2, term_2 = 2.0f * cos_theta_i * a;
    float r_s = (term_1 - term_2) / (term_1 + term_2);
    float term_3 = a_2_pb_2 * cos_theta_i_2 + sin_theta_i_4, term_4 = term_2 * sin_theta_i_2;
    float r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return 0.5f * (r_s + r_p);
}
// fresnel(cos_theta_i, eta) of a dielectric interface -> (r, cos_theta_t, eta_it, eta_ti)
inline void fresnel_dielectric(float cos_theta_i, float eta, float *r_out, float *cos_theta_t_out, float *eta_it_out, float *eta_ti_out) {
    bool outside = cos_theta_i >= 0.0f;
    float rcp_eta = 1.0f / eta, eta_it = outside ? eta : rcp_eta, eta_ti = outside ? rcp_eta : eta;
    float cos_theta_t_sqr = fmaf(-fmaf(-cos_theta_i, cos_theta_i, 1.0f), eta_ti * eta_ti, 1.0f);
    float cos_theta_i_abs = fabsf(cos_theta_i), cos_theta_t_abs = safe_sqrt(cos_theta_t_sqr);
    bool index_matched = eta == 1.0f, special_case = index_matched || cos_theta_i_abs == 0.0f;
    float r_sc = index_matched ? 0.0f : 1.0f;
    float a_s = fmaf(-eta_it, cos_theta_t_abs, cos_theta_i_abs) / fmaf(eta_it, cos_theta_t_abs, cos_theta_i_abs);
    float a_p = fmaf(-eta_it, cos_theta_i_abs, cos_theta_t_abs) / fmaf(eta_it, cos_theta_i_abs, cos_theta_t_abs);
    float r = 0.5f * (a_s * a_s + a_p * a_p);
    if (special_case) r = r_sc;
    *r_out = r; *cos_theta_t_out = cos_theta_i >= 0.0f ? -cos_theta_t_abs : cos_theta_t_abs;   // mulsign_neg
    *eta_it_out = eta_it; *eta_ti_out = eta_ti;
}

// ---- BSDFs in the local shading frame (wi = si.to_local(-ray.d)), as Mitsuba's plugins state them [upstream] -------
struct BsdfSampleRec { V3 wo, weight; float pdf, eta; bool delta; };
inline bool bsdf_is_smooth(const TrBsdfDesc &b) { return b.type == TR_BSDF_DIFFUSE; }
// bsdf.eval_pdf: diffuse (optionally two-sided): R/pi * cos_o, pdf cos_o/pi, iff both cosines are positive; 0 for delta lobes
inline V3 bsdf_eval_pdf(const TrBsdfDesc &b, V3 wi, V3 wo, float *pdf) {
    *pdf = 0.0f;
    if (b.type != TR_BSDF_DIFFUSE) return v3(0, 0, 0);
    float cos_i = wi.z, cos_o = wo.z;
    if ((b.flags & TR_BSDF_FLAG_TWOSIDED) && cos_i < 0.0f) { cos_i = -cos_i; cos_o = -cos_o; }
    if (!(cos_i > 0.0f && cos_o > 0.0f)) return v3(0, 0, 0);
    *pdf = kInvPi * cos_o;
    return (v3(b.reflectance[0], b.reflectance[1], b.reflectance[2]) * kInvPi) * cos_o;
}
// bsdf.sample(ctx, si, sample1, sample2): diffuse.cpp, conductor.cpp, dielectric.cpp (+ twosided.cpp)
inline BsdfSampleRec bsdf_sample(const TrBsdfDesc &b, V3 wi, float s1, float s2x, float s2y) {
    BsdfSampleRec r; r.wo = v3(0, 0, 0); r.weight = v3(0, 0, 0); r.pdf = 0.0f; r.eta = 1.0f; r.delta = false;
    const bool twosided = (b.flags & TR_BSDF_FLAG_TWOSIDED) != 0;
    if (b.type == TR_BSDF_NULL) {                              // null.cpp: pass straight through, bs.wo = -si.wi
        r.wo = -wi; r.weight = v3(1, 1, 1); r.pdf = 1.0f; r.delta = true;
    } else if (b.type == TR_BSDF_DIFFUSE) {
        const bool flip = twosided && wi.z < 0.0f;
        const float cos_i_eff = flip ? -wi.z : wi.z;
        V3 wo = square_to_cosine_hemisphere(s2x, s2y);
        r.pdf = kInvPi * wo.z;
        if (cos_i_eff > 0.0f && r.pdf > 0.0f) r.weight = v3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
        if (flip) wo.z = -wo.z;
        r.wo = wo;
    } else if (b.type == TR_BSDF_CONDUCTOR) {                  // SmoothConductor::sample (+ TwoSided::sample)
        const bool flip = twosided && wi.z < 0.0f;
        if (flip) wi.z = -wi.z;
        V3 wo = v3(-wi.x, -wi.y, wi.z);                        // reflect(wi)
        r.pdf = 1.0f; r.delta = true;
        if (wi.z > 0.0f)
            r.weight = v3(b.reflectance[0] * fresnel_conductor(wi.z, b.eta[0], b.k[0]),
                          b.reflectance[1] * fresnel_conductor(wi.z, b.eta[1], b.k[1]),
                          b.reflectance[2] * fresnel_conductor(wi.z, b.eta[2], b.k[2]));
        if (flip) wo.z = -wo.z;
        r.wo = wo;
    } else if (b.type == TR_BSDF_DIELECTRIC) {                 // SmoothDielectric::sample
        float r_i, cos_theta_t, eta_it, eta_ti;
        fresnel_dielectric(wi.z, b.eta[0], &r_i, &cos_theta_t, &eta_it, &eta_ti);
        const float t_i = 1.0f - r_i;
        const bool selected_r = s1 <= r_i;
        r.pdf = selected_r ? r_i : t_i; r.delta = true;
        if (selected_r) {
            r.wo = v3(-wi.x, -wi.y, wi.z);
            r.weight = v3(b.reflectance[0], b.reflectance[1], b.reflectance[2]);
        } else {
            r.wo = v3(-eta_ti * wi.x, -eta_ti * wi.y, cos_theta_t);            // refract(wi, cos_theta_t, eta_ti)
            r.eta = eta_it;
            const float f2 = eta_ti * eta_ti;                  // radiance transport: scale by eta_ti^2
            r.weight = v3(b.specular_transmittance[0] * f2, b.specular_transmittance[1] * f2, b.specular_transmittance[2] * f2);
        }
    }
    return r;
}

// mitransient/integrators/transientpath.py:109-326 (primal branch).  Returns L.
V3 trace_transient_path(const Scene &sc, const TrRenderParams &rp, Pcg32 &rng, Ray ray,
                        float px, float py, float sample_scale, float ray_weight, Film &film) {
    const bool camera_unwarp = (rp.flags & TR_FLAG_CAMERA_UNWARP) != 0;
    const bool discard_direct_light = (rp.flags & TR_FLAG_DISCARD_DIRECT_LIGHT) != 0;
    const uint32_t n_em = (uint32_t) sc.emitters.size();

    uint32_t depth = 0;                       // :119
    V3 L = v3(0, 0, 0);                       // :120
    V3 beta = v3(1, 1, 1);                    // utils.py:12 (unpolarized beta_init)
    float eta = 1.0f;                         // :125
    bool active = true;                       // :126
    float distance = 0.0f;                    // :127
    V3 prev_p = v3(0, 0, 0);                  // prev_si = zeros  :130
    float prev_bsdf_pdf = 1.0f;               // :131
    bool prev_bsdf_delta = true;              // :132

    if (camera_unwarp) {                      // :134-139
        Hit h0 = ray_intersect(sc, ray);
        if (h0.valid) distance = -h0.t;
    }

    for (uint32_t it = 0; active && it < rp.max_depth; ++it) {   // :141-143 (max_iterations = max_depth)
        film.cnt->vertices++;
        bool active_next = active;
        Hit si = ray_intersect(sc, ray);                          // :150-152
        distance += si.t * eta;                                   // :155 (inf on a miss)

        // ---------------- direct emission :159-181
        V3 Le = v3(0, 0, 0);
        if (si.valid) {
            const Shape &sh = sc.shapes[si.shape];
            if (sh.d.emitter >= 0) {
                const TrEmitterDesc &em = sc.emitters[sh.d.emitter];
                V3 rel = si.p - prev_p;                           // DirectionSample3f(scene, si, prev_si)
                float dist = sqrtf(dot(rel, rel));
                V3 dsd = rel * (1.0f / dist);
                float em_pdf = 0.0f;
                if (!prev_bsdf_delta) {
                    em_pdf = area_pdf_direction(sc, em, dsd, si.n, dist);
                    if (n_em > 1) em_pdf *= 1.0f / (float) n_em;  // Scene::pdf_emitter
                }
                float mis = mis_weight(prev_bsdf_pdf, em_pdf);    // :169-172
                V3 wi = v3(-dot(ray.d, si.s), -dot(ray.d, si.tt), -dot(ray.d, si.n));   // si.wi
                if (!discard_direct_light) Le = (beta * mis) * emitter_eval(em, wi);        // emitter.eval(si)  :175-177
            }
        }
        film.add_transient(px, py, Le, distance, sample_scale, ray_weight);   // :180-181

        // ---------------- emitter sampling :186-219
        active_next = active_next && (depth + 1 < rp.max_depth) && si.valid;  // :186
        const TrBsdfDesc *bsdf = si.valid ? &sc.bsdfs[sc.shapes[si.shape].d.bsdf] : nullptr;
        bool smooth = bsdf && bsdf->type == TR_BSDF_DIFFUSE;
        bool active_em = active_next && smooth;                               // :189-190

        float u0 = rng.next_1d(), u1 = rng.next_1d();                         // sampler.next_2d()  :194
        V3 Lr_dir = v3(0, 0, 0);
        float cos_i = si.valid ? -dot(ray.d, si.n) : 0.0f;
        bool back = cos_i < 0.0f;
        bool twosided = bsdf && (bsdf->flags & TR_BSDF_FLAG_TWOSIDED);
        float cos_i_eff = (twosided && back) ? -cos_i : cos_i;
        float ds_dist = 0.0f;
        if (active_em && n_em > 0) {
            DirectionSample ds; uint32_t ei;
            V3 em_weight = scene_sample_emitter_direction(sc, si.p, u0, u1, &ds, &ei);
            ds_dist = ds.dist;
            bool nz = (em_weight.x != 0.0f || em_weight.y != 0.0f || em_weight.z != 0.0f);
            if (ds.pdf != 0.0f && nz) {
                // bsdf.eval_pdf :208-211 (diffuse: R/pi * cos_o; pdf = cos_o/pi; needs both cosines > 0)
                float cos_o = dot(ds.d, si.n);
                float cos_o_eff = (twosided && back) ? -cos_o : cos_o;
                if (cos_i_eff > 0.0f && cos_o_eff > 0.0f) {
                    film.cnt->shadow_rays++;
                    bool occluded = ray_test(sc, spawn_ray_to(si.p, si.n, ds.p));
                    if (!occluded) {
                        float bsdf_pdf;       // only the z components of wi / wo matter to a diffuse lobe
                        V3 bsdf_val = bsdf_eval_pdf(*bsdf, v3(0.0f, 0.0f, cos_i), v3(0.0f, 0.0f, cos_o), &bsdf_pdf);
                        float mis_em = ds.delta ? 1.0f : mis_weight(ds.pdf, bsdf_pdf);  // :212-213
                        Lr_dir = ((beta * mis_em) * bsdf_val) * em_weight;              // :214
                    }
                }
            }
        }
        film.add_transient(px, py, Lr_dir, distance + ds_dist * eta, sample_scale, ray_weight);  // :217-219

        // ---------------- BSDF sampling :223-228
        float s1 = rng.next_1d();                                  // sampler.next_1d() (lobe selection of `dielectric`)
        float s2x = rng.next_1d(), s2y = rng.next_1d();            // sampler.next_2d()
        V3 bsdf_weight = v3(0, 0, 0); float bs_pdf = 0.0f; V3 wo_world = v3(0, 0, 0);
        float bs_eta = 1.0f; bool bs_delta = false;
        if (active_next) {
            // si.wi = si.to_local(-ray.d); the plugins sample in the local frame (diffuse / conductor / dielectric / null .cpp);
            // a null boundary is a delta lobe that passes straight through and still counts as a vertex (:318)
            V3 wi = v3(-dot(ray.d, si.s), -dot(ray.d, si.tt), -dot(ray.d, si.n));
            BsdfSampleRec bs = bsdf_sample(*bsdf, wi, s1, s2x, s2y);
            bsdf_weight = bs.weight; bs_pdf = bs.pdf; bs_eta = bs.eta; bs_delta = bs.delta;
            wo_world = fma3(si.n, bs.wo.z, fma3(si.tt, bs.wo.y, si.s * bs.wo.x));      // si.to_world(wo)
        }

        // ---------------- update loop variables :231-241
        L = L + Le + Lr_dir;                                       // :231
        if (si.valid) ray = spawn_ray(si.p, si.n, wo_world);       // :232
        eta *= bs_eta;                                             // :233 (== 1 except for a refraction)
        beta = beta * bsdf_weight;                                 // :234
        prev_p = si.p; prev_bsdf_pdf = bs_pdf; prev_bsdf_delta = bs_delta;   // :238-241

        // ---------------- stopping criterion :246-258
        float beta_max = max3(beta);
        active_next = active_next && (beta_max != 0.0f);           // :247
        float rr_prob = fminf(beta_max * (eta * eta), 0.95f);      // :251
        active_next = active_next && (rr_prob > 0.0f);             // :252
        bool rr_active = depth >= rp.rr_depth;                     // :255
        if (rr_active) { float r = (rr_prob > 0.0f) ? (1.0f / rr_prob) : 0.0f; beta = beta * r; }  // :256
        bool rr_continue = rng.next_1d() < rr_prob;                // :257
        active_next = active_next && (!rr_active || rr_continue);  // :258

        if (si.valid) depth += 1;                                  // :319
        active = active_next;                                      // :320
    }
    return L;
}


// =============================================================================================
//  transient_nlos_path  (mitransient/integrators/transientnlospath.py) + nlos_capture_meter
//  (mitransient/sensors/nloscapturemeter.py).  Single, Confocal and Exhaustive (equal grids) captures;
//  AD is out of scope.
// =============================================================================================
constexpr float kEps = 5.9604644775390625e-08f;   // dr.epsilon(Float) = 2^-24

struct NlosCtx {
    const TrEmitterDesc *em; V3 em_origin, em_forward;
    std::vector<V3> sensor_targets; V3 laser_target_single; bool laser_single_valid;
    std::vector<V3> laser_targets;                              // Exhaustive capture: the illuminated points (:345-379)
    std::vector<float> hg_pmf, hg_cdf; float hg_sum, hg_norm;   // DiscreteDistribution over shapes (:277-292)
    uint32_t sensor_shape;
    uint32_t laser_res_x = 0, laser_res_y = 0;                  // Exhaustive capture: illumination grid (:340-348)
};

// NLOSCaptureMeter._sample_direction + sample_ray(_differential)   nloscapturemeter.py:136-202
inline Ray nlos_sensor_ray(const Scene &sc, float sx, float sy) {
    const Shape &wall = sc.shapes[sc.sensor.shape];
    float Wf = (float) sc.film.width, Hf = (float) sc.film.height;
    float gx = (floorf(sx * Wf) + 0.5f) / Wf, gy = (floorf(sy * Hf) + 0.5f) / Hf;           // :149-150
    V3 target = xf_point(wall.d.to_world, fmaf(gx, 2.0f, -1.0f), fmaf(gy, 2.0f, -1.0f), 0.0f);   // :151-153
    if (sc.sensor.is_confocal) target = v3(sc.sensor.laser_target[0], sc.sensor.laser_target[1], sc.sensor.laser_target[2]);   // :142-144
    V3 origin = v3(sc.sensor.origin[0], sc.sensor.origin[1], sc.sensor.origin[2]);
    V3 dir = target - origin;                                                                  // :155-157
    float dist = sqrtf(dot(dir, dir));
    dir = v3(dir.x / dist, dir.y / dist, dir.z / dist);                                        // direction /= distance (Python-level: true division)
    return Ray{origin, dir, std::numeric_limits<float>::max()};
}

// TransientNLOSPath.prepare   transientnlospath.py:251-383
bool nlos_prepare(const Scene &sc, const TrRenderParams &rp, NlosCtx &ctx, std::string &err) {
    if (sc.emitters.size() != 1 || sc.emitters[0].type != TR_EMITTER_PROJECTOR) { err = "transient_nlos_path needs exactly one projector emitter"; return false; }
    const bool meter = sc.sensor.type == TR_SENSOR_NLOS_CAPTURE_METER;
    if (meter && (sc.sensor.shape < 0 || (uint32_t) sc.sensor.shape >= sc.shapes.size() ||
                  sc.shapes[sc.sensor.shape].d.type != TR_SHAPE_RECTANGLE)) { err = "transient_nlos_path needs an nlos_capture_meter on a rectangle"; return false; }
    if (!meter && sc.sensor.type != TR_SENSOR_PERSPECTIVE) { err = "transient_nlos_path: unknown sensor"; return false; }
    if (rp.capture_type != TR_CAPTURE_SINGLE && rp.capture_type != TR_CAPTURE_CONFOCAL && rp.capture_type != TR_CAPTURE_EXHAUSTIVE) {
        err = "unknown capture type"; return false;
    }
    if (rp.capture_type == TR_CAPTURE_EXHAUSTIVE) {
        // laser_resolution = (film.laser_scan_width, film.laser_scan_height) (:341-343); the sampling loop runs over the
        // sensor grid when force_equal_illumination_scanning is set (:599-601), over the laser grid otherwise (:602-604)
        const bool have = sc.film.laser_scan_width && sc.film.laser_scan_height;
        if (rp.force_equal_grids) {
            ctx.laser_res_x = sc.film.width; ctx.laser_res_y = sc.film.height;
            if ((rp.flags & TR_FLAG_NLOS_LASER_SAMPLING) && have &&
                (sc.film.laser_scan_width != sc.film.width || sc.film.laser_scan_height != sc.film.height)) {
                err = "Sensor and laser scan resolution must be equal if force_equal_illumination_scanning is set to True"; return false;
            }
        } else {
            if (!have) { err = "Exhaustive capture with its own laser grid needs film.laser_scan_width / laser_scan_height"; return false; }
            ctx.laser_res_x = sc.film.laser_scan_width; ctx.laser_res_y = sc.film.laser_scan_height;
        }
    }
    ctx.em = &sc.emitters[0];
    const float *m = ctx.em->to_world;
    ctx.em_origin = v3(m[3], m[7], m[11]);
    ctx.em_forward = v3(m[2], m[6], m[10]);                   // trafo @ Vector3f(0, 0, 1)
    ctx.sensor_shape = meter ? (uint32_t) sc.sensor.shape : 0xffffffffu;
    // hidden-geometry area distribution :277-292
    ctx.hg_pmf.clear(); ctx.hg_cdf.clear();
    double acc = 0.0;
    for (uint32_t i = 0; i < sc.shapes.size(); ++i) {
        bool wall = (sc.shapes[i].d.flags & TR_SHAPE_FLAG_SENSOR) != 0;
        float a = (wall && !(rp.flags & TR_FLAG_NLOS_HG_SAMPLING_INCLUDES_RELAY_WALL)) ? 0.0f : sc.shapes[i].area;
        acc += (double) a; ctx.hg_pmf.push_back(a); ctx.hg_cdf.push_back((float) acc);
    }
    ctx.hg_sum = (float) acc; ctx.hg_norm = (float) (1.0 / acc);
    if ((rp.flags & TR_FLAG_NLOS_HG_SAMPLING) && !(acc > 0.0)) { err = "hidden geometry sampling without hidden geometry"; return false; }
    // one sensor ray per film pixel -> scanned points :295-312
    const uint32_t W = sc.film.width, H = sc.film.height;
    ctx.sensor_targets.resize((size_t) W * H);
    float stepx = 1.0f / (float) W, stepy = 1.0f / (float) H;     // dr.linspace(0, 1, n, endpoint=False)
    for (uint32_t y = 0; y < H; ++y) for (uint32_t x = 0; x < W; ++x) {
        // sensor.sample_ray at the film position (x / W, y / H): the capture meter snaps to the pixel centre, a
        // perspective camera shoots through the pixel corner
        Hit h = ray_intersect(sc, meter ? nlos_sensor_ray(sc, fmaf((float) x, stepx, 0.0f), fmaf((float) y, stepy, 0.0f))
                                        : sample_camera_ray(sc, x, y, 0.0f, 0.0f));
        ctx.sensor_targets[(size_t) y * W + x] = h.p;             // si.p (zeros when the ray misses)
    }
    // Single capture: where the laser axis meets the scene :328-336
    Hit hl = ray_intersect(sc, Ray{ctx.em_origin, ctx.em_forward, std::numeric_limits<float>::max()});
    ctx.laser_target_single = hl.p; ctx.laser_single_valid = hl.valid;
    if (rp.capture_type == TR_CAPTURE_EXHAUSTIVE) {
        if (rp.force_equal_grids) ctx.laser_targets = ctx.sensor_targets;                      // :345-348
        else {
            // :349-379: a copy of the projector with fov = illumination_scan_fov shoots one ray per laser grid point
            // (meshgrid of linspace(0, 1, res, endpoint=False), x fastest); projector.sample_ray [upstream]: the direction
            // sample picks the point of the 1x1 irradiance texture the ray goes through
            const uint32_t lw = ctx.laser_res_x, lh = ctx.laser_res_y;
            ctx.laser_targets.resize((size_t) lw * lh);
            float sx_step = 1.0f / (float) lw, sy_step = 1.0f / (float) lh;
            const float *m = rp.laser_scan_sample_to_camera;
            bool any = false;
            for (uint32_t y = 0; y < lh; ++y) for (uint32_t x = 0; x < lw; ++x) {
                float u = fmaf((float) x, sx_step, 0.0f), v = fmaf((float) y, sy_step, 0.0f);
                float cx = fmaf(m[0], u, fmaf(m[1], v, m[3])), cy = fmaf(m[4], u, fmaf(m[5], v, m[7]));
                float cz = fmaf(m[8], u, fmaf(m[9], v, m[11])), cw = fmaf(m[12], u, fmaf(m[13], v, m[15]));
                float iw = 1.0f / cw;
                V3 dl = normalize(v3(cx * iw, cy * iw, cz * iw));
                V3 dw = xf_vector(ctx.em->to_world, dl);
                Hit h = ray_intersect(sc, Ray{ctx.em_origin, dw, std::numeric_limits<float>::max()});
                ctx.laser_targets[(size_t) y * lw + x] = h.p; any = any || h.valid;
            }
            // the whole laser-target block sits under `if self.laser_sampling:` (:322): without laser sampling nothing is checked
            if (!any && (rp.flags & TR_FLAG_NLOS_LASER_SAMPLING)) { err = "The emitter did not intersect any geometry in the scene."; return false; }
        }
    }
    if (rp.capture_type == TR_CAPTURE_SINGLE && (rp.flags & TR_FLAG_NLOS_LASER_SAMPLING) && !hl.valid) { err = "The emitter is not pointing at the scene!"; return false; }
    return true;
}

// diffuse (optionally twosided) bsdf.eval:  R/pi * cos_o  iff both cosines are positive
inline V3 bsdf_eval(const TrBsdfDesc &b, float cos_i, float cos_o) {
    if (b.type != TR_BSDF_DIFFUSE) return v3(0, 0, 0);
    if ((b.flags & TR_BSDF_FLAG_TWOSIDED) && cos_i < 0.0f) { cos_i = -cos_i; cos_o = -cos_o; }
    if (!(cos_i > 0.0f && cos_o > 0.0f)) return v3(0, 0, 0);
    return (v3(b.reflectance[0], b.reflectance[1], b.reflectance[2]) * kInvPi) * cos_o;
}

// Projector::sample_direction [upstream projector.cpp]: delta position, perspective uv lookup, weight
// irradiance * pi / (z_local^2 * cos(axis, -d)); pdf = 1.
inline V3 projector_sample_direction(const TrEmitterDesc &em, V3 itp, V3 *ds_p, float *ds_dist) {
    V3 loc = xf_point(em.world_to_local, itp.x, itp.y, itp.z);
    const float *c = em.camera_to_sample;
    float sx = fmaf(c[0], loc.x, fmaf(c[1], loc.y, fmaf(c[2], loc.z, c[3])));
    float sy = fmaf(c[4], loc.x, fmaf(c[5], loc.y, fmaf(c[6], loc.z, c[7])));
    float sw = fmaf(c[12], loc.x, fmaf(c[13], loc.y, fmaf(c[14], loc.z, c[15])));
    float iw = 1.0f / sw;
    float u = sx * iw, v = sy * iw;
    bool active = (u >= 0.0f) && (u <= 1.0f) && (v >= 0.0f) && (v <= 1.0f) && (loc.z > 0.0f);
    V3 p = v3(em.to_world[3], em.to_world[7], em.to_world[11]);
    V3 n = v3(em.to_world[2], em.to_world[6], em.to_world[10]);
    V3 d = p - itp;
    float dist = sqrtf(dot(d, d));
    d = d * (1.0f / dist);
    *ds_p = p; *ds_dist = dist;
    if (!active) return v3(0, 0, 0);
    float f = kPi / ((loc.z * loc.z) * (-dot(n, d)));
    return v3(em.radiance[0], em.radiance[1], em.radiance[2]) * f;
}

struct NlosVertex { V3 p, n; float cos_i; const TrBsdfDesc *bsdf; };

// emitter_nee_sample   transientnlospath.py:433-509
// `distance` is taken BY REFERENCE: the reference adds ds.dist to its argument in place (:497-498), and with plain emitter
// sampling that argument IS the loop's distance (Dr.Jit arrays are mutable objects) -- so with nlos_laser_sampling off and
// account_first_and_last_bounces on, every emitter connection also lengthens the rest of the path (SURVEY appendix B.11;
// reproduced because the executed reference does it).  The laser-sampling path passes a temporary (:562).
V3 nlos_emitter_nee_sample(const Scene &sc, const TrRenderParams &rp, const NlosCtx &ctx, Pcg32 &rng, const NlosVertex &vx,
                           V3 beta, float &distance, float eta, uint32_t depth, bool active_e, bool focus_laser,
                           float px, float py, float sample_scale, Film &film, uint32_t laser_x = 0, uint32_t laser_y = 0) {
    if (!active_e) {
        // a vertex without a smooth lobe (mirror, glass) asks for no emitter sample, yet ds.dist is filled and :497-498 runs
        // unmasked: the aliased distance still grows (paths that end here are not affected: nothing reads it afterwards)
        if ((rp.flags & TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES) && vx.bsdf && !focus_laser) {
            V3 dd = ctx.em_origin - vx.p;
            distance += sqrtf(dot(dd, dd)) * eta;
        }
        return v3(0, 0, 0);
    }
    film.cnt->shadow_rays++;
    V3 query = vx.p;
    if (focus_laser && (rp.capture_type == TR_CAPTURE_CONFOCAL || rp.capture_type == TR_CAPTURE_EXHAUSTIVE)) {   // :448-458
        V3 rel = ctx.em_origin - vx.p;
        float dte = sqrtf(dot(rel, rel));
        query = v3(ctx.em_origin.x + ctx.em_forward.x * dte, ctx.em_origin.y + ctx.em_forward.y * dte,
                   ctx.em_origin.z + ctx.em_forward.z * dte);
    }
    if (ray_test(sc, spawn_ray_to(vx.p, vx.n, ctx.em_origin))) {                              // :442
        // occluded: no deposit and no sampler draw, but the projector's sample_direction fills ds.dist without looking at
        // its mask (mitsuba projector.cpp, a direct call on the one emitter object), and :497-498 is unmasked too
        if (rp.flags & TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES) {
            V3 dd = ctx.em_origin - query;
            distance += sqrtf(dot(dd, dd)) * eta;
        }
        return v3(0, 0, 0);
    }
    rng.next_1d(); rng.next_1d();                                                             // sampler.next_2d(active_e)
    V3 dsp; float ds_dist;
    V3 em_weight = projector_sample_direction(*ctx.em, query, &dsp, &ds_dist);
    V3 wd = normalize(dsp - vx.p);                                                            // :484
    V3 bsdf_spec = bsdf_eval(*vx.bsdf, vx.cos_i, dot(wd, vx.n));
    if (rp.filter_depth != -1) active_e = active_e && ((int32_t) depth == rp.filter_depth);   // :489-490
    if (rp.flags & TR_FLAG_DISCARD_DIRECT_PATHS) active_e = active_e && (depth > 2);          // :491-492
    V3 Lr = active_e ? (beta * bsdf_spec) * em_weight : v3(0, 0, 0);                          // :493
    if (rp.flags & TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES) distance += ds_dist * eta;         // :497-498
    film.add_transient(px, py, Lr, distance, sample_scale, 1.0f, laser_x, laser_y);           // :500-507
    return Lr;
}

// emitter_laser_targets_sample   transientnlospath.py:512-564   (depth: already incremented by the caller, :635)
V3 nlos_emitter_laser_targets_sample(const Scene &sc, const TrRenderParams &rp, const NlosCtx &ctx, Pcg32 &rng, const NlosVertex &vx,
                                     V3 target, V3 beta, float distance, float eta, uint32_t depth, float px, float py,
                                     float sample_scale, Film &film, uint32_t laser_x, uint32_t laser_y) {
    V3 d = target - vx.p;                                                                     // :526
    float dl = sqrtf(dot(d, d));
    if (!(dl > 0.0f)) return v3(0, 0, 0);     // confocal depth-0 vertex: target == si.p, NaN direction in the reference (:526-528, :545)
    d = v3(d.x / dl, d.y / dl, d.z / dl);                                                     // :528  d /= distance_laser
    Ray ray_bsdf = spawn_ray_to(vx.p, vx.n, target);                                          // :529
    film.cnt->shadow_rays++;
    if (ray_test(sc, ray_bsdf)) return v3(0, 0, 0);                                           // :530
    V3 bsdf_spec = bsdf_eval(*vx.bsdf, vx.cos_i, dot(d, vx.n));                               // :533-535
    ray_bsdf.maxt = std::numeric_limits<float>::max();                                        // :537
    Hit sb = ray_intersect(sc, ray_bsdf);                                                     // :538-540
    if (!sb.valid) return v3(0, 0, 0);
    if (!(bsdf_spec.x > kEps || bsdf_spec.y > kEps || bsdf_spec.z > kEps)) return v3(0, 0, 0);   // :541-542
    float cos_wl = -dot(d, sb.n);                                                             // :544
    if (!(cos_wl > 0.0f)) return v3(0, 0, 0);                                                 // :545
    float pdf_ls = (dl * dl) / cos_wl;                                                        // :546-550
    bsdf_spec = v3(bsdf_spec.x / pdf_ls, bsdf_spec.y / pdf_ls, bsdf_spec.z / pdf_ls);         // :551
    NlosVertex vb{sb.p, sb.n, -dot(ray_bsdf.d, sb.n), &sc.bsdfs[sc.shapes[sb.shape].d.bsdf]}; // :553
    float distance_at_spot = distance + dl * eta;                                             // a temporary: not aliased
    return nlos_emitter_nee_sample(sc, rp, ctx, rng, vb, beta * bsdf_spec, distance_at_spot, eta, depth + 1, true, true,
                                   px, py, sample_scale, film, laser_x, laser_y);             // :560-564
}

// emitter_laser_sample   transientnlospath.py:567-635
V3 nlos_emitter_laser_sample(const Scene &sc, const TrRenderParams &rp, const NlosCtx &ctx, Pcg32 &rng, const NlosVertex &vx,
                             V3 beta, float distance, float eta, uint32_t depth, bool active_e, uint32_t pixel_index,
                             float px, float py, float sample_scale, Film &film) {
    if (!active_e) return v3(0, 0, 0);
    if (rp.capture_type == TR_CAPTURE_EXHAUSTIVE) {                                           // :597-627
        // every scanned point is illuminated by every laser point; the illuminated points are the sensor's
        const uint32_t n = ctx.laser_res_x * ctx.laser_res_y;
        V3 sum = v3(0, 0, 0);
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t laser_x = (uint32_t) ((float) i / (float) ctx.laser_res_y);              // :618  UInt(i / Float(res_y))
            uint32_t laser_y = i % ctx.laser_res_y;                                           // :619
            sum = sum + nlos_emitter_laser_targets_sample(sc, rp, ctx, rng, vx, ctx.laser_targets[i], beta, distance, eta,
                                                          depth + 1, px, py, sample_scale, film, laser_x, laser_y);
        }
        return sum * (1.0f / (float) n);                                                      // :627
    }
    V3 target = (rp.capture_type == TR_CAPTURE_CONFOCAL) ? ctx.sensor_targets[pixel_index] : ctx.laser_target_single;  // :588-596
    return nlos_emitter_laser_targets_sample(sc, rp, ctx, rng, vx, target, beta, distance, eta, depth + 1, px, py, sample_scale,
                                             film, 0, 0);                                     // :631-635
}

V3 trace_nlos_path(const Scene &sc, const TrRenderParams &rp, const NlosCtx &ctx, Pcg32 &rng, Ray ray, uint32_t pixel_index,
                   float px, float py, float sample_scale, Film &film) {
    const bool laser_sampling = (rp.flags & TR_FLAG_NLOS_LASER_SAMPLING) != 0;
    const bool hg_sampling = (rp.flags & TR_FLAG_NLOS_HG_SAMPLING) != 0;
    const bool hg_rr = (rp.flags & TR_FLAG_NLOS_HG_SAMPLING_DO_RROULETTE) != 0 && hg_sampling;
    const bool account_fl = (rp.flags & TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES) != 0;
    uint32_t depth = 0; V3 L = v3(0, 0, 0), beta = v3(1, 1, 1); float eta = 1.0f; bool active = true;
    float distance = 0.0f;                                                                    // ray.time  :718

    for (uint32_t it = 0; active && it < rp.max_depth; ++it) {                                // :738-740
        film.cnt->vertices++;
        Hit si = ray_intersect(sc, ray);                                                      // :746-749
        if (account_fl || depth > 0) distance += si.t * eta;                                  // :751-752
        // direct emission :757-777: the only emitter is the (delta) projector, no shape emits: Le = 0
        bool active_next = active && (depth + 1 < rp.max_depth) && si.valid;                  // :781
        const TrBsdfDesc *bsdf = si.valid ? &sc.bsdfs[sc.shapes[si.shape].d.bsdf] : nullptr;
        bool smooth = bsdf && bsdf->type == TR_BSDF_DIFFUSE;
        bool active_em = active_next && smooth;                                               // :784-785
        float cos_i = si.valid ? -dot(ray.d, si.n) : 0.0f;
        NlosVertex vx{si.p, si.n, cos_i, bsdf};
        V3 Lr = laser_sampling                                                                // :733-736, :788-793
            ? nlos_emitter_laser_sample(sc, rp, ctx, rng, vx, beta, distance, eta, depth, active_em, pixel_index, px, py, sample_scale, film)
            : nlos_emitter_nee_sample(sc, rp, ctx, rng, vx, beta, distance, eta, depth, active_em, false, px, py, sample_scale, film);

        // ---- direction sampling :797-827
        bool do_hg = hg_sampling; float pdf_method = 1.0f;
        if (hg_rr) { do_hg = rng.next_1d() < 0.5f; pdf_method = 0.5f; }                       // :799-806
        float h1 = rng.next_1d(); (void) h1;                                                  // arguments of hidden_geometry_sample
        float h2x = rng.next_1d(), h2y = rng.next_1d();                                       //   are drawn unconditionally (:814)
        float b1 = rng.next_1d(); (void) b1;                                                  // arguments of bsdf.sample (:821)
        float b2x = rng.next_1d(), b2y = rng.next_1d();
        V3 weight = v3(0, 0, 0), wo = v3(0, 0, 0); float bs_pdf = 0.0f, bs_eta = 1.0f;
        if (active_next && do_hg && hg_sampling) {                                            // hidden_geometry_sample :638-670
            // DiscreteDistribution.sample_reuse_pmf over shapes :420-423
            float value = h2x * ctx.hg_sum;
            uint32_t lo = 0, hi = (uint32_t) ctx.hg_cdf.size() - 1;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (ctx.hg_cdf[mid] < value) lo = mid + 1; else hi = mid; }
            float cdf_prev = lo ? ctx.hg_cdf[lo - 1] : 0.0f;
            float pmf = ctx.hg_pmf[lo] * ctx.hg_norm;
            float u0 = (h2x - cdf_prev * ctx.hg_norm) / pmf;
            V3 pp, pn; float ppdf;
            shape_sample_position(sc, sc.shapes[lo], u0, h2y, &pp, &pn, &ppdf);               // :427
            ppdf *= pmf;                                                                      // :428
            V3 d = pp - si.p;
            float dist = sqrtf(dot(d, d));
            d = v3(d.x / dist, d.y / dist, d.z / dist);                                       // :646  d /= dist
            float cos_ti = dot(si.n, d), cos_g = -dot(pn, d);                                 // :648-649
            bool act = (cos_ti > kEps) && (cos_g > kEps);
            wo = v3(dot(d, si.s), dot(d, si.tt), dot(d, si.n));                               // si.to_local(d)
            V3 spec = smooth ? bsdf_eval(*bsdf, cos_i, wo.z) : v3(0, 0, 0);
            bs_pdf = ppdf * (dist * dist) / fabsf(cos_g);                                     // :662
            act = act && (bs_pdf > kEps);
            if (act) weight = v3(spec.x / bs_pdf, spec.y / bs_pdf, spec.z / bs_pdf);
        } else if (active_next && !do_hg) {                                                   // bsdf.sample :819-824
            // in the local frame, like the plugins (diffuse: cosine hemisphere; conductor / dielectric: delta lobes; null)
            V3 wi = v3(-dot(ray.d, si.s), -dot(ray.d, si.tt), -dot(ray.d, si.n));
            BsdfSampleRec bs = bsdf_sample(*bsdf, wi, b1, b2x, b2y);
            wo = bs.wo; bs_pdf = bs.pdf; weight = bs.weight; bs_eta = bs.eta;
        }
        // ---- update loop variables :831-841
        L = L + Lr;
        V3 wo_world = fma3(si.n, wo.z, fma3(si.tt, wo.y, si.s * wo.x));
        if (si.valid) ray = spawn_ray(si.p, si.n, wo_world);
        eta = eta * bs_eta;                                                                   // :833  (hidden-geometry samples: 1)
        V3 bw = beta * weight;
        beta = v3(bw.x / pdf_method, bw.y / pdf_method, bw.z / pdf_method);                   // :834
        // ---- stopping criterion :846-858 (same as transient_path)
        float beta_max = max3(beta);
        active_next = active_next && (beta_max != 0.0f);
        float rr_prob = fminf(beta_max * (eta * eta), 0.95f);
        active_next = active_next && (rr_prob > 0.0f);
        bool rr_active = depth >= rp.rr_depth;
        if (rr_active) { float r = (rr_prob > 0.0f) ? (1.0f / rr_prob) : 0.0f; beta = beta * r; }
        bool rr_continue = rng.next_1d() < rr_prob;
        active_next = active_next && (!rr_active || rr_continue);
        if (si.valid) depth += 1;                                                             // :919
        active = active_next;
    }
    return L;
}


// =============================================================================================
//  transient_prbvolpath  (mitransient/integrators/transient_prbvolpath.py), primal, homogeneous media
//  with Henyey-Greenstein phase functions and `null` boundaries (the reference's participating-media
//  example, examples/transient/cornell-box/cbox_volumetric.xml).
//  Upstream pieces restated: Medium::sample_interaction / transmittance_eval_pdf (homogeneous),
//  HGPhaseFunction::sample / eval_pdf, Null BSDF, Interaction::spawn_ray(_to).
// =============================================================================================
struct MediumEvent { bool valid; float t, mint; V3 p, wi; V3 sigma_s, sigma_t; int medium; };

inline V3 medium_sigma_t(const TrMediumDesc &m) { return v3(m.sigma_t[0] * m.scale, m.sigma_t[1] * m.scale, m.sigma_t[2] * m.scale); }
inline float chan(V3 v, uint32_t c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }

// Medium::sample_interaction (homogeneous): free-flight distance by the chosen channel's extinction
inline MediumEvent medium_sample_interaction(const Scene &sc, int medium, const Ray &ray, float sample, uint32_t channel) {
    MediumEvent me; me.medium = medium; me.wi = -ray.d; me.mint = 0.0f;
    const TrMediumDesc &m = sc.media[medium];
    V3 ce = medium_sigma_t(m);
    float mm = chan(ce, channel);
    float sampled_t = 0.0f + (-log_cephes(1.0f - sample) / mm);
    me.valid = sampled_t <= ray.maxt;
    me.t = me.valid ? sampled_t : kInf;
    me.p = fma3(ray.d, sampled_t, ray.o);
    me.sigma_t = ce;
    me.sigma_s = v3(ce.x * m.albedo[0], ce.y * m.albedo[1], ce.z * m.albedo[2]);
    return me;
}
// Medium::transmittance_eval_pdf: tr = exp(-t * sigma_t), pdf = tr (surface first) or tr * sigma_t
inline void medium_transmittance_eval_pdf(const MediumEvent &me, float si_t, V3 *tr, V3 *pdf) {
    float t = fminf(me.t, si_t) - me.mint;
    *tr = v3(exp_cephes(-t * me.sigma_t.x), exp_cephes(-t * me.sigma_t.y), exp_cephes(-t * me.sigma_t.z));
    *pdf = (si_t < me.t) ? *tr : (*tr) * me.sigma_t;
}
// HGPhaseFunction
inline float hg_eval(float g, float cos_theta) {
    float temp = 1.0f + g * g + 2.0f * g * cos_theta;
    return 0.07957747154594766788f * (1.0f - g * g) / (temp * sqrtf(temp));
}
inline V3 hg_sample(float g, V3 wi, float s2x, float s2y, float *pdf) {
    float cos_theta;
    if (fabsf(g) < kEps) cos_theta = 1.0f - 2.0f * s2x;
    else {
        float sqr_term = (1.0f - g * g) / (1.0f - g + 2.0f * g * s2x);
        cos_theta = (1.0f + g * g - sqr_term * sqr_term) / (2.0f * g);
    }
    float sin_theta = sqrtf(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
    float sp, cp; sincos_cephes(2.0f * kPi * s2y, &sp, &cp);
    V3 s, t; coordinate_system(wi, &s, &t);                       // mei.sh_frame = Frame3f(mei.wi)
    V3 local = v3(sin_theta * cp, sin_theta * sp, -cos_theta);
    *pdf = hg_eval(g, -cos_theta);
    return fma3(wi, local.z, fma3(t, local.y, s * local.x));
}

inline int target_medium(const Scene &sc, const Hit &si, V3 d) {   // si.target_medium(d)
    const TrShapeDesc &sh = sc.shapes[si.shape].d;
    return dot(d, si.n) > 0.0f ? sh.exterior_medium : sh.interior_medium;
}
inline bool is_medium_transition(const Scene &sc, const Hit &si) {
    const TrShapeDesc &sh = sc.shapes[si.shape].d;
    return sh.interior_medium >= 0 || sh.exterior_medium >= 0;
}

// sample_emitter   transient_prbvolpath.py:414-512: emitter sample without visibility test, then walk the
// segment through null surfaces / media accumulating transmittance (homogeneous shortcut :463-468).
V3 vol_sample_emitter(const Scene &sc, Pcg32 &rng, bool from_medium, V3 ref_p, const Hit &si, int medium, uint32_t channel,
                      DirectionSample *ds_out, TrCounters *cnt) {
    const uint32_t n_em = (uint32_t) sc.emitters.size();
    float u0 = rng.next_1d(), u1 = rng.next_1d();                                  // :424-425
    ds_out->pdf = 0.0f; ds_out->dist = 0.0f; ds_out->delta = false; ds_out->d = v3(0, 0, 0);
    if (n_em == 0) return v3(0, 0, 0);
    uint32_t ei = 0; float emw = 1.0f, pdf_pick = 1.0f;
    if (n_em > 1) {
        float fi = u0 * (float) n_em;
        ei = (uint32_t) fi; if (ei > n_em - 1) ei = n_em - 1;
        u0 = fi - (float) ei; emw = (float) n_em; pdf_pick = 1.0f / (float) n_em;
    }
    DirectionSample ds;
    V3 emitter_val = area_sample_direction(sc, sc.emitters[ei], ref_p, u0, u1, &ds) * emw;
    ds.pdf *= pdf_pick;
    *ds_out = ds;
    if (ds.pdf == 0.0f) return v3(0, 0, 0);                                        // :429-431
    // A sample that carries no radiance (the light seen from behind: ds.pdf != 0, emitter_val == 0) still walks to the
    // light: the reference only stops on ds.pdf == 0, and the walk draws one sample per medium segment (:456-457), so
    // skipping it would shift the path's random stream.
    if (!from_medium && is_medium_transition(sc, si)) medium = target_medium(sc, si, ds.d);   // :434-435
    Ray ray;
    if (from_medium) {                                                             // Interaction::spawn_ray_to (no offset)
        V3 d = ds.p - ref_p; float dist = sqrtf(dot(d, d));
        ray = Ray{ref_p, d * (1.0f / dist), dist * (1.0f - kShadowEpsilon)};
    } else {
        ray = spawn_ray_to(si.p, si.n, ds.p);
    }
    const float max_dist = ray.maxt; float total_dist = 0.0f;                     // :438-439
    Hit si2; si2.valid = false; si2.t = kInf;
    bool needs_intersection = true, active = true;
    V3 transmittance = v3(1, 1, 1);
    while (active) {                                                               // :444
        float remaining = max_dist - total_dist;
        ray.maxt = remaining;
        active = remaining > 0.0f;
        if (!active) break;
        if (needs_intersection) { cnt->shadow_rays++; si2 = ray_intersect(sc, ray); needs_intersection = false; }
        bool active_medium = medium >= 0, active_surface = !active_medium;
        V3 tr_mult = v3(1, 1, 1);
        bool escaped_medium = false;
        if (active_medium) {                                                       // :456-471
            float u = rng.next_1d(); (void) u;                                     // medium.sample_interaction(ray, next_1d(active_medium), ...)
            MediumEvent me; me.mint = 0.0f; me.sigma_t = medium_sigma_t(sc.media[medium]);
            me.t = fminf(remaining, si2.t);
            V3 pdf; medium_transmittance_eval_pdf(me, si2.t, &tr_mult, &pdf);
            escaped_medium = true; active_medium = false;                          // mei.t = inf: never a real interaction
        }
        active_surface = (active_surface || escaped_medium) && si2.valid;          // :481-482
        if (active_surface) {
            const TrBsdfDesc &b = sc.bsdfs[sc.shapes[si2.shape].d.bsdf];
            float null_tr = (b.type == TR_BSDF_NULL) ? 1.0f : 0.0f;                // bsdf.eval_null_transmission
            tr_mult = tr_mult * null_tr;
        }
        transmittance = transmittance * tr_mult;                                   // :492
        if (active_surface) { ray = spawn_ray(si2.p, si2.n, ray.d); needs_intersection = true; }   // :495-497
        ray.maxt = remaining;
        active = (active_medium || active_surface) && (transmittance.x != 0.0f || transmittance.y != 0.0f || transmittance.z != 0.0f);
        if (active) total_dist += si2.t;                                           // :502
        if (active_surface && is_medium_transition(sc, si2)) medium = target_medium(sc, si2, ray.d);   // :505-506
    }
    return emitter_val * transmittance;
}

V3 trace_volpath(const Scene &sc, const TrRenderParams &rp, Pcg32 &rng, Ray ray, float px, float py, float sample_scale, Film &film) {
    const uint32_t n_em = (uint32_t) sc.emitters.size();
    const bool hide_emitters = (rp.flags & TR_FLAG_HIDE_EMITTERS) != 0;
    uint32_t depth = 0; V3 L = v3(0, 0, 0), beta = v3(1, 1, 1); float eta = 1.0f; bool active = true;
    Hit si; si.valid = false; si.t = kInf; si.p = v3(0, 0, 0); si.n = v3(0, 0, 0); si.shape = 0;
    bool needs_intersection = true;
    V3 last_scatter_p = v3(0, 0, 0); float last_scatter_direction_pdf = 1.0f;     // :150-151
    int medium = -1;                                                               // :154 (sensors inside media unsupported)
    bool specular_chain = true;
    float distance = 0.0f;
    if (rp.flags & TR_FLAG_CAMERA_UNWARP) {                                        // first_surface :515-528
        Ray r = ray; float dsum = 0.0f; bool act = true;
        while (act) {
            Hit h = ray_intersect(sc, r);
            act = h.valid;
            if (!act) break;
            dsum = dsum + h.t;
            act = sc.bsdfs[sc.shapes[h.shape].d.bsdf].type == TR_BSDF_NULL;
            const float e2 = kRayEpsilon * 2.0f;       // newOrigin = si.p + ray.d * RayEpsilon * 2  (:526, Python level: not fused)
            r = Ray{v3(h.p.x + r.d.x * e2, h.p.y + r.d.y * e2, h.p.z + r.d.z * e2), r.d, std::numeric_limits<float>::max()};
        }
        distance = -dsum;
    }
    float cf = 3.0f * rng.next_1d();                                               // :165-168 colour channel for free flights
    uint32_t channel = (uint32_t) fminf(cf, 2.0f);

    while (active) {                                                               // :170
        film.cnt->vertices++;
        active = active && (beta.x != 0.0f || beta.y != 0.0f || beta.z != 0.0f);   // :172
        if (!active) break;
        float q = fminf(max3(beta) * (eta * eta), 0.99f);                          // :176-180 russian roulette
        bool perform_rr = depth > rp.rr_depth;
        float rr_u = rng.next_1d();
        active = active && ((rr_u < q) || !perform_rr);
        if (perform_rr) beta = beta * (1.0f / q);
        if (!active) break;
        bool active_medium = medium >= 0, active_surface = !active_medium;         // :182-183
        bool act_medium_scatter = false, escaped_medium = false;
        MediumEvent mei; mei.valid = false; mei.t = kInf; mei.medium = -1; mei.p = v3(0, 0, 0); mei.wi = v3(0, 0, 0);
        if (active_medium) {                                                       // :189-230
            float u = rng.next_1d();
            mei = medium_sample_interaction(sc, medium, ray, u, channel);
            if (mei.valid) ray.maxt = mei.t;                                       // :194-195
            if (needs_intersection) si = ray_intersect(sc, ray);                   // :196-197
            needs_intersection = false;
            if (si.t < mei.t) mei.t = kInf;                                        // :200
            V3 tr, ffp; medium_transmittance_eval_pdf(mei, si.t, &tr, &ffp);       // :203-209
            float tr_pdf = chan(ffp, channel);
            V3 weight = tr_pdf > 0.0f ? v3(tr.x / tr_pdf, tr.y / tr_pdf, tr.z / tr_pdf) : v3(0, 0, 0);
            mei.valid = mei.t < kInf;
            escaped_medium = !mei.valid; active_medium = mei.valid;                // :211-212
            act_medium_scatter = active_medium;                                    // :225-226
            if (act_medium_scatter) { depth += 1; last_scatter_p = mei.p; distance += mei.t * eta; }   // :228-231
            active = active && (depth < rp.max_depth);                             // :234
            act_medium_scatter = act_medium_scatter && active;
            if (act_medium_scatter) weight = weight * mei.sigma_s;                 // :240-241
            beta = beta * weight;                                                  // :242
        } else {
            active = active && (depth < rp.max_depth);                             // :234 (unconditional in the reference)
        }
        // ---- surface interactions :258-283
        active_surface = active_surface || escaped_medium;
        if (active_surface && needs_intersection) { si = ray_intersect(sc, ray); needs_intersection = false; }
        if (active_surface) distance += si.t * eta;                                // :262
        bool count_direct = (active_surface && depth == 0) || specular_chain;      // :266-267
        if (active_surface && si.valid && sc.shapes[si.shape].d.emitter >= 0 && !(depth == 0 && hide_emitters)) {
            const TrEmitterDesc &em = sc.emitters[sc.shapes[si.shape].d.emitter];
            V3 rel = si.p - last_scatter_p;                                        // DirectionSample3f(scene, si, last_scatter_event)
            float dist = sqrtf(dot(rel, rel));
            V3 dsd = rel * (1.0f / dist);
            float emitter_pdf = area_pdf_direction(sc, em, dsd, si.n, dist);
            if (n_em > 1) emitter_pdf *= 1.0f / (float) n_em;
            float cos_i = -dot(ray.d, si.n);
            V3 emitted = cos_i > 0.0f ? v3(em.radiance[0], em.radiance[1], em.radiance[2]) : v3(0, 0, 0);
            V3 contrib = count_direct ? beta * emitted : (beta * mis_weight(last_scatter_direction_pdf, emitter_pdf)) * emitted;   // :277-278
            L = L + contrib;
            film.add_transient(px, py, contrib, distance, sample_scale, 1.0f);     // :282-283
        }
        active_surface = active_surface && si.valid;                               // :285
        const TrBsdfDesc *bsdf = active_surface ? &sc.bsdfs[sc.shapes[si.shape].d.bsdf] : nullptr;
        const bool smooth = bsdf && bsdf->type == TR_BSDF_DIFFUSE;
        const float cos_i = active_surface ? -dot(ray.d, si.n) : 0.0f;
        // ---- emitter sampling :291-331
        bool active_e_surface = active_surface && smooth && (depth + 1 < rp.max_depth);
        if (act_medium_scatter) specular_chain = false;                            // :295-296 (media sample emitters)
        bool active_e_medium = act_medium_scatter;
        if (active_e_surface || active_e_medium) {
            DirectionSample ds;
            V3 emitted = vol_sample_emitter(sc, rng, active_e_medium, active_e_medium ? mei.p : si.p, si, medium, channel, &ds, film.cnt);
            V3 nee_weight; float nee_pdf;
            if (active_e_surface) {
                float cos_o = dot(ds.d, si.n);
                nee_weight = bsdf_eval(*bsdf, cos_i, cos_o);
                bool two_back = (bsdf->flags & TR_BSDF_FLAG_TWOSIDED) && cos_i < 0.0f;
                float ci = two_back ? -cos_i : cos_i, co = two_back ? -cos_o : cos_o;
                nee_pdf = (ci > 0.0f && co > 0.0f) ? kInvPi * co : 0.0f;
            } else {
                float ph = hg_eval(sc.media[mei.medium].g, dot(ds.d, mei.wi));     // phase.eval_pdf(mei, ds.d)
                nee_weight = v3(ph, ph, ph); nee_pdf = ph;
            }
            V3 contrib = ((beta * nee_weight) * mis_weight(ds.pdf, nee_pdf)) * emitted;   // :322-323
            L = L + contrib;
            film.add_transient(px, py, contrib, distance + ds.dist * eta, sample_scale, 1.0f);   // :329-331
        }
        // ---- phase function sampling :335-360
        if (act_medium_scatter) {
            float p1 = rng.next_1d(); (void) p1;
            float p2x = rng.next_1d(), p2y = rng.next_1d();
            float ppdf;
            V3 wo = hg_sample(sc.media[mei.medium].g, mei.wi, p2x, p2y, &ppdf);
            act_medium_scatter = ppdf > 0.0f;
            if (act_medium_scatter) {                                              // phase_weight = 1
                ray = Ray{mei.p, wo, std::numeric_limits<float>::max()};           // mei.spawn_ray(wo)
                needs_intersection = true;
                last_scatter_direction_pdf = ppdf;
            }
        }
        // ---- BSDF sampling :364-408
        if (active_surface) {
            float b1 = rng.next_1d();
            float b2x = rng.next_1d(), b2y = rng.next_1d();
            const bool is_null = bsdf->type == TR_BSDF_NULL;
            // bsdf.sample in the local frame (null.cpp: bs.wo = -si.wi, so the world direction is a frame round trip of ray.d)
            V3 wi = v3(-dot(ray.d, si.s), -dot(ray.d, si.tt), -dot(ray.d, si.n));
            BsdfSampleRec bs = bsdf_sample(*bsdf, wi, b1, b2x, b2y);
            V3 bw = bs.weight; float bs_pdf = bs.pdf;
            V3 wo_world = fma3(si.n, bs.wo.z, fma3(si.tt, bs.wo.y, si.s * bs.wo.x));   // si.to_world(bs.wo)
            active_surface = bs_pdf > 0.0f;                                        // :370
            if (active_surface) {
                beta = beta * bw;                                                  // :386
                eta = eta * bs.eta;                                                // :387
                ray = spawn_ray(si.p, si.n, wo_world);                             // :388-389
                needs_intersection = true;
                if (!is_null) {                                                    // non_null_bsdf :392-399
                    depth += 1; last_scatter_p = si.p; last_scatter_direction_pdf = bs_pdf;
                    if (bs.delta) specular_chain = true;                           // :402-403 sampled_type has Delta
                }
                if (!is_null && !bs.delta) specular_chain = false;                 // :404-405 sampled_type has Smooth (diffuse lobes)
                if (is_medium_transition(sc, si)) medium = target_medium(sc, si, ray.d);   // :406-407
            }
        }
        active = active && (active_surface || active_medium);                      // :408
    }
    return L;
}

struct RenderJob {
    const Scene *sc; const TrRenderParams *rp; float *film; float *steady; int C;
    std::atomic<uint32_t> next_row{0};
    std::mutex mu; TrCounters total{}; std::vector<TrDeposit> *log = nullptr; uint64_t log_cap = 0;
    const NlosCtx *nlos = nullptr;
};

inline bool pixel_owned(const TrRenderParams &rp, uint32_t W, uint32_t x, uint32_t y) {
    if (rp.world <= 1) return true;
    uint32_t ts = rp.tile_size ? rp.tile_size : 32;
    uint32_t tiles_x = (W + ts - 1) / ts;
    uint32_t tile = (y / ts) * tiles_x + (x / ts);
    return (tile + y / ts) % rp.world == rp.rank;     // round-robin, skewed by one tile per tile row
}

void render_rows(RenderJob *job) {
    const Scene &sc = *job->sc; const TrRenderParams &rp = *job->rp;
    const uint32_t W = sc.film.width, H = sc.film.height;
    TrCounters cnt{}; std::vector<TrDeposit> local_log;
    Film film{sc.film, job->film, job->C, &cnt, job->log ? &local_log : nullptr};
    const float sample_scale = 1.0f / (float) rp.total_spp;        // common.py:173
    for (;;) {
        uint32_t y = job->next_row.fetch_add(1);
        if (y >= H) break;
        for (uint32_t x = 0; x < W; ++x) {
            if (!pixel_owned(rp, W, x, y)) continue;
            V3 Lsum = v3(0, 0, 0);
            for (uint32_t s = 0; s < rp.spp; ++s) {
                // lane index of this sample in the reference's wavefront (idx // spp == pixel)
                uint32_t idx = (y * W + x) * rp.spp + s;
                uint32_t v0, v1; sample_tea_32(rp.seed, idx, &v0, &v1);
                Pcg32 rng; rng.seed(v0, v1);
                float jx = rng.next_1d(), jy = rng.next_1d();      // pos_f = pos + sampler.next_2d()
                cnt.paths++;
                V3 L;
                if (rp.integrator == TR_INTEGRATOR_TRANSIENT_PRBVOLPATH) {
                    L = trace_volpath(sc, rp, rng, sample_camera_ray(sc, x, y, jx, jy), (float) x, (float) y, sample_scale, film);
                } else if (job->nlos) {
                    float sx = ((float) x + jx) * (1.0f / (float) W), sy = ((float) y + jy) * (1.0f / (float) H);
                    Ray sray = sc.sensor.type == TR_SENSOR_PERSPECTIVE ? sample_camera_ray(sc, x, y, jx, jy) : nlos_sensor_ray(sc, sx, sy);
                    L = trace_nlos_path(sc, rp, *job->nlos, rng, sray, y * W + x, (float) x, (float) y, sample_scale, film);
                } else {
                    Ray ray = sample_camera_ray(sc, x, y, jx, jy);
                    L = trace_transient_path(sc, rp, rng, ray, (float) x, (float) y, sample_scale, 1.0f, film);
                }
                Lsum = Lsum + L;                                   // steady block.put (weight 1)  common.py:196-203
            }
            if (job->steady) {
                float *o = job->steady + 3 * ((size_t) y * W + x);
                float inv = 1.0f / (float) rp.total_spp;
                if (rp.accumulate) { o[0] += Lsum.x * inv; o[1] += Lsum.y * inv; o[2] += Lsum.z * inv; }
                else { o[0] = Lsum.x * inv; o[1] = Lsum.y * inv; o[2] = Lsum.z * inv; }
            }
        }
    }
    std::lock_guard<std::mutex> g(job->mu);
    job->total.paths += cnt.paths; job->total.vertices += cnt.vertices; job->total.deposits += cnt.deposits;
    job->total.deposits_in_range += cnt.deposits_in_range; job->total.bin_checksum += cnt.bin_checksum;
    job->total.shadow_rays += cnt.shadow_rays;
    if (job->log) for (auto &d : local_log) if (job->log->size() < job->log_cap) job->log->push_back(d);
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char *oracle_last_error(void) { return g_err.c_str(); }

// Render into host buffers.  film: float[H][W][T][film_channels] (NOT cleared unless
// params->accumulate == 0, in which case owned pixels' columns are zeroed first).
// deposit_log (optional): receives up to log_capacity deposits in thread-completion order;
// *n_logged gets the count.  n_threads <= 0 -> hardware concurrency.
int oracle_render(const TrSceneDesc *desc, const TrRenderParams *params, float *film, float *steady,
                  TrCounters *counters, int n_threads, TrDeposit *deposit_log, uint64_t log_capacity,
                  uint64_t *n_logged) {
    Scene sc; std::string err;
    if (!params || !film) { g_err = "null argument"; return TR_ERR_INVALID_ARGUMENT; }
    if (!build_scene(desc, sc, err)) { g_err = err; return TR_ERR_INVALID_ARGUMENT; }
    NlosCtx nctx;
    if (params->integrator == TR_INTEGRATOR_TRANSIENT_NLOS_PATH) {
        if (!nlos_prepare(sc, *params, nctx, err)) { g_err = err; return TR_ERR_UNSUPPORTED; }
    } else if (params->integrator != TR_INTEGRATOR_TRANSIENT_PATH && params->integrator != TR_INTEGRATOR_TRANSIENT_PRBVOLPATH) {
        g_err = "oracle: integrator not implemented"; return TR_ERR_UNSUPPORTED;
    } else if (sc.sensor.type != TR_SENSOR_PERSPECTIVE) {
        g_err = "oracle: transient_path / transient_prbvolpath need a perspective sensor"; return TR_ERR_UNSUPPORTED;
    }
    const bool phasor = sc.film.type == TR_FILM_PHASOR_HDR;
    if (phasor) {
        if (!sc.film.frequencies || sc.film.n_frequencies == 0) { g_err = "phasor film without frequencies"; return TR_ERR_INVALID_ARGUMENT; }
        if (params->film_channels != (int) (2 * sc.film.n_frequencies + 1)) { g_err = "phasor film: film_channels must be 2F + 1"; return TR_ERR_INVALID_ARGUMENT; }
    } else if (params->film_channels != 3 && params->film_channels != 4) { g_err = "film_channels must be 3 or 4"; return TR_ERR_INVALID_ARGUMENT; }
    if (params->spp == 0 || params->total_spp == 0) { g_err = "spp must be > 0"; return TR_ERR_INVALID_ARGUMENT; }
    const uint32_t W = sc.film.width, H = sc.film.height, T = sc.film.temporal_bins;
    const int C = params->film_channels;
    // floats per pixel: T*C (plain), Lw*Lh*T*C (exhaustive_scan), 2F+1 (phasor)
    const size_t n_laser = (sc.film.laser_scan_width && sc.film.laser_scan_height) ? (size_t) sc.film.laser_scan_width * sc.film.laser_scan_height : 1;
    const size_t pixel_floats = phasor ? (size_t) C : n_laser * T * C;
    if (!params->accumulate) {
        for (uint32_t y = 0; y < H; ++y) for (uint32_t x = 0; x < W; ++x)
            if (pixel_owned(*params, W, x, y))
                memset(film + ((size_t) y * W + x) * pixel_floats, 0, sizeof(float) * pixel_floats);
    }
    RenderJob job; job.sc = &sc; job.rp = params; job.film = film; job.steady = steady; job.C = C;
    if (params->integrator == TR_INTEGRATOR_TRANSIENT_NLOS_PATH) job.nlos = &nctx;
    std::vector<TrDeposit> log;
    if (deposit_log && log_capacity) { job.log = &log; job.log_cap = log_capacity; log.reserve((size_t) std::min<uint64_t>(log_capacity, 1u << 22)); }
    int nt = n_threads > 0 ? n_threads : (int) std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::vector<std::thread> th;
    for (int i = 0; i < nt; ++i) th.emplace_back(render_rows, &job);
    for (auto &t : th) t.join();
    if (counters) *counters = job.total;
    if (deposit_log) { memcpy(deposit_log, log.data(), log.size() * sizeof(TrDeposit)); if (n_logged) *n_logged = log.size(); }
    return TR_OK;
}

// TransientImageBlock.put_ for a recorded deposit stream (sequential fp32 adds, stream order).
int oracle_film_scatter_add(float *film, const TrFilmDesc *fd, int channels, const TrDeposit *dep, uint64_t n) {
    if (!film || !fd || !dep) { g_err = "null argument"; return TR_ERR_INVALID_ARGUMENT; }
    TrCounters cnt{};
    Film f{*fd, film, channels, &cnt, nullptr};
    // recorded positions are film coordinates: the block subtracts its offset, floor(pos) - offset (:131)
    const float ox = (float) fd->crop_offset_x, oy = (float) fd->crop_offset_y;
    for (uint64_t i = 0; i < n; ++i)
        f.add_transient(floorf(dep[i].x) - ox, floorf(dep[i].y) - oy, v3(dep[i].r, dep[i].g, dep[i].b), dep[i].opl, 1.0f, 1.0f);
    return TR_OK;
}

// Same with the illuminated-point coordinates of an exhaustive_scan film (transient_image_block.py:134-140);
// laser_xy: uint32[n][2].
int oracle_film_scatter_add_laser(float *film, const TrFilmDesc *fd, int channels, const TrDeposit *dep, const uint32_t *laser_xy,
                                  uint64_t n) {
    if (!film || !fd || !dep || !laser_xy) { g_err = "null argument"; return TR_ERR_INVALID_ARGUMENT; }
    TrCounters cnt{};
    Film f{*fd, film, channels, &cnt, nullptr};
    for (uint64_t i = 0; i < n; ++i)
        f.add_transient(dep[i].x, dep[i].y, v3(dep[i].r, dep[i].g, dep[i].b), dep[i].opl, 1.0f, 1.0f, laser_xy[2 * i], laser_xy[2 * i + 1]);
    return TR_OK;
}

// TransientHDRFilm.develop_transient_ (transient_hdr_film.py:220-248)
int oracle_film_develop(const float *raw, float *out, uint64_t n_bins, int in_channels) {
    if (!raw || !out || in_channels < 2) { g_err = "bad argument"; return TR_ERR_INVALID_ARGUMENT; }
    int oc = in_channels - 1;
    for (uint64_t i = 0; i < n_bins; ++i) {
        float w = raw[i * in_channels + in_channels - 1];
        float div = (w == 0.0f) ? 1.0f : w;
        for (int c = 0; c < oc; ++c) out[i * oc + c] = raw[i * in_channels + c] / div;
    }
    return TR_OK;
}

// ---- small probes used by the known-answer tests ---------------------------------------
void oracle_pcg32(uint64_t initstate, uint64_t initseq, uint32_t *out, int n) {
    Pcg32 r; r.seed(initstate, initseq);
    for (int i = 0; i < n; ++i) out[i] = r.next_u32();
}
void oracle_tea32(uint32_t v0, uint32_t v1, uint32_t *o0, uint32_t *o1) { sample_tea_32(v0, v1, o0, o1); }
void oracle_sampler_floats(uint32_t seed, uint32_t lane, float *out, int n) {
    uint32_t v0, v1; sample_tea_32(seed, lane, &v0, &v1);
    Pcg32 r; r.seed(v0, v1);
    for (int i = 0; i < n; ++i) out[i] = r.next_1d();
}
void oracle_sincos(float x, float *s, float *c) { sincos_cephes(x, s, c); }
float oracle_fresnel_conductor(float cos_theta_i, float eta, float k) { return fresnel_conductor(cos_theta_i, eta, k); }
void oracle_fresnel_dielectric(float cos_theta_i, float eta, float *out4) {
    fresnel_dielectric(cos_theta_i, eta, out4, out4 + 1, out4 + 2, out4 + 3);
}
float oracle_exp(float x) { return exp_cephes(x); }
float oracle_log(float x) { return log_cephes(x); }
void oracle_cosine_hemisphere(float u0, float u1, float *xyz) {
    V3 w = square_to_cosine_hemisphere(u0, u1); xyz[0] = w.x; xyz[1] = w.y; xyz[2] = w.z;
}
// camera ray for pixel (px,py) with jitter (jx,jy): o[3], d[3], maxt
int oracle_camera_ray(const TrSceneDesc *desc, uint32_t px, uint32_t py, float jx, float jy, float *o, float *d, float *maxt) {
    Scene sc; std::string err;
    if (!build_scene(desc, sc, err)) { g_err = err; return TR_ERR_INVALID_ARGUMENT; }
    Ray r = sample_camera_ray(sc, px, py, jx, jy);
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; d[0] = r.d.x; d[1] = r.d.y; d[2] = r.d.z; *maxt = r.maxt;
    return TR_OK;
}
// closest hit of an arbitrary ray: returns 1 on hit; out = t, p[3], n[3], shape, prim
int oracle_ray_intersect(const TrSceneDesc *desc, const float *o, const float *d, float maxt, float *out_t, float *out_p,
                         float *out_n, uint32_t *out_shape, uint32_t *out_prim) {
    Scene sc; std::string err;
    if (!build_scene(desc, sc, err)) { g_err = err; return TR_ERR_INVALID_ARGUMENT; }
    Hit h = ray_intersect(sc, Ray{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), maxt});
    if (!h.valid) return 0;
    *out_t = h.t; out_p[0] = h.p.x; out_p[1] = h.p.y; out_p[2] = h.p.z;
    out_n[0] = h.n.x; out_n[1] = h.n.y; out_n[2] = h.n.z; *out_shape = h.shape; *out_prim = h.prim;
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Scene services ("osv_*"): the Mitsuba-level primitives the path loops above are built from -- scene.ray_intersect /
// ray_test, si.spawn_ray(_to) / to_local / to_world, Scene::sample_emitter_direction / pdf_emitter_direction,
// Emitter::eval, BSDF::flags / eval_pdf / sample, the perspective sensor -- exported lane-wise so that the
// REFERENCE'S OWN integrator code (transientpath.py ...) can be executed over them (tests/golden/refshim.py).  Each
// entry calls exactly the function the trace loops call, so both walk bit-identical arithmetic.
struct OsvScene { Scene sc; };

void *osv_open(const TrSceneDesc *desc) {
    OsvScene *h = new OsvScene(); std::string err;
    if (!build_scene(desc, h->sc, err)) { g_err = err; delete h; return nullptr; }
    return h;
}
void osv_close(void *h) { delete (OsvScene *) h; }

inline V3 ld3(const float *a, uint64_t i) { return v3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
inline void st3(float *a, uint64_t i, V3 v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }

void osv_camera_rays(void *h, uint64_t n, const uint32_t *px, const uint32_t *py, const float *jitter, float *o, float *d, float *maxt) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        Ray r = sample_camera_ray(sc, px[i], py[i], jitter[2 * i], jitter[2 * i + 1]);
        st3(o, i, r.o); st3(d, i, r.d); maxt[i] = r.maxt;
    }
}
// scene.ray_intersect: valid, t (inf on a miss), p, n, shading frame (s, tt), shape index, wi = to_local(-d)
void osv_ray_intersect(void *h, uint64_t n, const float *o, const float *d, const float *maxt, uint8_t *valid, float *t, float *p,
                       float *nrm, float *s, float *tt, int32_t *shape, float *wi) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        V3 rd = ld3(d, i);
        Hit hit = ray_intersect(sc, Ray{ld3(o, i), rd, maxt[i]});
        valid[i] = hit.valid; t[i] = hit.t; shape[i] = hit.valid ? (int32_t) hit.shape : -1;
        st3(p, i, hit.valid ? hit.p : v3(0, 0, 0)); st3(nrm, i, hit.valid ? hit.n : v3(0, 0, 0));
        st3(s, i, hit.valid ? hit.s : v3(0, 0, 0)); st3(tt, i, hit.valid ? hit.tt : v3(0, 0, 0));
        st3(wi, i, hit.valid ? v3(-dot(rd, hit.s), -dot(rd, hit.tt), -dot(rd, hit.n)) : v3(0, 0, 0));
    }
}
void osv_ray_test(void *h, uint64_t n, const float *o, const float *d, const float *maxt, uint8_t *hit) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) hit[i] = ray_test(sc, Ray{ld3(o, i), ld3(d, i), maxt[i]});
}
void osv_spawn_ray(void *, uint64_t n, const float *p, const float *nrm, const float *dir, float *o, float *maxt) {
    for (uint64_t i = 0; i < n; ++i) { Ray r = spawn_ray(ld3(p, i), ld3(nrm, i), ld3(dir, i)); st3(o, i, r.o); maxt[i] = r.maxt; }
}
void osv_spawn_ray_to(void *, uint64_t n, const float *p, const float *nrm, const float *target, float *o, float *d, float *maxt) {
    for (uint64_t i = 0; i < n; ++i) { Ray r = spawn_ray_to(ld3(p, i), ld3(nrm, i), ld3(target, i)); st3(o, i, r.o); st3(d, i, r.d); maxt[i] = r.maxt; }
}
void osv_to_world(void *, uint64_t n, const float *s, const float *tt, const float *nrm, const float *v, float *out) {
    for (uint64_t i = 0; i < n; ++i) { V3 w = ld3(v, i); st3(out, i, fma3(ld3(nrm, i), w.z, fma3(ld3(tt, i), w.y, ld3(s, i) * w.x))); }
}
void osv_to_local(void *, uint64_t n, const float *s, const float *tt, const float *nrm, const float *v, float *out) {
    for (uint64_t i = 0; i < n; ++i) { V3 w = ld3(v, i); st3(out, i, v3(dot(w, ld3(s, i)), dot(w, ld3(tt, i)), dot(w, ld3(nrm, i)))); }
}
// DirectionSample3f(scene, si, ref): d = normalize(si.p - ref.p), dist
void osv_direction_to(void *, uint64_t n, const float *p, const float *ref, float *d, float *dist) {
    for (uint64_t i = 0; i < n; ++i) {
        V3 rel = ld3(p, i) - ld3(ref, i);
        float l = sqrtf(dot(rel, rel));
        st3(d, i, rel * (1.0f / l)); dist[i] = l;
    }
}
// scene.sample_emitter_direction(ref, sample, test_visibility, active) -> ds (p, n, d, dist, pdf), weight
void osv_sample_emitter_direction(void *h, uint64_t n, const float *ref_p, const float *ref_n, const float *u, int test_visibility,
                                  const uint8_t *active, float *ds_p, float *ds_n, float *ds_d, float *ds_dist, float *ds_pdf,
                                  float *weight) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        DirectionSample ds; ds.p = ds.n = ds.d = v3(0, 0, 0); ds.dist = ds.pdf = 0.0f; ds.delta = false;
        V3 w = v3(0, 0, 0);
        if (active[i] && !sc.emitters.empty()) {
            uint32_t ei;
            w = scene_sample_emitter_direction(sc, ld3(ref_p, i), u[2 * i], u[2 * i + 1], &ds, &ei);
            bool nz = w.x != 0.0f || w.y != 0.0f || w.z != 0.0f;
            if (test_visibility && nz && ray_test(sc, spawn_ray_to(ld3(ref_p, i), ld3(ref_n, i), ds.p))) w = v3(0, 0, 0);
        }
        st3(ds_p, i, ds.p); st3(ds_n, i, ds.n); st3(ds_d, i, ds.d); ds_dist[i] = ds.dist; ds_pdf[i] = ds.pdf; st3(weight, i, w);
    }
}
// scene.pdf_emitter_direction(ref, ds, active) for the emitter attached to `shape` (the shape that was hit)
void osv_pdf_emitter_direction(void *h, uint64_t n, const int32_t *shape, const float *ds_d, const float *ds_n, const float *ds_dist,
                               const uint8_t *active, float *pdf) {
    const Scene &sc = ((OsvScene *) h)->sc;
    const uint32_t n_em = (uint32_t) sc.emitters.size();
    for (uint64_t i = 0; i < n; ++i) {
        pdf[i] = 0.0f;
        if (!active[i] || shape[i] < 0 || sc.shapes[shape[i]].d.emitter < 0) continue;
        float v = area_pdf_direction(sc, sc.emitters[sc.shapes[shape[i]].d.emitter], ld3(ds_d, i), ld3(ds_n, i), ds_dist[i]);
        if (n_em > 1) v *= 1.0f / (float) n_em;
        pdf[i] = v;
    }
}
// si.emitter(scene).eval(si, active): area light radiance on the front side (cos_theta(si.wi) > 0)
void osv_emitter_eval(void *h, uint64_t n, const int32_t *shape, const float *wi, const uint8_t *active, float *rgb) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        V3 r = v3(0, 0, 0);
        if (active[i] && shape[i] >= 0 && sc.shapes[shape[i]].d.emitter >= 0)
            r = emitter_eval(sc.emitters[sc.shapes[shape[i]].d.emitter], ld3(wi, i));
        st3(rgb, i, r);
    }
}
void osv_bsdf_smooth(void *h, uint64_t n, const int32_t *shape, uint8_t *smooth, uint8_t *is_null) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        const TrBsdfDesc *b = shape[i] >= 0 ? &sc.bsdfs[sc.shapes[shape[i]].d.bsdf] : nullptr;
        smooth[i] = b && bsdf_is_smooth(*b); is_null[i] = b && b->type == TR_BSDF_NULL;
    }
}
void osv_bsdf_eval_pdf(void *h, uint64_t n, const int32_t *shape, const float *wi, const float *wo, const uint8_t *active, float *value,
                       float *pdf) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        V3 v = v3(0, 0, 0); pdf[i] = 0.0f;
        if (active[i] && shape[i] >= 0) v = bsdf_eval_pdf(sc.bsdfs[sc.shapes[shape[i]].d.bsdf], ld3(wi, i), ld3(wo, i), &pdf[i]);
        st3(value, i, v);
    }
}
void osv_bsdf_sample(void *h, uint64_t n, const int32_t *shape, const float *wi, const float *s1, const float *s2, const uint8_t *active,
                     float *wo, float *pdf, float *eta, uint8_t *delta, float *weight) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        BsdfSampleRec r; r.wo = r.weight = v3(0, 0, 0); r.pdf = 0.0f; r.eta = 1.0f; r.delta = false;
        if (active[i] && shape[i] >= 0) r = bsdf_sample(sc.bsdfs[sc.shapes[shape[i]].d.bsdf], ld3(wi, i), s1[i], s2[2 * i], s2[2 * i + 1]);
        st3(wo, i, r.wo); st3(weight, i, r.weight); pdf[i] = r.pdf; eta[i] = r.eta; delta[i] = r.delta;
    }
}

// ---- services used by transient_nlos_path (transientnlospath.py) and nlos_capture_meter (nloscapturemeter.py)
// emitter.sample_direction(it, sample, active) for the projector: ds.p, ds.d, ds.dist, ds.pdf (= 1 when active), weight
void osv_projector_sample_direction(void *h, uint64_t n, const float *it_p, const uint8_t *active, float *ds_p, float *ds_d, float *ds_dist,
                                    float *ds_pdf, float *weight) {
    const Scene &sc = ((OsvScene *) h)->sc;
    const TrEmitterDesc &em = sc.emitters[0];
    for (uint64_t i = 0; i < n; ++i) {
        V3 p = v3(0, 0, 0), w = v3(0, 0, 0); float dist = 0.0f;
        V3 itp = ld3(it_p, i);
        // projector.cpp fills the DirectionSample without looking at `active`; only the returned spectrum is masked
        w = projector_sample_direction(em, itp, &p, &dist);
        if (!active[i]) w = v3(0, 0, 0);
        V3 d = p - itp; float l = sqrtf(dot(d, d));
        st3(ds_p, i, p); st3(ds_d, i, d * (1.0f / l)); ds_dist[i] = dist; st3(weight, i, w);
        ds_pdf[i] = 1.0f;                                            // projector.cpp: ds.pdf = 1 (delta position)
    }
}
// shape.sample_position(time, sample, active) for the shape index of every lane -> p, n, pdf
void osv_shape_sample_position(void *h, uint64_t n, const int32_t *shape, const float *u, const uint8_t *active, float *p, float *nrm, float *pdf) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        V3 pp = v3(0, 0, 0), pn = v3(0, 0, 0); float ppdf = 0.0f;
        if (active[i] && shape[i] >= 0) shape_sample_position(sc, sc.shapes[shape[i]], u[2 * i], u[2 * i + 1], &pp, &pn, &ppdf);
        st3(p, i, pp); st3(nrm, i, pn); pdf[i] = ppdf;
    }
}
// ---- services used by transient_prbvolpath (transient_prbvolpath.py): homogeneous media, HG phase function, medium transitions
void osv_medium_sample_interaction(void *h, uint64_t n, const int32_t *medium, const float *o, const float *d, const float *maxt,
                                   const float *u, const int32_t *channel, const uint8_t *active, uint8_t *valid, float *t, float *p,
                                   float *sigma_t, float *sigma_s) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        valid[i] = 0; t[i] = kInf; st3(p, i, v3(0, 0, 0)); st3(sigma_t, i, v3(0, 0, 0)); st3(sigma_s, i, v3(0, 0, 0));
        if (!active[i] || medium[i] < 0) continue;
        MediumEvent me = medium_sample_interaction(sc, medium[i], Ray{ld3(o, i), ld3(d, i), maxt[i]}, u[i], (uint32_t) channel[i]);
        valid[i] = me.valid; t[i] = me.t; st3(p, i, me.p); st3(sigma_t, i, me.sigma_t); st3(sigma_s, i, me.sigma_s);
    }
}
void osv_medium_transmittance(void *, uint64_t n, const float *mei_t, const float *sigma_t, const float *si_t, const uint8_t *active, float *tr,
                              float *pdf) {
    for (uint64_t i = 0; i < n; ++i) {
        V3 a = v3(0, 0, 0), b = v3(0, 0, 0);
        if (active[i]) { MediumEvent me; me.t = mei_t[i]; me.mint = 0.0f; me.sigma_t = ld3(sigma_t, i); medium_transmittance_eval_pdf(me, si_t[i], &a, &b); }
        st3(tr, i, a); st3(pdf, i, b);
    }
}
void osv_phase_eval(void *h, uint64_t n, const int32_t *medium, const float *wi, const float *wo, const uint8_t *active, float *val) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) val[i] = (active[i] && medium[i] >= 0) ? hg_eval(sc.media[medium[i]].g, dot(ld3(wo, i), ld3(wi, i))) : 0.0f;
}
void osv_phase_sample(void *h, uint64_t n, const int32_t *medium, const float *wi, const float *s2, const uint8_t *active, float *wo, float *pdf) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        V3 w = v3(0, 0, 0); pdf[i] = 0.0f;
        if (active[i] && medium[i] >= 0) w = hg_sample(sc.media[medium[i]].g, ld3(wi, i), s2[2 * i], s2[2 * i + 1], &pdf[i]);
        st3(wo, i, w);
    }
}
void osv_target_medium(void *h, uint64_t n, const int32_t *shape, const float *nrm, const float *d, int32_t *out, uint8_t *transition) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) {
        out[i] = -1; transition[i] = 0;
        if (shape[i] < 0) continue;
        Hit si; si.shape = (uint32_t) shape[i]; si.n = ld3(nrm, i);
        out[i] = target_medium(sc, si, ld3(d, i)); transition[i] = is_medium_transition(sc, si);
    }
}
void osv_shape_emitter(void *h, uint64_t n, const int32_t *shape, int32_t *emitter) {
    const Scene &sc = ((OsvScene *) h)->sc;
    for (uint64_t i = 0; i < n; ++i) emitter[i] = shape[i] >= 0 ? sc.shapes[shape[i]].d.emitter : -1;
}
int32_t osv_sensor_medium(void *h) { return ((OsvScene *) h)->sc.sensor.medium; }
void osv_acos(void *, uint64_t n, const float *x, float *out) { for (uint64_t i = 0; i < n; ++i) out[i] = acos_cephes(x[i]); }
void osv_frame_to_local(void *, uint64_t n, const float *nrm, const float *v, float *out) {
    for (uint64_t i = 0; i < n; ++i) st3(out, i, frame_to_local(ld3(nrm, i), ld3(v, i)));
}
// Shape::sample_direction of the shape carrying emitter 0 (area pdf -> solid angle) and Shape::pdf_direction
void osv_shape_sample_direction(void *h, uint64_t n, const float *ref_p, const float *u, const uint8_t *active, float *ds_p, float *ds_n,
                                float *ds_d, float *ds_dist, float *ds_pdf) {
    const Scene &sc = ((OsvScene *) h)->sc;
    TrEmitterDesc plain = sc.emitters[0]; plain.type = TR_EMITTER_AREA;
    for (uint64_t i = 0; i < n; ++i) {
        DirectionSample ds; ds.p = ds.n = ds.d = v3(0, 0, 0); ds.dist = ds.pdf = 0.0f;
        if (active[i]) area_sample_direction(sc, plain, ld3(ref_p, i), u[2 * i], u[2 * i + 1], &ds);
        st3(ds_p, i, ds.p); st3(ds_n, i, ds.n); st3(ds_d, i, ds.d); ds_dist[i] = ds.dist; ds_pdf[i] = ds.pdf;
    }
}
void osv_shape_pdf_direction(void *h, uint64_t n, const float *ds_d, const float *ds_n, const float *ds_dist, const uint8_t *active, float *pdf) {
    const Scene &sc = ((OsvScene *) h)->sc;
    TrEmitterDesc plain = sc.emitters[0]; plain.type = TR_EMITTER_AREA;
    for (uint64_t i = 0; i < n; ++i) pdf[i] = active[i] ? area_pdf_direction(sc, plain, ld3(ds_d, i), ld3(ds_n, i), ds_dist[i]) : 0.0f;
}
void osv_dot(void *, uint64_t n, const float *a, const float *b, float *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = dot(ld3(a, i), ld3(b, i));
}
// Projector::sample_ray [upstream] of a copy of the scene's projector whose sample -> camera matrix is s2c: one ray per (u, v)
void osv_projector_sample_ray(void *h, uint64_t n, const float *s2c, const float *uv, float *o, float *d) {
    const Scene &sc = ((OsvScene *) h)->sc;
    const TrEmitterDesc &em = sc.emitters[0];
    for (uint64_t i = 0; i < n; ++i) {
        const float *m = s2c; float u = uv[2 * i], v = uv[2 * i + 1];
        float cx = fmaf(m[0], u, fmaf(m[1], v, m[3])), cy = fmaf(m[4], u, fmaf(m[5], v, m[7]));
        float cz = fmaf(m[8], u, fmaf(m[9], v, m[11])), cw = fmaf(m[12], u, fmaf(m[13], v, m[15]));
        float iw = 1.0f / cw;
        V3 dl = normalize(v3(cx * iw, cy * iw, cz * iw));
        st3(o, i, v3(em.to_world[3], em.to_world[7], em.to_world[11])); st3(d, i, xf_vector(em.to_world, dl));
    }
}
uint32_t osv_n_shapes(void *h) { return (uint32_t) ((OsvScene *) h)->sc.shapes.size(); }
float osv_shape_area(void *h, uint32_t shape) { return ((OsvScene *) h)->sc.shapes[shape].area; }
int32_t osv_sensor_shape(void *h) { const Scene &sc = ((OsvScene *) h)->sc; return sc.sensor.type == TR_SENSOR_NLOS_CAPTURE_METER ? sc.sensor.shape : -1; }
// DiscreteDistribution(pmf).sample_reuse_pmf(u) with the oracle's arithmetic (nlos_prepare + trace_nlos_path)
void osv_discrete_sample_reuse_pmf(void *, uint64_t n, const float *pmf_in, uint32_t n_entries, const float *u, int32_t *index, float *u_new,
                                   float *pmf_out) {
    std::vector<float> cdf(n_entries); double acc = 0.0;
    for (uint32_t k = 0; k < n_entries; ++k) { acc += (double) pmf_in[k]; cdf[k] = (float) acc; }
    const float sum = (float) acc, norm = (float) (1.0 / acc);
    for (uint64_t i = 0; i < n; ++i) {
        float value = u[i] * sum;
        uint32_t lo = 0, hi = n_entries - 1;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cdf[mid] < value) lo = mid + 1; else hi = mid; }
        float cdf_prev = lo ? cdf[lo - 1] : 0.0f;
        float pmf = pmf_in[lo] * norm;
        index[i] = (int32_t) lo; pmf_out[i] = pmf; u_new[i] = (u[i] - cdf_prev * norm) / pmf;
    }
}

}  // extern "C"
