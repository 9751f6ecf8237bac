// Device-side maths for the transient path tracer (sm_100a).
//
// ARITHMETIC CONTRACT (DESIGN.md "Arithmetic contract"): IEEE fp32, compiled with -fmad=false so
// nothing is contracted implicitly; every fused multiply-add is spelled fmaf() here.  Division
// and sqrt are the correctly rounded ones (nvcc defaults -prec-div=true -prec-sqrt=true, no
// --use_fast_math).  The CPU oracle follows the same contract by specification, which makes the
// walk of every path bit-identical between the two; only the order of additions into a film
// bin differs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tr {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ V3 fma3(V3 a, float s, V3 c) { return v3(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z)); }
__device__ __forceinline__ V3 normalize(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
__device__ __forceinline__ float max3(V3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
__device__ __forceinline__ float mulsign(float a, float b) {
    return __uint_as_float(__float_as_uint(a) ^ (__float_as_uint(b) & 0x80000000u));
}
__device__ __forceinline__ V3 xyz(float4 v) { return v3(v.x, v.y, v.z); }

#define TR_PI 3.14159265358979323846f
#define TR_INV_PI 0.31830988618379067154f
#define TR_RAY_EPSILON 8.940696716308594e-05f
#define TR_SHADOW_EPSILON 8.940696716308594e-04f
#define TR_FLT_MAX 3.402823466e+38f

// Cephes single-precision sincos (same polynomial as the oracle).
__device__ __forceinline__ void sincos_cephes(float x, float *s_out, float *c_out) {
    float xa = fabsf(x);
    int j = (int) (xa * 1.2732395447351628f);
    j = (j + 1) & ~1;
    float y = (float) j;
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

// Cephes single-precision exp / log (same polynomials and fmaf placement as the oracle).
__device__ __forceinline__ float exp_cephes(float x) {
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
    return p * __uint_as_float((uint32_t) ((int) n + 127) << 23);
}
__device__ __forceinline__ float log_cephes(float x) {
    uint32_t bits = __float_as_uint(x);
    int e = (int) ((bits >> 23) & 0xffu) - 126;
    float m = __uint_as_float((bits & 0x007fffffu) | 0x3f000000u);
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

// Mitsuba `independent` sampler: PCG32 per lane seeded through TEA from (seed, lane index).
struct Pcg32 {
    uint64_t state, inc;
    __device__ __forceinline__ void advance() { state = state * 0x5851f42d4c957f2dULL + inc; }
    __device__ __forceinline__ uint32_t next_u32() {
        uint64_t old = state;
        advance();
        uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t) (old >> 59u);
        return __funnelshift_r(xorshifted, xorshifted, rot);  // rotate right
    }
    __device__ __forceinline__ float next_1d() {
        return __uint_as_float((next_u32() >> 9) | 0x3f800000u) - 1.0f;
    }
    __device__ __forceinline__ void seed(uint64_t initstate, uint64_t initseq) {
        state = 0; inc = (initseq << 1) | 1u; advance(); state += initstate; advance();
    }
};
__device__ __forceinline__ void sample_tea_32(uint32_t v0, uint32_t v1, uint32_t *o0, uint32_t *o1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    *o0 = v0; *o1 = v1;
}

__device__ __forceinline__ V3 square_to_cosine_hemisphere(float sx, float sy) {
    float x = fmaf(2.0f, sx, -1.0f), y = fmaf(2.0f, sy, -1.0f);
    bool is_zero = (x == 0.0f) && (y == 0.0f);
    bool q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = 0.25f * TR_PI * rp * (1.0f / r);
    if (q13) phi = 0.5f * TR_PI - phi;
    if (is_zero) phi = 0.0f;
    float s, c; sincos_cephes(phi, &s, &c);
    float px = r * c, py = r * s;
    float z2 = 1.0f - fmaf(py, py, px * px);
    return v3(px, py, sqrtf(fmaxf(z2, 0.0f)));
}

__device__ __forceinline__ float mis_weight(float pdf_a, float pdf_b) {
    float a2 = pdf_a * pdf_a, b2 = pdf_b * pdf_b;
    float w = a2 / (a2 + b2);
    return isfinite(w) ? w : 0.0f;
}

// Duff et al. orthonormal basis (Mitsuba coordinate_system); returns s, t = cross(n, s) is done by the caller.
__device__ __forceinline__ V3 coordinate_system_s(V3 n) {
    float sign = copysignf(1.0f, n.z);
    float a = -(1.0f / (sign + n.z));
    float b = n.x * n.y * a;
    return v3(mulsign(n.x * n.x * a, n.z) + 1.0f, mulsign(b, n.z), mulsign(-n.x, n.z));
}

__device__ __forceinline__ void coordinate_system(V3 n, V3 *s, V3 *t) {
    float sign = copysignf(1.0f, n.z);
    float a = -(1.0f / (sign + n.z));
    float b = n.x * n.y * a;
    *s = v3(mulsign(n.x * n.x * a, n.z) + 1.0f, mulsign(b, n.z), mulsign(-n.x, n.z));
    *t = v3(b, fmaf(n.y, n.y * a, sign), -n.y);
}

// acos for the angulararea falloff: Cephes asinf / acosf polynomial, same operations as the oracle's acos_cephes
__device__ __forceinline__ float asin_cephes(float x) {
    float a = fabsf(x), z, w; bool big = a > 0.5f;
    if (big) { z = 0.5f * (1.0f - a); w = sqrtf(z); } else { w = a; z = a * a; }
    float p = fmaf(4.2163199048e-2f, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f); p = fmaf(p, z, 7.4953002686e-2f); p = fmaf(p, z, 1.6666752422e-1f);
    float r = fmaf(p * z, w, w);
    if (big) r = 1.57079632679489661923f - (r + r);
    return x < 0.0f ? -r : r;
}
__device__ __forceinline__ float acos_cephes(float x) {
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * asin_cephes(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * asin_cephes(sqrtf(0.5f * (1.0f - x)));
    return 1.57079632679489661923f - asin_cephes(x);
}
// AngularAreaLight._fallof_curve (mitransient/emitters/angulararea.py:57-67, 73-81) for a direction in the emitter's frame
// ang = (cutoff_angle, beam_width, cos_cutoff, cos_beam) of the DevEmitter, inv_transition = 1 / (cutoff - beam)
__device__ __forceinline__ float angular_falloff(float4 ang, float inv_transition, V3 d) {
    const float cutoff = ang.x, cos_cutoff = ang.z, cos_beam = ang.w;
    V3 local_dir = normalize(d);
    float cos_theta = local_dir.z;
    float beam_res = cos_theta >= cos_beam ? 1.0f : (cutoff - acos_cephes(cos_theta)) * inv_transition;
    return cos_theta > cos_cutoff ? beam_res : 0.0f;
}
__device__ __forceinline__ V3 frame_to_local(V3 n, V3 v) { V3 s, t; coordinate_system(n, &s, &t); return v3(dot(v, s), dot(v, t), dot(v, n)); }

__device__ __forceinline__ V3 offset_p(V3 p, V3 n, V3 dir) {
    float mag = (1.0f + fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z))) * TR_RAY_EPSILON;
    mag = mulsign(mag, dot(n, dir));
    return fma3(n, mag, p);
}

}  // namespace tr
