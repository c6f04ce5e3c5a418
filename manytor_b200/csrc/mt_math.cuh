// fp32 math kernels of the step loop: degree-argument sin/cos, degree-valued
// atan2 on the positive quadrant, Philox4x32-10.  Coefficients come from
// tools/fit_polys.py (fp32 Horner error: sin/cos <= 1.6e-7 abs, atan2 <= 7.3e-5 deg).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mt {

// ---------------------------------------------------------------- sin / cos
// sin and cos of an angle given in DEGREES.  The reference converts with
// math.radians and calls np.sin/np.cos in fp64 (manytor.py:28-30,39); here the
// reduction is done in half-turns, which is exact for any fp32 input:
//   t = x/180, n = rint(2t), r = t - n/2 in [-1/4, 1/4], quadrant = n mod 4.
__device__ __forceinline__ void sincos_deg(float x, float &s, float &c) {
    const float inv_hi = 0x1.6c16c2p-8f;                 // fl(1/180)
    const float inv_lo = -0x1.27d27ep-33f;               // 1/180 - inv_hi
    float t = fmaf(x, inv_lo, x * inv_hi);
    float n = rintf(t + t);
    float r = fmaf(n, -0.5f, t);
    int q = __float2int_rn(n);
    float u = r * r;
    float sp = fmaf(fmaf(fmaf(-0.58907866f, u, 2.5497673f), u, -5.1677079f), u, 3.14159274f) * r;
    float cp = fmaf(fmaf(fmaf(fmaf(0.23132971f, u, -1.33504462f), u, 4.05870724f), u, -4.93480206f), u, 1.0f);
    float ss = (q & 1) ? cp : sp;
    float cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------- atan2
// degrees(atan2(y, x)) for y >= 0, x >= 0 (the reference only ever passes
// absolute values, manytor.py:18-21), given hyp = sqrt(x^2 + y^2), which the
// caller already has.  Half-angle form: atan2(y, x) = 2 atan(y / (x + hyp));
// the argument lies in [0, 1] over the whole quadrant, so there is no
// min/max/select; x + hyp adds two non-negative numbers (no cancellation).
// atan2(0, 0) = 0 like math.atan2.
__device__ __forceinline__ float atan2_deg_pos(float y, float x, float hyp) {
    float t = y * fast_rcp(fmaxf(x + hyp, 1e-30f));
    float s = t * t;
    float p = 0.917460918f;              // 2 atan(t) in degrees, degree 6 in s (tools/fit_polys.py: 7.3e-5 deg max error)
    p = fmaf(p, s, -4.29046488f);
    p = fmaf(p, s, 9.66606236f);
    p = fmaf(p, s, -15.4837093f);
    p = fmaf(p, s, 22.7891674f);
    p = fmaf(p, s, -38.1899414f);
    p = fmaf(p, s, 114.591492f);
    return p * t;
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------ packed fp32x2
// sm_100 issues two fp32 FMAs per lane from one instruction (SASS FFMA2 / FADD2 /
// FMUL2 on an aligned register pair, with free -x / |x| and scalar-broadcast
// operands).  The kernel is bound by instruction issue, not by the FMA pipe, so
// independent pairs of evaluations (two objectives, two joints, sin+cos of one
// angle, two cosine sequences) are packed: same flops, half the issue slots.
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// polynomial cores shared by the scalar and packed sin/cos: sin(pi r)/r and cos(pi r) in u = r^2, |r| <= 1/4
__device__ __forceinline__ float2 sinpi_poly2(float2 r, float2 u) {
    return mul2(fma2(fma2(fma2(bc2(-0.58907866f), u, bc2(2.5497673f)), u, bc2(-5.1677079f)), u, bc2(3.14159274f)), r);
}
__device__ __forceinline__ float2 cospi_poly2(float2 u) {
    return fma2(fma2(fma2(fma2(bc2(0.23132971f), u, bc2(-1.33504462f)), u, bc2(4.05870724f)), u, bc2(-4.93480206f)), u, bc2(1.0f));
}
// ... and on |r| <= 1/2 (one more term each; tools/fit_polys.py: fp32 Horner error 1.6e-7 / 1.1e-7 abs)
__device__ __forceinline__ float2 sinpi_poly2_half(float2 r, float2 u) {
    float2 p = fma2(bc2(0.0776594058f), u, bc2(-0.598292172f));
    p = fma2(p, u, bc2(2.55007768f));
    p = fma2(p, u, bc2(-5.1677103f));
    p = fma2(p, u, bc2(3.14159274f));
    return mul2(p, r);
}
__device__ __forceinline__ float2 cospi_poly2_half(float2 u) {
    float2 p = fma2(bc2(-0.0243967157f), u, bc2(0.234937564f));
    p = fma2(p, u, bc2(-1.33521211f));
    p = fma2(p, u, bc2(4.05870914f));
    p = fma2(p, u, bc2(-4.93480206f));
    return fma2(p, u, bc2(1.0f));
}

// sin/cos of TWO angles in degrees at once.  Reduction in half-turns, exact for any fp32 input below
// 3e8 degrees: t = x/180, n = rint(t) (the 1.5*2^23 magic constant makes the rounding a packed add and
// leaves n's parity in the low mantissa bit), r = t - n in [-1/2, 1/2]; sin(pi t) = (-1)^n sin(pi r),
// cos likewise -- ONE sign flip shared by both results.  (Round 1 reduced to quarter turns: polynomials
// one term shorter, but a swap + two sign fix-ups per angle, 14 instructions against 3 here.)
__device__ __forceinline__ void sincos_deg2(float2 x, float2 &s, float2 &c) {
    const float inv_hi = 0x1.6c16c2p-8f, inv_lo = -0x1.27d27ep-33f, magic = 12582912.0f;
    float2 t = fma2(x, bc2(inv_lo), mul2(x, bc2(inv_hi)));
    float2 m = add2(t, bc2(magic));                 // low mantissa bits of m = rint(t)
    float2 n = add2(m, bc2(-magic));
    float2 r = add2(t, neg2(n));
    float2 u = mul2(r, r);
    float2 sp = sinpi_poly2_half(r, u), cp = cospi_poly2_half(u);
    const int fx = __float_as_int(m.x) << 31, fy = __float_as_int(m.y) << 31;
    s = make_float2(__int_as_float(__float_as_int(sp.x) ^ fx), __int_as_float(__float_as_int(sp.y) ^ fy));
    c = make_float2(__int_as_float(__float_as_int(cp.x) ^ fx), __int_as_float(__float_as_int(cp.y) ^ fy));
}

// two SMALL angles (|x| <= 45 degrees): quadrant 0, no fix-up
__device__ __forceinline__ void sincos_deg_small2(float2 x, float2 &s, float2 &c) {
    const float inv_hi = 0x1.6c16c2p-8f, inv_lo = -0x1.27d27ep-33f;
    float2 r = fma2(x, bc2(inv_lo), mul2(x, bc2(inv_hi)));
    float2 u = mul2(r, r);
    s = sinpi_poly2(r, u);
    c = cospi_poly2(u);
}

// one SMALL angle (|x| <= 45 degrees), scalar
__device__ __forceinline__ void sincos_deg_small(float x, float &s, float &c) {
    const float inv_hi = 0x1.6c16c2p-8f, inv_lo = -0x1.27d27ep-33f;
    const float r = fmaf(x, inv_lo, x * inv_hi), u = r * r;
    s = fmaf(fmaf(fmaf(-0.58907866f, u, 2.5497673f), u, -5.1677079f), u, 3.14159274f) * r;
    c = fmaf(fmaf(fmaf(fmaf(0.23132971f, u, -1.33504462f), u, 4.05870724f), u, -4.93480206f), u, 1.0f);
}

// 2 atan(t) in degrees for two arguments in [0, 1] (the half-angle form of atan2_deg_pos)
__device__ __forceinline__ float2 atan_half_deg2(float2 t) {
    float2 s = mul2(t, t);
    float2 p = bc2(0.917460918f);
    p = fma2(p, s, bc2(-4.29046488f));
    p = fma2(p, s, bc2(9.66606236f));
    p = fma2(p, s, bc2(-15.4837093f));
    p = fma2(p, s, bc2(22.7891674f));
    p = fma2(p, s, bc2(-38.1899414f));
    p = fma2(p, s, bc2(114.591492f));
    return mul2(p, t);
}

// ------------------------------------------------------------------ Philox
// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator: no
// per-env RNG state is kept in HBM; streams are keyed by (seed, env id, index).
struct Philox {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox{c0, c1, c2, c3};
}

enum : uint32_t { STREAM_ACTIONS = 0x41435431u /* 'ACT1' */, STREAM_POINTS = 0x50545331u /* 'PTS1' */ };

// integer uniform on [low, low+span) from one 32-bit word (multiply-shift)
__host__ __device__ __forceinline__ int32_t uniform_int(uint32_t u, int32_t low, uint32_t span) {
    return low + (int32_t)mulhi32(u, span);
}

// float uniform on [0, 1) with 24 bits
__host__ __device__ __forceinline__ float uniform01(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

}  // namespace mt
