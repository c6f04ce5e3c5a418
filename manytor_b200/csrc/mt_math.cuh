// fp32 math kernels of the step loop: degree-argument sin/cos, degree-valued
// atan2 on the positive quadrant, Philox4x32-10.  Coefficients come from
// tools/fit_polys.py (fp32 Horner error: sin/cos <= 9e-8 abs, atan <= 1.1e-5 deg).
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

// ------------------------------------------------------------------- atan2
// degrees(atan2(y, x)) for y >= 0, x >= 0 (the reference only ever passes
// absolute values, manytor.py:18-21); atan2(0, 0) = 0 like math.atan2.
__device__ __forceinline__ float atan2_deg_pos(float y, float x) {
    float mx = fmaxf(x, y), mn = fminf(x, y);
    float t = __fdividef(mn, mx);
    t = (mx == 0.0f) ? 0.0f : t;
    float s = t * t;
    float p = -0.27388057f;
    p = fmaf(p, s, 1.40695274f);
    p = fmaf(p, s, -3.43219686f);
    p = fmaf(p, s, 5.6967206f);
    p = fmaf(p, s, -8.03824425f);
    p = fmaf(p, s, 11.4427519f);
    p = fmaf(p, s, -19.0978832f);
    p = fmaf(p, s, 57.2957726f);
    float r = p * t;
    return (y > x) ? 90.0f - r : r;
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
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
