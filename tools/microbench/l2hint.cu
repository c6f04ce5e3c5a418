// Does pinning part of the step loop's working set in B200's 126 MB L2 pay?
// Same 9-array no-math pattern as streams.cu, with per-access L2 eviction policies:
// state (goals/alive/total/counters, read+written every step) evict_last, a fraction of the
// objective tiles evict_last, everything streamed once (actions, obs, reward, done) evict_first.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ float4 ld4(const float4* a, uint64_t p) { float4 v; asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(p)); return v; }
__device__ __forceinline__ void st4(float4* a, float4 v, uint64_t p) { asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" :: "l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(p) : "memory"); }
__device__ __forceinline__ float2 ld2(const float2* a, uint64_t p) { float2 v; asm volatile("ld.global.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(p)); return v; }
__device__ __forceinline__ void st2(float2* a, float2 v, uint64_t p) { asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" :: "l"(a), "f"(v.x), "f"(v.y), "l"(p) : "memory"); }
__device__ __forceinline__ uint32_t ld1(const uint32_t* a, uint64_t p) { uint32_t v; asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(p)); return v; }
__device__ __forceinline__ void st1(uint32_t* a, uint32_t v, uint64_t p) { asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(a), "r"(v), "l"(p) : "memory"); }

__global__ void pattern_k(float4* goals, const float4* __restrict__ actions, uint32_t* alive, uint32_t* total,
                          uint32_t* cnt, const float2* __restrict__ points, float2* obs, uint32_t* reward,
                          uint8_t* done, long long tiles, int keep_pct, int hint_state) {
    const int lane = threadIdx.x & 31;
    const uint64_t PL = pol_last(), PF = pol_first();
    const uint64_t PS = hint_state ? PL : PF;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, W = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = w; t < tiles; t += W) {
        const long long env = t * 32 + lane;
        const uint64_t PP = ((t * 37) % 100) < keep_pct ? PL : PF;      // a fixed subset of tiles stays in L2
        float4 g = ld4(goals + env, PS), a = ld4(actions + env, PF);
        uint32_t al = ld1(alive + env, PS), tr = ld1(total + env, PS), c = ld1(cnt + env, PS);
        float2 p[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) p[i] = ld2(points + t * 480 + i * 32 + lane, PP);
        float s = g.x + a.y + __uint_as_float(tr);
#pragma unroll
        for (int i = 0; i < 15; ++i) { p[i].x += s; p[i].y -= s; }
#pragma unroll
        for (int i = 0; i < 15; ++i) st2(obs + t * 480 + i * 32 + lane, p[i], PF);
        st4(goals + env, a, PS); st1(alive + env, al ^ 1u, PS); st1(total + env, tr + 1u, PS); st1(cnt + env, c + 1u, PS);
        st1(reward + env, __float_as_uint(s), PF); done[env] = (uint8_t)(c & 1u);
    }
}

int main(int argc, char** argv) {
    const long long N = 1 << 20, tiles = N / 32;
    float4 *goals, *actions; uint32_t *alive, *cnt, *total, *reward; float2 *points, *obs; uint8_t* done;
    CK(cudaMalloc(&goals, N * 16)); CK(cudaMalloc(&actions, N * 16 * 8)); CK(cudaMalloc(&alive, N * 4));
    CK(cudaMalloc(&cnt, N * 4)); CK(cudaMalloc(&total, N * 4)); CK(cudaMalloc(&reward, N * 4));
    CK(cudaMalloc(&points, N * 120)); CK(cudaMalloc(&obs, N * 120)); CK(cudaMalloc(&done, N));
    CK(cudaMemset(goals, 0, N * 16)); CK(cudaMemset(actions, 0, N * 16 * 8)); CK(cudaMemset(points, 0, N * 120));
    CK(cudaMemset(alive, 0, N * 4)); CK(cudaMemset(cnt, 0, N * 4)); CK(cudaMemset(total, 0, N * 4));
    const double bytes = (double)N * 309;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 300; float ms;
    for (int setaside_mb : {0, 79}) {
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)setaside_mb << 20));
        for (int hint_state : {0, 1})
            for (int keep : {0, 20, 35, 45, 55, 70}) {
                if (!hint_state && keep) continue;
                for (int i = 0; i < 30; ++i) pattern_k<<<148 * 7, 128>>>(goals, actions + (i & 7) * N, alive, total, cnt, points, obs, reward, done, tiles, keep, hint_state);
                cudaEventRecord(e0);
                for (int i = 0; i < reps; ++i) pattern_k<<<148 * 7, 128>>>(goals, actions + (i & 7) * N, alive, total, cnt, points, obs, reward, done, tiles, keep, hint_state);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf("set-aside %2d MB  state %s  objectives kept %2d%%: %.2f us/launch  %.0f GB/s-equivalent (309 B/env)\n",
                       setaside_mb, hint_state ? "evict_last " : "evict_first", keep, ms * 1e3 / reps, bytes * reps / ms / 1e6);
            }
    }
    CK(cudaCtxResetPersistingL2Cache());
    return 0;
}
