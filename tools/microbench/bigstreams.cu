// What bounds the step kernel's access pattern once the per-env state no longer fits in L2
// (N >= 2^22 envs)?  No math; persistent warps, one 32-env tile at a time, like the step kernel.
//   copy      plain 2-stream copy of the same bytes (the roofline denominator's pattern)
//   soa       the kernel's arrays as they are: goals, alive, total, counters (read + write),
//             actions + points (read), obs + reward + done (write)
//   rec32     the four state arrays fused into one 32-byte record per env (2 x 128-bit per lane)
//   rec24     goals + one 64-bit word (alive | ep_len | total) = 24 B per env in two arrays
// each with L2 policies: 0 = none, 1 = state evict_last + streams evict_first, 2 = streams evict_first only,
// 3 = as 1 but only the first `keep` envs' state is evict_last (the rest evict_normal),
// 4 = as 1 plus the first `keep` envs' OBJECTIVES evict_last as well (they are re-read every step too)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bigstreams_bin bigstreams.cu && ./bigstreams_bin 22
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_normal() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }

template <class T> __device__ __forceinline__ T ldh(const T* a, uint64_t pol);
template <> __device__ __forceinline__ float4 ldh(const float4* a, uint64_t pol) {
    float4 v; asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol)); return v; }
template <> __device__ __forceinline__ float2 ldh(const float2* a, uint64_t pol) {
    float2 v; asm volatile("ld.global.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(pol)); return v; }
template <> __device__ __forceinline__ float ldh(const float* a, uint64_t pol) {
    float v; asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol)); return v; }
template <> __device__ __forceinline__ uint32_t ldh(const uint32_t* a, uint64_t pol) {
    uint32_t v; asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol)); return v; }
__device__ __forceinline__ void sth(float4* a, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory"); }
__device__ __forceinline__ void sth(float2* a, float2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(a), "f"(v.x), "f"(v.y), "l"(pol) : "memory"); }
__device__ __forceinline__ void sth(float* a, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory"); }
__device__ __forceinline__ void sth(uint32_t* a, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory"); }
__device__ __forceinline__ void sth(uint8_t* a, uint8_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(a), "r"((uint32_t)v), "l"(pol) : "memory"); }

__global__ void copy_k(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = in[i];
}

struct Arr {
    float4* goals; const float4* actions; uint32_t* alive; float* total; uint32_t* cnt;
    const float2* points; float2* obs; float* reward; uint8_t* done;
    float4* rec;      // rec32: [N][2] float4
    float2* word;     // rec24: [N] 64-bit
    long long tiles, keep_tiles;
};

// LAYOUT 0 = soa, 1 = rec32, 2 = rec24
template <int LAYOUT, int POL>
__global__ void __launch_bounds__(128) pattern_k(Arr A) {
    const int lane = threadIdx.x & 31;
    const uint64_t pn = pol_normal();
    const uint64_t ps = POL == 0 ? pn : pol_first();
    const uint64_t pl = (POL == 1 || POL == 3 || POL == 4) ? pol_last() : pn;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, W = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = w; t < A.tiles; t += W) {
        const long long env = t * 32 + lane;
        const uint64_t pk = (POL == 3 && t >= A.keep_tiles) ? pn : pl;
        float4 g, a = ldh(A.actions + env, ps);
        uint32_t al, c; float tr;
        if (LAYOUT == 0) {
            g = ldh(A.goals + env, pk); al = ldh(A.alive + env, pk); tr = ldh(A.total + env, pk); c = ldh(A.cnt + env, pk);
        } else if (LAYOUT == 1) {
            g = ldh(A.rec + env * 2, pk);
            float4 r = ldh(A.rec + env * 2 + 1, pk);
            al = __float_as_uint(r.x); tr = r.y; c = __float_as_uint(r.z);
        } else {
            g = ldh(A.goals + env, pk);
            float2 r = ldh(A.word + env, pk);
            al = __float_as_uint(r.x); tr = r.y; c = al >> 16;
        }
        float2 p[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) p[i] = ldh(A.points + t * 480 + i * 32 + lane, (POL == 4 && t < A.keep_tiles) ? pl : ps);
        float s = g.x + a.y + tr;
#pragma unroll
        for (int i = 0; i < 15; ++i) { p[i].x += s; p[i].y -= s; }
#pragma unroll
        for (int i = 0; i < 15; ++i) sth(A.obs + t * 480 + i * 32 + lane, p[i], ps);
        if (LAYOUT == 0) {
            sth(A.goals + env, a, pk); sth(A.alive + env, al ^ 1u, pk); sth(A.total + env, tr + 1.f, pk); sth(A.cnt + env, c + 1u, pk);
        } else if (LAYOUT == 1) {
            sth(A.rec + env * 2, a, pk);
            sth(A.rec + env * 2 + 1, make_float4(__uint_as_float(al ^ 1u), tr + 1.f, __uint_as_float(c + 1u), 0.f), pk);
        } else {
            sth(A.goals + env, a, pk);
            sth(A.word + env, make_float2(__uint_as_float(al ^ 1u), tr + 1.f), pk);
        }
        sth(A.reward + env, s, ps); sth(A.done + env, (uint8_t)(c & 1u), ps);
    }
}

template <int LAYOUT, int POL>
int run(const char* name, Arr A, long long N, int bps, int reps, double bytes_per_env) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const float4* act0 = A.actions;
    for (int i = 0; i < 5; ++i) { A.actions = act0 + (i & 3) * N; pattern_k<LAYOUT, POL><<<148 * bps, 128>>>(A); }
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) { A.actions = act0 + (i & 3) * N; pattern_k<LAYOUT, POL><<<148 * bps, 128>>>(A); }
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaGetLastError());
    printf("%-6s pol %d  %2d blk/SM: %8.2f us/launch  %6.0f GB/s real (%3.0f B/env)  %6.0f GB/s at 309 B/env\n", name, POL, bps,
           ms * 1e3 / reps, bytes_per_env * N * reps / ms / 1e6, bytes_per_env, 309.0 * N * reps / ms / 1e6);
    return 0;
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 22;
    const long long N = 1ll << lg, tiles = N / 32;
    const int reps = lg >= 22 ? 60 : 200;
    Arr A{};
    float4 *goals, *actions, *rec; uint32_t *alive, *cnt; float *total, *reward; float2 *points, *obs, *word; uint8_t* done;
    CK(cudaMalloc(&goals, N * 16)); CK(cudaMalloc(&actions, N * 16 * 4)); CK(cudaMalloc(&alive, N * 4));
    CK(cudaMalloc(&cnt, N * 4)); CK(cudaMalloc(&total, N * 4)); CK(cudaMalloc(&reward, N * 4));
    CK(cudaMalloc(&points, N * 120)); CK(cudaMalloc(&obs, N * 120)); CK(cudaMalloc(&done, N));
    CK(cudaMalloc(&rec, N * 32)); CK(cudaMalloc(&word, N * 8));
    CK(cudaMemset(goals, 0, N * 16)); CK(cudaMemset(actions, 0, N * 16 * 4)); CK(cudaMemset(points, 0, N * 120));
    CK(cudaMemset(alive, 0, N * 4)); CK(cudaMemset(cnt, 0, N * 4)); CK(cudaMemset(total, 0, N * 4));
    CK(cudaMemset(rec, 0, N * 32)); CK(cudaMemset(word, 0, N * 8));
    A.goals = goals; A.actions = actions; A.alive = alive; A.total = total; A.cnt = cnt; A.points = points; A.obs = obs;
    A.reward = reward; A.done = done; A.rec = rec; A.word = word; A.tiles = tiles;
    A.keep_tiles = (1ll << 20) / 32;    // policy 3: keep the first 2^20 envs' state (28-32 MB) in L2

    const double bytes = (double)N * 309;
    float4 *cin, *cout; const size_t cn = (size_t)(bytes / 2 / 16);
    CK(cudaMalloc(&cin, cn * 16)); CK(cudaMalloc(&cout, cn * 16)); CK(cudaMemset(cin, 0, cn * 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    printf("N = 2^%d envs\n", lg);
    for (int bps : {8, 16}) {
        for (int i = 0; i < 5; ++i) copy_k<<<148 * bps, 256>>>(cin, cout, cn);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) copy_k<<<148 * bps, 256>>>(cin, cout, cn);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("copy   %2d blk/SM: %8.2f us/launch  %6.0f GB/s\n", bps, ms * 1e3 / reps, bytes * reps / ms / 1e6);
    }
    const double kf = argc > 2 ? atof(argv[2]) : 1.0;
    if (argc > 2) {   // sweep mode: ./bigstreams_bin lg keep_fraction_of_2^20_envs
        A.keep_tiles = (long long)(kf * (1 << 20)) / 32;
        printf("keep = %.2f x 2^20 envs\n", kf);
        for (int bps : {7, 12}) {
            run<0, 1>("soa", A, N, bps, reps, 317);
            run<0, 3>("soa", A, N, bps, reps, 317);
            run<0, 4>("soa", A, N, bps, reps, 317);
            run<2, 1>("rec24", A, N, bps, reps, 309);
            run<2, 3>("rec24", A, N, bps, reps, 309);
            run<2, 4>("rec24", A, N, bps, reps, 309);
        }
        return 0;
    }
    for (int bps : {7, 12}) {
        run<0, 0>("soa", A, N, bps, reps, 317);
        run<0, 1>("soa", A, N, bps, reps, 317);
        run<0, 2>("soa", A, N, bps, reps, 317);
        run<0, 3>("soa", A, N, bps, reps, 317);
        run<1, 0>("rec32", A, N, bps, reps, 325);
        run<1, 1>("rec32", A, N, bps, reps, 325);
        run<1, 2>("rec32", A, N, bps, reps, 325);
        run<1, 3>("rec32", A, N, bps, reps, 325);
        run<2, 0>("rec24", A, N, bps, reps, 309);
        run<2, 1>("rec24", A, N, bps, reps, 309);
        run<2, 2>("rec24", A, N, bps, reps, 309);
        run<2, 3>("rec24", A, N, bps, reps, 309);
    }
    return 0;
}
