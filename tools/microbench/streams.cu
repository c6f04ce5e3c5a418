// How fast can B200 HBM stream the step kernel's access pattern with NO math?
//  (a) plain copy, 2 streams, same total bytes            -> the roofline denominator's pattern
//  (b) the step kernel's 9 arrays (4 read-write state arrays, actions + objectives read-only,
//      obs + reward + done write-only), one warp per 32-env tile, coalesced vector accesses
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void copy_k(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = in[i];
}

// persistent warps, tile = 32 envs, x = 10: per env reads 16+16+120+12 B, writes 16+12+120+5 B
__global__ void pattern_k(float4* goals, const float4* __restrict__ actions, uint32_t* alive, float* total,
                          uint32_t* cnt, const float2* __restrict__ points, float2* obs, float* reward,
                          uint8_t* done, long long tiles) {
    const int lane = threadIdx.x & 31;
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, W = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long t = w; t < tiles; t += W) {
        const long long env = t * 32 + lane;
        float4 g = goals[env], a = actions[env];
        uint32_t al = alive[env]; float tr = total[env]; uint32_t c = cnt[env];
        float2 p[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) p[i] = points[t * 480 + i * 32 + lane];   // 3840 B per tile, coalesced
        float s = g.x + a.y + tr;
#pragma unroll
        for (int i = 0; i < 15; ++i) { p[i].x += s; p[i].y -= s; }
#pragma unroll
        for (int i = 0; i < 15; ++i) obs[t * 480 + i * 32 + lane] = p[i];
        goals[env] = a; alive[env] = al ^ 1u; total[env] = tr + 1.f; cnt[env] = c + 1u;
        reward[env] = s; done[env] = (uint8_t)(c & 1u);
    }
}

int main() {
    const long long N = 1 << 20, tiles = N / 32;
    float4 *goals, *actions; uint32_t *alive, *cnt; float *total, *reward; float2 *points, *obs; uint8_t* done;
    CK(cudaMalloc(&goals, N * 16)); CK(cudaMalloc(&actions, N * 16 * 8)); CK(cudaMalloc(&alive, N * 4));
    CK(cudaMalloc(&cnt, N * 4)); CK(cudaMalloc(&total, N * 4)); CK(cudaMalloc(&reward, N * 4));
    CK(cudaMalloc(&points, N * 120)); CK(cudaMalloc(&obs, N * 120)); CK(cudaMalloc(&done, N));
    CK(cudaMemset(goals, 0, N * 16)); CK(cudaMemset(actions, 0, N * 16 * 8)); CK(cudaMemset(points, 0, N * 120));
    CK(cudaMemset(alive, 0, N * 4)); CK(cudaMemset(cnt, 0, N * 4)); CK(cudaMemset(total, 0, N * 4));
    const double bytes = (double)N * 309;                       // algorithmic bytes of one step
    float4 *cin, *cout; const size_t cn = (size_t)(bytes / 2 / 16);
    CK(cudaMalloc(&cin, cn * 16)); CK(cudaMalloc(&cout, cn * 16)); CK(cudaMemset(cin, 0, cn * 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 200; float ms;
    for (int blocks_per_sm : {4, 8, 16}) {
        for (int i = 0; i < 20; ++i) copy_k<<<148 * blocks_per_sm, 256>>>(cin, cout, cn);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) copy_k<<<148 * blocks_per_sm, 256>>>(cin, cout, cn);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("copy   %2d blk/SM: %.2f us/launch  %.0f GB/s\n", blocks_per_sm, ms * 1e3 / reps, bytes * reps / ms / 1e6);
    }
    for (int blocks_per_sm : {4, 7, 10, 16}) {
        for (int i = 0; i < 20; ++i) pattern_k<<<148 * blocks_per_sm, 128>>>(goals, actions + (i & 7) * N, alive, total, cnt, points, obs, reward, done, tiles);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) pattern_k<<<148 * blocks_per_sm, 128>>>(goals, actions + (i & 7) * N, alive, total, cnt, points, obs, reward, done, tiles);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("9-array step pattern %2d blk/SM x128 thr: %.2f us/launch  %.0f GB/s (309 B/env)\n", blocks_per_sm, ms * 1e3 / reps, bytes * reps / ms / 1e6);
    }
    return 0;
}
