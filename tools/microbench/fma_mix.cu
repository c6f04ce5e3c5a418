// Do packed fp32x2 (FFMA2) and scalar FFMA share one pipe on sm_100, or can a mix exceed either alone?
// (ncu shows the step kernel stalling on math_pipe_throttle with the "fmaheavy" pipe ~55 % busy.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_mix_bin fma_mix.cu && ./fma_mix_bin
// Each kernel keeps 8 independent chains per thread so that latency is covered; P packed + S scalar per iteration.
#include <cuda_runtime.h>
#include <cstdio>

template <int P, int S>
__global__ void k_mix(float *out, int iters, float a, float b) {
    float2 A = make_float2(a, a), B = make_float2(b, b);
    float2 x[8];
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x + i, threadIdx.x - i); y[i] = threadIdx.x * 0.5f + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < P) x[i] = __ffma2_rn(x[i], A, B);
            if (i < S) y[i] = fmaf(y[i], a, b);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i].x + x[i].y + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int P, int S>
static void run(float *out, int iters) {
    dim3 g(148 * 4), b(512);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<P, S><<<g, b>>>(out, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mix<P, S><<<g, b>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lanes = (2.0 * P + S) * iters * 148.0 * 4 * 512;      // fp32 FMA lane-operations
    const double insts = (double)(P + S) * iters * 148.0 * 4 * 16;     // warp instructions
    printf("%d FFMA2 + %d FFMA per iteration: %8.3f ms  %6.1f TFLOP/s  %6.2f G warp-inst/s\n", P, S, ms, 2 * lanes / ms / 1e9,
           insts / ms / 1e6);
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    const int iters = 200000;
    run<0, 8>(out, iters);
    run<8, 0>(out, iters);
    run<4, 4>(out, iters);
    run<4, 8>(out, iters);
    run<8, 4>(out, iters);
    run<8, 8>(out, iters);
    run<2, 8>(out, iters);
    return 0;
}
