#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_scalar(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_packed(float* out, int iters, float a, float b) {
    float2 A = make_float2(a, a), B = make_float2(b, b);
    float2 x0 = make_float2(threadIdx.x, threadIdx.x + 1), x1 = make_float2(threadIdx.x + 2, threadIdx.x + 3),
           x2 = make_float2(threadIdx.x + 4, threadIdx.x + 5), x3 = make_float2(threadIdx.x + 6, threadIdx.x + 7);
    for (int i = 0; i < iters; ++i) {
        x0 = __ffma2_rn(x0, A, B); x1 = __ffma2_rn(x1, A, B); x2 = __ffma2_rn(x2, A, B); x3 = __ffma2_rn(x3, A, B);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
}
// mixed: FFMA2 + ALU ops interleaved, to see whether packed frees issue slots
__global__ void k_mixed_scalar(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    float m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        m0 = fminf(m0, x0); m1 = fminf(m1, x1); m2 = fminf(m2, x2); m3 = fminf(m3, x3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = m0 + m1 + m2 + m3;
}
__global__ void k_mixed_packed(float* out, int iters, float a, float b) {
    float2 A = make_float2(a, a), B = make_float2(b, b);
    float2 x0 = make_float2(threadIdx.x, threadIdx.x + 1), x1 = make_float2(threadIdx.x + 2, threadIdx.x + 3);
    float m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < iters; ++i) {
        x0 = __ffma2_rn(x0, A, B); x1 = __ffma2_rn(x1, A, B);
        m0 = fminf(m0, x0.x); m1 = fminf(m1, x0.y); m2 = fminf(m2, x1.x); m3 = fminf(m3, x1.y);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = m0 + m1 + m2 + m3;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    int iters = 100000; dim3 g(148 * 8), b(256);
    float t1 = timeit([&] { k_scalar<<<g, b>>>(out, iters, 0.999f, 0.001f); });
    float t2 = timeit([&] { k_packed<<<g, b>>>(out, iters, 0.999f, 0.001f); });
    float t3 = timeit([&] { k_mixed_scalar<<<g, b>>>(out, iters, 0.999f, 0.001f); });
    float t4 = timeit([&] { k_mixed_packed<<<g, b>>>(out, iters, 0.999f, 0.001f); });
    double fl = 2.0 * 8 * iters * 148 * 8 * 256;
    printf("scalar FFMA x8 : %.3f ms  %.1f TFLOP/s\n", t1, fl / t1 / 1e9);
    printf("packed FFMA2 x4: %.3f ms  %.1f TFLOP/s\n", t2, fl / t2 / 1e9);
    printf("mixed scalar (4 FFMA + 4 FMNMX): %.3f ms\n", t3);
    printf("mixed packed (2 FFMA2 + 4 FMNMX): %.3f ms\n", t4);
    return 0;
}
