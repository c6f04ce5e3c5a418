#include <cuda_runtime.h>
#include <cstdio>
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("l2CacheSize %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB, SMs %d, smemPerSM %zu, smemOptin %zu\n",
           p.l2CacheSize >> 20, p.persistingL2CacheMaxSize >> 20, p.accessPolicyMaxWindowSize >> 20, p.multiProcessorCount,
           p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin);
    size_t lim; cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize); printf("current persisting limit %zu MB\n", lim >> 20);
    return 0;
}
