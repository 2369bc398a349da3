// Does an L2 prefetch to an unmapped global address fault?  (It must not, if speculative addresses are to be used as hints.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(const char *good, uint64_t bad, int mode, int *out) {
    if (mode == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(good));
    if (mode == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(bad));
    if (mode == 2) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(good), "r"(4096));
    if (mode == 3) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(bad), "r"(4096));
    if (mode == 4) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(bad + 16));
    out[0] = mode;
}
int main() {
    char *good; int *out;
    cudaMalloc(&good, 1 << 20); cudaMalloc(&out, 4);
    const uint64_t bads[] = {0x00007f0000000000ull, 0xdead00000000ull, 0x10ull, reinterpret_cast<uint64_t>(good) + (1ull << 40)};
    for (int mode = 0; mode < 5; ++mode)
        for (uint64_t bad : bads) {
            probe<<<1, 32>>>(good, bad, mode, out);
            cudaError_t e = cudaDeviceSynchronize();
            printf("mode %d bad=%llx -> %s\n", mode, (unsigned long long)bad, cudaGetErrorString(e));
            if (e != cudaSuccess) { printf("STICKY ERROR, stop\n"); return 1; }
        }
    return 0;
}
