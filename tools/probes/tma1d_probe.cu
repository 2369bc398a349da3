// Probe: rank-1 uint8 tensor-TMA loads with arbitrary (unaligned / negative) start coordinates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma1d_probe tma1d_probe.cu && ./tma1d_probe <gdim> <l2promo> <coord>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                        CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, uint8_t *out) {
    __shared__ __align__(128) uint8_t buf[256];
    __shared__ uint64_t bar;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" ::"r"(b));
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(d),
                     "l"(&tmap), "r"(x), "r"(b)
                     : "memory");
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W;\n}\n" ::"r"(b) : "memory");
    }
    __syncthreads();
    out[threadIdx.x] = buf[threadIdx.x];
}
int main(int argc, char **argv) {
    unsigned long long gdim = strtoull(argv[1], 0, 0);
    int promo = atoi(argv[2]);
    int x = atoi(argv[3]);
    const int N = 4096;
    uint8_t *d, *o, h[N], ho[256];
    for (int i = 0; i < N; ++i) h[i] = (uint8_t)(i * 7 + 3);
    cudaMalloc(&d, N); cudaMalloc(&o, 256);
    cudaMemcpy(d, h, N, cudaMemcpyHostToDevice);
    void *p = 0; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    alignas(64) CUtensorMap tm;
    cuuint64_t gd[1] = {gdim}, gs[1] = {0}; cuuint32_t box[1] = {256}, es[1] = {1};
    CUresult r = ((PFN)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("gdim=%llu promo=%d x=%d encode=%d ", gdim, promo, x, (int)r);
    if (r) { printf("\n"); return 0; }
    k<<<1, 256>>>(tm, x, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        cudaMemcpy(ho, o, 256, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 256; ++i) {
            long g = (long)x + i;
            uint8_t want = (g >= 0 && g < (long)gdim && g < N) ? h[g] : 0;
            bad += ho[i] != want;
        }
        printf("mismatch=%d", bad);
    }
    printf("\n");
    return 0;
}
