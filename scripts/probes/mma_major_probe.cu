// Micro-probe: tcgen05.mma issue rate for K-major vs MN-major bf16 operands (M=128, N=128/256, K=16), no loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ipr_gan_b200/csrc -I include -o /tmp/mma_probe scripts/probes/mma_major_probe.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace tc;

template <int N>
__global__ void __launch_bounds__(128) probe(int a_mn, int b_mn, int iters, long long *out)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 96 * 1024, slot = bar + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 1) tmem_alloc(slot, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N, a_mn, b_mn);
        const uint32_t sA = base, sB = base + 32 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t da = a_mn ? umma_desc_sw128(sA + k * 2048, 8192, 1024) : umma_desc_sw128(sA, 0, 1024) + 2 * k;
                const uint64_t db = b_mn ? umma_desc_sw128(sB + k * 2048, 8192, 1024) : umma_desc_sw128(sB, 0, 1024) + 2 * k;
                umma_bf16(tmem, da, db, idesc, 1u);
            }
        }
        umma_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

int main()
{
    long long *d; cudaMalloc(&d, 8);
    const int iters = 2000;
    cudaFuncSetAttribute(probe<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(probe<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int n = 128; n <= 256; n += 128)
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) {
                for (int rep = 0; rep < 2; rep++) {
                    if (n == 128) probe<128><<<1, 128, 100 * 1024>>>(a, b, iters, d);
                    else probe<256><<<1, 128, 100 * 1024>>>(a, b, iters, d);
                    cudaDeviceSynchronize();
                }
                long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("N=%d A=%s B=%s : %.1f cycles per MMA (ideal %d)  err=%s\n", n, a ? "MN" : "K ", b ? "MN" : "K ",
                       (double)h / (iters * 4), n / 2, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
