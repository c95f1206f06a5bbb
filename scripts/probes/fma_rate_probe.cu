// Issue-rate probe: 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) vs mixed, per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate_probe fma_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float *out, long long *cyc, int iters)
{
    float a[16]; float2 b[16];
    const float s = 1.0001f + threadIdx.x * 1e-7f, t = 0.9999f;
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = i + threadIdx.x; b[i] = make_float2(i, threadIdx.x); }
    unsigned long long s2, t2;
    asm("mov.b64 %0, {%1, %2};" : "=l"(s2) : "f"(s), "f"(s));
    asm("mov.b64 %0, {%1, %2};" : "=l"(t2) : "f"(t), "f"(t));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0 || MODE == 2) a[i] = fmaf(a[i], s, t);
            if (MODE == 1 || MODE == 2) {
                unsigned long long v;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(b[i].x), "f"(b[i].y));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(s2), "l"(t2));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(b[i].x), "=f"(b[i].y) : "l"(v));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) acc += a[i] + b[i].x + b[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps_per_sm)
{
    float *out; long long *cyc;
    const int iters = 4096, threads = warps_per_sm * 32, blocks = 148;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<MODE><<<blocks, threads>>>(out, cyc, iters);
    probe<MODE><<<blocks, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_warp = (MODE == 2 ? 32.0 : 16.0) * iters;           // instructions per warp
    const double instr_sm = per_warp * warps_per_sm;
    printf("%-12s warps/SM=%2d  cycles=%lld  warp-instr/clk/SM=%.3f  per SMSP=%.3f\n", name, warps_per_sm, h[0],
           instr_sm / h[0], instr_sm / h[0] / 4);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 16, 32}) { run<0>("FFMA", w); run<1>("FFMA2", w); run<2>("FFMA+FFMA2", w); }
    return 0;
}
