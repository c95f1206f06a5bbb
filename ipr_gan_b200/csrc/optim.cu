// optim.cu -- flat-arena optimizer and weight re-packing.
//   * ipr_adam_flat_f32: one launch updates EVERY parameter of a network (params / grads / moments live in four
//     contiguous fp32 arenas), replacing torch.optim.Adam's per-tensor foreach kernels (models/dcgan.py:21-24).
//     The step counter lives on the device so the launch is CUDA-graph replayable; the data-parallel 1/world factor
//     is folded in as grad_scale, and the gradient arena can be cleared as it is consumed (fused zero_grad).
//   * ipr_gather_pack_bf16: one launch rebuilds all bf16 GEMM operand layouts of a network from the fp32 master
//     arena through a precomputed index table (dst[i] = src[idx[i]], idx < 0 -> 0).
#include "ipr_common.cuh"
#include <cuda_bf16.h>

namespace {

// 128-bit load that does not allocate in L1 (NOT the read-only .nc path: the same thread rewrites the location below)
__device__ __forceinline__ float4 ld_once4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 long long n, float lr, float b1, float b2, float eps, float wd, float gscale, int zero_grad,
                 float *__restrict__ step, unsigned int *__restrict__ ticket)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    // every CTA reads the step counter, then takes a ticket; the CTA holding the last ticket knows all others have
    // read the old value and publishes step + 1 -- no separate increment launch, still CUDA-graph replayable
    __shared__ float step_sm;
    if (threadIdx.x == 0) {
        step_sm = *reinterpret_cast<volatile float *>(step);
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0u; *step = step_sm + 1.0f; }
    }
    __syncthreads();
    const float t = step_sm + 1.0f;
    const float bc1 = 1.0f - powf(b1, t);
    const float inv_bc2_sqrt = 1.0f / sqrtf(1.0f - powf(b2, t));
    const float step_size = lr / bc1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = n >> 2;
    // Two float4 per array and thread in flight: all eight 128-bit loads are issued before the first use (the first
    // version loaded, computed and stored one float4 per array at a time and reached 1.3-1.6 TB/s from HBM --
    // scripts/adam_probe.py -- a quarter of what a copy does; the arenas are cold when the optimizer runs).  The
    // moments are touched once per step: streaming loads / stores keep them from displacing the weights in L2.
    auto update = [&](float4 &pp, const float4 &gg, float4 &mm, float4 &vv) {
        float *pa = &pp.x; const float *ga = &gg.x; float *ma = &mm.x; float *va = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float gk = ga[k] * gscale + wd * pa[k];
            ma[k] = b1 * ma[k] + (1.0f - b1) * gk;
            va[k] = b2 * va[k] + (1.0f - b2) * gk * gk;
            // one approximate divide per element (<= 2 ulp) instead of two IEEE ones
            pa[k] -= __fdividef(step_size * ma[k], fmaf(sqrtf(va[k]), inv_bc2_sqrt, eps));
        }
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
        const long long j = i + stride;
        float4 p0 = reinterpret_cast<float4 *>(p)[i], p1 = reinterpret_cast<float4 *>(p)[j];
        const float4 g0 = reinterpret_cast<const float4 *>(g)[i], g1 = reinterpret_cast<const float4 *>(g)[j];
        float4 m0 = ld_once4(reinterpret_cast<const float4 *>(m) + i), m1 = ld_once4(reinterpret_cast<const float4 *>(m) + j);
        float4 v0 = ld_once4(reinterpret_cast<const float4 *>(v) + i), v1 = ld_once4(reinterpret_cast<const float4 *>(v) + j);
        update(p0, g0, m0, v0);
        update(p1, g1, m1, v1);
        if (zero_grad) { reinterpret_cast<float4 *>(g)[i] = zero4; reinterpret_cast<float4 *>(g)[j] = zero4; }
        reinterpret_cast<float4 *>(p)[i] = p0; reinterpret_cast<float4 *>(p)[j] = p1;
        ipr_stg_stream4(reinterpret_cast<float4 *>(m) + i, m0); ipr_stg_stream4(reinterpret_cast<float4 *>(m) + j, m1);
        ipr_stg_stream4(reinterpret_cast<float4 *>(v) + i, v0); ipr_stg_stream4(reinterpret_cast<float4 *>(v) + j, v1);
    }
    for (; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4 *>(p)[i];
        const float4 gg = reinterpret_cast<const float4 *>(g)[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        update(pp, gg, mm, vv);
        if (zero_grad) reinterpret_cast<float4 *>(g)[i] = zero4;
        reinterpret_cast<float4 *>(p)[i] = pp;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gk = g[i] * gscale + wd * p[i];
        if (zero_grad) g[i] = 0.0f;
        m[i] = b1 * m[i] + (1.0f - b1) * gk;
        v[i] = b2 * v[i] + (1.0f - b2) * gk * gk;
        p[i] -= __fdividef(step_size * m[i], fmaf(sqrtf(v[i]), inv_bc2_sqrt, eps));
    }
}

__global__ void __launch_bounds__(256)
gather_pack_kernel(const float *__restrict__ src, const int *__restrict__ idx, __nv_bfloat16 *__restrict__ dst, long long n,
                   const int *__restrict__ idx32, float *__restrict__ dst32, long long n32)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n8 = n >> 3;
    // Two dependent load levels (index table, then the gathered weights): two 8-element groups per thread, all four
    // index loads first, then all sixteen gathers, keep twice the bytes in flight (the launch ran at 1.3 TB/s).
    auto pack = [&](const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            w[k] = *reinterpret_cast<uint32_t *>(&t);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n8; i += 2 * stride) {
        const long long j = i + stride;
        const int4 a0 = __ldg(reinterpret_cast<const int4 *>(idx) + 2 * i), a1 = __ldg(reinterpret_cast<const int4 *>(idx) + 2 * i + 1);
        const int4 b0 = __ldg(reinterpret_cast<const int4 *>(idx) + 2 * j), b1 = __ldg(reinterpret_cast<const int4 *>(idx) + 2 * j + 1);
        const int ia[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const int ib[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float fa[8], fb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { fa[k] = ia[k] >= 0 ? __ldg(src + ia[k]) : 0.0f; fb[k] = ib[k] >= 0 ? __ldg(src + ib[k]) : 0.0f; }
        reinterpret_cast<uint4 *>(dst)[i] = pack(fa);
        reinterpret_cast<uint4 *>(dst)[j] = pack(fb);
    }
    for (; i < n8; i += stride) {
        const int4 i0 = __ldg(reinterpret_cast<const int4 *>(idx) + 2 * i);
        const int4 i1 = __ldg(reinterpret_cast<const int4 *>(idx) + 2 * i + 1);
        const int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = id[k] >= 0 ? __ldg(src + id[k]) : 0.0f;
        reinterpret_cast<uint4 *>(dst)[i] = pack(f);
    }
    // fp32 side table (permuted copies of the few parameters the kernels read in fp32: biases, the final GEMV row)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
        const int id = __ldg(idx32 + i);
        dst32[i] = id >= 0 ? __ldg(src + id) : 0.0f;
    }
}

}  // namespace

extern "C" int ipr_adam_flat_f32(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                                 float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                 int zero_grad, float *step, uint32_t *ticket, ipr_stream_t stream)
{
    IPR_REQUIRE(param && grad && exp_avg && exp_avg_sq && step && ticket, IPR_E_NULL);
    IPR_REQUIRE(n > 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(param) && ipr_aligned16(grad) && ipr_aligned16(exp_avg) && ipr_aligned16(exp_avg_sq),
                IPR_E_ALIGN);
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)ipr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    IPR_LAUNCH_PDL((adam_flat_kernel), (unsigned)blocks, 256, 0, ipr_cu(stream), param, grad, exp_avg, exp_avg_sq, (long long)n, lr,
                   beta1, beta2, eps, weight_decay, grad_scale, zero_grad, step, (unsigned int *)ticket);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_gather_pack_bf16(const float *src, const int32_t *index, void *dst, int64_t n,
                                    const int32_t *index_f32, float *dst_f32, int64_t n_f32, ipr_stream_t stream)
{
    IPR_REQUIRE(src && index && dst, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && n % 8 == 0 && n_f32 >= 0, IPR_E_SHAPE);
    IPR_REQUIRE(n_f32 == 0 || (index_f32 && dst_f32), IPR_E_NULL);
    IPR_REQUIRE(ipr_aligned16(index) && ipr_aligned16(dst), IPR_E_ALIGN);
    long long blocks = (n / 8 + 255) / 256;
    const long long cap = (long long)ipr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    IPR_LAUNCH_PDL((gather_pack_kernel), (unsigned)blocks, 256, 0, ipr_cu(stream), src, index, (__nv_bfloat16 *)dst, (long long)n,
                   index_f32, dst_f32, (long long)n_f32);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
