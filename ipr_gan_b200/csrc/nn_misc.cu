// nn_misc.cu -- the memory-bound layers around the tensor-core GEMMs of the DCGAN step:
//   * im2col3: 3-channel NCHW fp32 image (optionally times tanh') -> [pixel][64] bf16 patch matrix
//     (K = 27 is far too thin for an implicit GEMM; one 128-byte row per pixel feeds the tap GEMM);
//   * BatchNorm (training mode): statistics finalisation from the GEMM epilogue's column partials,
//     fused scale/shift + ReLU, and the two-pass backward with the ReLU mask and the white-box
//     sign-loss gradient folded into d(gamma) (tools/sign_model.py:48 -> zero extra passes);
//   * the discriminator's final Linear(8192 -> 1): GEMV forward, and backward fused with LeakyReLU'.
// All bf16 tensors are NHWC; vector width 8 bf16 (16 bytes) per thread.
#include "ipr_common.cuh"
#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t *>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void ldg8(const float *p, float (&f)[8]) {      // p 16-byte aligned (channel index % 8 == 0)
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ------------------------------------------------------------------------------------ im2col3
// out[(n,h,w)][k], k = (kh*3+kw)*3 + c  <-  x[n, c, h+kh-1, w+kw-1] (zero outside), k in [27,32) = 0: 64-byte rows.
// The GEMMs read them through 64-channel TMA boxes whose upper half is out-of-bounds zero fill.
// With `t` given (tanh output), the source value is x * (1 - t^2): the Tanh backward of the generator's
// last layer fused into the gather.
__global__ void __launch_bounds__(256)
im2col3_kernel(const float *__restrict__ x, const float *__restrict__ t, uint4 *__restrict__ out,
               long long pixels, int H, int W)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const size_t plane = (size_t)H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += stride) {
        const int w = (int)(p % W);
        const long long r = p / W;
        const int h = (int)(r % H);
        const long long n = r / H;
        float v[32];
#pragma unroll
        for (int k = 27; k < 32; k++) v[k] = 0.0f;
#pragma unroll
        for (int kh = 0; kh < 3; kh++)
#pragma unroll
            for (int kw = 0; kw < 3; kw++) {
                const int hh = h + kh - 1, ww = w + kw - 1;
                const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float val = 0.0f;
                    if (ok) {
                        const size_t idx = ((size_t)n * 3 + c) * plane + (size_t)hh * W + ww;
                        val = __ldg(x + idx);
                        if (t) { const float tv = __ldg(t + idx); val *= (1.0f - tv * tv); }
                    }
                    v[(kh * 3 + kw) * 3 + c] = val;
                }
            }
        uint4 *dst = out + p * 4;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; i++) f[i] = v[g * 8 + i];
            dst[g] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------------------------ BatchNorm forward
// partial: [rows][2][C] column sums / sums of squares.  One warp-column per channel group; double accumulation.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float *__restrict__ partial, int rows, int C, double count, float eps, float momentum,
                   const float *__restrict__ gamma, const float *__restrict__ beta,
                   float *__restrict__ running_mean, float *__restrict__ running_var, long long *__restrict__ num_batches,
                   float *__restrict__ scale, float *__restrict__ shift, float *__restrict__ mean_out,
                   float *__restrict__ rstd_out)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ double s1[8][33], s2[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double a = 0.0, b = 0.0;
    if (c < C)
        for (int r = threadIdx.y; r < rows; r += 8) {
            a += (double)partial[(size_t)r * 2 * C + c];
            b += (double)partial[(size_t)r * 2 * C + C + c];
        }
    s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int y = 1; y < 8; y++) { a += s1[y][threadIdx.x]; b += s2[y][threadIdx.x]; }
        const double mean = a / count;
        double var = b / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma[c], be = beta[c];
        scale[c] = g * rstd;
        shift[c] = be - (float)mean * g * rstd;
        mean_out[c] = (float)mean;
        rstd_out[c] = rstd;
        if (running_mean) {
            const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
    if (num_batches && blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) *num_batches += 1;
}

// y = relu(x * scale[c] + shift[c]) over [rows][C] bf16.
// The grid stride (gridDim * 256 vectors) is a multiple of c_vec (<= 256, a power of two times ... the host checks), so
// a thread sees the SAME eight channels in every iteration: scale / shift are loaded once, and four 16-byte vectors
// per thread are in flight (one load per iteration kept 2.4 MB in flight chip-wide: 2.2 TB/s of 6.5).
constexpr int BN_UNROLL = 4;
__global__ void __launch_bounds__(256)
bn_apply_relu_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, const float *__restrict__ scale,
                     const float *__restrict__ shift, long long n_vec, int c_vec)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (int)(i % c_vec) * 8;
    float sc[8], sh[8];
    ldg8(scale + c0, sc);
    ldg8(shift + c0, sh);
    auto one = [&](const uint4 &in) {
        float f[8];
        unpack8(in, f);
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.0f);
        return pack8(f);
    };
    for (; i + (BN_UNROLL - 1) * stride < n_vec; i += BN_UNROLL * stride) {
        uint4 v[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; u++) v[u] = __ldg(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < BN_UNROLL; u++) y[i + u * stride] = one(v[u]);
    }
    for (; i < n_vec; i += stride) y[i] = one(__ldg(x + i));
}

// ------------------------------------------------------------------------------------ BatchNorm backward
// pass 1: per-CTA partial sums over a slab of rows of  g = dy * (act > 0)  and  g * xhat.
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ xraw, const float *__restrict__ scale,
                     const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd,
                     long long rows, int c_vec, float *__restrict__ partial)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    extern __shared__ float sm[];                    // [ry][2][C]
    const int C = c_vec * 8;
    const int ry_n = blockDim.x / c_vec;
    const int cv = threadIdx.x % c_vec, ry = threadIdx.x / c_vec;
    float sg[8], sx[8];
#pragma unroll
    for (int k = 0; k < 8; k++) sg[k] = sx[k] = 0.0f;
    if (ry < ry_n) {
        float mu[8], rs[8], sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            mu[k] = mean[cv * 8 + k]; rs[k] = rstd[cv * 8 + k]; sc[k] = scale[cv * 8 + k]; sh[k] = shift[cv * 8 + k];
        }
        auto one = [&](const uint4 &dq, const uint4 &xq) {
            float d[8], xv[8];
            unpack8(dq, d);
            unpack8(xq, xv);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                // ReLU mask recomputed exactly as the forward computed its input (same fmaf): no activation re-read
                const float g = fmaf(xv[k], sc[k], sh[k]) > 0.0f ? d[k] : 0.0f;
                sg[k] += g;
                sx[k] += g * (xv[k] - mu[k]) * rs[k];
            }
        };
        // four row pairs in flight per thread (the CTAs are few -- two per SM -- so the loads per thread decide the
        // bytes in flight: one pair per iteration ran at 2.8 TB/s)
        const long long rstep = (long long)gridDim.x * ry_n;
        long long r = (long long)blockIdx.x * ry_n + ry;
        for (; r + (BN_UNROLL - 1) * rstep < rows; r += BN_UNROLL * rstep) {
            uint4 dq[BN_UNROLL], xq[BN_UNROLL];
#pragma unroll
            for (int u = 0; u < BN_UNROLL; u++) {
                dq[u] = __ldg(dy + (r + u * rstep) * c_vec + cv);
                xq[u] = __ldg(xraw + (r + u * rstep) * c_vec + cv);
            }
#pragma unroll
            for (int u = 0; u < BN_UNROLL; u++) one(dq[u], xq[u]);
        }
        for (; r < rows; r += rstep) one(__ldg(dy + r * c_vec + cv), __ldg(xraw + r * c_vec + cv));
#pragma unroll
        for (int k = 0; k < 8; k++) { sm[(ry * 2) * C + cv * 8 + k] = sg[k]; sm[(ry * 2 + 1) * C + cv * 8 + k] = sx[k]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float acc = 0.0f;
        for (int y = 0; y < ry_n; y++) acc += sm[y * 2 * C + i];
        partial[(size_t)blockIdx.x * 2 * C + i] = acc;
    }
}

// pass 1b: dbeta, dgamma (+ sign-loss gradient) and the per-channel coefficients of pass 2.
//   dx = a[c] * g + b[c] * xraw + d[c]   with  a = gamma*rstd,  b = -gamma*rstd^2*dgamma_bn/M,
//   d = -a*dbeta/M - b*mean
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float *__restrict__ partial, int rows, int C, double count,
                       const float *__restrict__ gamma, const float *__restrict__ mean, const float *__restrict__ rstd,
                       float *__restrict__ dgamma, float *__restrict__ dbeta, int accumulate,
                       const float *__restrict__ sign, float gamma0, float sign_scale,
                       float *__restrict__ coef)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ double s1[32][33], s2[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double sg = 0.0, sx = 0.0;
    if (c < C)
        for (int r = ty; r < rows; r += 32) { sg += (double)partial[(size_t)r * 2 * C + c]; sx += (double)partial[(size_t)r * 2 * C + C + c]; }
    s1[ty][tx] = sg; s2[ty][tx] = sx;
    __syncthreads();
    if (ty != 0 || c >= C) return;
    for (int y = 1; y < 32; y++) { sg += s1[y][tx]; sx += s2[y][tx]; }
    const float g = gamma[c], rs = rstd[c], mu = mean[c];
    float dg = (float)sx;
    const float a = g * rs;
    const float b = -g * rs * rs * (float)(sx / count);
    coef[c] = a;
    coef[C + c] = b;
    coef[2 * C + c] = -a * (float)(sg / count) - b * mu;
    if (sign) {                                      // d/dgamma of mean_c relu(gamma0 - gamma*sign)
        const float sv = sign[c];
        if (gamma0 - g * sv > 0.0f) dg += -sv * sign_scale / (float)C;
    }
    if (accumulate == 2) {
        // two backward passes of the same network (G(z) and G(trigger)) may run on different streams: a float add of
        // exactly two contributions onto a zeroed slot is order-independent, so the atomics stay deterministic
        atomicAdd(dgamma + c, dg);
        atomicAdd(dbeta + c, (float)sg);
    } else {
        dgamma[c] = accumulate ? dgamma[c] + dg : dg;
        dbeta[c] = accumulate ? dbeta[c] + (float)sg : (float)sg;
    }
}

// pass 2: dx = a*g + b*xraw + d
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ xraw, const float *__restrict__ scale,
                    const float *__restrict__ shift, const float *__restrict__ coef, uint4 *__restrict__ dx,
                    long long n_vec, int c_vec)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int C = c_vec * 8;
    const long long stride = (long long)gridDim.x * blockDim.x;        // a multiple of c_vec: same channels every iteration
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = (int)(i % c_vec) * 8;
    float sc[8], sh[8], ca[8], cb[8], cd[8];
    ldg8(scale + c0, sc); ldg8(shift + c0, sh);
    ldg8(coef + c0, ca); ldg8(coef + C + c0, cb); ldg8(coef + 2 * C + c0, cd);
    auto one = [&](const uint4 &dq, const uint4 &xq) {
        float d[8], xv[8], o[8];
        unpack8(dq, d);
        unpack8(xq, xv);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float g = fmaf(xv[k], sc[k], sh[k]) > 0.0f ? d[k] : 0.0f;
            o[k] = ca[k] * g + cb[k] * xv[k] + cd[k];
        }
        return pack8(o);
    };
    for (; i + (BN_UNROLL - 1) * stride < n_vec; i += BN_UNROLL * stride) {
        uint4 dq[BN_UNROLL], xq[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; u++) { dq[u] = __ldg(dy + i + u * stride); xq[u] = __ldg(xraw + i + u * stride); }
#pragma unroll
        for (int u = 0; u < BN_UNROLL; u++) dx[i + u * stride] = one(dq[u], xq[u]);
    }
    for (; i < n_vec; i += stride) dx[i] = one(__ldg(dy + i), __ldg(xraw + i));
}

// ------------------------------------------------------------------------------------ final Linear(K -> 1)
// logits[b] = dot(a[b,:], w) / sigma + bias          one CTA per sample (fixed-order block reduction)
__global__ void __launch_bounds__(256)
dfc_fwd_kernel(const uint4 *__restrict__ a, const float *__restrict__ w, const float *__restrict__ sigma,
               const float *__restrict__ bias, float *__restrict__ logits, int batch, int k_vec)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const int b = blockIdx.x;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < k_vec; i += blockDim.x) {
        float f[8];
        unpack8(__ldg(a + (size_t)b * k_vec + i), f);
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(w) + 2 * i), w1 = __ldg(reinterpret_cast<const float4 *>(w) + 2 * i + 1);
        acc += f[0] * w0.x + f[1] * w0.y + f[2] * w0.z + f[3] * w0.w + f[4] * w1.x + f[5] * w1.y + f[6] * w1.z + f[7] * w1.w;
    }
    acc = ipr_block_sum(acc, red);
    if (threadIdx.x == 0) logits[b] = acc / (sigma ? *sigma : 1.0f) + (bias ? *bias : 0.0f);
}

// da[b,k] = dlogit[b] * w[k] / sigma * lrelu'(a[b,k])      (gradient w.r.t. the previous conv's pre-activation)
__global__ void __launch_bounds__(256)
dfc_bwd_data_kernel(const uint4 *__restrict__ a, const float *__restrict__ w, const float *__restrict__ sigma,
                    const float *__restrict__ dlogit, uint4 *__restrict__ da, long long n_vec, int k_vec, float slope)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const float inv = 1.0f / (sigma ? *sigma : 1.0f);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const int kv = (int)(i % k_vec);
        const long long b = i / k_vec;
        const float dl = __ldg(dlogit + b) * inv;
        float f[8], o[8];
        unpack8(__ldg(a + i), f);
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(w) + 2 * kv), w1 = __ldg(reinterpret_cast<const float4 *>(w) + 2 * kv + 1);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = dl * wv[k] * (f[k] > 0.0f ? 1.0f : slope);
        da[i] = pack8(o);
    }
}

// dw[k] (+)= sum_b dlogit[b] * a[b,k]   (w.r.t. the normalised weight); CTA = 32 columns x 8 batch lanes, fixed order
__global__ void __launch_bounds__(256)
dfc_bwd_weight_kernel(const __nv_bfloat16 *__restrict__ a, const float *__restrict__ dlogit, float *__restrict__ dw,
                      int batch, int K, int accumulate, const int *__restrict__ dw_index, float *__restrict__ dbias)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float sm[8][33];
    if (dbias && blockIdx.x == 0 && threadIdx.x < 32) {          // d(bias) = sum_b dlogit[b], fixed order
        float s = 0.0f;
        for (int b = threadIdx.x; b < batch; b += 32) s += __ldg(dlogit + b);
        s = ipr_warp_sum(s);
        if (threadIdx.x == 0) atomicAdd(dbias, s);               // always accumulates (zeroed gradient arena)
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (k < K)
        for (int b = ty; b < batch; b += 8) acc += __ldg(dlogit + b) * __bfloat162float(a[(size_t)b * K + k]);
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && k < K) {
#pragma unroll
        for (int y = 1; y < 8; y++) acc += sm[y][tx];
        const int o = dw_index ? __ldg(dw_index + k) : k;            // NHWC feature -> position in the parameter
        dw[o] = accumulate ? dw[o] + acc : acc;
    }
}

inline unsigned grid_1d(long long items, int threads, int waves = 8) {
    long long blocks = (items + threads - 1) / threads;
    const long long cap = (long long)ipr_sm_count() * waves;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

// grid for the kernels that keep their per-channel constants in registers: gridDim * 256 must be a multiple of c_vec
inline unsigned grid_channels(long long n_vec, int c_vec) {
    int g = 256, c = c_vec;
    while (c) { const int t = g % c; g = c; c = t; }           // g = gcd(256, c_vec)
    const unsigned need = (unsigned)(c_vec / g);
    unsigned blocks = grid_1d(n_vec, 256);
    blocks = blocks / need * need;
    return blocks < need ? need : blocks;
}

}  // namespace

// ------------------------------------------------------------------------------------ col2im3
// The 3-channel ends of the networks as ONE pass over the 64-channel activation: a plain GEMM produces
// T[pixel][(kh*3+kw)*3 + c] = sum_k a[pixel][k] * W[k][c][kh][kw] (27 of 32 columns used), and this kernel folds the
// nine taps back:  out[b][c][h][w] = act( sum_{kh,kw} T[(b, h+1-kh, w+1-kw)][(kh*3+kw)*3 + c] ).
// Replaces a 9-tap implicit GEMM with N padded from 3 to 16 that re-streamed the activation nine times.
constexpr int C2I_TH = 8, C2I_TW = 32;                   // output tile of one CTA (256 threads, one pixel each)
constexpr int C2I_PITCH = 29;                            // floats per staged source pixel (27 used; odd => conflict-free)

__global__ void __launch_bounds__(C2I_TH * C2I_TW)
col2im3_kernel(const float *__restrict__ t, float *__restrict__ out, int height, int width, int ld, int tanh_out)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    // stage the (TH+2) x (TW+2) halo of source pixels (28 floats each, coalesced 128-bit loads), then every thread
    // sums its nine taps out of shared memory: the tap-expanded tensor is read ~1.3x instead of gathered 27x
    __shared__ float sm[(C2I_TH + 2) * (C2I_TW + 2) * C2I_PITCH];
    const int w0 = blockIdx.x * C2I_TW, h0 = blockIdx.y * C2I_TH;
    const long long b = blockIdx.z;
    const float *tb = t + b * (long long)height * width * ld;
    constexpr int HALO_W = C2I_TW + 2, HALO = (C2I_TH + 2) * HALO_W;
    // five 128-bit loads per thread in flight before the first shared-memory store (one at a time ran at 2.5 TB/s)
    constexpr int C2I_UN = 5;
    for (int i0 = threadIdx.x; i0 < HALO * 7; i0 += C2I_UN * (C2I_TH * C2I_TW)) {
        float4 v[C2I_UN];
#pragma unroll
        for (int u = 0; u < C2I_UN; u++) {
            const int i = i0 + u * (C2I_TH * C2I_TW);
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < HALO * 7) {
                const int pix = i / 7, v4 = i - pix * 7;
                const int hr = pix / HALO_W, hc = pix - hr * HALO_W;
                const int hh = h0 + hr - 1, ww = w0 + hc - 1;
                if (hh >= 0 && hh < height && ww >= 0 && ww < width)
                    v[u] = ipr_ldg_stream4(reinterpret_cast<const float4 *>(tb + ((long long)hh * width + ww) * ld) + v4);
            }
        }
#pragma unroll
        for (int u = 0; u < C2I_UN; u++) {
            const int i = i0 + u * (C2I_TH * C2I_TW);
            if (i < HALO * 7) {
                const int pix = i / 7, v4 = i - pix * 7;
                float *d = sm + pix * C2I_PITCH + v4 * 4;
                d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; if (v4 < 6) d[3] = v[u].w;       // column 27 is padding
            }
        }
    }
    __syncthreads();
    const int r = threadIdx.x / C2I_TW, c = threadIdx.x - r * C2I_TW;
    const int h = h0 + r, w = w0 + c;
    if (h >= height || w >= width) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
#pragma unroll
        for (int kw = 0; kw < 3; kw++) {
            // source pixel (h+1-kh, w+1-kw) -> halo coordinates (r+2-kh, c+2-kw); out-of-image pixels were staged as 0
            const float *src = sm + ((r + 2 - kh) * HALO_W + (c + 2 - kw)) * C2I_PITCH + (kh * 3 + kw) * 3;
            a0 += src[0]; a1 += src[1]; a2 += src[2];
        }
    }
    if (tanh_out) { a0 = tanhf(a0); a1 = tanhf(a1); a2 = tanhf(a2); }
    const long long plane = (long long)height * width;
    float *o = out + b * 3 * plane + (long long)h * width + w;
    o[0] = a0; o[plane] = a1; o[2 * plane] = a2;
}

extern "C" int ipr_col2im3_f32(const float *t, float *out, int64_t batch, int height, int width, int ld, int tanh_out,
                               ipr_stream_t stream)
{
    IPR_REQUIRE(t && out, IPR_E_NULL);
    IPR_REQUIRE(batch > 0 && batch < 65536 && height > 0 && width > 0 && ld >= 28 && (ld & 3) == 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(t), IPR_E_ALIGN);
    const dim3 grid((unsigned)((width + C2I_TW - 1) / C2I_TW), (unsigned)((height + C2I_TH - 1) / C2I_TH), (unsigned)batch);
    IPR_LAUNCH_PDL((col2im3_kernel), grid, C2I_TH * C2I_TW, 0, ipr_cu(stream), t, out, height, width, ld, tanh_out);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_im2col3_bf16(const float *x, const float *tanh_out, void *out, int64_t batch, int height,
                                int width, ipr_stream_t stream)
{
    IPR_REQUIRE(x && out, IPR_E_NULL);
    IPR_REQUIRE(batch > 0 && height > 0 && width > 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(out), IPR_E_ALIGN);
    const long long pixels = (long long)batch * height * width;
    IPR_LAUNCH_PDL((im2col3_kernel), grid_1d(pixels, 256), 256, 0, ipr_cu(stream), x, tanh_out, (uint4 *)out, pixels, height, width);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_bn_finalize_f32(const float *partial, int rows, int channels, double count, float eps,
                                   float momentum, const float *gamma, const float *beta, float *running_mean,
                                   float *running_var, int64_t *num_batches_tracked, float *scale, float *shift,
                                   float *mean, float *rstd, ipr_stream_t stream)
{
    IPR_REQUIRE(partial && gamma && beta && scale && shift && mean && rstd, IPR_E_NULL);
    IPR_REQUIRE(rows > 0 && channels > 0 && count > 0, IPR_E_SHAPE);
    dim3 block(32, 8);
    IPR_LAUNCH_PDL((bn_finalize_kernel), (channels + 31) / 32, block, 0, ipr_cu(stream), 
        partial, rows, channels, count, eps, momentum, gamma, beta, running_mean, running_var,
        (long long *)num_batches_tracked, scale, shift, mean, rstd);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_bn_apply_relu_bf16(const void *x, void *y, const float *scale, const float *shift,
                                      int64_t rows, int channels, ipr_stream_t stream)
{
    IPR_REQUIRE(x && y && scale && shift, IPR_E_NULL);
    IPR_REQUIRE(rows > 0 && channels > 0 && channels % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(x) && ipr_aligned16(y), IPR_E_ALIGN);
    const long long n_vec = (long long)rows * channels / 8;
    IPR_LAUNCH_PDL((bn_apply_relu_kernel), grid_channels(n_vec, channels / 8), 256, 0, ipr_cu(stream), (const uint4 *)x, (uint4 *)y, scale, shift,
                                                                        n_vec, channels / 8);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" size_t ipr_bn_bwd_workspace_bytes(int channels)
{
    // partial sums [ctas][2][C] + coefficients [3][C]
    return ((size_t)ipr_sm_count() * 2 * 2 * channels + 3 * (size_t)channels) * sizeof(float);
}

extern "C" int ipr_bn_relu_bwd_bf16(const void *dy, const void *xraw, const float *scale, const float *shift,
                                    const float *gamma,
                                    const float *mean, const float *rstd, void *dx, float *dgamma, float *dbeta,
                                    int accumulate, const float *sign, float gamma0, float sign_scale,
                                    void *workspace, size_t workspace_bytes, int64_t rows, int channels,
                                    ipr_stream_t stream)
{
    IPR_REQUIRE(dy && xraw && scale && shift && gamma && mean && rstd && dx && dgamma && dbeta && workspace, IPR_E_NULL);
    IPR_REQUIRE(rows > 0 && channels > 0 && channels % 8 == 0 && channels <= 2048, IPR_E_SHAPE);
    IPR_REQUIRE(workspace_bytes >= ipr_bn_bwd_workspace_bytes(channels), IPR_E_WORKSPACE);
    IPR_REQUIRE(ipr_aligned16(dy) && ipr_aligned16(xraw) && ipr_aligned16(dx), IPR_E_ALIGN);
    const int c_vec = channels / 8;
    IPR_REQUIRE(c_vec <= 256, IPR_E_UNSUPPORTED);
    const int ry_n = 256 / c_vec;
    long long ctas = (rows + ry_n - 1) / ry_n;
    const long long cap = (long long)ipr_sm_count() * 2;
    if (ctas > cap) ctas = cap;
    float *partial = (float *)workspace;
    float *coef = partial + (size_t)ipr_sm_count() * 2 * 2 * channels;
    cudaStream_t st = ipr_cu(stream);
    const size_t smem = (size_t)ry_n * 2 * channels * sizeof(float);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(bn_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    IPR_LAUNCH_PDL((bn_bwd_reduce_kernel), (unsigned)ctas, 256, smem, st, (const uint4 *)dy, (const uint4 *)xraw, scale, shift,
                                                            mean, rstd, rows, c_vec, partial);
    IPR_LAUNCH_CHECK();
    IPR_LAUNCH_PDL((bn_bwd_finalize_kernel), (channels + 31) / 32, 1024, 0, st, partial, (int)ctas, channels, (double)rows, gamma, mean,
                                                                  rstd, dgamma, dbeta, accumulate, sign, gamma0,
                                                                  sign_scale, coef);
    IPR_LAUNCH_CHECK();
    const long long n_vec = (long long)rows * c_vec;
    IPR_LAUNCH_PDL((bn_bwd_apply_kernel), grid_channels(n_vec, c_vec), 256, 0, st, (const uint4 *)dy, (const uint4 *)xraw, scale, shift,
                                                             coef, (uint4 *)dx, n_vec, c_vec);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_dfc_fwd_bf16(const void *a, const float *w, const float *sigma, const float *bias,
                                float *logits, int batch, int k, ipr_stream_t stream)
{
    IPR_REQUIRE(a && w && logits, IPR_E_NULL);
    IPR_REQUIRE(batch > 0 && k > 0 && k % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(a) && ipr_aligned16(w), IPR_E_ALIGN);
    IPR_LAUNCH_PDL((dfc_fwd_kernel), batch, 256, 0, ipr_cu(stream), (const uint4 *)a, w, sigma, bias, logits, batch, k / 8);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_dfc_bwd_bf16(const void *a, const float *w, const float *sigma, const float *dlogit,
                                void *da, float *dw, int accumulate_dw, float slope, int batch, int k,
                                const int32_t *dw_index, float *dbias, ipr_stream_t stream)
{
    IPR_REQUIRE(a && w && dlogit && da, IPR_E_NULL);
    IPR_REQUIRE(batch > 0 && k > 0 && k % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(a) && ipr_aligned16(w) && ipr_aligned16(da), IPR_E_ALIGN);
    const long long n_vec = (long long)batch * k / 8;
    IPR_LAUNCH_PDL((dfc_bwd_data_kernel), grid_1d(n_vec, 256), 256, 0, ipr_cu(stream), (const uint4 *)a, w, sigma, dlogit, (uint4 *)da,
                                                                        n_vec, k / 8, slope);
    IPR_LAUNCH_CHECK();
    if (dw) {
        IPR_LAUNCH_PDL((dfc_bwd_weight_kernel), (k + 31) / 32, 256, 0, ipr_cu(stream), (const __nv_bfloat16 *)a, dlogit, dw, batch, k,
                       accumulate_dw, dw_index, dbias);
        IPR_LAUNCH_CHECK();
    }
    return IPR_OK;
}

// ------------------------------------------------------------------------------------ column reductions
namespace {

// stage 1: out[g][c] = sum over the rows assigned to group g of in[r][c]   (fp32 partial rows, e.g. GEMM-epilogue
// statistics).  CTA = 32 columns x 8 row lanes; grid = (ncols/32, G).
__global__ void __launch_bounds__(256)
colsum_partials_stage1(const float *__restrict__ in, int rows, int ncols, int row_stride, float *__restrict__ out)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (c < ncols)
        for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8) acc += in[(size_t)r * row_stride + c];
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < ncols) {
#pragma unroll
        for (int y = 1; y < 8; y++) acc += sm[y][tx];
        out[(size_t)blockIdx.y * ncols + c] = acc;
    }
}
// few rows: out[c] (+)= scale * sum_r in[r][c] in one launch (32 columns x 8 row lanes per CTA, fixed order)
__global__ void __launch_bounds__(256)
colsum_small_kernel(const float *__restrict__ in, int rows, int ncols, int row_stride, float *__restrict__ out,
                    int accumulate, float scale)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (c < ncols)
        for (int r = ty; r < rows; r += 8) acc += in[(size_t)r * row_stride + c];
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < ncols) {
#pragma unroll
        for (int y = 1; y < 8; y++) acc += sm[y][tx];
        out[c] = accumulate ? out[c] + acc * scale : acc * scale;
    }
}
// stage 2: out[c] (+)= scale * sum_g in[g][c], double accumulation, fixed order
__global__ void __launch_bounds__(256)
colsum_stage2(const float *__restrict__ in, int rows, int ncols, float *__restrict__ out, int accumulate, float scale,
              const int *__restrict__ out_index)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    // 32 columns x 8 row lanes per CTA (the first version walked the rows with ONE thread per column: 10 us of pure
    // load latency per launch, 20 launches per step); fixed order: lane partials, then lanes 0..7
    __shared__ double sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double acc = 0.0;
    if (c < ncols)
        for (int r = ty; r < rows; r += 8) acc += (double)in[(size_t)r * ncols + c];
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty != 0 || c >= ncols) return;
#pragma unroll
    for (int y = 1; y < 8; y++) acc += sm[y][tx];
    const float v = (float)acc * scale;
    const int o = out_index ? __ldg(out_index + c) : c;
    if (accumulate == 2) atomicAdd(out + o, v);                   // two-contribution accumulation across streams
    else out[o] = accumulate ? out[o] + v : v;
}
// stage 1 for a bf16 [rows][C] tensor: each CTA owns a slab of rows; 8 channels per thread, 256 / c_vec row lanes per
// CTA combined through shared memory in a fixed order.  (With one row lane per CTA a 64-channel tensor kept 8 threads
// of each CTA busy: 41 us per launch at 147 k rows.)
__global__ void __launch_bounds__(256)
colsum_bf16_stage1(const uint4 *__restrict__ x, long long rows, int c_vec, int ry_n, float *__restrict__ out)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    extern __shared__ float cs_sm[];                 // [ry_n][cw * 8], cw = column vectors of this CTA
    // grid.x = column groups of `cw` vectors, grid.y = row slabs; thread = (row lane, column vector)
    const int cw = blockDim.x / ry_n;
    const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
    const int cv = blockIdx.x * cw + tx;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.0f;
    if (cv < c_vec && ty < ry_n)
        for (long long r = (long long)blockIdx.y * ry_n + ty; r < rows; r += (long long)gridDim.y * ry_n) {
            float f[8];
            unpack8(__ldg(x + r * c_vec + cv), f);
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k] += f[k];
        }
    if (ty < ry_n) {
#pragma unroll
        for (int k = 0; k < 8; k++) cs_sm[(ty * cw + tx) * 8 + k] = acc[k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cw * 8; i += blockDim.x) {
        const int col = blockIdx.x * cw * 8 + i;
        if (col >= c_vec * 8) continue;
        float t = 0.0f;
        for (int y = 0; y < ry_n; y++) t += cs_sm[y * cw * 8 + i];
        out[(size_t)blockIdx.y * c_vec * 8 + col] = t;
    }
}

}  // namespace

extern "C" size_t ipr_colsum_workspace_bytes(int ncols) { return (size_t)64 * ncols * sizeof(float); }

extern "C" int ipr_colsum_partials_f32(const float *partial, int rows, int ncols, int row_stride, float *out,
                                       int accumulate, float scale, void *workspace, size_t workspace_bytes,
                                       ipr_stream_t stream)
{
    IPR_REQUIRE(row_stride >= ncols, IPR_E_SHAPE);
    IPR_REQUIRE(partial && out && workspace, IPR_E_NULL);
    IPR_REQUIRE(rows > 0 && ncols > 0, IPR_E_SHAPE);
    IPR_REQUIRE(workspace_bytes >= ipr_colsum_workspace_bytes(ncols), IPR_E_WORKSPACE);
    if (rows <= 256) {                      // few rows: one launch, every CTA reduces its 32 columns completely
        IPR_LAUNCH_PDL((colsum_small_kernel), (ncols + 31) / 32, 256, 0, ipr_cu(stream), partial, rows, ncols, row_stride, out, accumulate, scale);
        IPR_LAUNCH_CHECK();
        return IPR_OK;
    }
    const int G = rows < 8 * 64 ? (rows + 7) / 8 : 64;
    dim3 grid((ncols + 31) / 32, G);
    IPR_LAUNCH_PDL((colsum_partials_stage1), grid, 256, 0, ipr_cu(stream), partial, rows, ncols, row_stride, (float *)workspace);
    IPR_LAUNCH_CHECK();
    IPR_LAUNCH_PDL((colsum_stage2), (ncols + 31) / 32, 256, 0, ipr_cu(stream), (const float *)workspace, G, ncols, out, accumulate, scale,
                   (const int *)nullptr);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_colsum_bf16(const void *x, int64_t rows, int channels, float *out, int accumulate, float scale,
                               const int32_t *out_index, void *workspace, size_t workspace_bytes, ipr_stream_t stream)
{
    IPR_REQUIRE(x && out && workspace, IPR_E_NULL);
    IPR_REQUIRE(rows > 0 && channels > 0 && channels % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(workspace_bytes >= ipr_colsum_workspace_bytes(channels), IPR_E_WORKSPACE);
    IPR_REQUIRE(ipr_aligned16(x), IPR_E_ALIGN);
    const int c_vec = channels / 8;
    const int G = rows < 64 ? (int)rows : 64;
    // 256 threads = cw column vectors x ry_n row lanes (cw = c_vec rounded up to a power of two, at most 256)
    int cw = 1;
    while (cw < c_vec && cw < 256) cw <<= 1;
    int ry_n = 256 / cw;
    while (ry_n > 1 && (long long)G * ry_n > rows) ry_n >>= 1;
    dim3 grid((c_vec + cw - 1) / cw, G);
    IPR_LAUNCH_PDL((colsum_bf16_stage1), grid, cw * ry_n, (size_t)ry_n * cw * 8 * sizeof(float), ipr_cu(stream), (const uint4 *)x, rows,
                   c_vec, ry_n, (float *)workspace);
    IPR_LAUNCH_CHECK();
    IPR_LAUNCH_PDL((colsum_stage2), (channels + 31) / 32, 256, 0, ipr_cu(stream), (const float *)workspace, G, channels, out,
                   accumulate, scale, out_index);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
