// pdq.cu -- watermark verification: bicubic up-sampling (torch CPU operation order), float->uint8,
// PDQ 256-bit perceptual hash, Hamming distance and binomial p-value lookup.
//
// Replaces tools/phash_pvalue.py:7-38 (a per-image Python loop over pdqhash.compute on the CPU) with
// batched kernels: one CTA per image for the hash.  Everything here is integer / ordered-fp32 work
// whose results must be BIT-EXACT against the CPU oracle, so FMA contraction is controlled
// explicitly: __fmul_rn/__fadd_rn where the CPU code rounds twice, __fmaf_rn where torch's
// AVX2/AVX512 build fuses (the pattern was established empirically, tests/test_gpu_ipr_ops.py::test_bicubic_bit_exact, tests/bicubic_ref.py).
#include "ipr_common.cuh"
#include <math.h>

namespace {

// ------------------------------------------------------------------------------------- bicubic
// Index/weight rule of at::native upsample_bicubic2d (align_corners = False, A = -0.75):
//   real = fma(scale, i + 0.5, -0.5), scale = in/out (fp32);  i0 = floor(real);  t = real - i0
//   w0 = cc2(t + 1), w1 = cc1(t), w2 = cc1(1 - t), w3 = cc2((1 - t) + 1)
//   cc1(x) = (fma(A+2, x, -(A+3)) x) x + 1           (first product fused, rest rounded separately)
//   cc2(x) = fma(fma(A, x, -5A), x, 8A) x - 4A       (first two fused, last product and add separate)
// taps at clamp(i0 - 1 + j, 0, in - 1).  Accumulation: r = fma(t0, w0, t1*w1); r = fma(t2, w2, r);
// r = fma(t3, w3, r); applied along W first (inner), then along H (outer).
__device__ __forceinline__ float cc1(float x) {
    float p = __fmaf_rn(1.25f, x, -2.25f);
    return __fadd_rn(__fmul_rn(__fmul_rn(p, x), x), 1.0f);
}
__device__ __forceinline__ float cc2(float x) {
    float t = __fmaf_rn(-0.75f, x, 3.75f);
    t = __fmaf_rn(t, x, -6.0f);
    return __fadd_rn(__fmul_rn(t, x), 3.0f);
}
__device__ __forceinline__ void cubic_taps(int i, float scale, int in_size, int idx[4], float w[4]) {
    const float real = __fmaf_rn(scale, __fadd_rn((float)i, 0.5f), -0.5f);
    int i0 = (int)floorf(real);
    i0 = min(i0, in_size - 1);
    float t = __fsub_rn(real, (float)i0);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    const float u = __fsub_rn(1.0f, t);
    w[0] = cc2(__fadd_rn(t, 1.0f)); w[1] = cc1(t); w[2] = cc1(u); w[3] = cc2(__fadd_rn(u, 1.0f));
#pragma unroll
    for (int j = 0; j < 4; j++) idx[j] = min(max(i0 - 1 + j, 0), in_size - 1);
}
__device__ __forceinline__ float cubic_mix(const float t[4], const float w[4]) {
    float r = __fmaf_rn(t[0], w[0], __fmul_rn(t[1], w[1]));
    r = __fmaf_rn(t[2], w[2], r);
    return __fmaf_rn(t[3], w[3], r);
}

__global__ void __launch_bounds__(256)
bicubic_kernel(const float *__restrict__ x, float *__restrict__ out, long long planes, int hin, int win,
               int hout, int wout, float scale_h, float scale_w)
{
    const long long total = planes * hout * wout;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int ow = (int)(e % wout);
        const long long t1 = e / wout;
        const int oh = (int)(t1 % hout);
        const long long pl = t1 / hout;
        int ih[4], iw[4]; float wh[4], ww[4];
        cubic_taps(oh, scale_h, hin, ih, wh);
        cubic_taps(ow, scale_w, win, iw, ww);
        const float *src = x + (size_t)pl * hin * win;
        float rows[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const float *r = src + (size_t)ih[a] * win;
            float t[4] = {__ldg(r + iw[0]), __ldg(r + iw[1]), __ldg(r + iw[2]), __ldg(r + iw[3])};
            rows[a] = cubic_mix(t, ww);
        }
        out[e] = cubic_mix(rows, wh);
    }
}

// ------------------------------------------------------------------------------------- PDQ hash
constexpr int PDQ_MAX = 64;       // max image edge handled in shared memory
constexpr int PDQ_THREADS = 256;

// sum / count with the division skipped for count == 1 (x / 1.0f == x exactly): images of 32..128 pixels have a window
// of ONE, where the filter still has to be executed (its running sum (a + b) - a is not b in floating point) but every
// IEEE division -- ~25 dependent instructions in a chain only H or W threads of the CTA execute -- is a no-op.
__device__ __forceinline__ float box_avg(float sum, int cur) { return cur == 1 ? sum : __fdiv_rn(sum, (float)cur); }

// 1-D running-sum box filter with ThreatExchange's window-growth rules; sequential by definition.
__device__ void box_1d(const float *in, float *out, int n, int stride, int win)
{
    const int half = (win + 2) / 2;
    const int n1 = half - 1, n2 = win - half + 1, n3 = n - win, n4 = half - 1;
    int li = 0, ri = 0, oi = 0, cur = 0;
    float sum = 0.0f;
    for (int k = 0; k < n1; k++) { sum = __fadd_rn(sum, in[ri]); cur++; ri += stride; }
    for (int k = 0; k < n2; k++) { sum = __fadd_rn(sum, in[ri]); cur++; out[oi] = box_avg(sum, cur); ri += stride; oi += stride; }
    for (int k = 0; k < n3; k++) {
        sum = __fadd_rn(sum, in[ri]); sum = __fsub_rn(sum, in[li]);
        out[oi] = box_avg(sum, cur); li += stride; ri += stride; oi += stride;
    }
    for (int k = 0; k < n4; k++) { sum = __fsub_rn(sum, in[li]); cur--; out[oi] = box_avg(sum, cur); li += stride; oi += stride; }
}

__device__ __forceinline__ float to_u8_float(float v) {
    // (v * 255) -> uint8 the way torch's .byte() does on CPU: truncate toward zero, wrap mod 256
    const int t = __float2int_rz(__fmul_rn(v, 255.0f));
    return (float)(t & 255);
}

__global__ void __launch_bounds__(PDQ_THREADS)
pdq_hash_kernel(const float *__restrict__ img, uint32_t *__restrict__ hash, float *__restrict__ coeffs_out,
                const float *__restrict__ dct, int H, int W)
{
    // the two image planes are sized by the actual image (32 x 32 after up-sampling the 16 x 16 crops: 8.4 KB instead
    // of 33 KB for the 64 x 64 maximum), which takes the CTAs per SM from 5 to the 8 the thread count allows
    extern __shared__ float pdq_planes[];
    float *b1 = pdq_planes, *b2 = pdq_planes + H * (W + 1);
    __shared__ float D[16 * 65];           // padded rows: D[i*65 + k]
    __shared__ float T[16 * 65];           // T[i*65 + j]
    __shared__ float Cf[256];
    __shared__ float med;
    const int tid = threadIdx.x;
    const int P = W + 1;                   // odd-ish pitch: row-sequential threads hit distinct banks
    const size_t plane = (size_t)H * W;
    const float *src = img + (size_t)blockIdx.x * 3 * plane;

    for (int e = tid; e < 16 * 64; e += PDQ_THREADS) D[(e >> 6) * 65 + (e & 63)] = __ldg(dct + e);
    // luma = 0.299 R + 0.587 G + 0.114 B on the uint8-converted image, three rounded products, two rounded adds
    for (int e = tid; e < H * W; e += PDQ_THREADS) {
        const int r = e / W, c = e - r * W;
        const float R = to_u8_float(__ldg(src + e)), G = to_u8_float(__ldg(src + plane + e)),
                    B = to_u8_float(__ldg(src + 2 * plane + e));
        float acc = __fmul_rn(0.299f, R);
        acc = __fadd_rn(acc, __fmul_rn(0.587f, G));
        acc = __fadd_rn(acc, __fmul_rn(0.114f, B));
        b1[r * P + c] = acc;
    }
    __syncthreads();
    if (!(H == 64 && W == 64)) {
        const int wr = (W + 127) / 128, wc = (H + 127) / 128;
        for (int rep = 0; rep < 2; rep++) {
            for (int r = tid; r < H; r += PDQ_THREADS) box_1d(b1 + r * P, b2 + r * P, W, 1, wr);
            __syncthreads();
            for (int c = tid; c < W; c += PDQ_THREADS) box_1d(b2 + c, b1 + c, H, P, wc);
            __syncthreads();
        }
    }
    // T = D (16x64) * A (64x64), A[k][j] = b1[dec(k)][dec(j)]; sequential k, separately rounded mul and add.
    // The decimation indices int((k + 0.5) * dim / 64) (double arithmetic, as ThreatExchange computes them) are
    // tabulated once per CTA: evaluating them inside the k loop cost a double-precision divide per multiply-add and
    // made this kernel ~10x slower than its arithmetic.  Each thread owns column j of rows i0, i0+4, i0+8, i0+12:
    // four independent accumulation chains share every A[k][j] load.
    __shared__ int dec_r[64], dec_c[64];
    if (tid < 64) {
        const bool full = (H == 64 && W == 64);
        dec_r[tid] = full ? tid : (int)(((tid + 0.5) * H) / 64);
        dec_c[tid] = full ? tid : (int)(((tid + 0.5) * W) / 64);
    }
    __syncthreads();
    {
        const int j = tid & 63, i0 = tid >> 6;
        const int jj = dec_c[j];
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll 8
        for (int k = 0; k < 64; k++) {
            const float a = b1[dec_r[k] * P + jj];
            s0 = __fadd_rn(s0, __fmul_rn(D[(i0 + 0) * 65 + k], a));
            s1 = __fadd_rn(s1, __fmul_rn(D[(i0 + 4) * 65 + k], a));
            s2 = __fadd_rn(s2, __fmul_rn(D[(i0 + 8) * 65 + k], a));
            s3 = __fadd_rn(s3, __fmul_rn(D[(i0 + 12) * 65 + k], a));
        }
        T[(i0 + 0) * 65 + j] = s0; T[(i0 + 4) * 65 + j] = s1; T[(i0 + 8) * 65 + j] = s2; T[(i0 + 12) * 65 + j] = s3;
    }
    __syncthreads();
    // B = T * D^T (16x16)
    {
        const int i = tid >> 4, j = tid & 15;
        float s = 0.0f;
        for (int k = 0; k < 64; k++) s = __fadd_rn(s, __fmul_rn(T[i * 65 + k], D[j * 65 + k]));
        Cf[tid] = s;
        if (coeffs_out) coeffs_out[(size_t)blockIdx.x * 256 + tid] = s;
    }
    __syncthreads();
    // median = 128th smallest of the 256 coefficients (what Torben's method returns for n = 256)
    {
        const float v = Cf[tid];
        int less = 0, leq = 0;
        for (int k = 0; k < 256; k++) { const float u = Cf[k]; less += (u < v); leq += (u <= v); }
        if (less < 128 && leq >= 128) med = v;      // all writers hold the same value
    }
    __syncthreads();
    const unsigned bits = __ballot_sync(0xffffffffu, Cf[tid] > med);
    if ((tid & 31) == 0) hash[(size_t)blockIdx.x * 8 + (tid >> 5)] = bits;
}

__global__ void __launch_bounds__(256)
hash_pvalue_kernel(const uint32_t *__restrict__ hx, const uint32_t *__restrict__ hy,
                   const float *__restrict__ ptable, float *__restrict__ p, int *__restrict__ r, long long batch)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= batch) return;
    const uint4 *a = reinterpret_cast<const uint4 *>(hx + n * 8);
    const uint4 *b = reinterpret_cast<const uint4 *>(hy + n * 8);
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                  __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
    const int rr = 256 - d;
    if (r) r[n] = rr;
    p[n] = __ldg(ptable + rr);
}

}  // namespace

extern "C" int ipr_bicubic_resize_f32(const float *x, float *out, int64_t planes,
                                      int hin, int win, int hout, int wout, ipr_stream_t stream)
{
    IPR_REQUIRE(x && out, IPR_E_NULL);
    IPR_REQUIRE(planes > 0 && hin > 0 && win > 0 && hout > 0 && wout > 0, IPR_E_SHAPE);
    const long long total = (long long)planes * hout * wout;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ipr_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const float sh = (float)hin / (float)hout, sw = (float)win / (float)wout;
    bicubic_kernel<<<(unsigned)blocks, 256, 0, ipr_cu(stream)>>>(x, out, planes, hin, win, hout, wout, sh, sw);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" void ipr_pdq_dct_matrix_host(float *d_host)
{
    const float scale = (float)sqrt(2.0 / 64.0);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 64; j++)
            d_host[i * 64 + j] = (float)(scale * cos((M_PI / 2.0 / 64.0) * (double)(i + 1) * (double)(2 * j + 1)));
}

extern "C" int ipr_pdq_hash_f32(const float *img, uint32_t *hash, float *coeffs, const float *dct,
                                int64_t batch, int height, int width, ipr_stream_t stream)
{
    IPR_REQUIRE(img && hash && dct, IPR_E_NULL);
    IPR_REQUIRE(batch > 0 && height > 0 && width > 0, IPR_E_SHAPE);
    IPR_REQUIRE(height <= PDQ_MAX && width <= PDQ_MAX && height >= 2 && width >= 2, IPR_E_UNSUPPORTED);
    IPR_REQUIRE(batch < (1LL << 31), IPR_E_UNSUPPORTED);
    const size_t smem = (size_t)2 * height * (width + 1) * sizeof(float);
    pdq_hash_kernel<<<(unsigned)batch, PDQ_THREADS, smem, ipr_cu(stream)>>>(img, hash, coeffs, dct, height, width);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_hash_pvalue(const uint32_t *hx, const uint32_t *hy, const float *ptable,
                               float *p, int32_t *r, int64_t batch, ipr_stream_t stream)
{
    IPR_REQUIRE(hx && hy && ptable && p, IPR_E_NULL);
    IPR_REQUIRE(batch > 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(hx) && ipr_aligned16(hy), IPR_E_ALIGN);
    hash_pvalue_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, ipr_cu(stream)>>>(hx, hy, ptable, p, r, batch);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
