// layers.cu -- the memory-bound layer kernels around the tensor-core GEMMs for the SRGAN / CycleGAN network
// families (networks/sr_resnet.py:3-44, discriminator_96.py:3-35, resnet_generator.py:3-59,
// conv_discriminator.py:3-21): every convolution there (k in {1,3,4,6,7,9}, stride 1/2, zero or reflection padding,
// transposed with output padding) is  patch matrix (this file) x packed weight (tcgen05 GEMM, gemm_tc.cu), and its
// data gradient is  dY x W^T (tcgen05 GEMM) folded back by the adjoint gather (this file).  Plus: BatchNorm /
// InstanceNorm (+ ReLU / LeakyReLU / PReLU / Tanh) forward and backward with the white-box sign-loss gradient
// (tools/sign_model.py:42-49) added inside the backward, PixelShuffle, layout changes at the module boundary.
// All activations NHWC bf16, 8 channels (16 bytes) per thread; statistics and coefficients fp32.
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

inline unsigned grid_1d(long long items, int threads, int waves = 16) {
    long long blocks = (items + threads - 1) / threads;
    const long long cap = (long long)ipr_sm_count() * waves;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

// ------------------------------------------------------------------------------------ patch matrix and its adjoint
struct ConvGeom {
    int n, h, w, c;          // input tensor (NHWC), c % 8 == 0
    int oh, ow;              // output grid
    int k, stride, pad;      // square kernel; pad is the zero / reflection border of the (virtual) input
    int up;                  // 1, or the zero-insertion factor of a transposed convolution run as a direct one
    int reflect;             // border mode (up must be 1)
    int kp;                  // columns per patch row (>= k*k*c, multiple of 8); columns past k*k*c are zero
};

// virtual coordinate v (in the zero-inserted grid of extent (size-1)*up + 1) -> input index, or -1 for a zero
__device__ __forceinline__ int src_index(int v, int size, int up, int reflect) {
    if (reflect) {
        if (v < 0) v = -v;
        if (v > size - 1) v = 2 * (size - 1) - v;
        return v;
    }
    if (v < 0 || v > (size - 1) * up) return -1;
    if (up == 1) return v;
    return (v % up == 0) ? v / up : -1;
}

// col[(n, oy, ox)][(ky*k + kx)*c + ch] = x[n, src(oy*stride + ky - pad), src(ox*stride + kx - pad), ch]
__global__ void __launch_bounds__(256)
im2col_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ col, const ConvGeom g, long long total_vec)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int cv = g.c >> 3, kpv = g.kp >> 3, taps = g.k * g.k;
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride_t) {
        const int j = (int)(i % kpv);
        const long long m = i / kpv;
        const int tap = j / cv, c8 = j - tap * cv;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (tap < taps) {
            const int ox = (int)(m % g.ow);
            const long long r = m / g.ow;
            const int oy = (int)(r % g.oh);
            const long long n = r / g.oh;
            const int ky = tap / g.k, kx = tap - ky * g.k;
            const int iy = src_index(oy * g.stride + ky - g.pad, g.h, g.up, g.reflect);
            const int ix = src_index(ox * g.stride + kx - g.pad, g.w, g.up, g.reflect);
            if (iy >= 0 && ix >= 0) v = __ldg(x + ((n * g.h + iy) * g.w + ix) * cv + c8);
        }
        col[i] = v;
    }
}

// dx[n, iy, ix, ch] = (addend) + sum over every (oy, ky), (ox, kx) whose source is (iy, ix) of dcol[(n,oy,ox)][(ky,kx,ch)]
__global__ void __launch_bounds__(256)
col2im_kernel(const uint4 *__restrict__ dcol, uint4 *__restrict__ dx, const uint4 *__restrict__ addend, const ConvGeom g,
              long long total_vec)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int cv = g.c >> 3, kpv = g.kp >> 3;
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride_t) {
        const int c8 = (int)(i % cv);
        const long long p = i / cv;
        const int ix = (int)(p % g.w);
        const long long r = p / g.w;
        const int iy = (int)(r % g.h);
        const long long n = r / g.h;
        // virtual coordinates that read input row iy / column ix (itself; its mirror images under reflection)
        int vy[3], vx[3], ny = 0, nx = 0;
        vy[ny++] = iy * g.up; vx[nx++] = ix * g.up;
        if (g.reflect) {
            if (iy >= 1 && iy <= g.pad) vy[ny++] = -iy;
            if (iy <= g.h - 2 && iy >= g.h - 1 - g.pad) vy[ny++] = 2 * (g.h - 1) - iy;
            if (ix >= 1 && ix <= g.pad) vx[nx++] = -ix;
            if (ix <= g.w - 2 && ix >= g.w - 1 - g.pad) vx[nx++] = 2 * (g.w - 1) - ix;
        }
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = 0.0f;
        if (addend) unpack8(__ldg(addend + i), acc);
        for (int a = 0; a < ny; a++)
            for (int ky = 0; ky < g.k; ky++) {
                const int ty = vy[a] + g.pad - ky;
                if (ty < 0 || ty % g.stride != 0) continue;
                const int oy = ty / g.stride;
                if (oy >= g.oh) continue;
                for (int b = 0; b < nx; b++)
                    for (int kx = 0; kx < g.k; kx++) {
                        const int tx = vx[b] + g.pad - kx;
                        if (tx < 0 || tx % g.stride != 0) continue;
                        const int ox = tx / g.stride;
                        if (ox >= g.ow) continue;
                        float f[8];
                        unpack8(__ldg(dcol + ((n * g.oh + oy) * g.ow + ox) * kpv + (ky * g.k + kx) * cv + c8), f);
#pragma unroll
                        for (int e = 0; e < 8; e++) acc[e] += f[e];
                    }
            }
        dx[i] = pack8(acc);
    }
}

int check_geom(const ConvGeom &g)
{
    IPR_REQUIRE(g.n > 0 && g.h > 0 && g.w > 0 && g.c > 0 && g.c % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(g.oh > 0 && g.ow > 0 && g.k > 0 && g.stride > 0 && g.pad >= 0 && g.up >= 1, IPR_E_SHAPE);
    IPR_REQUIRE(g.kp % 8 == 0 && g.kp >= g.k * g.k * g.c, IPR_E_SHAPE);
    if (g.reflect) {
        // one reflection only: the border is narrower than the image and no output reads past the mirrored border
        IPR_REQUIRE(g.up == 1 && g.pad < g.h && g.pad < g.w, IPR_E_UNSUPPORTED);
        IPR_REQUIRE((g.oh - 1) * g.stride + g.k - 1 - g.pad <= g.h - 1 + g.pad, IPR_E_SHAPE);
        IPR_REQUIRE((g.ow - 1) * g.stride + g.k - 1 - g.pad <= g.w - 1 + g.pad, IPR_E_SHAPE);
    }
    return IPR_OK;
}

// ------------------------------------------------------------------------------------ layout at the module boundary
// NCHW fp32 -> NHWC bf16 with the channel count padded to cp (zeros); optional per-element factor (1 - t^2) of a
// Tanh output t (the backward of a final Tanh fused into the gradient's layout change)
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float *__restrict__ x, const float *__restrict__ tanh_out, uint4 *__restrict__ y, long long pixels,
                    int c, int hw, int cp)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int cpv = cp >> 3;
    const long long total = pixels * cpv;
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride_t) {
        const int c8 = (int)(i % cpv);
        const long long p = i / cpv;
        const long long n = p / hw;
        const int s = (int)(p - n * hw);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int ch = c8 * 8 + e;
            float v = 0.0f;
            if (ch < c) {
                const size_t idx = ((size_t)n * c + ch) * hw + s;
                v = __ldg(x + idx);
                if (tanh_out) { const float t = __ldg(tanh_out + idx); v *= (1.0f - t * t); }
            }
            f[e] = v;
        }
        y[i] = pack8(f);
    }
}

// fp32 GEMM result [pixel][ld] (+ bias[ch]) -> NCHW fp32, first c channels, optional Tanh
__global__ void __launch_bounds__(256)
finish_nchw_kernel(const float *__restrict__ t, const float *__restrict__ bias, float *__restrict__ out, long long pixels,
                   int c, int hw, int ld, int tanh_out)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const long long total = pixels * c;
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride_t) {
        // consecutive threads -> consecutive pixels of one channel plane (coalesced NCHW writes)
        const long long n = i / ((long long)c * hw);
        const long long rem = i - n * (long long)c * hw;
        const int ch = (int)(rem / hw);
        const int s = (int)(rem - (long long)ch * hw);
        float v = __ldg(t + ((size_t)n * hw + s) * ld + ch) + (bias ? __ldg(bias + ch) : 0.0f);
        if (tanh_out) v = tanhf(v);
        out[i] = v;
    }
}

// PixelShuffle(2) on NHWC bf16: y[n, 2h+i, 2w+j, c] = x[n, h, w, c*4 + i*2 + j]   (inverse != 0: the other way round)
__global__ void __launch_bounds__(256)
pixel_shuffle2_kernel(const __nv_bfloat16 *__restrict__ x, __nv_bfloat16 *__restrict__ y, long long n_imgs, int h, int w,
                      int c_out, int inverse)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const long long total = n_imgs * h * w * c_out * 4;
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride_t) {
        // e enumerates the LARGE-grid tensor [n][2h][2w][c_out]
        const int c = (int)(e % c_out);
        long long r = e / c_out;
        const int X = (int)(r % (2 * w)); r /= (2 * w);
        const int Y = (int)(r % (2 * h));
        const long long n = r / (2 * h);
        const long long small = ((n * h + (Y >> 1)) * w + (X >> 1)) * (long long)(c_out * 4) + c * 4 + (Y & 1) * 2 + (X & 1);
        if (inverse) y[small] = x[e]; else y[e] = x[small];
    }
}

__global__ void __launch_bounds__(256)
add_bf16_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, uint4 *__restrict__ out, long long n_vec)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride_t) {
        float fa[8], fb[8];
        unpack8(__ldg(a + i), fa);
        unpack8(__ldg(b + i), fb);
#pragma unroll
        for (int e = 0; e < 8; e++) fa[e] += fb[e];
        out[i] = pack8(fa);
    }
}

// ------------------------------------------------------------------------------------ normalisation + activation
// groups = 1 (BatchNorm: statistics over all rows) or the batch size (InstanceNorm: per image); rows = pixels per group.
constexpr int ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_PRELU = 3, ACT_TANH = 4;

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
    if (act == ACT_RELU) return fmaxf(v, 0.0f);
    if (act == ACT_LRELU || act == ACT_PRELU) return v > 0.0f ? v : v * slope;
    if (act == ACT_TANH) return tanhf(v);
    return v;
}
__device__ __forceinline__ float act_bwd(float pre, int act, float slope) {
    if (act == ACT_RELU) return pre > 0.0f ? 1.0f : 0.0f;
    if (act == ACT_LRELU || act == ACT_PRELU) return pre > 0.0f ? 1.0f : slope;
    if (act == ACT_TANH) { const float t = tanhf(pre); return 1.0f - t * t; }
    return 1.0f;
}

// partial[(g * slabs + s)][2][C] = sum / sum of squares over the rows of slab s of group g
__global__ void __launch_bounds__(256)
norm_stats_kernel(const uint4 *__restrict__ x, float *__restrict__ partial, long long rows, int c_vec, int slabs)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    extern __shared__ float sm[];                    // [ry][2][C]
    const int C = c_vec * 8;
    const int ry_n = blockDim.x / c_vec;
    const int cv = threadIdx.x % c_vec, ry = threadIdx.x / c_vec;
    const int g = blockIdx.x / slabs, s = blockIdx.x - g * slabs;
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s1[k] = s2[k] = 0.0f;
    if (ry < ry_n) {
        const uint4 *base = x + (size_t)g * rows * c_vec;
        for (long long r = (long long)s * ry_n + ry; r < rows; r += (long long)slabs * ry_n) {
            float f[8];
            unpack8(__ldg(base + r * c_vec + cv), f);
#pragma unroll
            for (int k = 0; k < 8; k++) { s1[k] += f[k]; s2[k] += f[k] * f[k]; }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) { sm[(ry * 2) * C + cv * 8 + k] = s1[k]; sm[(ry * 2 + 1) * C + cv * 8 + k] = s2[k]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float acc = 0.0f;
        for (int y = 0; y < ry_n; y++) acc += sm[y * 2 * C + i];
        partial[(size_t)blockIdx.x * 2 * C + i] = acc;
    }
}

// scale / shift / mean / rstd per (group, channel) from the slab partials; BatchNorm running statistics (groups == 1).
// CTA = 32 (group, channel) columns x FIN_LANES slab lanes: every thread adds every FIN_LANES-th slab (double, fixed
// order), the lanes are combined in order through shared memory.  (The first version walked all <= 256 slabs with one
// thread per column: 47 us of serial load latency per launch, 87 launches in one SRGAN step.)
constexpr int FIN_LANES = 16;
__global__ void __launch_bounds__(32 * FIN_LANES)
norm_finalize_kernel(const float *__restrict__ partial, int groups, int slabs, int C, double count, float eps, float momentum,
                     const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ running_mean,
                     float *__restrict__ running_var, long long *__restrict__ num_batches, float *__restrict__ scale,
                     float *__restrict__ shift, float *__restrict__ mean_out, float *__restrict__ rstd_out)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ double sm[2][FIN_LANES][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches) *num_batches += 1;
    const bool live = i < groups * C;
    const int g = live ? i / C : 0, c = live ? i - g * C : 0;
    double a = 0.0, b = 0.0;
    if (live)
        for (int s = ty; s < slabs; s += FIN_LANES) {
            a += (double)partial[((size_t)(g * slabs + s) * 2) * C + c];
            b += (double)partial[((size_t)(g * slabs + s) * 2 + 1) * C + c];
        }
    sm[0][ty][tx] = a; sm[1][ty][tx] = b;
    __syncthreads();
    if (ty != 0 || !live) return;
#pragma unroll
    for (int y = 1; y < FIN_LANES; y++) { a += sm[0][y][tx]; b += sm[1][y][tx]; }
    const double mean = a / count;
    double var = b / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = gamma ? gamma[c] : 1.0f, be = beta ? beta[c] : 0.0f;
    scale[i] = ga * rstd;
    shift[i] = be - (float)mean * ga * rstd;
    mean_out[i] = (float)mean;
    rstd_out[i] = rstd;
    if (running_mean && g == 0) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// y = act(x * scale[g][c] + shift[g][c]) (+ residual);  scale == nullptr: y = act(x) (+ residual)
__global__ void __launch_bounds__(256)
norm_act_fwd_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, const float *__restrict__ scale,
                    const float *__restrict__ shift, const uint4 *__restrict__ residual, long long rows, int c_vec,
                    long long n_vec, int act, float slope, const float *__restrict__ slope_ptr)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int C = c_vec * 8;
    if (slope_ptr) slope = __ldg(slope_ptr);
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride_t) {
        const int c0 = (int)(i % c_vec) * 8;
        const long long g = (i / c_vec) / rows;
        float f[8];
        unpack8(__ldg(x + i), f);
        if (scale) {
            const float *sc = scale + g * C + c0, *sh = shift + g * C + c0;
#pragma unroll
            for (int k = 0; k < 8; k++) f[k] = fmaf(f[k], __ldg(sc + k), __ldg(sh + k));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = act_fwd(f[k], act, slope);
        if (residual) {
            float r[8];
            unpack8(__ldg(residual + i), r);
#pragma unroll
            for (int k = 0; k < 8; k++) f[k] += r[k];
        }
        y[i] = pack8(f);
    }
}

// backward pass 1: partial[(g*slabs+s)][3][C] = sum gin, sum gin*xhat, sum dy*pre*[pre<0] (PReLU slope gradient)
// with pre = x*scale + shift (or x), gin = dy * act'(pre), xhat = (x - mean) * rstd
__global__ void __launch_bounds__(256)
norm_bwd_reduce_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x, const float *__restrict__ scale,
                       const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd,
                       float *__restrict__ partial, long long rows, int c_vec, int slabs, int act, float slope,
                       const float *__restrict__ slope_ptr)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    extern __shared__ float sm[];                    // [ry][3][C]
    const int C = c_vec * 8;
    if (slope_ptr) slope = __ldg(slope_ptr);
    const int ry_n = blockDim.x / c_vec;
    const int cv = threadIdx.x % c_vec, ry = threadIdx.x / c_vec;
    const int g = blockIdx.x / slabs, s = blockIdx.x - g * slabs;
    float a[8], b[8], p[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = b[k] = p[k] = 0.0f;
    if (ry < ry_n) {
        float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ci = g * C + cv * 8 + k;
            sc[k] = scale ? scale[ci] : 1.0f; sh[k] = scale ? shift[ci] : 0.0f;
            mu[k] = scale ? mean[ci] : 0.0f;  rs[k] = scale ? rstd[ci] : 1.0f;
        }
        const size_t base = (size_t)g * rows * c_vec;
        for (long long r = (long long)s * ry_n + ry; r < rows; r += (long long)slabs * ry_n) {
            float d[8], xv[8];
            unpack8(__ldg(dy + base + r * c_vec + cv), d);
            unpack8(__ldg(x + base + r * c_vec + cv), xv);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float pre = fmaf(xv[k], sc[k], sh[k]);
                const float gin = d[k] * act_bwd(pre, act, slope);
                a[k] += gin;
                b[k] += gin * (xv[k] - mu[k]) * rs[k];
                if (act == ACT_PRELU && pre <= 0.0f) p[k] += d[k] * pre;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            sm[(ry * 3) * C + cv * 8 + k] = a[k]; sm[(ry * 3 + 1) * C + cv * 8 + k] = b[k]; sm[(ry * 3 + 2) * C + cv * 8 + k] = p[k];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
        float acc = 0.0f;
        for (int y = 0; y < ry_n; y++) acc += sm[y * 3 * C + i];
        partial[(size_t)blockIdx.x * 3 * C + i] = acc;
    }
}

// backward pass 2: per (group, channel) coefficients  dx = A*gin + B*x + D;  dgamma / dbeta summed over the groups
// (+ the sign-loss gradient); PReLU slope gradient summed over everything.
// CTAs 0 .. ceil(C/32)-1 (has_norm): 32 channels x FIN_LANES slab lanes each, groups walked in order; the LAST CTA (dslope)
// adds up the slope-gradient plane in a fixed order.  (The first version was ONE CTA with one thread per channel walking
// groups x slabs rows serially: 103 us per launch, a quarter of the SRGAN step.)
__global__ void __launch_bounds__(32 * FIN_LANES)
norm_bwd_finalize_kernel(const float *__restrict__ partial, int groups, int slabs, int C, double count,
                         const float *__restrict__ gamma, const float *__restrict__ mean, const float *__restrict__ rstd,
                         float *__restrict__ dgamma, float *__restrict__ dbeta, int accumulate,
                         const float *__restrict__ sign, float gamma0, float sign_scale, float *__restrict__ dslope,
                         int has_norm, float *__restrict__ coef, int norm_ctas)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ double sm[2][FIN_LANES][33];
    __shared__ float red[32];
    if ((int)blockIdx.x >= norm_ctas) {              // the slope-gradient CTA
        double acc = 0.0;
        const long long total = (long long)groups * slabs * C;
        for (long long e = threadIdx.x; e < total; e += blockDim.x) {
            const long long row = e / C;
            acc += (double)partial[(size_t)row * 3 * C + 2 * C + (int)(e - row * C)];
        }
        const float tot = ipr_block_sum((float)acc, red);
        if (threadIdx.x == 0) *dslope = accumulate ? *dslope + tot : tot;
        return;
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const bool live = c < C;
    double dg = 0.0, db = 0.0;
    for (int g = 0; g < groups; g++) {
        double sg = 0.0, sx = 0.0;
        if (live)
            for (int s = ty; s < slabs; s += FIN_LANES) {
                const float *row = partial + (size_t)(g * slabs + s) * 3 * C;
                sg += (double)row[c]; sx += (double)row[C + c];
            }
        __syncthreads();                             // previous group's lanes have been read
        sm[0][ty][tx] = sg; sm[1][ty][tx] = sx;
        __syncthreads();
        if (ty == 0 && live) {
#pragma unroll
            for (int y = 1; y < FIN_LANES; y++) { sg += sm[0][y][tx]; sx += sm[1][y][tx]; }
            const int ci = g * C + c;
            const float ga = gamma ? gamma[c] : 1.0f, rs = rstd[ci], mu = mean[ci];
            const float A = ga * rs;
            const float B = -ga * rs * rs * (float)(sx / count);
            coef[ci] = A;
            coef[groups * C + ci] = B;
            coef[2 * groups * C + ci] = -A * (float)(sg / count) - B * mu;
            dg += sx; db += sg;
        }
    }
    if (ty == 0 && live && dgamma) {
        float dgf = (float)dg;
        if (sign) {                                  // d/dgamma of mean_c relu(gamma0 - gamma*sign)
            const float sv = sign[c];
            if (gamma0 - gamma[c] * sv > 0.0f) dgf += -sv * sign_scale / (float)C;
        }
        dgamma[c] = accumulate ? dgamma[c] + dgf : dgf;
        dbeta[c] = accumulate ? dbeta[c] + (float)db : (float)db;
    }
}

// backward pass 3: dx = A*gin + B*x + D  (no norm: dx = gin)
__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x, const float *__restrict__ scale,
                      const float *__restrict__ shift, const float *__restrict__ coef, uint4 *__restrict__ dx, long long rows,
                      int c_vec, int groups, long long n_vec, int act, float slope, const float *__restrict__ slope_ptr)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int C = c_vec * 8;
    if (slope_ptr) slope = __ldg(slope_ptr);
    const long long stride_t = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride_t) {
        const int c0 = (int)(i % c_vec) * 8;
        const long long g = (i / c_vec) / rows;
        float d[8], xv[8], o[8];
        unpack8(__ldg(dy + i), d);
        unpack8(__ldg(x + i), xv);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int ci = (int)g * C + c0 + k;
            const float pre = scale ? fmaf(xv[k], __ldg(scale + ci), __ldg(shift + ci)) : xv[k];
            const float gin = d[k] * act_bwd(pre, act, slope);
            o[k] = scale ? __ldg(coef + ci) * gin + __ldg(coef + groups * C + ci) * xv[k] + __ldg(coef + 2 * groups * C + ci) : gin;
        }
        dx[i] = pack8(o);
    }
}

inline int norm_slabs(long long rows, int groups) {
    long long s = rows / 32;
    if (s < 1) s = 1;
    const long long cap = (long long)(4 * ipr_sm_count()) / groups;
    if (s > cap) s = cap;
    if (s > 256) s = 256;
    return (int)(s < 1 ? 1 : s);
}

}  // namespace

extern "C" int ipr_im2col_nhwc_bf16(const void *x, void *col, int n, int h, int w, int c, int oh, int ow, int k, int stride,
                                    int pad, int up, int reflect, int kp, ipr_stream_t stream)
{
    IPR_REQUIRE(x && col, IPR_E_NULL);
    ConvGeom g = {n, h, w, c, oh, ow, k, stride, pad, up, reflect, kp};
    int rc = check_geom(g);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(ipr_aligned16(x) && ipr_aligned16(col), IPR_E_ALIGN);
    const long long total = (long long)n * oh * ow * (kp / 8);
    IPR_LAUNCH_PDL((im2col_kernel), grid_1d(total, 256), 256, 0, ipr_cu(stream), (const uint4 *)x, (uint4 *)col, g, total);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_col2im_nhwc_bf16(const void *dcol, void *dx, const void *addend, int n, int h, int w, int c, int oh,
                                    int ow, int k, int stride, int pad, int up, int reflect, int kp, ipr_stream_t stream)
{
    IPR_REQUIRE(dcol && dx, IPR_E_NULL);
    ConvGeom g = {n, h, w, c, oh, ow, k, stride, pad, up, reflect, kp};
    int rc = check_geom(g);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(ipr_aligned16(dcol) && ipr_aligned16(dx) && ipr_aligned16(addend), IPR_E_ALIGN);
    const long long total = (long long)n * h * w * (c / 8);
    IPR_LAUNCH_PDL((col2im_kernel), grid_1d(total, 256), 256, 0, ipr_cu(stream), (const uint4 *)dcol, (uint4 *)dx,
                   (const uint4 *)addend, g, total);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_nchw_to_nhwc_bf16(const float *x, const float *tanh_out, void *y, int64_t n, int c, int h, int w, int cp,
                                     ipr_stream_t stream)
{
    IPR_REQUIRE(x && y, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && cp >= c && cp % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(y), IPR_E_ALIGN);
    const long long pixels = (long long)n * h * w;
    IPR_LAUNCH_PDL((nchw_to_nhwc_kernel), grid_1d(pixels * (cp / 8), 256), 256, 0, ipr_cu(stream), x, tanh_out, (uint4 *)y, pixels,
                   c, h * w, cp);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_finish_nchw_f32(const float *t, const float *bias, float *out, int64_t n, int c, int h, int w, int ld,
                                   int tanh_out, ipr_stream_t stream)
{
    IPR_REQUIRE(t && out, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && ld >= c, IPR_E_SHAPE);
    const long long pixels = (long long)n * h * w;
    IPR_LAUNCH_PDL((finish_nchw_kernel), grid_1d(pixels * c, 256), 256, 0, ipr_cu(stream), t, bias, out, pixels, c, h * w, ld,
                   tanh_out);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_pixel_shuffle2_nhwc_bf16(const void *x, void *y, int64_t n, int h, int w, int c_out, int inverse,
                                            ipr_stream_t stream)
{
    IPR_REQUIRE(x && y, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && h > 0 && w > 0 && c_out > 0, IPR_E_SHAPE);
    const long long total = (long long)n * h * w * c_out * 4;
    IPR_LAUNCH_PDL((pixel_shuffle2_kernel), grid_1d(total, 256), 256, 0, ipr_cu(stream), (const __nv_bfloat16 *)x,
                   (__nv_bfloat16 *)y, (long long)n, h, w, c_out, inverse);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_add_bf16(const void *a, const void *b, void *out, int64_t n, ipr_stream_t stream)
{
    IPR_REQUIRE(a && b && out, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && n % 8 == 0, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(a) && ipr_aligned16(b) && ipr_aligned16(out), IPR_E_ALIGN);
    IPR_LAUNCH_PDL((add_bf16_kernel), grid_1d(n / 8, 256), 256, 0, ipr_cu(stream), (const uint4 *)a, (const uint4 *)b, (uint4 *)out,
                   (long long)(n / 8));
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" size_t ipr_norm_workspace_bytes(int groups, int channels)
{
    // slab partials [groups*slabs][3][C] (slabs <= 256, groups*slabs <= 4*SMs + groups) + coefficients [3][groups][C]
    const size_t rows = (size_t)4 * ipr_sm_count() + (size_t)groups * 2;
    return (rows * 3 * channels + (size_t)3 * groups * channels) * sizeof(float);
}

extern "C" int ipr_norm_fwd_bf16(const void *x, void *y, const void *residual, int groups, int64_t rows, int channels,
                                 int has_norm, float eps, float momentum, const float *gamma, const float *beta,
                                 float *running_mean, float *running_var, int64_t *num_batches_tracked,
                                 float *scale, float *shift, float *mean, float *rstd, int act, float slope,
                                 const float *slope_ptr, void *workspace, size_t workspace_bytes, ipr_stream_t stream)
{
    IPR_REQUIRE(x && y, IPR_E_NULL);
    IPR_REQUIRE(groups > 0 && rows > 0 && channels > 0 && channels % 8 == 0 && channels <= 2048, IPR_E_SHAPE);
    IPR_REQUIRE(act >= 0 && act <= 4, IPR_E_SHAPE);
    IPR_REQUIRE(ipr_aligned16(x) && ipr_aligned16(y) && ipr_aligned16(residual), IPR_E_ALIGN);
    cudaStream_t st = ipr_cu(stream);
    const int c_vec = channels / 8;
    if (has_norm == 2) {                       // given statistics (eval-mode BatchNorm): the caller filled scale / shift
        IPR_REQUIRE(scale && shift && groups == 1, IPR_E_NULL);
    } else if (has_norm) {
        IPR_REQUIRE(scale && shift && mean && rstd && workspace, IPR_E_NULL);
        IPR_REQUIRE(workspace_bytes >= ipr_norm_workspace_bytes(groups, channels), IPR_E_WORKSPACE);
        IPR_REQUIRE(!running_mean || groups == 1, IPR_E_UNSUPPORTED);
        const int slabs = norm_slabs(rows, groups);
        const int threads = c_vec <= 256 ? 256 : c_vec;                 // >= one row lane per CTA
        IPR_REQUIRE(threads <= 1024, IPR_E_UNSUPPORTED);
        const int ry_n = threads / c_vec;
        const size_t smem = (size_t)ry_n * 2 * channels * sizeof(float);
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(norm_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr = true;
        }
        IPR_REQUIRE(smem <= 96 * 1024, IPR_E_UNSUPPORTED);
        float *partial = (float *)workspace;
        IPR_LAUNCH_PDL((norm_stats_kernel), groups * slabs, threads, smem, st, (const uint4 *)x, partial, (long long)rows, c_vec, slabs);
        IPR_LAUNCH_CHECK();
        IPR_LAUNCH_PDL((norm_finalize_kernel), (groups * channels + 31) / 32, 32 * FIN_LANES, 0, st, partial, groups, slabs, channels,
                       (double)rows, eps, momentum, gamma, beta, running_mean, running_var, (long long *)num_batches_tracked,
                       scale, shift, mean, rstd);
        IPR_LAUNCH_CHECK();
    }
    const long long n_vec = (long long)groups * rows * c_vec;
    IPR_LAUNCH_PDL((norm_act_fwd_kernel), grid_1d(n_vec, 256), 256, 0, st, (const uint4 *)x, (uint4 *)y,
                   has_norm ? scale : (const float *)nullptr, has_norm ? shift : (const float *)nullptr,
                   (const uint4 *)residual, (long long)rows, c_vec, n_vec, act, slope, slope_ptr);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_norm_bwd_bf16(const void *dy, const void *x, void *dx, int groups, int64_t rows, int channels,
                                 int has_norm, const float *gamma, const float *scale, const float *shift,
                                 const float *mean, const float *rstd, float *dgamma, float *dbeta, int accumulate,
                                 const float *sign, float gamma0, float sign_scale, int act, float slope,
                                 const float *slope_ptr, float *dslope, void *workspace, size_t workspace_bytes,
                                 ipr_stream_t stream)
{
    IPR_REQUIRE(dy && x && dx && workspace, IPR_E_NULL);
    IPR_REQUIRE(groups > 0 && rows > 0 && channels > 0 && channels % 8 == 0 && channels <= 2048, IPR_E_SHAPE);
    IPR_REQUIRE(act >= 0 && act <= 4, IPR_E_SHAPE);
    IPR_REQUIRE(!has_norm || (scale && shift && mean && rstd), IPR_E_NULL);
    IPR_REQUIRE(workspace_bytes >= ipr_norm_workspace_bytes(groups, channels), IPR_E_WORKSPACE);
    IPR_REQUIRE(ipr_aligned16(dy) && ipr_aligned16(x) && ipr_aligned16(dx), IPR_E_ALIGN);
    cudaStream_t st = ipr_cu(stream);
    const int c_vec = channels / 8;
    const long long n_vec = (long long)groups * rows * c_vec;
    const bool need_reduce = has_norm || (dslope != nullptr && act == ACT_PRELU);
    float *partial = (float *)workspace;
    const int slabs = norm_slabs(rows, groups);
    float *coef = partial + (size_t)groups * slabs * 3 * channels;
    if (need_reduce) {
        const int threads = c_vec <= 256 ? 256 : c_vec;
        IPR_REQUIRE(threads <= 1024, IPR_E_UNSUPPORTED);
        const int ry_n = threads / c_vec;
        const size_t smem = (size_t)ry_n * 3 * channels * sizeof(float);
        IPR_REQUIRE(smem <= 96 * 1024, IPR_E_UNSUPPORTED);
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(norm_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr = true;
        }
        IPR_LAUNCH_PDL((norm_bwd_reduce_kernel), groups * slabs, threads, smem, st, (const uint4 *)dy, (const uint4 *)x,
                       has_norm ? scale : (const float *)nullptr, shift, mean, rstd, partial, (long long)rows, c_vec, slabs, act,
                       slope, slope_ptr);
        IPR_LAUNCH_CHECK();
        const int norm_ctas = has_norm ? (channels + 31) / 32 : 0;
        const bool slope_cta = dslope != nullptr;
        IPR_LAUNCH_PDL((norm_bwd_finalize_kernel), norm_ctas + (slope_cta ? 1 : 0), 32 * FIN_LANES, 0, st, (const float *)partial,
                       groups, slabs, channels, (double)rows, gamma, mean, rstd, dgamma, dbeta, accumulate, sign, gamma0,
                       sign_scale, dslope, has_norm, coef, norm_ctas);
        IPR_LAUNCH_CHECK();
    }
    IPR_LAUNCH_PDL((norm_bwd_apply_kernel), grid_1d(n_vec, 256), 256, 0, st, (const uint4 *)dy, (const uint4 *)x,
                   has_norm ? scale : (const float *)nullptr, shift, (const float *)coef, (uint4 *)dx, (long long)rows, c_vec,
                   groups, n_vec, act, slope, slope_ptr);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
