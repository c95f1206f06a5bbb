// sn.cu -- spectral normalisation of the discriminator (torch.nn.utils.spectral_norm semantics,
// networks/sn_discriminator.py:1,9-21), batched over ALL layers of the network:
//   power iteration   v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v)      (4 launches, any #layers)
//   weight gradient   dW = (G - <G, W>/sigma * u v^T) / sigma                               (2 launches)
// The reference runs ~15 tiny PyTorch/cuBLAS kernels per layer and per forward for the former and an autograd
// chain for the latter.  W is the fp32 master weight viewed as (rows = out channels, cols = rest).
#include "ipr_common.cuh"

namespace {

struct SnTable {
    ipr_sn_layer_t layer[IPR_SN_MAX_LAYERS];
    int n_layers;
    int cta_begin[IPR_SN_MAX_LAYERS + 1];      // prefix sums of CTAs per layer for the current kernel
};

__device__ __forceinline__ int find_layer(const SnTable &t, int cta) {
    int l = 0;
    while (l + 1 < t.n_layers && cta >= t.cta_begin[l + 1]) l++;
    return l;
}

constexpr int ROW_CHUNK = 32;

// A1: partial_t[layer][row_chunk][col] = sum_{rows in chunk} W[row][col] * u[row]
__global__ void __launch_bounds__(256)
sn_wtu_kernel(const __grid_constant__ SnTable tab, float *__restrict__ scratch)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    const int l = find_layer(tab, blockIdx.x);
    const ipr_sn_layer_t L = tab.layer[l];
    const int local = blockIdx.x - tab.cta_begin[l];
    const int col_ctas = (L.cols + 255) / 256;
    const int rc = local / col_ctas, cc = local - rc * col_ctas;
    const int col = cc * 256 + threadIdx.x;
    if (col >= L.cols) return;
    const int r0 = rc * ROW_CHUNK, r1 = min(L.rows, r0 + ROW_CHUNK);
    float acc = 0.0f;
#pragma unroll 4
    for (int r = r0; r < r1; r++) acc += L.w[(size_t)r * L.cols + col] * L.u[r];
    scratch[L.scratch_off + (size_t)rc * L.cols + col] = acc;
}

// A2 (one CTA per layer): t = sum of partials; v = t / max(||t||, eps)
__global__ void __launch_bounds__(1024)
sn_v_kernel(const __grid_constant__ SnTable tab, const float *__restrict__ scratch, float eps)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const ipr_sn_layer_t L = tab.layer[blockIdx.x];
    const int chunks = (L.rows + ROW_CHUNK - 1) / ROW_CHUNK;
    float ss = 0.0f;
    for (int c = threadIdx.x; c < L.cols; c += blockDim.x) {
        float t = 0.0f;
        for (int k = 0; k < chunks; k++) t += scratch[L.scratch_off + (size_t)k * L.cols + c];
        L.v[c] = t;
        ss += t * t;
    }
    const float tot = ipr_block_sum(ss, red);
    const float inv = 1.0f / fmaxf(sqrtf(tot), eps);
    for (int c = threadIdx.x; c < L.cols; c += blockDim.x) {
        const float nv = L.v[c] * inv;
        L.v[c] = nv;
        if (L.v_snap) L.v_snap[c] = nv;                  // copy for this forward's backward (the buffer moves on)
    }
}

// B1: s[row] = W[row,:] . v     (one CTA per row, 128-bit loads where the row is 16-byte aligned)
__global__ void __launch_bounds__(256)
sn_wv_kernel(const __grid_constant__ SnTable tab, float *__restrict__ scratch)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const int l = find_layer(tab, blockIdx.x);
    const ipr_sn_layer_t L = tab.layer[l];
    const int row = blockIdx.x - tab.cta_begin[l];
    const float *w = L.w + (size_t)row * L.cols;
    float acc = 0.0f;
    if ((L.cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(L.v)) & 15) == 0) {
        const float4 *w4 = reinterpret_cast<const float4 *>(w), *v4 = reinterpret_cast<const float4 *>(L.v);
        for (int c = threadIdx.x; c < (L.cols >> 2); c += blockDim.x) {
            const float4 a = w4[c], b = v4[c];
            acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
    } else {
        for (int c = threadIdx.x; c < L.cols; c += blockDim.x) acc += w[c] * L.v[c];
    }
    const float tot = ipr_block_sum(acc, red);
    if (threadIdx.x == 0) scratch[L.scratch_off + row] = tot;
}

// B2 (one CTA per layer): update: u = s / max(||s||, eps), sigma = u . s ; no update: sigma = u_old . s
__global__ void __launch_bounds__(1024)
sn_u_sigma_kernel(const __grid_constant__ SnTable tab, const float *__restrict__ scratch, float eps, int update)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const ipr_sn_layer_t L = tab.layer[blockIdx.x];
    float acc = 0.0f;
    for (int r = threadIdx.x; r < L.rows; r += blockDim.x) {
        const float s = scratch[L.scratch_off + r];
        acc += update ? s * s : s * L.u[r];
    }
    const float tot = ipr_block_sum(acc, red);
    if (update) {
        const float inv = 1.0f / fmaxf(sqrtf(tot), eps);
        for (int r = threadIdx.x; r < L.rows; r += blockDim.x) {
            const float nu = scratch[L.scratch_off + r] * inv;
            L.u[r] = nu;
            if (L.u_snap) L.u_snap[r] = nu;
        }
        if (threadIdx.x == 0) *L.sigma = tot * inv;          // u . s = ||s||^2 / max(||s||, eps)
    } else {
        if (L.u_snap) for (int r = threadIdx.x; r < L.rows; r += blockDim.x) L.u_snap[r] = L.u[r];
        if (L.v_snap) for (int c = threadIdx.x; c < L.cols; c += blockDim.x) L.v_snap[c] = L.v[c];
        if (threadIdx.x == 0) *L.sigma = tot;
    }
}

// C1: per-CTA partial of <G, W> (each CTA covers DOT_SPAN elements with 128-bit loads);
// C2: G <- (G - (<G,W>/sigma) u v^T) / sigma   (in place, or added into grad_out)
constexpr int DOT_SPAN = 256 * 16;

__global__ void __launch_bounds__(256)
sn_dot_kernel(const __grid_constant__ SnTable tab, float *__restrict__ scratch)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const int l = find_layer(tab, blockIdx.x);
    const ipr_sn_layer_t L = tab.layer[l];
    const int part = blockIdx.x - tab.cta_begin[l];
    const long long n = (long long)L.rows * L.cols;
    const long long e0 = (long long)part * DOT_SPAN;
    float acc = 0.0f;
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(L.grad) | reinterpret_cast<uintptr_t>(L.w)) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long i = e0 + ((long long)j * 256 + threadIdx.x) * 4;
            if (i < n) {
                const float4 a = *reinterpret_cast<const float4 *>(L.grad + i), b = *reinterpret_cast<const float4 *>(L.w + i);
                acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
            }
        }
    } else {
        for (int j = 0; j < 16; j++) {
            const long long i = e0 + (long long)j * 256 + threadIdx.x;
            if (i < n) acc += L.grad[i] * L.w[i];
        }
    }
    const float tot = ipr_block_sum(acc, red);
    if (threadIdx.x == 0) scratch[L.scratch_off + part] = tot;
}

__global__ void __launch_bounds__(256)
sn_grad_kernel(const __grid_constant__ SnTable tab, const float *__restrict__ scratch)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const int l = find_layer(tab, blockIdx.x);
    const ipr_sn_layer_t L = tab.layer[l];
    const long long n = (long long)L.rows * L.cols;
    const int parts = (int)((n + DOT_SPAN - 1) / DOT_SPAN);
    float d = 0.0f;
    for (int k = threadIdx.x; k < parts; k += blockDim.x) d += scratch[L.scratch_off + k];
    const float dot = ipr_block_sum(d, red);                 // same fixed order in every CTA
    const float sigma = *L.sigma;
    const float coef = dot / sigma;
    const float inv = 1.0f / sigma;
    const long long base = (long long)(blockIdx.x - tab.cta_begin[l]) * DOT_SPAN;
    for (int j = 0; j < 16; j++) {
        const long long i = base + (long long)j * 256 + threadIdx.x;
        if (i >= n) break;
        const int r = (int)(i / L.cols), c = (int)(i - (long long)r * L.cols);
        const float out = (L.grad[i] - coef * L.u[r] * L.v[c]) * inv;
        if (L.grad_out) L.grad_out[i] += out; else L.grad[i] = out;
    }
}

int fill(SnTable &t, const ipr_sn_layer_t *layers, int n)
{
    IPR_REQUIRE(layers, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && n <= IPR_SN_MAX_LAYERS, IPR_E_SHAPE);
    for (int i = 0; i < n; i++) {
        IPR_REQUIRE(layers[i].w && layers[i].u && layers[i].v && layers[i].sigma, IPR_E_NULL);
        IPR_REQUIRE(layers[i].rows > 0 && layers[i].cols > 0, IPR_E_SHAPE);
        t.layer[i] = layers[i];
    }
    t.n_layers = n;
    return IPR_OK;
}

}  // namespace

extern "C" size_t ipr_sn_scratch_floats(int rows, int cols)
{
    const size_t a = (size_t)((rows + ROW_CHUNK - 1) / ROW_CHUNK) * cols;
    const size_t parts = ((size_t)rows * cols + DOT_SPAN - 1) / DOT_SPAN;
    const size_t b = (size_t)rows > parts ? (size_t)rows : parts;
    return (a > b ? a : b) + 32;
}

extern "C" int ipr_sn_power_iter_f32(const ipr_sn_layer_t *layers_host, int n_layers, int update, float eps,
                                     float *scratch, ipr_stream_t stream)
{
    SnTable t;
    int rc = fill(t, layers_host, n_layers);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(scratch, IPR_E_NULL);
    cudaStream_t st = ipr_cu(stream);
    if (update) {
        int total = 0;
        for (int i = 0; i < n_layers; i++) {
            t.cta_begin[i] = total;
            total += ((t.layer[i].rows + ROW_CHUNK - 1) / ROW_CHUNK) * ((t.layer[i].cols + 255) / 256);
        }
        t.cta_begin[n_layers] = total;
        IPR_LAUNCH_PDL((sn_wtu_kernel), total, 256, 0, st, t, scratch);
        IPR_LAUNCH_CHECK();
        IPR_LAUNCH_PDL((sn_v_kernel), n_layers, 1024, 0, st, t, scratch, eps);
        IPR_LAUNCH_CHECK();
    }
    int total = 0;
    for (int i = 0; i < n_layers; i++) { t.cta_begin[i] = total; total += t.layer[i].rows; }
    t.cta_begin[n_layers] = total;
    IPR_LAUNCH_PDL((sn_wv_kernel), total, 256, 0, st, t, scratch);
    IPR_LAUNCH_CHECK();
    IPR_LAUNCH_PDL((sn_u_sigma_kernel), n_layers, 1024, 0, st, t, scratch, eps, update);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_sn_weight_grad_f32(const ipr_sn_layer_t *layers_host, int n_layers, float *scratch,
                                      ipr_stream_t stream)
{
    SnTable t;
    int rc = fill(t, layers_host, n_layers);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(scratch, IPR_E_NULL);
    for (int i = 0; i < n_layers; i++) IPR_REQUIRE(t.layer[i].grad, IPR_E_NULL);
    cudaStream_t st = ipr_cu(stream);
    int total = 0;
    for (int i = 0; i < n_layers; i++) {
        t.cta_begin[i] = total;
        total += (int)(((long long)t.layer[i].rows * t.layer[i].cols + DOT_SPAN - 1) / DOT_SPAN);
    }
    t.cta_begin[n_layers] = total;
    IPR_LAUNCH_PDL((sn_dot_kernel), total, 256, 0, st, t, scratch);
    IPR_LAUNCH_CHECK();
    IPR_LAUNCH_PDL((sn_grad_kernel), total, 256, 0, st, t, scratch);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
