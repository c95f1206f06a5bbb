// step_misc.cu -- the scalar ends of the protected training step, one launch each, so that the captured step holds
// no framework (ATen) arithmetic at all:
//   * hinge discriminator loss + its logit gradients          models/dcgan.py:31-35  (relu(1-r).mean + relu(1+f).mean)
//   * generator adversarial loss + its logit gradient         models/dcgan.py:37-40  (-mean D(G(z)))
//   * latent draws on the device (Philox4x32-10 + Box-Muller) experiments/image_generation.py:94 (z = randn(B,128),
//     made on the CPU and copied every step by the reference)
// Losses land in caller-provided slots (the metrics board that rides in the gradient all-reduce buffer), gradients
// in caller-provided vectors that the step feeds straight into the networks' backward passes.
#include "ipr_common.cuh"

namespace {

// losses[0] = LossD = LossR + LossF, [1] = LossR, [2] = LossF;  d_real / d_fake = d LossD / d logits
__global__ void __launch_bounds__(1024)
hinge_d_kernel(const float *__restrict__ real, const float *__restrict__ fake, int batch, float loss_scale,
               float *__restrict__ losses, float *__restrict__ d_real, float *__restrict__ d_fake)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const float inv = 1.0f / (float)batch;
    float sr = 0.0f, sf = 0.0f;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
        const float mr = 1.0f - real[i], mf = 1.0f + fake[i];
        sr += mr > 0.0f ? mr : 0.0f;
        sf += mf > 0.0f ? mf : 0.0f;
        if (d_real) d_real[i] = mr > 0.0f ? -inv : 0.0f;
        if (d_fake) d_fake[i] = mf > 0.0f ? inv : 0.0f;
    }
    const float tr = ipr_block_sum(sr, red);
    const float tf = ipr_block_sum(sf, red);
    if (threadIdx.x == 0) {
        const float lr = tr * inv * loss_scale, lf = tf * inv * loss_scale;
        losses[0] = lr + lf; losses[1] = lr; losses[2] = lf;
    }
}

// loss = -mean(logits), dlogits = -1/B
__global__ void __launch_bounds__(1024)
gen_adv_kernel(const float *__restrict__ logits, int batch, float loss_scale, float *__restrict__ loss,
               float *__restrict__ dlogits)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const float inv = 1.0f / (float)batch;
    float s = 0.0f;
    for (int i = threadIdx.x; i < batch; i += blockDim.x) {
        s += logits[i];
        if (dlogits) dlogits[i] = -inv;
    }
    const float t = ipr_block_sum(s, red);
    if (threadIdx.x == 0) *loss = -t * inv * loss_scale;
}

// ---- Philox4x32-10 (Salmon et al. 2011), the counter-based generator cuRAND and PyTorch use
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

// out[4 i .. 4 i + 3] = four N(0,1) draws from counter (base + i); *counter advances by ceil(n / 4) once every CTA
// has read it (arrival ticket), so the launch is CUDA-graph replayable and every replay draws fresh numbers.
__global__ void __launch_bounds__(256)
randn_kernel(float *__restrict__ out, long long n, unsigned long long seed, unsigned long long *__restrict__ counter,
             unsigned int *__restrict__ ticket)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ unsigned long long base_sm;
    if (threadIdx.x == 0) {
        base_sm = *reinterpret_cast<volatile unsigned long long *>(counter);
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {                    // every CTA holds the old value: publish the next one
            *ticket = 0u;
            *counter = base_sm + (unsigned long long)((n + 3) >> 2);
        }
    }
    __syncthreads();
    const unsigned long long base = base_sm;
    const long long quads = (n + 3) >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
        const unsigned long long ctr = base + (unsigned long long)q;
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float v[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float r = sqrtf(-2.0f * logf(u01(c[2 * h])));
            float s, co;
            sincosf(6.28318530717958647692f * u01(c[2 * h + 1]), &s, &co);
            v[2 * h] = r * co; v[2 * h + 1] = r * s;
        }
        const long long e = q << 2;
        if (e + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            reinterpret_cast<float4 *>(out)[q] = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) if (e + k < n) out[e + k] = v[k];
        }
    }
}

// ---- pointwise losses of the SRGAN / CycleGAN steps, value and gradient in one pass
//   kind 0: mean (x - y)^2        (F.mse_loss / nn.MSELoss: models/srgan.py:49,59, models/cyclegan.py:122-143)
//   kind 1: mean |x - y|          (nn.L1Loss: models/cyclegan.py:125-133)
//   kind 2: mean BCE-with-logits against the constant y0    (models/srgan.py:36-56)
// y == nullptr: the target is the constant y0 (ones_like / zeros_like in the reference).
// partial[cta] = sum of the per-element losses of that CTA's elements; dx = weight * dloss/dx.
__global__ void __launch_bounds__(256)
pointwise_loss_kernel(const float *__restrict__ x, const float *__restrict__ y, float y0, long long n, int kind, float weight,
                      float *__restrict__ dx, float *__restrict__ partial)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    const float gscale = weight / (float)n;
    float acc = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float xv = x[i], t = y ? y[i] : y0;
        float l, g;
        if (kind == 0) { const float d = xv - t; l = d * d; g = 2.0f * d; }
        else if (kind == 1) { const float d = xv - t; l = fabsf(d); g = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f); }
        else {
            // max(x, 0) - x t + log(1 + exp(-|x|)); d/dx = sigmoid(x) - t
            l = fmaxf(xv, 0.0f) - xv * t + log1pf(expf(-fabsf(xv)));
            g = 1.0f / (1.0f + expf(-xv)) - t;
        }
        acc += l;
        if (dx) dx[i] = g * gscale;
    }
    const float tot = ipr_block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256)
pointwise_loss_finalize_kernel(const float *__restrict__ partial, int n_partial, float scale, float *__restrict__ loss)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    float acc = 0.0f;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
    const float tot = ipr_block_sum(acc, red);
    if (threadIdx.x == 0) *loss = tot * scale;
}

}  // namespace

extern "C" size_t ipr_pointwise_loss_workspace_bytes(void) { return (size_t)1024 * sizeof(float); }

extern "C" int ipr_pointwise_loss_f32(const float *x, const float *y, float y0, int64_t n, int kind, float weight, float *loss,
                                      float *dx, void *workspace, size_t workspace_bytes, ipr_stream_t stream)
{
    IPR_REQUIRE(x && loss && workspace, IPR_E_NULL);
    IPR_REQUIRE(n > 0 && kind >= 0 && kind <= 2, IPR_E_SHAPE);
    IPR_REQUIRE(workspace_bytes >= ipr_pointwise_loss_workspace_bytes(), IPR_E_WORKSPACE);
    long long ctas = (n + 1023) / 1024;
    if (ctas > 1024) ctas = 1024;
    if (ctas > 2LL * ipr_sm_count()) ctas = 2LL * ipr_sm_count();
    IPR_LAUNCH_PDL((pointwise_loss_kernel), (unsigned)ctas, 256, 0, ipr_cu(stream), x, y, y0, (long long)n, kind, weight, dx,
                   (float *)workspace);
    IPR_LAUNCH_CHECK();
    IPR_LAUNCH_PDL((pointwise_loss_finalize_kernel), 1, 256, 0, ipr_cu(stream), (const float *)workspace, (int)ctas,
                   weight / (float)n, loss);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_hinge_d_loss_f32(const float *real_logits, const float *fake_logits, int batch, float loss_scale,
                                    float *losses, float *d_real, float *d_fake, ipr_stream_t stream)
{
    IPR_REQUIRE(real_logits && fake_logits && losses, IPR_E_NULL);
    IPR_REQUIRE(batch > 0, IPR_E_SHAPE);
    IPR_LAUNCH_PDL((hinge_d_kernel), 1, 1024, 0, ipr_cu(stream), real_logits, fake_logits, batch, loss_scale, losses, d_real, d_fake);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_gen_adv_loss_f32(const float *logits, int batch, float loss_scale, float *loss, float *dlogits,
                                    ipr_stream_t stream)
{
    IPR_REQUIRE(logits && loss, IPR_E_NULL);
    IPR_REQUIRE(batch > 0, IPR_E_SHAPE);
    IPR_LAUNCH_PDL((gen_adv_kernel), 1, 1024, 0, ipr_cu(stream), logits, batch, loss_scale, loss, dlogits);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_randn_f32(float *out, int64_t n, uint64_t seed, uint64_t *counter, uint32_t *ticket,
                             ipr_stream_t stream)
{
    IPR_REQUIRE(out && counter && ticket, IPR_E_NULL);
    IPR_REQUIRE(n > 0, IPR_E_SHAPE);
    long long blocks = ((n + 3) / 4 + 255) / 256;
    const long long cap = (long long)ipr_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    IPR_LAUNCH_PDL((randn_kernel), (unsigned)blocks, 256, 0, ipr_cu(stream), out, (long long)n, (unsigned long long)seed,
                   (unsigned long long *)counter, (unsigned int *)ticket);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
