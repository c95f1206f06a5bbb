// sign.cu -- white-box signature: sign (hinge) loss over normalisation-layer gammas with its gradient,
// and the signature bit-error count.  The whole signature is a few KB (448 ... 5 248 gammas), so one
// CTA walks a by-value layer table: one launch replaces ~10 tiny kernels per layer of the reference
// (tools/sign_model.py:42-60).  Latency-bound by construction; no meaningful roofline.
#include "ipr_common.cuh"

namespace {

struct SignTable {
    ipr_sign_layer_t layer[IPR_SIGN_MAX_LAYERS];
    int n_layers;
};

__global__ void __launch_bounds__(512)
sign_loss_kernel(const __grid_constant__ SignTable tab, float gamma0, float grad_scale, int accumulate,
                 float loss_scale, float *__restrict__ loss)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    float total = 0.0f;
    for (int l = 0; l < tab.n_layers; l++) {
        const ipr_sign_layer_t L = tab.layer[l];
        const float inv_n = 1.0f / (float)L.n;
        float part = 0.0f;
        for (int c = threadIdx.x; c < L.n; c += blockDim.x) {
            const float s = L.sign[c];
            const float m = gamma0 - L.gamma[c] * s;          // hinge margin
            const bool active = m > 0.0f;
            part += active ? m : 0.0f;
            if (L.grad) {
                const float g = active ? -s * inv_n * grad_scale : 0.0f;
                L.grad[c] = accumulate ? L.grad[c] + g : g;
            }
        }
        const float layer_sum = ipr_block_sum(part, red);
        total += layer_sum * inv_n;                            // per-layer mean, summed over layers
    }
    if (threadIdx.x == 0 && loss) *loss = total * loss_scale;
}

__global__ void __launch_bounds__(512)
sign_ber_kernel(const __grid_constant__ SignTable tab, int *__restrict__ counts)
{
    __shared__ int red[32];
    int wrong = 0, bits = 0;
    for (int l = 0; l < tab.n_layers; l++) {
        const ipr_sign_layer_t L = tab.layer[l];
        for (int c = threadIdx.x; c < L.n; c += blockDim.x) {
            const float g = L.gamma[c];
            const float sg = (float)((g > 0.0f) - (g < 0.0f));   // torch.sign: 0 for 0 (and never == +-1 for NaN)
            wrong += (sg != L.sign[c]) ? 1 : 0;
        }
        bits += L.n;
    }
    wrong = ipr_warp_sum_i(wrong);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = wrong;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0;
        v = ipr_warp_sum_i(v);
        if (threadIdx.x == 0) { counts[0] = v; counts[1] = bits; }
    }
}

int fill_table(SignTable &t, const ipr_sign_layer_t *layers, int n_layers)
{
    IPR_REQUIRE(layers, IPR_E_NULL);
    IPR_REQUIRE(n_layers > 0, IPR_E_SHAPE);
    IPR_REQUIRE(n_layers <= IPR_SIGN_MAX_LAYERS, IPR_E_UNSUPPORTED);
    for (int i = 0; i < n_layers; i++) {
        IPR_REQUIRE(layers[i].gamma && layers[i].sign, IPR_E_NULL);
        IPR_REQUIRE(layers[i].n > 0, IPR_E_SHAPE);
        t.layer[i] = layers[i];
    }
    t.n_layers = n_layers;
    return IPR_OK;
}

}  // namespace

extern "C" int ipr_sign_loss_fwd_bwd_f32(const ipr_sign_layer_t *layers_host, int n_layers, float gamma0,
                                         float grad_scale, int accumulate, float loss_scale, float *loss,
                                         ipr_stream_t stream)
{
    SignTable t;
    int rc = fill_table(t, layers_host, n_layers);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(loss, IPR_E_NULL);
    IPR_LAUNCH_PDL((sign_loss_kernel), 1, 512, 0, ipr_cu(stream), t, gamma0, grad_scale, accumulate, loss_scale, loss);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_sign_ber_i32(const ipr_sign_layer_t *layers_host, int n_layers, int32_t *counts,
                                ipr_stream_t stream)
{
    SignTable t;
    int rc = fill_table(t, layers_host, n_layers);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(counts, IPR_E_NULL);
    sign_ber_kernel<<<1, 512, 0, ipr_cu(stream)>>>(t, counts);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
