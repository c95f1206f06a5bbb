// ssim.cu -- watermark reconstruction loss: SSIM (11-tap sigma-1.5 valid separable Gaussian window)
// forward AND backward in one pass over HBM.
//
// Replaces tools/loss.py:10-20,82-85 -> pytorch_msssim.SSIM(data_range=1) (10 depthwise conv2d +
// ~20 elementwise kernels forward, about twice that in autograd backward) with ONE kernel that reads
// x and y once and writes d(loss)/dx once:   algorithmic traffic 36*B*H*W bytes (3 tensors x 3 ch x 4 B).
//
// Layout / tiling
//   * one CTA = one 32x32 tile of dX for NP consecutive (n,c) planes; halo recomputed per tile
//     (a 32x32 image is a single tile with no halo, the DCGAN case);
//   * shared memory holds interleaved (x,y) pairs so that every pass moves 64-bit words:
//       pass 1 (vertical)   : (x,y) -> sum g*(x,y), sum g*(x^2+y^2, x*y)
//       pass 2 (horizontal) : -> mu_x, mu_y, E[x^2+y^2], E[xy] -> SSIM map S and dS/dp, dS/dq, dS/dr
//       pass 3 (vertical^T) , pass 4 (horizontal^T): full correlation of the three derivative maps,
//       epilogue dX = coef * (Fp + 2 x Fq + y Fr), 128-bit stores;
//   * each thread keeps a sliding window of L+10 inputs in registers and produces L outputs, so the
//     11-tap filter costs (L+10)/L shared loads per output instead of 11; taps are compile-time
//     immediates (FFMA with immediate operand).
//   * sigma_x^2 + sigma_y^2 only ever appears as a sum, so x^2 and y^2 are filtered together
//     (4 filtered quantities instead of 5).
// The SSIM-map sum is reduced per plane inside the CTA (fixed order) into a workspace and a second
// single-CTA launch adds the partials in a fixed order: results are run-to-run deterministic.
#include "ipr_common.cuh"

namespace {

constexpr int RAD = 10;            // window - 1
constexpr int TILE = 32;           // dX tile edge
constexpr int LCH = 8;             // outputs per thread per pass (sliding-window length)
constexpr float SSIM_C1 = 1.0e-4f; // (0.01 * 1)^2
constexpr float SSIM_C2 = 9.0e-4f; // (0.03 * 1)^2

// exp(-(i-5)^2 / (2*1.5^2)) / sum, evaluated in fp32 exactly as pytorch_msssim._fspecial_gauss_1d does
// (tests/test_ssim_oracle.py checks these bit patterns against torch).
__device__ constexpr float kTap[11] = {
    0x1.0d957p-10f, 0x1.f1fe02p-8f, 0x1.26eb18p-5f, 0x1.bff0fep-4f, 0x1.b43c3ep-3f, 0x1.10656p-2f,
    0x1.b43c3ep-3f, 0x1.bff0fep-4f, 0x1.26eb18p-5f, 0x1.f1fe02p-8f, 0x1.0d957p-10f};

struct SsimParams {
    const float *x, *y;
    float *dx;
    float *partial;        // [planes][tiles]
    long long planes;
    int H, W, Hv, Wv;      // Hv = H-10, Wv = W-10 (valid SSIM map)
    int tiles_r, tiles_c;
    int np;                // planes per CTA
    int normalized;
    float coef;            // -grad_scale * (normalized ? .5 : 1) / (planes*Hv*Wv)
};

struct Geom {
    int r0, r1, c0, c1;    // dX tile (global coords)
    int i0, i1, j0, j1;    // SSIM-map rows/cols needed by this tile
    int IR, IC, SR, SC, OR, OC;
    int pP, pD;            // pitches (in elements) of the input-pair and map arrays, both odd
};

__device__ __forceinline__ Geom make_geom(const SsimParams &p, int tile)
{
    Geom g;
    const int tr = tile / p.tiles_c, tc = tile - tr * p.tiles_c;
    g.r0 = tr * TILE; g.r1 = min(p.H, g.r0 + TILE);
    g.c0 = tc * TILE; g.c1 = min(p.W, g.c0 + TILE);
    g.i0 = max(0, g.r0 - RAD); g.i1 = min(p.Hv, g.r1);
    g.j0 = max(0, g.c0 - RAD); g.j1 = min(p.Wv, g.c1);
    g.SR = g.i1 - g.i0; g.SC = g.j1 - g.j0;
    g.IR = g.SR + RAD;  g.IC = g.SC + RAD;
    g.OR = g.r1 - g.r0; g.OC = g.c1 - g.c0;
    g.pP = g.IC | 1;
    g.pD = g.SC | 1;
    return g;
}

// shared-memory carve-up per plane slot (units: floats)
struct Carve {
    int P, Va, Vb, Dpq, Dr, Epq, Er, total;
};
__host__ __device__ inline Carve make_carve(int IR, int SR, int OR, int pP, int pD)
{
    Carve c;
    int o = 0;
    c.P = o;   o += 2 * IR * pP;
    c.Dpq = o; o += 2 * SR * pD;
    c.Dr = o;  o += SR * pD;
    o = (o + 1) & ~1;
    // V (passes 1-2) and E (passes 3-4) are never live together: alias them
    const int v = 4 * SR * pP, e = 3 * OR * pD + 1;
    c.Va = o; c.Vb = o + 2 * SR * pP;
    c.Epq = o; c.Er = o + 2 * OR * pD;
    c.Er = (c.Er + 1) & ~1;
    o += (v > e ? v : e) + 2;
    c.total = (o + 1) & ~1;
    return c;
}

template <bool WITH_GRAD>
__global__ void __launch_bounds__(256)
ssim_tile_kernel(const SsimParams p)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ float red[32];
    const int tid = threadIdx.x, nth = blockDim.x;
    const Geom g = make_geom(p, blockIdx.y);
    const Carve cv = make_carve(g.IR, g.SR, g.OR, g.pP, g.pD);
    const long long plane0 = (long long)blockIdx.x * p.np;
    const int np = p.np;
    const size_t plane_elems = (size_t)p.H * p.W;

    // ---------------- phase 0: global -> shared, interleaving (x, y) and de-normalising
    {
        const bool fast = (g.IC == p.W) && ((p.W & 3) == 0) && ((plane_elems & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y)) & 15) == 0;
        if (fast) {
            const int w4 = g.IC >> 2, per = g.IR * w4;
            for (int e = tid; e < np * per; e += nth) {
                const int slot = e / per, rem = e - slot * per;
                const int rr = rem / w4, c4 = rem - rr * w4;
                const long long pl = min(plane0 + slot, p.planes - 1);
                const size_t off = (size_t)pl * plane_elems + (size_t)(g.i0 + rr) * p.W + (c4 << 2);
                float4 a = ipr_ldg_stream4(reinterpret_cast<const float4 *>(p.x + off));
                float4 b = ipr_ldg_stream4(reinterpret_cast<const float4 *>(p.y + off));
                if (p.normalized) {
                    a.x = (a.x + 1.f) * .5f; a.y = (a.y + 1.f) * .5f; a.z = (a.z + 1.f) * .5f; a.w = (a.w + 1.f) * .5f;
                    b.x = (b.x + 1.f) * .5f; b.y = (b.y + 1.f) * .5f; b.z = (b.z + 1.f) * .5f; b.w = (b.w + 1.f) * .5f;
                }
                float2 *dst = reinterpret_cast<float2 *>(smem + slot * cv.total + cv.P) + rr * g.pP + (c4 << 2);
                dst[0] = make_float2(a.x, b.x); dst[1] = make_float2(a.y, b.y);
                dst[2] = make_float2(a.z, b.z); dst[3] = make_float2(a.w, b.w);
            }
        } else {
            const int per = g.IR * g.IC;
            for (int e = tid; e < np * per; e += nth) {
                const int slot = e / per, rem = e - slot * per;
                const int rr = rem / g.IC, cc = rem - rr * g.IC;
                const long long pl = min(plane0 + slot, p.planes - 1);
                const size_t off = (size_t)pl * plane_elems + (size_t)(g.i0 + rr) * p.W + (g.j0 + cc);
                float a = __ldg(p.x + off), b = __ldg(p.y + off);
                if (p.normalized) { a = (a + 1.f) * .5f; b = (b + 1.f) * .5f; }
                reinterpret_cast<float2 *>(smem + slot * cv.total + cv.P)[rr * g.pP + cc] = make_float2(a, b);
            }
        }
    }
    __syncthreads();

    // ---------------- pass 1: vertical filter of (x,y) and (x^2+y^2, xy); lanes -> columns
    {
        const int nch = (g.SR + LCH - 1) / LCH, per = g.IC * nch;
        for (int it = tid; it < np * per; it += nth) {
            const int slot = it / per, rem = it - slot * per;
            const int ch = rem / g.IC, col = rem - ch * g.IC;
            const int rb = ch * LCH;
            const float2 *src = reinterpret_cast<const float2 *>(smem + slot * cv.total + cv.P) + col;
            float2 w[LCH + RAD], q[LCH + RAD];
#pragma unroll
            for (int t = 0; t < LCH + RAD; t++) {
                w[t] = src[min(rb + t, g.IR - 1) * g.pP];
                q[t] = make_float2(fmaf(w[t].x, w[t].x, w[t].y * w[t].y), w[t].x * w[t].y);
            }
            float2 *va = reinterpret_cast<float2 *>(smem + slot * cv.total + cv.Va) + col;
            float2 *vb = reinterpret_cast<float2 *>(smem + slot * cv.total + cv.Vb) + col;
#pragma unroll
            for (int o = 0; o < LCH; o++) {
                float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    a.x = fmaf(kTap[k], w[o + k].x, a.x); a.y = fmaf(kTap[k], w[o + k].y, a.y);
                    b.x = fmaf(kTap[k], q[o + k].x, b.x); b.y = fmaf(kTap[k], q[o + k].y, b.y);
                }
                if (rb + o < g.SR) { va[(rb + o) * g.pP] = a; vb[(rb + o) * g.pP] = b; }
            }
        }
    }
    __syncthreads();

    // ---------------- pass 2: horizontal filter + SSIM map + derivative maps; lanes -> rows
    float ssum[4] = {0.f, 0.f, 0.f, 0.f};   // per plane slot (np <= 4)
    {
        const int nch = (g.SC + LCH - 1) / LCH, per = g.SR * nch;
        for (int it = tid; it < np * per; it += nth) {
            const int slot = it / per, rem = it - slot * per;
            const int ch = rem / g.SR, row = rem - ch * g.SR;
            const int cb = ch * LCH;
            const float2 *va = reinterpret_cast<const float2 *>(smem + slot * cv.total + cv.Va) + row * g.pP;
            const float2 *vb = reinterpret_cast<const float2 *>(smem + slot * cv.total + cv.Vb) + row * g.pP;
            float2 wa[LCH + RAD], wb[LCH + RAD];
#pragma unroll
            for (int t = 0; t < LCH + RAD; t++) {
                const int cc = min(cb + t, g.IC - 1);
                wa[t] = va[cc]; wb[t] = vb[cc];
            }
            float2 *dpq = reinterpret_cast<float2 *>(smem + slot * cv.total + cv.Dpq) + row * g.pD;
            float *dr = smem + slot * cv.total + cv.Dr + row * g.pD;
            const bool row_owned = (g.i0 + row) >= g.r0;       // (< r1 holds because i1 <= r1)
            float acc = 0.f;
#pragma unroll
            for (int o = 0; o < LCH; o++) {
                float mx = 0.f, my = 0.f, e2 = 0.f, exy = 0.f;
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    mx = fmaf(kTap[k], wa[o + k].x, mx);  my = fmaf(kTap[k], wa[o + k].y, my);
                    e2 = fmaf(kTap[k], wb[o + k].x, e2);  exy = fmaf(kTap[k], wb[o + k].y, exy);
                }
                const float mxx = mx * mx, myy = my * my, mxy = mx * my;
                const float A1 = 2.f * mxy + SSIM_C1;
                const float B1 = mxx + myy + SSIM_C1;
                const float A2 = 2.f * (exy - mxy) + SSIM_C2;
                const float B2 = (e2 - mxx - myy) + SSIM_C2;
                const float iB1 = 1.0f / B1, iB2 = 1.0f / B2;
                const float iB12 = iB1 * iB2;
                const float S = A1 * A2 * iB12;
                if (cb + o < g.SC) {
                    if (WITH_GRAD) {
                        const float dS_dq = -S * iB2;
                        const float dS_dr = 2.f * A1 * iB12;
                        const float dS_dp = 2.f * my * (A2 - A1) * iB12 + 2.f * mx * S * (iB2 - iB1);
                        dpq[cb + o] = make_float2(dS_dp, dS_dq);
                        dr[cb + o] = dS_dr;
                    }
                    if (row_owned && (g.j0 + cb + o) >= g.c0) acc += S;
                }
            }
#pragma unroll
            for (int s = 0; s < 4; s++) if (s == slot) ssum[s] += acc;
        }
    }
    // per-plane partial sums of the SSIM map (fixed-order block reduction)
    for (int s = 0; s < np; s++) {
        const float tot = ipr_block_sum(ssum[s], red);
        if (tid == 0 && plane0 + s < p.planes)
            p.partial[(size_t)(plane0 + s) * gridDim.y + blockIdx.y] = tot;
    }
    if (!WITH_GRAD) return;
    __syncthreads();

    // ---------------- pass 3: vertical transposed filter of (dp,dq,dr); lanes -> columns
    {
        const int nch = (g.OR + LCH - 1) / LCH, per = g.SC * nch;
        for (int it = tid; it < np * per; it += nth) {
            const int slot = it / per, rem = it - slot * per;
            const int ch = rem / g.SC, col = rem - ch * g.SC;
            const int ob = ch * LCH;                     // local output row
            const float2 *dpq = reinterpret_cast<const float2 *>(smem + slot * cv.total + cv.Dpq) + col;
            const float *dr = smem + slot * cv.total + cv.Dr + col;
            float2 w[LCH + RAD]; float wr[LCH + RAD];
#pragma unroll
            for (int t = 0; t < LCH + RAD; t++) {
                const int li = (g.r0 + ob + t - RAD) - g.i0;     // local SSIM-map row feeding output rows
                const bool ok = (li >= 0) && (li < g.SR);
                const int lc = ok ? li : 0;
                const float2 v = dpq[lc * g.pD]; const float vr = dr[lc * g.pD];
                w[t] = ok ? v : make_float2(0.f, 0.f); wr[t] = ok ? vr : 0.f;
            }
            float2 *epq = reinterpret_cast<float2 *>(smem + slot * cv.total + cv.Epq) + col;
            float *er = smem + slot * cv.total + cv.Er + col;
#pragma unroll
            for (int o = 0; o < LCH; o++) {
                float2 a = make_float2(0.f, 0.f); float b = 0.f;
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    a.x = fmaf(kTap[k], w[o + k].x, a.x); a.y = fmaf(kTap[k], w[o + k].y, a.y);
                    b = fmaf(kTap[k], wr[o + k], b);
                }
                if (ob + o < g.OR) { epq[(ob + o) * g.pD] = a; er[(ob + o) * g.pD] = b; }
            }
        }
    }
    __syncthreads();

    // ---------------- pass 4: horizontal transposed filter + epilogue; lanes -> rows
    {
        const int nch = (g.OC + LCH - 1) / LCH, per = g.OR * nch;
        const bool vec_ok = ((p.W & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dx) & 15) == 0);
        for (int it = tid; it < np * per; it += nth) {
            const int slot = it / per, rem = it - slot * per;
            const int ch = rem / g.OR, row = rem - ch * g.OR;
            const int ob = ch * LCH;                     // local output col
            if (plane0 + slot >= p.planes) continue;
            const float2 *epq = reinterpret_cast<const float2 *>(smem + slot * cv.total + cv.Epq) + row * g.pD;
            const float *er = smem + slot * cv.total + cv.Er + row * g.pD;
            float2 w[LCH + RAD]; float wr[LCH + RAD];
#pragma unroll
            for (int t = 0; t < LCH + RAD; t++) {
                const int lj = (g.c0 + ob + t - RAD) - g.j0;
                const bool ok = (lj >= 0) && (lj < g.SC);
                const int lc = ok ? lj : 0;
                const float2 v = epq[lc]; const float vr = er[lc];
                w[t] = ok ? v : make_float2(0.f, 0.f); wr[t] = ok ? vr : 0.f;
            }
            const float2 *pin = reinterpret_cast<const float2 *>(smem + slot * cv.total + cv.P) +
                                (g.r0 + row - g.i0) * g.pP + (g.c0 + ob - g.j0);
            float outv[LCH];
#pragma unroll
            for (int o = 0; o < LCH; o++) {
                float fp = 0.f, fq = 0.f, fr = 0.f;
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    fp = fmaf(kTap[k], w[o + k].x, fp); fq = fmaf(kTap[k], w[o + k].y, fq);
                    fr = fmaf(kTap[k], wr[o + k], fr);
                }
                float2 xy = make_float2(0.f, 0.f);
                if (ob + o < g.OC) xy = pin[o];
                outv[o] = p.coef * (fp + 2.f * xy.x * fq + xy.y * fr);
            }
            float *dst = p.dx + (size_t)(plane0 + slot) * plane_elems + (size_t)(g.r0 + row) * p.W + (g.c0 + ob);
            if (vec_ok && ob + LCH <= g.OC) {
                ipr_stg_stream4(reinterpret_cast<float4 *>(dst), make_float4(outv[0], outv[1], outv[2], outv[3]));
                ipr_stg_stream4(reinterpret_cast<float4 *>(dst) + 1, make_float4(outv[4], outv[5], outv[6], outv[7]));
            } else {
#pragma unroll
                for (int o = 0; o < LCH; o++) if (ob + o < g.OC) dst[o] = outv[o];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 32x32 planes (the DCGAN / CIFAR shape): one WARP owns one plane through all four passes, so there is no block
// barrier at all -- warps only __syncwarp() between passes and drift apart, overlapping one plane's loads with
// another's arithmetic.  Lane = column (vertical passes) or row (horizontal passes); every pass is a fully unrolled
// 1-D filter over a whole line kept in registers (each shared-memory value is read exactly once), zero taps of the
// transposed filters are dropped at compile time, and the paired quantities ((x,y), (x^2+y^2, xy), (dS/dp, dS/dq))
// go through Blackwell's packed fma.rn.f32x2 (two fp32 FMAs per issue slot).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));      // <= 1 ulp; the map denominators are >= C1*C2 > 0
    return r;
}

// Per-warp shared-memory carve (bytes).  Every array is indexed [line][lane] or [lane][line], so the aliases below are
// same-lane aliases only (no cross-lane hazard beyond the __syncwarp between passes):
//   P   float2 [32][33]  @0       (x, y) pairs; alive for the whole plane (pass 4 reads x, y from it and parks dX in .x)
//   V   float2 [22][33]  @8448    vertically filtered pairs, written twice (means, then second moments)
//   M   rows of 264 B    @8448    row r = { r-map[22] float @0, (p,q)-maps[22] float2 @88 }: rows 0..21 first hold
//                                 D = dS/d(p,q,r) (row r overlays V row r, written behind the read position of the lane
//                                 that owns that row), then all 32 rows hold E = G^T * D (row r overlays D row r).
constexpr int S32_PF = 4;                        // shared-memory loads run this many lines ahead of the stores
constexpr int S32_PITCH = 33;                    // float2 per row of P and V
constexpr int S32_OFF_M = 32 * S32_PITCH * 8;    // 8448
constexpr int S32_ROW_M = 22 * 4 + 22 * 8;       // 264 = one V row
constexpr int S32_WARP_BYTES = S32_OFF_M + 32 * S32_ROW_M;           // 16896 -> 12 warps per SM
static_assert(S32_ROW_M == S32_PITCH * 8, "a D/E row overlays exactly one V row");
static_assert(S32_PF >= 1, "pass 2b stores D(p,q)[j] over V[j+11], which must already be in registers");
static_assert(S32_WARP_BYTES % 16 == 0, "warp regions stay 16-byte aligned");

// The transposed passes filter three maps: packed FFMA2 for (p,q) next to scalar FFMA for r.  FFMA2 takes the FMA pipe
// for two cycles (profiles/r1_fma_rate_probe.txt: same 128 FMA/clk/SM as scalar FFMA) but frees an issue slot for the
// LDS/STS traffic; measured 1302 us (packed) vs 1364 us (all scalar) at 65536x3x32x32.
#ifndef IPR_SSIM_BWD_SCALAR
#define S32_MUL2(k, d) fmul2(tap2(k), d)
#define S32_FMA2(k, d, a) ffma2(tap2(k), d, a)
#else
#define S32_MUL2(k, d) make_float2(kTap[k] * (d).x, kTap[k] * (d).y)
#define S32_FMA2(k, d, a) make_float2(fmaf(kTap[k], (d).x, (a).x), fmaf(kTap[k], (d).y, (a).y))
#endif

template <bool WITH_GRAD, bool NORM>
__global__ void __launch_bounds__(128, 3)
ssim32_warp_kernel(const SsimParams p)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw + warp * S32_WARP_BYTES;
    float2 *P = reinterpret_cast<float2 *>(base);
    float2 *V = reinterpret_cast<float2 *>(base + S32_OFF_M);
    auto Mr = [&](int row) { return reinterpret_cast<float *>(base + S32_OFF_M + row * S32_ROW_M); };
    auto Mpq = [&](int row) { return reinterpret_cast<float2 *>(base + S32_OFF_M + row * S32_ROW_M + 88); };
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    float2 g2[6];
#pragma unroll
    for (int k = 0; k < 6; k++) g2[k] = make_float2(kTap[k], kTap[k]);
    auto tap2 = [&](int k) { return g2[k <= 5 ? k : 10 - k]; };
    const int l22 = lane < 22 ? lane : 21;       // lanes 22..31 shadow line 21 in the 22-line passes (stores masked)
    const bool act = lane < 22;

    for (long long plane = (long long)blockIdx.x * (blockDim.x >> 5) + warp; plane < p.planes; plane += warps_total) {
        const float *xg = p.x + plane * 1024, *yg = p.y + plane * 1024;
        if (plane + warps_total < p.planes) {                  // pull this warp's next plane into L2 (one 128-B line per lane)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(xg + warps_total * 1024 + lane * 32));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(yg + warps_total * 1024 + lane * 32));
        }
        // ---- phase 0: 8 + 8 coalesced 128-bit loads per lane, interleaved into (x, y) pairs
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int v4 = it * 32 + lane;                     // float4 index inside the plane
            float4 a = ipr_ldg_stream4(reinterpret_cast<const float4 *>(xg) + v4);
            float4 b = ipr_ldg_stream4(reinterpret_cast<const float4 *>(yg) + v4);
            if (NORM) {
                a.x = (a.x + 1.f) * .5f; a.y = (a.y + 1.f) * .5f; a.z = (a.z + 1.f) * .5f; a.w = (a.w + 1.f) * .5f;
                b.x = (b.x + 1.f) * .5f; b.y = (b.y + 1.f) * .5f; b.z = (b.z + 1.f) * .5f; b.w = (b.w + 1.f) * .5f;
            }
            float2 *dst = P + (v4 >> 3) * S32_PITCH + ((v4 & 7) << 2);
            dst[0] = make_float2(a.x, b.x); dst[1] = make_float2(a.y, b.y);
            dst[2] = make_float2(a.z, b.z); dst[3] = make_float2(a.w, b.w);
        }
        __syncwarp();
        // Every pass below streams its input line once and scatters each value into the (at most 11) outputs it
        // feeds: 11 independent FMA chains in flight, ~11 live accumulators instead of a whole line in registers.
        // ---- pass 1a: vertical filter of (x, y); lane = column
        {
            float2 acc[22], buf[32];
#pragma unroll
            for (int r = 0; r < S32_PF; r++) buf[r] = P[r * S32_PITCH + lane];
#pragma unroll
            for (int r = 0; r < 32; r++) {
                if (r + S32_PF < 32) buf[r + S32_PF] = P[(r + S32_PF) * S32_PITCH + lane];
                const float2 v = buf[r];
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    const int i = r - k;
                    if (i >= 0 && i < 22) acc[i] = (k == 0) ? fmul2(tap2(0), v) : ffma2(tap2(k), v, acc[i]);
                }
                if (r >= RAD) V[(r - RAD) * S32_PITCH + lane] = acc[r - RAD];
            }
        }
        __syncwarp();
        // ---- pass 2a: horizontal filter -> local means (mu_x, mu_y) kept in registers; lane = map row
        float2 m[22];
        {
            float2 buf[32];
#pragma unroll
            for (int c = 0; c < S32_PF; c++) buf[c] = V[l22 * S32_PITCH + c];
#pragma unroll
            for (int c = 0; c < 32; c++) {
                if (c + S32_PF < 32) buf[c + S32_PF] = V[l22 * S32_PITCH + c + S32_PF];
                const float2 v = buf[c];
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    const int j = c - k;
                    if (j >= 0 && j < 22) m[j] = (k == 0) ? fmul2(tap2(0), v) : ffma2(tap2(k), v, m[j]);
                }
            }
        }
        __syncwarp();
        // ---- pass 1b: vertical filter of (x^2 + y^2, x*y); lane = column
        {
            float2 acc[22], buf[32];
#pragma unroll
            for (int r = 0; r < S32_PF; r++) buf[r] = P[r * S32_PITCH + lane];
#pragma unroll
            for (int r = 0; r < 32; r++) {
                if (r + S32_PF < 32) buf[r + S32_PF] = P[(r + S32_PF) * S32_PITCH + lane];
                const float2 w = buf[r];
                const float2 v = make_float2(fmaf(w.x, w.x, w.y * w.y), w.x * w.y);
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    const int i = r - k;
                    if (i >= 0 && i < 22) acc[i] = (k == 0) ? fmul2(tap2(0), v) : ffma2(tap2(k), v, acc[i]);
                }
                if (r >= RAD) V[(r - RAD) * S32_PITCH + lane] = acc[r - RAD];
            }
        }
        __syncwarp();
        // ---- pass 2b: horizontal filter of the second moments + SSIM map + its derivatives; lane = map row
        float ssum = 0.f;
        {
            float2 e[22], buf[32];
            float *drow = Mr(l22);
            float2 *dpqrow = Mpq(l22);
#pragma unroll
            for (int c = 0; c < S32_PF; c++) buf[c] = V[l22 * S32_PITCH + c];
#pragma unroll
            for (int c = 0; c < 32; c++) {
                if (c + S32_PF < 32) buf[c + S32_PF] = V[l22 * S32_PITCH + c + S32_PF];
                const float2 v = buf[c];
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    const int j = c - k;
                    if (j >= 0 && j < 22) e[j] = (k == 0) ? fmul2(tap2(0), v) : ffma2(tap2(k), v, e[j]);
                }
                if (c >= RAD) {
                    const int j = c - RAD;
                    const float mx = m[j].x, my = m[j].y;
                    const float mxy = mx * my;
                    const float B1 = fmaf(my, my, fmaf(mx, mx, SSIM_C1));
                    const float A1 = fmaf(2.f, mxy, SSIM_C1);
                    const float A2 = fmaf(2.f, e[j].y - mxy, SSIM_C2);
                    const float B2 = (e[j].x + (SSIM_C1 + SSIM_C2)) - B1;
                    const float iB12 = rcp_approx(B1 * B2);
                    const float S = A1 * A2 * iB12;
                    ssum += S;
                    if (WITH_GRAD) {
                        // dS/dp = 2 mu_y (A2 - A1)/(B1 B2) + 2 mu_x S (1/B2 - 1/B1);  1/B2 - 1/B1 = (B1 - B2)/(B1 B2)
                        const float t = 2.f * iB12;
                        const float dp = t * fmaf(my, A2 - A1, mx * S * (B1 - B2));
                        const float dq = -S * B1 * iB12;
                        if (act) {                             // lands on V[row][0..c+1] of this lane's own row: already read
                            drow[j] = t * A1;
                            dpqrow[j] = make_float2(dp, dq);
                        }
                    }
                }
            }
        }
        ssum = ipr_warp_sum(act ? ssum : 0.f);
        if (lane == 0) p.partial[plane] = ssum;
        __syncwarp();
        if (!WITH_GRAD) continue;
        // ---- pass 3: vertical transposed filter; lane = map column.  E row r overlays D row r of the same lane.
        {
            float2 acc[32], buf[22]; float accr[32], bufr[22];
#pragma unroll
            for (int i = 0; i < S32_PF; i++) { buf[i] = Mpq(i)[l22]; bufr[i] = Mr(i)[l22]; }
#pragma unroll
            for (int i = 0; i < 22; i++) {
                if (i + S32_PF < 22) { buf[i + S32_PF] = Mpq(i + S32_PF)[l22]; bufr[i + S32_PF] = Mr(i + S32_PF)[l22]; }
                const float2 d = buf[i];
                const float dr = bufr[i];
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    const int r = i + k;
                    if (k == RAD || i == 0) { acc[r] = S32_MUL2(k, d); accr[r] = kTap[k] * dr; }
                    else { acc[r] = S32_FMA2(k, d, acc[r]); accr[r] = fmaf(kTap[k], dr, accr[r]); }
                }
                if (act) { Mpq(i)[lane] = acc[i]; Mr(i)[lane] = accr[i]; }     // row i is complete (fed by d[i-10..i])
            }
            if (act) {
#pragma unroll
                for (int r = 22; r < 32; r++) { Mpq(r)[lane] = acc[r]; Mr(r)[lane] = accr[r]; }
            }
        }
        __syncwarp();
        // ---- pass 4: horizontal transposed filter + chain rule with (x, y) from P; lane = image row; dX parks in P.x
        {
            float2 acc[32], buf[22]; float accr[32], bufr[22];
            const float2 *epq = Mpq(lane);
            const float *er = Mr(lane);
            float2 *prow = P + lane * S32_PITCH;
#pragma unroll
            for (int j = 0; j < S32_PF; j++) { buf[j] = epq[j]; bufr[j] = er[j]; }
#pragma unroll
            for (int j = 0; j < 22; j++) {
                if (j + S32_PF < 22) { buf[j + S32_PF] = epq[j + S32_PF]; bufr[j + S32_PF] = er[j + S32_PF]; }
                const float2 d = buf[j];
                const float dr = bufr[j];
#pragma unroll
                for (int k = 0; k <= RAD; k++) {
                    const int c = j + k;
                    if (k == RAD || j == 0) { acc[c] = S32_MUL2(k, d); accr[c] = kTap[k] * dr; }
                    else { acc[c] = S32_FMA2(k, d, acc[c]); accr[c] = fmaf(kTap[k], dr, accr[c]); }
                }
                {                                              // column j is complete
                    const float2 xy = prow[j];
                    prow[j].x = p.coef * fmaf(xy.y, accr[j], fmaf(2.f * xy.x, acc[j].y, acc[j].x));
                }
            }
#pragma unroll
            for (int c = 22; c < 32; c++) {
                const float2 xy = prow[c];
                prow[c].x = p.coef * fmaf(xy.y, accr[c], fmaf(2.f * xy.x, acc[c].y, acc[c].x));
            }
        }
        __syncwarp();
        // ---- copy-out: coalesced 128-bit stores of dX
        {
            float4 *dx4 = reinterpret_cast<float4 *>(p.dx + plane * 1024);
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int v4 = it * 32 + lane;
                const float2 *src = P + (v4 >> 3) * S32_PITCH + ((v4 & 7) << 2);
                ipr_stg_stream4(dx4 + v4, make_float4(src[0].x, src[1].x, src[2].x, src[3].x));
            }
        }
        __syncwarp();
    }
}

// loss = 1 - (sum of all partials) / count        (single CTA, fixed order)
__global__ void __launch_bounds__(1024)
ssim_finalize_loss_kernel(const float *__restrict__ partial, long long n, float inv_count, float loss_scale,
                          float *__restrict__ loss)
{
    ipr_pdl_wait();
    ipr_pdl_trigger();
    __shared__ float red[32];
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    const float tot = ipr_block_sum(acc, red);
    if (threadIdx.x == 0) *loss = (1.0f - tot * inv_count) * loss_scale;
}

// out[n] = (sum over the sample's C planes and tiles) / (C*Hv*Wv)
__global__ void __launch_bounds__(256)
ssim_finalize_sample_kernel(const float *__restrict__ partial, long long batch, int per_sample, float inv_count,
                            float *__restrict__ out)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= batch) return;
    const float *src = partial + n * per_sample;
    float acc = 0.f;
    for (int i = 0; i < per_sample; i++) acc += src[i];
    out[n] = acc * inv_count;
}

struct Plan {
    int tiles_r, tiles_c, np, threads;
    size_t smem;
};

Plan make_plan(int H, int W)
{
    Plan pl;
    pl.tiles_r = (H + TILE - 1) / TILE;
    pl.tiles_c = (W + TILE - 1) / TILE;
    // worst-case tile geometry on this image
    const int maxOR = H < TILE ? H : TILE, maxOC = W < TILE ? W : TILE;
    int maxSR = 0, maxSC = 0;
    for (int t = 0; t < pl.tiles_r; t++) {
        int r0 = t * TILE, r1 = r0 + TILE < H ? r0 + TILE : H;
        int i0 = r0 - RAD > 0 ? r0 - RAD : 0, i1 = r1 < H - RAD ? r1 : H - RAD;
        if (i1 - i0 > maxSR) maxSR = i1 - i0;
    }
    for (int t = 0; t < pl.tiles_c; t++) {
        int c0 = t * TILE, c1 = c0 + TILE < W ? c0 + TILE : W;
        int j0 = c0 - RAD > 0 ? c0 - RAD : 0, j1 = c1 < W - RAD ? c1 : W - RAD;
        if (j1 - j0 > maxSC) maxSC = j1 - j0;
    }
    const Carve cv = make_carve(maxSR + RAD, maxSR, maxOR, (maxSC + RAD) | 1, maxSC | 1);
    const size_t per_plane = (size_t)cv.total * sizeof(float);
    // small images: several planes per CTA so that every pass has enough work items for its warps
    const long long pix = (long long)H * W;
    pl.np = pix <= 1024 ? 2 : 1;
    pl.threads = 128;
    if (pix > 1024) pl.threads = 256;
    pl.smem = per_plane * pl.np;
    return pl;
}

int check_common(const float *x, const float *y, int64_t batch, int C, int H, int W)
{
    IPR_REQUIRE(x && y, IPR_E_NULL);
    IPR_REQUIRE(batch > 0 && C > 0 && H > 0 && W > 0, IPR_E_SHAPE);
    IPR_REQUIRE(H > RAD && W > RAD, IPR_E_UNSUPPORTED);   // the window must fit (pytorch_msssim skips the pass otherwise)
    IPR_REQUIRE((long long)batch * C < (1LL << 31), IPR_E_UNSUPPORTED);
    return IPR_OK;
}

template <bool WITH_GRAD, bool NORM>
int launch_warp32_impl(const SsimParams &p, cudaStream_t st)
{
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * S32_WARP_BYTES;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(ssim32_warp_kernel<WITH_GRAD, NORM>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    long long ctas = (p.planes + WARPS - 1) / WARPS;
    const long long cap = (long long)ipr_sm_count() * 3;          // 3 CTAs (12 warps) per SM, persistent over planes
    if (ctas > cap) ctas = cap;
    IPR_LAUNCH_PDL((ssim32_warp_kernel<WITH_GRAD, NORM>), (unsigned)ctas, WARPS * 32, smem, st, p);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

template <bool WITH_GRAD>
int launch_warp32(const SsimParams &p, cudaStream_t st)
{
    return p.normalized ? launch_warp32_impl<WITH_GRAD, true>(p, st) : launch_warp32_impl<WITH_GRAD, false>(p, st);
}

inline bool use_warp32(const SsimParams &p) {
    return p.H == 32 && p.W == 32 && ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y) |
                                      reinterpret_cast<uintptr_t>(p.dx)) & 15) == 0;
}

template <bool WITH_GRAD>
int launch_tiles(const SsimParams &p, const Plan &pl, cudaStream_t st)
{
    if (use_warp32(p)) return launch_warp32<WITH_GRAD>(p, st);
    static bool attr_done[2] = {false, false};
    if (!attr_done[WITH_GRAD]) {
        cudaError_t e = cudaFuncSetAttribute(ssim_tile_kernel<WITH_GRAD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done[WITH_GRAD] = true;
    }
    const long long groups = (p.planes + pl.np - 1) / pl.np;
    IPR_REQUIRE(pl.tiles_r * pl.tiles_c <= 65535, IPR_E_UNSUPPORTED);
    dim3 grid((unsigned)groups, (unsigned)(pl.tiles_r * pl.tiles_c));
    ssim_tile_kernel<WITH_GRAD><<<grid, pl.threads, pl.smem, st>>>(p);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

}  // namespace

extern "C" size_t ipr_ssim_workspace_bytes(int64_t batch, int channels, int height, int width)
{
    if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return 0;
    const size_t tiles = (size_t)((height + TILE - 1) / TILE) * ((width + TILE - 1) / TILE);
    return (size_t)batch * channels * tiles * sizeof(float);
}

extern "C" int ipr_ssim_fwd_bwd_f32(const float *x, const float *y, float *dx, float *loss,
                                    void *workspace, size_t workspace_bytes,
                                    int64_t batch, int channels, int height, int width,
                                    int normalized, float grad_scale, float loss_scale, ipr_stream_t stream)
{
    int rc = check_common(x, y, batch, channels, height, width);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(loss && workspace, IPR_E_NULL);
    IPR_REQUIRE(workspace_bytes >= ipr_ssim_workspace_bytes(batch, channels, height, width), IPR_E_WORKSPACE);
    const Plan pl = make_plan(height, width);
    SsimParams p;
    p.x = x; p.y = y; p.dx = dx; p.partial = (float *)workspace;
    p.planes = (long long)batch * channels;
    p.H = height; p.W = width; p.Hv = height - RAD; p.Wv = width - RAD;
    p.tiles_r = pl.tiles_r; p.tiles_c = pl.tiles_c; p.np = pl.np; p.normalized = normalized;
    const double count = (double)p.planes * p.Hv * p.Wv;
    p.coef = (float)(-(double)grad_scale * (normalized ? 0.5 : 1.0) / count);
    rc = dx ? launch_tiles<true>(p, pl, ipr_cu(stream)) : launch_tiles<false>(p, pl, ipr_cu(stream));
    if (rc != IPR_OK) return rc;
    const long long n = p.planes * pl.tiles_r * pl.tiles_c;
    IPR_LAUNCH_PDL((ssim_finalize_loss_kernel), 1, 1024, 0, ipr_cu(stream), p.partial, n, (float)(1.0 / count), loss_scale, loss);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}

extern "C" int ipr_ssim_per_sample_f32(const float *x, const float *y, float *out,
                                       void *workspace, size_t workspace_bytes,
                                       int64_t batch, int channels, int height, int width, ipr_stream_t stream)
{
    int rc = check_common(x, y, batch, channels, height, width);
    if (rc != IPR_OK) return rc;
    IPR_REQUIRE(out && workspace, IPR_E_NULL);
    IPR_REQUIRE(workspace_bytes >= ipr_ssim_workspace_bytes(batch, channels, height, width), IPR_E_WORKSPACE);
    const Plan pl = make_plan(height, width);
    SsimParams p;
    p.x = x; p.y = y; p.dx = nullptr; p.partial = (float *)workspace;
    p.planes = (long long)batch * channels;
    p.H = height; p.W = width; p.Hv = height - RAD; p.Wv = width - RAD;
    p.tiles_r = pl.tiles_r; p.tiles_c = pl.tiles_c; p.np = pl.np; p.normalized = 0;
    p.coef = 0.f;
    rc = launch_tiles<false>(p, pl, ipr_cu(stream));
    if (rc != IPR_OK) return rc;
    const int per_sample = channels * pl.tiles_r * pl.tiles_c;
    const float inv = (float)(1.0 / ((double)channels * p.Hv * p.Wv));
    ssim_finalize_sample_kernel<<<(unsigned)((batch + 255) / 256), 256, 0, ipr_cu(stream)>>>(
        p.partial, batch, per_sample, inv, out);
    IPR_LAUNCH_CHECK();
    return IPR_OK;
}
