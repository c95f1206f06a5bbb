/*
 * oracle/pdq_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * CPU restatement of the PDQ perceptual hash that the reference reaches through the
 * un-vendored third-party package `pdqhash==0.2.2` (requirements.txt:5; call site
 * /root/reference/tools/phash_pvalue.py:13  `h, q = pdqhash.compute(x)`).
 * pdqhash 0.2.2 is a Cython binding over facebook/ThreatExchange `pdq/cpp`; its source is
 * NOT present in /root/reference and cannot be fetched (no network), so this file restates
 * the published algorithm (hashing/pdqhashing.cpp, downscaling/downscaling.cpp,
 * hashing/torben.cpp):
 *
 *   luma  = 0.299 R + 0.587 G + 0.114 B              (float, 0..255 scale)
 *   2 x { box filter along rows ; box filter along cols }   (Jarosz; window = (dim+127)/128)
 *   decimate to 64x64 at int((i+0.5)*dim/64)
 *   B = D A D^T, D[i][j] = sqrt(2/64) cos(pi/(2*64) (i+1)(2j+1)), i<16   (two plain triple loops)
 *   median = Torben median of the 256 coefficients ; bit(i*16+j) = B[i][j] > median
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path and the real pdqhash
 * binary is unavailable, so bit values are pinned to THIS restatement.
 * Only the Hamming distance between two hashes made by the same implementation is consumed
 * (tools/phash_pvalue.py:34-35), so bit ORDER is irrelevant; bit VALUES are what is compared.
 *
 * All float arithmetic is sequential fp32 without FMA contraction (build: -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PDQ_WIN_DIV 128

static int pdq_window(int dim) { return (dim + PDQ_WIN_DIV - 1) / PDQ_WIN_DIV; }

/* One 1-D running-sum box filter (ThreatExchange downscaling.cpp box1DFloat semantics):
 * the window grows at the head, slides in the middle, shrinks at the tail. */
static void box_1d(const float *in, float *out, int n, int stride, int win)
{
    int half = (win + 2) / 2;
    int n1 = half - 1;            /* accumulate only            */
    int n2 = win - half + 1;      /* growing window, writes     */
    int n3 = n - win;             /* full window, writes        */
    int n4 = half - 1;            /* shrinking window, writes   */
    int li = 0, ri = 0, oi = 0, cur = 0;
    float sum = 0.0f;
    for (int k = 0; k < n1; k++) { sum += in[ri]; cur++; ri += stride; }
    for (int k = 0; k < n2; k++) { sum += in[ri]; cur++; out[oi] = sum / (float)cur; ri += stride; oi += stride; }
    for (int k = 0; k < n3; k++) { sum += in[ri]; sum -= in[li]; out[oi] = sum / (float)cur; li += stride; ri += stride; oi += stride; }
    for (int k = 0; k < n4; k++) { sum -= in[li]; cur--; out[oi] = sum / (float)cur; li += stride; oi += stride; }
}

static void box_rows(const float *in, float *out, int rows, int cols, int win)
{
    for (int i = 0; i < rows; i++) box_1d(in + (size_t)i * cols, out + (size_t)i * cols, cols, 1, win);
}

static void box_cols(const float *in, float *out, int rows, int cols, int win)
{
    for (int j = 0; j < cols; j++) box_1d(in + j, out + j, rows, cols, win);
}

/* D (16x64, row-major).  Exported so tests can compare the product's table bit-for-bit. */
void pdq_oracle_dct_matrix(float *d)
{
    const float scale = (float)sqrt(2.0 / 64.0);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 64; j++)
            d[i * 64 + j] = (float)(scale * cos((M_PI / 2.0 / 64.0) * (double)(i + 1) * (double)(2 * j + 1)));
}

static float torben_median(const float *m, int n)
{
    float lo = m[0], hi = m[0];
    for (int i = 1; i < n; i++) { if (m[i] < lo) lo = m[i]; if (m[i] > hi) hi = m[i]; }
    int less, greater, equal;
    float guess, maxlt, mingt;
    const int half = (n + 1) / 2;
    for (;;) {
        guess = (lo + hi) / 2;
        less = greater = equal = 0;
        maxlt = lo; mingt = hi;
        for (int i = 0; i < n; i++) {
            if (m[i] < guess) { less++; if (m[i] > maxlt) maxlt = m[i]; }
            else if (m[i] > guess) { greater++; if (m[i] < mingt) mingt = m[i]; }
            else equal++;
        }
        if (less <= half && greater <= half) break;
        else if (less > greater) hi = maxlt;
        else lo = mingt;
    }
    if (less >= half) return maxlt;
    else if (less + equal >= half) return guess;
    return mingt;
}

/* rgb: HWC uint8.  bits: 256 bytes of 0/1, index i*16+j.  coeffs (optional): the 16x16 DCT block. */
void pdq_oracle_hash_rgb(const uint8_t *rgb, int rows, int cols, uint8_t *bits, float *coeffs_out)
{
    size_t n = (size_t)rows * cols;
    float *b1 = (float *)malloc(n * sizeof(float));
    float *b2 = (float *)malloc(n * sizeof(float));
    static float a64[64][64];
    static float t16[16][64];
    float c16[256];
    float D[16 * 64];
    const float cr = 0.299f, cg = 0.587f, cb = 0.114f;

    for (size_t p = 0; p < n; p++) {
        float r = (float)rgb[3 * p], g = (float)rgb[3 * p + 1], b = (float)rgb[3 * p + 2];
        float acc = cr * r;
        acc = acc + cg * g;
        acc = acc + cb * b;
        b1[p] = acc;
    }
    if (rows == 64 && cols == 64) {
        memcpy(a64, b1, sizeof(a64));
    } else {
        int wr = pdq_window(cols), wc = pdq_window(rows);
        for (int rep = 0; rep < 2; rep++) {
            box_rows(b1, b2, rows, cols, wr);
            box_cols(b2, b1, rows, cols, wc);
        }
        for (int i = 0; i < 64; i++) {
            int ii = (int)(((i + 0.5) * rows) / 64);
            for (int j = 0; j < 64; j++) {
                int jj = (int)(((j + 0.5) * cols) / 64);
                a64[i][j] = b1[(size_t)ii * cols + jj];
            }
        }
    }
    pdq_oracle_dct_matrix(D);
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 64; j++) {
            float s = 0.0f;
            for (int k = 0; k < 64; k++) s += D[i * 64 + k] * a64[k][j];
            t16[i][j] = s;
        }
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 16; j++) {
            float s = 0.0f;
            for (int k = 0; k < 64; k++) s += t16[i][k] * D[j * 64 + k];
            c16[i * 16 + j] = s;
        }
    float med = torben_median(c16, 256);
    for (int k = 0; k < 256; k++) bits[k] = c16[k] > med ? 1 : 0;
    if (coeffs_out) memcpy(coeffs_out, c16, sizeof(c16));
    free(b1); free(b2);
}

/* Batched helper: imgs (B, rows, cols, 3) uint8 -> bits (B, 256). */
void pdq_oracle_hash_batch(const uint8_t *imgs, int batch, int rows, int cols, uint8_t *bits)
{
    for (int b = 0; b < batch; b++)
        pdq_oracle_hash_rgb(imgs + (size_t)b * rows * cols * 3, rows, cols, bits + (size_t)b * 256, NULL);
}
