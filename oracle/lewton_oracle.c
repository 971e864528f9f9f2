/*
 * lewton_oracle.c -- CPU restatement of lewton's packet-synthesis arithmetic.
 * TEST INFRASTRUCTURE ONLY -- see lewton_oracle.h for the rules and the pinning.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * All float arithmetic is binary32, one rounding per operation, in the operand
 * order of the cited reference lines.  Integer code uses the wrapping semantics
 * a Rust release build has.
 */
#include "lewton_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* std::f32::consts::PI */
static const float PI_F = 3.14159265358979323846f;

/* ------------------------------------------------------------------------- */
/* header_cached.rs:43-110                                                    */
/* ------------------------------------------------------------------------- */

/* header_cached.rs:43-54 */
static float win_slope(unsigned x, unsigned n)
{
    float v = sinf(0.5f * PI_F * ((float)x + 0.5f) / (float)n);
    return sinf(0.5f * PI_F * v * v);
}

/* lib.rs:174-176 */
static uint32_t bit_reverse32(uint32_t v)
{
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) {
        r = (r << 1) | (v & 1u);
        v >>= 1;
    }
    return r;
}

lwo_tables *lwo_tables_new(int bs)
{
    if (bs < 6 || bs > 13) return NULL;
    lwo_tables *t = (lwo_tables *)calloc(1, sizeof(*t));
    int n = 1 << bs, n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
    t->bs = bs;
    t->n = n;
    t->a = (float *)malloc(sizeof(float) * n2);
    t->b = (float *)malloc(sizeof(float) * n2);
    t->c = (float *)malloc(sizeof(float) * n4);
    t->window = (float *)malloc(sizeof(float) * n2);
    t->bitrev = (uint32_t *)malloc(sizeof(uint32_t) * n8);

    /* header_cached.rs:56-62 generate_window(n/2) */
    for (int i = 0; i < n2; i++) t->window[i] = win_slope((unsigned)i, (unsigned)n2);

    /* header_cached.rs:64-99 compute_twiddle_factors */
    float pi_4_n = 4.0f * PI_F / (float)n;
    float pi_05_n = 0.5f * PI_F / (float)n;
    float pi_2_n = 2.0f * PI_F / (float)n;
    int k2 = 0;
    for (int k = 0; k < n4; k++) {
        t->a[2 * k] = cosf((float)k * pi_4_n);
        t->a[2 * k + 1] = -sinf((float)k * pi_4_n);
        t->b[2 * k] = cosf((float)(k2 + 1) * pi_05_n) * 0.5f;
        t->b[2 * k + 1] = sinf((float)(k2 + 1) * pi_05_n) * 0.5f;
        k2 += 2;
    }
    k2 = 0;
    for (int k = 0; k < n8; k++) {
        t->c[2 * k] = cosf((float)(k2 + 1) * pi_2_n);
        t->c[2 * k + 1] = -sinf((float)(k2 + 1) * pi_2_n);
        k2 += 2;
    }
    /* header_cached.rs:101-110 compute_bitreverse */
    for (int i = 0; i < n8; i++)
        t->bitrev[i] = (bit_reverse32((uint32_t)i) >> (32 - bs + 3)) << 2;
    return t;
}

void lwo_tables_free(lwo_tables *t)
{
    if (!t) return;
    free(t->a); free(t->b); free(t->c); free(t->window); free(t->bitrev);
    free(t);
}
const float *lwo_tables_a(const lwo_tables *t) { return t->a; }
const float *lwo_tables_b(const lwo_tables *t) { return t->b; }
const float *lwo_tables_c(const lwo_tables *t) { return t->c; }
const float *lwo_tables_window(const lwo_tables *t) { return t->window; }
const uint32_t *lwo_tables_bitrev(const lwo_tables *t) { return t->bitrev; }

/* ------------------------------------------------------------------------- */
/* imdct.rs                                                                   */
/* ------------------------------------------------------------------------- */

/* One complex rotate-and-sum butterfly shared by imdct.rs:36-41, 94-99, 161-166:
 * hi is the "ee0"/"e0" pair (odd index i, then i-1), lo the "ee2"/"e2" pair. */
static inline void bfly(float *e, long hi, long lo, float w0, float w1)
{
    float k00 = e[hi] - e[lo];
    float k01 = e[hi - 1] - e[lo - 1];
    e[hi] = e[hi] + e[lo];
    e[hi - 1] = e[hi - 1] + e[lo - 1];
    e[lo] = k00 * w0 - k01 * w1;
    e[lo - 1] = k01 * w0 + k00 * w1;
}

/* imdct.rs:14-71 imdct_step3_iter0_loop: twiddle stride fixed at 8 */
static void step3_iter0(int n, float *e, long i_off, long k_off, const float *a)
{
    long hi = i_off, lo = i_off + k_off;
    long ao = 0;
    for (int it = 0; it < (n >> 2); it++) {
        for (int q = 0; q < 4; q++) {
            bfly(e, hi - 2 * q, lo - 2 * q, a[ao], a[ao + 1]);
            ao += 8;
        }
        hi -= 8;
        lo -= 8;
    }
}

/* imdct.rs:73-133 imdct_step3_inner_r_loop: twiddle stride k1 */
static void step3_r(int lim, float *e, long d0, long k_off, const float *a, int k1)
{
    long hi = d0, lo = d0 + k_off;
    long ao = 0;
    for (int it = 0; it < (lim >> 2); it++) {
        for (int q = 0; q < 4; q++) {
            bfly(e, hi - 2 * q, lo - 2 * q, a[ao], a[ao + 1]);
            ao += k1;
        }
        hi -= 8;
        lo -= 8;
    }
}

/* imdct.rs:135-199 imdct_step3_inner_s_loop: 4 twiddle pairs hoisted, walk by k0 */
static void step3_s(int n, float *e, long i_off, long k_off, const float *a, int a_off, int k0)
{
    float w[8];
    for (int q = 0; q < 4; q++) {
        w[2 * q] = a[a_off * q];
        w[2 * q + 1] = a[a_off * q + 1];
    }
    long hi = i_off, lo = i_off + k_off;
    for (int it = 0;;) {
        for (int q = 0; q < 4; q++) bfly(e, hi - 2 * q, lo - 2 * q, w[2 * q], w[2 * q + 1]);
        it++;
        if (it >= n) break;
        hi -= k0;
        lo -= k0;
    }
}

/* imdct.rs:201-232 iter_54; z7 points at the reference's zm7[7] */
static inline void iter_54(float *z7)
{
#define Z(i) z7[(i) - 7]
    float k00 = Z(7) - Z(3);
    float y0 = Z(7) + Z(3);
    float y2 = Z(5) + Z(1);
    float k22 = Z(5) - Z(1);
    Z(7) = y0 + y2;
    Z(5) = y0 - y2;
    float k33 = Z(4) - Z(0);
    Z(3) = k00 + k33;
    Z(1) = k00 - k33;
    float k11 = Z(6) - Z(2);
    float y1 = Z(6) + Z(2);
    float y3 = Z(4) + Z(0);
    Z(6) = y1 + y3;
    Z(4) = y1 - y3;
    Z(2) = k11 - k22;
    Z(0) = k11 + k22;
#undef Z
}

/* imdct.rs:234-288 imdct_step3_inner_s_loop_ld654 */
static void step3_ld654(int n, float *e, long i_off, const float *a, int base_n)
{
    float a2 = a[base_n >> 3];
    long z = i_off;
    long stop = i_off - 16L * (n - 1);
    for (;;) {
        float k00, k11;
        k00 = e[z] - e[z - 8];
        k11 = e[z - 1] - e[z - 9];
        e[z] = e[z] + e[z - 8];
        e[z - 1] = e[z - 1] + e[z - 9];
        e[z - 8] = k00;
        e[z - 9] = k11;

        k00 = e[z - 2] - e[z - 10];
        k11 = e[z - 3] - e[z - 11];
        e[z - 2] = e[z - 2] + e[z - 10];
        e[z - 3] = e[z - 3] + e[z - 11];
        e[z - 10] = (k00 + k11) * a2;
        e[z - 11] = (k11 - k00) * a2;

        k00 = e[z - 12] - e[z - 4];
        k11 = e[z - 5] - e[z - 13];
        e[z - 4] = e[z - 4] + e[z - 12];
        e[z - 5] = e[z - 5] + e[z - 13];
        e[z - 12] = k11;
        e[z - 13] = k00;

        k00 = e[z - 14] - e[z - 6];
        k11 = e[z - 7] - e[z - 15];
        e[z - 6] = e[z - 6] + e[z - 14];
        e[z - 7] = e[z - 7] + e[z - 15];
        e[z - 14] = (k00 + k11) * a2;
        e[z - 15] = (k00 - k11) * a2;

        iter_54(e + z);
        iter_54(e + z - 8);
        if (z <= stop) break;
        z -= 16;
    }
}

static void inverse_mdct_scratch(const lwo_tables *t, float *buffer, float *buf2)
{
    const int n = t->n, ld = t->bs;
    const int n2 = n >> 1, n4 = n >> 2, n8 = n >> 3;
    const float *a = t->a, *b = t->b, *c = t->c;

    /* imdct.rs:337-371 "copy and reflect spectral data" + step 0 */
    {
        long d = n2 - 2, ao = 0;
        for (long e = 0; e != n2; e += 4) {
            buf2[d + 1] = buffer[e] * a[ao] - buffer[e + 2] * a[ao + 1];
            buf2[d] = buffer[e] * a[ao + 1] + buffer[e + 2] * a[ao];
            d -= 2;
            ao += 2;
        }
        for (long e = n2 - 3;; e -= 4) {
            buf2[d + 1] = (-buffer[e + 2]) * a[ao] - (-buffer[e]) * a[ao + 1];
            buf2[d] = (-buffer[e + 2]) * a[ao + 1] + (-buffer[e]) * a[ao];
            if (d < 2) break;
            d -= 2;
            ao += 2;
        }
    }

    float *u = buffer, *v = buf2;
    /* imdct.rs:385-430 step 2 */
    {
        long ao = n2 - 8, d0 = n4, d1 = 0, e0 = n4, e1 = 0;
        for (;;) {
            float v41_21 = v[e0 + 1] - v[e1 + 1];
            float v40_20 = v[e0] - v[e1];
            u[d0 + 1] = v[e0 + 1] + v[e1 + 1];
            u[d0] = v[e0] + v[e1];
            u[d1 + 1] = v41_21 * a[ao + 4] - v40_20 * a[ao + 5];
            u[d1] = v40_20 * a[ao + 4] + v41_21 * a[ao + 5];

            v41_21 = v[e0 + 3] - v[e1 + 3];
            v40_20 = v[e0 + 2] - v[e1 + 2];
            u[d0 + 3] = v[e0 + 3] + v[e1 + 3];
            u[d0 + 2] = v[e0 + 2] + v[e1 + 2];
            u[d1 + 3] = v41_21 * a[ao] - v40_20 * a[ao + 1];
            u[d1 + 2] = v40_20 * a[ao] + v41_21 * a[ao + 1];

            if (ao < 8) break;
            ao -= 8;
            d0 += 4; d1 += 4; e0 += 4; e1 += 4;
        }
    }

    /* imdct.rs:445-484 step 3, literal schedule (incl. the bs 6/7 behaviour) */
    step3_iter0(n >> 4, u, n2 - 1 - n4 * 0, -(long)(n >> 3), a);
    step3_iter0(n >> 4, u, n2 - 1 - n4 * 1, -(long)(n >> 3), a);

    for (int i = 0; i < 4; i++)
        step3_r(n >> 5, u, n2 - 1 - n8 * i, -(long)(n >> 4), a, 16);

    int l = 2;
    for (; l < ((ld - 3) >> 1); l++) {
        int k0 = n >> (l + 2), k0_2 = k0 >> 1;
        int lim = 1 << (l + 1);
        for (int i = 0; i < lim; i++)
            step3_r(n >> (l + 4), u, n2 - 1 - (long)k0 * i, -(long)k0_2, a, 1 << (l + 3));
    }
    for (l = (ld - 3) >> 1; l < ld - 6; l++) {
        int k0 = n >> (l + 2), k1 = 1 << (l + 3), k0_2 = k0 >> 1;
        int rlim = n >> (l + 6), lim = 1 << (l + 1);
        long i_off = n2 - 1;
        long a_off = 0;
        for (int r = 0; r < rlim; r++) {
            step3_s(lim, u, i_off, -(long)k0_2, a + a_off, k1, k0);
            a_off += k1 * 4;
            i_off -= 8;
        }
    }
    step3_ld654(n >> 5, u, n2 - 1, a, n);

    /* imdct.rs:490-528 steps 4,5,6: bit-reverse shuffle u -> v */
    {
        long d0 = n4 - 4, d1 = n2 - 4;
        const uint32_t *br = t->bitrev;
        for (;;) {
            long k4 = br[0];
            v[d1 + 3] = u[k4 + 0];
            v[d1 + 2] = u[k4 + 1];
            v[d0 + 3] = u[k4 + 2];
            v[d0 + 2] = u[k4 + 3];
            k4 = br[1];
            v[d1 + 1] = u[k4 + 0];
            v[d1 + 0] = u[k4 + 1];
            v[d0 + 1] = u[k4 + 2];
            v[d0 + 0] = u[k4 + 3];
            if (d0 < 4) break;
            d0 -= 4; d1 -= 4; br += 2;
        }
    }

    /* imdct.rs:533-580 step 7 */
    {
        long co = 0, d = 0, e = n2 - 4;
        while (d < e) {
            float a02 = v[d] - v[e + 2];
            float a11 = v[d + 1] + v[e + 3];
            float b0 = c[co + 1] * a02 + c[co] * a11;
            float b1 = c[co + 1] * a11 - c[co] * a02;
            float b2 = v[d] + v[e + 2];
            float b3 = v[d + 1] - v[e + 3];
            v[d] = b2 + b0;
            v[d + 1] = b3 + b1;
            v[e + 2] = b2 - b0;
            v[e + 3] = b1 - b3;

            a02 = v[d + 2] - v[e];
            a11 = v[d + 3] + v[e + 1];
            b0 = c[co + 3] * a02 + c[co + 2] * a11;
            b1 = c[co + 3] * a11 - c[co + 2] * a02;
            b2 = v[d + 2] + v[e];
            b3 = v[d + 3] - v[e + 1];
            v[d + 2] = b2 + b0;
            v[d + 3] = b3 + b1;
            v[e] = b2 - b0;
            v[e + 1] = b1 - b3;

            co += 4; d += 4; e -= 4;
        }
    }

    /* imdct.rs:589-658 step 8 + decode */
    {
        long d0 = 0, d1 = n2 - 4, d2 = n2, d3 = n - 4;
        long bo = n2 - 8, e = n2 - 8;
        for (;;) {
            for (int q = 0; q < 4; q++) {
                /* q = 0 uses pair 6/7, q = 1 pair 4/5, ... (imdct.rs:619-649) */
                int p = 6 - 2 * q;
                float p_odd = buf2[e + p] * b[bo + p + 1] - buf2[e + p + 1] * b[bo + p];
                float p_even = (-buf2[e + p]) * b[bo + p] - buf2[e + p + 1] * b[bo + p + 1];
                buffer[d0 + q] = p_odd;
                buffer[d1 + 3 - q] = -p_odd;
                buffer[d2 + q] = p_even;
                buffer[d3 + 3 - q] = p_even;
            }
            if (e < 8) break;
            e -= 8; bo -= 8;
            d0 += 4; d2 += 4; d1 -= 4; d3 -= 4;
        }
    }
}

void lwo_inverse_mdct_with_scratch(const lwo_tables *t, float *buffer, float *buf2)
{
    inverse_mdct_scratch(t, buffer, buf2);
}

void lwo_inverse_mdct(const lwo_tables *t, float *buffer)
{
    /* imdct.rs:302: the reference heap-allocates this per call */
    float *buf2 = (float *)calloc((size_t)(t->n >> 1), sizeof(float));
    inverse_mdct_scratch(t, buffer, buf2);
    free(buf2);
}

/* audio.rs:792-806 dct_iv_slow (f32, like the reference) */
static void dct_iv_slow(float *buffer, int n)
{
    float *x = (float *)malloc(sizeof(float) * n);
    memcpy(x, buffer, sizeof(float) * n);
    unsigned nmask = ((unsigned)n << 3) - 1;
    float *mcos = (float *)malloc(sizeof(float) * 8 * n);
    const float frac_pi_4 = 0.785398163397448309615660845819875721f;
    for (int i = 0; i < 8 * n; i++) mcos[i] = cosf(frac_pi_4 * (float)i / (float)n);
    for (int i = 0; i < n; i++) {
        float acc = 0.f;
        for (int j = 0; j < n; j++)
            acc += x[j] * mcos[((unsigned)(2 * i + 1) * (unsigned)(2 * j + 1)) & nmask];
        buffer[i] = acc;
    }
    free(x); free(mcos);
}

/* audio.rs:808-825 inverse_mdct_slow */
void lwo_inverse_mdct_slow(float *buffer, int n)
{
    int n4 = n >> 2, n2 = n >> 1, n3_4 = n - n4;
    float *temp = (float *)malloc(sizeof(float) * n2);
    memcpy(temp, buffer, sizeof(float) * n2);
    dct_iv_slow(temp, n2);
    for (int i = 0; i < n4; i++) buffer[i] = temp[i + n4];
    for (int i = n4; i < n3_4; i++) buffer[i] = -temp[n3_4 - i - 1];
    for (int i = n3_4; i < n; i++) buffer[i] = -temp[i - n3_4];
    free(temp);
}

void lwo_inverse_mdct_f64(const float *spectrum, double *out, int n)
{
    int n4 = n >> 2, n2 = n >> 1, n3_4 = n - n4;
    double *temp = (double *)malloc(sizeof(double) * n2);
    const double pi = 3.14159265358979323846264338327950288;
    for (int i = 0; i < n2; i++) {
        double acc = 0.0;
        for (int j = 0; j < n2; j++)
            acc += (double)spectrum[j] * cos(pi / 4.0 * (double)(2 * i + 1) * (double)(2 * j + 1) / (double)n2);
        temp[i] = acc;
    }
    for (int i = 0; i < n4; i++) out[i] = temp[i + n4];
    for (int i = n4; i < n3_4; i++) out[i] = -temp[n3_4 - i - 1];
    for (int i = n3_4; i < n; i++) out[i] = -temp[i - n3_4];
    free(temp);
}

/* ------------------------------------------------------------------------- */
/* floor type 1 -- audio.rs:253-555                                           */
/* ------------------------------------------------------------------------- */

/* audio.rs:253-283 extr_neighbor: first index in v[..x] whose value is the
 * extreme one among those on the wanted side of v[x]. */
static int extr_neighbor(const uint32_t *v, int x, int want_low, int *idx, uint32_t *val)
{
    uint32_t bound = v[x];
    int best = -1;
    for (int i = 0; i < x; i++) {
        int ok = want_low ? (v[i] < bound) : (v[i] > bound);
        if (!ok) continue;
        if (best < 0) best = i;
        else if (want_low ? (v[i] > v[best]) : (v[i] < v[best])) best = i;
    }
    if (best < 0) return -1;  /* the reference panics here */
    *idx = best;
    *val = v[best];
    return 0;
}
int lwo_low_neighbor(const uint32_t *v, int x, int *idx, uint32_t *val)
{
    return extr_neighbor(v, x, 1, idx, val);
}
int lwo_high_neighbor(const uint32_t *v, int x, int *idx, uint32_t *val)
{
    return extr_neighbor(v, x, 0, idx, val);
}

/* audio.rs:354-367 (u32/i32 wrapping like a release build) */
uint32_t lwo_render_point(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t x)
{
    int32_t dy = (int32_t)(y1 - y0);
    uint32_t adx = x1 - x0;
    uint32_t ady = (uint32_t)(dy < 0 ? -dy : dy);
    uint32_t err = ady * (x - x0);
    uint32_t off = err / adx;
    return dy < 0 ? y0 - off : y0 + off;
}

void lwo_floor1_sort(lwo_floor1 *fl)
{
    for (int i = 0; i < fl->nposts; i++) fl->sorted_idx[i] = i;
    /* stable insertion sort by x (header.rs:887-889) */
    for (int i = 1; i < fl->nposts; i++) {
        int k = fl->sorted_idx[i], j = i - 1;
        while (j >= 0 && fl->x_list[fl->sorted_idx[j]] > fl->x_list[k]) {
            fl->sorted_idx[j + 1] = fl->sorted_idx[j];
            j--;
        }
        fl->sorted_idx[j + 1] = k;
    }
}

/* audio.rs:391-435 */
int lwo_floor1_amplitude(const lwo_floor1 *fl, const uint32_t *floor1_y,
                         uint32_t *final_y, uint8_t *step2)
{
    static const int ranges[4] = {256, 128, 86, 64};
    int32_t range = ranges[fl->multiplier - 1];
    step2[0] = 1; step2[1] = 1;
    final_y[0] = floor1_y[0];
    final_y[1] = floor1_y[1];
    for (int i = 2; i < fl->nposts; i++) {
        int li, hi; uint32_t lx, hx;
        if (lwo_low_neighbor(fl->x_list, i, &li, &lx)) return -1;
        if (lwo_high_neighbor(fl->x_list, i, &hi, &hx)) return -1;
        int32_t predicted = (int32_t)lwo_render_point(lx, final_y[li], hx, final_y[hi], fl->x_list[i]);
        int32_t val = (int32_t)floor1_y[i];
        int32_t highroom = range - predicted;
        int32_t lowroom = predicted;
        int32_t room = (highroom < lowroom ? highroom : lowroom) * 2;
        if (val > 0) {
            step2[li] = 1;
            step2[hi] = 1;
            step2[i] = 1;
            int32_t r;
            if (val >= room) {
                if (highroom > lowroom) r = predicted + val - lowroom;
                else r = predicted - val + highroom - 1;
            } else {
                /* audio.rs:422-423: `if val % 2 == 1 { -val - 1 } else { val } >> 1`
                 * -- the shift applies to the whole if-expression (arithmetic on i32) */
                int32_t t = (val % 2 == 1) ? (-val - 1) : val;
                r = predicted + (t >> 1);
            }
            final_y[i] = (uint32_t)r;
        } else {
            final_y[i] = (uint32_t)predicted;
            step2[i] = 0;
        }
    }
    for (int i = 0; i < fl->nposts; i++)
        if (final_y[i] > (uint32_t)range - 1) final_y[i] = (uint32_t)range - 1;
    return 0;
}

static const float INVERSE_DB_TABLE[256] = {
#include "floor1_inverse_db.inc"
};
const float *lwo_inverse_db_table(void) { return INVERSE_DB_TABLE; }

/* audio.rs:503-524 render_line, appending to out (bounded by cap) */
static int render_line(int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                       uint32_t *out, int len, int cap)
{
    int32_t dy = y1 - y0;
    int32_t adx = x1 - x0;
    int32_t ady = dy < 0 ? -dy : dy;
    int32_t base = dy / adx;
    int32_t y = y0;
    int32_t err = 0;
    int32_t sy = base + (dy < 0 ? -1 : 1);
    ady = ady - (base < 0 ? -base : base) * adx;
    if (len < cap) out[len] = (uint32_t)y;
    len++;
    for (int32_t x = x0 + 1; x < x1; x++) {
        err += ady;
        if (err >= adx) { err -= adx; y += sy; }
        else y += base;
        if (len < cap) out[len] = (uint32_t)y;
        len++;
    }
    return len;
}

/* audio.rs:526-551, the integer curve before the table lookup */
int lwo_floor1_curve_y(const lwo_floor1 *fl, const uint32_t *final_y,
                       const uint8_t *step2, int n2, uint32_t *out_y)
{
    uint32_t mult = (uint32_t)fl->multiplier;
    uint32_t hx = 0, lx = 0, hy = 0;
    uint32_t ly = final_y[fl->sorted_idx[0]] * mult;
    int len = 0;
    /* the reference renders into a growing Vec and truncates to n2 afterwards;
     * writing only the first n2 entries is the same thing */
    for (int i = 1; i < fl->nposts; i++) {
        int si = fl->sorted_idx[i];
        if (step2[si]) {
            hy = final_y[si] * mult;
            hx = fl->x_list[si];
            len = render_line((int32_t)lx, (int32_t)ly, (int32_t)hx, (int32_t)hy, out_y, len, n2);
            lx = hx;
            ly = hy;
        }
    }
    if (hx < (uint32_t)n2)
        len = render_line((int32_t)hx, (int32_t)hy, n2, (int32_t)hy, out_y, len, n2);
    return len >= n2 ? 0 : -1;
}

int lwo_floor1_synthesis(const lwo_floor1 *fl, const uint32_t *final_y,
                         const uint8_t *step2, int n2, float *out)
{
    uint32_t *y = (uint32_t *)malloc(sizeof(uint32_t) * n2);
    int rc = lwo_floor1_curve_y(fl, final_y, step2, n2, y);
    if (rc == 0)
        for (int i = 0; i < n2; i++) out[i] = INVERSE_DB_TABLE[y[i] & 255u];
    free(y);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* audio.rs:762-777                                                           */
/* ------------------------------------------------------------------------- */
void lwo_inverse_couple(float *mag, float *ang, int len)
{
    for (int i = 0; i < len; i++) {
        float m = mag[i], a = ang[i];
        float nm, na;
        if (m > 0.f) {
            if (a > 0.f) { nm = m; na = m - a; }
            else { nm = m + a; na = m; }
        } else {
            if (a > 0.f) { nm = m; na = m + a; }
            else { nm = m - a; na = m; }
        }
        mag[i] = nm;
        ang[i] = na;
    }
}

/* samples.rs:92-103 (Rust `as i16`: truncate toward zero, NaN -> 0) */
int16_t lwo_sample_i16(float v)
{
    float fl = v * 32768.0f;
    if (fl > 32767.f) return 32767;
    if (fl < -32768.f) return -32768;
    if (fl != fl) return 0;
    return (int16_t)fl;
}

/* audio.rs:1056-1073 */
void lwo_window_geometry(int bs0, int bs1, int blockflag, int prev_flag, int next_flag,
                         lwo_window_geom *g)
{
    int bs = blockflag ? bs1 : bs0;
    int n = 1 << bs;
    int n0 = 1 << bs0;
    int wc = n >> 1;
    g->n = n;
    /* short blocks carry no flags: map_or(true, ..) */
    int prev = blockflag ? prev_flag : 1;
    int next = blockflag ? next_flag : 1;
    if (prev) { g->left_start = 0; g->left_end = wc; g->left_use_bs1 = blockflag; }
    else { g->left_start = (n - n0) >> 2; g->left_end = (n + n0) >> 2; g->left_use_bs1 = 0; }
    if (next) { g->right_start = wc; g->right_end = n; }
    else { g->right_start = (n * 3 - n0) >> 2; g->right_end = (n * 3 + n0) >> 2; }
}

/* ------------------------------------------------------------------------- */
/* PreviousWindowRight, audio.rs:847-861                                      */
/* ------------------------------------------------------------------------- */
lwo_pwr *lwo_pwr_new(int channels, int cap)
{
    lwo_pwr *p = (lwo_pwr *)calloc(1, sizeof(*p));
    p->channels = channels;
    p->cap = cap;
    p->data = (float *)calloc((size_t)channels * cap, sizeof(float));
    return p;
}
void lwo_pwr_reset(lwo_pwr *p) { p->has = 0; p->len = 0; }
void lwo_pwr_free(lwo_pwr *p) { if (p) { free(p->data); free(p); } }
int lwo_pwr_has(const lwo_pwr *p) { return p->has; }
int lwo_pwr_len(const lwo_pwr *p) { return p->len; }
float *lwo_pwr_data(lwo_pwr *p, int ch) { return p->data + (size_t)ch * p->cap; }
void lwo_pwr_set(lwo_pwr *p, int len) { p->has = 1; p->len = len; }

/* audio.rs:1079-1154: window/overlap-add/slice/state for all channels.
 * x = [channels][n] IMDCT output (modified), out receives [channels][n] rows. */
static int overlap_add(const lwo_tables *t0, const lwo_tables *t1, int channels,
                       int blockflag, int prev_flag, int next_flag,
                       float *x, lwo_pwr *pwr, float *out, int *out_len)
{
    lwo_window_geom g;
    lwo_window_geometry(t0->bs, t1->bs, blockflag, prev_flag, next_flag, &g);
    int n = g.n;
    int keep = g.right_end - g.right_start;
    if (keep > pwr->cap) return 2;
    if (pwr->has) {
        if (pwr->channels != channels) return 3;         /* audio.rs:1086 assert */
        const lwo_tables *tw = g.left_use_bs1 ? t1 : t0;
        int slope_len = tw->n >> 1;
        int plen = pwr->len;
        /* audio.rs:1083 `pwr.data.take()` has already emptied the state when the
         * guard at audio.rs:1107-1111 fires, so the error leaves it empty */
        if (slope_len < plen) { pwr->has = 0; pwr->len = 0; return 1; }
        /* a prev longer than the block would index out of range in the reference (panic) */
        if (g.left_start + plen > n) return 3;
        const float *w = tw->window;
        float *tmp = (float *)malloc(sizeof(float) * keep);
        for (int ch = 0; ch < channels; ch++) {
            float *xc = x + (size_t)ch * n;
            float *prev = lwo_pwr_data(pwr, ch);
            for (int i = 0; i < plen; i++)
                xc[g.left_start + i] = (xc[g.left_start + i] * w[i]) + (prev[i] * w[plen - 1 - i]);
            memcpy(tmp, xc + g.right_start, sizeof(float) * keep);   /* future prev half */
            int olen = g.right_start - g.left_start;
            memcpy(out + (size_t)ch * n, xc + g.left_start, sizeof(float) * olen);
            memcpy(prev, tmp, sizeof(float) * keep);
        }
        free(tmp);
        *out_len = g.right_start - g.left_start;
    } else {
        for (int ch = 0; ch < channels; ch++)
            memcpy(lwo_pwr_data(pwr, ch), x + (size_t)ch * n + g.right_start, sizeof(float) * keep);
        *out_len = 0;                                     /* audio.rs:1140-1151 */
    }
    pwr->has = 1;
    pwr->len = keep;
    pwr->channels = channels;
    return 0;
}

int lwo_synth_spectrum(const lwo_tables *t0, const lwo_tables *t1, int channels,
                       int blockflag, int prev_flag, int next_flag,
                       const float *spectrum, lwo_pwr *pwr, float *out, int *out_len)
{
    const lwo_tables *t = blockflag ? t1 : t0;
    int n = t->n, n2 = n >> 1;
    float *x = (float *)calloc((size_t)channels * n, sizeof(float));
    float *scratch = (float *)malloc(sizeof(float) * n2);
    for (int ch = 0; ch < channels; ch++) {
        /* audio.rs:1044-1051: spectrum extended with n/2 zeros, then inverse_mdct */
        memcpy(x + (size_t)ch * n, spectrum + (size_t)ch * n2, sizeof(float) * n2);
        inverse_mdct_scratch(t, x + (size_t)ch * n, scratch);
    }
    int rc = overlap_add(t0, t1, channels, blockflag, prev_flag, next_flag, x, pwr, out, out_len);
    free(scratch);
    free(x);
    return rc;
}

int lwo_synth_packet(const lwo_tables *t0, const lwo_tables *t1, int channels,
                     int blockflag, int prev_flag, int next_flag,
                     int n_coupling, const uint8_t *mag, const uint8_t *ang,
                     lwo_channel_in *chans, lwo_pwr *pwr, float *out, int *out_len)
{
    const lwo_tables *t = blockflag ? t1 : t0;
    int n = t->n, n2 = n >> 1;
    /* audio.rs:991-1002 inverse coupling, steps in reverse order */
    for (int s = n_coupling - 1; s >= 0; s--) {
        if (mag[s] == ang[s]) return 3;                   /* dual_mut_idx assert, audio.rs:783 */
        lwo_inverse_couple(chans[mag[s]].residue, chans[ang[s]].residue, n2);
    }
    /* audio.rs:1006-1039 floor curve, times residue */
    float *spec = (float *)malloc(sizeof(float) * (size_t)channels * n2);
    int rc = 0;
    for (int ch = 0; ch < channels && rc == 0; ch++) {
        float *fl = spec + (size_t)ch * n2;
        switch (chans[ch].floor_kind) {
        case LWO_FLOOR_ONE: {
            uint32_t final_y[LWO_MAX_POSTS];
            uint8_t step2[LWO_MAX_POSTS];
            if (lwo_floor1_amplitude(chans[ch].fl, chans[ch].floor1_y, final_y, step2)) { rc = 3; break; }
            if (lwo_floor1_synthesis(chans[ch].fl, final_y, step2, n2, fl)) { rc = 3; break; }
            break;
        }
        case LWO_FLOOR_DENSE:
            memcpy(fl, chans[ch].dense_floor, sizeof(float) * n2);
            break;
        default:
            for (int i = 0; i < n2; i++) fl[i] = 0.f;      /* audio.rs:1021-1024 */
        }
        if (rc) break;
        for (int i = 0; i < n2; i++) fl[i] = fl[i] * chans[ch].residue[i];
    }
    if (rc == 0)
        rc = lwo_synth_spectrum(t0, t1, channels, blockflag, prev_flag, next_flag, spec, pwr, out, out_len);
    free(spec);
    return rc;
}
