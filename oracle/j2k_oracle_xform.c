/*
 * oracle/j2k_oracle_xform.c -- TEST INFRASTRUCTURE ONLY (see j2k_oracle.h).
 *
 * Restatement of SURVEY.md section 8(a) rows W1-W4 (inverse DWT, libavcodec/jpeg2000dwt.c:309-537,
 * :539-620) and M1-M3 (inverse component transforms, libavcodec/jpeg2000dsp.c:43-91).
 *
 * Border handling: the reference extends each line by sequential copies (jpeg2000dwt.c:49-74); here the
 * out-of-range neighbours are fetched by index reflection (whole-sample symmetric, period 2(n-1)),
 * which SURVEY.md Appendix D showed to be bit-identical, and which is also how the CUDA kernels do it.
 * Floating point: every step is written as separate add / multiply / subtract on `float` and the file is
 * built with -ffp-contract=off, matching the reference's strict-IEEE, no-FMA build.
 */
#include <stdlib.h>
#include <string.h>

#include "j2k_oracle.h"

#define ALPHA 1.586134342059924f
#define BETA  0.052980118572961f
#define GAMMA 0.882911075530934f
#define DELTA 0.443506852043971f
#define KGAIN 1.230174104914001f
#define XGAIN 0.812893066115961f

/* the reference's lines are filled by sequential copies: reading them back is equivalent to
 * reflecting the index into [i0, i1) */
static inline int reflect(int i, int i0, int i1)
{
    int n = i1 - i0;
    if (n == 1)
        return i0;
    int period = 2 * (n - 1);
    int k = (i - i0) % period;
    if (k < 0)
        k += period;
    return i0 + (k < n ? k : period - k);
}

/* ---- 5/3 (jpeg2000dwt.c:309-325): p[] holds samples i0..i1-1 at p[i0..], plus 2 on each side ------ */
static void lift53(int32_t *p, int i0, int i1)
{
    if (i1 <= i0 + 1) {
        if (i0 == 1)
            p[1] >>= 1;
        return;
    }
    for (int k = 1; k <= 2; k++) {
        p[i0 - k] = p[reflect(i0 - k, i0, i1)];
        p[i1 - 1 + k] = p[reflect(i1 - 1 + k, i0, i1)];
    }
    uint32_t *u = (uint32_t *)p;
    for (int i = i0 >> 1; i < (i1 >> 1) + 1; i++)
        u[2 * i] -= (uint32_t)((int32_t)(u[2 * i - 1] + u[2 * i + 1] + 2) >> 2);
    for (int i = i0 >> 1; i < i1 >> 1; i++)
        u[2 * i + 1] += (uint32_t)((int32_t)(u[2 * i] + u[2 * i + 2]) >> 1);
}

/* ---- 9/7 float (jpeg2000dwt.c:376-401) -------------------------------------------------------------- */
static void lift97f(float *p, int i0, int i1)
{
    if (i1 <= i0 + 1) {
        if (i0 == 1)
            p[1] *= KGAIN / 2;
        else
            p[0] *= XGAIN;
        return;
    }
    for (int k = 1; k <= 4; k++) {
        p[i0 - k] = p[reflect(i0 - k, i0, i1)];
        p[i1 - 1 + k] = p[reflect(i1 - 1 + k, i0, i1)];
    }
    for (int i = (i0 >> 1) - 1; i < (i1 >> 1) + 2; i++)
        p[2 * i] -= DELTA * (p[2 * i - 1] + p[2 * i + 1]);
    for (int i = (i0 >> 1) - 1; i < (i1 >> 1) + 1; i++)
        p[2 * i + 1] -= GAMMA * (p[2 * i] + p[2 * i + 2]);
    for (int i = i0 >> 1; i < (i1 >> 1) + 1; i++)
        p[2 * i] += BETA * (p[2 * i - 1] + p[2 * i + 1]);
    for (int i = i0 >> 1; i < i1 >> 1; i++)
        p[2 * i + 1] += ALPHA * (p[2 * i] + p[2 * i + 2]);
}

/* ---- 9/7 16.16 fixed point (jpeg2000dwt.c:453-481) -------------------------------------------------- */
static void lift97i(int32_t *p, int i0, int i1)
{
    if (i1 <= i0 + 1) {
        if (i0 == 1)
            p[1] = (int32_t)((p[1] * 80621ll + (1 << 16)) >> 17);
        else
            p[0] = (int32_t)((p[0] * 53274ll + (1 << 15)) >> 16);
        return;
    }
    for (int k = 1; k <= 4; k++) {
        p[i0 - k] = p[reflect(i0 - k, i0, i1)];
        p[i1 - 1 + k] = p[reflect(i1 - 1 + k, i0, i1)];
    }
    for (int i = (i0 >> 1) - 1; i < (i1 >> 1) + 2; i++)
        p[2 * i] -= (int32_t)((29066ll * (p[2 * i - 1] + (int64_t)p[2 * i + 1]) + (1 << 15)) >> 16);
    for (int i = (i0 >> 1) - 1; i < (i1 >> 1) + 1; i++)
        p[2 * i + 1] -= (int32_t)((57862ll * (p[2 * i] + (int64_t)p[2 * i + 2]) + (1 << 15)) >> 16);
    for (int i = i0 >> 1; i < (i1 >> 1) + 1; i++)
        p[2 * i] += (int32_t)((3472ll * (p[2 * i - 1] + (int64_t)p[2 * i + 1]) + (1 << 15)) >> 16);
    for (int i = i0 >> 1; i < i1 >> 1; i++) {
        int64_t sum = p[2 * i] + (int64_t)p[2 * i + 2];
        p[2 * i + 1] += (int32_t)sum;
        p[2 * i + 1] += (int32_t)((38413ll * sum + (1 << 15)) >> 16);
    }
}

/*
 * One resolution step in place on a Mallat-layout plane of row stride W: rows first, then columns
 * (jpeg2000dwt.c:327-374).  Inside a line the low-pass samples come first, then the high-pass ones;
 * interleaving puts low-pass sample k at odd/even position according to the parity of the origin.
 */
static void level_pass(void *plane, int W, int lh, int lv, int mh, int mv, int type, void *linebuf)
{
    int32_t *pi = plane, *li = (int32_t *)linebuf + 8;
    float *pf = plane, *lf = (float *)linebuf + 8;
    for (int pass = 0; pass < 2; pass++) {
        int nlines = pass == 0 ? lv : lh, len = pass == 0 ? lh : lv, m = pass == 0 ? mh : mv;
        size_t step = pass == 0 ? 1 : (size_t)W, lstep = pass == 0 ? (size_t)W : 1;
        for (int l = 0; l < nlines; l++) {
            size_t base = l * lstep;
            int j = 0;
            if (type == ORACLE_DWT97) {
                for (int i = m; i < len; i += 2, j++) lf[m + i] = pf[base + j * step];
                for (int i = 1 - m; i < len; i += 2, j++) lf[m + i] = pf[base + j * step];
                lift97f(lf, m, m + len);
                for (int i = 0; i < len; i++) pf[base + i * step] = lf[m + i];
            } else {
                for (int i = m; i < len; i += 2, j++) li[m + i] = pi[base + j * step];
                for (int i = 1 - m; i < len; i += 2, j++) li[m + i] = pi[base + j * step];
                if (type == ORACLE_DWT53)
                    lift53(li, m, m + len);
                else
                    lift97i(li, m, m + len);
                for (int i = 0; i < len; i++) pi[base + i * step] = li[m + i];
            }
        }
    }
}

/* W1: per-level sizes and origin parities by repeated halving of the full-resolution bounds
 * (jpeg2000dwt.c:539-560); then coarsest level first (:601-620). */
int oracle_idwt(int type, const int *border, int levels, void *data)
{
    if (levels == 0)
        return 0;
    if (levels > 32 || (type != ORACLE_DWT97 && type != ORACLE_DWT53 && type != ORACLE_DWT97_INT))
        return -1;
    int b[2][2] = { { border[0], border[1] }, { border[2], border[3] } };
    int len[33][2] = { { 0 } }, mod[33][2] = { { 0 } };
    int maxlen = b[0][1] - b[0][0] > b[1][1] - b[1][0] ? b[0][1] - b[0][0] : b[1][1] - b[1][0];
    for (int lev = levels - 1; lev >= 0; lev--)
        for (int i = 0; i < 2; i++) {
            len[lev][i] = b[i][1] - b[i][0];
            mod[lev][i] = b[i][0] & 1;
            b[i][0] = (b[i][0] + 1) >> 1;
            b[i][1] = (b[i][1] + 1) >> 1;
        }
    const int W = len[levels - 1][0], H = len[levels - 1][1];
    void *linebuf = malloc(sizeof(int32_t) * ((size_t)maxlen + 32));
    if (!linebuf)
        return ORACLE_ERR_NOMEM;
    for (int lev = 0; lev < levels; lev++)
        level_pass(data, W, len[lev][0], len[lev][1], mod[lev][0], mod[lev][1], type, linebuf);
    if (type == ORACLE_DWT97_INT) {
        int32_t *p = data;
        for (size_t i = 0; i < (size_t)W * H; i++)
            p[i] = (int32_t)(p[i] + 128ll) >> 8;     /* undo I_PRESHIFT (:534-536) */
    }
    free(linebuf);
    return 0;
}

/* M1-M3 */
int oracle_mct(int type, void *p0, void *p1, void *p2, int n)
{
    if (type == ORACLE_DWT53) {                      /* rct_int, jpeg2000dsp.c:78-91 */
        uint32_t *y = p0, *cb = p1, *cr = p2;
        for (int i = 0; i < n; i++) {
            uint32_t g = y[i] - (uint32_t)((int32_t)(cr[i] + cb[i]) >> 2);
            uint32_t r = g + cr[i], b = g + cb[i];
            y[i] = r; cb[i] = g; cr[i] = b;
        }
    } else if (type == ORACLE_DWT97) {               /* ict_float, :43-59 */
        float *y = p0, *cb = p1, *cr = p2;
        for (int i = 0; i < n; i++) {
            float r = y[i] + 1.402f * cr[i];
            float g = y[i] - 0.34413f * cb[i];
            g = g - 0.71414f * cr[i];
            float b = y[i] + 1.772f * cb[i];
            y[i] = r; cb[i] = g; cr[i] = b;
        }
    } else if (type == ORACLE_DWT97_INT) {           /* ict_int, :61-76 */
        int32_t *y = p0, *cb = p1, *cr = p2;
        for (int i = 0; i < n; i++) {
            int32_t r = y[i] + cr[i] + ((int32_t)(26345u * (uint32_t)cr[i] + (1 << 15)) >> 16);
            int32_t g = y[i] - ((int32_t)(22553u * (uint32_t)cb[i] + (1 << 15)) >> 16)
                             - ((int32_t)(46802u * (uint32_t)cr[i] + (1 << 15)) >> 16);
            int32_t b = y[i] + 2 * cb[i] + ((int32_t)((uint32_t)-14942 * (uint32_t)cb[i] + (1 << 15)) >> 16);
            y[i] = r; cb[i] = g; cr[i] = b;
        }
    } else {
        return -1;
    }
    return 0;
}
