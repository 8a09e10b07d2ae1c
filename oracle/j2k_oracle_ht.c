/*
 * oracle/j2k_oracle_ht.c -- TEST INFRASTRUCTURE ONLY (see j2k_oracle.h).
 *
 * Restatement of the reference's HT block decoder, SURVEY.md section 8(a) rows H1-H8:
 *   H1 orchestration        libavcodec/jpeg2000htdec.c:1188-1336
 *   H2 the four bit readers :123-131, :145-221, :283-295, :429-460
 *   H3 MEL                  :462-495
 *   H4 CxtVLC               :300-331, :510-531 (tables: ffmpeg-ht_b200/csrc/cxtvlc_tables.h, generated from the reference's)
 *   H5 U-VLC                :338-388, combination rules :671-712 and :833-854
 *   H6 cleanup pass         :547-1014
 *   H7 SigProp              :1016-1131
 *   H8 MagRef               :1136-1185
 * The reference's behaviour on damaged segments (refill granularity of the backward reader, what the
 * readers return past the end of their data) is restated too, so that random-byte fuzzing against
 * oracle/_ref stays bit-exact.
 */
#include <stdlib.h>
#include <string.h>

#include "cxtvlc_tables.h"
#include "j2k_oracle.h"

/* ---- H2: forward LSB-first reader with "7 bits after 0xFF" (MagSgn) ------------------------------- */
typedef struct FwdBits {
    const uint8_t *d;
    int len, pos, n;
    unsigned last;
    uint64_t acc;
} FwdBits;

static void fwd_fill(FwdBits *s)
{
    while (s->n < 32) {
        unsigned byte = 0xFF;                       /* past the end: 0xFF, `last` is not updated (:207-221) */
        int nb = s->last == 0xFF ? 7 : 8;
        if (s->pos < s->len) {
            byte = s->d[s->pos++];
            s->last = byte;
        }
        s->acc |= (uint64_t)byte << s->n;           /* OR: bit 7 of a 7-bit byte overlaps the next byte */
        s->n += nb;
    }
}

static uint32_t fwd_get(FwdBits *s, int nb)
{
    if (s->n <= nb)
        fwd_fill(s);
    uint32_t v = (uint32_t)(s->acc & ((1ull << nb) - 1));
    s->acc >>= nb;
    s->n -= nb;
    return v;
}

/* ---- H2: backward LSB-first reader (VLC, MagRef) ---------------------------------------------------
 * Bytes are pulled in groups of four from base[pos] downwards.  A byte whose low seven bits are all
 * ones gives only seven bits when the byte read before it (one address higher) is > 0x8F (:145-201).
 * When fewer than four bytes remain the group is zero-filled and pos is parked at 0, which makes a
 * later group start at base[0] again -- kept, it only matters for damaged segments. */
typedef struct BwdBits {
    const uint8_t *base;
    int pos, n;
    uint64_t acc;
} BwdBits;

static void bwd_fill(BwdBits *s)
{
    if (s->n >= 32)
        return;
    unsigned neighbour = s->base[s->pos + 1];
    unsigned g[4] = { 0, 0, 0, 0 };
    if (s->pos >= 3) {
        for (int k = 0; k < 4; k++)
            g[k] = s->base[s->pos - k];
        s->pos -= 4;
    } else {
        for (int k = 0; k <= s->pos; k++)
            g[k] = s->base[s->pos - k];
        s->pos = 0;
    }
    uint64_t bits = 0;
    int nb = 0;
    for (int k = 0; k < 4; k++) {
        int take = ((g[k] & 0x7F) == 0x7F && neighbour > 0x8F) ? 7 : 8;
        bits |= (uint64_t)(g[k] & ((1u << take) - 1)) << nb;
        nb += take;
        neighbour = g[k];
    }
    s->acc |= bits << s->n;
    s->n += nb;
}

static inline void bwd_drop(BwdBits *s, int nb) { s->acc >>= nb; s->n -= nb; }

static uint32_t bwd_get(BwdBits *s, int nb)
{
    if (s->n < nb)
        bwd_fill(s);
    uint32_t v = (uint32_t)(s->acc & ((1ull << nb) - 1));
    bwd_drop(s, nb);
    return v;
}

/* ---- H2/H3: MEL bit reader (MSB-first, 7 bits after 0xFF, 0xFF past the end) and run decoder ------ */
typedef struct Mel {
    const uint8_t *d;
    int len, pos, left;
    unsigned cur;
    int k, run, one;
} Mel;

static const uint8_t MEL_EXP[13] = { 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5 };

static int mel_bit(Mel *m)
{
    if (m->left == 0) {
        m->left = m->cur == 0xFF ? 7 : 8;
        if (m->pos < m->len)
            m->cur = m->d[m->pos++];
        else
            m->cur = 0xFF;
    }
    m->left--;
    return (m->cur >> m->left) & 1;
}

static int mel_symbol(Mel *m)
{
    if (m->run == 0 && !m->one) {
        int e = MEL_EXP[m->k];
        if (mel_bit(m)) {
            m->run = 1 << e;
            m->k = m->k < 12 ? m->k + 1 : 12;
        } else {
            int r = 0;
            while (e-- > 0)
                r = 2 * r + mel_bit(m);
            m->run = r & 0xFF;
            m->k = m->k > 0 ? m->k - 1 : 0;
            m->one = 1;
        }
    }
    if (m->run > 0) {
        m->run--;
        return 0;
    }
    m->one = 0;
    return 1;
}

/* ---- H5: U-VLC pieces ------------------------------------------------------------------------------ */
static int uvlc_prefix(BwdBits *v)
{
    static const uint8_t val[8] = { 5, 1, 2, 1, 3, 1, 2, 1 }, len[8] = { 3, 1, 2, 1, 3, 1, 2, 1 };
    if (v->n < 3)
        bwd_fill(v);
    int b = (int)(v->acc & 7);
    bwd_drop(v, len[b]);
    return val[b];
}

static int uvlc_suffix(BwdBits *v, int pfx)
{
    if (pfx < 3)
        return 0;
    if (v->n < 5)
        bwd_fill(v);
    int b = (int)(v->acc & 31);
    if (pfx == 3) {
        bwd_drop(v, 1);
        return b & 1;
    }
    bwd_drop(v, 5);
    return b;
}

static int uvlc_ext(BwdBits *v, int sfx) { return (int)bwd_get(v, sfx >= 28 ? 4 : 0); }

static int uvlc_whole(BwdBits *v)
{
    int p = uvlc_prefix(v);
    int s = uvlc_suffix(v, p);
    int e = uvlc_ext(v, s);
    return p + s + 4 * e;
}

/* ---- SigProp bit reader (:442-460) ---------------------------------------------------------------- */
typedef struct SpBits {
    const uint8_t *d;
    int len, pos, left;
    unsigned cur, last;
} SpBits;

static int sp_bit(SpBits *s)
{
    if (s->left == 0) {
        s->left = s->last == 0xFF ? 7 : 8;
        s->cur = s->pos < s->len ? s->d[s->pos++] : 0;
        s->last = s->cur;
    }
    int b = s->cur & 1;
    s->cur >>= 1;
    s->left--;
    return b;
}

static inline int clz32(uint32_t v) { return __builtin_clz(v); }

/* state bits kept per sample (same meaning as :62-65) */
#define ST_SIGMA 1
#define ST_REFIND 4
#define ST_NEWSIG 8
#define ST_SCANNED 16

typedef struct Quad {
    int rho, uoff, ek, e1;
} Quad;

/* H4: one quad's significance pattern and EMB patterns */
static void decode_quad(Mel *mel, BwdBits *vlc, const unsigned short *table, int ctx, Quad *q)
{
    q->rho = q->uoff = q->ek = q->e1 = 0;
    if (ctx == 0 && mel_symbol(mel) == 0)
        return;
    bwd_fill(vlc);
    unsigned e = table[(vlc->acc & 0x7F) | ((unsigned)ctx << 7)];
    q->uoff = (e >> 3) & 1;
    q->rho = (e >> 4) & 15;
    q->ek = (e >> 8) & 15;
    q->e1 = (e >> 12) & 15;
    bwd_drop(vlc, e & 7);
}

/* H6.  Returns 1 or a negative error. */
static int cleanup_pass(const uint8_t *Dcup, int Lcup, int Pcup, int pLSB, int maxbp, int w, int h,
                        int S, int32_t *samples, uint8_t *states)
{
    const int QW = (w + 1) >> 1, QH = (h + 1) >> 1;
    const int nq = QW * QH;
    if (maxbp >= 32)
        return ORACLE_ERR_INVALID;

    uint8_t *sig = calloc((size_t)4 * nq + 8, 1);
    uint8_t *E = calloc((size_t)4 * nq + 8, 1);
    uint32_t *mu = calloc((size_t)4 * nq + 8, sizeof(uint32_t));
    int ret = 1;

    FwdBits ms = { Dcup, Pcup, 0, 0, 0, 0 };
    fwd_fill(&ms);
    Mel mel = { Dcup, Lcup, Pcup, 0, 0, 0, 0, 0 };
    BwdBits vlc = { Dcup + Pcup, Lcup - 2 - Pcup, 0, 0 };
    bwd_fill(&vlc);
    bwd_drop(&vlc, 4);                               /* the Scup nibble (:283-295) */

    for (int qy = 0; qy < QH && ret > 0; qy++) {
        const unsigned short *table = qy == 0 ? CXTVLC_FIRST_ROW : CXTVLC_OTHER_ROWS;
        int ctx_run = 0;                             /* first row: context carried from the previous quad */
        for (int qx0 = 0; qx0 < QW; qx0 += 2) {
            const int np = QW - qx0 >= 2 ? 2 : 1;
            Quad qd[2];
            int u[2] = { 0, 0 }, kappa[2] = { 1, 1 }, U[2];
            for (int j = 0; j < np; j++) {
                const int qx = qx0 + j, q = qy * QW + qx;
                int ctx;
                if (qy == 0) {
                    ctx = ctx_run;                                            /* :646-649 */
                } else {
                    const uint8_t *n = sig + 4 * (q - QW);                    /* :788-796 */
                    ctx = n[1] + (n[3] << 2);
                    if (qx > 0) {
                        ctx |= n[-1];
                        ctx += (sig[4 * q - 1] | sig[4 * q - 2]) << 1;
                    }
                    if (qx < QW - 1)
                        ctx |= n[5] << 2;
                }
                decode_quad(&mel, &vlc, table, ctx, &qd[j]);
                for (int i = 0; i < 4; i++)
                    sig[4 * q + i] = (qd[j].rho >> i) & 1;
                ctx_run = (sig[4 * q] | sig[4 * q + 1]) + (sig[4 * q + 2] << 1) + (sig[4 * q + 3] << 2);
            }
            /* H5: residuals u */
            if (np == 2) {
                bwd_fill(&vlc);
                if (qd[0].uoff && qd[1].uoff) {
                    if (qy == 0) {                                            /* :671-704 */
                        if (mel_symbol(&mel)) {
                            int p0 = uvlc_prefix(&vlc), p1 = uvlc_prefix(&vlc);
                            int s0 = uvlc_suffix(&vlc, p0), s1 = uvlc_suffix(&vlc, p1);
                            int e0 = uvlc_ext(&vlc, s0), e1 = uvlc_ext(&vlc, s1);
                            u[0] = 2 + p0 + s0 + 4 * e0;
                            u[1] = 2 + p1 + s1 + 4 * e1;
                        } else {
                            int p0 = uvlc_prefix(&vlc);
                            if (p0 > 2) {
                                u[1] = (int)bwd_get(&vlc, 1) + 1;
                                int s0 = uvlc_suffix(&vlc, p0);
                                u[0] = p0 + s0 + 4 * uvlc_ext(&vlc, s0);
                            } else {
                                int p1 = uvlc_prefix(&vlc);
                                int s0 = uvlc_suffix(&vlc, p0), s1 = uvlc_suffix(&vlc, p1);
                                int e0 = uvlc_ext(&vlc, s0), e1 = uvlc_ext(&vlc, s1);
                                u[0] = p0 + s0 + 4 * e0;
                                u[1] = p1 + s1 + 4 * e1;
                            }
                        }
                    } else {                                                  /* :833-845 */
                        int p0 = uvlc_prefix(&vlc), p1 = uvlc_prefix(&vlc);
                        int s0 = uvlc_suffix(&vlc, p0), s1 = uvlc_suffix(&vlc, p1);
                        int e0 = uvlc_ext(&vlc, s0), e1 = uvlc_ext(&vlc, s1);
                        u[0] = p0 + s0 + 4 * e0;
                        u[1] = p1 + s1 + 4 * e1;
                    }
                } else if (qd[0].uoff) {
                    u[0] = uvlc_whole(&vlc);
                } else if (qd[1].uoff) {
                    u[1] = uvlc_whole(&vlc);
                }
            } else if (qd[0].uoff) {
                u[0] = uvlc_whole(&vlc);
            }
            /* exponent predictor kappa (:855-885); first row keeps kappa = 1 */
            for (int j = 0; j < np; j++) {
                const int qx = qx0 + j, q = qy * QW + qx;
                if (qy > 0) {
                    const uint8_t *en = E + 4 * (q - QW);
                    int gamma = (qd[j].rho & (qd[j].rho - 1)) != 0;           /* more than one significant sample */
                    int m = en[1] > en[3] ? en[1] : en[3];
                    if (qx > 0 && en[-1] > m)
                        m = en[-1];
                    if (qx < QW - 1 && en[5] > m)
                        m = en[5];
                    int k = gamma * (m - 1);
                    kappa[j] = k > 1 ? k : 1;
                }
                U[j] = kappa[j] + u[j];
            }
            for (int j = 0; j < np; j++)
                if (U[j] > maxbp)
                    ret = ORACLE_ERR_INVALID;                                 /* :715, :889 */
            if (ret < 0)
                break;
            /* MagSgn (:394-427) */
            for (int j = 0; j < np; j++) {
                const int q = qy * QW + qx0 + j;
                for (int i = 0; i < 4; i++) {
                    int m = sig[4 * q + i] * U[j] - ((qd[j].ek >> i) & 1);
                    uint32_t v = 0;
                    if (m > 0)
                        v = fwd_get(&ms, m) + ((uint32_t)((qd[j].e1 >> i) & 1) << m);
                    if (m != 0) {
                        E[4 * q + i] = (uint8_t)(32 - clz32(v | 1));
                        uint32_t val = (uint32_t)(((int32_t)v >> 1) + 1);
                        val <<= pLSB;
                        val |= 1u << (pLSB - 1);
                        val |= (v & 1u) << 31;
                        mu[4 * q + i] = val;
                    }
                }
            }
        }
    }

    if (ret > 0) {
        /* quad order -> raster, with the reference's masking of the odd row/column (:975-1007, quirk Q2) */
        for (int qy = 0; qy < QH; qy++)
            for (int qx = 0; qx < QW; qx++) {
                const int q = qy * QW + qx, y = 2 * qy, x = 2 * qx;
                const int keep_row = qy != QH - 1 || !(h & 1);
                const int keep_col = qx != QW - 1 || !(w & 1);
                const int keep[4] = { 1, keep_row, keep_col, keep_row | keep_col };
                for (int i = 0; i < 4; i++) {
                    int yy = y + (i & 1), xx = x + (i >> 1);
                    samples[yy * S + xx] = (int32_t)mu[4 * q + i] * keep[i];
                    states[(yy + 1) * S + xx + 1] |= sig[4 * q + i] * keep[i];
                }
            }
    }
    free(sig);
    free(E);
    free(mu);
    return ret;
}

/* H7: one group of up to 4 columns x up to 4 rows of a stripe (:1034-1077) */
static void sigprop_group(SpBits *sp, int y0, int x0, int gw, int gh, int S, int p, int causal,
                          int32_t *samples, uint8_t *states)
{
    for (int x = x0; x < x0 + gw; x++)
        for (int y = y0; y < y0 + gh; y++) {
            uint8_t *c = &states[(y + 1) * S + x + 1];
            int member = 0;
            if (!(c[0] & ST_SIGMA)) {
                int below_ok = !causal || y != y0 + gh - 1;
                unsigned up = c[-S - 1] | c[-S] | c[-S + 1];
                unsigned mid = c[-1] | c[1];
                unsigned dn = c[S - 1] | c[S] | c[S + 1];
                unsigned m = up | mid | (below_ok ? dn : 0);
                m |= ((up >> 3) & (up >> 4)) | ((mid >> 3) & (mid >> 4));
                if (below_ok)
                    m |= (dn >> 3) & (dn >> 4);
                member = m & 1;
            }
            unsigned st = c[0] | ST_SCANNED;
            if (member) {
                st |= ST_REFIND;
                int bit = sp_bit(sp);
                st |= bit ? ST_NEWSIG : 0;
                samples[y * S + x] |= (bit << p) | (bit << (p - 1));
            }
            c[0] |= (uint8_t)st;
        }
    for (int x = x0; x < x0 + gw; x++)
        for (int y = y0; y < y0 + gh; y++)
            if (states[(y + 1) * S + x + 1] & ST_NEWSIG)
                samples[y * S + x] |= (int32_t)((uint32_t)sp_bit(sp) << 31);
}

static void sigprop_pass(const uint8_t *Dref, int Lref, int w, int h, int S, int p, int causal,
                         int32_t *samples, uint8_t *states)
{
    SpBits sp = { Dref, Lref, 0, 0, 0, 0 };
    const int full = h / 4 * 4;
    for (int y0 = 0; y0 < full; y0 += 4)
        for (int x0 = 0; x0 < w; x0 += 4)
            sigprop_group(&sp, y0, x0, w - x0 < 4 ? w - x0 : 4, 4, S, p, causal, samples, states);
    /* the last, shorter stripe is walked even when it is empty (:1118-1130) */
    for (int x0 = 0; x0 < w; x0 += 4)
        sigprop_group(&sp, full, x0, w - x0 < 4 ? w - x0 : 4, h - full, S, p, causal, samples, states);
}

/* H8 (:1136-1185) */
static void magref_pass(const uint8_t *Dref, int Lref, int w, int h, int S, int p, int32_t *samples,
                        uint8_t *states)
{
    BwdBits mr = { Dref, Lref - 1, 0, 0 };
    for (int y0 = 0; y0 < h; y0 += 4) {
        int sh = h - y0 < 4 ? h - y0 : 4;
        for (int x = 0; x < w; x++)
            for (int y = y0; y < y0 + sh; y++) {
                uint8_t *c = &states[(y + 1) * S + x + 1];
                if (!(c[0] & ST_SIGMA))
                    continue;
                c[0] |= ST_REFIND;
                uint32_t bit = bwd_get(&mr, 1);
                uint32_t keep = (0xFFFFFFFEu | bit) << p;
                samples[y * S + x] = (int32_t)(((uint32_t)samples[y * S + x] & keep) | (1u << (p - 1)));
            }
    }
}

/* H1 */
int oracle_ht_block(const uint8_t *data, int length, int npasses, int Lcup, int Lref, int zbp,
                    int w, int h, int M_b, int roi_shift, int vsc, int32_t *out)
{
    memset(out, 0, sizeof(int32_t) * (size_t)w * h);
    if (npasses == 0)
        return 0;
    /* placeholder passes come in triples before the HT set (:1240-1247) */
    int rem = npasses % 3;
    int placeholders = rem ? npasses - rem : npasses - 3;
    int p0 = placeholders / 3;
    int z_blk = npasses - placeholders;
    if (z_blk <= 0)
        return 0;
    if (Lcup < 2)
        return ORACLE_ERR_INVALID;

    /* private copy: the decoder patches the segment in place (:1258, :1277-1278) */
    int alloc = (length > Lcup + Lref ? length : Lcup + Lref) + 16;
    uint8_t *buf = calloc((size_t)alloc, 1);
    memcpy(buf, data, (size_t)length);
    buf[length] = 0xFF;
    uint8_t *Dcup = buf, *Dref = buf + Lcup;

    int S_blk = (p0 + zbp) & 0xFF;
    int pLSB = (30 - S_blk) & 0xFF;
    int maxbp = S_blk + 1;                          /* cblk->zbp (an int) was overwritten with S_blk - 1, then + 2 (:605, :1263) */
    int ret;
    int Scup = (Dcup[Lcup - 1] << 4) + (Dcup[Lcup - 2] & 0x0F);
    if (Scup < 2 || Scup > Lcup || Scup > 4079) {
        free(buf);
        return ORACLE_ERR_INVALID;
    }
    int Pcup = Lcup - Scup;
    Dcup[Lcup - 1] = 0xFF;
    Dcup[Lcup - 2] |= 0x0F;

    const int S = w + 4;
    int32_t *samples = calloc((size_t)S * (h + 4), sizeof(int32_t));
    uint8_t *states = calloc((size_t)S * (h + 4), 1);
    ret = cleanup_pass(Dcup, Lcup, Pcup, pLSB, maxbp, w, h, S, samples, states);
    if (ret > 0) {
        if (z_blk > 1)
            sigprop_pass(Dref, Lref, w, h, S, pLSB - 1, vsc, samples, states);
        if (z_blk > 2)
            magref_pass(Dref, Lref, w, h, S, pLSB - 1, samples, states);
        const uint32_t mask = UINT32_MAX >> (M_b + 1);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                uint32_t v = (uint32_t)samples[y * S + x];
                uint32_t sign = v & 0x80000000u;
                v &= 0x7FFFFFFFu;
                if (roi_shift && (v & ~mask) == 0)
                    v <<= roi_shift;
                out[y * w + x] = (int32_t)(v | sign);
            }
    }
    free(samples);
    free(states);
    free(buf);
    return ret;
}
