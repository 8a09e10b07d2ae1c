/*
 * oracle/j2k_oracle_mq.c -- TEST INFRASTRUCTURE ONLY (see oracle/j2k_oracle.h).
 *
 * CPU restatement of the reference's Part-1 (EBCOT / MQ) code-block decoder, the path its jpeg2000 decoder
 * takes for code-blocks that are not HT:
 *   libavcodec/mqc.c:28-90          state table (ITU-T T.800 Table C.2; normative data, generated below from
 *                                   the 47-row form of the table)
 *   libavcodec/mqcdec.c:30-140      MQ decoder in the software convention (inverted C register, the low byte
 *                                   of C counts the shifts left before the next byte)
 *   libavcodec/jpeg2000.c:90-170    context labels: significance (Table D.1), sign (D.2/D.3), refinement (D.4)
 *   libavcodec/jpeg2000dec.c:1872-1990  significance propagation, magnitude refinement and clean-up passes
 *   libavcodec/jpeg2000dec.c:1993-2089  decode_cblk: pass loop, bit-plane bookkeeping, ROI, sign-magnitude output
 * Restated: the styles with one codeword segment -- 0, context reset (0x02), vertically causal contexts (0x08),
 * predictable termination (0x10, nothing to do in a decoder) and segmentation symbols (0x20); the T2 side
 * refuses arithmetic-coding bypass (0x01) and termination on each pass (0x04).
 *
 * Output as the reference leaves it in t1->data: bit 31 sign, magnitude left-aligned with the binary point at
 * bit 31 - M_b and the reconstruction half-bit set -- the same format the HT decoder produces, so the same
 * dequantisation follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "j2k_oracle.h"

/* ---- T.800 Table C.2 ------------------------------------------------------------------------------- */
static const struct { uint16_t qe; uint8_t nmps, nlps, sw; } mq_rows[47] = {
    { 0x5601, 1, 1, 1 },   { 0x3401, 2, 6, 0 },   { 0x1801, 3, 9, 0 },   { 0x0AC1, 4, 12, 0 },
    { 0x0521, 5, 29, 0 },  { 0x0221, 38, 33, 0 }, { 0x5601, 7, 6, 1 },   { 0x5401, 8, 14, 0 },
    { 0x4801, 9, 14, 0 },  { 0x3801, 10, 14, 0 }, { 0x3001, 11, 17, 0 }, { 0x2401, 12, 18, 0 },
    { 0x1C01, 13, 20, 0 }, { 0x1601, 29, 21, 0 }, { 0x5601, 15, 14, 1 }, { 0x5401, 16, 14, 0 },
    { 0x5101, 17, 15, 0 }, { 0x4801, 18, 16, 0 }, { 0x3801, 19, 17, 0 }, { 0x3401, 20, 18, 0 },
    { 0x3001, 21, 19, 0 }, { 0x2801, 22, 19, 0 }, { 0x2401, 23, 20, 0 }, { 0x2201, 24, 21, 0 },
    { 0x1C01, 25, 22, 0 }, { 0x1801, 26, 23, 0 }, { 0x1601, 27, 24, 0 }, { 0x1401, 28, 25, 0 },
    { 0x1201, 29, 26, 0 }, { 0x1101, 30, 27, 0 }, { 0x0AC1, 31, 28, 0 }, { 0x09C1, 32, 29, 0 },
    { 0x08A1, 33, 30, 0 }, { 0x0521, 34, 31, 0 }, { 0x0441, 35, 32, 0 }, { 0x02A1, 36, 33, 0 },
    { 0x0221, 37, 34, 0 }, { 0x0141, 38, 35, 0 }, { 0x0111, 39, 36, 0 }, { 0x0085, 40, 37, 0 },
    { 0x0049, 41, 38, 0 }, { 0x0025, 42, 39, 0 }, { 0x0015, 43, 40, 0 }, { 0x0009, 44, 41, 0 },
    { 0x0005, 45, 42, 0 }, { 0x0001, 45, 43, 0 }, { 0x5601, 46, 46, 0 },
};

enum { CX_SIG = 0, CX_SGN = 9, CX_REF = 14, CX_UNI = 17, CX_RL = 18, CX_COUNT = 19 };

typedef struct Mq {
    const uint8_t *data;
    int len, pos;              /* pos = index of the byte last folded into C */
    uint32_t a, c;
    uint8_t cx[CX_COUNT];      /* 2 * state + MPS */
} Mq;

/* the decoder plants 0xFF 0xFF behind the segment (jpeg2000dec.c:2014-2015): a marker it never steps over */
static inline unsigned mq_byte(const Mq *m, int i) { return i < m->len ? m->data[i] : 0xFF; }

static void mq_bytein(Mq *m)
{
    if (mq_byte(m, m->pos) == 0xFF) {
        if (mq_byte(m, m->pos + 1) > 0x8F)
            m->c++;                                        /* marker: feed ones, stay */
        else {
            m->pos++;
            m->c += 2 + 0xFE00 - (mq_byte(m, m->pos) << 9);   /* stuffed byte: 7 bits */
        }
    } else {
        m->pos++;
        m->c += 1 + 0xFF00 - (mq_byte(m, m->pos) << 8);
    }
}

static void mq_reset_contexts(Mq *m)                       /* ff_mqc_init_contexts (mqcdec.c) */
{
    memset(m->cx, 0, sizeof(m->cx));
    m->cx[CX_UNI] = 2 * 46;
    m->cx[CX_RL] = 2 * 3;
    m->cx[0] = 2 * 4;
}

static void mq_init(Mq *m, const uint8_t *data, int len)
{
    mq_reset_contexts(m);
    m->data = data;
    m->len = len;
    m->pos = 0;
    m->c = (mq_byte(m, 0) ^ 0xFFu) << 16;
    mq_bytein(m);
    m->c <<= 7;
    m->a = 0x8000;
}

static int mq_decode(Mq *m, int ctx)
{
    uint8_t *s = &m->cx[ctx];
    const unsigned idx = *s >> 1, mps = *s & 1, qe = mq_rows[idx].qe;
    int d;
    m->a -= qe;
    if ((m->c >> 16) < m->a) {
        if (m->a & 0x8000)
            return (int)mps;
        /* MPS path with conditional exchange */
        if (m->a < qe) {
            d = 1 - (int)mps;
            *s = (uint8_t)(2 * mq_rows[idx].nlps + (mps ^ mq_rows[idx].sw));
        } else {
            d = (int)mps;
            *s = (uint8_t)(2 * mq_rows[idx].nmps + mps);
        }
    } else {
        /* LPS path with conditional exchange */
        m->c -= m->a << 16;
        if (m->a < qe) {
            d = (int)mps;
            *s = (uint8_t)(2 * mq_rows[idx].nmps + mps);
        } else {
            d = 1 - (int)mps;
            *s = (uint8_t)(2 * mq_rows[idx].nlps + (mps ^ mq_rows[idx].sw));
        }
        m->a = qe;
    }
    do {                                                   /* RENORMD */
        if (!(m->c & 0xFF)) {
            m->c -= 0x100;
            mq_bytein(m);
        }
        m->a += m->a;
        m->c += m->c;
    } while (!(m->a & 0x8000));
    return d;
}

/* ---- sample state: one byte per sample inside a frame of width w + 2 ---------------------------------- */
enum { ST_SIG = 1, ST_NEG = 2, ST_VIS = 4, ST_REF = 8 };

typedef struct Blk {
    int w, h, stride;          /* stride = w + 2 */
    uint8_t *st;               /* (h + 2) rows */
    int32_t *val;              /* w * h, row stride w */
    int bandpos;
    int vsc;                   /* JPEG2000_CBLK_VSC: the last row of a stripe does not see the stripe below */
} Blk;

#define ST(b, x, y) ((b)->st[((y) + 1) * (b)->stride + (x) + 1])

static inline int sig(const Blk *b, int x, int y) { return ST(b, x, y) & ST_SIG; }

/* flags_mask of the coding passes (jpeg2000dec.c:1882-1883, :1916-1917, :1966-1968): with vertically causal
 * contexts the fourth row of a stripe reads its S, SW, SE neighbours as insignificant (and their sign as absent) */
static inline int blind_south(const Blk *b, int y) { return b->vsc && (y & 3) == 3; }

/* neighbourhood sums: horizontal, vertical, diagonal significant neighbours */
static void counts(const Blk *b, int x, int y, int *h, int *v, int *d)
{
    const int south = !blind_south(b, y);
    *h = !!sig(b, x - 1, y) + !!sig(b, x + 1, y);
    *v = !!sig(b, x, y - 1) + (south && sig(b, x, y + 1));
    *d = !!sig(b, x - 1, y - 1) + !!sig(b, x + 1, y - 1) + (south && sig(b, x - 1, y + 1)) + (south && sig(b, x + 1, y + 1));
}

/* Table D.1; bandpos 0 LL, 1 HL, 2 LH, 3 HH */
static int sig_context(const Blk *b, int x, int y)
{
    int h, v, d;
    counts(b, x, y, &h, &v, &d);
    if (b->bandpos == 3) {
        const int hv = h + v;
        if (d >= 3) return 8;
        if (d == 2) return hv >= 1 ? 7 : 6;
        if (d == 1) return hv >= 2 ? 5 : hv == 1 ? 4 : 3;
        return hv >= 2 ? 2 : hv;
    }
    if (b->bandpos == 1) { const int t = h; h = v; v = t; }
    if (h == 2) return 8;
    if (h == 1) return v >= 1 ? 7 : d >= 1 ? 6 : 5;
    if (v == 2) return 4;
    if (v == 1) return 3;
    return d >= 2 ? 2 : d;
}

/* Tables D.2 / D.3: context 9..13 and the bit the decoded decision is XORed with */
static int sign_context(const Blk *b, int x, int y, int *xorbit)
{
    static const int8_t contrib[4] = { 0, 1, 0, -1 };     /* index = SIG | NEG: insignificant 0, positive +1, negative -1 */
    int hc = contrib[ST(b, x - 1, y) & 3] + contrib[ST(b, x + 1, y) & 3];
    int vc = contrib[ST(b, x, y - 1) & 3] + (blind_south(b, y) ? 0 : contrib[ST(b, x, y + 1) & 3]);
    hc = hc > 1 ? 1 : hc < -1 ? -1 : hc;
    vc = vc > 1 ? 1 : vc < -1 ? -1 : vc;
    if (hc < 0 || (hc == 0 && vc < 0)) {
        *xorbit = 1;
        hc = -hc;
        vc = -vc;
    } else {
        *xorbit = 0;
    }
    if (hc == 1)
        return CX_SGN + 3 + vc;                            /* (1,-1) 11, (1,0) 12, (1,1) 13 */
    return CX_SGN + vc;                                    /* (0,0) 9, (0,1) 10 */
}

static int ref_context(const Blk *b, int x, int y)
{
    int h, v, d;
    if (ST(b, x, y) & ST_REF)
        return CX_REF + 2;
    counts(b, x, y, &h, &v, &d);
    return CX_REF + (h + v + d ? 1 : 0);
}

static inline int has_sig_neighbour(const Blk *b, int x, int y)
{
    int h, v, d;
    counts(b, x, y, &h, &v, &d);
    return h + v + d;
}

static void become_significant(Blk *b, Mq *m, int x, int y, int32_t mask)
{
    int xorbit;
    const int ctx = sign_context(b, x, y, &xorbit);
    const int neg = mq_decode(m, ctx) ^ xorbit;
    b->val[y * b->w + x] |= (int32_t)((uint32_t)neg << 31) | mask;
    ST(b, x, y) |= ST_SIG | (neg ? ST_NEG : 0);
}

/* decode_sigpass (jpeg2000dec.c:1872-1904) */
static void pass_sigprop(Blk *b, Mq *m, int bp)
{
    const int32_t mask = 3 << (bp - 1);
    for (int y0 = 0; y0 < b->h; y0 += 4)
        for (int x = 0; x < b->w; x++)
            for (int y = y0; y < b->h && y < y0 + 4; y++) {
                if ((ST(b, x, y) & (ST_SIG | ST_VIS)) || !has_sig_neighbour(b, x, y))
                    continue;
                if (mq_decode(m, CX_SIG + sig_context(b, x, y)))
                    become_significant(b, m, x, y, mask);
                ST(b, x, y) |= ST_VIS;
            }
}

/* decode_refpass (:1906-1931) */
static void pass_magref(Blk *b, Mq *m, int bp)
{
    const int32_t phalf = 1 << (bp - 1);
    for (int y0 = 0; y0 < b->h; y0 += 4)
        for (int x = 0; x < b->w; x++)
            for (int y = y0; y < b->h && y < y0 + 4; y++) {
                if ((ST(b, x, y) & (ST_SIG | ST_VIS)) != ST_SIG)
                    continue;
                int32_t *p = &b->val[y * b->w + x];
                const int ctx = ref_context(b, x, y);
                *p |= phalf;
                if (mq_decode(m, ctx))
                    *p |= phalf << 1;
                else
                    *p &= ~(phalf << 1);
                ST(b, x, y) |= ST_REF;
            }
}

/* decode_clnpass (:1933-1990); the segmentation symbol behind it is read and, like in the reference, only logged */
static void pass_cleanup(Blk *b, Mq *m, int bp, int segsym)
{
    const int32_t mask = 3 << (bp - 1);
    for (int y0 = 0; y0 < b->h; y0 += 4)
        for (int x = 0; x < b->w; x++) {
            int runlen = 0, dec = 0;
            if (y0 + 3 < b->h) {
                int quiet = 1;                              /* a full column, nothing significant, visited or near */
                for (int y = y0; y < y0 + 4 && quiet; y++)
                    quiet = !(ST(b, x, y) & (ST_SIG | ST_VIS)) && !has_sig_neighbour(b, x, y);
                if (quiet) {
                    if (!mq_decode(m, CX_RL))
                        continue;
                    runlen = mq_decode(m, CX_UNI);
                    runlen = (runlen << 1) | mq_decode(m, CX_UNI);
                    dec = 1;
                }
            }
            for (int y = y0 + runlen; y < y0 + 4 && y < b->h; y++) {
                if (!dec && !(ST(b, x, y) & (ST_SIG | ST_VIS)))
                    dec = mq_decode(m, CX_SIG + sig_context(b, x, y));
                if (dec)
                    become_significant(b, m, x, y, mask);
                dec = 0;
                ST(b, x, y) &= ~ST_VIS;
            }
        }
    if (segsym)
        for (int i = 0; i < 4; i++)
            mq_decode(m, CX_UNI);
}

/* decode_cblk (:1993-2089).  nonzerobits = expn + guard bits - 1 - (zero bit-planes - roi shift of component 0).
 * out: w * h words, row stride w.  Returns 0 = no data (block stays zero), 1 = decoded,
 * ORACLE_ERR_INVALID = "bpno became invalid": the reference stops there and still dequantises what it has. */
int oracle_mq_block(const uint8_t *data, int length, int npasses, int nonzerobits, int w, int h, int bandpos,
                    int M_b, int roi_shift, int32_t *out)
{
    return oracle_mq_block_style(data, length, npasses, nonzerobits, w, h, bandpos, M_b, roi_shift, 0, out);
}

/* style = SPcod code-block style: 0x02 reset, 0x08 causal, 0x20 segmentation symbols (0x10 changes nothing) */
int oracle_mq_block_style(const uint8_t *data, int length, int npasses, int nonzerobits, int w, int h, int bandpos,
                          int M_b, int roi_shift, int style, int32_t *out)
{
    memset(out, 0, sizeof(int32_t) * (size_t)w * h);
    if (!length)
        return 0;
    Blk b = { w, h, w + 2, calloc((size_t)(w + 2) * (h + 3), 1), out, bandpos, style & 0x08 };
    Mq m;
    mq_init(&m, data, length);
    int bpno = nonzerobits - 1 + 31 - M_b - 1 - roi_shift, pass_t = 2, ret = 1;
    while (npasses--) {
        if (bpno < 0 || bpno > 29) {
            ret = ORACLE_ERR_INVALID;
            break;
        }
        if (pass_t == 0) pass_sigprop(&b, &m, bpno + 1);
        else if (pass_t == 1) pass_magref(&b, &m, bpno + 1);
        else pass_cleanup(&b, &m, bpno + 1, style & 0x20);
        if (style & 0x02)                                   /* JPEG2000_CBLK_RESET (:2038-2039) */
            mq_reset_contexts(&m);
        if (++pass_t == 3) {
            bpno--;
            pass_t = 0;
        }
    }
    free(b.st);
    if (ret == 1 && roi_shift) {
        const uint32_t mask = UINT32_MAX >> (M_b + 1);    /* ROI detection (:2075-2087) */
        for (int i = 0; i < w * h; i++) {
            uint32_t v = (uint32_t)out[i] & 0x7FFFFFFFu, s = (uint32_t)out[i] & 0x80000000u;
            if ((v & ~mask) == 0)
                v <<= roi_shift;
            out[i] = (int32_t)(v | s);
        }
    }
    return ret;
}
