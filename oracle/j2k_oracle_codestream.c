/*
 * oracle/j2k_oracle_codestream.c -- TEST INFRASTRUCTURE ONLY (see j2k_oracle.h).
 *
 * A deliberately plain, pointer-tree decoder for the HT subset of the reference's native jpeg2000
 * decoder: marker segments, geometry, tier-2 packet headers, then per tile the hot path
 * (H0 tile_codeblocks, D1-D3 dequantisation, W ff_dwt_decode, M mct_decode, P1 write_frame).
 *
 *   marker segments      libavcodec/jpeg2000dec.c:197-1014, main loop :2425-2637
 *   geometry             libavcodec/jpeg2000.c:214-577, init_tile jpeg2000dec.c:1016-1070
 *   bit reader / tag tree jpeg2000dec.c:70-131
 *   packet headers (HT)  jpeg2000dec.c:1136-1542
 *   progression orders   jpeg2000dec.c:1544-1833
 *   dequantisation       jpeg2000dec.c:2098-2181
 *   tile_codeblocks      jpeg2000dec.c:2212-2299
 *   write_frame          jpeg2000dec.c:2301-2366
 *
 * Not handled (returns ORACLE_ERR_PATCH): Part-1 MQ code-blocks, MIXED/HTDECLARED sets, tile-part POC, PPM/PPT,
 * tile-part COD/QCD.  Those are outside SURVEY.md section 8's first bar.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "j2k_oracle.h"

#define MAXC 4
#define MAXRES 33

typedef struct Rect { int x0, x1, y0, y1; } Rect;

typedef struct TagNode { int val, vis; struct TagNode *parent; } TagNode;

typedef struct Cblk {
    int x0, x1, y0, y1;          /* plane coordinates relative to band origin, Mallat offset included */
    int npasses, zbp, incl, lblock, plhd;
    int modes;                   /* cblk->modes: SPcod copy; bit 6 cleared once a MIXED block turns out to be a Part-1 block */
    int Lcup, Lref;
    uint8_t *data;
    int length, cap;
} Cblk;

typedef struct Prec {
    int ncw, nch, decoded_layers;
    Cblk *cb;
    TagNode *incl, *zb;
} Prec;

typedef struct Band {
    Rect c;
    int l2cw, l2ch;
    float fstep;
    int istep;
    Prec *prec;
} Band;

typedef struct Res {
    Rect c;
    int ppx, ppy, npx, npy, nbands;
    Band band[3];
} Res;

typedef struct CodSty {
    int nres, l2cw, l2ch, cblk_style, transform, csty, nlayers, prog, mct, init;
    int ppx[MAXRES], ppy[MAXRES];
} CodSty;

typedef struct QntSty {
    int expn[100], mant[100], style, nguard;
} QntSty;

typedef struct Comp {
    Rect c;                                         /* comp->coord_o: full-resolution bounds */
    Rect cr;                                        /* comp->coord: reduced by lowres (jpeg2000dec.c:1047-1050) */
    Res *res;
    void *data;
    int roi_shift;
} Comp;

typedef struct TilePart {
    const uint8_t *p, *end;                         /* tpg: packets of this tile-part */
    const uint8_t *hp, *hend;                       /* header_tpg: its run of PPM packet headers */
} TilePart;

typedef struct Tile {
    Rect c;
    Comp comp[MAXC];
    TilePart part[32];
    int nparts;
    /* PPT: packed packet headers of this tile (get_ppt, jpeg2000dec.c:982-1014) */
    int first_seen;                                 /* a tile-part with TPsot == 0 set this tile's coding styles (get_sot, :862-869) */
    /* the tile's own styles: the main header's at that point, then what COD / COC / QCD / QCC / POC segments of its
     * tile-part headers change (tile->codsty / qntsty / poc / properties) */
    CodSty cod[MAXC];
    QntSty qnt[MAXC];
    int had_coc[MAXC], had_qcc[MAXC];
    int npoc, poc_default;
    struct { int RS, CS, LYE, RE, CE, P; } poc[32];
    int has_ppt;
    uint8_t *ppt;
    size_t ppt_live;                                /* size of it when that stream was set up (SOD of tile-part 0) */
    size_t ppt_size, ppt_pos;                       /* ppt_pos = tile->packed_headers_stream */
} Tile;

typedef struct Dec {
    const uint8_t *cs, *end;
    int bitexact;
    int red;                                        /* reduction_factor = AVCodecContext.lowres (jpeg2000dec.c:2807-2823) */
    int W, H, X0, Y0, TW, TH, TX0, TY0, ncomp, precision;
    int cbps[MAXC], cdx[MAXC], cdy[MAXC], roi[MAXC];
    int isHT, cap_b05, cap_b11, cap_b12, cap_b14;
    CodSty cod[MAXC];
    QntSty qnt[MAXC];
    int had_coc[MAXC], had_qcc[MAXC];
    int ntx, nty;
    Tile *tile;
    int pix, out_w, out_h;
    int colour_space, pal8, cdef[4];                /* JP2 wrapper state (jp2_find_codestream) */
    int rsiz;                                       /* avctx->profile: 3 / 4 = D-Cinema 2K / 4K */
    uint32_t palette[256];                          /* s->palette (jpeg2000dec.h:94): 0xAARRGGBB per entry */
    int sar_num, sar_den;                           /* s->sar from the 'res ' box (:2762-2796), 0/0 = none */
    int npoc;                                       /* main-header POC entries (get_poc) */
    struct { int RS, CS, LYE, RE, CE, P; } poc[32];
    int chroma_lw, chroma_lh;                       /* log2_chroma_w/h of the chosen AVPixelFormat */
    /* PPM: packed packet headers of the whole codestream (get_ppm, jpeg2000dec.c:958-980) */
    int has_ppm;
    uint8_t *ppm;
    size_t ppm_size, ppm_pos;                       /* ppm_pos = s->packed_headers_stream */
    /* packet header bit reader */
    const uint8_t *bp, *bend;
    int bit_index;
    int64_t nblocks;
} Dec;

static inline int ceildiv(int a, int b) { return (int)(((int64_t)a + b - 1) / b); }
static inline int ceildivpow2(int a, int b) { return (int)(-((-(int64_t)a) >> b)); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clip(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int clip64(int64_t v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : (int)v; }
static inline int ilog2(unsigned v) { int n = 0; while (v >>= 1) n++; return n; }

static inline unsigned rd16(const uint8_t *p) { return (p[0] << 8) | p[1]; }
static inline uint32_t rd32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

/* ---- packet-header bits (jpeg2000dec.c:70-90) ------------------------------------------------------ */
static inline unsigned hb_getbyte(Dec *d) { return d->bp < d->bend ? *d->bp++ : 0; }
static inline unsigned hb_peekbyte(Dec *d) { return d->bp < d->bend ? *d->bp : 0; }

static int hb_bits(Dec *d, int n)
{
    int res = 0;
    while (--n >= 0) {
        res <<= 1;
        if (d->bit_index == 0)
            d->bit_index = 7 + (hb_getbyte(d) != 0xFF);
        d->bit_index--;
        res |= (hb_peekbyte(d) >> d->bit_index) & 1;
    }
    return res;
}

static void hb_flush(Dec *d)
{
    if (hb_getbyte(d) == 0xFF && d->bp < d->bend)
        d->bp++;
    d->bit_index = 8;
}

/* ---- tag trees (jpeg2000.c:41-94, jpeg2000dec.c:93-131) ------------------------------------------- */
static TagNode *tagtree_new(int w, int h)
{
    int total = 1, ww = w, hh = h;
    while (ww > 1 || hh > 1) {
        total += ww * hh;
        ww = (ww + 1) >> 1;
        hh = (hh + 1) >> 1;
    }
    TagNode *t = calloc((size_t)total, sizeof(*t)), *lvl = t;
    while (w > 1 || h > 1) {
        int pw = (w + 1) >> 1;
        TagNode *up = lvl + w * h;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                lvl[y * w + x].parent = &up[(y >> 1) * pw + (x >> 1)];
        lvl = up;
        w = pw;
        h = (h + 1) >> 1;
    }
    return t;
}

static int tagtree_decode(Dec *d, TagNode *node, int threshold)
{
    TagNode *stack[32];
    int sp = -1;
    while (node && !node->vis) {
        stack[++sp] = node;
        node = node->parent;
    }
    int cur = node ? node->val : stack[sp]->val;
    while (cur < threshold && sp >= 0) {
        if (cur < stack[sp]->val)
            cur = stack[sp]->val;
        while (cur < threshold) {
            if (hb_bits(d, 1)) {
                stack[sp]->vis++;
                break;
            }
            cur++;
        }
        stack[sp]->val = cur;
        sp--;
    }
    return cur;
}

/* ---- marker segments -------------------------------------------------------------------------------- */
static int parse_cox(Dec *d, const uint8_t *p, const uint8_t *end, CodSty *c)
{
    if (end - p < 5)
        return ORACLE_ERR_INVALID;
    c->nres = p[0] + 1;
    if (c->nres > MAXRES)                           /* JPEG2000_MAX_RESLEVELS = 34: up to 32 decomposition levels */
        return ORACLE_ERR_INVALID;
    if (c->nres <= d->red)                          /* "reduction_factor too large for this bitstream" (:509-517) */
        return ORACLE_ERR_EINVAL;
    c->l2cw = (p[1] & 15) + 2;
    c->l2ch = (p[2] & 15) + 2;
    if (c->l2cw > 10 || c->l2ch > 10 || c->l2cw + c->l2ch > 12)
        return ORACLE_ERR_INVALID;
    c->cblk_style = p[3];
    c->transform = p[4];                                  /* 0 = 9/7, 1 = 5/3 */
    if (d->bitexact && c->transform == ORACLE_DWT97)
        c->transform = ORACLE_DWT97_INT;                  /* jpeg2000dec.c:543-544 */
    p += 5;
    if (c->csty & 1) {
        for (int i = 0; i < c->nres; i++) {
            if (p >= end)
                return ORACLE_ERR_INVALID;
            c->ppx[i] = *p & 15;
            c->ppy[i] = *p >> 4;
            p++;
            if (i && (!c->ppx[i] || !c->ppy[i]))
                return ORACLE_ERR_INVALID;
        }
    } else {
        for (int i = 0; i < MAXRES; i++)
            c->ppx[i] = c->ppy[i] = 15;
    }
    return 0;
}

/* get_qcx (jpeg2000dec.c:675-717).  n = bytes the segment declares for Sqcx + step sizes, avail = bytes to the end
 * of the packet: what is read depends on the style alone and is checked against the packet; the marker loop then
 * compares the consumption with the declared length (:2629).  -> bytes consumed or an error */
static int parse_qcx(const uint8_t *p, int n, int64_t avail, QntSty *q)
{
    if (avail < 1)
        return ORACLE_ERR_INVALID;
    q->nguard = p[0] >> 5;
    q->style = p[0] & 31;
    p++;
    avail--;
    int used = 1;
    if (q->style == 0) {
        n -= 1;
        if (avail < n || n > 99)
            return ORACLE_ERR_INVALID;
        for (int i = 0; i < n; i++, used++)
            q->expn[i] = p[i] >> 3;
    } else if (q->style == 1) {
        if (avail < 2)
            return ORACLE_ERR_INVALID;
        unsigned x = rd16(p);
        used += 2;
        q->expn[0] = x >> 11;
        q->mant[0] = x & 0x7FF;
        for (int i = 1; i < 99; i++) {
            q->expn[i] = imax(0, q->expn[0] - (i - 1) / 3);
            q->mant[i] = q->mant[0];
        }
    } else {                                              /* 2 and the undefined styles alike */
        n = (n - 1) >> 1;
        if (avail < 2 * n || n > 99)
            return ORACLE_ERR_INVALID;
        for (int i = 0; i < n; i++, used += 2) {
            unsigned x = rd16(p + 2 * i);
            q->expn[i] = x >> 11;
            q->mant[i] = x & 0x7FF;
        }
    }
    return used;
}

/* pixel format choice for a raw codestream (jpeg2000dec.c:134-193, :335-363) */
/* ---- pixel format selection (get_siz, jpeg2000dec.c:134-193, :328-412) -----------------------------------
 * The fields of libavutil's AVPixFmtDescriptor that pix_fmt_match() reads, for the formats of the reference's
 * rgb / gray / yuv / xyz lists, in list order. */
typedef struct PixDesc {
    const char *name;
    int klass, ncomp, depth, lw, lh, pal;
} PixDesc;
#define P8 ORACLE_PIX_PLANAR8
#define P16 ORACLE_PIX_PLANAR16
static const PixDesc RGB_FMTS[] = {
    { "pal8", ORACLE_PIX_GRAY8, 1, 8, 0, 0, 1 },    /* index plane written like gray8 (write_frame_8), then the palette */
    { "rgb24", ORACLE_PIX_RGB24, 3, 8, 0, 0, 0 }, { "rgba", ORACLE_PIX_RGBA, 4, 8, 0, 0, 0 },
    { "rgb48le", ORACLE_PIX_RGB48, 3, 16, 0, 0, 0 }, { "rgba64le", ORACLE_PIX_RGBA64, 4, 16, 0, 0, 0 },
};
static const PixDesc GRAY_FMTS[] = {
    { "gray", ORACLE_PIX_GRAY8, 1, 8, 0, 0, 0 }, { "ya8", ORACLE_PIX_YA8, 2, 8, 0, 0, 0 },
    { "gray16le", ORACLE_PIX_GRAY16, 1, 16, 0, 0, 0 }, { "ya16le", ORACLE_PIX_YA16, 2, 16, 0, 0, 0 },
};
static const PixDesc YUV_FMTS[] = {
    { "yuv410p", P8, 3, 8, 2, 2, 0 }, { "yuv411p", P8, 3, 8, 2, 0, 0 }, { "yuva420p", P8, 4, 8, 1, 1, 0 },
    { "yuv420p", P8, 3, 8, 1, 1, 0 }, { "yuv422p", P8, 3, 8, 1, 0, 0 }, { "yuva422p", P8, 4, 8, 1, 0, 0 },
    { "yuv440p", P8, 3, 8, 0, 1, 0 }, { "yuv444p", P8, 3, 8, 0, 0, 0 }, { "yuva444p", P8, 4, 8, 0, 0, 0 },
    { "yuv420p9le", P16, 3, 9, 1, 1, 0 }, { "yuv422p9le", P16, 3, 9, 1, 0, 0 }, { "yuv444p9le", P16, 3, 9, 0, 0, 0 },
    { "yuva420p9le", P16, 4, 9, 1, 1, 0 }, { "yuva422p9le", P16, 4, 9, 1, 0, 0 }, { "yuva444p9le", P16, 4, 9, 0, 0, 0 },
    { "yuv420p10le", P16, 3, 10, 1, 1, 0 }, { "yuv422p10le", P16, 3, 10, 1, 0, 0 }, { "yuv444p10le", P16, 3, 10, 0, 0, 0 },
    { "yuva420p10le", P16, 4, 10, 1, 1, 0 }, { "yuva422p10le", P16, 4, 10, 1, 0, 0 }, { "yuva444p10le", P16, 4, 10, 0, 0, 0 },
    { "yuv420p12le", P16, 3, 12, 1, 1, 0 }, { "yuv422p12le", P16, 3, 12, 1, 0, 0 }, { "yuv444p12le", P16, 3, 12, 0, 0, 0 },
    { "yuv420p14le", P16, 3, 14, 1, 1, 0 }, { "yuv422p14le", P16, 3, 14, 1, 0, 0 }, { "yuv444p14le", P16, 3, 14, 0, 0, 0 },
    { "yuv420p16le", P16, 3, 16, 1, 1, 0 }, { "yuv422p16le", P16, 3, 16, 1, 0, 0 }, { "yuv444p16le", P16, 3, 16, 0, 0, 0 },
    { "yuva420p16le", P16, 4, 16, 1, 1, 0 }, { "yuva422p16le", P16, 4, 16, 1, 0, 0 }, { "yuva444p16le", P16, 4, 16, 0, 0, 0 },
};
static const PixDesc XYZ_FMTS[] = { { "xyz12le", ORACLE_PIX_RGB48, 3, 12, 0, 0, 0 } };
#undef P8
#undef P16
#define NELEM(a) ((int)(sizeof(a) / sizeof((a)[0])))

static int log2c(int v) { return v >= 4 ? 2 : v >= 2 ? 1 : 0; }

/* pix_fmt_match (:134-166) */
static int pix_match(const PixDesc *f, const Dec *d)
{
    if (f->ncomp != d->ncomp || f->depth < d->precision)
        return 0;
    int ok = 1;
    switch (d->ncomp) {
    case 4: ok = ok && d->cdx[3] == 1 && d->cdy[3] == 1;                                /* fall through */
    case 3: ok = ok && log2c(d->cdx[2]) == f->lw && log2c(d->cdy[2]) == f->lh;         /* fall through */
    case 2: ok = ok && log2c(d->cdx[1]) == f->lw && log2c(d->cdy[1]) == f->lh;         /* fall through */
    case 1: ok = ok && d->cdx[0] == 1 && d->cdy[0] == 1 && f->pal == (d->pal8 ? 1 : 0);
    }
    return ok;
}

static const PixDesc *first_match(const PixDesc *list, int n, const Dec *d)
{
    for (int i = 0; i < n; i++)
        if (pix_match(&list[i], d))
            return &list[i];
    return NULL;
}

static _Thread_local const char *g_pix_name = "none";
const char *oracle_last_pix_fmt_name(void) { return g_pix_name; }

static int choose_pix(Dec *d)
{
    static const PixDesc YUVA420P_FALLBACK = { "yuva420p", ORACLE_PIX_PLANAR8, 4, 8, 1, 1, 0 };
    const PixDesc *f = NULL;
    if (d->precision > 16)
        return 0;
    const int dcinema = d->rsiz == 3 || d->rsiz == 4;
    switch (dcinema ? -1 : d->colour_space) {       /* :330-352 */
    case -1:                                        /* D-Cinema profiles: XYZ12, then the YUV list */
        if (d->precision <= 8 && d->ncomp == 3)     /* would be written bytewise into half of a 16-bit frame: the other
                                                     * half is whatever the frame buffer held -- not restated */
            return 0;
        f = first_match(XYZ_FMTS, NELEM(XYZ_FMTS), d);
        if (!f) f = first_match(YUV_FMTS, NELEM(YUV_FMTS), d);
        break;
    case 16: f = first_match(RGB_FMTS, NELEM(RGB_FMTS), d); break;
    case 17: f = first_match(GRAY_FMTS, NELEM(GRAY_FMTS), d); break;
    case 18: f = first_match(YUV_FMTS, NELEM(YUV_FMTS), d); break;
    default:
        f = first_match(RGB_FMTS, NELEM(RGB_FMTS), d);
        if (!f) f = first_match(GRAY_FMTS, NELEM(GRAY_FMTS), d);
        if (!f) f = first_match(YUV_FMTS, NELEM(YUV_FMTS), d);
        if (!f) f = first_match(XYZ_FMTS, NELEM(XYZ_FMTS), d);
    }
    if (!f) {                                       /* the last resort, :364-398 */
        const int same01 = d->ncomp >= 2 && d->cdx[0] == d->cdx[1] && d->cdy[0] == d->cdy[1];
        if (d->ncomp == 4 && d->cdx[0] == 1 && d->cdy[0] == 1 && d->cdx[1] == 1 && d->cdy[1] == 1 &&
            d->cdx[2] == d->cdx[3] && d->cdy[2] == d->cdy[3]) {
            if (d->precision == 8 && d->cdx[2] == 2 && d->cdy[2] == 2 && !d->pal8) {
                f = &YUVA420P_FALLBACK;
                for (int i = 0; i < 4; i++)
                    d->cdef[i] = i;
            }
        } else if (d->ncomp == 3 && d->precision == 8 && same01 && d->cdx[0] == d->cdx[2] && d->cdy[0] == d->cdy[2]) {
            f = &RGB_FMTS[1];
        } else if (d->ncomp == 2 && d->precision == 8 && same01) {
            f = &GRAY_FMTS[1];
        } else if (d->ncomp == 2 && d->precision == 16 && same01) {
            f = &GRAY_FMTS[3];
        } else if (d->ncomp == 1 && d->precision == 8) {
            f = &GRAY_FMTS[0];
        } else if (d->ncomp == 1 && d->precision == 12) {
            f = &GRAY_FMTS[2];
        }
    }
    if (!f || !f->klass)
        return 0;
    g_pix_name = f->name;
    d->chroma_lw = f->lw;
    d->chroma_lh = f->lh;
    return f->klass;
}

/* av_reduce(dst_num, dst_den, num, den, max) of libavutil (rational.c:35-78) for non-negative arguments: the
 * fraction in lowest terms when it fits, else the best approximation with both terms <= max -- the last convergent
 * of the continued fraction that fits, or the last semi-convergent when that one is closer. */
static void reduce_fraction(int *out_num, int *out_den, int64_t num, int64_t den, int64_t max)
{
    int64_t a = num, b = den;
    while (b) { int64_t t = a % b; a = b; b = t; }           /* gcd */
    if (a) { num /= a; den /= a; }
    int64_t p0 = 0, q0 = 1, p1 = 1, q1 = 0;                   /* convergents p/q: previous and current */
    if (num <= max && den <= max) {
        p1 = num; q1 = den; den = 0;
    }
    while (den) {
        const uint64_t x = (uint64_t)(num / den);
        const int64_t rest = num - den * (int64_t)x;
        const int64_t p2 = (int64_t)x * p1 + p0, q2 = (int64_t)x * q1 + q0;
        if (p2 > max || q2 > max) {
            uint64_t k = x;
            if (p1) k = (uint64_t)((max - p0) / p1);
            if (q1) { const uint64_t kq = (uint64_t)((max - q0) / q1); if (kq < k) k = kq; }
            if (den * (2 * (int64_t)k * q1 + q0) > num * q1) {
                const int64_t np = (int64_t)k * p1 + p0, nq = (int64_t)k * q1 + q0;
                p1 = np; q1 = nq;
            }
            break;
        }
        p0 = p1; q0 = q1; p1 = p2; q1 = q2;
        num = den; den = rest;
    }
    *out_num = (int)p1;
    *out_den = (int)q1;
}

/* ---- JP2 wrapper: jp2_find_codestream (:2658-2804).  Returns 1 with *ppos at the codestream box payload,
 * 0 when no codestream box was found, < 0 for a malformed box. */
static uint32_t jp2_be(const Dec *d, int64_t *pos, int n)
{
    uint32_t v = 0;
    const int64_t size = d->end - d->cs;
    for (int i = 0; i < n; i++, (*pos)++)
        v = v << 8 | (*pos < size ? d->cs[*pos] : 0);   /* unchecked reads land in the zero padding */
    return v;
}

static int jp2_find_codestream(Dec *d, int64_t *ppos)
{
    const int64_t size = d->end - d->cs;
    int64_t pos = *ppos;
    int search_range = 10;
#define SEEK(p) (pos = (p) < 0 ? 0 : (p) > size ? size : (p))
    while (search_range && size - pos >= 8) {
        uint32_t atom_size = jp2_be(d, &pos, 4);
        const uint32_t atom = jp2_be(d, &pos, 4);
        uint32_t atom_end;
        if (atom_size == 1) {
            if (jp2_be(d, &pos, 4))
                return 0;
            atom_size = jp2_be(d, &pos, 4);
            if (atom_size < 16 || pos + atom_size - 16 > INT32_MAX) {
                *ppos = pos;
                return ORACLE_ERR_INVALID;
            }
            atom_end = (uint32_t)(pos + atom_size - 16);
        } else {
            if (atom_size < 8 || pos + atom_size - 8 > INT32_MAX) {
                *ppos = pos;
                return ORACLE_ERR_INVALID;
            }
            atom_end = (uint32_t)(pos + atom_size - 8);
        }
        if (atom == 0x6A703263u) {                  /* jp2c */
            *ppos = pos;
            return 1;
        }
        if (size - pos < (int64_t)atom_size || atom_end < atom_size)
            return 0;
        if (atom == 0x6A703268u && atom_size >= 16) {   /* jp2h */
            uint32_t atom2_end;
            do {
                if (size - pos < 8)
                    break;
                uint32_t atom2_size = jp2_be(d, &pos, 4);
                const uint32_t atom2 = jp2_be(d, &pos, 4);
                atom2_end = (uint32_t)(pos + atom2_size - 8);
                if (atom2_size < 8 || atom2_end > atom_end || atom2_end < atom2_size)
                    break;
                atom2_size -= 8;
                if (atom2 == 0x6A703263u) {
                    *ppos = pos;
                    return 1;
                } else if (atom2 == 0x636F6C72u && atom2_size >= 7) {          /* colr */
                    const int method = (int)jp2_be(d, &pos, 1);
                    pos += 2;
                    if (method == 1)
                        d->colour_space = (int)jp2_be(d, &pos, 4);
                } else if (atom2 == 0x70636C72u && atom2_size >= 6) {          /* pclr */
                    const int count = (int)jp2_be(d, &pos, 2), channels = (int)jp2_be(d, &pos, 1);
                    int depth[3], bytes = 0;
                    for (int i = 0; i < 3; i++) {
                        depth[i] = (int)(jp2_be(d, &pos, 1) & 0x7F) + 1;
                        bytes += ((depth[i] + 7) >> 3) * count;
                    }
                    if (!(count > 256 || channels != 3 || depth[0] > 16 || depth[1] > 16 || depth[2] > 16 ||
                          (int64_t)atom2_size < bytes)) {
                        d->pal8 = 1;
                        for (int i = 0; i < count; i++) {                      /* :2731-2752, uint32 arithmetic */
                            uint32_t v[3];
                            for (int k = 0; k < 3; k++) {
                                if (depth[k] <= 8) {
                                    v[k] = jp2_be(d, &pos, 1) << (8 - depth[k]);
                                    v[k] |= v[k] >> depth[k];
                                } else {
                                    v[k] = jp2_be(d, &pos, 2) >> (depth[k] - 8);
                                }
                            }
                            d->palette[i] = 0xffu << 24 | v[0] << 16 | v[1] << 8 | v[2];
                        }
                    }
                } else if (atom2 == 0x63646566u && atom2_size >= 2) {          /* cdef */
                    for (int n = (int)jp2_be(d, &pos, 2); n > 0; n--) {
                        const int cn = (int)jp2_be(d, &pos, 2);
                        jp2_be(d, &pos, 2);
                        const int asoc = (int)jp2_be(d, &pos, 2);
                        if (cn < 4 && asoc < 4)
                            d->cdef[cn] = asoc;
                        if (pos >= size + 6)
                            break;
                    }
                } else if (atom2 == 0x72657320u && atom2_size >= 18) {         /* 'res ' -> sample aspect ratio */
                    pos += 4;
                    const uint32_t resx = jp2_be(d, &pos, 4);
                    if (resx == 0x72657363u || resx == 0x72657364u) {          /* resc / resd */
                        int64_t vnum = jp2_be(d, &pos, 2), vden = jp2_be(d, &pos, 2);
                        int64_t hnum = jp2_be(d, &pos, 2), hden = jp2_be(d, &pos, 2);
                        int64_t vexp = jp2_be(d, &pos, 1), hexp = jp2_be(d, &pos, 1);
                        if (vnum && vden && hnum && hden) {
                            if (vexp > hexp) { vexp -= hexp; hexp = 0; } else { hexp -= vexp; vexp = 0; }
                            if (INT64_MAX / (hnum * vden) > pow(10, hexp) && INT64_MAX / (vnum * hden) > pow(10, vexp))
                                reduce_fraction(&d->sar_den, &d->sar_num, (int64_t)(hnum * vden * pow(10, hexp)),
                                                (int64_t)(vnum * hden * pow(10, vexp)), INT32_MAX);
                        }
                    }
                }
                SEEK(atom2_end);
            } while (atom_end - atom2_end >= 8);
        } else {
            search_range--;
        }
        SEEK(atom_end);
    }
#undef SEEK
    return 0;
}

static int parse_headers(Dec *d)
{
    const uint8_t *p = d->cs;
    for (int i = 0; i < 4; i++)
        d->cdef[i] = -1;                            /* jpeg2000dec.c:2834 */
    if (d->end - p >= 12 && rd32(p) == 12 && rd32(p + 4) == 0x6A502020u && rd32(p + 8) == 0x0D0A870Au) {
        int64_t pos = 12;                           /* JP2 / JPH file (:2842-2854) */
        /* "if (!jp2_find_codestream(s))" (:2846): an absurd box size makes it return an error code, which is not
         * zero -- the search for SOC simply goes on behind the box header that was read */
        const int r = jp2_find_codestream(d, &pos);
        if (!r)
            return ORACLE_ERR_INVALID;
        p = d->cs + pos;
    }
    while (d->end - p >= 3 && rd16(p) != 0xFF4F)
        p++;
    if (d->end - p < 2 || rd16(p) != 0xFF4F)
        return ORACLE_ERR_INVALID;
    p += 2;
    int cur_tile = -1, cur_tp = 0;
    const uint8_t *tp_end = NULL;
    for (;;) {
        if (d->end - p < 2)
            break;                                        /* "Missing EOC" is tolerated */
        unsigned marker = rd16(p);
        p += 2;
        if (marker >= 0xFF30 && marker <= 0xFF3F)
            continue;
        if (marker == 0xFF93) {                           /* SOD */
            if (!d->tile || cur_tile < 0 || tp_end < p)
                return ORACLE_ERR_INVALID;
            Tile *t = &d->tile[cur_tile];
            if (d->has_ppm) {                             /* Nppm, then that many bytes of headers (:2467-2472) */
                uint32_t n = 0;
                if (d->ppm_size - d->ppm_pos >= 4) {
                    n = rd32(d->ppm + d->ppm_pos);
                    d->ppm_pos += 4;
                }
                if (d->ppm_size - d->ppm_pos < n)
                    return ORACLE_ERR_INVALID;
                t->part[cur_tp].hp = d->ppm + d->ppm_pos;
                t->part[cur_tp].hend = t->part[cur_tp].hp + n;
                d->ppm_pos += n;
            }
            if (t->has_ppt && cur_tp == 0) {              /* :2474-2476: the packed headers go live here */
                t->ppt_pos = 0;
                t->ppt_live = t->ppt_size;
            }
            t->part[cur_tp].p = p;
            t->part[cur_tp].end = tp_end;
            if (cur_tp + 1 > t->nparts)
                t->nparts = cur_tp + 1;
            p = tp_end;
            continue;
        }
        if (marker == 0xFFD9)
            break;
        if (d->end - p < 2)
            break;
        int len = rd16(p);
        if (len < 2 || d->end - p < len)
            break;
        const uint8_t *seg = p + 2, *segend = p + len;
        int ret = 0;
        switch (marker) {
        case 0xFF51: {                                    /* SIZ */
            if (d->ncomp)
                return ORACLE_ERR_INVALID;
            if (segend - seg < 36)
                return ORACLE_ERR_INVALID;
            unsigned rsiz = rd16(seg);
            d->W = rd32(seg + 2); d->H = rd32(seg + 6); d->X0 = rd32(seg + 10); d->Y0 = rd32(seg + 14);
            d->TW = rd32(seg + 18); d->TH = rd32(seg + 22); d->TX0 = rd32(seg + 26); d->TY0 = rd32(seg + 30);
            int nc = rd16(seg + 34);
            /* av_image_check_size2(w, h, max_pixels = INT_MAX, AV_PIX_FMT_NONE) (imgutils.c:289-316) */
            if (d->W <= 0 || d->H <= 0 || 8LL * d->W + 1024 >= INT32_MAX ||
                (uint64_t)(8LL * d->W + 1024) * ((uint64_t)d->H + 128) >= (uint64_t)INT32_MAX || (int64_t)d->W * d->H > INT32_MAX)
                return ORACLE_ERR_PATCH;
            if (nc <= 0)
                return ORACLE_ERR_INVALID;
            if (nc > 4)
                return ORACLE_ERR_PATCH;
            d->rsiz = (int)rsiz;
            if (d->TX0 < 0 || d->TY0 < 0 || d->X0 < d->TX0 || d->Y0 < d->TY0 ||
                d->TW + (int64_t)d->TX0 <= d->X0 || d->TH + (int64_t)d->TY0 <= d->Y0)
                return ORACLE_ERR_INVALID;
            if (d->X0 >= d->W || d->Y0 >= d->H || d->TW <= 0 || d->TH <= 0)
                return ORACLE_ERR_INVALID;
            if (segend - seg < 36 + 3 * nc)
                return ORACLE_ERR_INVALID;
            d->ncomp = nc;
            for (int i = 0; i < nc; i++) {
                unsigned x = seg[36 + 3 * i];
                d->cbps[i] = (x & 0x7F) + 1;
                d->precision = imax(d->precision, d->cbps[i]);
                d->cdx[i] = seg[37 + 3 * i];
                d->cdy[i] = seg[38 + 3 * i];
                if (!d->cdx[i] || d->cdx[i] == 3 || d->cdx[i] > 4 || !d->cdy[i] || d->cdy[i] == 3 || d->cdy[i] > 4)
                    return ORACLE_ERR_INVALID;
            }
            d->ntx = ceildiv(d->W - d->TX0, d->TW);
            d->nty = ceildiv(d->H - d->TY0, d->TH);
            if ((int64_t)d->ntx * d->nty * 14 > d->end - d->cs)
                return ORACLE_ERR_EINVAL;
            d->tile = calloc((size_t)d->ntx * d->nty, sizeof(Tile));
            if (d->red && (d->X0 || d->Y0))
                return ORACLE_ERR_PATCH;                  /* :256-259 */
            int ow = ceildivpow2(d->W - d->X0, d->red), oh = ceildivpow2(d->H - d->Y0, d->red);   /* :314-320 */
            d->out_w = ceildiv(ow, d->cdx[0]);
            d->out_h = ceildiv(oh, d->cdy[0]);
            for (int i = 1; i < nc; i++) {
                d->out_w = imax(d->out_w, ceildiv(ow, d->cdx[i]));
                d->out_h = imax(d->out_h, ceildiv(oh, d->cdy[i]));
            }
            d->pix = choose_pix(d);
            if (!d->pix)
                return ORACLE_ERR_PATCH;
            seg += 36 + 3 * nc;
            break;
        }
        case 0xFF50: {                                    /* CAP (jpeg2000dec.c:424-489) */
            if (segend - seg < 6)
                return ORACLE_ERR_INVALID;
            uint32_t pcap = rd32(seg);
            seg += 4;
            unsigned ccap15 = 0;
            d->isHT = (pcap >> 17) & 1;
            for (int i = 0; i < 32; i++)
                if ((pcap >> (31 - i)) & 1) {
                    if (segend - seg < 2)
                        return ORACLE_ERR_INVALID;
                    if (i == 14)
                        ccap15 = rd16(seg);
                    seg += 2;
                }
            if (d->isHT) {
                d->cap_b14 = (ccap15 >> 14) & 3;
                if (d->cap_b14 == 2)
                    return ORACLE_ERR_EINVAL;
                if ((ccap15 >> 13) & 1)
                    return ORACLE_ERR_PATCH;
                d->cap_b05 = (ccap15 >> 5) & 1;
                d->cap_b11 = (ccap15 >> 11) & 1;
                d->cap_b12 = (ccap15 >> 12) & 1;
                int P = ccap15 & 31, B = !P ? 8 : P < 20 ? P + 8 : P < 31 ? 4 * (P - 19) + 27 : 74;
                if (B > 31)
                    return ORACLE_ERR_PATCH;
            }
            break;
        }
        case 0xFF52: {                                    /* COD */
            if (cur_tile >= 0 && d->isHT && !d->cap_b11)  /* HOMOGENEOUS set: no tile-part COD (:2514-2520) */
                return ORACLE_ERR_INVALID;
            if (segend - seg < 5)
                return ORACLE_ERR_INVALID;
            CodSty tmp;
            memset(&tmp, 0, sizeof(tmp));
            tmp.csty = seg[0];
            tmp.prog = seg[1];
            tmp.nlayers = rd16(seg + 2) & 0xFF;           /* uint8_t nlayers (jpeg2000.h:152): 256 layers are none */
            tmp.mct = seg[4];
            if (tmp.mct && d->ncomp < 3)
                return ORACLE_ERR_INVALID;
            if ((ret = parse_cox(d, seg + 5, segend, &tmp)) < 0)
                return ret;
            tmp.init = 1;
            for (int c = 0; c < d->ncomp; c++) {
                if (cur_tile >= 0) {
                    if (!d->tile[cur_tile].had_coc[c])
                        d->tile[cur_tile].cod[c] = tmp;
                } else if (!d->had_coc[c])
                    d->cod[c] = tmp;
            }
            seg += 10 + ((tmp.csty & 1) ? tmp.nres : 0);
            break;
        }
        case 0xFF53: {                                    /* COC */
            if (cur_tile >= 0 && d->isHT && !d->cap_b11)
                return ORACLE_ERR_INVALID;
            if (segend - seg < 2)
                return ORACLE_ERR_INVALID;
            int c = seg[0];
            if (c >= d->ncomp)
                return ORACLE_ERR_INVALID;
            CodSty *cs = cur_tile >= 0 ? &d->tile[cur_tile].cod[c] : &d->cod[c];
            int keep = cs->csty & 6;
            cs->csty = seg[1] | keep;
            if ((ret = parse_cox(d, seg + 2, segend, cs)) < 0)
                return ret;
            if (cur_tile >= 0)
                d->tile[cur_tile].had_coc[c] = 1;
            else
                d->had_coc[c] = 1;
            cs->init = 1;
            seg += 7 + ((cs->csty & 1) ? cs->nres : 0);
            break;
        }
        case 0xFF5C: {                                    /* QCD */
            if (cur_tile >= 0 && d->isHT && !d->cap_b11)
                return ORACLE_ERR_INVALID;
            QntSty tmp;
            memset(&tmp, 0, sizeof(tmp));
            if ((ret = parse_qcx(seg, len - 2, d->end - seg, &tmp)) < 0)
                return ret;
            seg += ret;
            for (int c = 0; c < d->ncomp; c++) {
                if (cur_tile >= 0) {
                    if (!d->tile[cur_tile].had_qcc[c])
                        d->tile[cur_tile].qnt[c] = tmp;
                } else if (!d->had_qcc[c])
                    d->qnt[c] = tmp;
            }
            break;
        }
        case 0xFF5D: {                                    /* QCC */
            if (cur_tile >= 0 && d->isHT && !d->cap_b11)
                return ORACLE_ERR_INVALID;
            if (segend - seg < 1)
                return ORACLE_ERR_INVALID;
            int c = seg[0];
            if (c >= d->ncomp)
                return ORACLE_ERR_INVALID;
            if (cur_tile >= 0)
                d->tile[cur_tile].had_qcc[c] = 1;
            else
                d->had_qcc[c] = 1;
            if ((ret = parse_qcx(seg + 1, len - 3, d->end - (seg + 1), cur_tile >= 0 ? &d->tile[cur_tile].qnt[c] : &d->qnt[c])) < 0)
                return ret;
            seg += 1 + ret;
            break;
        }
        case 0xFF5E: {                                    /* RGN (get_rgn, :641-672) */
            if (cur_tile >= 0 && d->isHT && !d->cap_b11)
                return ORACLE_ERR_INVALID;
            if (d->isHT && !d->cap_b12)                   /* "belongs to the RGNFREE set" (:2536-2540) */
                return ORACLE_ERR_INVALID;
            if (segend - seg < 3 || seg[1] != 0 || seg[0] >= d->ncomp)
                return ORACLE_ERR_INVALID;
            if (cur_tile >= 0 && cur_tp != 0)             /* only in the first tile-part of a tile */
                return ORACLE_ERR_INVALID;
            if (seg[2] > 30)
                return ORACLE_ERR_PATCH;
            if (cur_tile < 0)
                d->roi[seg[0]] = seg[2];
            else
                d->tile[cur_tile].comp[seg[0]].roi_shift = seg[2];
            seg += 3;
            break;
        }
        case 0xFF90: {                                    /* SOT (:822-873) */
            if (!d->tile || segend - seg < 8)
                return ORACLE_ERR_INVALID;
            int isot = rd16(seg);
            if (isot >= d->ntx * d->nty)
                return ORACLE_ERR_INVALID;
            uint32_t psot = rd32(seg + 2);
            int tpsot = seg[6];
            int64_t left = d->end - (seg + 8);
            /* get_sot measures against "bytes left - 2 + n + 2" AFTER it has read the 8 bytes (:845-850): a
             * tile-part has to end two bytes (the EOC marker) before the end of the packet */
            if (!psot)
                psot = (uint32_t)(left + len);
            if ((int64_t)psot > left + len)
                return ORACLE_ERR_INVALID;
            if (tpsot >= 32)
                return ORACLE_ERR_PATCH;
            cur_tile = isot;
            cur_tp = tpsot;
            if (!tpsot) {
                Tile *t = &d->tile[isot];
                t->first_seen = 1;
                memcpy(t->cod, d->cod, sizeof(t->cod));
                memcpy(t->qnt, d->qnt, sizeof(t->qnt));
                memcpy(t->poc, d->poc, sizeof(t->poc));
                t->npoc = d->npoc;
                t->poc_default = 1;
            }
            tp_end = seg + 8 + psot - len - 2;
            seg += 8;
            break;
        }
        case 0xFF5F: {                                    /* POC: get_poc (jpeg2000dec.c:760-818) */
            if (cur_tile >= 0 && d->isHT && !d->cap_b11)  /* :2558 refuses it for HOMOGENEOUS HT streams */
                return ORACLE_ERR_INVALID;
            if (d->end - seg < 5 || len < 2 + 7)
                return ORACLE_ERR_INVALID;
            int n = (len - 2) / 7;
            if (n > 32)
                return ORACLE_ERR_PATCH;
            struct { int RS, CS, LYE, RE, CE, P; } tmp[32];
            for (int i = 0; i < n; i++, seg += 7) {
                int RS = seg[0], CS = seg[1], LYE = (int)rd16(seg + 2), RE = seg[4], CE = seg[5], P = seg[6];
                if (!CE)
                    CE = 256;
                if (CE > d->ncomp)
                    CE = d->ncomp;
                if (RS >= RE || RE > 33 || CS >= CE || !LYE)
                    return ORACLE_ERR_INVALID;
                tmp[i].RS = RS; tmp[i].CS = CS; tmp[i].LYE = LYE; tmp[i].RE = RE; tmp[i].CE = CE; tmp[i].P = P;
            }
            if (cur_tile >= 0) {
                /* the tile's list is a copy of the main header's until its first own segment, which replaces it;
                 * further ones append (:806-816) */
                Tile *t = &d->tile[cur_tile];
                if (!t->npoc || t->poc_default)
                    t->npoc = 0;
                else if (t->npoc + n > 32)
                    return ORACLE_ERR_INVALID;
                memcpy(&t->poc[t->npoc], tmp, sizeof(tmp[0]) * (size_t)n);
                t->npoc += n;
                t->poc_default = 0;
            } else {
                if (d->npoc + n > 32)
                    return ORACLE_ERR_INVALID;
                memcpy(&d->poc[d->npoc], tmp, sizeof(tmp[0]) * (size_t)n);
                d->npoc += n;                             /* a further POC segment appends (:805-812) */
            }
            break;
        }
        case 0xFF60: {                                    /* PPM: main header only (:2597-2604) */
            if (cur_tile >= 0 || len < 3)
                return ORACLE_ERR_INVALID;
            const size_t n = (size_t)(segend - seg) - 1;  /* Zppm is skipped */
            d->ppm = realloc(d->ppm, d->ppm_size + n + 1);
            memcpy(d->ppm + d->ppm_size, seg + 1, n);
            d->ppm_size += n;
            d->has_ppm = 1;
            seg = segend;
            break;
        }
        case 0xFF61: {                                    /* PPT: first tile-part of a tile (:2605-2617, :982-1014) */
            if (d->has_ppm || (d->isHT && !d->cap_b11) || len < 3 || cur_tile < 0 || cur_tp != 0)
                return ORACLE_ERR_INVALID;
            Tile *t = &d->tile[cur_tile];
            const size_t n = (size_t)(segend - seg) - 1;
            t->ppt = realloc(t->ppt, t->ppt_size + n + 1);
            memcpy(t->ppt + t->ppt_size, seg + 1, n);
            t->ppt_size += n;
            t->has_ppt = 1;
            seg = segend;
            break;
        }
        case 0xFF55: {                                    /* TLM: only walked over (get_tlm, :901-934) */
            const unsigned stlm = d->end - seg > 1 ? seg[1] : 0;
            const int st = (stlm >> 4) & 3, sp = (stlm >> 6) & 1;
            if (st == 3)
                return ORACLE_ERR_INVALID;
            const int entry = (sp + 1) * 2 + st;
            seg += 2 + (int)(uint8_t)((len - 4) / entry) * entry;   /* the entry count lives in a uint8_t */
            break;
        }
        case 0xFF58:                                      /* PLT (get_plt, :936-955): the last byte ends a length */
            if (len < 4 || (segend[-1] & 0x80))
                return ORACLE_ERR_INVALID;
            seg = segend;
            break;
        case 0xFF63:                                      /* CRG (read_crg, :875-883) */
            if (d->ncomp * 4 != len - 2)
                return ORACLE_ERR_INVALID;
            seg = segend;
            break;
        default:                                          /* COM, PLM, CPF, unknown: skipped */
            seg = segend;
            break;
        }
        if (seg != segend)
            return -1;                                    /* declared length must equal what was consumed */
        p = segend;
    }
    return 0;
}

/* ---- geometry -------------------------------------------------------------------------------------- */

/* jpeg2000.c:214-272 -- same operations in the same types so that the float result is identical */
static void band_stepsize(Band *b, const CodSty *c, const QntSty *q, int bandno, int gbandno, int r, int cbps, int red)
{
    float f;
    if (q->style == 0) {
        f = 1;
    } else if (q->style <= 2) {
        union { uint32_t u; float f; } e;
        e.u = (uint32_t)((cbps - q->expn[gbandno]) + 127) << 23;
        f = e.f;
        f *= q->mant[gbandno] / 2048.0 + 1.0;
    } else {
        f = 0;                                            /* "Unknown quantization format" (jpeg2000.c:246-249) */
    }
    if (c->transform != ORACLE_DWT53) {
        int lband = 0;
        switch (bandno + (r > 0)) {
        case 1: case 2:
            f *= 0.812893066115961f * 2;
            lband = 1;
            break;
        case 3:
            f *= 0.812893066115961f * 0.812893066115961f * 4;
            break;
        }
        f *= pow(1.230174104914001f, 2 * (c->nres - red - r) + lband - 2);   /* nreslevels2decode (jpeg2000.c:263) */
    }
    if (f > (INT32_MAX >> 15))
        f = 0;
    b->fstep = f;
    b->istep = (int)floorf(f * (1 << 15));
}

/* init_tile (jpeg2000dec.c:1016-1070) + ff_jpeg2000_init_component / init_band / init_prec
 * (jpeg2000.c:274-577) */
static int init_tile(Dec *d, int tileno)
{
    Tile *t = &d->tile[tileno];
    /* a tile whose first tile-part never showed up has no coding style unless a COD of its own supplies one: the
     * per-component check on cs->init below (jpeg2000dec.c:1054-1055) fails the frame then */
    const int tx = tileno % d->ntx, ty = tileno / d->ntx;
    t->c.x0 = clip64((int64_t)tx * d->TW + d->TX0, d->X0, d->W);
    t->c.x1 = clip64((int64_t)(tx + 1) * d->TW + d->TX0, d->X0, d->W);
    t->c.y0 = clip64((int64_t)ty * d->TH + d->TY0, d->Y0, d->H);
    t->c.y1 = clip64((int64_t)(ty + 1) * d->TH + d->TY0, d->Y0, d->H);

    for (int c = 0; c < d->ncomp; c++) {
        Comp *comp = &t->comp[c];
        const CodSty *cs = &t->cod[c];
        const QntSty *qs = &t->qnt[c];
        if (!cs->init)
            return ORACLE_ERR_INVALID;
        if (d->isHT && !d->cap_b05 && !cs->transform)
            return ORACLE_ERR_INVALID;                    /* 9/7 inside an HTREV set */
        if (d->isHT && d->cap_b14 != (cs->cblk_style >> 6) && d->cap_b14 != 0)
            return ORACLE_ERR_INVALID;
        comp->c.x0 = ceildiv(t->c.x0, d->cdx[c]);
        comp->c.x1 = ceildiv(t->c.x1, d->cdx[c]);
        comp->c.y0 = ceildiv(t->c.y0, d->cdy[c]);
        comp->c.y1 = ceildiv(t->c.y1, d->cdy[c]);
        if (!comp->roi_shift)
            comp->roi_shift = d->roi[c];
        comp->cr.x0 = ceildivpow2(comp->c.x0, d->red);
        comp->cr.x1 = ceildivpow2(comp->c.x1, d->red);
        comp->cr.y0 = ceildivpow2(comp->c.y0, d->red);
        comp->cr.y1 = ceildivpow2(comp->c.y1, d->red);
        if (cs->nres - d->red <= 0)
            return ORACLE_ERR_EINVAL;
        if (cs->transform > ORACLE_DWT97_INT)             /* ff_jpeg2000_dwt_init: "default: return -1" (jpeg2000dwt.c:590) */
            return -1;
        const int cw = comp->cr.x1 - comp->cr.x0, ch = comp->cr.y1 - comp->cr.y0;
        if (cw > 32768 || ch > 32768)
            return ORACLE_ERR_PATCH;
        comp->data = calloc((size_t)cw * ch + 16, 4);
        comp->res = calloc((size_t)cs->nres, sizeof(Res));
        int gband = 0;
        for (int r = 0; r < cs->nres; r++) {
            Res *rl = &comp->res[r];
            const int decl = cs->nres - r;               /* N_L - r + 1 */
            rl->c.x0 = ceildivpow2(comp->c.x0, decl - 1);
            rl->c.x1 = ceildivpow2(comp->c.x1, decl - 1);
            rl->c.y0 = ceildivpow2(comp->c.y0, decl - 1);
            rl->c.y1 = ceildivpow2(comp->c.y1, decl - 1);
            rl->ppx = cs->ppx[r];
            rl->ppy = cs->ppy[r];
            rl->nbands = r == 0 ? 1 : 3;
            rl->npx = rl->c.x1 == rl->c.x0 ? 0 : ceildivpow2(rl->c.x1, rl->ppx) - (rl->c.x0 >> rl->ppx);
            rl->npy = rl->c.y1 == rl->c.y0 ? 0 : ceildivpow2(rl->c.y1, rl->ppy) - (rl->c.y0 >> rl->ppy);
            for (int b = 0; b < rl->nbands; b++, gband++) {
                Band *bd = &rl->band[b];
                band_stepsize(bd, cs, qs, b, gband, r, d->cbps[c], d->red);
                int bppx, bppy;
                if (r == 0) {
                    bd->c = rl->c;
                    bppx = rl->ppx;
                    bppy = rl->ppy;
                } else {
                    const int xo = (b + 1) & 1, yo = ((b + 1) >> 1) & 1;
                    bd->c.x0 = ceildivpow2((int)(comp->c.x0 - ((int64_t)xo << (decl - 1))), decl);
                    bd->c.x1 = ceildivpow2((int)(comp->c.x1 - ((int64_t)xo << (decl - 1))), decl);
                    bd->c.y0 = ceildivpow2((int)(comp->c.y0 - ((int64_t)yo << (decl - 1))), decl);
                    bd->c.y1 = ceildivpow2((int)(comp->c.y1 - ((int64_t)yo << (decl - 1))), decl);
                    bppx = rl->ppx - 1;
                    bppy = rl->ppy - 1;
                }
                bd->l2cw = imin(cs->l2cw, bppx);
                bd->l2ch = imin(cs->l2ch, bppy);
                const int nprec = rl->npx * rl->npy;
                bd->prec = calloc((size_t)imax(nprec, 1), sizeof(Prec));
                for (int pn = 0; pn < nprec; pn++) {
                    Prec *pr = &bd->prec[pn];
                    Rect pc;
                    pc.x0 = ((rl->c.x0 >> rl->ppx) + pn % rl->npx) * (1 << bppx);
                    pc.y0 = ((rl->c.y0 >> rl->ppy) + pn / rl->npx) * (1 << bppy);
                    pc.x1 = pc.x0 + (1 << bppx);
                    pc.y1 = pc.y0 + (1 << bppy);
                    pc.x0 = imax(pc.x0, bd->c.x0); pc.x1 = imin(pc.x1, bd->c.x1);
                    pc.y0 = imax(pc.y0, bd->c.y0); pc.y1 = imin(pc.y1, bd->c.y1);
                    pr->ncw = ceildivpow2(pc.x1, bd->l2cw) - (pc.x0 >> bd->l2cw);
                    pr->nch = ceildivpow2(pc.y1, bd->l2ch) - (pc.y0 >> bd->l2ch);
                    if (pr->ncw < 0 || pr->nch < 0)
                        pr->ncw = pr->nch = 0;             /* empty intersection (the reference would fail to allocate) */
                    pr->incl = tagtree_new(pr->ncw, pr->nch);
                    pr->zb = tagtree_new(pr->ncw, pr->nch);
                    const int n = pr->ncw * pr->nch;
                    pr->cb = calloc((size_t)imax(n, 1), sizeof(Cblk));
                    for (int k = 0; k < n; k++) {
                        Cblk *cb = &pr->cb[k];
                        int gx = ((pc.x0 >> bd->l2cw) << bd->l2cw) + ((k % pr->ncw) << bd->l2cw);
                        int gy = ((pc.y0 >> bd->l2ch) << bd->l2ch) + ((k / pr->ncw) << bd->l2ch);
                        cb->x0 = imax(gx, pc.x0); cb->x1 = imin(gx + (1 << bd->l2cw), pc.x1);
                        cb->y0 = imax(gy, pc.y0); cb->y1 = imin(gy + (1 << bd->l2ch), pc.y1);
                        /* Mallat placement: shift by the size of the next-lower resolution (jpeg2000.c:365-376) */
                        if ((b + !!r) & 1) {
                            int off = comp->res[r - 1].c.x1 - comp->res[r - 1].c.x0;
                            cb->x0 += off; cb->x1 += off;
                        }
                        if ((b + !!r) & 2) {
                            int off = comp->res[r - 1].c.y1 - comp->res[r - 1].c.y0;
                            cb->y0 += off; cb->y1 += off;
                        }
                        cb->lblock = 3;
                    }
                }
            }
        }
    }
    return 0;
}

static void free_tile(Dec *d, Tile *t)
{
    for (int c = 0; c < d->ncomp; c++) {
        Comp *comp = &t->comp[c];
        if (comp->res) {
            for (int r = 0; r < t->cod[c].nres; r++)
                for (int b = 0; b < comp->res[r].nbands; b++) {
                    Band *bd = &comp->res[r].band[b];
                    if (!bd->prec)
                        continue;
                    for (int pn = 0; pn < comp->res[r].npx * comp->res[r].npy; pn++) {
                        Prec *pr = &bd->prec[pn];
                        for (int k = 0; k < pr->ncw * pr->nch; k++)
                            free(pr->cb[k].data);
                        free(pr->cb);
                        free(pr->incl);
                        free(pr->zb);
                    }
                    free(bd->prec);
                }
            free(comp->res);
        }
        free(comp->data);
        comp->res = NULL;
        comp->data = NULL;
    }
}

/* ---- tier-2 ---------------------------------------------------------------------------------------- */

typedef struct PktState { Tile *t; int tp; } PktState;

/* select_stream (jpeg2000dec.c:1110-1133): continue in the current tile-part, hop over exhausted ones */
static void select_part(Dec *d, PktState *ps, int csty)
{
    Tile *t = ps->t;
    d->bp = t->part[ps->tp].p;
    d->bend = t->part[ps->tp].end;
    /* an exhausted or missing tile-part: on to the next slot, over gaps in the TPsot numbering as well, up to the
     * last slot of the array (:1117-1126) */
    while (d->bp >= d->bend && d->bit_index == 8 && ps->tp < 31) {
        ps->tp++;
        d->bp = t->part[ps->tp].p;
        d->bend = t->part[ps->tp].end;
    }
    if (!d->bp)
        d->bp = d->bend = d->end;
    if (csty & 2)                                         /* SOP */
        if (d->bend - d->bp >= 6 && rd32(d->bp) == 0xFF910004u)
            d->bp += 6;
}

/* select_header (jpeg2000dec.c:1099-1108): read the packet header from the tile-part's PPM run; once that is
 * used up the reference moves to the next tile-part and reads from its data */
static void select_header(Dec *d, PktState *ps)
{
    Tile *t = ps->t;
    d->bp = t->part[ps->tp].hp;
    d->bend = t->part[ps->tp].hend;
    if (!d->bp)
        d->bp = d->bend = d->end;
    if (d->bp >= d->bend && d->bit_index == 8 && ps->tp < 31) {
        ps->tp++;
        d->bp = t->part[ps->tp].p;
        d->bend = t->part[ps->tp].end;
        if (!d->bp)
            d->bp = d->bend = d->end;
    }
}

static int getnpasses(Dec *d)
{
    if (!hb_bits(d, 1)) return 1;
    if (!hb_bits(d, 1)) return 2;
    int n = hb_bits(d, 2);
    if (n != 3) return 3 + n;
    n = hb_bits(d, 5);
    if (n != 31) return 6 + n;
    return 37 + hb_bits(d, 7);
}

/* jpeg2000_decode_packet, HT-only paths (jpeg2000dec.c:1136-1542) */
static int decode_packet(Dec *d, PktState *ps, const CodSty *cs, Res *rl, int precno, int layno,
                         const int *expn, int nguard, int roi0)
{
    if (layno < rl->band[0].prec[precno].decoded_layers)
        return 0;
    rl->band[0].prec[precno].decoded_layers = layno + 1;
    Tile *pt = ps->t;
    if (d->has_ppm)                                       /* :1147-1153 */
        select_header(d, ps);
    else if (pt->has_ppt) {
        d->bp = pt->ppt + pt->ppt_pos;
        d->bend = pt->ppt + pt->ppt_live;
    } else
        select_part(d, ps, cs->csty);

    struct Inc { Cblk *cb; int nseg; int seg[100]; } *incs = NULL;
    int ninc = 0, capinc = 0;
    int ret = 0;

    if (!hb_bits(d, 1)) {
        hb_flush(d);
        goto finish;
    }
    for (int b = 0; b < rl->nbands; b++) {
        Band *bd = &rl->band[b];
        if (bd->c.x0 == bd->c.x1 || bd->c.y0 == bd->c.y1)
            continue;
        Prec *pr = &bd->prec[precno];
        for (int k = 0; k < pr->ncw * pr->nch; k++) {
            Cblk *cb = &pr->cb[k];
            int incl;
            if (!cb->incl) {
                cb->modes = cs->cblk_style;               /* :1177-1179 */
                if (cb->modes >= 0x40)
                    cb->plhd = 1;
                if (layno > 0)
                    tagtree_decode(d, pr->incl + k, 1);
                incl = tagtree_decode(d, pr->incl + k, layno + 1) == layno;
                if (incl) {
                    int zbp = tagtree_decode(d, pr->zb + k, 100);
                    int v = expn[b] + nguard - 1 - (zbp - roi0);
                    if (v < 0 || v > 30) {
                        ret = ORACLE_ERR_INVALID;
                        goto fail;
                    }
                    cb->incl = 1;
                    cb->zbp = zbp;
                    cb->lblock = 3;
                }
            } else {
                incl = hb_bits(d, 1);
            }
            if (!incl)
                continue;
            int newpasses = getnpasses(d);
            if (cb->npasses + newpasses >= 100) {
                ret = ORACLE_ERR_PATCH;
                goto fail;
            }
            int llen = 0;
            while (hb_bits(d, 1))
                llen++;
            if (cb->lblock + llen + ilog2(newpasses) > 16) {
                ret = ORACLE_ERR_PATCH;
                goto fail;
            }
            cb->lblock += llen;
            if (ninc == capinc) {
                capinc = capinc ? 2 * capinc : 64;
                incs = realloc(incs, sizeof(*incs) * capinc);
            }
            struct Inc *in = &incs[ninc++];
            in->cb = cb;
            in->nseg = 0;

            int segment_passes, next_segment_passes = 0, bits;
            unsigned segment_bytes = 0;
            /* a MIXED block (SPcod bit 7) that the length rules below unmask as a Part-1 block: HT bit off for good
             * (:1271-1274, :1303-1305, :1331-1333).  Part-1 blocks with bypass / termination on each pass are not restated. */
#define TURNS_PART1(cb) do { \
                if (cs->cblk_style & 0x05) { ret = ORACLE_ERR_PATCH; goto fail; } \
                (cb)->plhd = 0; (cb)->modes &= ~0x40; } while (0)
            if (cb->plhd) {                               /* :1256-1340 */
                int href = (cb->npasses + newpasses - 1) % 3;
                segment_passes = newpasses - href;
                int pass_bound = 2;
                bits = cb->lblock;
                if (segment_passes < 1) {
                    segment_passes = newpasses;
                    while (pass_bound <= segment_passes) { bits++; pass_bound += pass_bound; }
                    segment_bytes = hb_bits(d, bits);
                    if (segment_bytes && (cb->modes & 0x80))
                        TURNS_PART1(cb);
                } else {
                    while (pass_bound <= segment_passes) { bits++; pass_bound += pass_bound; }
                    segment_bytes = hb_bits(d, bits);
                    if (segment_bytes) {
                        if (!(cb->modes & 0x80) ||
                            (cb->lblock > 3 && segment_bytes > 1 && (segment_bytes >> (bits - 1)) == 0)) {
                            next_segment_passes = 2;      /* the first HT cleanup pass */
                            cb->plhd = 0;
                            cb->Lcup = segment_bytes;
                        } else {                          /* an original (Part-1) coding pass */
                            TURNS_PART1(cb);
                            segment_passes = newpasses;
                            while (pass_bound <= segment_passes) {
                                bits++;
                                pass_bound += pass_bound;
                                segment_bytes = (segment_bytes << 1) + hb_bits(d, 1);
                            }
                        }
                    } else {
                        segment_passes = newpasses;
                        if (pass_bound <= segment_passes) {
                            for (;;) {
                                bits++;
                                pass_bound += pass_bound;
                                segment_bytes = (segment_bytes << 1) + hb_bits(d, 1);
                                if (pass_bound > segment_passes)
                                    break;
                            }
                            if (segment_bytes && (cb->modes & 0x80))
                                TURNS_PART1(cb);
                        }
                    }
                }
            } else if (!(cb->modes & 0x40)) {
                /* Part-1 block: with neither restart nor bypass there is one codeword segment and one length
                 * of lblock + floor(log2(passes)) bits per packet (:1364-1368); the others are not restated */
                if (cs->cblk_style & (0x01 | 0x04)) {
                    ret = ORACLE_ERR_PATCH;
                    goto fail;
                }
                segment_bytes = (unsigned)hb_bits(d, cb->lblock + ilog2(newpasses));
                segment_passes = newpasses;
            } else {                                      /* :1341-1362 later contributions of an HT block */
                segment_passes = cb->npasses % 3;
                bits = 0;
                if (segment_passes == 0) {
                    segment_passes = 1;
                    next_segment_passes = 2;
                } else {
                    segment_passes = newpasses > 1 ? 3 - segment_passes : 1;
                    next_segment_passes = 1;
                    bits = ilog2(segment_passes);
                }
                bits += cb->lblock;
                segment_bytes = hb_bits(d, bits);
                cb->Lref += segment_bytes;
            }
            cb->npasses = (cb->npasses + segment_passes) & 0xFF;
            in->seg[in->nseg++] = segment_bytes;
            newpasses -= segment_passes & 0xFF;
            if ((cb->modes & 0x40) && !cb->plhd) {        /* :1404-1418 */
                while (newpasses > 0) {
                    segment_passes = newpasses > 1 ? next_segment_passes : 1;
                    next_segment_passes = 3 - next_segment_passes;
                    bits = cb->lblock + ilog2(segment_passes);
                    segment_bytes = hb_bits(d, bits);
                    newpasses -= segment_passes;
                    cb->Lref += segment_bytes;
                    cb->npasses = (cb->npasses + segment_passes) & 0xFF;
                    if (in->nseg < 100)
                        in->seg[in->nseg++] = segment_bytes;
                }
            } else {
                while (newpasses > 0) {                   /* :1419-1438 with no termination modes */
                    segment_bytes = hb_bits(d, cb->lblock);
                    newpasses -= 1;
                    cb->npasses = (cb->npasses + 1) & 0xFF;
                    if (in->nseg < 100)
                        in->seg[in->nseg++] = segment_bytes;
                }
            }
        }
    }
    hb_flush(d);
finish:
    if (cs->csty & 4)                                     /* EPH */
        if (d->bend - d->bp >= 2 && rd16(d->bp) == 0xFF92)
            d->bp += 2;
    if (d->has_ppm) {                                     /* save the header stream, go on in the data (:1474-1480) */
        pt->part[ps->tp].hp = d->bp;
        pt->part[ps->tp].hend = d->bend;
        select_part(d, ps, cs->csty);
    } else if (pt->has_ppt) {
        pt->ppt_pos = (size_t)(d->bp - pt->ppt);
        select_part(d, ps, cs->csty);
    }
    for (int i = 0; i < ninc; i++) {
        Cblk *cb = incs[i].cb;
        for (int s = 0; s < incs[i].nseg; s++) {
            int n = incs[i].seg[s];
            if (d->bend - d->bp < n) {
                ret = ORACLE_ERR_INVALID;
                goto fail;
            }
            if (cb->length + n + 8 > cb->cap) {
                cb->cap = imax(2 * cb->cap, cb->length + n + 8);
                cb->data = realloc(cb->data, (size_t)cb->cap);
            }
            memcpy(cb->data + cb->length, d->bp, (size_t)n);
            d->bp += n;
            cb->length = (cb->length + n) & 0xFFFF;       /* uint16_t length (jpeg2000.h:188): the bytes are all there,
                                                           * the count wraps, the next segment lands on top of them */
        }
    }
    ps->t->part[ps->tp].p = d->bp;
    free(incs);
    return 0;
fail:
    free(incs);
    return ret;
}

/* does (x, y) on the reference grid start a precinct of this resolution?  (jpeg2000dec.c:1700-1722) */
static int precinct_at(Dec *d, Tile *t, int c, int r, int x, int y, int cprl_rule)
{
    const CodSty *cs = &t->cod[c];
    Res *rl = &t->comp[c].res[r];
    const int red = cs->nres - 1 - r;
    if (cprl_rule) {                                      /* the CPRL loop uses a weaker test (:1642-1648) */
        int xc = x / d->cdx[c], yc = y / d->cdy[c];
        if (yc % (1LL << (rl->ppy + red)) && y != t->c.y0)
            return -1;
        if (xc % (1LL << (rl->ppx + red)) && x != t->c.x0)
            return -1;
        unsigned px = (unsigned)(ceildivpow2(xc, red) >> rl->ppx) - (unsigned)(ceildivpow2(t->comp[c].c.x0, red) >> rl->ppx);
        unsigned py = (unsigned)(ceildivpow2(yc, red) >> rl->ppy) - (unsigned)(ceildivpow2(t->comp[c].c.y0, red) >> rl->ppy);
        if (px >= (unsigned)rl->npx || py >= (unsigned)rl->npy)
            return -2;
        return (int)(px + rl->npx * py);
    }
    int64_t trx0 = ((int64_t)t->c.x0 + ((int64_t)d->cdx[c] << red) - 1) / ((int64_t)d->cdx[c] << red);
    int64_t try0 = ((int64_t)t->c.y0 + ((int64_t)d->cdy[c] << red) - 1) / ((int64_t)d->cdy[c] << red);
    if (!(y % ((uint64_t)d->cdy[c] << (rl->ppy + red)) == 0 ||
          (y == t->c.y0 && ((try0 << red) % (1LL << (red + rl->ppy))))))
        return -1;
    if (!(x % ((uint64_t)d->cdx[c] << (rl->ppx + red)) == 0 ||
          (x == t->c.x0 && ((trx0 << red) % (1LL << (red + rl->ppx))))))
        return -1;
    unsigned px = (unsigned)(((int64_t)x + ((int64_t)d->cdx[c] << red) - 1) / ((int64_t)d->cdx[c] << red) >> rl->ppx);
    unsigned py = (unsigned)(((int64_t)y + ((int64_t)d->cdy[c] << red) - 1) / ((int64_t)d->cdy[c] << red) >> rl->ppy);
    px -= (unsigned)(ceildivpow2(t->comp[c].c.x0, red) >> rl->ppx);
    py -= (unsigned)(ceildivpow2(t->comp[c].c.y0, red) >> rl->ppy);
    if (px >= (unsigned)rl->npx || py >= (unsigned)rl->npy)
        return -2;                                        /* a precinct position, but none exists there */
    return (int)(px + rl->npx * py);
}

#define PACKET(c, r, pn, l)                                                                           \
    do {                                                                                              \
        int e_ = decode_packet(d, &ps, &t->cod[c], &t->comp[c].res[r], pn, l,                         \
                               t->qnt[c].expn + ((r) ? 3 * ((r) - 1) + 1 : 0), t->qnt[c].nguard,      \
                               t->comp[0].roi_shift);                                                 \
        if (e_ < 0)                                                                                   \
            return e_;                                                                                \
    } while (0)

/* jpeg2000_decode_packets_po_iteration (jpeg2000dec.c:1544-1828): resolutions [RS, RE), components [CS, CE),
 * layers [0, LYE) in progression P.  The resolution loops of LRCP/RLCP/RPCL stop at the first level that no
 * component of the range owns ("ok_reslevel"). */
static int po_iteration(Dec *d, Tile *t, PktState *pps, int RS, int CS, int LYE, int RE, int CE, int P)
{
#define ps (*pps)
    int ok;
    switch (P) {
    case 0: /* LRCP */
        for (int l = 0; l < LYE; l++) {
            ok = 1;
            for (int r = RS; ok && r < RE; r++) {
                ok = 0;
                for (int c = CS; c < CE; c++) {
                    if (r >= t->cod[c].nres)
                        continue;
                    ok = 1;
                    for (int pn = 0; pn < t->comp[c].res[r].npx * t->comp[c].res[r].npy; pn++)
                        PACKET(c, r, pn, l);
                }
            }
        }
        break;
    case 1: /* RLCP */
        ok = 1;
        for (int r = RS; ok && r < RE; r++) {
            ok = 0;
            for (int l = 0; l < LYE; l++)
                for (int c = CS; c < CE; c++) {
                    if (r >= t->cod[c].nres)
                        continue;
                    ok = 1;
                    for (int pn = 0; pn < t->comp[c].res[r].npx * t->comp[c].res[r].npy; pn++)
                        PACKET(c, r, pn, l);
                }
        }
        break;
    case 2: /* RPCL */
        ok = 1;
        for (int r = RS; ok && r < RE; r++) {
            int sx = 30, sy = 30;
            ok = 0;
            for (int c = CS; c < CE; c++)
                if (r < t->cod[c].nres) {
                    int red = t->cod[c].nres - 1 - r;
                    sx = imin(sx, t->comp[c].res[r].ppx + red);
                    sy = imin(sy, t->comp[c].res[r].ppy + red);
                }
            sx = 1 << sx;
            sy = 1 << sy;
            for (int y = t->c.y0; y < t->c.y1; y = (y / sy + 1) * sy)
                for (int x = t->c.x0; x < t->c.x1; x = (x / sx + 1) * sx)
                    for (int c = CS; c < CE; c++) {
                        if (r >= t->cod[c].nres)
                            continue;
                        int pn = precinct_at(d, t, c, r, x, y, 0);
                        if (pn == -1)
                            continue;
                        ok = 1;                           /* :1730, ahead of the range test */
                        if (pn < 0)
                            continue;
                        for (int l = 0; l < LYE; l++)
                            PACKET(c, r, pn, l);
                    }
        }
        break;
    case 3: { /* PCRL */
        int sx = 32, sy = 32;
        for (int c = CS; c < CE; c++)
            for (int r = RS; r < imin(t->cod[c].nres, RE); r++) {
                int red = t->cod[c].nres - 1 - r;
                sx = imin(sx, t->comp[c].res[r].ppx + red);
                sy = imin(sy, t->comp[c].res[r].ppy + red);
            }
        if (sx >= 31 || sy >= 31)
            return ORACLE_ERR_PATCH;
        sx = 1 << sx;
        sy = 1 << sy;
        for (int y = t->c.y0; y < t->c.y1; y = (y / sy + 1) * sy)
            for (int x = t->c.x0; x < t->c.x1; x = (x / sx + 1) * sx)
                for (int c = CS; c < CE; c++)
                    for (int r = RS; r < imin(t->cod[c].nres, RE); r++) {
                        int pn = precinct_at(d, t, c, r, x, y, 0);
                        if (pn < 0)
                            continue;
                        for (int l = 0; l < LYE; l++)
                            PACKET(c, r, pn, l);
                    }
        break;
    }
    case 4: /* CPRL */
        for (int c = CS; c < CE; c++) {
            int sx = 32, sy = 32;
            const int rend = imin(t->cod[c].nres, RE);
            if (RS >= rend)
                continue;
            for (int r = RS; r < rend; r++) {
                int red = t->cod[c].nres - 1 - r;
                sx = imin(sx, t->comp[c].res[r].ppx + red);
                sy = imin(sy, t->comp[c].res[r].ppy + red);
            }
            if (sx >= 31 || sy >= 31)
                return ORACLE_ERR_PATCH;
            sx = 1 << sx;
            sy = 1 << sy;
            for (int y = t->c.y0; y < t->c.y1; y = (y / sy + 1) * sy)
                for (int x = t->c.x0; x < t->c.x1; x = (x / sx + 1) * sx)
                    for (int r = RS; r < rend; r++) {
                        int pn = precinct_at(d, t, c, r, x, y, 1);
                        if (pn < 0)
                            continue;
                        for (int l = 0; l < LYE; l++)
                            PACKET(c, r, pn, l);
                    }
        }
        break;
    default:
        break;
    }
    return 0;
#undef ps
}

/* jpeg2000_decode_packets (jpeg2000dec.c:1835-1869): one iteration per POC entry, else the COD progression */
static int decode_packets(Dec *d, Tile *t)
{
    PktState ps = { t, 0 };
    const int nlayers = t->cod[0].nlayers;
    d->bit_index = 8;
    if (!t->npoc)
        return po_iteration(d, t, &ps, 0, 0, nlayers, 33, d->ncomp, t->cod[0].prog);
    for (int i = 0; i < t->npoc; i++) {
        int ret = po_iteration(d, t, &ps, t->poc[i].RS, t->poc[i].CS, imin(t->poc[i].LYE, nlayers), t->poc[i].RE,
                               imin(t->poc[i].CE, d->ncomp), t->poc[i].P);
        if (ret < 0)
            return ret;
    }
    return 0;
}

/* ---- the hot path, per tile ------------------------------------------------------------------------- */

/* D1-D3 (jpeg2000dec.c:2098-2181): t1 = sign-magnitude block, scatter into the component plane */
static void dequantise(const int32_t *t1, int bw, int bh, void *plane, int W, int x, int y,
                       int transform, const Band *band, int M_b)
{
    const int downshift = 31 - M_b;
    if (transform == ORACLE_DWT97) {
        float fscale = band->fstep;
        fscale /= (float)(1 << downshift);
        float *dst = plane;
        for (int j = 0; j < bh; j++)
            for (int i = 0; i < bw; i++) {
                int32_t v = t1[j * bw + i];
                if (v < 0)
                    v = -(v & INT32_MAX);
                dst[(size_t)W * (y + j) + x + i] = (float)v * fscale;
            }
    } else if (transform == ORACLE_DWT97_INT) {
        float fscale = band->fstep;
        fscale /= (float)(1 << downshift);
        fscale *= (float)(1 << 6);
        fscale *= (float)(1 << (16 + 8));
        int scale = (int)(fscale + 0.5);
        int32_t *dst = plane;
        for (int j = 0; j < bh; j++)
            for (int i = 0; i < bw; i++) {
                int32_t v = t1[j * bw + i];
                if (v < 0)
                    v = -(v & INT32_MAX);
                v = (v + (1 << 5)) >> 6;
                int64_t m = v * (int64_t)scale;
                /* RSHIFT(m, 16): libavutil/common.h, rounds half away from zero */
                dst[(size_t)W * (y + j) + x + i] = (int32_t)(m > 0 ? (m + (1 << 15)) >> 16 : (m + (1 << 15) - 1) >> 16);
            }
    } else {
        int32_t *dst = plane;
        for (int j = 0; j < bh; j++)
            for (int i = 0; i < bw; i++) {
                int32_t v = t1[j * bw + i];
                if (v < 0)
                    v = -((v & INT32_MAX) >> downshift);
                else
                    v >>= downshift;
                if (band->istep != 32768)
                    v = (int32_t)((v * (int64_t)band->istep) / 65536);
                dst[(size_t)W * (y + j) + x + i] = v;
            }
    }
}

/* H0 tile_codeblocks + M mct_decode + P1 write_frame */
static int decode_tile(Dec *d, Tile *t, uint8_t *const planes[4], const int strides[4])
{
    int32_t *t1 = malloc(sizeof(int32_t) * 4096);
    for (int c = 0; c < d->ncomp; c++) {
        Comp *comp = &t->comp[c];
        const CodSty *cs = &t->cod[c];
        const QntSty *qs = &t->qnt[c];
        const int W = comp->cr.x1 - comp->cr.x0;
        int coded = 0, subband = 0;
        for (int r = 0; r < cs->nres - d->red; r++) {      /* nreslevels2decode (:2230) */
            Res *rl = &comp->res[r];
            for (int b = 0; b < rl->nbands; b++, subband++) {
                Band *bd = &rl->band[b];
                const int M_b = qs->expn[subband] + qs->nguard - 1;
                if (bd->c.x0 == bd->c.x1 || bd->c.y0 == bd->c.y1)
                    continue;
                if (M_b >= 31) {
                    free(t1);
                    return ORACLE_ERR_PATCH;
                }
                for (int pn = 0; pn < rl->npx * rl->npy; pn++) {
                    Prec *pr = &bd->prec[pn];
                    for (int k = 0; k < pr->ncw * pr->nch; k++) {
                        Cblk *cb = &pr->cb[k];
                        const int bw = cb->x1 - cb->x0, bh = cb->y1 - cb->y0;
                        static const uint8_t none[8];
                        if (!(cb->modes & 0x40)) {        /* Part-1 block (decode_cblk, :2264-2275); cb->modes is 0 for a
                                                           * block no packet ever reached, as in the reference */
                            if (!cb->length)
                                continue;
                            if (cs->cblk_style & 0x05) {  /* bypass / termination on each pass: not restated */
                                free(t1);
                                return ORACLE_ERR_PATCH;
                            }
                            const int nonzerobits = M_b - (cb->zbp - t->comp[0].roi_shift);
                            if (oracle_mq_block_style(cb->data ? cb->data : none, cb->length, cb->npasses, nonzerobits, bw,
                                                      bh, b + (r > 0), M_b, comp->roi_shift, cs->cblk_style, t1)) {
                                coded = 1;
                                d->nblocks++;
                                dequantise(t1, bw, bh, comp->data, W, cb->x0 - bd->c.x0, cb->y0 - bd->c.y0,
                                           cs->transform, bd, M_b);
                            }
                            continue;
                        }
                        int ret = oracle_ht_block(cb->data ? cb->data : none, cb->length, cb->npasses, cb->Lcup,
                                                  cb->Lref, cb->zbp, bw, bh, M_b, comp->roi_shift,
                                                  cs->cblk_style & 8, t1);
                        if (!ret)
                            continue;
                        coded = 1;
                        d->nblocks++;
                        dequantise(t1, bw, bh, comp->data, W, cb->x0 - bd->c.x0, cb->y0 - bd->c.y0,
                                   cs->transform, bd, M_b);
                    }
                }
            }
        }
        if (coded) {
            int border[4] = { comp->cr.x0, comp->cr.x1, comp->cr.y0, comp->cr.y1 };
            oracle_idwt(cs->transform, border, cs->nres - d->red - 1, comp->data);
        }
    }
    free(t1);

    if (t->cod[0].mct) {                                  /* mct_decode (jpeg2000dec.c:2183-2209) */
        int ok = 1;
        for (int i = 1; i < 3; i++)
            ok &= t->cod[0].transform == t->cod[i].transform && !memcmp(&t->comp[0].cr, &t->comp[i].cr, sizeof(Rect));
        if (ok)
            oracle_mct(t->cod[0].transform, t->comp[0].data, t->comp[1].data, t->comp[2].data,
                       (t->comp[0].cr.x1 - t->comp[0].cr.x0) * (t->comp[0].cr.y1 - t->comp[0].cr.y0));
    }

    /* write_frame_8/16 (jpeg2000dec.c:2301-2366) */
    const int planar = d->pix == ORACLE_PIX_PLANAR8 || d->pix == ORACLE_PIX_PLANAR16;
    const int wide = d->precision > 8;
    const int pixelsize = planar ? 1 : d->ncomp;
    const int msb16 = d->pix == ORACLE_PIX_RGB48 || d->pix == ORACLE_PIX_RGBA64 || d->pix == ORACLE_PIX_GRAY16;
    const int precision = !wide ? 8 : msb16 ? 16 : d->precision;
    for (int c = 0; c < d->ncomp; c++) {
        Comp *comp = &t->comp[c];
        const int cbps = d->cbps[c];
        const int plane = planar ? (d->cdef[c] ? d->cdef[c] - 1 : d->ncomp - 1) : 0;    /* :2325-2326 */
        const int w = comp->cr.x1 - ceildiv(d->X0, d->cdx[c]);
        const int h = comp->cr.y1 - ceildiv(d->Y0, d->cdy[c]);
        const float *fp = comp->data;
        const int32_t *ip = comp->data;
        size_t k = 0;
        for (int y = comp->cr.y0 - ceildiv(d->Y0, d->cdy[c]); y < h; y++) {
            uint8_t *line = planes[plane] + (size_t)y * strides[plane];
            for (int x = comp->cr.x0 - ceildiv(d->X0, d->cdx[c]); x < w; x++, k++) {
                int val = t->cod[c].transform == ORACLE_DWT97 ? (int)lrintf(fp[k]) : ip[k];
                val += 1 << (cbps - 1);
                val = clip(val, 0, (1 << cbps) - 1);
                val <<= precision - cbps;
                size_t idx = (size_t)x * pixelsize + (planar ? 0 : c);
                if (wide)
                    ((uint16_t *)line)[idx] = (uint16_t)val;
                else
                    line[idx] = (uint8_t)val;
            }
        }
    }
    return 0;
}

static void free_dec(Dec *d)
{
    if (d->tile) {
        for (int i = 0; i < d->ntx * d->nty; i++) {
            free_tile(d, &d->tile[i]);
            free(d->tile[i].ppt);
        }
        free(d->tile);
    }
    free(d->ppm);
}

int oracle_decode(const uint8_t *cs, int64_t len, int bitexact, int64_t *info, uint8_t *out, int64_t cap)
{
    Dec d;
    memset(&d, 0, sizeof(d));
    d.cs = cs;
    d.end = cs + len;
    d.bitexact = bitexact & 1;                      /* the argument carries flags: bit 0 BITEXACT, bits 8-11 lowres */
    d.red = (bitexact >> 8) & 15;
    int ret = parse_headers(&d);
    if (ret < 0)
        goto end;
    if (!d.tile || !d.ncomp) {
        ret = ORACLE_ERR_INVALID;
        goto end;
    }
    /* output planes, tightly packed */
    {
        const int bps = d.precision > 8 ? 2 : 1;
        const int planar = d.pix == ORACLE_PIX_PLANAR8 || d.pix == ORACLE_PIX_PLANAR16;
        uint8_t *planes[4] = { 0 };
        int strides[4] = { 0 };
        int64_t pos = 0;
        if (!planar) {
            strides[0] = d.out_w * d.ncomp * bps;
            planes[0] = out;
            pos = (int64_t)strides[0] * d.out_h;
        } else {
            for (int c = 0; c < d.ncomp; c++) {
                const int chroma = c == 1 || c == 2;    /* planes of the AVPixelFormat, not of the components */
                int pw = chroma ? (d.out_w + (1 << d.chroma_lw) - 1) >> d.chroma_lw : d.out_w;
                int ph = chroma ? (d.out_h + (1 << d.chroma_lh) - 1) >> d.chroma_lh : d.out_h;
                strides[c] = pw * bps;
                planes[c] = out + pos;
                pos += (int64_t)strides[c] * ph;
            }
        }
        const int64_t pal_at = pos;
        if (!strcmp(g_pix_name, "pal8"))            /* data[1] of a PAL8 frame: 256 native-endian words (:2900-2901) */
            pos += 1024;
        if (pos > cap) {
            ret = ORACLE_ERR_NOMEM;
            goto end;
        }
        memset(out, 0, (size_t)pos);
        if (pos > pal_at)
            memcpy(out + pal_at, d.palette, 1024);
        for (int x = 0; x < d.ncomp; x++)           /* default channel definitions (:2883-2892) */
            if (d.cdef[x] < 0) {
                for (x = 0; x < d.ncomp; x++)
                    d.cdef[x] = x + 1;
                if ((d.ncomp & 1) == 0)
                    d.cdef[d.ncomp - 1] = 0;
                break;
            }
        for (int i = 0; i < d.ntx * d.nty; i++) {
            if ((ret = init_tile(&d, i)) < 0)
                goto end;
            if ((ret = decode_packets(&d, &d.tile[i])) < 0)
                goto end;
        }
        for (int i = 0; i < d.ntx * d.nty; i++)
            if ((ret = decode_tile(&d, &d.tile[i], planes, strides)) < 0)
                goto end;
        info[0] = d.out_w;
        info[1] = d.out_h;
        info[2] = d.pix;
        info[3] = pos;
        info[4] = d.nblocks;
        info[5] = d.ncomp;
        info[6] = d.precision;
        info[7] = d.sar_num && d.sar_den ? ((int64_t)d.sar_num << 32) | (uint32_t)d.sar_den : 0;   /* :2867-2868 */
        ret = 0;
    }
end:
    free_dec(&d);
    return ret;
}

/* ---- CPU baseline ("port") -------------------------------------------------------------------------- */
typedef struct Job {
    const uint8_t *const *cs;
    const int64_t *len;
    int nframes, first, step, reps, bitexact, errors;
} Job;

static void *bench_worker(void *arg)
{
    Job *j = arg;
    int64_t info[8];
    size_t cap = 0;
    uint8_t *out = NULL;
    for (int r = 0; r < j->reps; r++)
        for (int i = j->first; i < j->nframes; i += j->step) {
            if (!out) {
                cap = 1u << 20;
                out = malloc(cap);
            }
            int ret = oracle_decode(j->cs[i], j->len[i], j->bitexact, info, out, (int64_t)cap);
            while (ret == ORACLE_ERR_NOMEM && cap < ((size_t)1 << 32)) {
                cap *= 4;
                out = realloc(out, cap);
                ret = oracle_decode(j->cs[i], j->len[i], j->bitexact, info, out, (int64_t)cap);
            }
            if (ret < 0)
                j->errors++;
        }
    free(out);
    return NULL;
}

double oracle_decode_bench(const uint8_t *const *cs, const int64_t *len, int nframes, int threads,
                           int reps, int bitexact)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    Job jobs[256];
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (Job){ cs, len, nframes, t, threads, reps, bitexact, 0 };
        pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
    }
    int errors = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        errors += jobs[t].errors;
    }
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (errors)
        return -1.0;
    return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}
