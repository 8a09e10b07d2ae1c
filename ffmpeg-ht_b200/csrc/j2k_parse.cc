/*
 * j2k_parse.cc -- codestream -> flat frame plan (host, serial work).  See j2k_parse.h.
 *
 * Reference behaviour restated here (file:line in libavcodec/ of sandflow/ffmpeg-ht):
 *   SIZ jpeg2000dec.c:197-422   CAP :424-489   COD/COC :492-641   QCD/QCC :676-758   RGN :643-674
 *   SOT :822-873   main loop / SOD bookkeeping :2425-2637   pixel format choice :134-193, :335-363
 *   init_tile :1016-1070   component/band/precinct/code-block geometry jpeg2000.c:274-577
 *   step sizes jpeg2000.c:214-272   packet headers jpeg2000dec.c:1136-1542 (HT paths)
 *   progression orders :1544-1833
 * Unsupported on purpose (B200J2K_ERR_PATCHWELCOME): Part-1 blocks with bypass / termination on each pass -- SURVEY.md section 8(f) "next".
 */
#include "j2k_parse.h"

#include <cmath>
#include <cstring>

#include "../../include/b200j2k.h"

namespace b200j2k {

namespace {

inline int ceildiv(int a, int b) { return (int)(((int64_t)a + b - 1) / b); }
inline int ceildivpow2(int a, int b) { return (int)(-((-(int64_t)a) >> b)); }
inline int imin(int a, int b) { return a < b ? a : b; }
inline int imax(int a, int b) { return a > b ? a : b; }
inline int clip64(int64_t v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : (int)v; }
inline int ilog2(unsigned v) { int n = 0; while (v >>= 1) n++; return n; }
inline unsigned rd16(const uint8_t *p) { return (unsigned)(p[0] << 8) | p[1]; }
inline uint32_t rd32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int parse_cox(const uint8_t *p, const uint8_t *end, CodSty *c, int flags)
{
    if (end - p < 5)
        return B200J2K_ERR_INVALIDDATA;
    c->nres = p[0] + 1;
    if (c->nres > J2K_MAX_LEVELS + 1)                     /* JPEG2000_MAX_RESLEVELS = 34: up to 32 decomposition levels */
        return B200J2K_ERR_INVALIDDATA;
    /* lowres (jpeg2000dec.c:509-520): more levels dropped than the stream has is an error */
    if (c->nres <= B200J2K_FLAG_GET_LOWRES(flags))
        return B200J2K_ERR_EINVAL;
    c->l2cw = (p[1] & 15) + 2;
    c->l2ch = (p[2] & 15) + 2;
    if (c->l2cw > 10 || c->l2ch > 10 || c->l2cw + c->l2ch > 12)
        return B200J2K_ERR_INVALIDDATA;
    c->cblk_style = p[3];
    c->transform = p[4];                                  /* an unknown one fails when a tile is set up with it */
    if ((flags & B200J2K_FLAG_BITEXACT) && c->transform == J2K_DWT97)
        c->transform = J2K_DWT97_INT;
    p += 5;
    if (c->csty & 1) {
        for (int i = 0; i < c->nres; i++) {
            if (p >= end)
                return B200J2K_ERR_INVALIDDATA;
            c->ppx[i] = *p & 15;
            c->ppy[i] = *p >> 4;
            p++;
            if (i && (!c->ppx[i] || !c->ppy[i]))
                return B200J2K_ERR_INVALIDDATA;
        }
    } else {
        memset(c->ppx, 15, sizeof(c->ppx));
        memset(c->ppy, 15, sizeof(c->ppy));
    }
    return 0;
}

/* Sqcd / Sqcc and the step sizes behind it (get_qcx, jpeg2000dec.c:675-717).  n: bytes the segment declares for
 * this part, avail: bytes to the end of the packet -- the reference reads by what the STYLE needs, checked against
 * the packet only, and the caller then compares what was consumed with the declared length (:2629).
 * -> bytes consumed, or an error. */
int parse_qcx(const uint8_t *p, int n, int64_t avail, QntSty *q)
{
    if (avail < 1)
        return B200J2K_ERR_INVALIDDATA;
    q->nguard = p[0] >> 5;
    q->style = p[0] & 31;
    p++;
    avail--;
    if (q->style == 0) {
        const int m = n - 1;
        if (avail < m || m > 99)                          /* JPEG2000_MAX_DECLEVELS * 3 */
            return B200J2K_ERR_INVALIDDATA;
        for (int i = 0; i < m; i++)
            q->expn[i] = p[i] >> 3;
        return 1 + (m > 0 ? m : 0);
    }
    if (q->style == 1) {
        if (avail < 2)
            return B200J2K_ERR_INVALIDDATA;
        unsigned x = rd16(p);
        q->expn[0] = x >> 11;
        q->mant[0] = x & 0x7FF;
        for (int i = 1; i < 99; i++) {
            q->expn[i] = (uint8_t)imax(0, q->expn[0] - (i - 1) / 3);
            q->mant[i] = q->mant[0];
        }
        return 3;
    }
    /* style 2, and every undefined one (their bands get a step size of zero, band_stepsize) */
    const int m = (n - 1) >> 1;
    if (avail < 2 * m || m > 99)
        return B200J2K_ERR_INVALIDDATA;
    for (int i = 0; i < m; i++) {
        unsigned x = rd16(p + 2 * i);
        q->expn[i] = x >> 11;
        q->mant[i] = x & 0x7FF;
    }
    return 1 + (m > 0 ? 2 * m : 0);
}

/* ---- pixel format selection (get_siz, jpeg2000dec.c:134-193, :328-412) --------------------------------
 * The reference walks a list of AVPixelFormats chosen by the JP2 colour specification (16 sRGB, 17 grey,
 * 18 sYCC, anything else: all lists) and takes the first one pix_fmt_match() accepts.  The descriptors
 * below are the fields of libavutil's AVPixFmtDescriptor that pix_fmt_match() reads. */
struct PixDesc {
    const char *name;       /* libavutil name */
    uint8_t klass;          /* enum b200j2k_pix_fmt: the memory layout the pack kernels write */
    uint8_t ncomp, depth, lw, lh, pal;
};
#define P8 B200J2K_PIX_PLANAR8
#define P16 B200J2K_PIX_PLANAR16
const PixDesc RGB_FMTS[] = {
    { "pal8", B200J2K_PIX_GRAY8, 1, 8, 0, 0, 1 },   /* index plane laid out like gray8 + the palette as plane 1 */
    { "rgb24", B200J2K_PIX_RGB24, 3, 8, 0, 0, 0 }, { "rgba", B200J2K_PIX_RGBA, 4, 8, 0, 0, 0 },
    { "rgb48le", B200J2K_PIX_RGB48, 3, 16, 0, 0, 0 }, { "rgba64le", B200J2K_PIX_RGBA64, 4, 16, 0, 0, 0 },
};
const PixDesc GRAY_FMTS[] = {
    { "gray", B200J2K_PIX_GRAY8, 1, 8, 0, 0, 0 }, { "ya8", B200J2K_PIX_YA8, 2, 8, 0, 0, 0 },
    { "gray16le", B200J2K_PIX_GRAY16, 1, 16, 0, 0, 0 }, { "ya16le", B200J2K_PIX_YA16, 2, 16, 0, 0, 0 },
};
const PixDesc YUV_FMTS[] = {
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
const PixDesc XYZ_FMTS[] = { { "xyz12le", B200J2K_PIX_RGB48, 3, 12, 0, 0, 0 } };
#undef P8
#undef P16

int ilog2c(int v) { return v >= 4 ? 2 : v >= 2 ? 1 : 0; }     /* cdx, cdy are 1, 2 or 4 */

/* pix_fmt_match (jpeg2000dec.c:134-166) */
bool pix_match(const PixDesc &d, const Header &h)
{
    if (d.ncomp != h.ncomp || d.depth < h.precision)
        return false;
    bool ok = true;
    switch (h.ncomp) {
    case 4:
        ok = ok && h.cdx[3] == 1 && h.cdy[3] == 1;
        /* fall through */
    case 3:
        ok = ok && ilog2c(h.cdx[2]) == d.lw && ilog2c(h.cdy[2]) == d.lh;
        /* fall through */
    case 2:
        ok = ok && ilog2c(h.cdx[1]) == d.lw && ilog2c(h.cdy[1]) == d.lh;
        /* fall through */
    case 1:
        ok = ok && h.cdx[0] == 1 && h.cdy[0] == 1 && d.pal == (h.pal8 ? 1 : 0);
    }
    return ok;
}

template <size_t N> const PixDesc *first_match(const PixDesc (&list)[N], const Header &h)
{
    for (const PixDesc &d : list)
        if (pix_match(d, h))
            return &d;
    return nullptr;
}

/* -> layout class (0 = no format: AVERROR_PATCHWELCOME); *name = the libavutil pixel format name */
int choose_pix(Header &h, const char **name)
{
    const PixDesc *d = nullptr;
    static const PixDesc YUVA420P_FALLBACK = { "yuva420p", B200J2K_PIX_PLANAR8, 4, 8, 1, 1, 0 };
    if (h.precision > 16)
        return B200J2K_PIX_NONE;
    const bool dcinema = h.rsiz == 3 || h.rsiz == 4;      /* AV_PROFILE_JPEG2000_DCINEMA_2K / _4K: the XYZ list, whatever
                                                           * the colour specification says (jpeg2000dec.c:330-333) */
    switch (dcinema ? -1 : h.colour_space) {
    case -1:
        /* a stream of at most 8 bits would be written with one byte per sample into the 16-bit XYZ12 frame
         * (write_frame_8), leaving half of every row untouched: not served */
        if (h.precision <= 8 && h.ncomp == 3)
            return B200J2K_PIX_NONE;
        d = first_match(XYZ_FMTS, h);
        if (!d) d = first_match(YUV_FMTS, h);
        break;
    case 16: d = first_match(RGB_FMTS, h); break;
    case 17: d = first_match(GRAY_FMTS, h); break;
    case 18: d = first_match(YUV_FMTS, h); break;
    default:
        d = first_match(RGB_FMTS, h);
        if (!d) d = first_match(GRAY_FMTS, h);
        if (!d) d = first_match(YUV_FMTS, h);
        if (!d) d = first_match(XYZ_FMTS, h);
        break;
    }
    if (!d) {
        /* the reference's last resort (jpeg2000dec.c:364-398) */
        const bool same01 = h.ncomp >= 2 && h.cdx[0] == h.cdx[1] && h.cdy[0] == h.cdy[1];
        if (h.ncomp == 4 && h.cdx[0] == 1 && h.cdy[0] == 1 && h.cdx[1] == 1 && h.cdy[1] == 1 &&
            h.cdx[2] == h.cdx[3] && h.cdy[2] == h.cdy[3]) {
            if (h.precision == 8 && h.cdx[2] == 2 && h.cdy[2] == 2 && !h.pal8) {
                d = &YUVA420P_FALLBACK;
                for (int i = 0; i < 4; i++)
                    h.cdef[i] = i;
            }
        } else if (h.ncomp == 3 && h.precision == 8 && same01 && h.cdx[0] == h.cdx[2] && h.cdy[0] == h.cdy[2]) {
            d = &RGB_FMTS[1];
        } else if (h.ncomp == 2 && h.precision == 8 && same01) {
            d = &GRAY_FMTS[1];
        } else if (h.ncomp == 2 && h.precision == 16 && same01) {
            d = &GRAY_FMTS[3];
        } else if (h.ncomp == 1 && h.precision == 8) {
            d = &GRAY_FMTS[0];
        } else if (h.ncomp == 1 && h.precision == 12) {
            d = &GRAY_FMTS[2];
        }
    }
    if (!d || d->klass == B200J2K_PIX_NONE)
        return B200J2K_PIX_NONE;
    *name = d->name;
    h.chroma_lw = d->lw;
    h.chroma_lh = d->lh;
    return d->klass;
}

/* ---- JP2 wrapper: jp2_find_codestream (jpeg2000dec.c:2658-2804) --------------------------------------
 * Walks the boxes up to the contiguous codestream box; from the JP2 header box it takes the enumerated
 * colour space ('colr' method 1), a valid palette ('pclr': steers the format choice and, for a one-component
 * stream of at most 8 bits outside the greyscale list, makes the output PAL8) and the channel definitions
 * ('cdef': component -> plane for planar output) and the sample aspect ratio ('res ').
 * Returns 1 with *pos at the codestream, 0 when there is none, < 0 on a malformed box. */
struct Cursor {
    const uint8_t *b;
    int64_t pos, size;
    int64_t left() const { return size - pos; }
    uint32_t be32()                               /* the reference reads unchecked into the zero padding */
    {
        uint32_t v = 0;
        for (int i = 0; i < 4; i++, pos++)
            v = v << 8 | (pos < size ? b[pos] : 0);
        return v;
    }
    uint32_t be16() { uint32_t v = 0; for (int i = 0; i < 2; i++, pos++) v = v << 8 | (pos < size ? b[pos] : 0); return v; }
    uint32_t u8() { uint32_t v = pos < size ? b[pos] : 0; pos++; return v; }
    void seek(int64_t p) { pos = p < 0 ? 0 : p > size ? size : p; }
};

/* av_reduce() of libavutil (rational.c:35-78) for non-negative arguments: lowest terms when both fit under max,
 * else the closest fraction with both terms <= max (last fitting convergent of the continued fraction, or the
 * last semi-convergent when that is closer). */
void reduce_fraction(int *out_num, int *out_den, int64_t num, int64_t den, int64_t max)
{
    int64_t g = num, t = den;
    while (t) { const int64_t r = g % t; g = t; t = r; }
    if (g) { num /= g; den /= g; }
    int64_t pn = 0, pd = 1, cn = 1, cd = 0;                 /* previous / current convergent */
    if (num <= max && den <= max) {
        cn = num; cd = den; den = 0;
    }
    while (den) {
        const uint64_t x = (uint64_t)(num / den);
        const int64_t next_den = num - den * (int64_t)x;
        const int64_t nn = (int64_t)x * cn + pn, nd = (int64_t)x * cd + pd;
        if (nn > max || nd > max) {
            uint64_t k = x;
            if (cn) k = (uint64_t)((max - pn) / cn);
            if (cd && (uint64_t)((max - pd) / cd) < k) k = (uint64_t)((max - pd) / cd);
            if (den * (2 * (int64_t)k * cd + pd) > num * cd) {
                const int64_t a = (int64_t)k * cn + pn, b = (int64_t)k * cd + pd;
                cn = a; cd = b;
            }
            break;
        }
        pn = cn; pd = cd; cn = nn; cd = nd;
        num = den; den = next_den;
    }
    *out_num = (int)cn;
    *out_den = (int)cd;
}

int jp2_find_codestream(Cursor &g, Header &h)
{
    int search_range = 10;
    while (search_range && g.left() >= 8) {
        uint32_t atom_size = g.be32();
        const uint32_t atom = g.be32();
        uint32_t atom_end;
        if (atom_size == 1) {
            if (g.be32())
                return 0;                                  /* "Huge atom" */
            atom_size = g.be32();
            if (atom_size < 16 || g.pos + atom_size - 16 > INT32_MAX)
                return B200J2K_ERR_INVALIDDATA;
            atom_end = (uint32_t)(g.pos + atom_size - 16);
        } else {
            if (atom_size < 8 || g.pos + atom_size - 8 > INT32_MAX)
                return B200J2K_ERR_INVALIDDATA;
            atom_end = (uint32_t)(g.pos + atom_size - 8);
        }
        if (atom == 0x6A703263u)                           /* 'jp2c' */
            return 1;
        if (g.left() < (int64_t)atom_size || atom_end < atom_size)
            return 0;
        if (atom == 0x6A703268u && atom_size >= 16) {      /* 'jp2h' */
            uint32_t atom2_end;
            do {
                if (g.left() < 8)
                    break;
                uint32_t atom2_size = g.be32();
                const uint32_t atom2 = g.be32();
                atom2_end = (uint32_t)(g.pos + atom2_size - 8);
                if (atom2_size < 8 || atom2_end > atom_end || atom2_end < atom2_size)
                    break;
                atom2_size -= 8;
                if (atom2 == 0x6A703263u) {
                    return 1;
                } else if (atom2 == 0x636F6C72u && atom2_size >= 7) {          /* 'colr' */
                    const int method = (int)g.u8();
                    g.pos += 2;
                    if (method == 1)
                        h.colour_space = (int)g.be32();
                } else if (atom2 == 0x70636C72u && atom2_size >= 6) {          /* 'pclr' */
                    const int count = (int)g.be16(), channels = (int)g.u8();
                    int depth[3], bytes = 0;
                    for (int i = 0; i < 3; i++) {
                        depth[i] = (int)(g.u8() & 0x7F) + 1;
                        bytes += ((depth[i] + 7) >> 3) * count;
                    }
                    if (!(count > 256 || channels != 3 || depth[0] > 16 || depth[1] > 16 || depth[2] > 16 ||
                          (int64_t)atom2_size < bytes)) {
                        h.pal8 = 1;
                        for (int i = 0; i < count; i++) {                      /* :2731-2752, uint32 arithmetic */
                            uint32_t v[3];
                            for (int k = 0; k < 3; k++) {
                                if (depth[k] <= 8) {
                                    v[k] = g.u8() << (8 - depth[k]);
                                    v[k] |= v[k] >> depth[k];
                                } else {
                                    v[k] = g.be16() >> (depth[k] - 8);
                                }
                            }
                            h.palette[i] = 0xffu << 24 | v[0] << 16 | v[1] << 8 | v[2];
                        }
                    }
                } else if (atom2 == 0x63646566u && atom2_size >= 2) {          /* 'cdef' */
                    for (int n = (int)g.be16(); n > 0; n--) {
                        const int cn = (int)g.be16();
                        g.be16();
                        const int asoc = (int)g.be16();
                        if (cn < 4 && asoc < 4)
                            h.cdef[cn] = asoc;
                        if (g.pos >= g.size + 6)           /* the checked reads return 0 forever: stop early */
                            break;
                    }
                } else if (atom2 == 0x72657320u && atom2_size >= 18) {         /* 'res ': capture / display resolution */
                    g.pos += 4;
                    const uint32_t resx = g.be32();
                    if (resx == 0x72657363u || resx == 0x72657364u) {          /* 'resc' / 'resd' */
                        int64_t vnum = g.be16(), vden = g.be16(), hnum = g.be16(), hden = g.be16();
                        int64_t vexp = g.u8(), hexp = g.u8();
                        if (vnum && vden && hnum && hden) {
                            if (vexp > hexp) { vexp -= hexp; hexp = 0; } else { hexp -= vexp; vexp = 0; }
                            /* same double arithmetic as the reference (:2788-2794) */
                            if (INT64_MAX / (hnum * vden) > pow(10, hexp) && INT64_MAX / (vnum * hden) > pow(10, vexp))
                                reduce_fraction(&h.sar_den, &h.sar_num, (int64_t)(hnum * vden * pow(10, hexp)),
                                                (int64_t)(vnum * hden * pow(10, vexp)), INT32_MAX);
                        }
                    }
                }
                g.seek(atom2_end);
            } while (atom_end - atom2_end >= 8);
        } else {
            search_range--;
        }
        g.seek(atom_end);
    }
    return 0;
}

/* band step size, float path, same operation order as jpeg2000.c:214-272 */
float band_stepsize(const CodSty &c, const QntSty &q, int bandno, int gbandno, int r, int cbps, int red)
{
    float f;
    if (q.style == 0) {
        f = 1;
    } else if (q.style <= 2) {
        union { uint32_t u; float f; } e;
        e.u = (uint32_t)((cbps - q.expn[gbandno]) + 127) << 23;
        f = e.f;
        f *= q.mant[gbandno] / 2048.0 + 1.0;
    } else {
        f = 0;                                            /* "Unknown quantization format" (jpeg2000.c:246-249) */
    }
    if (c.transform != J2K_DWT53) {
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
        f *= pow(1.230174104914001f, 2 * (c.nres - red - r) + lband - 2);   /* nreslevels2decode (jpeg2000.c:263) */
    }
    if (f > (INT32_MAX >> 15))
        f = 0;
    return f;
}

uint32_t tagtree_nodes(int w, int h)
{
    uint32_t n = 1;
    while (w > 1 || h > 1) {
        n += (uint32_t)w * h;
        w = (w + 1) >> 1;
        h = (h + 1) >> 1;
    }
    return n;
}

}  // namespace

const char *pix_fmt_name(int pix)
{
    switch (pix) {
    case B200J2K_PIX_GRAY8: return "gray";
    case B200J2K_PIX_GRAY16: return "gray16le";
    case B200J2K_PIX_YA8: return "ya8";
    case B200J2K_PIX_YA16: return "ya16le";
    case B200J2K_PIX_RGB24: return "rgb24";
    case B200J2K_PIX_RGB48: return "rgb48le";
    case B200J2K_PIX_RGBA: return "rgba";
    case B200J2K_PIX_RGBA64: return "rgba64le";
    case B200J2K_PIX_PLANAR8: return "planar8";
    case B200J2K_PIX_PLANAR16: return "planar16";
    }
    return "none";
}

/* ---- marker segments -------------------------------------------------------------------------------- */
int Parser::parse_headers(const uint8_t *pkt, int64_t size, int flags, size_t *main_header_len)
{
    const uint8_t *p = pkt, *end = pkt + size;
    Header &h = hdr;
    h = Header();
    parts_.clear();
    ppm_.clear();
    has_ppm_ = false;
    ppt_.clear();
    tsty_.clear();
    geom_extra_.clear();
    tile_first_.clear();
    has_ppt_.clear();
    ppt_pos_.clear();
    size_t ppm_pos = 0;                                    /* s->packed_headers_stream */
    /* tile-part header segments that change what build_geometry lays out: part of the geometry cache key */
    auto note_tile_segment = [&](int tile, const uint8_t *from, const uint8_t *to) {
        const uint8_t t[4] = { (uint8_t)(tile >> 24), (uint8_t)(tile >> 16), (uint8_t)(tile >> 8), (uint8_t)tile };
        geom_extra_.insert(geom_extra_.end(), t, t + 4);
        geom_extra_.insert(geom_extra_.end(), from, to);
    };
    *main_header_len = 0;
    if (size >= 12 && rd32(p) == 12 && rd32(p + 4) == 0x6A502020u && rd32(p + 8) == 0x0D0A870Au) {
        /* JP2 / JPH file (jpeg2000dec.c:2842-2854) */
        Cursor g = { pkt, 12, size };
        /* "if (!jp2_find_codestream(s))" (:2846): the error code an absurd box size produces is not zero, so the
         * reference carries on and looks for SOC behind the box header it has read (the cursor stands there) */
        const int r = jp2_find_codestream(g, h);
        if (!r)
            return B200J2K_ERR_INVALIDDATA;               /* "Could not find Jpeg2000 codestream atom." */
        p = pkt + g.pos;
    }
    while (end - p >= 3 && rd16(p) != 0xFF4F)
        p++;
    if (end - p < 2 || rd16(p) != 0xFF4F)
        return B200J2K_ERR_INVALIDDATA;
    p += 2;
    int cur_tile = -1, cur_tp = 0;
    const uint8_t *tp_end = nullptr;
    for (;;) {
        if (end - p < 2)
            break;
        unsigned marker = rd16(p);
        const uint8_t *marker_pos = p;
        p += 2;
        if (marker >= 0xFF30 && marker <= 0xFF3F)
            continue;
        if (marker == 0xFF93) {                           /* SOD */
            if (parts_.empty() || cur_tile < 0 || tp_end < p)
                return B200J2K_ERR_INVALIDDATA;
            auto &pv = parts_[cur_tile];
            if ((int)pv.size() <= cur_tp)
                pv.resize(cur_tp + 1, TilePart{ nullptr, nullptr, nullptr, nullptr, nullptr });
            const uint8_t *hp = nullptr, *hend = nullptr;
            if (has_ppm_) {                               /* this tile-part's run: Nppm, then the headers (:2467-2472) */
                uint32_t n = 0;
                if (ppm_.size() - ppm_pos >= 4) {
                    n = rd32(ppm_.data() + ppm_pos);
                    ppm_pos += 4;
                }
                if (ppm_.size() - ppm_pos < n)
                    return B200J2K_ERR_INVALIDDATA;
                hp = ppm_.data() + ppm_pos;
                hend = hp + n;
                ppm_pos += n;
            }
            /* the tile's packed headers become readable HERE, as much of them as has arrived (:2474-2476); a tile
             * whose first tile-part never reaches its SOD reads an empty header stream: all packets empty */
            if (has_ppt_[cur_tile] && cur_tp == 0) {
                ppt_pos_[cur_tile] = 0;
                ppt_live_[cur_tile] = ppt_[cur_tile].size();
            }
            pv[cur_tp] = TilePart{ p, tp_end, p, hp, hend };
            p = tp_end;
            continue;
        }
        if (marker == 0xFFD9)
            break;
        /* "Missing EOC Marker" (:2485-2494): the length field has been read (a checked read: short of two bytes it
         * moves to the end) before the loop gives up -- the position counts for the bytes consumed */
        if (end - p < 2) {
            p = end;
            break;
        }
        int len = (int)rd16(p);
        if (len < 2 || end - p < len) {
            p += 2;
            break;
        }
        const uint8_t *seg = p + 2, *segend = p + len;
        int ret = 0;
        switch (marker) {
        case 0xFF51: {
            if (h.ncomp || segend - seg < 36)
                return B200J2K_ERR_INVALIDDATA;
            h.rsiz = (int)rd16(seg);
            h.W = (int)rd32(seg + 2); h.H = (int)rd32(seg + 6); h.X0 = (int)rd32(seg + 10); h.Y0 = (int)rd32(seg + 14);
            h.TW = (int)rd32(seg + 18); h.TH = (int)rd32(seg + 22); h.TX0 = (int)rd32(seg + 26); h.TY0 = (int)rd32(seg + 30);
            int nc = (int)rd16(seg + 34);
            /* av_image_check_size2() without a pixel format (imgutils.c:289-316): rows of 8 bytes per pixel plus
             * padding, 128 rows of padding, all below INT_MAX; then max_pixels = INT_MAX */
            if (h.W <= 0 || h.H <= 0 || 8LL * h.W + 1024 >= INT32_MAX ||
                (uint64_t)(8LL * h.W + 1024) * ((uint64_t)h.H + 128) >= (uint64_t)INT32_MAX || (int64_t)h.W * h.H > INT32_MAX)
                return B200J2K_ERR_PATCHWELCOME;
            if (nc <= 0)
                return B200J2K_ERR_INVALIDDATA;
            if (nc > 4)
                return B200J2K_ERR_PATCHWELCOME;
            if (h.TX0 < 0 || h.TY0 < 0 || h.X0 < h.TX0 || h.Y0 < h.TY0 ||
                h.TW + (int64_t)h.TX0 <= h.X0 || h.TH + (int64_t)h.TY0 <= h.Y0)
                return B200J2K_ERR_INVALIDDATA;
            if (h.X0 >= h.W || h.Y0 >= h.H || h.TW <= 0 || h.TH <= 0)
                return B200J2K_ERR_INVALIDDATA;
            if (segend - seg < 36 + 3 * nc)
                return B200J2K_ERR_INVALIDDATA;
            h.ncomp = nc;
            for (int i = 0; i < nc; i++) {
                unsigned x = seg[36 + 3 * i];
                h.cbps[i] = (int)(x & 0x7F) + 1;
                h.precision = imax(h.precision, h.cbps[i]);
                h.cdx[i] = seg[37 + 3 * i];
                h.cdy[i] = seg[38 + 3 * i];
                if (!h.cdx[i] || h.cdx[i] == 3 || h.cdx[i] > 4 || !h.cdy[i] || h.cdy[i] == 3 || h.cdy[i] > 4)
                    return B200J2K_ERR_INVALIDDATA;
            }
            h.ntx = ceildiv(h.W - h.TX0, h.TW);
            h.nty = ceildiv(h.H - h.TY0, h.TH);
            if ((int64_t)h.ntx * h.nty * 14 > size)
                return B200J2K_ERR_EINVAL;
            parts_.assign((size_t)h.ntx * h.nty, {});
            tile_first_.assign((size_t)h.ntx * h.nty, 0);
            tsty_.assign((size_t)h.ntx * h.nty, TileSty());
            ppt_.assign((size_t)h.ntx * h.nty, {});
            has_ppt_.assign((size_t)h.ntx * h.nty, 0);
            ppt_pos_.assign((size_t)h.ntx * h.nty, 0);
            ppt_live_.assign((size_t)h.ntx * h.nty, 0);
            const int red = B200J2K_FLAG_GET_LOWRES(flags);
            if (red && (h.X0 || h.Y0))
                return B200J2K_ERR_PATCHWELCOME;          /* jpeg2000dec.c:256-259 */
            /* image size with the reduction factor (jpeg2000dec.c:314-320) */
            int ow = ceildivpow2(h.W - h.X0, red), oh = ceildivpow2(h.H - h.Y0, red);
            h.out_w = ceildiv(ow, h.cdx[0]);
            h.out_h = ceildiv(oh, h.cdy[0]);
            for (int i = 1; i < nc; i++) {
                h.out_w = imax(h.out_w, ceildiv(ow, h.cdx[i]));
                h.out_h = imax(h.out_h, ceildiv(oh, h.cdy[i]));
            }
            h.pix = choose_pix(h, &h.pix_name);
            if (!h.pix)
                return B200J2K_ERR_PATCHWELCOME;
            seg += 36 + 3 * nc;
            break;
        }
        case 0xFF50: {
            if (segend - seg < 6)
                return B200J2K_ERR_INVALIDDATA;
            uint32_t pcap = rd32(seg);
            seg += 4;
            unsigned ccap15 = 0;
            h.isHT = (pcap >> 17) & 1;
            for (int i = 0; i < 32; i++)
                if ((pcap >> (31 - i)) & 1) {
                    if (segend - seg < 2)
                        return B200J2K_ERR_INVALIDDATA;
                    if (i == 14)
                        ccap15 = rd16(seg);
                    seg += 2;
                }
            if (h.isHT) {
                h.cap_b14 = (ccap15 >> 14) & 3;
                if (h.cap_b14 == 2)
                    return B200J2K_ERR_EINVAL;
                if ((ccap15 >> 13) & 1)
                    return B200J2K_ERR_PATCHWELCOME;
                h.cap_b05 = (ccap15 >> 5) & 1;
                h.cap_b11 = (ccap15 >> 11) & 1;
                h.cap_b12 = (ccap15 >> 12) & 1;
                int P = ccap15 & 31, B = !P ? 8 : P < 20 ? P + 8 : P < 31 ? 4 * (P - 19) + 27 : 74;
                if (B > 31)
                    return B200J2K_ERR_PATCHWELCOME;
            }
            break;
        }
        case 0xFF52: {
            /* in a tile-part header: the styles of that tile; INVALIDDATA for the HOMOGENEOUS HT set
             * (jpeg2000dec.c:2514-2556) */
            if (cur_tile >= 0 && h.isHT && !h.cap_b11)
                return B200J2K_ERR_INVALIDDATA;
            if (segend - seg < 5)
                return B200J2K_ERR_INVALIDDATA;
            CodSty *codp = cur_tile >= 0 ? tsty_[cur_tile].cod : h.cod;
            CodSty tmp;
            tmp.csty = seg[0];
            tmp.prog = seg[1];
            tmp.nlayers = (int)(rd16(seg + 2) & 0xFF);          /* the reference keeps it in a uint8_t (jpeg2000.h:152) */
            tmp.mct = seg[4];
            if (tmp.mct && h.ncomp < 3)
                return B200J2K_ERR_INVALIDDATA;
            if ((ret = parse_cox(seg + 5, segend, &tmp, flags)) < 0)
                return ret;
            tmp.init = 1;
            for (int c = 0; c < h.ncomp; c++)
                if (!(cur_tile >= 0 ? tsty_[cur_tile].had_coc[c] : h.had_coc[c]))
                    codp[c] = tmp;
            if (cur_tile >= 0)
                note_tile_segment(cur_tile, marker_pos, segend);
            seg += 10 + ((tmp.csty & 1) ? tmp.nres : 0);
            break;
        }
        case 0xFF53: {
            if (cur_tile >= 0 && h.isHT && !h.cap_b11)
                return B200J2K_ERR_INVALIDDATA;
            if (segend - seg < 2)
                return B200J2K_ERR_INVALIDDATA;
            int c = seg[0];
            if (c >= h.ncomp)
                return B200J2K_ERR_INVALIDDATA;
            CodSty *cs = cur_tile >= 0 ? &tsty_[cur_tile].cod[c] : &h.cod[c];
            int keep = cs->csty & 6;
            cs->csty = seg[1] | keep;
            if ((ret = parse_cox(seg + 2, segend, cs, flags)) < 0)
                return ret;
            if (cur_tile >= 0) {
                tsty_[cur_tile].had_coc[c] = 1;
                note_tile_segment(cur_tile, marker_pos, segend);
            } else
                h.had_coc[c] = 1;
            cs->init = 1;
            seg += 7 + ((cs->csty & 1) ? cs->nres : 0);
            break;
        }
        case 0xFF5C: {
            if (cur_tile >= 0 && h.isHT && !h.cap_b11)
                return B200J2K_ERR_INVALIDDATA;
            QntSty tmp;
            memset(&tmp, 0, sizeof(tmp));
            if ((ret = parse_qcx(seg, len - 2, end - seg, &tmp)) < 0)
                return ret;
            seg += ret;
            for (int c = 0; c < h.ncomp; c++) {
                if (cur_tile >= 0) {
                    if (!tsty_[cur_tile].had_qcc[c])
                        tsty_[cur_tile].qnt[c] = tmp;
                } else if (!h.had_qcc[c])
                    h.qnt[c] = tmp;
            }
            if (cur_tile >= 0)
                note_tile_segment(cur_tile, marker_pos, segend);
            break;
        }
        case 0xFF5D: {
            if (cur_tile >= 0 && h.isHT && !h.cap_b11)
                return B200J2K_ERR_INVALIDDATA;
            if (segend - seg < 1)
                return B200J2K_ERR_INVALIDDATA;
            int c = seg[0];
            if (c >= h.ncomp)
                return B200J2K_ERR_INVALIDDATA;
            if (cur_tile >= 0) {
                tsty_[cur_tile].had_qcc[c] = 1;
                note_tile_segment(cur_tile, marker_pos, segend);
            } else
                h.had_qcc[c] = 1;
            if ((ret = parse_qcx(seg + 1, len - 3, end - (seg + 1), cur_tile >= 0 ? &tsty_[cur_tile].qnt[c] : &h.qnt[c])) < 0)
                return ret;
            seg += 1 + ret;
            break;
        }
        case 0xFF5E: {                                    /* RGN (get_rgn, jpeg2000dec.c:641-672; :2530-2534) */
            if (cur_tile >= 0 && h.isHT && !h.cap_b11)
                return B200J2K_ERR_INVALIDDATA;
            if (h.isHT && !h.cap_b12)                     /* "belongs to the RGNFREE set" (:2536-2540) */
                return B200J2K_ERR_INVALIDDATA;
            if (segend - seg < 3 || seg[1] != 0 || seg[0] >= h.ncomp)
                return B200J2K_ERR_INVALIDDATA;
            if (cur_tile >= 0 && cur_tp != 0)             /* only in the first tile-part of a tile */
                return B200J2K_ERR_INVALIDDATA;
            if (seg[2] > 30)
                return B200J2K_ERR_PATCHWELCOME;
            if (cur_tile >= 0) {
                tsty_[cur_tile].roi[seg[0]] = seg[2];
                note_tile_segment(cur_tile, marker_pos, segend);
            } else
                h.roi[seg[0]] = seg[2];
            seg += 3;
            break;
        }
        case 0xFF90: {
            if (parts_.empty() || segend - seg < 8)
                return B200J2K_ERR_INVALIDDATA;
            if (cur_tile < 0)
                *main_header_len = (size_t)(marker_pos - pkt);
            int isot = (int)rd16(seg);
            if (isot >= h.ntx * h.nty)
                return B200J2K_ERR_INVALIDDATA;
            uint32_t psot = rd32(seg + 2);
            int tpsot = seg[6];
            int64_t left = end - (seg + 8);
            /* get_sot measures against "bytes left - 2 + n + 2" AFTER it has read the 8 bytes (:845-850): a
             * tile-part has to end two bytes (the EOC marker) before the end of the packet */
            if (!psot)
                psot = (uint32_t)(left + len);
            if ((int64_t)psot > left + len)
                return B200J2K_ERR_INVALIDDATA;
            if (tpsot >= 32)
                return B200J2K_ERR_PATCHWELCOME;
            cur_tile = isot;
            cur_tp = tpsot;
            if (!tpsot) {                                 /* get_sot copies the styles to the tile (:862-869) */
                tile_first_[isot] = 1;
                TileSty &T = tsty_[isot];
                for (int c = 0; c < J2K_MAX_COMP; c++) {
                    T.cod[c] = h.cod[c];
                    T.qnt[c] = h.qnt[c];
                }
                T.npoc = h.npoc;
                for (int i = 0; i < h.npoc; i++)
                    T.poc[i] = h.poc[i];
                T.poc_default = true;
            }
            tp_end = seg + 8 + psot - len - 2;
            seg += 8;
            break;
        }
        case 0xFF5F: {                                    /* POC (get_poc, jpeg2000dec.c:760-818) */
            if (cur_tile >= 0 && h.isHT && !h.cap_b11)    /* HOMOGENEOUS HT set: no progression change per tile (:2558) */
                return B200J2K_ERR_INVALIDDATA;
            if (end - seg < 5 || len < 2 + 7)
                return B200J2K_ERR_INVALIDDATA;
            const int n = (len - 2) / 7;
            if (n > 32)
                return B200J2K_ERR_PATCHWELCOME;
            Header::Poc tmp[32];
            for (int i = 0; i < n; i++) {
                Header::Poc &e = tmp[i];
                const uint8_t *q = seg + 7 * i;
                e.RS = q[0]; e.CS = q[1]; e.LYE = (int)rd16(q + 2); e.RE = q[4]; e.CE = q[5]; e.P = q[6];
                if (!e.CE)
                    e.CE = 256;
                if (e.CE > h.ncomp)
                    e.CE = h.ncomp;
                if (e.RS >= e.RE || e.RE > 33 || e.CS >= e.CE || e.CE > h.ncomp || !e.LYE)
                    return B200J2K_ERR_INVALIDDATA;
            }
            if (cur_tile >= 0) {
                /* the tile's list starts as a copy of the main header's ("default"): the first POC segment of
                 * the tile replaces it, further ones extend it (:806-816) */
                TileSty &T = tsty_[cur_tile];
                if (!T.npoc || T.poc_default)
                    T.npoc = 0;
                else if (T.npoc + n > 32)
                    return B200J2K_ERR_INVALIDDATA;
                for (int i = 0; i < n; i++)
                    T.poc[T.npoc++] = tmp[i];
                T.poc_default = false;
            } else {
                if (h.npoc + n > 32)
                    return B200J2K_ERR_INVALIDDATA;
                for (int i = 0; i < n; i++)               /* a second POC segment extends the list */
                    h.poc[h.npoc++] = tmp[i];
            }
            seg += 7 * n;
            break;
        }
        case 0xFF60:                                      /* PPM (get_ppm, jpeg2000dec.c:958-980, :2597-2604) */
            if (cur_tile >= 0 || len < 3)
                return B200J2K_ERR_INVALIDDATA;
            has_ppm_ = true;
            ppm_.insert(ppm_.end(), seg + 1, segend);     /* Zppm is not used */
            seg = segend;
            break;
        case 0xFF61:                                      /* PPT (get_ppt, :982-1014, :2605-2617) */
            if (has_ppm_ || (h.isHT && !h.cap_b11) || len < 3 || cur_tile < 0 || cur_tp != 0)
                return B200J2K_ERR_INVALIDDATA;
            has_ppt_[cur_tile] = 1;
            ppt_[cur_tile].insert(ppt_[cur_tile].end(), seg + 1, segend);
            seg = segend;
            break;
        case 0xFF55: {                                    /* TLM: only walked over (get_tlm, jpeg2000dec.c:901-934) */
            const unsigned stlm = end - seg > 1 ? seg[1] : 0;
            const int st = (stlm >> 4) & 3, sp = (stlm >> 6) & 1;
            if (st == 3)
                return B200J2K_ERR_INVALIDDATA;
            const int entry = (sp + 1) * 2 + st;
            seg += 2 + (int)(uint8_t)((len - 4) / entry) * entry;   /* the entry count lives in a uint8_t */
            break;
        }
        case 0xFF58:                                      /* PLT (get_plt, :936-955): its last byte ends a length */
            if (len < 4 || (segend[-1] & 0x80))
                return B200J2K_ERR_INVALIDDATA;
            seg = segend;
            break;
        case 0xFF63:                                      /* CRG (read_crg, :875-883) */
            if (h.ncomp * 4 != len - 2)
                return B200J2K_ERR_INVALIDDATA;
            seg = segend;
            break;
        default:                                          /* COM, PLM, CPF, unknown markers: skipped */
            seg = segend;
            break;
        }
        if (seg != segend)
            return B200J2K_ERR_UNSPECIFIED;              /* declared length != what was consumed (:2629-2633) */
        p = segend;
    }
    consumed = (int64_t)(p - pkt);
    if (!h.ncomp || parts_.empty())
        return B200J2K_ERR_INVALIDDATA;
    if (!*main_header_len)
        *main_header_len = (size_t)(p - pkt);
    for (int c = 0; c < h.ncomp; c++) {
        if (!h.cod[c].init)
            return B200J2K_ERR_INVALIDDATA;
        if (h.cod[c].transform == J2K_DWT53)
            h.lossless = 1;
    }
    /* channel definitions default to the identity, the last of an even number of components being alpha
     * (jpeg2000dec.c:2883-2892) */
    for (int x = 0; x < h.ncomp; x++)
        if (h.cdef[x] < 0) {
            for (x = 0; x < h.ncomp; x++)
                h.cdef[x] = x + 1;
            if ((h.ncomp & 1) == 0)
                h.cdef[h.ncomp - 1] = 0;
            break;
        }
    /* output planes */
    const int bps = h.precision > 8 ? 2 : 1;
    const bool planar = h.pix == B200J2K_PIX_PLANAR8 || h.pix == B200J2K_PIX_PLANAR16;
    picture_bytes = 0;
    if (!planar) {
        nplanes = 1;
        plane_w[0] = h.out_w;
        plane_h[0] = h.out_h;
        plane_row_bytes[0] = h.out_w * h.ncomp * bps;
        plane_off[0] = 0;
        picture_bytes = (uint64_t)plane_row_bytes[0] * h.out_h;
        if (h.pix_name && !strcmp(h.pix_name, "pal8")) {
            /* AVFrame.data[1] of a PAL8 frame: 256 native-endian 0xAARRGGBB words (jpeg2000dec.c:2900-2901) */
            nplanes = 2;
            plane_w[1] = 256;
            plane_h[1] = 1;
            plane_row_bytes[1] = 1024;
            plane_off[1] = picture_bytes;
            picture_bytes = plane_off[1] + 1024;
        }
    } else {
        /* planes of the chosen AVPixelFormat: chroma planes 1 and 2 shrink by its log2_chroma_w/h */
        nplanes = h.ncomp;
        for (int c = 0; c < h.ncomp; c++) {
            const bool chroma = c == 1 || c == 2;
            plane_w[c] = chroma ? (h.out_w + (1 << h.chroma_lw) - 1) >> h.chroma_lw : h.out_w;
            plane_h[c] = chroma ? (h.out_h + (1 << h.chroma_lh) - 1) >> h.chroma_lh : h.out_h;
            plane_row_bytes[c] = plane_w[c] * bps;
            plane_off[c] = picture_bytes;
            picture_bytes += (uint64_t)plane_row_bytes[c] * plane_h[c];
        }
    }
    return 0;
}

/* ---- geometry (cached) ------------------------------------------------------------------------------ */
int Parser::build_geometry(int flags)
{
    const Header &h = hdr;
    tiles_.clear(); res_.clear(); precs_.clear(); cblk_template_.clear();
    bands.clear(); band_negative_Mb_.clear(); tcs.clear(); groups.clear();
    tag_total_ = 0;
    coef_elems = 0;
    ll_elems = 0;
    max_levels = 0;
    max_w = max_h = 0;
    max_area = max_state = 0;
    const int ntiles = h.ntx * h.nty;
    const int red = B200J2K_FLAG_GET_LOWRES(flags);
    tiles_.resize(ntiles);
    const bool planar = h.pix == B200J2K_PIX_PLANAR8 || h.pix == B200J2K_PIX_PLANAR16;
    const bool wide = h.precision > 8;
    const bool msb16 = h.pix == B200J2K_PIX_RGB48 || h.pix == B200J2K_PIX_RGBA64 || h.pix == B200J2K_PIX_GRAY16;
    const int out_precision = !wide ? 8 : msb16 ? 16 : h.precision;

    for (int tileno = 0; tileno < ntiles; tileno++) {
        TileGeom &t = tiles_[tileno];
        const int tx = tileno % h.ntx, ty = tileno / h.ntx;
        t.c.x0 = clip64((int64_t)tx * h.TW + h.TX0, h.X0, h.W);
        t.c.x1 = clip64((int64_t)(tx + 1) * h.TW + h.TX0, h.X0, h.W);
        t.c.y0 = clip64((int64_t)ty * h.TH + h.TY0, h.Y0, h.H);
        t.c.y1 = clip64((int64_t)(ty + 1) * h.TH + h.TY0, h.Y0, h.H);
        const int first_tc = (int)tcs.size();
        for (int c = 0; c < h.ncomp; c++) {
            const CodSty &cs = tsty_[tileno].cod[c];
            const QntSty &qs = tsty_[tileno].qnt[c];
            if (!cs.init)
                return B200J2K_ERR_INVALIDDATA;
            if (h.isHT && !h.cap_b05 && !cs.transform)
                return B200J2K_ERR_INVALIDDATA;
            if (h.isHT && h.cap_b14 != (cs.cblk_style >> 6) && h.cap_b14 != 0)
                return B200J2K_ERR_INVALIDDATA;
            /* Part-1 blocks with arithmetic-coding bypass or termination on each pass (several codeword segments per
             * block) are not served: refused in tier 2, when the first such block contributes to a packet -- a stream
             * that only announces the style decodes like in the reference */
            /* cc = comp->coord_o: every resolution, band, precinct and code-block is laid out from it;
             * cr = comp->coord = cc reduced by lowres: the plane, the DWT and the output use it
             * (jpeg2000dec.c:1047-1050, jpeg2000.c:483-530) */
            Rect cc, cr;
            cc.x0 = ceildiv(t.c.x0, h.cdx[c]); cc.x1 = ceildiv(t.c.x1, h.cdx[c]);
            cc.y0 = ceildiv(t.c.y0, h.cdy[c]); cc.y1 = ceildiv(t.c.y1, h.cdy[c]);
            t.comp[c] = cc;
            cr.x0 = ceildivpow2(cc.x0, red); cr.x1 = ceildivpow2(cc.x1, red);
            cr.y0 = ceildivpow2(cc.y0, red); cr.y1 = ceildivpow2(cc.y1, red);
            const int nres2decode = cs.nres - red;
            if (nres2decode <= 0)
                return B200J2K_ERR_EINVAL;
            const int cw = cr.x1 - cr.x0, ch = cr.y1 - cr.y0;
            if (cw > 32768 || ch > 32768)
                return B200J2K_ERR_PATCHWELCOME;

            TileCompDesc tc;
            memset(&tc, 0, sizeof(tc));
            tc.coef_off = (uint32_t)coef_elems;
            tc.width = cw; tc.height = ch;
            tc.x0 = cr.x0; tc.y0 = cr.y0;
            tc.nlevels = nres2decode - 1;
            /* a tile outside this context's subset keeps its place in the tables, but no kernel reconstructs it:
             * no transform type matches, and its code-blocks are flagged like those lowres drops */
            tc.transform = tile_selected(flags, tileno) ? cs.transform : -1;
            tc.comp = c;
            tc.out_x = cr.x0 - ceildiv(h.X0, h.cdx[c]);
            tc.out_y = cr.y0 - ceildiv(h.Y0, h.cdy[c]);
            tc.cbps = h.cbps[c];
            {
                int b[4] = { cr.x0, cr.x1, cr.y0, cr.y1 };
                for (int k = tc.nlevels; k >= 0; k--) {
                    tc.lev[k] = LevelDesc{ b[0], b[1], b[2], b[3] };
                    for (int i = 0; i < 4; i++)
                        b[i] = (b[i] + 1) >> 1;
                }
            }
            max_levels = imax(max_levels, tc.nlevels);
            if ((uint64_t)coef_elems + (uint64_t)cw * ch + 64 > 0xFFFFFFFFull)
                return B200J2K_ERR_PATCHWELCOME;
            coef_elems += ((uint64_t)cw * ch + 63) & ~63ull;

            t.res_base[c] = (uint32_t)res_.size();
            res_.resize(res_.size() + cs.nres);
            int gband = 0;
            for (int r = 0; r < cs.nres; r++) {
                ResGeom &rl = res_[t.res_base[c] + r];
                const int decl = cs.nres - r;
                rl.c.x0 = ceildivpow2(cc.x0, decl - 1); rl.c.x1 = ceildivpow2(cc.x1, decl - 1);
                rl.c.y0 = ceildivpow2(cc.y0, decl - 1); rl.c.y1 = ceildivpow2(cc.y1, decl - 1);
                rl.ppx = cs.ppx[r]; rl.ppy = cs.ppy[r];
                rl.nbands = r == 0 ? 1 : 3;
                rl.npx = rl.c.x1 == rl.c.x0 ? 0 : ceildivpow2(rl.c.x1, rl.ppx) - (rl.c.x0 >> rl.ppx);
                rl.npy = rl.c.y1 == rl.c.y0 ? 0 : ceildivpow2(rl.c.y1, rl.ppy) - (rl.c.y0 >> rl.ppy);
                for (int b = 0; b < rl.nbands; b++, gband++) {
                    BandGeom &bd = rl.band[b];
                    int bppx, bppy;
                    if (r == 0) {
                        bd.c = rl.c;
                        bppx = rl.ppx; bppy = rl.ppy;
                    } else {
                        const int xo = (b + 1) & 1, yo = ((b + 1) >> 1) & 1;
                        bd.c.x0 = ceildivpow2((int)(cc.x0 - ((int64_t)xo << (decl - 1))), decl);
                        bd.c.x1 = ceildivpow2((int)(cc.x1 - ((int64_t)xo << (decl - 1))), decl);
                        bd.c.y0 = ceildivpow2((int)(cc.y0 - ((int64_t)yo << (decl - 1))), decl);
                        bd.c.y1 = ceildivpow2((int)(cc.y1 - ((int64_t)yo << (decl - 1))), decl);
                        bppx = rl.ppx - 1; bppy = rl.ppy - 1;
                    }
                    bd.empty = bd.c.x0 == bd.c.x1 || bd.c.y0 == bd.c.y1;
                    const int l2cw = imin(cs.l2cw, bppx), l2ch = imin(cs.l2ch, bppy);
                    const int M_b = qs.expn[gband] + qs.nguard - 1;
                    if (!bd.empty && M_b >= 31 && r < nres2decode)   /* tile_codeblocks only walks the levels it decodes */
                        return B200J2K_ERR_PATCHWELCOME;   /* jpeg2000dec.c:2246-2249 */
                    /* M_b < 0 (no QCD at all: zero exponents, zero guard bits): nothing in the reference minds until
                     * a block of the band decodes -- and then it shifts by 32; refused once tier 2 finds such a block */
                    /* dequantisation constants: D1-D3 */
                    BandDesc bdesc;
                    memset(&bdesc, 0, sizeof(bdesc));
                    {
                        float fstep = band_stepsize(cs, qs, b, gband, r, h.cbps[c], red);
                        bdesc.i_stepsize = (int32_t)floorf(fstep * (1 << 15));
                        const int downshift = 31 - imax(M_b, 0);
                        float fscale = fstep;
                        fscale /= (float)(1 << imax(0, imin(downshift, 30)));   /* an empty band may carry any M_b: its constants are never used */
                        bdesc.fscale = fscale;
                        float s = fscale;
                        s *= (float)(1 << 6);
                        s *= (float)(1 << (16 + 8));
                        bdesc.iscale = (int32_t)(s + 0.5);
                    }
                    bd.band_index = (int)bands.size();
                    bands.push_back(bdesc);
                    band_negative_Mb_.push_back(M_b < 0);
                    bd.prec_base = (uint32_t)precs_.size();
                    const int nprec = rl.npx * rl.npy;
                    /* Mallat placement of the band inside the plane (jpeg2000.c:365-376) */
                    int offx = 0, offy = 0;
                    if (r > 0) {
                        const ResGeom &lo = res_[t.res_base[c] + r - 1];
                        if ((b + 1) & 1) offx = lo.c.x1 - lo.c.x0;
                        if ((b + 1) & 2) offy = lo.c.y1 - lo.c.y0;
                    }
                    for (int pn = 0; pn < nprec; pn++) {
                        PrecGeom pg;
                        memset(&pg, 0, sizeof(pg));
                        pg.cblk_base = (uint32_t)cblk_template_.size();
                        if (!bd.empty) {
                            Rect pc;
                            pc.x0 = ((rl.c.x0 >> rl.ppx) + pn % rl.npx) * (1 << bppx);
                            pc.y0 = ((rl.c.y0 >> rl.ppy) + pn / rl.npx) * (1 << bppy);
                            pc.x1 = pc.x0 + (1 << bppx);
                            pc.y1 = pc.y0 + (1 << bppy);
                            pc.x0 = imax(pc.x0, bd.c.x0); pc.x1 = imin(pc.x1, bd.c.x1);
                            pc.y0 = imax(pc.y0, bd.c.y0); pc.y1 = imin(pc.y1, bd.c.y1);
                            pg.ncw = ceildivpow2(pc.x1, l2cw) - (pc.x0 >> l2cw);
                            pg.nch = ceildivpow2(pc.y1, l2ch) - (pc.y0 >> l2ch);
                            if (pg.ncw < 0 || pg.nch < 0 || pc.x1 <= pc.x0 || pc.y1 <= pc.y0)
                                pg.ncw = pg.nch = 0;
                            for (int k = 0; k < pg.ncw * pg.nch; k++) {
                                int gx = ((pc.x0 >> l2cw) << l2cw) + ((k % pg.ncw) << l2cw);
                                int gy = ((pc.y0 >> l2ch) << l2ch) + ((k / pg.ncw) << l2ch);
                                int x0 = imax(gx, pc.x0), x1 = imin(gx + (1 << l2cw), pc.x1);
                                int y0 = imax(gy, pc.y0), y1 = imin(gy + (1 << l2ch), pc.y1);
                                CblkDesc d;
                                memset(&d, 0, sizeof(d));
                                d.w = (uint16_t)(x1 - x0);
                                d.h = (uint16_t)(y1 - y0);
                                d.dst_stride = (uint32_t)cw;
                                d.dst_off = tc.coef_off + (uint32_t)(y0 - bd.c.y0 + offy) * (uint32_t)cw + (uint32_t)(x0 - bd.c.x0 + offx);
                                d.M_b = (uint8_t)M_b;
                                /* bits 4-5 only mean something to mq_block_kernel: context reset, segmentation symbols */
                                d.flags = (uint8_t)(((cs.cblk_style & 8) ? 1 : 0) | (cs.transform << 1) |
                                                    ((cs.cblk_style & 0x02) ? J2K_CBLK_MQ_RESET : 0) |
                                                    ((cs.cblk_style & 0x20) ? J2K_CBLK_MQ_SEGSYM : 0));
                                d.pad = (uint8_t)(b + (r > 0));              /* bandpos (jpeg2000dec.c:2240) */
                                if (!(cs.cblk_style & 0xC0)) {
                                    d.flags |= J2K_CBLK_MQ;                  /* MIXED sets: decided per block in tier 2 */
                                    d.Lref = d.pad;
                                }
                                if (r >= nres2decode || !tile_selected(flags, tileno)) {
                                    /* dropped by lowres: the packets are still parsed, the block is never decoded
                                     * (tile_codeblocks walks nreslevels2decode levels, jpeg2000dec.c:2230) */
                                    d.flags |= J2K_CBLK_SKIP;
                                    d.dst_off = tc.coef_off;
                                }
                                d.roi_shift = (uint8_t)roi_of(tileno, c);
                                d.band = (uint16_t)bd.band_index;
                                if (bands.size() > 65535)
                                    return B200J2K_ERR_PATCHWELCOME;
                                max_w = max_w > d.w ? max_w : d.w;
                                max_h = max_h > d.h ? max_h : d.h;
                                if ((uint32_t)d.w * d.h > max_area) max_area = (uint32_t)d.w * d.h;
                                if ((uint32_t)(d.w + 3) * (d.h + 3) > max_state) max_state = (uint32_t)(d.w + 3) * (d.h + 3);
                                cblk_template_.push_back(d);
                            }
                        }
                        pg.tag_nodes = tagtree_nodes(pg.ncw, pg.nch);
                        pg.tag_base = tag_total_;
                        tag_total_ += 2 * pg.tag_nodes;
                        precs_.push_back(pg);
                    }
                }
            }
            /* LL ping-pong space: the largest intermediate resolution of this plane */
            if (tc.nlevels >= 1) {
                const LevelDesc &l = tc.lev[tc.nlevels - 1];
                uint64_t n = (uint64_t)(l.u1 - l.u0) * (uint64_t)(l.v1 - l.v0);
                (void)n;
            }
            tcs.push_back(tc);
        }
        /* pack groups (mct_decode + write_frame, jpeg2000dec.c:2183-2209, :2301-2366) */
        if (!planar) {
            PackGroup g;
            memset(&g, 0, sizeof(g));
            g.ncomp = h.ncomp;
            g.plane = 0;
            g.pixel_stride = h.ncomp;
            g.wide = wide;
            for (int c = 0; c < h.ncomp; c++) {
                g.tc[c] = first_tc + c;
                g.sample_off[c] = c;
                g.shift[c] = out_precision - h.cbps[c];
                tcs[first_tc + c].group = (int)groups.size();
            }
            const CodSty *tcod = tsty_[tileno].cod;
            bool ok = tcod[0].mct != 0 && h.ncomp >= 3;
            for (int i = 1; i < 3 && ok; i++)
                ok = tcod[0].transform == tcod[i].transform && !memcmp(&t.comp[0], &t.comp[i], sizeof(Rect));
            g.mct = ok;
            groups.push_back(g);
        } else {
            /* planar: component c goes to plane cdef[c] - 1, an unassociated one to the last plane
             * (write_frame, jpeg2000dec.c:2325-2326).  The component transform does not care about the
             * output format (mct_decode, :2183-2209): a 4:4:4 stream with the MCT bit set keeps comps 0-2 in
             * one group so that RCT / ICT runs before the samples part ways. */
            const CodSty *tcod = tsty_[tileno].cod;
            bool mct_ok = tcod[0].mct != 0 && h.ncomp >= 3;
            for (int i = 1; i < 3 && mct_ok; i++)
                mct_ok = tcod[0].transform == tcod[i].transform && !memcmp(&t.comp[0], &t.comp[i], sizeof(Rect));
            int c = 0;
            if (mct_ok) {
                PackGroup g;
                memset(&g, 0, sizeof(g));
                g.ncomp = 3;
                g.mct = 1;
                g.pixel_stride = 1;
                g.wide = wide;
                for (; c < 3; c++) {
                    g.tc[c] = first_tc + c;
                    g.comp_plane[c] = h.cdef[c] ? h.cdef[c] - 1 : h.ncomp - 1;
                    g.shift[c] = out_precision - h.cbps[c];
                    tcs[first_tc + c].group = (int)groups.size();
                }
                g.plane = g.comp_plane[0];
                groups.push_back(g);
            }
            for (; c < h.ncomp; c++) {
                PackGroup g;
                memset(&g, 0, sizeof(g));
                g.ncomp = 1;
                g.tc[0] = first_tc + c;
                g.plane = g.comp_plane[0] = h.cdef[c] ? h.cdef[c] - 1 : h.ncomp - 1;
                g.pixel_stride = 1;
                g.shift[0] = out_precision - h.cbps[c];
                g.wide = wide;
                tcs[first_tc + c].group = (int)groups.size();
                groups.push_back(g);
            }
        }
    }
    ll_elems = coef_elems;   /* same offsets are used in the ping-pong buffers; they only hold <= 1/4 of each plane */
    /* Every component has to fit the plane it is written to.  The plane of a planar format comes from the 'cdef'
     * box of the file: a crafted one can send a full-size component to a subsampled chroma plane.  The reference's
     * write_frame then writes past the end of the AVFrame plane on the host; here the same write would leave
     * the picture allocation on the device, so such a frame is refused. */
    for (const PackGroup &g : groups)
        for (int c = 0; c < g.ncomp; c++) {
            const TileCompDesc &tc = tcs[g.tc[c]];
            const LevelDesc &o = tc.lev[tc.nlevels];
            const int pl = g.pixel_stride == g.ncomp && g.ncomp > 1 ? g.plane : (g.ncomp > 1 ? g.comp_plane[c] : g.plane);
            if (pl < 0 || pl >= nplanes || tc.out_x < 0 || tc.out_y < 0 ||
                (int64_t)tc.out_x + (o.u1 - o.u0) > plane_w[pl] || (int64_t)tc.out_y + (o.v1 - o.v0) > plane_h[pl])
                return B200J2K_ERR_INVALIDDATA;
        }
    return 0;
}

/* ---- packet-header bit reader (jpeg2000dec.c:70-90) ------------------------------------------------ */
inline int Parser::bits(int n)
{
    int res = 0;
    while (--n >= 0) {
        res <<= 1;
        if (bit_index_ == 0) {
            unsigned c = bp_ < bend_ ? *bp_++ : 0;
            bit_index_ = 7 + (c != 0xFF);
        }
        bit_index_--;
        res |= ((bp_ < bend_ ? *bp_ : 0) >> bit_index_) & 1;
    }
    return res;
}

inline void Parser::flush()
{
    unsigned c = bp_ < bend_ ? *bp_++ : 0;
    if (c == 0xFF && bp_ < bend_)
        bp_++;
    bit_index_ = 8;
}

/* tag tree walk on flat per-frame node state (jpeg2000dec.c:93-131) */
int Parser::tagtree_decode(uint32_t base, int ncw, int nch, int leaf, int threshold)
{
    uint32_t stack[16];
    int sp = -1;
    int w = ncw, h = nch, x = leaf % ncw, y = leaf / ncw;
    uint32_t lvl = base;
    int cur;
    bool found = false;
    uint32_t node = 0;
    for (;;) {
        node = lvl + (uint32_t)y * w + x;
        if (tags_[node].vis) {
            found = true;
            break;
        }
        stack[++sp] = node;
        if (w <= 1 && h <= 1)
            break;
        lvl += (uint32_t)w * h;
        w = (w + 1) >> 1; h = (h + 1) >> 1;
        x >>= 1; y >>= 1;
    }
    cur = found ? tags_[node].val : tags_[stack[sp]].val;
    while (cur < threshold && sp >= 0) {
        TagNodeState &n = tags_[stack[sp]];
        if (cur < n.val)
            cur = n.val;
        while (cur < threshold) {
            if (bits(1)) {
                n.vis++;
                break;
            }
            cur++;
        }
        n.val = (int16_t)cur;
        sp--;
    }
    return cur;
}

/* select_header (jpeg2000dec.c:1099-1108): the tile-part's run of PPM packet headers; when it is used up the
 * reference moves on to the NEXT tile-part and reads the header from that one's data */
void Parser::select_header(int tileno, int *tp)
{
    auto &pv = parts_[tileno];
    const uint8_t *none = pkt_ + pkt_size_;
    if (*tp < (int)pv.size() && pv[*tp].hp) {
        bp_ = pv[*tp].hp;
        bend_ = pv[*tp].hend;
    } else {
        bp_ = bend_ = none;
    }
    if (bp_ >= bend_ && bit_index_ == 8 && *tp < 31) {
        ++*tp;
        if (*tp < (int)pv.size() && pv[*tp].p) {
            bp_ = pv[*tp].p;
            bend_ = pv[*tp].end;
        } else {
            bp_ = bend_ = none;
        }
    }
}

void Parser::select_part(int tileno, int *tp, int csty)
{
    /* select_stream (jpeg2000dec.c:1111-1134): the tile's parts are 32 slots indexed by TPsot, a slot that never saw
     * its SOD is an empty stream; an exhausted or empty one hands over to the next slot, over gaps in the numbering,
     * up to the last slot.  What the decoder reports as consumed at the end is the position in the stream selected
     * last -- zero for an empty slot. */
    auto &pv = parts_[tileno];
    auto slot = [&](int i) -> const TilePart * { return i < (int)pv.size() && pv[i].p ? &pv[i] : nullptr; };
    const TilePart *cur = slot(*tp);
    while ((!cur || cur->p >= cur->end) && bit_index_ == 8 && *tp < 31)
        cur = slot(++*tp);
    if (cur) {
        bp_ = cur->p;
        bend_ = cur->end;
        cur_part_start_ = cur->start;
    } else {
        bp_ = bend_ = cur_part_start_ = pkt_ + pkt_size_;
    }
    if (csty & 2)
        if (bend_ - bp_ >= 6 && rd32(bp_) == 0xFF910004u)
            bp_ += 6;
}

/* jpeg2000_decode_packet, HT-only paths */
int Parser::decode_packet(int tileno, int c, int r, int precno, int layno, int *tp)
{
    const CodSty &cs = tsty_[tileno].cod[c];
    const QntSty &qs = tsty_[tileno].qnt[c];
    const TileGeom &t = tiles_[tileno];
    const ResGeom &rl = res_[t.res_base[c] + r];
    const uint8_t *expn = qs.expn + (r ? 3 * (r - 1) + 1 : 0);
    const int roi0 = roi_of(tileno, 0);                  /* quirk Q6: component 0's shift for every component */

    uint8_t &dl = decoded_layers_[rl.band[0].prec_base + precno];
    if (layno < dl)
        return 0;
    dl = (uint8_t)(layno + 1);
    /* where the packet header is read from (jpeg2000dec.c:1147-1153) */
    const bool ppm = has_ppm_, ppt = !ppm && !has_ppt_.empty() && has_ppt_[tileno];
    if (ppm)
        select_header(tileno, tp);
    else if (ppt) {
        bp_ = ppt_[tileno].data() + ppt_pos_[tileno];
        bend_ = ppt_[tileno].data() + ppt_live_[tileno];
    } else
        select_part(tileno, tp, cs.csty);

    /* blocks this packet contributes to, each with its run of segment lengths in `seglen` (as many as the
     * contribution has segments: the reference sizes lengthinc[] by the pass count, jpeg2000dec.c:1227) */
    struct Inc { uint32_t cb; uint32_t first, nseg; };
    static thread_local std::vector<Inc> incs;
    static thread_local std::vector<uint16_t> seglen;
    incs.clear();
    seglen.clear();

    if (!bits(1)) {
        flush();
    } else {
        for (int b = 0; b < rl.nbands; b++) {
            const BandGeom &bd = rl.band[b];
            if (bd.empty)
                continue;
            const PrecGeom &pg = precs_[bd.prec_base + precno];
            const int n = pg.ncw * pg.nch;
            for (int k = 0; k < n; k++) {
                CblkState &cb = state_[pg.cblk_base + k];
                int incl;
                if (!cb.incl) {
                    cb.modes = (uint8_t)cs.cblk_style;    /* :1177-1179 */
                    if (cb.modes >= 0x40)
                        cb.plhd = 1;
                    if (layno > 0)
                        tagtree_decode(pg.tag_base, pg.ncw, pg.nch, k, 1);
                    incl = tagtree_decode(pg.tag_base, pg.ncw, pg.nch, k, layno + 1) == layno;
                    if (incl) {
                        int zbp = tagtree_decode(pg.tag_base + pg.tag_nodes, pg.ncw, pg.nch, k, 100);
                        int v = expn[b] + qs.nguard - 1 - (zbp - roi0);
                        if (v < 0 || v > 30)
                            return B200J2K_ERR_INVALIDDATA;
                        cb.incl = 1;
                        cb.zbp = (uint8_t)zbp;
                        cb.nzb = (uint8_t)v;
                        cb.lblock = 3;
                    }
                } else {
                    incl = bits(1);
                }
                if (!incl)
                    continue;
                int newpasses;
                if (!bits(1)) newpasses = 1;
                else if (!bits(1)) newpasses = 2;
                else {
                    int nn = bits(2);
                    if (nn != 3) newpasses = 3 + nn;
                    else {
                        nn = bits(5);
                        newpasses = nn != 31 ? 6 + nn : 37 + bits(7);
                    }
                }
                if (cb.npasses + newpasses >= 100)
                    return B200J2K_ERR_PATCHWELCOME;
                int llen = 0;
                while (bits(1))
                    llen++;
                if (cb.lblock + llen + ilog2((unsigned)newpasses) > 16)
                    return B200J2K_ERR_PATCHWELCOME;
                cb.lblock = (uint8_t)(cb.lblock + llen);
                incs.push_back(Inc{ pg.cblk_base + (uint32_t)k, (uint32_t)seglen.size(), 0 });
                Inc &in = incs.back();

                /* Codeword-segment lengths of this contribution (jpeg2000dec.c:1256-1362; T.800 B.10.7):
                 * every length is a field of lblock + floor(log2(passes in the segment)) bits.  What differs between
                 * the three kinds of block is only how the `left` passes are cut into segments:
                 *   Part-1, style 0     one segment with all of them;
                 *   HT                  the passes come in sets of three -- cleanup | SigProp MagRef -- and a segment
                 *                       ends with the cleanup pass and with the set: HT_RUN[position in the set];
                 *   HT, placeholder     nothing of the block has bytes yet (all passes so far were empty sets): the
                 *                       segment runs up to the last cleanup pass of the contribution, and its length
                 *                       decides what the block is (below). */
                static const uint8_t HT_RUN[3] = { 1, 2, 1 };
                const bool mixed = (cb.modes & 0x80) != 0;
                int left = newpasses;
                auto field = [&](int passes) { return cb.lblock + ilog2((unsigned)passes); };
                auto take = [&](int passes, unsigned len) {
                    cb.npasses = (uint8_t)(cb.npasses + passes);
                    left -= passes;
                    seglen.push_back((uint16_t)len);
                    in.nseg++;
                };
                /* a MIXED block (SPcod bit 7) whose lengths show Part-1 passes: HT bit off for good (:1271-1274,
                 * :1303-1305, :1331-1333); bypass / termination on each pass are not served for those */
                auto unmask_part1 = [&]() -> bool {
                    if (cs.cblk_style & 0x05)
                        return false;
                    cb.plhd = 0;
                    cb.modes &= (uint8_t)~0x40;
                    return true;
                };
                if (cb.plhd) {
                    const int behind = (cb.npasses + left - 1) % 3;    /* refinement passes behind the last cleanup */
                    const int upto = left - behind;                    /* passes up to and including that cleanup */
                    unsigned len;
                    if (upto < 1) {
                        /* no cleanup pass in here: HT has nothing to say (zero), bytes mean Part-1 passes */
                        len = (unsigned)bits(field(left));
                        if (len && mixed && !unmask_part1())
                            return B200J2K_ERR_PATCHWELCOME;
                        take(left, len);
                    } else {
                        const int nb = field(upto);
                        len = (unsigned)bits(nb);
                        /* the first HT cleanup with bytes: always so in an HT-only stream; in a MIXED one the
                         * reference tells by the field -- grown lblock, more than one byte, top bit clear */
                        const bool cleanup = len && (!mixed || (cb.lblock > 3 && len > 1 && (len >> (nb - 1)) == 0));
                        if (cleanup) {
                            cb.plhd = 0;
                            cb.Lcup = len;
                            take(upto, len);
                        } else {
                            /* the field was that of ALL the passes: the missing low bits follow */
                            const bool bytes_seen = len != 0;
                            if (bytes_seen && !unmask_part1())
                                return B200J2K_ERR_PATCHWELCOME;
                            const int more = field(left) - nb;
                            len = (len << more) | (unsigned)bits(more);
                            if (!bytes_seen && len && mixed && !unmask_part1())
                                return B200J2K_ERR_PATCHWELCOME;
                            take(left, len);
                        }
                    }
                } else if (!(cb.modes & 0x40)) {
                    if (cs.cblk_style & 0x05)                           /* bypass / termination on each pass: not served */
                        return B200J2K_ERR_PATCHWELCOME;
                    take(left, (unsigned)bits(field(left)));            /* :1364-1368 */
                }
                /* HT block past its placeholders (possibly since two lines ago): cut by sets */
                while (left > 0) {
                    const int run = HT_RUN[cb.npasses % 3] < left ? HT_RUN[cb.npasses % 3] : left;
                    const unsigned len = (unsigned)bits(field(run));
                    cb.Lref += len;
                    take(run, len);
                }
            }
        }
        flush();
    }
    if (cs.csty & 4)
        if (bend_ - bp_ >= 2 && rd16(bp_) == 0xFF92)
            bp_ += 2;
    /* packed headers: remember where the header stream stands, continue in the tile-part's data (:1474-1480) */
    if (ppm) {
        if (*tp < (int)parts_[tileno].size()) {
            parts_[tileno][*tp].hp = bp_;
            parts_[tileno][*tp].hend = bend_;
        }
        select_part(tileno, tp, cs.csty);
    } else if (ppt) {
        ppt_pos_[tileno] = (size_t)(bp_ - ppt_[tileno].data());
        select_part(tileno, tp, cs.csty);
    }
    /* packet body: code-block bytes are referenced in place; a block that already has bytes elsewhere
     * is gathered into `extra` so that Dcup||Dref stays contiguous for the kernel */
    for (const Inc &in : incs) {
        CblkState &cb = state_[in.cb];
        for (uint32_t s = 0; s < in.nseg; s++) {
            uint32_t n = seglen[in.first + s];
            if ((uint32_t)(bend_ - bp_) < n)
                return B200J2K_ERR_INVALIDDATA;
            if (n) {
                const uint32_t here = (uint32_t)(bp_ - pkt_);
                if (cb.length == 0) {
                    cb.data_off = here;
                } else if (cb.data_off + cb.length != here) {
                    const uint8_t *old = cb.data_off < (uint64_t)pkt_size_ ? pkt_ + cb.data_off
                                                                           : extra.data() + (cb.data_off - pkt_size_);
                    std::vector<uint8_t> tmp(old, old + cb.length);
                    size_t at = extra.size();
                    extra.insert(extra.end(), tmp.begin(), tmp.end());
                    extra.insert(extra.end(), bp_, bp_ + n);
                    extra.resize((extra.size() + 15) & ~(size_t)15);
                    cb.data_off = (uint32_t)(pkt_size_ + at);
                }
                /* cblk->length is a uint16_t (jpeg2000.h:188): past 65535 bytes the count wraps, the block is what
                 * its first `length` bytes say and the next segment lands on top of the rest */
                cb.length = (cb.length + n) & 0xFFFF;
                bp_ += n;
            }
        }
    }
    if (*tp < (int)parts_[tileno].size() && parts_[tileno][*tp].start)      /* an empty slot stays empty */
        parts_[tileno][*tp].p = bp_;
    return 0;
}

int Parser::precinct_at(int tileno, int c, int r, int x, int y, bool cprl_rule) const
{
    const Header &h = hdr;
    const TileGeom &t = tiles_[tileno];
    const ResGeom &rl = res_[t.res_base[c] + r];
    const int red = tsty_[tileno].cod[c].nres - 1 - r;
    const Rect &cc = t.comp[c];
    if (cprl_rule) {
        int xc = x / h.cdx[c], yc = y / h.cdy[c];
        if (yc % (1LL << (rl.ppy + red)) && y != t.c.y0)
            return -1;
        if (xc % (1LL << (rl.ppx + red)) && x != t.c.x0)
            return -1;
        unsigned px = (unsigned)(ceildivpow2(xc, red) >> rl.ppx) - (unsigned)(ceildivpow2(cc.x0, red) >> rl.ppx);
        unsigned py = (unsigned)(ceildivpow2(yc, red) >> rl.ppy) - (unsigned)(ceildivpow2(cc.y0, red) >> rl.ppy);
        if (px >= (unsigned)rl.npx || py >= (unsigned)rl.npy)
            return -1;
        return (int)(px + rl.npx * py);
    }
    const int64_t sx = (int64_t)h.cdx[c] << red, sy = (int64_t)h.cdy[c] << red;
    int64_t trx0 = ((int64_t)t.c.x0 + sx - 1) / sx, try0 = ((int64_t)t.c.y0 + sy - 1) / sy;
    if (!(y % ((uint64_t)h.cdy[c] << (rl.ppy + red)) == 0 ||
          (y == t.c.y0 && ((try0 << red) % (1LL << (red + rl.ppy))))))
        return -1;
    if (!(x % ((uint64_t)h.cdx[c] << (rl.ppx + red)) == 0 ||
          (x == t.c.x0 && ((trx0 << red) % (1LL << (red + rl.ppx))))))
        return -1;
    unsigned px = (unsigned)((((int64_t)x + sx - 1) / sx) >> rl.ppx);
    unsigned py = (unsigned)((((int64_t)y + sy - 1) / sy) >> rl.ppy);
    px -= (unsigned)(ceildivpow2(cc.x0, red) >> rl.ppx);
    py -= (unsigned)(ceildivpow2(cc.y0, red) >> rl.ppy);
    if (px >= (unsigned)rl.npx || py >= (unsigned)rl.npy)
        return -2;                                        /* position accepted, precinct outside the grid */
    return (int)(px + rl.npx * py);
}

#define PACKET(c, r, pn, l)                                                       \
    do {                                                                          \
        int e_ = decode_packet(tileno, c, r, pn, l, &tp);                         \
        if (e_ < 0)                                                               \
            return e_;                                                            \
    } while (0)

/* jpeg2000_decode_packets_po_iteration (jpeg2000dec.c:1544-1828): one progression over the resolution range
 * [RS, RE), the component range [CS, CE) and the layers [0, LYE).  Packets a previous iteration already
 * delivered are skipped inside decode_packet (decoded_layers). */
int Parser::po_iteration(int tileno, int RS, int CS, int LYE, int RE, int CE, int P, int *tpp)
{
    const Header &h = hdr;
    const CodSty *tcod = tsty_[tileno].cod;          /* this tile's coding styles */
    const TileGeom &t = tiles_[tileno];
    int &tp = *tpp;
    auto R = [&](int c, int r) -> const ResGeom & { return res_[t.res_base[c] + r]; };
    switch (P) {
    case 0:
        for (int l = 0; l < LYE; l++) {
            bool ok = true;
            for (int r = RS; ok && r < RE; r++) {
                ok = false;
                for (int c = CS; c < CE; c++)
                    if (r < tcod[c].nres) {
                        ok = true;
                        for (int pn = 0; pn < R(c, r).npx * R(c, r).npy; pn++)
                            PACKET(c, r, pn, l);
                    }
            }
        }
        break;
    case 1: {
        bool ok = true;
        for (int r = RS; ok && r < RE; r++) {
            ok = false;
            for (int l = 0; l < LYE; l++)
                for (int c = CS; c < CE; c++)
                    if (r < tcod[c].nres) {
                        ok = true;
                        for (int pn = 0; pn < R(c, r).npx * R(c, r).npy; pn++)
                            PACKET(c, r, pn, l);
                    }
        }
        break;
    }
    case 2: {
        bool ok = true;
        for (int r = RS; ok && r < RE; r++) {
            ok = false;
            int sx = 30, sy = 30;
            for (int c = CS; c < CE; c++)
                if (r < tcod[c].nres) {
                    int red = tcod[c].nres - 1 - r;
                    sx = imin(sx, R(c, r).ppx + red);
                    sy = imin(sy, R(c, r).ppy + red);
                }
            sx = 1 << sx; sy = 1 << sy;
            for (int y = t.c.y0; y < t.c.y1; y = (y / sy + 1) * sy)
                for (int x = t.c.x0; x < t.c.x1; x = (x / sx + 1) * sx)
                    for (int c = CS; c < CE; c++) {
                        if (r >= tcod[c].nres)
                            continue;
                        int pn = precinct_at(tileno, c, r, x, y, false);
                        if (pn == -1)
                            continue;
                        ok = true;                        /* set before the "precinct exists" test (:1730) */
                        if (pn < 0)
                            continue;
                        for (int l = 0; l < LYE; l++)
                            PACKET(c, r, pn, l);
                    }
        }
        break;
    }
    case 3: {
        int sx = 32, sy = 32;
        for (int c = CS; c < CE; c++)
            for (int r = RS; r < imin(tcod[c].nres, RE); r++) {
                int red = tcod[c].nres - 1 - r;
                sx = imin(sx, R(c, r).ppx + red);
                sy = imin(sy, R(c, r).ppy + red);
            }
        if (sx >= 31 || sy >= 31)
            return B200J2K_ERR_PATCHWELCOME;
        sx = 1 << sx; sy = 1 << sy;
        for (int y = t.c.y0; y < t.c.y1; y = (y / sy + 1) * sy)
            for (int x = t.c.x0; x < t.c.x1; x = (x / sx + 1) * sx)
                for (int c = CS; c < CE; c++)
                    for (int r = RS; r < imin(tcod[c].nres, RE); r++) {
                        int pn = precinct_at(tileno, c, r, x, y, false);
                        if (pn < 0)
                            continue;
                        for (int l = 0; l < LYE; l++)
                            PACKET(c, r, pn, l);
                    }
        break;
    }
    case 4:
        for (int c = CS; c < CE; c++) {
            int sx = 32, sy = 32;
            if (RS >= imin(tcod[c].nres, RE))
                continue;
            for (int r = RS; r < imin(tcod[c].nres, RE); r++) {
                int red = tcod[c].nres - 1 - r;
                sx = imin(sx, R(c, r).ppx + red);
                sy = imin(sy, R(c, r).ppy + red);
            }
            if (sx >= 31 || sy >= 31)
                return B200J2K_ERR_PATCHWELCOME;
            sx = 1 << sx; sy = 1 << sy;
            for (int y = t.c.y0; y < t.c.y1; y = (y / sy + 1) * sy)
                for (int x = t.c.x0; x < t.c.x1; x = (x / sx + 1) * sx)
                    for (int r = RS; r < imin(tcod[c].nres, RE); r++) {
                        int pn = precinct_at(tileno, c, r, x, y, true);
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
}

/* jpeg2000_decode_packets (jpeg2000dec.c:1830-1869) */
int Parser::decode_packets(int tileno)
{
    const Header &h = hdr;
    const TileSty &T = tsty_[tileno];
    const int nlayers = T.cod[0].nlayers;
    int tp = 0, ret = 0;
    bit_index_ = 8;
    if (T.npoc) {
        for (int i = 0; i < T.npoc && ret >= 0; i++) {
            const Header::Poc &e = T.poc[i];
            ret = po_iteration(tileno, e.RS, e.CS, imin(e.LYE, nlayers), e.RE, imin(e.CE, h.ncomp), e.P, &tp);
        }
    } else {
        ret = po_iteration(tileno, 0, 0, nlayers, 33, h.ncomp, T.cod[0].prog, &tp);
    }
    return ret;
}
#undef PACKET

bool Parser::tile_selected(int flags, int tileno)
{
    const int step = B200J2K_FLAG_GET_TILE_STEP(flags);
    return step <= 1 || tileno % step == (int)B200J2K_FLAG_GET_TILE_FIRST(flags) % step;
}

int Parser::parse(const uint8_t *pkt, int64_t size, int flags, bool headers_only)
{
    if (!pkt || size < 2)
        return B200J2K_ERR_INVALIDDATA;
    if (size > 0x7FFFFFF0ll)
        return B200J2K_ERR_PATCHWELCOME;
    pkt_ = pkt;
    pkt_size_ = size;
    size_t mh = 0;
    int ret = parse_headers(pkt, size, flags, &mh);
    if (ret < 0)
        return ret;
    if (headers_only)
        return 0;
    /* geometry only depends on the main header: reuse it across a sequence */
    if (geom_flags_ != flags || geom_key_.size() != mh || memcmp(geom_key_.data(), pkt, mh) ||
        geom_extra_key_ != geom_extra_) {
        geom_flags_ = -1;
        if ((ret = build_geometry(flags)) < 0)
            return ret;
        geom_key_.assign(pkt, pkt + mh);
        geom_extra_key_ = geom_extra_;
        geom_flags_ = flags;
        geom_serial++;
    }
    /* per-frame tier-2 state */
    state_.assign(cblk_template_.size(), CblkState{});
    tags_.assign(tag_total_, TagNodeState{ 0, 0 });
    decoded_layers_.assign(precs_.size(), 0);
    extra.clear();
    /* s->g of the reference, whose position is what decode() returns in the end (jpeg2000dec.c:2903): the packet as the
     * marker loop left it, until a tile selects one of its tile-part streams (select_part) */
    cur_part_start_ = pkt_;
    bp_ = pkt_ + consumed;
    bend_ = pkt_ + pkt_size_;
    tile_subset = B200J2K_FLAG_GET_TILE_STEP(flags) > 1;
    needed_ranges.clear();
    for (int tileno = 0; tileno < hdr.ntx * hdr.nty; tileno++) {
        /* a tile whose first tile-part (TPsot = 0) never showed up did not inherit the main header's coding styles:
         * unless a COD of its own filled them in, init_tile fails on !codsty->init (jpeg2000dec.c:1054-1055) and the
         * frame with it */
        for (int c = 0; c < hdr.ncomp; c++)
            if (!tsty_[tileno].cod[c].init)
                return B200J2K_ERR_INVALIDDATA;
        /* a transform byte other than 0 / 1: ff_jpeg2000_dwt_init has no such type and init_tile fails with its
         * bare -1 (jpeg2000dwt.c:551-591 via jpeg2000.c:477) */
        for (int c = 0; c < hdr.ncomp; c++)
            if (tsty_[tileno].cod[c].transform > J2K_DWT97_INT)
                return B200J2K_ERR_UNSPECIFIED;
        if (!tile_selected(flags, tileno))
            continue;                                     /* another context's tile: its packets are not even read */
        if (tile_subset)
            for (const TilePart &tp : parts_[tileno])
                if (tp.start && tp.end > tp.start)
                    needed_ranges.emplace_back((uint32_t)(tp.start - pkt), (uint32_t)(tp.end - tp.start));
        if ((ret = decode_packets(tileno)) < 0)
            return ret;
        /* "EOC marker reached": every tile ends with bytestream2_skip(&s->g, 2) on whatever stream is current -- the
         * tile-part stream its last packet came from, or still the previous one when the tile had no packet to read
         * (:1865-1866) */
        if (bp_ && bend_ > bp_)
            bp_ += bend_ - bp_ < 2 ? bend_ - bp_ : 2;
    }
    if (cur_part_start_ && bp_ && bp_ >= cur_part_start_)
        consumed = (int64_t)(bp_ - cur_part_start_);
    /* emit the descriptor table */
    cblk = cblk_template_;
    max_Lcup = max_Lref = 0;
    any_refine = 0;
    for (size_t i = 0; i < cblk.size(); i++) {
        const CblkState &s = state_[i];
        CblkDesc &d = cblk[i];
        d.data_off = s.data_off;
        d.npasses = s.npasses;
        d.zbp = s.zbp;
        d.length = s.length;
        if (s.npasses && !(d.flags & J2K_CBLK_SKIP) && band_negative_Mb_[d.band])
            return B200J2K_ERR_PATCHWELCOME;
        if ((d.flags & J2K_CBLK_MQ) || (s.incl && !(s.modes & 0x40))) {
            d.flags |= J2K_CBLK_MQ;                       /* also the Part-1 blocks of a MIXED set */
            d.Lref = d.pad;
            /* cblk->nonzerobits = expn + guard bits - 1 - (zero bit-planes - ROI shift of component 0)
             * (jpeg2000dec.c:1186-1194); Lref keeps the band orientation set with the geometry */
            d.Lcup = s.nzb;
            continue;
        }
        d.Lcup = (uint16_t)s.Lcup;
        d.Lref = (uint16_t)s.Lref;
        if (s.npasses) {
            /* the decoder reads Dcup[0..Lcup) and Dref[0..Lref): keep both inside the gathered bytes */
            if ((uint32_t)d.Lcup + d.Lref > s.length) {
                /* damaged lengths: clamp so that the kernel never leaves the buffer (block decodes as the
                 * reference would on short data: garbage or error => zeros) */
                if (d.Lcup > s.length) { d.Lcup = (uint16_t)s.length; d.Lref = 0; }
                else d.Lref = (uint16_t)(s.length - d.Lcup);
            }
            if (d.Lcup > max_Lcup) max_Lcup = d.Lcup;
            if (d.Lref > max_Lref) max_Lref = d.Lref;
            int rem = s.npasses % 3;
            if (rem != 1)
                any_refine = 1;
        }
    }
    return 0;
}

}  // namespace b200j2k
