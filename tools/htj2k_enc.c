/*
 * tools/htj2k_enc.c -- synthetic HTJ2K (Rec. ITU-T T.814 / ISO/IEC 15444-15) codestream generator.
 *
 * TEST / BENCHMARK WORKLOAD INFRASTRUCTURE.  The reference (sandflow/ffmpeg-ht) has no HT encoder and
 * ships no .j2k sample (hard part #1), so every codestream the tests and bench.py
 * decode is produced here, on the CPU, from a seeded synthetic image.  Nothing in this file is on the
 * product's decode path.
 *
 * What it writes is exactly the subset of the syntax the reference's native decoder consumes on its HT
 * path: SOC SIZ CAP COD QCD, one tile-part per tile (SOT SOD), packets in
 * LRCP/RLCP/RPCL/PCRL/CPRL order with optional user precincts, SOP/EPH, one quality layer, HT code-blocks
 * with a cleanup pass and optionally one SigProp + one MagRef pass, EOC.
 *
 * The block coder follows T.814 clause 7 as read back from the reference decoder
 * (libavcodec/jpeg2000htdec.c:547-1185) and the encoder-side procedures of SURVEY.md Appendix D: MEL run
 * coder, CxtVLC entry choice by (context, rho, EMB consistency), U-VLC, MagSgn, backward VLC/MagRef byte
 * streams with their stuffing rule.  Wavelets: forward 5/3 (exactly reversible against
 * libavcodec/jpeg2000dwt.c:309-374) and forward 9/7 lifting without K scaling (the convention the
 * reference's inverse, :376-451, and its step-size folding, libavcodec/jpeg2000.c:251-264, expect).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cxtvlc_tables.h"

typedef struct J2kGenParams {
    int32_t width, height, ncomp, bitdepth;
    int32_t tile_w, tile_h;      /* 0 = one tile */
    int32_t levels;              /* decomposition levels N_L */
    int32_t cbw_log2, cbh_log2;  /* code-block size exponents (2..10, sum <= 12) */
    int32_t reversible;          /* 1: 5/3 (+RCT)  0: 9/7 float (+ICT) */
    int32_t mct;                 /* 1: component transform on comps 0..2 */
    int32_t prog;                /* 0 LRCP 1 RLCP 2 RPCL 3 PCRL 4 CPRL */
    int32_t use_precincts;       /* 0: default 2^15 precincts */
    int32_t prec_exps[33];       /* PPx | PPy<<4 for r = 0..levels */
    int32_t guard_bits;
    int32_t npasses;             /* 1 cleanup, 2 +SigProp, 3 +SigProp+MagRef */
    int32_t sop, eph;
    int32_t vsc;                 /* stripe-causal code-block style bit (0x08) */
    float   base_step;           /* irreversible: quantiser step in sample units, same for every band */
    int32_t x_off, y_off;        /* XOsiz, YOsiz (also used as XTOsiz, YTOsiz) */
    int32_t stats_blocks;        /* out: number of non-empty code-blocks */
    int32_t stats_max_lcup;      /* out */
    int32_t error;               /* out: 0 ok, 1 exponent bound exceeded (raise guard_bits), 2 segment too long */
    int32_t layers;              /* 0/1: one quality layer.  2: every block is spread over two layers in one of
                                  * four ways chosen by a hash of its position -- all passes in layer 0; cleanup in
                                  * layer 0 and refinement in layer 1; first included in layer 1; a triple of empty
                                  * placeholder passes in layer 0 and the HT set in layer 1 -- and half of the blocks
                                  * with spare zero bit-planes carry an extra placeholder triple in front of the HT set
                                  * (jpeg2000dec.c:1256-1362) */
    int32_t sub_x[4], sub_y[4];  /* XRsiz / YRsiz per component (0 = 1).  With subsampling: no offset, no component
                                  * transform, LRCP / RLCP only; plane c holds ceil(W/sub_x) x ceil(H/sub_y) samples */
    int32_t npoc;                /* > 0: a POC marker segment in the main header with these entries; the packets
                                  * follow them in order, each packet once (the first entry that reaches it).
                                  * poc_split > 0 puts the first poc_split entries into a POC segment of their own */
    int32_t poc[8][6];           /* RSpoc, CSpoc, LYEpoc, REpoc, CEpoc, Ppoc */
    int32_t poc_split;
    int32_t packed_headers;      /* 0: packet headers in front of their bodies.  1: PPT -- the packet headers of a tile
                                  * (EPH markers included) move into PPT marker segments of its tile-part header and
                                  * Ccap15 bit 11 is set (jpeg2000dec.c:2605-2617).  2: PPM -- into PPM marker segments
                                  * of the main header, one Nppm-prefixed run per tile-part (:958-980, :2467-2472) */
    int32_t ppx_chunk;           /* payload bytes per PPT / PPM marker segment (0 = as many as fit: 65532) */
    int32_t tp_packets;          /* > 0: a new tile-part after every tp_packets packets (at most 32 parts per tile, the
                                  * reference's limit); PPT headers all go to the first part (jpeg2000dec.c:995), PPM
                                  * gets one Nppm-prefixed run per tile-part */
    int32_t tp_interleave;       /* with tp_packets: write part 0 of every tile, then part 1 of every tile, ... */
    int32_t tile_var;            /* != 0: every odd tile is coded differently and says so in its first tile-part header
                                  * (Ccap15 bit 11 set: not the HOMOGENEOUS set, jpeg2000dec.c:2514-2563).  Bits:
                                  * 1 another progression order (COD), 2 one decomposition level less (COD + QCD),
                                  * 4 narrower code-blocks (COD), 8 one more guard bit / a coarser step (QCD),
                                  * 16 a POC marker segment with two entries (only without a main-header POC) */
    int32_t ht_set;              /* Ccap15 bits 14-15 / SPcod bits 6-7 (jpeg2000dec.c:443-459, :1060, :1256-1340):
                                  * 0 HTONLY, style 0x40.  1 HTDECLARED (Ccap15 0x4000), style 0x40.
                                  * 3 MIXED (0xC000), style 0xC0, cleanup lengths written the way a MIXED stream must
                                  *   (Lblock > 3, the top bit of the length field 0) so every block reads as HT.
                                  * 4 MIXED, style 0xC0, lengths written as for HTONLY: the reference then takes the
                                  *   blocks whose field has Lblock == 3 or its top bit set for Part-1 (MQ) blocks and
                                  *   decodes their HT bytes with the MQ decoder (parity material, not a picture).
                                  * 5 like 4 with Ccap15 saying HTONLY (the reference does not cross-check that) */
} J2kGenParams;

/* ------------------------------------------------------------------------------------------------ */
/* growable byte buffer                                                                             */

typedef struct Buf {
    uint8_t *p;
    size_t n, cap;
} Buf;

static void buf_reserve(Buf *b, size_t extra)
{
    if (b->n + extra > b->cap) {
        size_t c = b->cap ? b->cap * 2 : 4096;
        while (c < b->n + extra)
            c *= 2;
        b->p = realloc(b->p, c);
        b->cap = c;
    }
}
static void buf_put(Buf *b, uint8_t v) { buf_reserve(b, 1); b->p[b->n++] = v; }
static void buf_put16(Buf *b, unsigned v) { buf_put(b, v >> 8); buf_put(b, v & 255); }
static void buf_put32(Buf *b, uint32_t v) { buf_put16(b, v >> 16); buf_put16(b, v & 0xFFFF); }
static void buf_append(Buf *b, const uint8_t *s, size_t n) { buf_reserve(b, n); memcpy(b->p + b->n, s, n); b->n += n; }

static inline int ceildiv(int a, int b) { return (a + b - 1) / b; }
static inline int ceildivpow2(int a, int b) { return (int)(((int64_t)a + ((int64_t)1 << b) - 1) >> b); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------------ */
/* HT block coder: bit-stream writers                                                               */

/* forward, LSB-first, the byte after 0xFF carries 7 bits (MagSgn and SigProp) */
typedef struct FwdLsb {
    Buf b;
    uint32_t acc;
    int nbits, limit;
} FwdLsb;

static void fwd_init(FwdLsb *s) { memset(s, 0, sizeof(*s)); s->limit = 8; }
static void fwd_put(FwdLsb *s, uint32_t v, int n)
{
    while (n > 0) {
        int take = imin(n, s->limit - s->nbits);
        s->acc |= (v & ((1u << take) - 1)) << s->nbits;
        s->nbits += take;
        v >>= take;
        n -= take;
        if (s->nbits == s->limit) {
            buf_put(&s->b, (uint8_t)s->acc);
            s->limit = (s->acc == 0xFF) ? 7 : 8;
            s->acc = 0;
            s->nbits = 0;
        }
    }
}
/* MagSgn termination: pad with ones, drop a final 0xFF (the decoder synthesises 0xFF past the end) */
static void fwd_finish_magsgn(FwdLsb *s)
{
    if (s->nbits) {
        uint32_t pad = (1u << (s->limit - s->nbits)) - 1;
        fwd_put(s, pad, s->limit - s->nbits);
    }
    if (s->b.n && s->b.p[s->b.n - 1] == 0xFF)
        s->b.n--;
}
/* SigProp termination: pad with zeros (the decoder feeds zeros past the end) */
static void fwd_finish_zero(FwdLsb *s)
{
    if (s->nbits)
        fwd_put(s, 0, s->limit - s->nbits);
}

/* MEL: forward, MSB-first, the byte after 0xFF carries 7 bits */
typedef struct MelEnc {
    Buf b;
    uint32_t acc;
    int nbits, limit;
    int k, run;
} MelEnc;

static const uint8_t MEL_E[13] = { 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5 };

static void mel_init(MelEnc *m) { memset(m, 0, sizeof(*m)); m->limit = 8; }
static void mel_bit(MelEnc *m, int bit)
{
    m->acc = (m->acc << 1) | (bit & 1);
    if (++m->nbits == m->limit) {
        buf_put(&m->b, (uint8_t)m->acc);
        m->limit = (m->acc == 0xFF) ? 7 : 8;
        m->acc = 0;
        m->nbits = 0;
    }
}
static void mel_sym(MelEnc *m, int sym)
{
    int e = MEL_E[m->k];
    if (!sym) {
        if (++m->run == (1 << e)) {
            mel_bit(m, 1);
            m->run = 0;
            m->k = imin(12, m->k + 1);
        }
    } else {
        mel_bit(m, 0);
        for (int i = e - 1; i >= 0; i--)
            mel_bit(m, (m->run >> i) & 1);
        m->run = 0;
        m->k = imax(0, m->k - 1);
    }
}
static void mel_finish(MelEnc *m)
{
    if (m->run > 0)
        mel_bit(m, 1);
    if (m->nbits) {
        int pad = m->limit - m->nbits;
        m->acc <<= pad;
        buf_put(&m->b, (uint8_t)m->acc);
        m->nbits = 0;
    }
}

/* backward, LSB-first (VLC and MagRef).  A byte closes at 8 bits, or at 7 when the previously closed
 * byte is > 0x8F and the 7 bits are all ones (the MSB stays 0 = stuffing).  Bytes are produced in
 * reverse memory order. */
typedef struct BwdLsb {
    Buf b;      /* bytes in emission order (reverse of memory order) */
    uint32_t acc;
    int nbits;
    unsigned prev;
} BwdLsb;

static void bwd_init(BwdLsb *s, int vlc)
{
    memset(s, 0, sizeof(*s));
    s->prev = 0xFF;
    if (vlc) {          /* low nibble of Dcup[Lcup-2] is reserved for Scup and reads as 0xF */
        s->acc = 0xF;
        s->nbits = 4;
    }
}
static void bwd_put(BwdLsb *s, uint32_t v, int n)
{
    for (int i = 0; i < n; i++) {
        s->acc |= ((v >> i) & 1u) << s->nbits;
        s->nbits++;
        if (s->nbits == 8 || (s->nbits == 7 && s->prev > 0x8F && s->acc == 0x7F)) {
            buf_put(&s->b, (uint8_t)s->acc);
            s->prev = s->acc;
            s->acc = 0;
            s->nbits = 0;
        }
    }
}
static void bwd_finish(BwdLsb *s)
{
    if (s->nbits) {
        buf_put(&s->b, (uint8_t)s->acc);
        s->prev = s->acc;
        s->acc = 0;
        s->nbits = 0;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* CxtVLC encoder tables, built by inverting the decode tables                                      */

typedef struct VlcCand {
    uint8_t cw, len, emb_k, emb_1;
} VlcCand;

/* [table][context][rho][u_off] -> candidates */
static VlcCand g_cand[2][8][16][2][16];
static uint8_t g_ncand[2][8][16][2];
static int g_tables_ready;

static void build_enc_tables(void)
{
    if (g_tables_ready)
        return;
    for (int t = 0; t < 2; t++) {
        const unsigned short *tab = t == 0 ? CXTVLC_FIRST_ROW : CXTVLC_OTHER_ROWS;
        for (int ctx = 0; ctx < 8; ctx++)
            for (int cw = 0; cw < 128; cw++) {
                unsigned e = tab[cw | (ctx << 7)];
                int len = e & 7, u_off = (e >> 3) & 1, rho = (e >> 4) & 15;
                int ek = (e >> 8) & 15, e1 = (e >> 12) & 15;
                if (cw >> len)
                    continue;          /* keep the canonical (zero-extended) index only */
                uint8_t *n = &g_ncand[t][ctx][rho][u_off];
                if (*n < 16)
                    g_cand[t][ctx][rho][u_off][(*n)++] = (VlcCand){ (uint8_t)cw, (uint8_t)len, (uint8_t)ek, (uint8_t)e1 };
            }
    }
    g_tables_ready = 1;
}

static inline int popcount4(int v) { return (v & 1) + ((v >> 1) & 1) + ((v >> 2) & 1) + ((v >> 3) & 1); }
static inline int bitlen(uint32_t v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

static void put_uvlc_prefix(BwdLsb *v, int u)
{
    if (u == 1) bwd_put(v, 1, 1);
    else if (u == 2) bwd_put(v, 2, 2);
    else if (u <= 4) bwd_put(v, 4, 3);
    else bwd_put(v, 0, 3);
}
static void put_uvlc_suffix(BwdLsb *v, int u)
{
    if (u < 3) return;
    if (u <= 4) { bwd_put(v, u - 3, 1); return; }
    if (u <= 32) { bwd_put(v, u - 5, 5); return; }
    bwd_put(v, 28 + ((u - 33) & 3), 5);
}
static void put_uvlc_ext(BwdLsb *v, int u)
{
    if (u >= 33) {
        int sfx = 28 + ((u - 33) & 3);
        bwd_put(v, (u - 5 - sfx) >> 2, 4);
    }
}

typedef struct HtBlockOut {
    Buf data;           /* Dcup || Dref */
    int Lcup, Lref;
    int error;
} HtBlockOut;

/*
 * Encode one code-block.  mag/sgn: w*h magnitudes (already reduced to the cleanup bit-plane, i.e. the
 * full magnitude >> (npasses > 1)) and signs; fullmag: full magnitudes (for the refinement passes).
 * maxbp: the decoder's bound on U (S_blk + 1).
 */
static void ht_encode_block(const uint32_t *fullmag, const uint8_t *sgn, int w, int h, int npasses,
                            int vsc, int maxbp, HtBlockOut *out)
{
    const int QW = (w + 1) >> 1, QH = (h + 1) >> 1;
    const int refine = npasses > 1;
    FwdLsb ms; MelEnc mel; BwdLsb vlc;
    fwd_init(&ms); mel_init(&mel); bwd_init(&vlc, 1);
    build_enc_tables();

    uint8_t *sig = calloc((size_t)4 * QW * QH, 1);   /* sigma per quad sample */
    uint8_t *Eq = calloc((size_t)4 * QW * QH, 1);    /* exponent per quad sample */
    out->error = 0;

#define MAG(y, x) (((y) < h && (x) < w) ? (fullmag[(y) * w + (x)] >> refine) : 0u)
#define SGN(y, x) (((y) < h && (x) < w) ? sgn[(y) * w + (x)] : 0)

    for (int qy = 0; qy < QH; qy++) {
        int ctx_prev_first = 0;  /* running context inside the first quad row */
        for (int qx0 = 0; qx0 < QW; qx0 += 2) {
            int npair = imin(2, QW - qx0);
            int rho[2] = { 0, 0 }, ctx[2] = { 0, 0 }, U[2] = { 0, 0 }, u[2] = { 0, 0 }, uoff[2] = { 0, 0 };
            int kappa[2] = { 1, 1 };
            uint32_t v[2][4];
            int E[2][4];
            VlcCand chosen[2];
            int coded[2] = { 0, 0 };   /* has a CxtVLC codeword */
            for (int j = 0; j < npair; j++) {
                int qx = qx0 + j, q = qy * QW + qx;
                int emax = 0;
                for (int n = 0; n < 4; n++) {
                    int y = 2 * qy + (n & 1), x = 2 * qx + (n >> 1);
                    uint32_t m = MAG(y, x);
                    if (m) {
                        rho[j] |= 1 << n;
                        v[j][n] = 2 * (m - 1) + SGN(y, x);
                        E[j][n] = bitlen(v[j][n] | 1);
                    } else {
                        v[j][n] = 0;
                        E[j][n] = 0;
                    }
                    sig[4 * q + n] = m != 0;
                    Eq[4 * q + n] = (uint8_t)E[j][n];
                    emax = imax(emax, E[j][n]);
                }
                /* context (libavcodec/jpeg2000htdec.c:646-649 first row, :788-796 other rows) */
                if (qy == 0) {
                    ctx[j] = ctx_prev_first;
                    ctx_prev_first = (sig[4 * q] | sig[4 * q + 1]) + 2 * sig[4 * q + 2] + 4 * sig[4 * q + 3];
                } else {
                    int qn = q - QW;
                    int c = sig[4 * qn + 1] + 4 * sig[4 * qn + 3];
                    if (qx > 0) {
                        c |= sig[4 * (qn - 1) + 3];
                        c += 2 * (sig[4 * (q - 1) + 3] | sig[4 * (q - 1) + 2]);
                    }
                    if (qx < QW - 1)
                        c |= 4 * sig[4 * (qn + 1) + 1];
                    ctx[j] = c;
                    /* exponent predictor (:855-885) */
                    int gamma = popcount4(rho[j]) > 1;
                    int en = Eq[4 * qn + 1], ene = Eq[4 * qn + 3];
                    int enw = qx > 0 ? Eq[4 * (qn - 1) + 3] : 0;
                    int enf = qx < QW - 1 ? Eq[4 * (qn + 1) + 1] : 0;
                    int me = imax(imax(en, ene), imax(enw, enf));
                    kappa[j] = imax(1, gamma * (me - 1));
                }
                U[j] = imax(kappa[j], emax);
                u[j] = U[j] - kappa[j];
                uoff[j] = u[j] > 0;
                if (U[j] > maxbp)
                    out->error = 1;
                /* MEL symbol for a context-0 quad, then the CxtVLC codeword */
                if (ctx[j] == 0)
                    mel_sym(&mel, rho[j] != 0);
                if (ctx[j] != 0 || rho[j] != 0) {
                    int t = qy == 0 ? 0 : 1;
                    int etrue = 0;
                    for (int n = 0; n < 4; n++)
                        if (E[j][n] == U[j] && (rho[j] >> n & 1))
                            etrue |= 1 << n;
                    const VlcCand *cands = g_cand[t][ctx[j]][rho[j]][uoff[j]];
                    int nc = g_ncand[t][ctx[j]][rho[j]][uoff[j]], best = -1;
                    for (int i = 0; i < nc; i++) {
                        if (uoff[j] && cands[i].emb_1 != (etrue & cands[i].emb_k))
                            continue;
                        if (best < 0 || popcount4(cands[i].emb_k) > popcount4(cands[best].emb_k))
                            best = i;
                    }
                    if (best < 0) {     /* cannot happen for the T.814 tables */
                        out->error = 3;
                        best = 0;
                    }
                    chosen[j] = cands[best];
                    coded[j] = 1;
                    bwd_put(&vlc, chosen[j].cw, chosen[j].len);
                } else {
                    chosen[j] = (VlcCand){ 0, 0, 0, 0 };
                }
            }
            /* U-VLC (:671-712 first row, :833-854 other rows) */
            if (npair == 2 && uoff[0] && uoff[1]) {
                if (qy == 0) {
                    if (u[0] > 2 && u[1] > 2) {
                        mel_sym(&mel, 1);
                        int a = u[0] - 2, b = u[1] - 2;
                        put_uvlc_prefix(&vlc, a); put_uvlc_prefix(&vlc, b);
                        put_uvlc_suffix(&vlc, a); put_uvlc_suffix(&vlc, b);
                        put_uvlc_ext(&vlc, a); put_uvlc_ext(&vlc, b);
                    } else {
                        mel_sym(&mel, 0);
                        put_uvlc_prefix(&vlc, u[0]);
                        if (u[0] > 2) {
                            bwd_put(&vlc, u[1] - 1, 1);
                            put_uvlc_suffix(&vlc, u[0]);
                            put_uvlc_ext(&vlc, u[0]);
                        } else {
                            put_uvlc_prefix(&vlc, u[1]);
                            put_uvlc_suffix(&vlc, u[0]); put_uvlc_suffix(&vlc, u[1]);
                            put_uvlc_ext(&vlc, u[0]); put_uvlc_ext(&vlc, u[1]);
                        }
                    }
                } else {
                    put_uvlc_prefix(&vlc, u[0]); put_uvlc_prefix(&vlc, u[1]);
                    put_uvlc_suffix(&vlc, u[0]); put_uvlc_suffix(&vlc, u[1]);
                    put_uvlc_ext(&vlc, u[0]); put_uvlc_ext(&vlc, u[1]);
                }
            } else {
                for (int j = 0; j < npair; j++)
                    if (uoff[j]) {
                        put_uvlc_prefix(&vlc, u[j]);
                        put_uvlc_suffix(&vlc, u[j]);
                        put_uvlc_ext(&vlc, u[j]);
                    }
            }
            /* MagSgn (:394-427) */
            for (int j = 0; j < npair; j++)
                for (int n = 0; n < 4; n++)
                    if (rho[j] >> n & 1) {
                        int ek = chosen[j].emb_k >> n & 1, e1 = chosen[j].emb_1 >> n & 1;
                        int m = U[j] - ek;
                        uint32_t val = v[j][n] - ((uint32_t)e1 << m);
                        if (m > 0) {
                            if (m > 16) {
                                fwd_put(&ms, val & 0xFFFF, 16);
                                fwd_put(&ms, val >> 16, m - 16);
                            } else {
                                fwd_put(&ms, val, m);
                            }
                        }
                    }
            (void)coded;
        }
    }
#undef MAG
#undef SGN

    fwd_finish_magsgn(&ms);
    mel_finish(&mel);
    bwd_finish(&vlc);

    /* assemble Dcup = MagSgn || MEL || reversed VLC || Scup byte */
    Buf *d = &out->data;
    d->n = 0;
    buf_append(d, ms.b.p, ms.b.n);
    buf_append(d, mel.b.p, mel.b.n);
    for (size_t i = vlc.b.n; i-- > 0;)
        buf_put(d, vlc.b.p[i]);
    int Scup = (int)(mel.b.n + vlc.b.n + 1);
    buf_put(d, (uint8_t)(Scup >> 4));
    d->p[d->n - 2] = (d->p[d->n - 2] & 0xF0) | (Scup & 15);
    out->Lcup = (int)d->n;
    out->Lref = 0;
    if (Scup > 4079 || out->Lcup > 65535)
        out->error = 2;

    if (refine) {
        /* SigProp (libavcodec/jpeg2000htdec.c:1034-1131) and MagRef (:1136-1185), emulating the
         * decoder's scan so that the same samples are visited in the same order */
        FwdLsb sp; BwdLsb mr;
        fwd_init(&sp); bwd_init(&mr, 0);
        const int S = w + 2;
        uint8_t *st = calloc((size_t)(h + 2) * S, 1);   /* bit0 sigma, bit1 newly significant */
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                st[(y + 1) * S + x + 1] = (fullmag[y * w + x] >> 1) != 0;
        for (int y0 = 0; y0 < h; y0 += 4) {
            int sh = imin(4, h - y0);
            for (int x0 = 0; x0 < w; x0 += 4) {
                int sw = imin(4, w - x0);
                for (int x = x0; x < x0 + sw; x++)
                    for (int y = y0; y < y0 + sh; y++) {
                        uint8_t *c = &st[(y + 1) * S + x + 1];
                        if (c[0] & 1)
                            continue;
                        int causal = !vsc || y != y0 + sh - 1;
                        int nb = c[-S - 1] | c[-S] | c[-S + 1] | c[-1] | c[1];
                        if (causal)
                            nb |= c[S - 1] | c[S] | c[S + 1];
                        if (nb & 3) {
                            int bit = (fullmag[y * w + x] & 1) != 0;   /* magnitude is exactly 1 here */
                            fwd_put(&sp, bit, 1);
                            if (bit)
                                c[0] |= 2;
                        }
                    }
                for (int x = x0; x < x0 + sw; x++)
                    for (int y = y0; y < y0 + sh; y++)
                        if (st[(y + 1) * S + x + 1] & 2)
                            fwd_put(&sp, sgn[y * w + x], 1);
            }
        }
        fwd_finish_zero(&sp);
        if (npasses > 2) {
            for (int y0 = 0; y0 < h; y0 += 4) {
                int sh = imin(4, h - y0);
                for (int x = 0; x < w; x++)
                    for (int y = y0; y < y0 + sh; y++)
                        if (st[(y + 1) * S + x + 1] & 1)
                            bwd_put(&mr, fullmag[y * w + x] & 1, 1);
            }
            bwd_finish(&mr);
        }
        buf_append(d, sp.b.p, sp.b.n);
        for (size_t i = mr.b.n; i-- > 0;)
            buf_put(d, mr.b.p[i]);
        out->Lref = (int)(sp.b.n + mr.b.n);
        if (out->Lref == 0) {       /* keep the refinement segment non-empty */
            buf_put(d, 0);
            out->Lref = 1;
        }
        if (out->Lref > 65535)
            out->error = 2;
        free(st);
        free(sp.b.p);
        free(mr.b.p);
    }
    free(sig);
    free(Eq);
    free(ms.b.p);
    free(mel.b.p);
    free(vlc.b.p);
}

/* exported for block-level tests: coef = signed integers (w*h), returns total bytes or <0 */
int j2kgen_encode_block(const int32_t *coef, int w, int h, int npasses, int vsc, int maxbp,
                        uint8_t *out, int cap, int *Lcup, int *Lref)
{
    uint32_t *mag = malloc(sizeof(uint32_t) * w * h);
    uint8_t *sg = malloc((size_t)w * h);
    for (int i = 0; i < w * h; i++) {
        mag[i] = coef[i] < 0 ? (uint32_t)(-(int64_t)coef[i]) : (uint32_t)coef[i];
        sg[i] = coef[i] < 0;
    }
    HtBlockOut o;
    memset(&o, 0, sizeof(o));
    ht_encode_block(mag, sg, w, h, npasses, vsc, maxbp, &o);
    int n = (int)o.data.n;
    int ret = o.error ? -o.error : n;
    if (n > cap)
        ret = -9;
    else
        memcpy(out, o.data.p, n);
    *Lcup = o.Lcup;
    *Lref = o.Lref;
    free(o.data.p);
    free(mag);
    free(sg);
    return ret;
}

/* ------------------------------------------------------------------------------------------------ */
/* forward wavelet transforms on a Mallat-layout plane (stride = plane width)                       */

static void fdwt53_1d(int32_t *p, int i0, int i1)
{
    if (i1 <= i0 + 1) {
        if (i0 & 1)
            p[i0] *= 2;
        return;
    }
    /* whole-sample symmetric extension by reflection (period 2(n-1)) */
    int n = i1 - i0;
    for (int k = 1; k <= 2; k++) {
        int a = k % (2 * (n - 1));
        p[i0 - k] = p[i0 + (a < n ? a : 2 * (n - 1) - a)];
        p[i1 - 1 + k] = p[i1 - 1 - (a < n ? a : 2 * (n - 1) - a)];
    }
    for (int i = ((i0 + 1) >> 1) - 1; i < (i1 + 1) >> 1; i++)
        p[2 * i + 1] -= (p[2 * i] + p[2 * i + 2]) >> 1;
    for (int i = (i0 + 1) >> 1; i < (i1 + 1) >> 1; i++)
        p[2 * i] += (p[2 * i - 1] + p[2 * i + 1] + 2) >> 2;
}

#define A97 1.586134342059924f
#define B97 0.052980118572961f
#define G97 0.882911075530934f
#define D97 0.443506852043971f
#define K97 1.230174104914001f
#define X97 0.812893066115961f

static void fdwt97_1d(float *p, int i0, int i1)
{
    if (i1 <= i0 + 1) {
        /* inverse multiplies by K/2 (odd) or X (even): libavcodec/jpeg2000dwt.c:380-386 */
        if (i0 & 1)
            p[i0] *= 2.0f / K97;
        else
            p[i0] *= 1.0f / X97;
        return;
    }
    int n = i1 - i0;
    for (int k = 1; k <= 4; k++) {
        int a = k % (2 * (n - 1));
        int r = a < n ? a : 2 * (n - 1) - a;
        p[i0 - k] = p[i0 + r];
        p[i1 - 1 + k] = p[i1 - 1 - r];
    }
    int lo = i0 - 4, hi = i1 + 4;   /* valid (extended) index range [lo, hi) */
    for (int i = lo + 1; i < hi - 1; i++)
        if (i & 1) p[i] -= A97 * (p[i - 1] + p[i + 1]);
    for (int i = lo + 2; i < hi - 2; i++)
        if (!(i & 1)) p[i] -= B97 * (p[i - 1] + p[i + 1]);
    for (int i = lo + 3; i < hi - 3; i++)
        if (i & 1) p[i] += G97 * (p[i - 1] + p[i + 1]);
    for (int i = lo + 4; i < hi - 4; i++)
        if (!(i & 1)) p[i] += D97 * (p[i - 1] + p[i + 1]);
}

/* one plane, absolute coordinates [x0,x1) x [y0,y1), `levels` decompositions, in place */
static void fdwt_plane(void *plane, int is_float, int x0, int x1, int y0, int y1, int levels)
{
    const int W = x1 - x0;
    int bx0[34], bx1[34], by0[34], by1[34];
    bx0[0] = x0; bx1[0] = x1; by0[0] = y0; by1[0] = y1;
    for (int l = 1; l <= levels; l++) {
        bx0[l] = (bx0[l - 1] + 1) >> 1; bx1[l] = (bx1[l - 1] + 1) >> 1;
        by0[l] = (by0[l - 1] + 1) >> 1; by1[l] = (by1[l - 1] + 1) >> 1;
    }
    int maxlen = imax(x1 - x0, y1 - y0) + 16;
    int32_t *li = malloc(sizeof(int32_t) * maxlen);
    float *lf = (float *)li;
    int32_t *pi = plane;
    float *pf = plane;
    for (int l = 0; l < levels; l++) {
        int lh = bx1[l] - bx0[l], lv = by1[l] - by0[l];
        int mh = bx0[l] & 1, mv = by0[l] & 1;
        /* columns */
        for (int c = 0; c < lh; c++) {
            int j = 0;
            if (is_float) {
                float *q = lf + 6;
                for (int i = 0; i < lv; i++) q[mv + i] = pf[(size_t)W * i + c];
                fdwt97_1d(q, mv, mv + lv);
                for (int i = 0; i < lv; i++) if (((mv + i) & 1) == 0) pf[(size_t)W * j++ + c] = q[mv + i];
                for (int i = 0; i < lv; i++) if (((mv + i) & 1) == 1) pf[(size_t)W * j++ + c] = q[mv + i];
            } else {
                int32_t *q = li + 6;
                for (int i = 0; i < lv; i++) q[mv + i] = pi[(size_t)W * i + c];
                fdwt53_1d(q, mv, mv + lv);
                for (int i = 0; i < lv; i++) if (((mv + i) & 1) == 0) pi[(size_t)W * j++ + c] = q[mv + i];
                for (int i = 0; i < lv; i++) if (((mv + i) & 1) == 1) pi[(size_t)W * j++ + c] = q[mv + i];
            }
        }
        /* rows */
        for (int r = 0; r < lv; r++) {
            int j = 0;
            if (is_float) {
                float *q = lf + 6;
                for (int i = 0; i < lh; i++) q[mh + i] = pf[(size_t)W * r + i];
                fdwt97_1d(q, mh, mh + lh);
                for (int i = 0; i < lh; i++) if (((mh + i) & 1) == 0) pf[(size_t)W * r + j++] = q[mh + i];
                for (int i = 0; i < lh; i++) if (((mh + i) & 1) == 1) pf[(size_t)W * r + j++] = q[mh + i];
            } else {
                int32_t *q = li + 6;
                for (int i = 0; i < lh; i++) q[mh + i] = pi[(size_t)W * r + i];
                fdwt53_1d(q, mh, mh + lh);
                for (int i = 0; i < lh; i++) if (((mh + i) & 1) == 0) pi[(size_t)W * r + j++] = q[mh + i];
                for (int i = 0; i < lh; i++) if (((mh + i) & 1) == 1) pi[(size_t)W * r + j++] = q[mh + i];
            }
        }
    }
    free(li);
}

/* ------------------------------------------------------------------------------------------------ */
/* tier-2: packet headers                                                                           */

typedef struct BitW {
    Buf *b;
    unsigned acc;
    int nbits, limit;
} BitW;

static void bw_init(BitW *w, Buf *b) { w->b = b; w->acc = 0; w->nbits = 0; w->limit = 8; }
static void bw_put(BitW *w, unsigned v, int n)
{
    for (int i = n - 1; i >= 0; i--) {
        w->acc = (w->acc << 1) | ((v >> i) & 1);
        if (++w->nbits == w->limit) {
            buf_put(w->b, (uint8_t)w->acc);
            w->limit = (w->acc == 0xFF) ? 7 : 8;
            w->acc = 0;
            w->nbits = 0;
        }
    }
}
static void bw_flush(BitW *w)
{
    if (w->nbits) {
        w->acc <<= (w->limit - w->nbits);
        buf_put(w->b, (uint8_t)w->acc);
        w->limit = (w->acc == 0xFF) ? 7 : 8;
        w->acc = 0;
        w->nbits = 0;
    }
    if (w->limit == 7)      /* last header byte was 0xFF: a stuffed zero byte follows */
        buf_put(w->b, 0);
    w->limit = 8;
}

typedef struct TagNode {
    int value, low, known, parent;
} TagNode;

typedef struct TagTree {
    TagNode *n;
    int w, h, count;
} TagTree;

static void tt_init(TagTree *t, int w, int h)
{
    int total = 0, ww = w, hh = h;
    for (;;) {
        total += ww * hh;
        if (ww <= 1 && hh <= 1) break;
        ww = (ww + 1) >> 1; hh = (hh + 1) >> 1;
    }
    t->n = calloc(total, sizeof(TagNode));
    t->w = w; t->h = h; t->count = total;
    int base = 0;
    ww = w; hh = h;
    for (;;) {
        int nw = (ww + 1) >> 1, nh = (hh + 1) >> 1;
        int nbase = base + ww * hh;
        for (int y = 0; y < hh; y++)
            for (int x = 0; x < ww; x++)
                t->n[base + y * ww + x].parent = (ww <= 1 && hh <= 1) ? -1 : nbase + (y >> 1) * nw + (x >> 1);
        if (ww <= 1 && hh <= 1) break;
        base = nbase; ww = nw; hh = nh;
    }
    for (int i = 0; i < total; i++)
        t->n[i].value = 1 << 30;
}
static void tt_set(TagTree *t, int leaf, int value)
{
    for (int i = leaf; i >= 0; i = t->n[i].parent) {
        if (t->n[i].value > value)
            t->n[i].value = value;
        else
            break;
    }
}
static void tt_encode(TagTree *t, BitW *w, int leaf, int threshold)
{
    int stack[32], sp = 0;
    for (int i = leaf; i >= 0; i = t->n[i].parent)
        stack[sp++] = i;
    int low = 0;
    while (sp > 0) {
        TagNode *nd = &t->n[stack[--sp]];
        if (low > nd->low) nd->low = low; else low = nd->low;
        while (low < threshold) {
            if (low >= nd->value) {
                if (!nd->known) {
                    bw_put(w, 1, 1);
                    nd->known = 1;
                }
                break;
            }
            bw_put(w, 0, 1);
            low++;
        }
        nd->low = low;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* geometry                                                                                         */

typedef struct CodedBlock {
    uint8_t *data;
    int Lcup, Lref, len;
    int zbp;
    int included;
    /* quality-layer plan and tier-2 state (layers == 2) */
    int mode;                   /* 0 all in layer 0, 1 refinement in layer 1, 2 first included in layer 1, 3 placeholders first */
    int p0_same;                /* placeholder triples signalled together with the HT set */
    int zbp_sig;                /* zero bit-planes as signalled: zbp minus all placeholder triples */
    int lblock, sent;           /* Lblock state; 1 once the block has appeared in a packet header */
} CodedBlock;

typedef struct Precinct {
    TagTree incl, zbpt;         /* tag trees live as long as the precinct: their state spans the layers */
    int tt_ready;
    int ncbw, ncbh;             /* code-block grid */
    int cx0, cy0;               /* index of first code-block column/row in the band grid */
    CodedBlock *cb;
} Precinct;

typedef struct Band {
    int x0, x1, y0, y1;         /* band coordinates */
    int offx, offy;             /* Mallat offset inside the tile-component plane */
    int cbw, cbh;               /* log2 code-block size actually used */
    int expn, mant;
    float step;                 /* decoder's f_stepsize incl. folded gains (irreversible) */
    Precinct *prec;
} Band;

typedef struct ResLevel {
    int x0, x1, y0, y1;
    int ppx, ppy, npx, npy;
    int nbands;
    Band band[3];
    int *written;               /* per precinct: layers already sequenced (POC entries may overlap) */
} ResLevel;

/* one progression over resolutions [RS, RE), components [CS, CE) and layers [0, LYE): the packet order of
 * T.800 B.12 for that sub-volume.  A packet that an earlier progression already carried is not repeated. */
/* where write_packet puts packet headers when they are packed (PPT / PPM); NULL = in front of the body */
static __thread Buf *packed_hdr_target;
/* tile-part cuts: body / packed-header sizes after every packet (filled by sequence_one when cut_pos != NULL) */
static __thread size_t *cut_pos;
static __thread int cut_n, cut_cap;

static int write_packet(Buf *out, const J2kGenParams *P, ResLevel *rl, int precno, int layer, int *sop_counter);
static void sequence_one(Buf *out, const J2kGenParams *P, ResLevel *rl, int precno, int layer, int *sopc)
{
    if (!rl->written)
        rl->written = calloc((size_t)imax(rl->npx * rl->npy, 1), sizeof(int));
    if (layer < rl->written[precno])
        return;
    rl->written[precno] = layer + 1;
    write_packet(out, P, rl, precno, layer, sopc);
    if (cut_pos && cut_n < cut_cap) {
        cut_pos[2 * cut_n] = out->n;
        cut_pos[2 * cut_n + 1] = packed_hdr_target ? packed_hdr_target->n : 0;
        cut_n++;
    }
}

static void sequence_packets(Buf *out, const J2kGenParams *P, ResLevel **res, int C, int L, int tx0, int ty0,
                             int tx1, int ty1, int RS, int CS, int LYE, int RE, int CE, int prog, int *sopc)
{
    RE = imin(RE, L + 1);
    switch (prog) {
    case 0: /* LRCP */
        for (int l = 0; l < LYE; l++)
            for (int r = RS; r < RE; r++)
                for (int c = CS; c < CE; c++)
                    for (int p = 0; p < res[c][r].npx * res[c][r].npy; p++)
                        sequence_one(out, P, &res[c][r], p, l, sopc);
        break;
    case 1: /* RLCP */
        for (int r = RS; r < RE; r++)
            for (int l = 0; l < LYE; l++)
                for (int c = CS; c < CE; c++)
                    for (int p = 0; p < res[c][r].npx * res[c][r].npy; p++)
                        sequence_one(out, P, &res[c][r], p, l, sopc);
        break;
    case 2: /* RPCL (components share one sampling grid here) */
        for (int r = RS; r < RE; r++)
            for (int p = 0; p < res[0][r].npx * res[0][r].npy; p++)
                for (int c = CS; c < CE; c++)
                    for (int l = 0; l < LYE; l++)
                        sequence_one(out, P, &res[c][r], p, l, sopc);
        break;
    case 3: /* PCRL: positions ordered by their top-left corner on the reference grid */
    case 4: /* CPRL */
        for (int cc = (prog == 4 ? CS : 0); cc < (prog == 4 ? CE : 1); cc++) {
            /* walk reference-grid positions at the finest precinct granularity */
            int minx = 30, miny = 30;
            for (int r = 0; r <= L; r++) {
                minx = imin(minx, res[0][r].ppx + L - r);
                miny = imin(miny, res[0][r].ppy + L - r);
            }
            int sx = 1 << imin(minx, 30), sy = 1 << imin(miny, 30);
            for (int y = ty0; y < ty1; y = (y / sy + 1) * sy)
                for (int x = tx0; x < tx1; x = (x / sx + 1) * sx)
                    for (int c = (prog == 4 ? cc : CS); c < (prog == 4 ? cc + 1 : CE); c++)
                        for (int r = RS; r < RE; r++) {
                            ResLevel *rl = &res[c][r];
                            int nl = L - r;
                            if (!rl->npx || !rl->npy) continue;
                            int okx = (x % (1 << (rl->ppx + nl)) == 0) || (x == tx0 && ((rl->x0 << nl) % (1 << (rl->ppx + nl))));
                            int oky = (y % (1 << (rl->ppy + nl)) == 0) || (y == ty0 && ((rl->y0 << nl) % (1 << (rl->ppy + nl))));
                            if (!okx || !oky) continue;
                            int pxi = (ceildivpow2(x, nl) >> rl->ppx) - (rl->x0 >> rl->ppx);
                            int pyi = (ceildivpow2(y, nl) >> rl->ppy) - (rl->y0 >> rl->ppy);
                            if (pxi < 0 || pyi < 0 || pxi >= rl->npx || pyi >= rl->npy) continue;
                            for (int l = 0; l < LYE; l++)
                                sequence_one(out, P, rl, pxi + pyi * rl->npx, l, sopc);
                        }
        }
        break;
    }
}

/* the decoder's step size for a band: libavcodec/jpeg2000.c:214-272 (float path), same op order */
static float decoder_stepsize(int cbps, int expn, int mant, int bandpos, int nres, int reslevelno)
{
    union { uint32_t u; float f; } e;
    e.u = (uint32_t)((cbps - expn) + 127) << 23;
    float f = e.f;
    f *= mant / 2048.0 + 1.0;
    int lband = 0;
    switch (bandpos) {
    case 1: case 2: f *= X97 * 2; lband = 1; break;
    case 3: f *= X97 * X97 * 4; break;
    }
    f *= pow(K97, 2 * (nres - reslevelno) + lband - 2);
    return f;
}

static void put_npasses(BitW *w, int n)
{
    /* jpeg2000dec.c:1073-1086 */
    if (n == 1) bw_put(w, 0, 1);
    else if (n == 2) bw_put(w, 2, 2);
    else if (n <= 5) bw_put(w, 0xC | (n - 3), 4);
    else { bw_put(w, 0xF, 4); bw_put(w, n - 6, 5); }
}

static int floor_log2(int v) { int n = 0; while (v > 1) { v >>= 1; n++; } return n; }

/* what a block contributes to the packet of `layer`: 0 nothing, 1 placeholder triple, 2 HT set (cleanup and,
 * unless they follow in the next layer, the refinement passes), 3 refinement passes only */
static int contribution(const J2kGenParams *P, const CodedBlock *cb, int layer)
{
    if (!cb->included)
        return 0;
    switch (cb->mode) {
    case 0: return layer == 0 ? 2 : 0;
    case 1: return layer == 0 ? 2 : (P->npasses > 1 ? 3 : 0);
    case 2: return layer == 1 ? 2 : 0;
    default: return layer == 0 ? 1 : 2;
    }
}


static int write_packet(Buf *out, const J2kGenParams *P, ResLevel *rl, int precno, int layer, int *sop_counter)
{
    Buf *hb = packed_hdr_target ? packed_hdr_target : out;
    if (P->sop) {
        buf_put16(out, 0xFF91); buf_put16(out, 4); buf_put16(out, (*sop_counter)++ & 0xFFFF);
    }
    int any = 0;
    for (int b = 0; b < rl->nbands; b++) {
        Band *bd = &rl->band[b];
        if (bd->x0 == bd->x1 || bd->y0 == bd->y1) continue;
        Precinct *pr = &bd->prec[precno];
        for (int i = 0; i < pr->ncbw * pr->ncbh; i++)
            any |= contribution(P, &pr->cb[i], layer) != 0;
    }
    BitW w;
    bw_init(&w, hb);
    if (!any) {
        bw_put(&w, 0, 1);
        bw_flush(&w);
        if (P->eph) buf_put16(hb, 0xFF92);
        return 0;
    }
    bw_put(&w, 1, 1);
    for (int b = 0; b < rl->nbands; b++) {
        Band *bd = &rl->band[b];
        if (bd->x0 == bd->x1 || bd->y0 == bd->y1) continue;
        Precinct *pr = &bd->prec[precno];
        int n = pr->ncbw * pr->ncbh;
        if (!n) continue;
        if (!pr->tt_ready) {
            tt_init(&pr->incl, pr->ncbw, pr->ncbh);
            tt_init(&pr->zbpt, pr->ncbw, pr->ncbh);
            for (int i = 0; i < n; i++) {
                CodedBlock *cb = &pr->cb[i];
                /* leaf of the inclusion tree = first layer with a contribution; blocks without data never appear */
                tt_set(&pr->incl, i, cb->included ? (cb->mode == 2 ? 1 : 0) : 1 << 20);
                if (cb->included)
                    tt_set(&pr->zbpt, i, cb->zbp_sig);
            }
            pr->tt_ready = 1;
        }
        for (int i = 0; i < n; i++) {
            CodedBlock *cb = &pr->cb[i];
            const int what = contribution(P, cb, layer);
            if (!cb->sent) {
                /* the decoder asks the tree twice from the second layer on (jpeg2000dec.c:1163-1170) */
                if (layer > 0)
                    tt_encode(&pr->incl, &w, i, 1);
                tt_encode(&pr->incl, &w, i, layer + 1);
                if (!what)
                    continue;
                tt_encode(&pr->zbpt, &w, i, cb->zbp_sig + 1);   /* emits up to and including the terminating 1 */
                cb->sent = 1;
                cb->lblock = 3;
            } else {
                bw_put(&w, what != 0, 1);
                if (!what)
                    continue;
            }
            const int np = P->npasses;
            if (what == 1) {
                /* three empty placeholder passes: a zero length in Lblock bits, then the one more bit the decoder
                 * reads while it widens the field for three passes (jpeg2000dec.c:1301-1321) */
                put_npasses(&w, 3);
                bw_put(&w, 0, 1);
                bw_put(&w, 0, cb->lblock);
                bw_put(&w, 0, 1);
            } else if (what == 2) {
                const int npc = (cb->mode == 1 && np > 1) ? 1 : np;          /* passes of the HT set sent now */
                const int newpasses = 3 * cb->p0_same + npc;
                const int wide = floor_log2(3 * cb->p0_same + 1);            /* extra bits of the cleanup length field */
                put_npasses(&w, newpasses);
                int need = bitlen(cb->Lcup) - wide;
                if (npc > 1)
                    need = imax(need, bitlen(cb->Lref) - (npc == 3));
                if (P->ht_set == 3)       /* MIXED: an HT cleanup length needs Lblock > 3 and a 0 on top (:1293-1298) */
                    need = imax(imax(need, bitlen(cb->Lcup) - wide + 1), 4);
                while (cb->lblock < need) { bw_put(&w, 1, 1); cb->lblock++; }
                bw_put(&w, 0, 1);
                bw_put(&w, cb->Lcup, cb->lblock + wide);
                if (npc > 1)
                    bw_put(&w, cb->Lref, cb->lblock + (npc == 3));
            } else {
                const int newpasses = np - 1;                                /* SigProp [+ MagRef] */
                put_npasses(&w, newpasses);
                const int need = bitlen(cb->Lref) - (newpasses == 2);
                while (cb->lblock < need) { bw_put(&w, 1, 1); cb->lblock++; }
                bw_put(&w, 0, 1);
                bw_put(&w, cb->Lref, cb->lblock + (newpasses == 2));
            }
        }
    }
    bw_flush(&w);
    if (P->eph) buf_put16(hb, 0xFF92);
    for (int b = 0; b < rl->nbands; b++) {
        Band *bd = &rl->band[b];
        if (bd->x0 == bd->x1 || bd->y0 == bd->y1) continue;
        Precinct *pr = &bd->prec[precno];
        for (int i = 0; i < pr->ncbw * pr->ncbh; i++) {
            CodedBlock *cb = &pr->cb[i];
            const int what = contribution(P, cb, layer);
            if (what == 2) {
                const int all = !(cb->mode == 1 && P->npasses > 1);
                buf_append(out, cb->data, all ? cb->len : cb->Lcup);
            } else if (what == 3) {
                buf_append(out, cb->data + cb->Lcup, cb->Lref);
            }
        }
    }
    return 0;
}

static void quant_exps(const J2kGenParams *P, int compno, int r, int bandpos, int *expn, int *mant);

/* COD and QCD marker segments from the parameters: main header, or a tile-part header when a tile differs */
static void put_cod(Buf *out, const J2kGenParams *P, int C)
{
    const int L = P->levels;
    buf_put16(out, 0xFF52);
    buf_put16(out, 12 + (P->use_precincts ? L + 1 : 0));
    buf_put(out, (P->use_precincts ? 1 : 0) | (P->sop ? 2 : 0) | (P->eph ? 4 : 0));
    buf_put(out, P->prog); buf_put16(out, P->layers == 2 ? 2 : 1); buf_put(out, (P->mct && C >= 3) ? 1 : 0);
    buf_put(out, L); buf_put(out, P->cbw_log2 - 2); buf_put(out, P->cbh_log2 - 2);
    buf_put(out, (P->ht_set >= 3 ? 0xC0 : 0x40) | (P->vsc ? 0x08 : 0));
    buf_put(out, P->reversible ? 1 : 0);
    if (P->use_precincts)
        for (int r = 0; r <= L; r++) buf_put(out, P->prec_exps[r]);
}

static void put_qcd(Buf *out, const J2kGenParams *P)
{
    const int nb = 3 * P->levels + 1;
    buf_put16(out, 0xFF5C);
    if (P->reversible) {
        buf_put16(out, 3 + nb); buf_put(out, (P->guard_bits << 5) | 0);
        for (int b = 0; b < nb; b++) {
            int r = b == 0 ? 0 : (b - 1) / 3 + 1, bp = b == 0 ? 0 : (b - 1) % 3 + 1, e, m;
            quant_exps(P, 0, r, bp, &e, &m);
            buf_put(out, e << 3);
        }
    } else {
        buf_put16(out, 3 + 2 * nb); buf_put(out, (P->guard_bits << 5) | 2);
        for (int b = 0; b < nb; b++) {
            int r = b == 0 ? 0 : (b - 1) / 3 + 1, bp = b == 0 ? 0 : (b - 1) % 3 + 1, e, m;
            quant_exps(P, 0, r, bp, &e, &m);
            buf_put16(out, (e << 11) | m);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */

static void quant_exps(const J2kGenParams *P, int compno, int r, int bandpos, int *expn, int *mant)
{
    if (P->reversible) {
        /* epsilon_b = R_I + log2(gain_b) (+1 for the RCT chroma range) */
        int gain = bandpos == 0 ? 0 : (bandpos == 3 ? 2 : 1);
        *expn = P->bitdepth + gain + ((P->mct && P->ncomp >= 3) ? 1 : 0);
        *mant = 0;
        (void)compno; (void)r;
    } else {
        /* Delta = 2^(cbps - expn) * (1 + mant/2048)  (libavcodec/jpeg2000.c:243-244) */
        double d = P->base_step;
        int e = (int)floor(log2(d));
        int m = (int)floor((d / ldexp(1.0, e) - 1.0) * 2048.0 + 0.5);
        if (m >= 2048) { m = 0; e++; }
        *expn = P->bitdepth - e;
        *mant = m;
        if (*expn < 0) *expn = 0;
        if (*expn > 31) *expn = 31;
    }
}

int64_t j2kgen_encode(const int32_t *const *planes, J2kGenParams *P, uint8_t *outbuf, int64_t cap)
{
    Buf out = { 0 };
    const int C = P->ncomp, L = P->levels;
    const int X0 = P->x_off, Y0 = P->y_off;
    const int X1 = X0 + P->width, Y1 = Y0 + P->height;
    const int TW = P->tile_w > 0 ? P->tile_w : P->width + X0;   /* tile grid origin = (X0,Y0) */
    const int TH = P->tile_h > 0 ? P->tile_h : P->height + Y0;
    const int ntx = ceildiv(X1 - X0, TW), nty = ceildiv(Y1 - Y0, TH);
    P->error = 0; P->stats_blocks = 0; P->stats_max_lcup = 0;
    build_enc_tables();

    /* ---- main header ---- */
    buf_put16(&out, 0xFF4F);
    buf_put16(&out, 0xFF51); buf_put16(&out, 38 + 3 * C); buf_put16(&out, 0x4000);
    buf_put32(&out, X1); buf_put32(&out, Y1); buf_put32(&out, X0); buf_put32(&out, Y0);
    buf_put32(&out, TW); buf_put32(&out, TH); buf_put32(&out, X0); buf_put32(&out, Y0);
    buf_put16(&out, C);
    int any_sub = 0;
    for (int c = 0; c < C && c < 4; c++) any_sub |= P->sub_x[c] > 1 || P->sub_y[c] > 1;
    if (any_sub && (X0 || Y0 || (P->mct && C >= 3) || P->prog > 1)) { P->error = 3; return -3; }
#define SUBX(c) ((c) < 4 && P->sub_x[c] > 1 ? P->sub_x[c] : 1)
#define SUBY(c) ((c) < 4 && P->sub_y[c] > 1 ? P->sub_y[c] : 1)
    for (int c = 0; c < C; c++) { buf_put(&out, P->bitdepth - 1); buf_put(&out, SUBX(c)); buf_put(&out, SUBY(c)); }
    /* CAP: Pcap bit 15 (Part 15), Ccap15: HTONLY, bit 5 = irreversible */
    {
        int magb_bits = 0;   /* the reference only range-checks MAGB (jpeg2000dec.c:473-486) */
        buf_put16(&out, 0xFF50); buf_put16(&out, 8); buf_put32(&out, 0x00020000);
        buf_put16(&out, (P->reversible ? 0 : 0x0020) | ((P->packed_headers == 1 || P->tile_var) ? 0x0800 : 0) | magb_bits |
                        (P->ht_set == 1 ? 0x4000 : P->ht_set == 3 || P->ht_set == 4 ? 0xC000 : 0));
    }
    const int NLAY = P->layers == 2 ? 2 : 1;
    put_cod(&out, P, C);
    put_qcd(&out, P);
    if (P->npoc > 0) {
        if (P->npoc > 8) { P->error = 4; return -4; }
        int first = P->poc_split > 0 && P->poc_split < P->npoc ? P->poc_split : P->npoc;
        for (int a = 0; a < P->npoc; ) {
            int n = a == 0 ? first : P->npoc - a;
            buf_put16(&out, 0xFF5F); buf_put16(&out, 2 + 7 * n);
            for (int i = a; i < a + n; i++) {
                buf_put(&out, P->poc[i][0]); buf_put(&out, P->poc[i][1]); buf_put16(&out, P->poc[i][2]);
                buf_put(&out, P->poc[i][3]); buf_put(&out, P->poc[i][4]); buf_put(&out, P->poc[i][5]);
            }
            a += n;
        }
    }

    const size_t main_end = out.n;       /* PPM marker segments are inserted here once the tiles are written */
    Buf ppm = { 0 };
    /* every tile-part is assembled on its own, so that the parts of different tiles can be interleaved */
    const int ntiles = ntx * nty;
    Buf *tparts = calloc((size_t)ntiles * 32, sizeof(Buf)), *tphdr = calloc((size_t)ntiles * 32, sizeof(Buf));
    int *tnparts = calloc((size_t)ntiles, sizeof(int));

    /* ---- tiles ---- */
    for (int ty = 0; ty < nty; ty++)
        for (int tx = 0; tx < ntx; tx++) {
            const int tileno = ty * ntx + tx;
            /* the parameters this tile is coded with: the main header's, or a variation announced in its header */
            J2kGenParams *const Pmain = P;
            J2kGenParams Pt = *Pmain;
            const int tile_differs = Pmain->tile_var && (tileno & 1);
            if (tile_differs) {
                if (Pmain->tile_var & 1) Pt.prog = (Pmain->prog + 2) % 5;
                if (Pmain->tile_var & 2) Pt.levels = imax(1, Pmain->levels - 1);
                if (Pmain->tile_var & 4) Pt.cbw_log2 = imax(2, Pmain->cbw_log2 - 1);
                if (Pmain->tile_var & 8) {
                    if (Pmain->reversible) Pt.guard_bits = imin(7, Pmain->guard_bits + 1);
                    else Pt.base_step = Pmain->base_step * 2.0f;
                }
                if ((Pmain->tile_var & 16) && !Pmain->npoc) {
                    const int nl = Pmain->layers == 2 ? 2 : 1;
                    Pt.npoc = 2;
                    Pt.poc[0][0] = 0; Pt.poc[0][1] = 0; Pt.poc[0][2] = nl; Pt.poc[0][3] = imax(1, (Pt.levels + 1) / 2);
                    Pt.poc[0][4] = C; Pt.poc[0][5] = 1;
                    Pt.poc[1][0] = 0; Pt.poc[1][1] = 0; Pt.poc[1][2] = nl; Pt.poc[1][3] = Pt.levels + 1;
                    Pt.poc[1][4] = C; Pt.poc[1][5] = 0;
                }
            }
            {
            J2kGenParams *P = &Pt;
            const int L = Pt.levels;
            const int tx0 = imax(X0 + tx * TW, X0), tx1 = imin(X0 + (tx + 1) * TW, X1);
            const int ty0 = imax(Y0 + ty * TH, Y0), ty1 = imin(Y0 + (ty + 1) * TH, Y1);
            const int W = tx1 - tx0, H = ty1 - ty0;
            /* component planes of the tile, level shifted and colour transformed */
            void **tp = calloc(C, sizeof(void *));
            const int shift = 1 << (P->bitdepth - 1);
            const int use_mct = P->mct && C >= 3;
            if (any_sub) {
                /* every component on its own grid: tile-component = ceil(tile / (XRsiz, YRsiz)) */
                for (int c = 0; c < C; c++) {
                    const int dx = SUBX(c), dy = SUBY(c);
                    const int cx0 = ceildiv(tx0, dx), cx1 = ceildiv(tx1, dx), cy0 = ceildiv(ty0, dy), cy1 = ceildiv(ty1, dy);
                    const int cw = cx1 - cx0, ch = cy1 - cy0, pw = ceildiv(P->width, dx);
                    tp[c] = malloc(sizeof(int32_t) * (size_t)imax(cw * ch, 1) + 64);
                    for (int y = 0; y < ch; y++)
                        for (int x = 0; x < cw; x++) {
                            const int32_t v = planes[c][(size_t)(cy0 + y) * pw + (cx0 + x)] - shift;
                            if (P->reversible) ((int32_t *)tp[c])[(size_t)y * cw + x] = v;
                            else ((float *)tp[c])[(size_t)y * cw + x] = (float)v;
                        }
                }
            } else {
            for (int c = 0; c < C; c++) tp[c] = malloc(sizeof(int32_t) * (size_t)W * H + 64);
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    size_t src = (size_t)(ty0 - Y0 + y) * P->width + (tx0 - X0 + x), dst = (size_t)y * W + x;
                    if (P->reversible) {
                        int32_t v[4];
                        for (int c = 0; c < C && c < 4; c++) v[c] = planes[c][src] - shift;
                        if (use_mct) {
                            int32_t r = v[0], g = v[1], b = v[2];
                            v[0] = (r + 2 * g + b) >> 2; v[1] = b - g; v[2] = r - g;
                        }
                        for (int c = 0; c < C && c < 4; c++) ((int32_t *)tp[c])[dst] = v[c];
                    } else {
                        float v[4];
                        for (int c = 0; c < C && c < 4; c++) v[c] = (float)(planes[c][src] - shift);
                        if (use_mct) {
                            float r = v[0], g = v[1], b = v[2];
                            v[0] = 0.299f * r + 0.587f * g + 0.114f * b;
                            v[1] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                            v[2] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                        }
                        for (int c = 0; c < C && c < 4; c++) ((float *)tp[c])[dst] = v[c];
                    }
                }
            }
            /* per component: DWT, geometry, block coding */
            ResLevel **res = calloc(C, sizeof(ResLevel *));
            const int tile_x0 = tx0, tile_x1 = tx1, tile_y0 = ty0, tile_y1 = ty1;
            for (int c = 0; c < C; c++) {
                /* the component's own rectangle shadows the tile's from here to the end of this loop body */
                const int tx0 = ceildiv(tile_x0, SUBX(c)), tx1 = ceildiv(tile_x1, SUBX(c));
                const int ty0 = ceildiv(tile_y0, SUBY(c)), ty1 = ceildiv(tile_y1, SUBY(c));
                const int W = tx1 - tx0;
                fdwt_plane(tp[c], !P->reversible, tx0, tx1, ty0, ty1, L);
                res[c] = calloc(L + 1, sizeof(ResLevel));
                for (int r = 0; r <= L; r++) {
                    ResLevel *rl = &res[c][r];
                    int nl = L - r;
                    rl->x0 = ceildivpow2(tx0, nl); rl->x1 = ceildivpow2(tx1, nl);
                    rl->y0 = ceildivpow2(ty0, nl); rl->y1 = ceildivpow2(ty1, nl);
                    rl->ppx = P->use_precincts ? (P->prec_exps[r] & 15) : 15;
                    rl->ppy = P->use_precincts ? (P->prec_exps[r] >> 4) : 15;
                    rl->npx = rl->x1 == rl->x0 ? 0 : ceildivpow2(rl->x1, rl->ppx) - (rl->x0 >> rl->ppx);
                    rl->npy = rl->y1 == rl->y0 ? 0 : ceildivpow2(rl->y1, rl->ppy) - (rl->y0 >> rl->ppy);
                    rl->nbands = r == 0 ? 1 : 3;
                    for (int b = 0; b < rl->nbands; b++) {
                        Band *bd = &rl->band[b];
                        int bandpos = b + (r > 0);
                        int xob = bandpos & 1, yob = bandpos >> 1;
                        if (r == 0) {
                            bd->x0 = rl->x0; bd->x1 = rl->x1; bd->y0 = rl->y0; bd->y1 = rl->y1;
                            bd->offx = bd->offy = 0;
                        } else {
                            int nb = L - r + 1;
                            bd->x0 = ceildivpow2(tx0 - (xob << (nb - 1)), nb); bd->x1 = ceildivpow2(tx1 - (xob << (nb - 1)), nb);
                            bd->y0 = ceildivpow2(ty0 - (yob << (nb - 1)), nb); bd->y1 = ceildivpow2(ty1 - (yob << (nb - 1)), nb);
                            bd->offx = xob ? res[c][r - 1].x1 - res[c][r - 1].x0 : 0;
                            bd->offy = yob ? res[c][r - 1].y1 - res[c][r - 1].y0 : 0;
                        }
                        int bppx = rl->ppx - (r > 0), bppy = rl->ppy - (r > 0);
                        bd->cbw = imin(P->cbw_log2, bppx);
                        bd->cbh = imin(P->cbh_log2, bppy);
                        quant_exps(P, c, r, bandpos, &bd->expn, &bd->mant);
                        bd->step = P->reversible ? 1.0f : decoder_stepsize(P->bitdepth, bd->expn, bd->mant, bandpos, L + 1, r);
                        const int M_b = bd->expn + P->guard_bits - 1;
                        const int refine = P->npasses > 1;
                        const int zbp = M_b - 1 - refine;
                        const int maxbp = zbp + 1;
                        bd->prec = calloc((size_t)imax(1, rl->npx * rl->npy), sizeof(Precinct));
                        if (bd->x0 == bd->x1 || bd->y0 == bd->y1)
                            continue;
                        for (int pno = 0; pno < rl->npx * rl->npy; pno++) {
                            Precinct *pr = &bd->prec[pno];
                            int px0 = ((rl->x0 >> rl->ppx) + pno % rl->npx) << bppx, px1 = px0 + (1 << bppx);
                            int py0 = ((rl->y0 >> rl->ppy) + pno / rl->npx) << bppy, py1 = py0 + (1 << bppy);
                            px0 = imax(px0, bd->x0); px1 = imin(px1, bd->x1);
                            py0 = imax(py0, bd->y0); py1 = imin(py1, bd->y1);
                            if (px1 <= px0 || py1 <= py0) { pr->ncbw = pr->ncbh = 0; continue; }
                            pr->ncbw = ceildivpow2(px1, bd->cbw) - (px0 >> bd->cbw);
                            pr->ncbh = ceildivpow2(py1, bd->cbh) - (py0 >> bd->cbh);
                            pr->cb = calloc((size_t)pr->ncbw * pr->ncbh, sizeof(CodedBlock));
                            for (int k = 0; k < pr->ncbw * pr->ncbh; k++) {
                                int gx = ((px0 >> bd->cbw) + k % pr->ncbw) << bd->cbw;
                                int gy = ((py0 >> bd->cbh) + k / pr->ncbw) << bd->cbh;
                                int cx0 = imax(gx, px0), cx1 = imin(gx + (1 << bd->cbw), px1);
                                int cy0 = imax(gy, py0), cy1 = imin(gy + (1 << bd->cbh), py1);
                                int bw = cx1 - cx0, bh = cy1 - cy0;
                                CodedBlock *cb = &pr->cb[k];
                                uint32_t *mag = malloc(sizeof(uint32_t) * bw * bh);
                                uint8_t *sg = malloc((size_t)bw * bh);
                                int nz = 0;
                                for (int y = 0; y < bh; y++)
                                    for (int x = 0; x < bw; x++) {
                                        size_t pidx = (size_t)(cy0 - bd->y0 + bd->offy + y) * W + (cx0 - bd->x0 + bd->offx + x);
                                        int64_t q;
                                        if (P->reversible) {
                                            q = ((int32_t *)tp[c])[pidx];
                                        } else {
                                            float f = ((float *)tp[c])[pidx] / bd->step;
                                            q = (int64_t)f;     /* dead-zone quantiser: truncate toward zero */
                                        }
                                        uint32_t m = (uint32_t)(q < 0 ? -q : q);
                                        mag[y * bw + x] = m;
                                        sg[y * bw + x] = q < 0;
                                        nz |= (m >> refine) != 0;
                                    }
                                if (nz) {
                                    HtBlockOut o;
                                    memset(&o, 0, sizeof(o));
                                    ht_encode_block(mag, sg, bw, bh, P->npasses, P->vsc, maxbp, &o);
                                    if (o.error) P->error = o.error;
                                    cb->data = o.data.p; cb->len = (int)o.data.n;
                                    cb->Lcup = o.Lcup; cb->Lref = o.Lref; cb->zbp = zbp; cb->included = 1;
                                    cb->zbp_sig = zbp;
                                    if (P->layers == 2) {
                                        uint32_t hsh = (uint32_t)(((((tileno * 131 + c) * 31 + r) * 17 + b) * 7 + pno) * 3 + k);
                                        hsh = hsh * 2654435761u;
                                        hsh ^= hsh >> 15;
                                        cb->mode = (int)(hsh & 3);
                                        cb->p0_same = (int)((hsh >> 2) & 1);
                                        int p0 = cb->p0_same + (cb->mode == 3);
                                        if (zbp < p0) { cb->mode = 0; cb->p0_same = 0; p0 = 0; }
                                        cb->zbp_sig = zbp - p0;
                                    }
                                    P->stats_blocks++;
                                    if (o.Lcup > P->stats_max_lcup) P->stats_max_lcup = o.Lcup;
                                }
                                free(mag); free(sg);
                            }
                        }
                    }
                }
            }
            /* tile-part(s): SOT, [PPT], SOD, packets */
            Buf body = { 0 }, phdr = { 0 };
            packed_hdr_target = P->packed_headers ? &phdr : NULL;
            int sopc = 0;
            int maxnp = 0, total_packets = 0;
            for (int c = 0; c < C; c++) for (int r = 0; r <= L; r++) {
                maxnp = imax(maxnp, res[c][r].npx * res[c][r].npy);
                total_packets += res[c][r].npx * res[c][r].npy * NLAY;
            }
            cut_cap = total_packets + 1;
            cut_pos = P->tp_packets > 0 ? calloc((size_t)cut_cap * 2, sizeof(size_t)) : NULL;
            cut_n = 0;
            if (P->npoc > 0) {
                for (int i = 0; i < P->npoc; i++)
                    sequence_packets(&body, P, res, C, L, tx0, ty0, tx1, ty1, P->poc[i][0], P->poc[i][1],
                                     imin(P->poc[i][2], NLAY), P->poc[i][3], P->poc[i][4] ? imin(P->poc[i][4], C) : C,
                                     P->poc[i][5], &sopc);
            } else {
                sequence_packets(&body, P, res, C, L, tx0, ty0, tx1, ty1, 0, 0, NLAY, L + 1, C, P->prog, &sopc);
            }
            packed_hdr_target = NULL;
            /* part k holds packets [k * per, (k + 1) * per); the last one takes the rest */
            int nparts = 1, per = cut_n;
            if (cut_pos && cut_n > 0) {
                per = imax(P->tp_packets, (cut_n + 31) / 32);
                nparts = (cut_n + per - 1) / per;
            }
            for (int k = 0; k < nparts; k++) {
                size_t b0 = 0, b1 = body.n, h0 = 0, h1 = phdr.n;
                if (nparts > 1) {
                    if (k > 0) { b0 = cut_pos[2 * (k * per - 1)]; h0 = cut_pos[2 * (k * per - 1) + 1]; }
                    if (k + 1 < nparts) { b1 = cut_pos[2 * ((k + 1) * per - 1)]; h1 = cut_pos[2 * ((k + 1) * per - 1) + 1]; }
                }
                Buf *tp_out = &tparts[tileno * 32 + k];
                buf_put16(tp_out, 0xFF90); buf_put16(tp_out, 10); buf_put16(tp_out, tileno); buf_put32(tp_out, 0);
                buf_put(tp_out, k); buf_put(tp_out, nparts);
                if (tile_differs && k == 0) {             /* this tile's own coding / quantisation / progression */
                    if (Pmain->tile_var & 7) put_cod(tp_out, P, C);
                    if (Pmain->tile_var & (2 | 8)) put_qcd(tp_out, P);
                    if ((Pmain->tile_var & 16) && !Pmain->npoc) {
                        buf_put16(tp_out, 0xFF5F); buf_put16(tp_out, 2 + 7 * P->npoc);
                        for (int i = 0; i < P->npoc; i++) {
                            buf_put(tp_out, P->poc[i][0]); buf_put(tp_out, P->poc[i][1]); buf_put16(tp_out, P->poc[i][2]);
                            buf_put(tp_out, P->poc[i][3]); buf_put(tp_out, P->poc[i][4]); buf_put(tp_out, P->poc[i][5]);
                        }
                    }
                }
                if (P->packed_headers == 1 && k == 0) {
                    const size_t chunk = P->ppx_chunk > 0 ? (size_t)imin(P->ppx_chunk, 65532) : 65532;
                    int z = 0;
                    for (size_t at = 0; at < phdr.n || z == 0; at += chunk, z++) {
                        const size_t n = phdr.n - at < chunk ? phdr.n - at : chunk;
                        buf_put16(tp_out, 0xFF61); buf_put16(tp_out, (unsigned)(3 + n)); buf_put(tp_out, z & 255);
                        buf_append(tp_out, phdr.p + at, n);
                    }
                } else if (P->packed_headers == 2) {
                    buf_append(&tphdr[tileno * 32 + k], phdr.p + h0, h1 - h0);
                }
                buf_put16(tp_out, 0xFF93);
                buf_append(tp_out, body.p + b0, b1 - b0);
                const uint32_t psot = (uint32_t)tp_out->n;
                tp_out->p[6] = psot >> 24; tp_out->p[7] = psot >> 16; tp_out->p[8] = psot >> 8; tp_out->p[9] = psot;
            }
            tnparts[tileno] = nparts;
            free(cut_pos); cut_pos = NULL;
            free(body.p); free(phdr.p);
            /* free */
            for (int c = 0; c < C; c++) {
                for (int r = 0; r <= L; r++)
                    for (int b = 0; b < res[c][r].nbands; b++) {
                        Band *bd = &res[c][r].band[b];
                        if (!bd->prec) continue;
                        for (int p = 0; p < res[c][r].npx * res[c][r].npy; p++) {
                            Precinct *pr = &bd->prec[p];
                            for (int k = 0; k < pr->ncbw * pr->ncbh; k++) free(pr->cb[k].data);
                            free(pr->cb);
                            if (pr->tt_ready) { free(pr->incl.n); free(pr->zbpt.n); }
                        }
                        free(bd->prec);
                    }
                for (int r = 0; r <= L; r++) free(res[c][r].written);
                free(res[c]);
                free(tp[c]);
            }
            free(res);
            free(tp);
            }
            Pmain->stats_blocks = Pt.stats_blocks;
            Pmain->stats_max_lcup = Pt.stats_max_lcup;
            if (Pt.error) Pmain->error = Pt.error;
        }
    /* emit the tile-parts: tile after tile, or part k of every tile before part k + 1 of any */
    for (int pass = 0; pass < (P->tp_interleave ? 32 : 1); pass++)
        for (int t = 0; t < ntiles; t++)
            for (int k = P->tp_interleave ? pass : 0; k < (P->tp_interleave ? imin(pass + 1, tnparts[t]) : tnparts[t]); k++) {
                buf_append(&out, tparts[t * 32 + k].p, tparts[t * 32 + k].n);
                if (P->packed_headers == 2) {            /* PPM runs follow the order of the tile-parts in the stream */
                    buf_put32(&ppm, (uint32_t)tphdr[t * 32 + k].n);
                    buf_append(&ppm, tphdr[t * 32 + k].p, tphdr[t * 32 + k].n);
                }
            }
    for (int i = 0; i < ntiles * 32; i++) { free(tparts[i].p); free(tphdr[i].p); }
    free(tparts); free(tphdr); free(tnparts);
    buf_put16(&out, 0xFFD9);
    if (P->packed_headers == 2) {
        /* PPM marker segments go to the end of the main header: Zppm, then the Nppm-prefixed runs cut into chunks */
        Buf seg = { 0 };
        const size_t chunk = P->ppx_chunk > 0 ? (size_t)imin(P->ppx_chunk, 65532) : 65532;
        int z = 0;
        for (size_t at = 0; at < ppm.n; at += chunk, z++) {
            const size_t n = ppm.n - at < chunk ? ppm.n - at : chunk;
            buf_put16(&seg, 0xFF60); buf_put16(&seg, (unsigned)(3 + n)); buf_put(&seg, z & 255);
            buf_append(&seg, ppm.p + at, n);
        }
        buf_reserve(&out, seg.n);
        memmove(out.p + main_end + seg.n, out.p + main_end, out.n - main_end);
        memcpy(out.p + main_end, seg.p, seg.n);
        out.n += seg.n;
        free(seg.p);
    }
    free(ppm.p);
    int64_t n = (int64_t)out.n;
    if (n <= cap)
        memcpy(outbuf, out.p, out.n);
    free(out.p);
    if (P->error)
        return -P->error;
    return n <= cap ? n : -100;
}
