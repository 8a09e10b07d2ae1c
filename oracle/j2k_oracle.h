/*
 * oracle/j2k_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C, of the reference's HTJ2K decode path (sandflow/ffmpeg-ht,
 * libavcodec/jpeg2000{dec,htdec,dwt,dsp}.c and jpeg2000.c).  It is the checker the CUDA path is
 * compared with; it is never linked, loaded or called by the product (ffmpeg-ht_b200/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every entry point below against the
 * unmodified reference compiled by oracle/Makefile (oracle/_ref/libj2kref.so) on seeded inputs, and
 * tests/golden/ holds reference-generated vectors for boxes where /root/reference is absent.
 */
#ifndef J2K_ORACLE_H
#define J2K_ORACLE_H

#include <stdint.h>

#define ORACLE_ERR_INVALID   (-1094995529)   /* AVERROR_INVALIDDATA */
#define ORACLE_ERR_PATCH     (-1163346256)   /* AVERROR_PATCHWELCOME */
#define ORACLE_ERR_NOMEM     (-12)
#define ORACLE_ERR_EINVAL    (-22)

/* transform ids follow the reference's enum DWTType (libavcodec/jpeg2000dwt.h:37-41) */
enum { ORACLE_DWT97 = 0, ORACLE_DWT53 = 1, ORACLE_DWT97_INT = 2 };

/* pixel formats the path can produce (values are this repo's own, not AVPixelFormat) */
enum {
    ORACLE_PIX_GRAY8 = 1, ORACLE_PIX_GRAY16, ORACLE_PIX_YA8, ORACLE_PIX_YA16,
    ORACLE_PIX_RGB24, ORACLE_PIX_RGB48, ORACLE_PIX_RGBA, ORACLE_PIX_RGBA64,
    ORACLE_PIX_PLANAR8, ORACLE_PIX_PLANAR16   /* planar YUV(A), any subsampling: see oracle_last_pix_fmt_name() */
};

/* H1-H8: one HT code-block.  data = Dcup||Dref.  out[h*w] = sign-magnitude int32 as left in
 * t1->data by ff_jpeg2000_decode_htj2k (libavcodec/jpeg2000htdec.c:1188-1336).
 * Returns 1 decoded, 0 nothing to decode, <0 error (out is all zero then, quirk Q1). */
int oracle_ht_block(const uint8_t *data, int length, int npasses, int Lcup, int Lref, int zbp,
                    int w, int h, int M_b, int roi_shift, int vsc, int32_t *out);

/* W1-W4: inverse DWT of one Mallat-layout plane, border = {x0,x1,y0,y1} (jpeg2000dwt.c:539-620) */
/* Part-1 (EBCOT / MQ) code-block of style 0 -> the same sign-magnitude words (decode_cblk, jpeg2000dec.c:1993-2089).
 * nonzerobits = expn + guard bits - 1 - (zero bit-planes - ROI shift of component 0); bandpos 0 LL, 1 HL, 2 LH, 3 HH;
 * out: w * h words, row stride w.  0 = no data, 1 = decoded, ORACLE_ERR_INVALID = the bit-plane counter left its
 * range ("bpno became invalid"): out holds what was decoded until then, which the reference dequantises as it is. */
int oracle_mq_block(const uint8_t *data, int length, int npasses, int nonzerobits, int w, int h, int bandpos,
                    int M_b, int roi_shift, int32_t *out);

/* the same under a code-block style with one codeword segment: 0x02 context reset after each pass, 0x08 vertically
 * causal contexts, 0x20 segmentation symbols (jpeg2000dec.c:1999, :2034-2039) */
int oracle_mq_block_style(const uint8_t *data, int length, int npasses, int nonzerobits, int w, int h, int bandpos,
                          int M_b, int roi_shift, int style, int32_t *out);

int oracle_idwt(int type, const int *border, int levels, void *data);

/* M1-M3: inverse component transform in place (jpeg2000dsp.c:43-91) */
int oracle_mct(int type, void *p0, void *p1, void *p2, int n);

/* whole codestream -> tightly packed pixels.
 * info[0]=width info[1]=height info[2]=ORACLE_PIX_* info[3]=bytes written
 * info[4]=number of code-blocks decoded info[5]=ncomponents info[6]=precision
 * info[7]=sample aspect ratio from a JP2 'res ' box, num << 32 | den, 0 = none (jpeg2000dec.c:2762-2796, :2867).
 * Returns 0 or a negative error. */
int oracle_decode(const uint8_t *cs, int64_t len, int bitexact, int64_t *info, uint8_t *out, int64_t cap);
/* libavutil name of the AVPixelFormat the last oracle_decode() of this thread chose ("rgb48le", "yuv420p10le") */
const char *oracle_last_pix_fmt_name(void);

/* CPU baseline ("port"): decode every codestream `reps` times on `threads` threads; seconds or <0 */
double oracle_decode_bench(const uint8_t *const *cs, const int64_t *len, int nframes, int threads,
                           int reps, int bitexact);

#endif
