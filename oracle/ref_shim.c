/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin harness around the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libj2kref.so).  It replaces the three
 * entry points of libavcodec's generic layer that the native jpeg2000 decoder needs
 * (ff_thread_get_buffer, ff_set_dimensions, av_codec_is_encoder) and exposes plain-C entry
 * points that tests/ and bench.py's reference arm load with ctypes:
 *
 *   ref_decode()        whole codestream -> pixels, through REF_CODEC.cb.decode
 *                       (libavcodec/jpeg2000dec.c:2825-2908, FFCodec at :2926-2939)
 *   ref_ht_block()      one HT code-block -> int32 sign-magnitude coefficients, through
 *                       ff_jpeg2000_decode_htj2k (libavcodec/jpeg2000htdec.c:1188-1336)
 *   ref_idwt()          ff_jpeg2000_dwt_init + ff_dwt_decode (libavcodec/jpeg2000dwt.c:539-620)
 *   ref_mct()           Jpeg2000DSPContext.mct_decode[] (libavcodec/jpeg2000dsp.c:43-104)
 *   ref_parser_*()      ff_jpeg2000_parser.parser_parse (libavcodec/jpeg2000_parser.c:177-215) on a hand-made
 *                       AVCodecParserContext: splits a byte stream into codestreams / JP2 files
 *   ref_decode_bench()  the CPU baseline: N codestreams decoded by T threads, one decoder
 *                       context per thread (what `ffmpeg -threads T` frame threading does,
 *                       libavcodec/pthread_frame.c:248-295), wall-clock seconds returned.
 *
 * Nothing in the product (ffmpeg-ht_b200/) may link or load this file.
 */
#define _GNU_SOURCE            /* pthread_barrier_t under -std=c17 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "libavutil/imgutils.h"
#include "libavutil/buffer.h"
#include "libavutil/log.h"
#include "libavutil/opt.h"
#include "libavutil/mem.h"
#include "libavutil/pixdesc.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/codec_internal.h"
#include "libavcodec/jpeg2000.h"
#include "libavcodec/jpeg2000dec.h"
#include "libavcodec/jpeg2000dsp.h"
#include "libavcodec/jpeg2000dwt.h"
#include "libavcodec/jpeg2000htdec.h"
#include "libavcodec/thread.h"

/* The same harness is compiled a second time around the product's FFCodec shim
 * (ffmpeg-ht_b200/ffcodec/jpeg2000_b200.c, -DREF_CODEC=ff_jpeg2000_b200_decoder -DREF_PREFIX_B200) so that
 * tests can drive the B200 path through the very FFCodec callbacks libavcodec would call. */
#ifndef REF_CODEC
#define REF_CODEC ff_jpeg2000_decoder
#define REF_NAME(x) ref_##x
#define REF_STAGE_API 1
#else
#define REF_NAME(x) b200shim_##x
#include "b200j2k.h"
#endif
extern const FFCodec REF_CODEC;

/* ---- the three generic-layer functions the decoder calls ------------------------------ */

int av_codec_is_encoder(const AVCodec *codec) { (void)codec; return 0; }

int ff_set_dimensions(AVCodecContext *s, int width, int height)
{
    int ret = av_image_check_size2(width, height, s->max_pixels, AV_PIX_FMT_NONE, 0, s);
    if (ret < 0)
        width = height = 0;
    s->coded_width  = width;
    s->coded_height = height;
    s->width        = AV_CEIL_RSHIFT(width,  s->lowres);     /* as libavcodec/utils.c:94-107 */
    s->height       = AV_CEIL_RSHIFT(height, s->lowres);
    return ret;
}

int ff_thread_get_buffer(AVCodecContext *avctx, AVFrame *f, int flags)
{
    (void)flags;
    f->width  = avctx->width;
    f->height = avctx->height;
    f->format = avctx->pix_fmt;
    /* 64-byte line alignment: same class of padded linesize the real pool hands out */
    int ret = av_image_alloc(f->data, f->linesize, f->width, f->height, avctx->pix_fmt, 64);
    return ret < 0 ? ret : 0;
}

/* the FFCodec shim of the product owns a pinned frame pool while the caller uses the default allocator: the two
 * generic-layer symbols that path touches */
int avcodec_default_get_buffer2(AVCodecContext *s, AVFrame *frame, int flags)
{
    return ff_thread_get_buffer(s, frame, flags);
}

int ff_decode_frame_props(AVCodecContext *avctx, AVFrame *frame)
{
    (void)avctx;
    frame->pts = AV_NOPTS_VALUE;
    return 0;
}

#ifndef REF_STAGE_API
/* decode.c's format negotiation, as much of it as a decoder proposing a hardware format needs: the caller's
 * get_format picks from the list, the default takes the first software format (avcodec_default_get_format) */
#include "libavutil/hwcontext.h"
int ff_get_format(AVCodecContext *avctx, const enum AVPixelFormat *fmt)
{
    if (avctx->get_format)
        return avctx->get_format(avctx, fmt);
    while (*fmt != AV_PIX_FMT_NONE && (av_pix_fmt_desc_get(*fmt)->flags & AV_PIX_FMT_FLAG_HWACCEL))
        fmt++;
    return *fmt;
}
#endif

static void release_frame(AVFrame *frame)
{
    if (frame->buf[0])
        av_buffer_unref(&frame->buf[0]);          /* pool buffer (the product's shim) */
    else if (frame->data[0])
        av_freep(&frame->data[0]);                /* av_image_alloc of the ff_thread_get_buffer stand-in above */
}

static int serial_execute2(AVCodecContext *c,
                           int (*func)(AVCodecContext *c2, void *arg, int jobnr, int threadnr),
                           void *arg2, int *ret, int count)
{
    for (int i = 0; i < count; i++) {
        int r = func(c, arg2, i, 0);
        if (ret)
            ret[i] = r;
    }
    return 0;
}

/* ---- decoder context handling ---------------------------------------------------------- */

static int g_loglevel_set;

static AVCodecContext *ctx_open(int bitexact)
{
    if (!g_loglevel_set) {
        av_log_set_level(getenv("REF_LOG") ? atoi(getenv("REF_LOG")) : AV_LOG_FATAL);   /* REF_LOG=32: the reference's own messages */
        g_loglevel_set = 1;
    }
    AVCodecContext *avctx = av_mallocz(sizeof(*avctx));
    if (!avctx)
        return NULL;
    avctx->codec      = &REF_CODEC.p;
    avctx->codec_id   = AV_CODEC_ID_JPEG2000;
    avctx->codec_type = AVMEDIA_TYPE_VIDEO;
    avctx->pix_fmt    = AV_PIX_FMT_NONE;
    avctx->max_pixels = INT32_MAX;
    avctx->execute2   = serial_execute2;
    avctx->flags      = (bitexact & 1) ? AV_CODEC_FLAG_BITEXACT : 0;   /* argument = flags: bit 0 BITEXACT, bits 8-11 lowres */
    avctx->lowres     = (bitexact >> 8) & 15;
    avctx->priv_data  = av_mallocz(REF_CODEC.priv_data_size);
    if (!avctx->priv_data) {
        av_free(avctx);
        return NULL;
    }
    *(const AVClass **)avctx->priv_data = REF_CODEC.p.priv_class;
    /* av_opt_set_defaults() for the integer / boolean private options (avcodec_open2 does this, avcodec.c:209-216);
     * REF_OPT_<name>=<value> in the environment overrides one, like -<name> <value> on the ffmpeg command line */
    for (const AVOption *o = REF_CODEC.p.priv_class ? REF_CODEC.p.priv_class->option : NULL; o && o->name; o++)
        if (o->type == AV_OPT_TYPE_INT || o->type == AV_OPT_TYPE_BOOL) {
            char key[64];
            snprintf(key, sizeof(key), "REF_OPT_%s", o->name);
            *(int *)((uint8_t *)avctx->priv_data + o->offset) = getenv(key) ? atoi(getenv(key)) : (int)o->default_val.i64;
        }
    if (REF_CODEC.init(avctx) < 0) {
        av_free(avctx->priv_data);
        av_free(avctx);
        return NULL;
    }
    return avctx;
}

static void ctx_close(AVCodecContext *avctx)
{
    if (!avctx)
        return;
    if (REF_CODEC.close)
        REF_CODEC.close(avctx);
    av_free(avctx->priv_data);
    av_free(avctx);
}

/* copy the visible bytes of every plane, tightly packed, into out */
static int64_t pack_frame(const AVCodecContext *avctx, const AVFrame *f, uint8_t *out, int64_t cap)
{
    int64_t pos = 0;
    int linesizes[4];
    if (av_image_fill_linesizes(linesizes, avctx->pix_fmt, f->width) < 0)
        return -1;
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(avctx->pix_fmt);
    for (int p = 0; p < 4 && f->data[p]; p++) {
        int h = f->height;
        if (p == 1 || p == 2)
            h = AV_CEIL_RSHIFT(f->height, d->log2_chroma_h);
        if ((d->flags & AV_PIX_FMT_FLAG_PAL) && p == 1) {
            if (pos + 1024 > cap)
                return -2;
            memcpy(out + pos, f->data[1], 1024);
            pos += 1024;
            continue;
        }
        for (int y = 0; y < h; y++) {
            if (pos + linesizes[p] > cap)
                return -2;
            memcpy(out + pos, f->data[p] + (int64_t)y * f->linesize[p], linesizes[p]);
            pos += linesizes[p];
        }
    }
    return pos;
}

/*
 * info[0]=width info[1]=height info[2]=AVPixelFormat info[3]=bytes written info[4]=decode return
 * info[5]/info[6]=AVCodecContext.sample_aspect_ratio num/den
 * returns 0 on success, the negative AVERROR otherwise.
 */
int REF_NAME(decode)(const uint8_t *cs, int64_t len, int bitexact, int64_t *info, uint8_t *out, int64_t cap)
{
    AVCodecContext *avctx = ctx_open(bitexact);
    if (!avctx)
        return -12;
    uint8_t *padded = av_mallocz(len + AV_INPUT_BUFFER_PADDING_SIZE);
    memcpy(padded, cs, len);
    AVPacket pkt;
    memset(&pkt, 0, sizeof(pkt));
    pkt.data = padded;
    pkt.size = (int)len;
    AVFrame *frame = av_mallocz(sizeof(AVFrame));
    int got = 0;
    int ret = REF_CODEC.cb.decode(avctx, frame, &got, &pkt);
    info[4] = ret;
    if (ret >= 0 && got) {
        info[0] = frame->width;
        info[1] = frame->height;
        info[2] = avctx->pix_fmt;
        info[3] = pack_frame(avctx, frame, out, cap);
        info[5] = avctx->sample_aspect_ratio.num;
        info[6] = avctx->sample_aspect_ratio.den;
        ret = info[3] < 0 ? -1 : 0;
    } else if (ret >= 0) {
        ret = -1;
    }
    release_frame(frame);
    av_free(frame);
    av_free(padded);
    ctx_close(avctx);
    return ret;
}

/*
 * A sequence through ONE decoder context, frames collected in output order -- also from a decoder with delay (empty
 * packets drain it at the end, as libavcodec does for AV_CODEC_CAP_DELAY decoders).  info: 8 words per OUTPUT frame
 * ([0] width [1] height [2] format [3] bytes written), *nout = frames that came out; out: the frames back to back.
 */
int REF_NAME(decode_seq)(const uint8_t *const *cs, const int64_t *len, int n, int bitexact, int64_t *info, int *nout,
                         uint8_t *out, int64_t cap)
{
    AVCodecContext *avctx = ctx_open(bitexact);
    if (!avctx)
        return AVERROR(ENOMEM);
    int ret = 0, k = 0;
    int64_t pos = 0;
    for (int i = 0; ret >= 0 && (i < n || k < n); i++) {
        AVPacket pkt;
        memset(&pkt, 0, sizeof(pkt));
        uint8_t *padded = NULL;
        if (i < n) {
            padded = av_mallocz(len[i] + AV_INPUT_BUFFER_PADDING_SIZE);
            memcpy(padded, cs[i], len[i]);
            pkt.data = padded;
            pkt.size = (int)len[i];
        }
        AVFrame *frame = av_mallocz(sizeof(AVFrame));
        frame->extended_data = frame->data;
        int got = 0;
        avctx->pix_fmt = AV_PIX_FMT_NONE;
        ret = REF_CODEC.cb.decode(avctx, frame, &got, &pkt);
        if (ret >= 0 && got) {
            int64_t *fi = info + 8 * k;
            fi[0] = frame->width;
            fi[1] = frame->height;
            fi[2] = frame->format;
            if (frame->format >= 0)
                avctx->pix_fmt = frame->format;
            const int64_t w = pack_frame(avctx, frame, out + pos, cap - pos);
            if (w < 0)
                ret = AVERROR(EINVAL);
            else {
                fi[3] = w;
                pos += w;
                k++;
            }
        } else if (ret >= 0 && i >= n)
            ret = AVERROR_EOF;                            /* drained dry before every frame came out */
        release_frame(frame);
        av_free(frame);
        av_free(padded);
    }
    *nout = k;
    ctx_close(avctx);
    return ret < 0 && ret != AVERROR_EOF ? ret : 0;
}

const char *REF_NAME(pix_fmt_name)(int fmt) { return av_get_pix_fmt_name(fmt); }

#ifndef REF_STAGE_API
/* ---- the shim with a CUDA device / frames context ------------------------------------------ */
int standin_download(const AVFrame *f, enum AVPixelFormat sw_format, uint8_t *out, int64_t cap);

static enum AVPixelFormat pick_cuda(AVCodecContext *c, const enum AVPixelFormat *fmt)
{
    (void)c;
    for (; *fmt != AV_PIX_FMT_NONE; fmt++)
        if (*fmt == AV_PIX_FMT_CUDA)
            return *fmt;
    return AV_PIX_FMT_NONE;
}
static enum AVPixelFormat pick_software(AVCodecContext *c, const enum AVPixelFormat *fmt)
{
    (void)c;
    while (*fmt != AV_PIX_FMT_NONE && (av_pix_fmt_desc_get(*fmt)->flags & AV_PIX_FMT_FLAG_HWACCEL))
        fmt++;
    return *fmt;
}

/*
 * Decodes `n` codestreams of one sequence with ONE decoder context the way an application with CUDA frames would.
 * mode 0: the caller owns the frames context (avctx->hw_frames_ctx, sized from sw_name / width / height)
 * mode 1: the caller only offers the device (avctx->hw_device_ctx): the decoder makes the frames context
 * mode 2: a device is offered but get_format declines AV_PIX_FMT_CUDA: ordinary software frames
 * info per frame: [0]=width [1]=height [2]=AVFrame.format [3]=bytes written [4]=1 when data[0] is device memory
 *                 [5]=avctx->sw_pix_fmt; out receives the frames one after the other, tightly packed.
 */
int REF_NAME(decode_cuda)(const uint8_t *const *cs, const int64_t *len, int n, int mode, const char *sw_name, int width,
                          int height, int64_t *info, uint8_t *out, int64_t cap)
{
    AVBufferRef *device = NULL, *frames = NULL;
    int ret = av_hwdevice_ctx_create(&device, AV_HWDEVICE_TYPE_CUDA, "0", NULL, 0);
    if (ret < 0)
        return ret;
    AVCodecContext *avctx = ctx_open(0);
    if (!avctx) {
        av_buffer_unref(&device);
        return AVERROR(ENOMEM);
    }
    avctx->get_format = mode == 2 ? pick_software : pick_cuda;
    if (mode == 0) {
        frames = av_hwframe_ctx_alloc(device);
        AVHWFramesContext *fc = (AVHWFramesContext *)frames->data;
        fc->format    = AV_PIX_FMT_CUDA;
        fc->sw_format = av_get_pix_fmt(sw_name);
        fc->width     = width;
        fc->height    = height;
        if ((ret = av_hwframe_ctx_init(frames)) < 0)
            goto end;
        avctx->hw_frames_ctx = av_buffer_ref(frames);
    } else {
        avctx->hw_device_ctx = av_buffer_ref(device);
    }
    int64_t pos = 0;
    for (int i = 0; i < n && ret >= 0; i++) {
        AVPacket pkt;
        memset(&pkt, 0, sizeof(pkt));
        uint8_t *padded = av_mallocz(len[i] + AV_INPUT_BUFFER_PADDING_SIZE);
        memcpy(padded, cs[i], len[i]);
        pkt.data = padded;
        pkt.size = (int)len[i];
        AVFrame *frame = av_mallocz(sizeof(AVFrame));
        frame->extended_data = frame->data;
        int got = 0;
        ret = REF_CODEC.cb.decode(avctx, frame, &got, &pkt);
        if (ret >= 0 && got) {
            int64_t *fi = info + 8 * i;
            fi[0] = frame->width;
            fi[1] = frame->height;
            fi[2] = frame->format;
            fi[4] = b200j2k_pointer_device(frame->data[0]) >= 0;
            fi[5] = avctx->sw_pix_fmt;
            int64_t w = frame->format == AV_PIX_FMT_CUDA ? standin_download(frame, avctx->sw_pix_fmt, out + pos, cap - pos)
                                                         : pack_frame(avctx, frame, out + pos, cap - pos);
            if (w < 0)
                ret = AVERROR(EINVAL);
            else {
                fi[3] = w;
                pos += w;
            }
        } else if (ret >= 0)
            ret = AVERROR_INVALIDDATA;
        release_frame(frame);
        av_free(frame);
        av_free(padded);
    }
end:
    av_buffer_unref(&avctx->hw_frames_ctx);
    av_buffer_unref(&avctx->hw_device_ctx);
    ctx_close(avctx);
    av_buffer_unref(&frames);
    av_buffer_unref(&device);
    return ret < 0 ? ret : 0;
}
#endif

#ifdef REF_STAGE_API
/* ---- stage-level entry points ---------------------------------------------------------- */

/*
 * One HT code-block.  data = Dcup||Dref (length bytes); returns the reference's return code and
 * fills out[h*w] with t1.data (sign-magnitude, libavcodec/jpeg2000htdec.c:1318-1331).
 */
int ref_ht_block(const uint8_t *data, int length, int npasses, int Lcup, int Lref, int zbp,
                 int w, int h, int M_b, int roi_shift, int vsc, int32_t *out)
{
    static _Thread_local Jpeg2000T1Context t1;
    Jpeg2000DecoderContext s;
    Jpeg2000CodingStyle codsty;
    Jpeg2000Cblk cblk;
    AVCodecContext *avctx = ctx_open(0);
    memset(&s, 0, sizeof(s));
    memset(&codsty, 0, sizeof(codsty));
    memset(&cblk, 0, sizeof(cblk));
    s.avctx = avctx;
    int l2w = 2;
    while ((1 << l2w) < w)
        l2w++;
    t1.stride = (1 << l2w) + 2;           /* jpeg2000dec.c:2227 */
    cblk.data = av_mallocz(length + 8);
    memcpy(cblk.data, data, length);
    cblk.length          = length;
    cblk.npasses         = npasses;
    cblk.pass_lengths[0] = Lcup;
    cblk.pass_lengths[1] = Lref;
    cblk.zbp             = zbp;
    cblk.modes           = JPEG2000_CTSY_HTJ2K_F | (vsc ? JPEG2000_CBLK_VSC : 0);
    int ret = ff_jpeg2000_decode_htj2k(&s, &codsty, &t1, &cblk, w, h, M_b, roi_shift);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            out[y * w + x] = t1.data[y * t1.stride + x];
    av_free(cblk.data);
    ctx_close(avctx);
    return ret;
}

/* type: 0 = FF_DWT97 (float), 1 = FF_DWT53, 2 = FF_DWT97_INT; border = {x0,x1,y0,y1} */
int ref_idwt(int type, const int *border, int levels, void *data)
{
    DWTContext d;
    int b[2][2] = { { border[0], border[1] }, { border[2], border[3] } };
    memset(&d, 0, sizeof(d));
    int ret = ff_jpeg2000_dwt_init(&d, b, levels, type);
    if (ret < 0)
        return ret;
    ret = ff_dwt_decode(&d, data);
    ff_dwt_destroy(&d);
    return ret;
}

int ref_mct(int type, void *p0, void *p1, void *p2, int n)
{
    Jpeg2000DSPContext c;
    ff_jpeg2000dsp_init(&c);
    c.mct_decode[type](p0, p1, p2, n);
    return 0;
}

#endif /* REF_STAGE_API */

/* ---- the frame splitter (AVCodecParser "jpeg2000") --------------------------------------- */
#include "libavcodec/parser.h"
#ifndef REF_PARSER
#define REF_PARSER ff_jpeg2000_parser
#endif
extern const AVCodecParser REF_PARSER;
const AVCodecParser *av_parser_iterate(void **opaque) { (void)opaque; return NULL; }   /* parser.c's only loose end */

void *REF_NAME(parser_new)(int complete_frames)
{
    AVCodecParserContext *s = av_mallocz(sizeof(*s));
    s->priv_data = av_mallocz(REF_PARSER.priv_data_size);
    s->parser = &REF_PARSER;
    s->flags = complete_frames ? PARSER_FLAG_COMPLETE_FRAMES : 0;
    return s;
}

/* one parser_parse call: returns its return value; *out / *out_size as the parser set them (the pointer is
 * reported as an offset into `buf` when it points there, else -1 = the parser's own buffer, copied to `copy`) */
int REF_NAME(parser_parse)(void *p, const uint8_t *buf, int size, int64_t *out_off, int *out_size, uint8_t *copy, int64_t cap)
{
    AVCodecParserContext *s = p;
    const uint8_t *o = NULL;
    int osz = 0;
    int r = REF_PARSER.parser_parse(s, NULL, &o, &osz, buf, size);
    *out_size = osz;
    if (o && o >= buf && o <= buf + size)
        *out_off = o - buf;
    else
        *out_off = -1;
    if (o && osz > 0 && osz <= cap)
        memcpy(copy, o, osz);
    return r;
}

void REF_NAME(parser_free)(void *p)
{
    AVCodecParserContext *s = p;
    REF_PARSER.parser_close(s);
    av_free(s->priv_data);
    av_free(s);
}


/* ---- CPU baseline ---------------------------------------------------------------------- */

typedef struct BenchJob {
    const uint8_t *const *cs;
    const int64_t *len;
    int nframes, first, step, bitexact, reps;
    int errors;
    pthread_barrier_t *start;
} BenchJob;

/* one packet in, at most one frame out (a decoder with delay -- the shim with async_depth > 1 -- returns the frame of an
 * earlier packet, or none yet); data == NULL: an empty packet, which drains.  Returns < 0 on error, else got_frame. */
static int bench_decode_one(AVCodecContext *avctx, const uint8_t *data, int64_t len)
{
    AVPacket pkt;
    memset(&pkt, 0, sizeof(pkt));
    pkt.data = (uint8_t *)data;                           /* padded copies made by decode_bench, outside the timing */
    pkt.size = (int)len;
    AVFrame *frame = av_mallocz(sizeof(AVFrame));
    frame->extended_data = frame->data;
    int got = 0;
    avctx->pix_fmt = AV_PIX_FMT_NONE;
    const int ret = REF_CODEC.cb.decode(avctx, frame, &got, &pkt);
    release_frame(frame);
    av_free(frame);
    return ret < 0 ? ret : got;
}

/* sends the packets first, first + step, ... `reps` times and drains; errors + missing frames are counted */
static void bench_decode_run(BenchJob *j, AVCodecContext *avctx, int reps)
{
    int sent = 0, got = 0;
    for (int r = 0; r < reps; r++)
        for (int i = j->first; i < j->nframes; i += j->step) {
            const int g = bench_decode_one(avctx, j->cs[i], j->len[i]);
            sent++;
            if (g < 0)
                j->errors++;
            else
                got += g;
        }
    while (got < sent) {                                  /* frames still inside a decoder with delay */
        const int g = bench_decode_one(avctx, NULL, 0);
        if (g <= 0)
            break;
        got += g;
    }
    if (got != sent)
        j->errors++;
}

static void *bench_worker(void *arg)
{
    BenchJob *j = arg;
    AVCodecContext *avctx = ctx_open(j->bitexact);
    /* steady state is what is measured: opening the context and the first frame of a sequence (tables, buffers
     * that grow once) happen before the clock starts, for the reference and for the shim alike */
    if (j->first < j->nframes) {
        BenchJob warm = *j;
        warm.nframes = j->first + 1;
        bench_decode_run(&warm, avctx, 1);                /* buffers, tables */
        bench_decode_run(&warm, avctx, 5);                /* frame pools at their steady-state depth */
        j->errors += warm.errors;
    }
    pthread_barrier_wait(j->start);
    bench_decode_run(j, avctx, j->reps);
    ctx_close(avctx);
    return NULL;
}

/* decodes every codestream `reps` times using `threads` decoder contexts; returns seconds (<0 on error).
 * REF_PINNED_PACKETS=1 (shim build only): the packets live in page-locked memory, as a caller's own AVBuffer pool
 * could provide them; default: ordinary av_malloc'ed packets, what libavformat hands over. */
double REF_NAME(decode_bench)(const uint8_t *const *cs, const int64_t *len, int nframes, int threads,
                        int reps, int bitexact)
{
    if (threads < 1)
        threads = 1;
    if (threads > 256)
        threads = 256;
    pthread_t th[256];
    BenchJob jobs[256];
    struct timespec t0, t1;
    pthread_barrier_t start;
    /* make sure one-time init (LUTs, log level) is outside the timed region */
    ctx_close(ctx_open(bitexact));
    /* the packets as a demuxer hands them over: padded (AV_INPUT_BUFFER_PADDING_SIZE), allocated once */
    uint8_t **padded = av_calloc(nframes, sizeof(*padded));
    int pinned = 0;
    for (int i = 0; i < nframes; i++) {
        padded[i] = av_mallocz(len[i] + AV_INPUT_BUFFER_PADDING_SIZE);
        memcpy(padded[i], cs[i], len[i]);
    }
#ifndef REF_STAGE_API
    if (getenv("REF_PINNED_PACKETS") && atoi(getenv("REF_PINNED_PACKETS"))) {
        pinned = 1;
        for (int i = 0; i < nframes; i++)
            if (b200j2k_host_register(padded[i], len[i] + AV_INPUT_BUFFER_PADDING_SIZE) < 0)
                pinned = -1;
    }
#endif
    pthread_barrier_init(&start, NULL, threads + 1);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (BenchJob){ (const uint8_t *const *)padded, len, nframes, t, threads, bitexact, reps, 0, &start };
        pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
    }
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int errors = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        errors += jobs[t].errors;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    pthread_barrier_destroy(&start);
    for (int i = 0; i < nframes; i++) {
#ifndef REF_STAGE_API
        if (pinned)
            b200j2k_host_unregister(padded[i]);
#endif
        av_free(padded[i]);
    }
    av_free(padded);
    if (pinned < 0)
        return -2.0;
    if (errors)
        return -1.0;
    return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
