/*
 * jpeg2000_b200.c -- FFCodec shim: the reference-side half of the drop-in boundary.
 *
 * This file is meant to be added to libavcodec/ of sandflow/ffmpeg-ht.  It registers
 * a decoder with the SAME identity as the native one (name "jpeg2000", AV_CODEC_ID_JPEG2000, same
 * capabilities, same "lowres" private option; cf. ff_jpeg2000_decoder, libavcodec/jpeg2000dec.c:2913-2939)
 * whose decode callback hands the packet to the B200 library through the C ABI of include/b200j2k.h
 * instead of running tile_codeblocks / mct_decode / write_frame on the CPU
 * (libavcodec/jpeg2000dec.c:2212-2395).
 *
 * Contract kept:
 *   init    int (*)(AVCodecContext *)                                         (jpeg2000dec.c:2807)
 *   decode  int (*)(AVCodecContext *, AVFrame *, int *got_frame, AVPacket *)  (jpeg2000dec.c:2825)
 *   sets avctx->width/height (ff_set_dimensions), pix_fmt, profile, bits_per_raw_sample, LOSSLESS property,
 *   takes the picture from ff_thread_get_buffer(), returns the bytes consumed or a negative AVERROR,
 *   never AVERROR(EAGAIN); one decoder context per AVCodecContext, so frame threading (one context per
 *   worker, libavcodec/pthread_frame.c:248-295) maps to one CUDA stream + staging buffers per worker.
 * Added: .close (device buffers must be released) and FF_CODEC_CAP_INIT_CLEANUP.
 * What it refuses (AVERROR_PATCHWELCOME, as the library reports): Part-1 code-blocks with mode switches (bypass,
 * restart, ...), MIXED HT sets, coding-style / progression changes in tile-part headers, PAL8 -- a build that wants
 * those keeps the native decoder registered as fallback
 * under a different name.
 */
#include "libavutil/imgutils.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "avcodec.h"
#include "codec_internal.h"
#include "decode.h"
#include "thread.h"

#include "b200j2k.h"

typedef struct B200J2KContext {
    const AVClass *class;
    b200j2k_decoder *dec;
    int device;             /* CUDA device index; -1 = round-robin over frame-thread workers */
    int reduction_factor;   /* the "lowres" private option; AVCodecContext.lowres feeds it (jpeg2000dec.c:2811-2818) */
    int flags;              /* B200J2K_FLAG_* this context decodes with */
} B200J2KContext;

static av_cold int b200j2k_ff_init(AVCodecContext *avctx)
{
    B200J2KContext *s = avctx->priv_data;
    int ndev = b200j2k_device_count();
    int flags = (avctx->flags & AV_CODEC_FLAG_BITEXACT) ? B200J2K_FLAG_BITEXACT : 0;
    int device = s->device;

    /* jpeg2000_decode_init (jpeg2000dec.c:2807-2823) */
    if (avctx->lowres)
        av_log(avctx, AV_LOG_WARNING, "lowres is overriden by reduction_factor but set anyway\n");
    if (!s->reduction_factor && avctx->lowres < 33)
        s->reduction_factor = avctx->lowres;
    if (avctx->lowres != s->reduction_factor && avctx->lowres)
        return AVERROR(EINVAL);
    if (s->reduction_factor > 15)
        return AVERROR(EINVAL);
    flags |= B200J2K_FLAG_LOWRES(s->reduction_factor);
    s->flags = flags;
    if (ndev <= 0) {
        av_log(avctx, AV_LOG_ERROR, "no CUDA device: the B200 HTJ2K path has no CPU fallback\n");
        return AVERROR(ENODEV);
    }
    if (device < 0) {
        static int next_device;              /* frame threads open one context each: spread them */
        device = next_device++ % ndev;
    }
    return b200j2k_create(&s->dec, device, 1, flags);
}

static av_cold int b200j2k_ff_close(AVCodecContext *avctx)
{
    B200J2KContext *s = avctx->priv_data;
    b200j2k_destroy(s->dec);
    s->dec = NULL;
    return 0;
}

static int b200j2k_ff_decode(AVCodecContext *avctx, AVFrame *picture, int *got_frame, AVPacket *avpkt)
{
    B200J2KContext *s = avctx->priv_data;
    b200j2k_frame_info info;
    uint8_t *data[4] = { NULL };
    int linesize[4] = { 0 };
    enum AVPixelFormat fmt;
    int64_t consumed;
    int ret, flags = s->flags;

    /* header pass: everything jpeg2000_read_main_headers() publishes before the picture exists */
    if ((ret = b200j2k_probe(avpkt->data, avpkt->size, flags, &info)) < 0)
        return ret;
    /* the generic layer shifts the dimensions back down by lowres (jpeg2000dec.c:326, utils.c:94-107) */
    if ((ret = ff_set_dimensions(avctx, info.width << avctx->lowres, info.height << avctx->lowres)) < 0)
        return ret;
    /* the library names the format get_siz() would pick (jpeg2000dec.c:328-412) */
    fmt = av_get_pix_fmt(info.av_pix_fmt_name);
    if (fmt == AV_PIX_FMT_NONE)
        return AVERROR_PATCHWELCOME;
    avctx->pix_fmt             = fmt;
    avctx->profile             = info.profile;
    avctx->bits_per_raw_sample = info.precision;
    if (info.lossless)
        avctx->properties |= FF_CODEC_PROPERTY_LOSSLESS;
    if (info.sar_num && info.sar_den)                      /* jpeg2000dec.c:2867-2868 */
        avctx->sample_aspect_ratio = (AVRational){ info.sar_num, info.sar_den };

    if (avctx->skip_frame >= AVDISCARD_ALL)
        return avpkt->size;

    if ((ret = ff_thread_get_buffer(avctx, picture, 0)) < 0)
        return ret;
    for (int i = 0; i < info.nplanes; i++) {
        data[i]     = picture->data[i];
        linesize[i] = picture->linesize[i];
    }
    /* pinned frame pools (get_buffer2 backed by b200j2k_host_alloc) receive the DMA directly; ordinary
     * av_malloc'ed frames go through the library's pinned staging buffer */
    consumed = b200j2k_decode(s->dec, avpkt->data, avpkt->size, data, linesize, &info);
    if (consumed < 0)
        return (int)consumed;
    *got_frame = 1;
    return (int)consumed;
}

#define OFFSET(x) offsetof(B200J2KContext, x)
#define VD AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_DECODING_PARAM
static const AVOption b200j2k_options[] = {
    { "lowres", "Lower the decoding resolution by a power of two",
        OFFSET(reduction_factor), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 32, VD },
    { "device", "CUDA device index (-1: spread decoder contexts over all devices)",
        OFFSET(device), AV_OPT_TYPE_INT, { .i64 = -1 }, -1, 64, VD },
    { NULL },
};

static const AVClass b200j2k_class = {
    .class_name = "jpeg2000",
    .item_name  = av_default_item_name,
    .option     = b200j2k_options,
    .version    = LIBAVUTIL_VERSION_INT,
};

const FFCodec ff_jpeg2000_b200_decoder = {
    .p.name         = "jpeg2000",
    CODEC_LONG_NAME("JPEG 2000 (HTJ2K decode path on NVIDIA B200)"),
    .p.type         = AVMEDIA_TYPE_VIDEO,
    .p.id           = AV_CODEC_ID_JPEG2000,
    .p.capabilities = AV_CODEC_CAP_SLICE_THREADS | AV_CODEC_CAP_FRAME_THREADS | AV_CODEC_CAP_DR1,
    .priv_data_size = sizeof(B200J2KContext),
    .init           = b200j2k_ff_init,
    .close          = b200j2k_ff_close,
    FF_CODEC_DECODE_CB(b200j2k_ff_decode),
    .p.priv_class   = &b200j2k_class,
    .p.max_lowres   = 5,
    .caps_internal  = FF_CODEC_CAP_SKIP_FRAME_FILL_PARAM | FF_CODEC_CAP_INIT_CLEANUP,
};
