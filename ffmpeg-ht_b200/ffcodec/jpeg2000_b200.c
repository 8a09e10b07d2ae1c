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
 * Added: .close (device buffers must be released), FF_CODEC_CAP_INIT_CLEANUP, and a pinned frame pool: as long
 * as the caller has not installed its own get_buffer2, the pictures come from an AVBufferPool of page-locked host
 * memory owned by this decoder (the way libdav1d.c owns its picture pool), so the device->host copy lands in the
 * AVFrame itself instead of going through the library's staging buffer and a 50 MB memcpy per 4K frame
 * (private option "pinned_frames", default on).  With a caller-supplied get_buffer2 the picture is taken from
 * ff_thread_get_buffer() as in the reference.
 * Added: device-resident pictures.  When the caller offers a CUDA device or frames
 * context (AVCodecContext.hw_device_ctx / hw_frames_ctx, libavutil/hwcontext_cuda.c:81-260) the decoder proposes
 * AV_PIX_FMT_CUDA through ff_get_format() the way cuviddec.c does; if that is what get_format picks, the picture is
 * taken from the frames context with av_hwframe_get_buffer() -- AVFrame.data[] are device pointers, sw_format is the
 * format the reference would have produced -- and the kernels' output is copied there device to device: it never
 * crosses PCIe.  Without such a context nothing changes.
 * What it refuses (AVERROR_PATCHWELCOME, as the library reports): Part-1 code-blocks with arithmetic-coding bypass or
 * termination on each pass -- a build that wants those keeps the native decoder registered as fallback under a
 * different name.
 */
#include <stdatomic.h>
#include <string.h>
#include <time.h>

#include "libavutil/buffer.h"
#include "libavutil/hwcontext.h"
#include "libavutil/imgutils.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "avcodec.h"
#include "codec_internal.h"
#include "decode.h"
#include "profiles.h"
#include "thread.h"

#include "b200j2k.h"

#define B200J2K_MAX_ASYNC 4

typedef struct B200J2KContext {
    const AVClass *class;
    b200j2k_decoder *dec;
    int device;             /* CUDA device index; -1 = round-robin over frame-thread workers */
    int reduction_factor;   /* the "lowres" private option; AVCodecContext.lowres feeds it (jpeg2000dec.c:2811-2818) */
    int flags;              /* B200J2K_FLAG_* this context decodes with */
    int pinned_frames;      /* own the frame pool (page-locked memory) while the caller uses the default get_buffer2 */
    AVBufferPool *pool;
    size_t pool_size;
    AVBufferRef *hwframes;  /* frames context this decoder created from the caller's hw_device_ctx */
    int cur_device;         /* device the library context lives on */
    /* "async_depth" > 1: up to that many pictures are in flight inside this context, the frame of packet N comes out
     * while packet N + depth - 1 goes in (a decoder with delay, like every hardware decoder of libavcodec) */
    int async_depth;
    struct {
        AVFrame *frame;
        int slot;
    } pending[B200J2K_MAX_ASYNC];
    int npending;
    double t_phase[6];      /* B200SHIM_TRACE=1: seconds spent in probe / get frame / submit / upload wait / output wait */
    int n_calls;
} B200J2KContext;

static int b200j2k_drain(AVCodecContext *avctx, B200J2KContext *s);
static double b200j2k_now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* ---- pinned frame pool ------------------------------------------------------------------------------------ */
static void b200j2k_pinned_free(void *opaque, uint8_t *data)
{
    b200j2k_host_free(data);
}

static AVBufferRef *b200j2k_pinned_alloc(size_t size)
{
    uint8_t *p = b200j2k_host_alloc((int64_t)size);
    AVBufferRef *ref;
    if (!p)
        return NULL;
    ref = av_buffer_create(p, size, b200j2k_pinned_free, NULL, 0);
    if (!ref)
        b200j2k_host_free(p);
    return ref;
}

/* what ff_get_buffer() + avcodec_default_get_buffer2() do for a video frame (decode.c, get_buffer.c:86-125), with the
 * one buffer of the picture taken from page-locked memory: 64-byte aligned rows, planes back to back */
static int b200j2k_get_pinned_frame(AVCodecContext *avctx, B200J2KContext *s, AVFrame *f)
{
    int linesize[4], ret;
    ptrdiff_t ls[4];
    size_t sizes[4], total = 0;

    f->width  = avctx->width;
    f->height = avctx->height;
    f->format = avctx->pix_fmt;
    if ((ret = ff_decode_frame_props(avctx, f)) < 0)
        return ret;
    if ((ret = av_image_fill_linesizes(linesize, f->format, f->width)) < 0)
        return ret;
    for (int i = 0; i < 4; i++)
        ls[i] = linesize[i] = FFALIGN(linesize[i], 64);
    if ((ret = av_image_fill_plane_sizes(sizes, f->format, f->height, ls)) < 0)
        return ret;
    for (int i = 0; i < 4; i++) {
        if (sizes[i] > SIZE_MAX / 8 - total)
            return AVERROR(EINVAL);
        total += sizes[i];
    }
    total += 64;
    if (!s->pool || s->pool_size != total) {
        av_buffer_pool_uninit(&s->pool);          /* buffers still referenced by frames stay alive until released */
        s->pool = av_buffer_pool_init(total, b200j2k_pinned_alloc);
        if (!s->pool)
            return AVERROR(ENOMEM);
        s->pool_size = total;
    }
    f->buf[0] = av_buffer_pool_get(s->pool);
    if (!f->buf[0])
        return AVERROR(ENOMEM);
    if ((ret = av_image_fill_pointers(f->data, f->format, f->height, f->buf[0]->data, linesize)) < 0) {
        av_buffer_unref(&f->buf[0]);
        return ret;
    }
    for (int i = 0; i < 4; i++)
        f->linesize[i] = linesize[i];
    f->extended_data = f->data;
    return 0;
}

/* Device-resident picture.  Returns 1 with `f` filled from a CUDA frames context, 0 when the caller's get_format did
 * not pick AV_PIX_FMT_CUDA (the software path goes on), < 0 on error.  Mirrors cuviddec.c: the frames context is the
 * caller's (hw_frames_ctx), else one made here on the caller's device (hw_device_ctx) for this picture size. */
static int b200j2k_get_cuda_frame(AVCodecContext *avctx, B200J2KContext *s, AVFrame *f, enum AVPixelFormat sw_fmt)
{
    const enum AVPixelFormat fmts[] = { AV_PIX_FMT_CUDA, sw_fmt, AV_PIX_FMT_NONE };
    AVBufferRef *ref = avctx->hw_frames_ctx;
    int ret;

    avctx->pix_fmt = sw_fmt;
    ret = ff_get_format(avctx, fmts);
    if (ret < 0)
        return ret;
    if (ret != AV_PIX_FMT_CUDA) {
        avctx->pix_fmt = sw_fmt;
        return 0;
    }
    if (!ref) {
        AVHWFramesContext *fc = s->hwframes ? (AVHWFramesContext *)s->hwframes->data : NULL;
        if (!fc || fc->sw_format != sw_fmt || fc->width != avctx->width || fc->height != avctx->height) {
            av_buffer_unref(&s->hwframes);
            s->hwframes = av_hwframe_ctx_alloc(avctx->hw_device_ctx);
            if (!s->hwframes)
                return AVERROR(ENOMEM);
            fc = (AVHWFramesContext *)s->hwframes->data;
            fc->format    = AV_PIX_FMT_CUDA;
            fc->sw_format = sw_fmt;
            fc->width     = avctx->width;
            fc->height    = avctx->height;
            if ((ret = av_hwframe_ctx_init(s->hwframes)) < 0) {
                av_buffer_unref(&s->hwframes);
                return ret;
            }
        }
        ref = s->hwframes;
    } else {
        const AVHWFramesContext *fc = (const AVHWFramesContext *)ref->data;
        if (fc->format != AV_PIX_FMT_CUDA || fc->sw_format != sw_fmt || fc->width < avctx->width ||
            fc->height < avctx->height) {
            av_log(avctx, AV_LOG_ERROR, "hw_frames_ctx does not fit the picture (%s %dx%d wanted)\n",
                   av_get_pix_fmt_name(sw_fmt), avctx->width, avctx->height);
            return AVERROR(EINVAL);
        }
    }
    avctx->pix_fmt    = AV_PIX_FMT_CUDA;
    avctx->sw_pix_fmt = sw_fmt;
    if ((ret = ff_decode_frame_props(avctx, f)) < 0)
        return ret;
    if ((ret = av_hwframe_get_buffer(ref, f, 0)) < 0)
        return ret;
    f->width  = avctx->width;
    f->height = avctx->height;
    return 1;
}

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
        static atomic_uint next_device;      /* frame threads open one context each: spread them */
        device = (int)(atomic_fetch_add(&next_device, 1) % (unsigned)ndev);
    }
    s->cur_device = device;
    s->async_depth = FFMAX(1, FFMIN(s->async_depth, B200J2K_MAX_ASYNC));
    return b200j2k_create(&s->dec, device, s->async_depth, flags);
}

static av_cold int b200j2k_ff_close(AVCodecContext *avctx)
{
    B200J2KContext *s = avctx->priv_data;
    if (getenv("B200SHIM_TRACE") && s->n_calls)
        av_log(avctx, AV_LOG_ERROR, "b200 shim trace: %d calls, ms per call: probe %.3f frame %.3f submit %.3f upload-wait %.3f "
               "output-wait %.3f\n", s->n_calls, 1e3 * s->t_phase[0] / s->n_calls, 1e3 * s->t_phase[1] / s->n_calls,
               1e3 * s->t_phase[2] / s->n_calls, 1e3 * s->t_phase[3] / s->n_calls, 1e3 * s->t_phase[4] / s->n_calls);
    b200j2k_drain(avctx, s);
    b200j2k_destroy(s->dec);
    s->dec = NULL;
    av_buffer_pool_uninit(&s->pool);
    av_buffer_unref(&s->hwframes);
    return 0;
}

/* waits for the oldest picture in flight and hands it out */
static int b200j2k_output_oldest(AVCodecContext *avctx, B200J2KContext *s, AVFrame *picture, int *got_frame)
{
    int ret, failed;
    if (!s->npending)
        return 0;
    ret = b200j2k_wait(s->dec, s->pending[0].slot);
    failed = ret < 0 ? 0 : b200j2k_slot_blocks_failed(s->dec, s->pending[0].slot);
    if (ret >= 0) {
        av_frame_move_ref(picture, s->pending[0].frame);
        *got_frame = 1;
        if (failed > 0)
            av_log(avctx, AV_LOG_ERROR, "%d code-blocks could not be decoded (zeros in their place)\n", failed);
    }
    av_frame_free(&s->pending[0].frame);
    s->npending--;
    memmove(&s->pending[0], &s->pending[1], s->npending * sizeof(s->pending[0]));
    return ret;
}

/* drops everything in flight (close, seek, change of device) */
static int b200j2k_drain(AVCodecContext *avctx, B200J2KContext *s)
{
    while (s->npending) {
        b200j2k_wait(s->dec, s->pending[0].slot);
        av_frame_free(&s->pending[0].frame);
        s->npending--;
        memmove(&s->pending[0], &s->pending[1], s->npending * sizeof(s->pending[0]));
    }
    return 0;
}

static av_cold void b200j2k_ff_flush(AVCodecContext *avctx)
{
    b200j2k_drain(avctx, avctx->priv_data);
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

    double t0, t1;
    if (!avpkt->size)                                      /* end of stream: hand out what is still in flight */
        return b200j2k_output_oldest(avctx, s, picture, got_frame);
    t0 = b200j2k_now();
    /* header pass: everything jpeg2000_read_main_headers() publishes before the picture exists */
    if ((ret = b200j2k_probe(avpkt->data, avpkt->size, flags, &info)) < 0)
        return ret;
    t1 = b200j2k_now(); s->t_phase[0] += t1 - t0; t0 = t1; s->n_calls++;
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

    if (avctx->hw_frames_ctx || avctx->hw_device_ctx) {
        if ((ret = b200j2k_get_cuda_frame(avctx, s, picture, fmt)) < 0)
            return ret;
        if (ret > 0) {
            /* decode on the GPU the caller's frames live on */
            const int dev = b200j2k_pointer_device(picture->data[0]);
            if (dev >= 0 && dev != s->cur_device) {
                if ((ret = b200j2k_drain(avctx, s)) < 0)
                    return ret;
                b200j2k_destroy(s->dec);
                s->dec = NULL;
                if ((ret = b200j2k_create(&s->dec, dev, s->async_depth, s->flags)) < 0)
                    return ret;
                s->cur_device = dev;
            }
            goto have_picture;
        }
    }
    if (s->pinned_frames && (!avctx->get_buffer2 || avctx->get_buffer2 == avcodec_default_get_buffer2))
        ret = b200j2k_get_pinned_frame(avctx, s, picture);
    else
        ret = ff_thread_get_buffer(avctx, picture, 0);
    if (ret < 0)
        return ret;
have_picture:
    t1 = b200j2k_now(); s->t_phase[1] += t1 - t0; t0 = t1;
    for (int i = 0; i < info.nplanes; i++) {
        data[i]     = picture->data[i];
        linesize[i] = picture->linesize[i];
    }
    /* pinned frame pools (get_buffer2 backed by b200j2k_host_alloc) receive the DMA directly; ordinary
     * av_malloc'ed frames go through the library's pinned staging buffer */
    if (s->async_depth <= 1) {
        consumed = b200j2k_decode(s->dec, avpkt->data, avpkt->size, data, linesize, &info);
        if (consumed < 0)
            return (int)consumed;
        t1 = b200j2k_now(); s->t_phase[2] += t1 - t0;
        /* the reference logs every damaged HT block and carries on with zeros in its place (jpeg2000dec.c:2275-2278) */
        if (info.nblocks_failed > 0)
            av_log(avctx, AV_LOG_ERROR, "%d code-blocks could not be decoded (zeros in their place)\n", info.nblocks_failed);
        *got_frame = 1;
        return (int)consumed;
    }
    /* pictures in flight: queue this one (host parsing happens here, copies and kernels are asynchronous), then hand out
     * the oldest one once `async_depth` are under way.  The packet is only read during the call unless it lies in
     * page-locked memory, in which case the reference held by the queued frame's side keeps it alive: libavcodec
     * keeps avpkt valid only for the call, so a pinned packet is staged like a pageable one here. */
    {
        AVFrame *queued = av_frame_alloc();
        int slot;
        if (!queued)
            return AVERROR(ENOMEM);
        av_frame_move_ref(queued, picture);
        slot = b200j2k_submit(s->dec, avpkt->data, avpkt->size, data, linesize, &info);
        if (slot < 0) {
            av_frame_free(&queued);
            return slot;
        }
        t1 = b200j2k_now(); s->t_phase[2] += t1 - t0; t0 = t1;
        /* the packet must not be referenced after this call returns: if the library DMAs straight from a page-locked
         * packet, wait for the upload before giving the packet back */
        if (b200j2k_pointer_is_pinned(avpkt->data) && (ret = b200j2k_wait_upload(s->dec, slot)) < 0) {
            av_frame_free(&queued);
            return ret;
        }
        t1 = b200j2k_now(); s->t_phase[3] += t1 - t0; t0 = t1;
        s->pending[s->npending].frame = queued;
        s->pending[s->npending].slot  = slot;
        s->npending++;
        consumed = info.bytes_consumed;
        if (s->npending >= s->async_depth && (ret = b200j2k_output_oldest(avctx, s, picture, got_frame)) < 0)
            return ret;
        t1 = b200j2k_now(); s->t_phase[4] += t1 - t0;
        return (int)consumed;
    }
}

#define OFFSET(x) offsetof(B200J2KContext, x)
#define VD AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_DECODING_PARAM
static const AVOption b200j2k_options[] = {
    { "lowres", "Lower the decoding resolution by a power of two",
        OFFSET(reduction_factor), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 32, VD },
    { "device", "CUDA device index (-1: spread decoder contexts over all devices)",
        OFFSET(device), AV_OPT_TYPE_INT, { .i64 = -1 }, -1, 64, VD },
    { "pinned_frames", "Take the pictures from a pool of page-locked memory owned by the decoder unless the caller "
                       "installed a get_buffer2",
        OFFSET(pinned_frames), AV_OPT_TYPE_BOOL, { .i64 = 1 }, 0, 1, VD },
    { "async_depth", "Pictures in flight inside one decoder context (1 = synchronous, like the native decoder; more "
                     "adds that many minus one frames of delay and overlaps host parsing with the transfers)",
        OFFSET(async_depth), AV_OPT_TYPE_INT, { .i64 = 1 }, 1, B200J2K_MAX_ASYNC, VD },
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
    /* the native decoder's capabilities plus DELAY: only with async_depth > 1 does a frame ever stay behind, but the
     * flag is what makes libavcodec drain the decoder at the end of the stream */
    .p.capabilities = AV_CODEC_CAP_SLICE_THREADS | AV_CODEC_CAP_FRAME_THREADS | AV_CODEC_CAP_DR1 | AV_CODEC_CAP_DELAY,
    .priv_data_size = sizeof(B200J2KContext),
    .init           = b200j2k_ff_init,
    .close          = b200j2k_ff_close,
    .flush          = b200j2k_ff_flush,
    FF_CODEC_DECODE_CB(b200j2k_ff_decode),
    .p.priv_class   = &b200j2k_class,
    .p.max_lowres   = 5,
    .p.profiles     = NULL_IF_CONFIG_SMALL(ff_jpeg2000_profiles),
    .caps_internal  = FF_CODEC_CAP_SKIP_FRAME_FILL_PARAM | FF_CODEC_CAP_INIT_CLEANUP,
};
